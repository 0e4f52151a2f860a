"""Generate tests/golden/golden_l2.npz from the UNMODIFIED reference C (oracle/_ref/libgpsref.so).

Run in the authoring container only (needs /root/reference to build oracle/_ref):

    python tests/golden/make_golden.py

The GPU box has no /root/reference; the committed .npz is what pins the oracle and the CUDA path
there.  Every array below is an output of reference code (Firmware/project_main/GPS/gps_misc.c,
tracking.c, acquisition.c; Firmware/project_single_sat/GPS/simulator.c) on the inputs stored next
to it.  Noise > 0 fixtures depend on glibc rand(), so the noisy buffers themselves are stored.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

from oracle_lib import Reference, f32_bits  # noqa: E402
from stm32f4_sdr_gps_b200.signal_synth import Satellite, Scene, synthesize  # noqa: E402

IF_HZ = 4092000


def main() -> None:
    ref = Reference()
    g = {}
    chans = ref.channels(40)

    # ---- C/A codes, PRN 1..210 (gps_misc.c:317-372), bit-packed
    codes = np.zeros((210, 1023), np.uint8)
    ch0 = ref.channel_at(chans, 0)
    for prn in range(1, 211):
        ref.channel_init(ch0, prn, 0)
        codes[prn - 1] = ref.prn_code(ch0)
    g["ca_codes_packed"] = np.packbits(codes, axis=1)

    # ---- replicas (gps_misc.c:282-300) for PRN 1 and 19, bits 0..15
    rep = np.zeros((2, 16, 1023), np.uint16)
    for i, prn in enumerate((1, 19)):
        ref.channel_init(ch0, prn, 0)
        for b in range(16):
            buf = np.zeros(1024, np.uint16)
            ref.lib.gps_generate_prn_data2(ch0, buf.ctypes.data, b)
            rep[i, b] = buf[:1023]
    g["replica_prns"] = np.array([1, 19])
    g["replica_words"] = rep

    # ---- simulator KAT (SS/main.c:59-68): noise 0 / 15 / 30 / 45 after srand(1)
    ref.channel_init(ch0, 1, 0)
    noises = np.array([0, 15, 30, 45])
    sim = np.stack([ref.sim_buffer(int(n), 1) for n in noises])
    g["sim_noise"] = noises
    g["sim_buffers"] = sim
    g["sim_search"] = np.array([ref.search_cell(ch0, sim[i], 2000, 0, 0, 2046) for i in range(4)], np.int32)
    g["sim_iq_all"] = np.stack([ref.iq_cell(ch0, sim[i], float(IF_HZ + 2000), 0, 0, 2046) for i in range(4)])
    # sub-byte shifts on the noisy buffer: all 2046 offsets x bits 0..15
    g["sim15_iq_bits"] = np.stack([ref.iq_cell(ch0, sim[1], float(IF_HZ + 2000), b, 0, 2046) for b in range(16)])

    # ---- NCO words (gps_misc.c:219): fp32 divide + truncation
    freqs = np.array([IF_HZ + d for d in range(-7000, 7001, 250)] + [IF_HZ + 0.5, IF_HZ - 1234.567, IF_HZ + 4999.99],
                     np.float32)
    nco = np.zeros(freqs.size, np.uint32)
    for i, f in enumerate(freqs):
        di = np.zeros(2048, np.uint8)
        dq = np.zeros(2048, np.uint8)
        # recover acc_step*32 from the mixer's behaviour is awkward; use the tracking variant's accumulator
        _, acc = ref.epl_cell(ch0, sim[0], float(np.float32(f) - np.float32(IF_HZ)), 0, 0.0)
        nco[i] = acc  # = 511 * step32 mod 2^32
    g["nco_freq_offsets"] = (freqs - np.float32(IF_HZ)).astype(np.float32)
    g["nco_acc_after_511"] = nco

    # ---- mixer output (gps_misc.c:211-240) on a random buffer for a Doppler sweep
    rng = np.random.default_rng(20231101)
    rnd = rng.integers(0, 256, 2046, dtype=np.uint8)
    g["rnd_signal"] = rnd
    mix_f = np.array([IF_HZ - 7000, IF_HZ - 500, IF_HZ, IF_HZ + 2000, IF_HZ + 6500], np.float32)
    mix_i = np.zeros((mix_f.size, 2044), np.uint8)
    mix_q = np.zeros((mix_f.size, 2044), np.uint8)
    for i, f in enumerate(mix_f):
        di = np.full(2048, 0xAA, np.uint8)
        dq = np.full(2048, 0x55, np.uint8)
        ref.lib.gps_shift_to_zero_freq(rnd.ctypes.data, di.ctypes.data, dq.ctypes.data, np.float32(f))
        assert di[2044] == 0xAA and dq[2045] == 0x55  # bytes 2044..2045 never written
        mix_i[i] = di[:2044]
        mix_q[i] = dq[:2044]
    g["mix_freqs"] = mix_f
    g["mix_i"] = mix_i
    g["mix_q"] = mix_q

    # ---- raw correlator on fully random buffers incl. non-zero tail bytes (gps_misc.c:48-145)
    prn_r = rng.integers(0, 65536, 1024, dtype=np.uint16)
    d_i = rng.integers(0, 65536, 1024, dtype=np.uint16)
    d_q = rng.integers(0, 65536, 1024, dtype=np.uint16)
    iq = np.zeros((2046, 2), np.int16)
    c8 = np.zeros(2046, np.int16)
    import ctypes as C
    for off in range(2046):
        a, b = C.c_int16(), C.c_int16()
        ref.lib.gps_correlation_iq(prn_r.ctypes.data, d_i.ctypes.data, d_q.ctypes.data, off, C.byref(a), C.byref(b))
        iq[off] = (a.value, b.value)
        c8[off] = ref.lib.gps_correlation8(prn_r.ctypes.data, d_i.ctypes.data, d_q.ctypes.data, off)
    g["raw_prn"] = prn_r[:1023]
    g["raw_i"] = d_i[:1023]
    g["raw_q"] = d_q[:1023]
    g["raw_iq"] = iq
    g["raw_corr8"] = c8
    wins = np.array([[0, 2046], [0, 1], [2045, 2046], [100, 107], [750, 1250], [1990, 2046], [5, 5], [9, 3]])
    sr = np.zeros((len(wins), 3), np.int32)
    for k, (a0, a1) in enumerate(wins):
        avr, ph = C.c_uint16(), C.c_uint16()
        mx = ref.lib.correlation_search(prn_r.ctypes.data, d_i.ctypes.data, d_q.ctypes.data, int(a0), int(a1),
                                        C.byref(avr), C.byref(ph))
        sr[k] = (mx, ph.value, avr.value)
    g["raw_windows"] = wins
    g["raw_search"] = sr

    # ---- a small multi-satellite scene: sweep cells and a closed-loop tracking trace
    scene = Scene(sats=[Satellite(prn=5, doppler_hz=1020.0, code_phase_samples=15920.4, cn0_dbhz=50.0,
                                  carrier_phase_rad=0.3, nav_bit_offset_ms=7),
                        Satellite(prn=14, doppler_hz=-2480.0, code_phase_samples=811.7, cn0_dbhz=50.0,
                                  carrier_phase_rad=2.1, nav_bit_offset_ms=13)],
                  n_ms=600, seed=0x5D120009)
    sig = synthesize(scene)
    g["scene_signal"] = sig
    g["scene_prns"] = np.array([5, 14, 20])          # PRN 20 is absent
    for i, prn in enumerate((5, 14, 20)):
        ref.channel_init(ref.channel_at(chans, i), prn, 0)
    g["scene_sweep"] = ref.sweep_cells(chans, 3, sig, 4, -5000, 500, 21, 0)      # (3, 21, 4, 3)
    g["scene_sweep_bits3"] = ref.sweep_cells(chans, 3, sig, 2, -3000, 500, 3, 3)  # sub-byte replica shift

    trk_iq, trk_nav, trk_state, trk_final = [], [], [], []
    for i, s in enumerate(scene.sats):
        ch = ref.channel_at(chans, 10 + i)
        ref.channel_init(ch, s.prn, 0)
        st = ref.snapshot(ch)
        st.acq_state = 9                               # GPS_ACQ_DONE
        st.found_freq_offset_hz = int(round(s.doppler_hz / 500.0) * 500)
        st.found_code_phase = int(round(s.code_phase_samples / 8.0)) % 2046
        st.trk_state = 1                               # GPS_NEED_PRE_TRACK
        ref.restore(ch, st)
        iq_log, nav_log, st_log = ref.track_run(ch, sig, 0, 600)
        trk_iq.append(iq_log)
        trk_nav.append(nav_log)
        trk_state.append(st_log.view(np.uint32))
        fin = ref.snapshot(ch)
        trk_final.append(np.frombuffer(bytes(fin), np.uint8).copy())
    g["track_found_freq"] = np.array([int(round(s.doppler_hz / 500.0) * 500) for s in scene.sats])
    g["track_found_phase"] = np.array([int(round(s.code_phase_samples / 8.0)) % 2046 for s in scene.sats])
    g["track_iq"] = np.stack(trk_iq)
    g["track_nav"] = np.stack(trk_nav)
    g["track_state_bits"] = np.stack(trk_state)
    g["track_final_flat"] = np.stack(trk_final)

    out = HERE / "golden_l2.npz"
    np.savez_compressed(out, **g)
    print("wrote", out, out.stat().st_size, "bytes;", len(g), "arrays")


if __name__ == "__main__":
    main()
