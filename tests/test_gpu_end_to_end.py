"""BASELINE configs[3] in miniature, as a user of the library runs it on one GPU: cold start over a list of PRNs
(one launch), the reference's code-search rounds and pre-track for the satellites found, then closed-loop tracking
with the loop filters on the device, streamed from host memory.  Checked against the GROUND TRUTH of the synthetic
scene (which satellites, Doppler, code phase, data bits) - parity of every stage with the reference is the business
of the other test modules; this one shows the stages fit together."""
import ctypes as C

import numpy as np
import pytest

from stm32f4_sdr_gps_b200 import Channels, Receiver, load_host_library
from stm32f4_sdr_gps_b200.signal_synth import Satellite, Scene, synthesize

pytestmark = pytest.mark.gpu

MS_SAMPLES = 16368


def test_cold_start_to_data_bits(host_engine):
    rng = np.random.default_rng(404)
    searched = list(range(1, 13))
    present = {3: (-3210.0, 4000.5), 7: (1475.0, 12001.2), 11: (4620.0, 777.7)}      # prn: (doppler Hz, code phase samples)
    sats = [Satellite(prn=p, doppler_hz=d, code_phase_samples=c, cn0_dbhz=48.0, carrier_phase_rad=float(rng.uniform(0, 6.28)),
                      nav_bit_offset_ms=int(rng.integers(0, 20))) for p, (d, c) in present.items()]
    n_ms = 3000
    scene = Scene(sats=sats, n_ms=n_ms, seed=4040)
    sig = synthesize(scene)
    lib = load_host_library()
    ring = host_engine.ring_ms

    # 1. cold start: every (PRN, Doppler bin, ms) cell in one launch, the reference's chain vote per bin
    host_engine.upload_signal(0, sig[:ring])
    ch_all = Channels(searched)
    rx_all = Receiver(host_engine, ch_all)
    votes, _ = rx_all.cold_sweep(-5000, 500, 21, 0, 10)
    # The reference's decision (one bin with a chain of >= 3 like phases, acquisition.c:365-416) also fires on noise now
    # and then and relies on the later rounds to time such a channel out.  With all ten cells of every bin in hand a
    # caller can ask for more: a satellite that is there collects chains in its bin AND the neighbouring one (its
    # Doppler lies between two bins), noise does not; the Doppler hint is the vote-weighted mean of the two.
    found = {}
    for i, p in enumerate(searched):
        v = votes[i].astype(int)
        pair = v[:-1] + v[1:]
        b = int(np.argmax(pair))
        if ch_all.snapshot(i).acq_state == 2 and pair[b] >= 6:                          # GPS_ACQ_FREQ_SEARCH_DONE
            found[p] = (-5000 + 500 * b) + 500.0 * v[b + 1] / pair[b]
    rx_all.close()
    ch_all.free()
    assert set(found) == set(present), found
    for p, f in found.items():
        assert abs(f - present[p][0]) <= 300, (p, f)

    # 2. the reference's own sequencing from there: code-search rounds 1..3 for all found satellites side by side
    #    (gps_master_handling + one launch per snapshot), then pre-track, then tracking
    prns = sorted(found)
    ch = Channels(prns, [int(found[p]) for p in prns])
    rx = Receiver(host_engine, ch)
    lib.gpsb_host_set_sat_cnt(len(prns))
    lib.gpsb_host_master_reset()
    lib.gps_master_handling.argtypes = [C.c_void_p, C.c_uint8]
    ms = 10
    while ms < 900:
        lib.gpsb_host_set_packet_cnt(ms)
        lib.gps_master_handling(ch.at(0), ms % 4)
        if not lib.gps_master_need_acq():
            break
        rx.acquire_ms(ms)
        ms += 1
    assert not lib.gps_master_need_acq(), [ch.snapshot(i).acq_state for i in range(len(prns))]
    for i, p in enumerate(prns):                                                     # half-chip code phase of the epoch
        truth = (present[p][1] / 8.0) % 2046
        got = ch.snapshot(i).found_code_phase
        assert min(abs(got - truth), 2046 - abs(got - truth)) <= 3, (p, got, truth)
    lib.gps_master_handling(ch.at(0), ms % 4)                                        # IDLE -> NEED_PRE_TRACK for everybody
    assert all(ch.snapshot(i).trk_state == 1 for i in range(len(prns)))

    # 3. tracking: pre-track on the per-millisecond path, then the device-resident loop, the recording streamed
    #    from host memory through the 1024-ms ring
    t0 = ms
    iq_a, _ = rx.track_run(t0, 200)                                                  # frames still resident from step 1
    assert all(ch.snapshot(i).trk_state == 4 for i in range(len(prns)))              # GPS_TRACKING_RUN
    t1 = t0 + 200
    iq_b, nav_b = rx.track_stream(t1, np.ascontiguousarray(sig[t1:n_ms]), chunk_ms=64)
    dev_ms, host_ms = rx.loop_stats()
    assert dev_ms >= len(prns) * (n_ms - t1)

    # 4. against the truth: Doppler, code phase, and the data bits in the sign of the prompt in-phase sum
    for i, p in enumerate(prns):
        st = ch.snapshot(i)
        f = float(np.uint32(st.if_freq_offset_hz_bits).view(np.float32))
        assert abs(f - present[p][0]) < 25.0, (p, f)
        dop, cp0 = present[p]
        want_fine = (cp0 - (n_ms - 1) * MS_SAMPLES * dop / 1_575_420_000.0) % MS_SAMPLES
        fine = float(np.uint32(st.code_phase_fine_bits).view(np.float32))
        err = abs(fine - want_fine)
        assert min(err, MS_SAMPLES - err) < 6.0, (p, fine, want_fine)
        truth = scene.truth[p]
        ip = iq_b[-1000:, i, 2].astype(np.int64)
        m = np.arange(n_ms - 1000, n_ms)
        # the code epoch inside ms m starts about cp0 samples in; the data bit of ms m is the one of the epoch that
        # covers most of it; skip the ms next to a bit edge
        epoch = m - (1 if cp0 > MS_SAMPLES / 2 else 0)
        bit_idx = (epoch - truth["nav_bit_offset_ms"]) // 20 + 1
        pos = (epoch - truth["nav_bit_offset_ms"]) % 20
        keep = (pos >= 2) & (pos <= 17)
        d = truth["nav_bits"][np.clip(bit_idx, 0, truth["nav_bits"].size - 1)].astype(np.int64) * 2 - 1
        agree = np.sum(np.sign(ip[keep]) * d[keep]) / keep.sum()
        assert abs(agree) > 0.97, (p, agree)                     # +-1: the Costas loop may lock upside down
        assert np.abs(ip).mean() > 300                           # and the prompt arm carries the power
    rx.close()
    ch.free()
    lib.gpsb_host_set_sat_cnt(4)
    lib.gpsb_host_master_reset()
