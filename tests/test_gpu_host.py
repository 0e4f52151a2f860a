"""The host-side mirror (libgpsb_host.so) running on the GPU, against the reference's closed-loop
traces (committed golden fixtures produced by the unmodified reference) - bit-exact."""
import ctypes as C

import numpy as np
import pytest

from stm32f4_sdr_gps_b200 import Channels, Receiver, load_host_library

pytestmark = pytest.mark.gpu
IF_HZ = 4092000


def _armed_channels(golden, prns=(5, 14)):
    ch = Channels(list(prns))
    for i in range(len(prns)):
        st = ch.snapshot(i)
        st.acq_state, st.trk_state = 9, 1                  # GPS_ACQ_DONE, GPS_NEED_PRE_TRACK
        st.found_freq_offset_hz = int(golden["track_found_freq"][i])
        st.found_code_phase = int(golden["track_found_phase"][i])
        ch.restore(i, st)
    return ch


@pytest.mark.parametrize("mode", ["device", "host-threads", "host-lockstep"])
def test_batched_receiver_closed_loop_equals_reference(host_engine, golden, mode):
    """600 ms closed loop (pre-track, E/P/L loops, nav-bit sync) for two satellites in one receiver: I/Q sums,
    nav bits and the final channel records equal the reference run per satellite, whichever way the loop is
    closed.  device: pre-track on the per-millisecond path, then ONE k_track_run launch per channel for the
    rest (loop filters on the GPU).  host-threads: filters on the host, one worker per channel driving its own
    resident-kernel slot.  host-lockstep: filters on the host, single thread, all channels per millisecond."""
    sig = golden["scene_signal"]
    host_engine.upload_signal(0, sig)
    ch = _armed_channels(golden)
    rx = Receiver(host_engine, ch)
    rx.set_loop_site(2 if mode == "device" else 1)
    rx.set_threads(1 if mode == "host-lockstep" else 0)
    launches0 = host_engine.launch_count
    iq, nav = rx.track_run(0, 600)
    assert host_engine.launch_count - launches0 <= 600 * 2          # never more than one launch per kind per ms
    on_device, on_host = rx.loop_stats()
    if mode == "device":
        assert on_device > 2 * 400 and on_host < 2 * 200            # everything after pre-track ran inside k_track_run
        assert host_engine.launch_count - launches0 <= on_host + 4
    else:
        assert on_device == 0
    for s in range(2):
        assert np.array_equal(iq[:, s, :], golden["track_iq"][s]), s
        assert np.array_equal(nav[:, s], golden["track_nav"][s]), s
        assert bytes(ch.snapshot(s)) == golden["track_final_flat"][s].tobytes(), s
    rx.close()
    ch.free()


def test_reference_named_tracking_entry_point(host_engine, golden):
    """gps_tracking_process(channel, data, index) - the drop-in signature (tracking.h:6) - one satellite."""
    lib = load_host_library()
    sig = golden["scene_signal"]
    ch = _armed_channels(golden, prns=(5,))
    lib.gpsb_host_attach(host_engine.handle)
    lib.gps_channell_prepare(ch.at(0))
    assert lib.gpsb_host_last_status() == 0
    for ms in range(200):
        lib.gpsb_host_set_packet_cnt(ms)
        lib.gps_tracking_process(ch.at(0), sig[ms].ctypes.data, ms % 4)
        assert lib.gpsb_host_last_status() == 0
        want = golden["track_state_bits"][0][ms]
        got = ch.snapshot(0)
        assert (got.code_phase_fine_bits, got.if_freq_offset_hz_bits) == (int(want[0]), int(want[1])), ms
    ch.free()


def test_level0_reference_signatures(host_engine, golden):
    """gps_misc.h:198-216 primitives, exact reference signatures, through the GPU."""
    lib = load_host_library()
    lib.gpsb_host_attach(host_engine.handle)
    pad = lambda w: np.concatenate([np.asarray(w, np.uint16), np.zeros(1, np.uint16)])
    prn, di, dq = pad(golden["raw_prn"]), pad(golden["raw_i"]), pad(golden["raw_q"])
    for off in (0, 1, 2, 1023, 2044, 2045):
        a, b = C.c_int16(), C.c_int16()
        lib.gps_correlation_iq(prn.ctypes.data, di.ctypes.data, dq.ctypes.data, off, C.byref(a), C.byref(b))
        assert (a.value, b.value) == tuple(golden["raw_iq"][off])
        assert lib.gps_correlation8(prn.ctypes.data, di.ctypes.data, dq.ctypes.data, off) == golden["raw_corr8"][off]
    avr, ph = C.c_uint16(), C.c_uint16()
    mx = lib.correlation_search(prn.ctypes.data, di.ctypes.data, dq.ctypes.data, 750, 1250, C.byref(avr), C.byref(ph))
    assert (mx, ph.value, avr.value) == tuple(golden["raw_search"][4])
    sig = golden["rnd_signal"]
    di2 = np.full(2048, 0xAA, np.uint8)
    dq2 = np.full(2048, 0x55, np.uint8)
    lib.gps_shift_to_zero_freq(sig.ctypes.data, di2.ctypes.data, dq2.ctypes.data, C.c_float(float(golden["mix_freqs"][3])))
    assert np.array_equal(di2[:2044], golden["mix_i"][3]) and np.array_equal(dq2[:2044], golden["mix_q"][3])
    assert di2[2044] == 0xAA and dq2[2045] == 0x55
    ch = Channels([19])
    for b in (0, 3, 15):
        buf = np.zeros(1024, np.uint16)
        lib.gps_generate_prn_data2(ch.at(0), buf.ctypes.data, b)
        assert np.array_equal(buf[:1023], golden["replica_words"][1, b]), b
    ch.free()


def test_cold_sweep_votes_match_reference_vote(host_engine, golden, reference):
    """One-launch cold start: cells equal the reference cells (golden), and the per-bin chain vote equals
    the reference's acquisition_process_single_freq_data() on those cells (acquisition.c:322-360)."""
    sig = golden["scene_signal"]
    host_engine.upload_signal(0, sig[:16])
    prns = [int(p) for p in golden["scene_prns"]]
    ch = Channels(prns)
    rx = Receiver(host_engine, ch)
    votes, phases = rx.cold_sweep(-5000, 500, 21, 0, 4)
    want_cells = golden["scene_sweep"]                               # (3, 21, 4, 3) from the reference
    rlib = reference.lib
    ref_phases = (C.c_uint16 * 25).in_dll(rlib, "acq_single_freq_phases")
    ref_hist = (C.c_uint32 * 29).in_dll(rlib, "acq_freq_histogram")
    rlib.acquisition_process_single_freq_data.argtypes = [C.c_void_p, C.c_uint8]
    rchans = reference.channels(1)
    rch = reference.channel_at(rchans, 0)
    reference.channel_init(rch, 1, 0)
    for s in range(3):
        for b in range(21):
            for m in range(4):
                ref_phases[m] = int(want_cells[s, b, m, 1])
            for k in range(29):
                ref_hist[k] = 0
            st = reference.snapshot(rch)
            st.freq_index = b
            reference.restore(rch, st)
            rlib.acquisition_process_single_freq_data(rch, 4)
            expect = ref_hist[b]                                      # chain length if >= 2 else 0
            got = int(votes[s, b]) if votes[s, b] >= 2 else 0
            assert got == expect, (s, b, got, expect)
    # the two present satellites collect their votes in the true Doppler bin at the true code phase
    assert votes[0, 12] >= 2 and abs(int(phases[0, 12]) - 1990) <= 2
    assert ch.snapshot(2).acq_state in (1, 2)                         # absent PRN: still searching or (falsely) voted
    rx.close()
    ch.free()


def test_batched_acquisition_equals_reference_per_channel(host_engine, golden, reference):
    """gpsb_rx_acquire_ms: code-phase search rounds 1 and 2 for two channels with a Doppler hint, one
    launch per snapshot for both, against the reference's acquisition_process_channel() per channel."""
    sig = golden["scene_signal"]
    host_engine.upload_signal(0, sig)
    prns, hints = [5, 14], [1000, -2500]
    ch = Channels(prns, hints)
    rx = Receiver(host_engine, ch)
    lib = load_host_library()
    rchans = reference.channels(2)
    reference.lib.acquisition_start_code_search_channel.argtypes = [C.c_void_p]
    reference.set_ms(0)
    lib.gpsb_host_set_packet_cnt(0)
    for i in range(2):
        rch = reference.channel_at(rchans, i)
        reference.channel_init(rch, prns[i], hints[i])
        reference.lib.acquisition_start_channel(rch)                  # hint -> GPS_ACQ_FREQ_SEARCH_DONE
        reference.lib.acquisition_start_code_search_channel(rch)
        lib.acquisition_start_channel(ch.at(i))
        lib.acquisition_start_code_search_channel(ch.at(i))
        assert bytes(ch.snapshot(i)) == bytes(reference.snapshot(rch))
    done = False
    for ms in range(400):
        rx.acquire_ms(ms)
        reference.set_ms(ms)
        for i in range(2):
            reference.lib.acquisition_process_channel(reference.channel_at(rchans, i), sig[ms].ctypes.data)
        for i in range(2):
            assert bytes(ch.snapshot(i)) == bytes(reference.snapshot(reference.channel_at(rchans, i))), (ms, i)
        if all(ch.snapshot(i).acq_state == 6 for i in range(2)):      # GPS_ACQ_CODE_PHASE_SEARCH2_DONE
            done = True
            break
    assert done
    assert abs(ch.snapshot(0).found_code_phase - 1990) <= 16
    rx.close()
    ch.free()
