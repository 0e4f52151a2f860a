"""Host-side state machines (libgpsb_host.so) against the UNMODIFIED reference, on CPU.

The library has no CPU correlator, so the test drives its split-phase API: PLAN says which cell the
channel needs, the test computes that cell with the oracle (standing in for the GPU - the checker,
not the product), FINISH consumes it.  After every millisecond the complete channel state - including
the IEEE bit patterns of all loop-filter floats - must equal the reference's gps_ch_t after the same
millisecond (Firmware/project_main/GPS/tracking.c, acquisition.c, nav_data.c, gps_master.c)."""
import ctypes as C

import numpy as np
import pytest

from stm32f4_sdr_gps_b200 import Channels, Plan, SearchRes, load_host_library
from stm32f4_sdr_gps_b200.host_api import WANT_EPL, WANT_NOTHING, WANT_SEARCH

IF_HZ = 4092000


def oracle_search(oracle, chips, sig, rq):
    """The cell a gpsb_search_req describes: replica(bits) + mix(acc0, step32) + search [start, stop)."""
    if rq.start >= rq.stop:
        return SearchRes(0, 0, 0, 0)
    rep = oracle.replica(chips, rq.off_bits)
    di, dq, _ = oracle.mix(sig, rq.acc0, rq.step32)
    mx, ph, avg = oracle.correlation_search(rep, di, dq, rq.start, rq.stop)
    return SearchRes(mx, ph, avg, 0)


def oracle_epl(oracle, chips, sig, rq):
    return oracle.epl_explicit(chips, sig, rq.acc0, rq.step32, rq.off_e, rq.off_p, rq.off_l, rq.off_bits)


def states_equal(a, b):
    return bytes(a) == bytes(b)


def diff_fields(a, b):
    out = []
    for name, _ in a._fields_:
        va, vb = getattr(a, name), getattr(b, name)
        va = list(va) if hasattr(va, "__len__") else va
        vb = list(vb) if hasattr(vb, "__len__") else vb
        if va != vb:
            out.append((name, va, vb))
    return out


def host_track_ms(lib, oracle, ch, chips, sig_ms, ms, index):
    lib.gpsb_host_set_packet_cnt(ms)
    plan = Plan()
    lib.gpsb_host_plan_track(ch, ms, index, C.byref(plan))
    if plan.want == WANT_SEARCH:
        res = oracle_search(oracle, chips, sig_ms, plan.search)
        lib.gpsb_host_finish_track(ch, index, C.byref(plan), C.byref(res), None)
    elif plan.want == WANT_EPL:
        iq = oracle_epl(oracle, chips, sig_ms, plan.epl)
        lib.gpsb_host_finish_track(ch, index, C.byref(plan), None, iq.ctypes.data)
        return iq
    return None


def test_channel_record_layout_matches_reference(reference):
    lib = load_host_library()
    assert lib.gpsb_host_sizeof_channel() == reference.lib.ref_sizeof_channel()
    # a state written through the reference's struct reads back identically through ours
    chans = reference.channels(1)
    ch = reference.channel_at(chans, 0)
    reference.channel_init(ch, 23, -1500)
    st = reference.snapshot(ch)
    st.trk_state, st.code_phase_fine_bits, st.word_cnt_test, st.subframe_cnt = 4, 0x45123456, 77, 9
    st.pre_track_phases[29], st.subframe_data[37], st.pll_check_buf[3] = 1234, 0xA5, -77
    reference.restore(ch, st)
    from stm32f4_sdr_gps_b200 import FlatState
    mine = FlatState()
    lib.gpsb_host_snapshot(ch, C.byref(mine))          # our accessor on the reference's memory
    assert states_equal(mine, reference.snapshot(ch))
    code = np.zeros(1023, np.uint8)
    lib.gps_generate_prn(code.ctypes.data, 23)
    assert np.array_equal(code, reference.prn_code(ch))


def test_code_generator_all_prns(oracle):
    lib = load_host_library()
    for prn in range(1, 211):
        code = np.zeros(1023, np.uint8)
        lib.gps_generate_prn(code.ctypes.data, prn)
        assert np.array_equal(code, oracle.ca_code(prn)), prn


@pytest.mark.parametrize("sat", [0, 1])
def test_tracking_closed_loop_bit_exact(oracle, reference, golden, sat):
    """Pre-track -> tracking -> bit sync over 600 ms: state identical to the reference after EVERY ms."""
    lib = load_host_library()
    sig = golden["scene_signal"]
    prn = (5, 14)[sat]
    chips = oracle.ca_code(prn)
    rchans = reference.channels(1)
    rch = reference.channel_at(rchans, 0)
    reference.channel_init(rch, prn, 0)
    st = reference.snapshot(rch)
    st.acq_state, st.trk_state = 9, 1
    st.found_freq_offset_hz = int(golden["track_found_freq"][sat])
    st.found_code_phase = int(golden["track_found_phase"][sat])
    reference.restore(rch, st)

    mine = Channels([prn])
    mch = mine.at(0)
    mst = mine.snapshot(0)
    mst.acq_state, mst.trk_state = 9, 1
    mst.found_freq_offset_hz, mst.found_code_phase = st.found_freq_offset_hz, st.found_code_phase
    mine.restore(0, mst)
    lib.gpsb_host_attach(None)            # resets the shared scratch; no GPU is touched in this test

    n_epl = 0
    for ms in range(600):
        index = ms % 4
        reference.set_ms(ms)
        reference.lib.gps_tracking_process(rch, sig[ms].ctypes.data, index)
        iq = host_track_ms(lib, oracle, mch, chips, sig[ms], ms, index)
        a, b = mine.snapshot(0), reference.snapshot(rch)
        assert states_equal(a, b), (ms, diff_fields(a, b))
        if iq is not None:
            n_epl += 1
            assert np.array_equal(iq, golden["track_iq"][sat][ms]), ms     # reference's logged sums
    assert n_epl > 400
    fin = mine.snapshot(0)
    assert bytes(fin) == golden["track_final_flat"][sat].tobytes()        # committed fixture
    assert fin.trk_state == 4                                             # GPS_TRACKING_RUN
    mine.free()


def test_tracking_reference_tdm_schedule(oracle, reference, golden):
    """The reference's own 17-ms time-division schedule (main.c:134-155): two channels share the shared
    scratch, gaps between a channel's slots are bridged by gps_rewind_if_phase, index 0xFF is a dummy."""
    lib = load_host_library()
    sig = golden["scene_signal"]
    prns = (5, 14)
    rchans = reference.channels(4)
    mine = Channels(list(prns) + [0, 0])
    lib.gpsb_host_attach(None)
    for i, prn in enumerate(prns):
        rch = reference.channel_at(rchans, i)
        reference.channel_init(rch, prn, 0)
        st = reference.snapshot(rch)
        st.acq_state, st.trk_state = 9, 1
        st.found_freq_offset_hz = int(golden["track_found_freq"][i])
        st.found_code_phase = int(golden["track_found_phase"][i])
        reference.restore(rch, st)
        mst = mine.snapshot(i)
        mst.acq_state, mst.trk_state = 9, 1
        mst.found_freq_offset_hz, mst.found_code_phase = st.found_freq_offset_hz, st.found_code_phase
        mine.restore(i, mst)
    for ms in range(600):
        big = ms % 17
        sat = big // 4
        if sat >= 4:
            sat = 0
        index = 0xFF if big == 16 else big % 4
        if sat >= 2:
            continue
        reference.set_ms(ms)
        reference.lib.gps_tracking_process(reference.channel_at(rchans, sat), sig[ms].ctypes.data, index)
        host_track_ms(lib, oracle, mine.at(sat), oracle.ca_code(prns[sat]), sig[ms], ms, index)
        a, b = mine.snapshot(sat), reference.snapshot(reference.channel_at(rchans, sat))
        assert states_equal(a, b), (ms, sat, diff_fields(a, b))
    assert mine.snapshot(0).trk_state == 4
    mine.free()


def test_false_lock_kicker_and_rand(oracle, reference):
    """Noise-only input drives the bad-lock counters into the random re-seed (tracking.c:300-326);
    both sides draw from the same libc rand() sequence."""
    lib = load_host_library()
    rng = np.random.default_rng(77)
    sig = rng.integers(0, 256, (1500, 2046), dtype=np.uint8)
    prn = 9
    chips = oracle.ca_code(prn)
    libc = C.CDLL(None)
    results = []
    for side in ("ref", "mine"):
        libc.srand(12345)
        if side == "ref":
            chans = reference.channels(1)
            ch = reference.channel_at(chans, 0)
            reference.channel_init(ch, prn, 0)
            st = reference.snapshot(ch)
        else:
            mine = Channels([prn])
            ch = mine.at(0)
            st = mine.snapshot(0)
            lib.gpsb_host_attach(None)
        st.acq_state, st.trk_state, st.found_freq_offset_hz = 9, 4, 1500
        st.if_freq_offset_hz_bits = int(np.float32(1500.0).view(np.uint32))
        st.code_phase_fine_bits = int(np.float32(4000.0).view(np.uint32))
        st.pll_bad_state_cnt, st.pll_bad_state_master_cnt = 10, 80     # one more bad slot trips the kicker
        if side == "ref":
            reference.restore(ch, st)
        else:
            mine.restore(0, st)
        trace = []
        for ms in range(1500):
            if side == "ref":
                reference.set_ms(ms)
                reference.lib.gps_tracking_process(ch, sig[ms].ctypes.data, ms % 4)
                trace.append(bytes(reference.snapshot(ch)))
            else:
                host_track_ms(lib, oracle, ch, chips, sig[ms], ms, ms % 4)
                trace.append(bytes(mine.snapshot(0)))
        results.append(trace)
    first_bad = next((i for i, (a, b) in enumerate(zip(*results)) if a != b), None)
    assert first_bad is None, first_bad
    # the kicker actually fired: the master counter was cleared from above the threshold region
    from stm32f4_sdr_gps_b200 import FlatState
    master = [FlatState.from_buffer_copy(t).pll_bad_state_master_cnt for t in results[0]]
    assert max(master) >= 80 and any(a >= 80 and b == 0 for a, b in zip(master, master[1:])), max(master)


def _run_acquisition(lib, oracle, reference, sig, prns, given, n_snap):
    """gps_master-sequenced acquisition (main.c:163-168) on both sides; returns per-snapshot equality."""
    n = len(prns)
    assert n == 4                                  # the reference build has GPS_SAT_CNT == 4
    rchans = reference.channels(n)
    mine = Channels(prns, given)
    for i in range(n):
        reference.channel_init(reference.channel_at(rchans, i), prns[i], given[i])
    return rchans, mine


def test_acquisition_with_master_sequencing(oracle, reference, golden):
    """Full acquisition of four channels under gps_master_handling(): Doppler search (chain votes over
    10 snapshots per bin) for the channels without a hint, then code search 1/2/3 for all."""
    lib = load_host_library()
    sig = golden["scene_signal"]
    prns = [5, 14, 5, 14]
    given = [0, -2500, 1000, 0]
    # the reference's sequencing state lives in file-scope flags (gps_master.c:44-46): reload for a clean start
    from oracle_lib import REF_SO
    import shutil, tempfile, os
    tmp = tempfile.mkdtemp()
    fresh = os.path.join(tmp, "libgpsref_fresh.so")
    shutil.copy(REF_SO, fresh)
    rlib = C.CDLL(fresh)
    rlib.gps_fill_summ_table()
    rlib.ref_channels_alloc.restype = C.c_void_p
    rlib.ref_channel_at.restype = C.c_void_p
    rlib.ref_channel_at.argtypes = [C.c_void_p, C.c_uint32]
    rlib.ref_channel_init.argtypes = [C.c_void_p, C.c_uint32, C.c_int32]
    rlib.acquisition_process.argtypes = [C.c_void_p, C.c_void_p]
    rlib.gps_master_handling.argtypes = [C.c_void_p, C.c_uint8]
    from oracle_lib import FlatState as RefFlat
    rlib.ref_channel_snapshot.argtypes = [C.c_void_p, C.POINTER(RefFlat)]
    rchans = rlib.ref_channels_alloc(4)
    mine = Channels(prns, given)
    for i in range(4):
        rlib.ref_channel_init(rlib.ref_channel_at(rchans, i), prns[i], given[i])
    lib.gpsb_host_attach(None)
    lib.gpsb_host_master_reset()
    lib.gpsb_host_set_sat_cnt(4)
    chips = [oracle.ca_code(p) for p in prns]

    done_at = None
    for snap in range(1200):
        ms = snap
        s = sig[snap % sig.shape[0]]
        rlib.ref_set_packet_cnt(ms)
        rlib.acquisition_process(rchans, s.ctypes.data)
        rlib.gps_master_handling(rchans, 0)
        lib.gpsb_host_set_packet_cnt(ms)
        for i in range(4):                                  # acquisition_process (acquisition.c:51-57)
            plan = Plan()
            lib.gpsb_host_plan_acq(mine.at(i), ms, C.byref(plan))
            if plan.want == WANT_SEARCH:
                res = oracle_search(oracle, chips[i], s, plan.search)
                lib.gpsb_host_finish_acq(mine.at(i), C.byref(plan), C.byref(res))
        lib.gps_master_handling(mine.base, 0)
        for i in range(4):
            r = RefFlat()
            rlib.ref_channel_snapshot(rlib.ref_channel_at(rchans, i), C.byref(r))
            a = mine.snapshot(i)
            assert bytes(a) == bytes(r), (snap, i, diff_fields(a, r))
        assert lib.gps_master_need_acq() == rlib.gps_master_need_acq()
        if not lib.gps_master_need_acq():
            done_at = snap
            break
    assert done_at is not None, "acquisition did not finish"
    for i in range(4):
        st = mine.snapshot(i)
        assert st.acq_state == 9 and st.trk_state == 1     # GPS_ACQ_DONE, GPS_NEED_PRE_TRACK
    # physically sensible too (the reference's Doppler vote is coarse: an adjacent 500-Hz bin may win)
    assert abs(mine.snapshot(0).found_freq_offset_hz - 1020) <= 520 and abs(mine.snapshot(0).found_code_phase - 1990) <= 3
    mine.free()


def test_chain_vote_and_empty_windows():
    lib = load_host_library()
    mine = Channels([3])
    lib.gpsb_host_attach(None)
    st = mine.snapshot(0)
    st.acq_state, st.acq_code_search_start, st.acq_code_search_stop, st.code_hist_step = 3, 7, 7, 64
    mine.restore(0, st)
    plan = Plan()
    lib.gpsb_host_plan_acq(mine.at(0), 5, C.byref(plan))
    assert plan.want == WANT_SEARCH and plan.search.start == plan.search.stop == 7
    lib.gpsb_host_finish_acq(mine.at(0), C.byref(plan), None)      # empty window: nothing voted
    assert list(mine.snapshot(0).code_phase_histogram) == [0] * 32
    # prn 0 = unused channel: nothing is planned (acquisition.c:136)
    idle = Channels([0])
    lib.gpsb_host_plan_acq(idle.at(0), 5, C.byref(plan))
    assert plan.want == WANT_NOTHING
    lib.gpsb_host_plan_track(idle.at(0), 5, 0, C.byref(plan))
    assert plan.want == WANT_NOTHING
    mine.free()
    idle.free()


def test_prompt_offset_2046_edge(oracle, reference, golden):
    """code_phase_fine in (16368, 16376): prompt byte offset 2046, early 2045, late 0 (tracking.c:115-130).
    The reference evaluates offset 2046 as offset 0; so must we - state compared after the step."""
    lib = load_host_library()
    sig = golden["scene_signal"]
    prn = 5
    chips = oracle.ca_code(prn)
    for fine in (16369.25, 16375.9, 16368.0, 0.4, 7.99):
        rchans = reference.channels(1)
        rch = reference.channel_at(rchans, 0)
        reference.channel_init(rch, prn, 0)
        mine = Channels([prn])
        lib.gpsb_host_attach(None)
        st = reference.snapshot(rch)
        st.acq_state, st.trk_state = 9, 4
        st.if_freq_offset_hz_bits = int(np.float32(1020.0).view(np.uint32))
        st.code_phase_fine_bits = int(np.float32(fine).view(np.uint32))
        st.prev_track_timestamp = 99
        reference.restore(rch, st)
        mst = mine.snapshot(0)
        for name, _ in st._fields_:
            setattr(mst, name, getattr(st, name))
        mine.restore(0, mst)
        for ms in (100, 101):
            reference.set_ms(ms)
            reference.lib.gps_tracking_process(rch, sig[ms].ctypes.data, ms % 4)
            host_track_ms(lib, oracle, mine.at(0), chips, sig[ms], ms, ms % 4)
            a, b = mine.snapshot(0), reference.snapshot(rch)
            assert states_equal(a, b), (fine, ms, diff_fields(a, b))
        mine.free()
