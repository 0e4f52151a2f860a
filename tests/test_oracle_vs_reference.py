"""Pin the oracle (oracle/gps_oracle.c) against the UNMODIFIED reference C compiled into
oracle/_ref/libgpsref.so, function by function, on the reference simulator buffer and on random
inputs.  CPU only.  Reference citations: Firmware/project_main/GPS/gps_misc.c."""
import ctypes as C

import numpy as np
import pytest

IF_HZ = 4092000


def _ref_corr(ref, prn_w, di_w, dq_w, off):
    a, b = C.c_int16(), C.c_int16()
    ref.lib.gps_correlation_iq(prn_w.ctypes.data, di_w.ctypes.data, dq_w.ctypes.data, off, C.byref(a), C.byref(b))
    return a.value, b.value


def test_ca_codes_all_prns(oracle, reference):
    chans = reference.channels(1)
    ch = reference.channel_at(chans, 0)
    for prn in range(1, 211):
        reference.channel_init(ch, prn, 0)
        assert np.array_equal(oracle.ca_code(prn), reference.prn_code(ch)), prn
    # IS-GPS-200: PRN 1 starts 1100100000 (octal 1440)
    assert "".join(map(str, oracle.ca_code(1)[:10])) == "1100100000"
    with pytest.raises(ValueError):
        oracle.ca_code(0)
    with pytest.raises(ValueError):
        oracle.ca_code(211)


def test_replica_all_bit_shifts(oracle, reference):
    chans = reference.channels(1)
    ch = reference.channel_at(chans, 0)
    for prn in (1, 7, 32):
        reference.channel_init(ch, prn, 0)
        chips = reference.prn_code(ch)
        for bits in range(16):
            buf = np.zeros(1024, np.uint16)
            reference.lib.gps_generate_prn_data2(ch, buf.ctypes.data, bits)
            mine = oracle.replica(chips, bits)
            assert np.array_equal(mine[:2046], buf.view(np.uint8)[:2046]), (prn, bits)


def test_nco_words(oracle, reference):
    assert oracle.nco_step(4092000.0) == 1073741824
    assert oracle.nco_step(4094000.0) == 1074266624
    chans = reference.channels(1)
    ch = reference.channel_at(chans, 0)
    reference.channel_init(ch, 1, 0)
    sig = reference.sim_buffer(0)
    rng = np.random.default_rng(3)
    for off in list(rng.uniform(-7000, 7000, 200).astype(np.float32)) + [0.0, 0.25, -0.25]:
        _, acc_ref = reference.epl_cell(ch, sig, float(off), 12345, 0.0)
        f = np.float32(IF_HZ) + np.float32(off)
        assert (12345 + 511 * oracle.nco_step32(f)) & 0xFFFFFFFF == acc_ref


def test_mixer_matches_and_leaves_tail_untouched(oracle, reference):
    rng = np.random.default_rng(11)
    for trial in range(20):
        sig = rng.integers(0, 256, 2046, dtype=np.uint8)
        f = np.float32(IF_HZ + rng.uniform(-7000, 7000))
        di = np.full(2048, 0xEE, np.uint8)
        dq = np.full(2048, 0x11, np.uint8)
        reference.lib.gps_shift_to_zero_freq(sig.ctypes.data, di.ctypes.data, dq.ctypes.data, f)
        oi, oq, _ = oracle.mix(sig, 0, oracle.nco_step32(f))
        assert np.array_equal(oi[:2044], di[:2044]) and np.array_equal(oq[:2044], dq[:2044])
        assert di[2044] == 0xEE and di[2045] == 0xEE and dq[2044] == 0x11   # loop bound 511 words
        assert oi[2044] == 0 and oi[2045] == 0


def test_correlator_all_offsets_random(oracle, reference):
    """Raw sums incl. the odd-offset word exclusions and non-zero tail bytes (gps_misc.c:48-93)."""
    rng = np.random.default_rng(5)
    for trial in range(2):
        prn_w = rng.integers(0, 65536, 1024, dtype=np.uint16)
        di_w = rng.integers(0, 65536, 1024, dtype=np.uint16)
        dq_w = rng.integers(0, 65536, 1024, dtype=np.uint16)
        rep, di, dq = (x.view(np.uint8).copy() for x in (prn_w, di_w, dq_w))
        for off in range(2046):
            assert oracle.correlation_iq(rep, di, dq, off) == _ref_corr(reference, prn_w, di_w, dq_w, off), off
            c8 = reference.lib.gps_correlation8(prn_w.ctypes.data, di_w.ctypes.data, dq_w.ctypes.data, off)
            assert oracle.correlation8(rep, di, dq, off) == c8, off


def test_detector_rounding_extremes(oracle, reference):
    """Large sums: squares exceed 2^24 so the int->float conversions round (gps_misc.c:116-118)."""
    prn_w = np.zeros(1024, np.uint16)
    for fill_i, fill_q in ((0xFFFF, 0xFFFF), (0xFFFF, 0x0000), (0xFFFE, 0xFF7F), (0x7FFF, 0xFFFF)):
        di_w = np.full(1024, fill_i, np.uint16)
        dq_w = np.full(1024, fill_q, np.uint16)
        rep, di, dq = (x.view(np.uint8).copy() for x in (prn_w, di_w, dq_w))
        for off in (0, 1, 2, 777, 2045):
            c8 = reference.lib.gps_correlation8(prn_w.ctypes.data, di_w.ctypes.data, dq_w.ctypes.data, off)
            assert oracle.correlation8(rep, di, dq, off) == c8


def test_search_windows(oracle, reference):
    rng = np.random.default_rng(8)
    prn_w = rng.integers(0, 65536, 1024, dtype=np.uint16)
    di_w = rng.integers(0, 65536, 1024, dtype=np.uint16)
    dq_w = rng.integers(0, 65536, 1024, dtype=np.uint16)
    rep, di, dq = (x.view(np.uint8).copy() for x in (prn_w, di_w, dq_w))
    for a0, a1 in ((0, 2046), (0, 1), (2045, 2046), (100, 107), (750, 1250), (5, 5), (9, 3)):
        avr, ph = C.c_uint16(), C.c_uint16()
        mx = reference.lib.correlation_search(prn_w.ctypes.data, di_w.ctypes.data, dq_w.ctypes.data, a0, a1,
                                              C.byref(avr), C.byref(ph))
        assert oracle.correlation_search(rep, di, dq, a0, a1) == (mx, ph.value, avr.value)


def test_simulator_known_answers(oracle, reference):
    """SS/main.c:59-68 recipe; BASELINE.md section 4 table."""
    chans = reference.channels(1)
    ch = reference.channel_at(chans, 0)
    reference.channel_init(ch, 1, 0)
    chips = oracle.ca_code(1)
    expect = {0: (7904, 100, 65), 15: (5490, 100, 73), 30: (3093, 100, 65), 45: (692, 100, 46)}
    for noise, triple in expect.items():
        sig = reference.sim_buffer(noise, 1)
        assert reference.search_cell(ch, sig, 2000, 0, 0, 2046) == triple
        assert oracle.search_cell(chips, sig, float(IF_HZ + 2000), 0, 0, 2046) == triple
    sig = reference.sim_buffer(15, 1)
    iq = reference.iq_cell(ch, sig, float(IF_HZ + 2000), 0, 97, 104)
    assert iq.tolist() == [[-47, -81], [20, 78], [13, 2879], [4, 5490], [10, 2674], [44, 54], [36, -66]]


def test_fused_cells_random(oracle, reference):
    rng = np.random.default_rng(17)
    chans = reference.channels(1)
    ch = reference.channel_at(chans, 0)
    for prn in (3, 22):
        reference.channel_init(ch, prn, 0)
        chips = oracle.ca_code(prn)
        for trial in range(40):
            sig = rng.integers(0, 256, 2046, dtype=np.uint8)
            fo = float(np.float32(rng.uniform(-6000, 6000)))
            acc = int(rng.integers(0, 2**32))
            fine = float(np.float32(rng.uniform(0, 16368)))
            if trial == 0:
                fine = 3.0          # prompt offset 0 -> early wraps to 2045 (tracking.c:127-128)
            if trial == 1:
                fine = 16367.5      # late offset 2046 -> 0 (tracking.c:129-130)
            out_ref, acc_ref = reference.epl_cell(ch, sig, fo, acc, fine)
            out_orc, acc_orc = oracle.track_epl(chips, sig, fo, acc, fine)
            assert np.array_equal(out_ref, out_orc) and acc_ref == acc_orc, (prn, trial)
        for bits in (0, 5, 15):
            sig = rng.integers(0, 256, 2046, dtype=np.uint8)
            f = float(np.float32(IF_HZ + rng.integers(-14, 15) * 500))
            a0 = int(rng.integers(0, 1900))
            assert reference.search_cell_f(ch, sig, f, bits, a0, a0 + 120) == \
                oracle.search_cell(chips, sig, f, bits, a0, a0 + 120)


def test_rewind(oracle, reference):
    import ctypes
    lib = reference.lib
    chans = reference.channels(1)
    ch = reference.channel_at(chans, 0)
    reference.channel_init(ch, 1, 0)
    rng = np.random.default_rng(23)
    for _ in range(50):
        st = reference.snapshot(ch)
        acc = int(rng.integers(0, 2**32))
        fo = np.float32(rng.uniform(-6000, 6000))
        steps = int(rng.integers(1, 50))
        st.if_freq_accum = acc
        st.if_freq_offset_hz_bits = int(fo.view(np.uint32))
        reference.restore(ch, st)
        # gps_tracking_t sits inside gps_ch_t; drive gps_rewind_if_phase through a tracking step with a
        # time gap instead of poking at struct offsets: diff_ticks = steps + 1 (tracking.c:102-113)
        st.trk_state = 4
        st.prev_track_timestamp = 1000
        st.code_phase_fine_bits = int(np.float32(800.0).view(np.uint32))
        reference.restore(ch, st)
        reference.set_ms(1000 + steps + 1)
        sig = reference.sim_buffer(0)
        lib.gps_tracking_process(ch, sig.ctypes.data, 0)
        after = reference.snapshot(ch).if_freq_accum
        rew = oracle.lib.orc_rewind_if_phase(acc, fo, steps)
        step32 = oracle.nco_step32(np.float32(IF_HZ) + fo)
        assert (rew + 511 * step32) & 0xFFFFFFFF == after
