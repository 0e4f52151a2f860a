"""Parity of the CUDA path (through the C ABI, libgpsb_cuda.so) against the oracle and the golden
vectors.  Integer work: everything here is compared bit-exact (no tolerances).  Needs a B200."""
import numpy as np
import pytest

from stm32f4_sdr_gps_b200 import EPL_REQ, SEARCH_REQ, GpsbError, nco_step32

pytestmark = pytest.mark.gpu
IF_HZ = 4092000


def _pad_words(w):
    return np.concatenate([np.asarray(w, np.uint16), np.zeros(1, np.uint16)])


def test_device_code_generator_all_prns(engine, oracle):
    for prn in list(range(1, 38)) + [64, 120, 158, 210]:
        engine.set_code_prn(0, prn)
        assert np.array_equal(engine.get_code(0), oracle.ca_code(prn)), prn
    with pytest.raises(GpsbError):
        engine.set_code_prn(0, 0)
    with pytest.raises(GpsbError):
        engine.set_code_prn(0, 211)
    with pytest.raises(GpsbError):
        engine.set_code_prn(engine.max_sv, 1)


def test_set_code_roundtrip(engine, oracle):
    chips = oracle.ca_code(17)
    engine.set_code(3, chips)
    assert np.array_equal(engine.get_code(3), chips)


def test_l0_replica_all_shifts(engine, oracle, golden):
    for i, prn in enumerate(golden["replica_prns"]):
        chips = oracle.ca_code(int(prn))
        for b in range(16):
            got = engine.l0_generate_prn_data2(chips, b)
            assert np.array_equal(got, golden["replica_words"][i, b]), (prn, b)


def test_l0_mixer(engine, oracle, golden):
    sig = golden["rnd_signal"]
    for f, gi, gq in zip(golden["mix_freqs"], golden["mix_i"], golden["mix_q"]):
        di = np.full(2046, 0xAB, np.uint8)
        dq = np.full(2046, 0xCD, np.uint8)
        acc = engine.l0_shift_to_zero_freq(sig, 77, nco_step32(f), di, dq)
        oi, oq, oacc = oracle.mix(sig, 77, nco_step32(f))
        assert np.array_equal(di[:2044], oi[:2044]) and np.array_equal(dq[:2044], oq[:2044])
        assert di[2044] == 0xAB and dq[2045] == 0xCD and acc == oacc
        di0 = np.zeros(2046, np.uint8)
        dq0 = np.zeros(2046, np.uint8)
        engine.l0_shift_to_zero_freq(sig, 0, nco_step32(f), di0, dq0)
        assert np.array_equal(di0[:2044], gi) and np.array_equal(dq0[:2044], gq)


def test_l0_raw_correlator_golden(engine, golden):
    prn, di, dq = _pad_words(golden["raw_prn"]), _pad_words(golden["raw_i"]), _pad_words(golden["raw_q"])
    for off in list(range(0, 2046, 61)) + [1, 2, 3, 1021, 1022, 1023, 2043, 2044, 2045]:
        assert engine.l0_correlation_iq(prn, di, dq, off) == tuple(golden["raw_iq"][off]), off
        assert engine.l0_correlation8(prn, di, dq, off) == golden["raw_corr8"][off], off
    for (a0, a1), want in zip(golden["raw_windows"], golden["raw_search"]):
        assert engine.l0_correlation_search(prn, di, dq, int(a0), int(a1)) == tuple(want), (a0, a1)


def test_simulator_kat_through_fused_search(engine, golden):
    """SS/main.c:59-68: best_phase == 100 on the reference simulator buffer."""
    engine.set_code_prn(1, 1)
    engine.upload_signal(0, golden["sim_buffers"])
    rq = np.zeros(4, SEARCH_REQ)
    rq["sv_slot"], rq["ms_index"] = 1, np.arange(4)
    rq["step32"], rq["stop"] = nco_step32(IF_HZ + 2000), 2046
    res = engine.search(rq)
    for i in range(4):
        assert (res["max"][i], res["phase"][i], res["avg"][i]) == tuple(golden["sim_search"][i])
    assert (res["max"][0], res["phase"][0], res["avg"][0]) == (7904, 100, 65)


def test_search_iq_every_offset_every_shift(engine, golden):
    engine.set_code_prn(1, 1)
    engine.upload_signal(0, golden["sim_buffers"])
    rq = np.zeros(1, SEARCH_REQ)
    rq["sv_slot"], rq["ms_index"], rq["step32"], rq["stop"] = 1, 1, nco_step32(IF_HZ + 2000), 2046
    for b in range(16):
        rq["off_bits"] = b
        assert np.array_equal(engine.search_iq(rq), golden["sim15_iq_bits"][b]), b
    for i in range(4):
        rq["off_bits"], rq["ms_index"] = 0, i
        assert np.array_equal(engine.search_iq(rq), golden["sim_iq_all"][i])


def test_epl_random_requests_vs_oracle(engine, oracle):
    rng = np.random.default_rng(42)
    n_ms, prns = 16, [2, 9, 31]
    sig = rng.integers(0, 256, (n_ms, 2046), dtype=np.uint8)
    engine.upload_signal(32, sig)
    for s, prn in enumerate(prns):
        engine.set_code_prn(s, prn)
    n = 600
    rq = np.zeros(n, EPL_REQ)
    rq["sv_slot"] = rng.integers(0, 3, n)
    rq["ms_index"] = 32 + rng.integers(0, n_ms, n)
    rq["acc0"] = rng.integers(0, 2**32, n, dtype=np.uint64)
    rq["step32"] = rng.integers(0, 2**32, n, dtype=np.uint64)
    rq["off_p"] = rng.integers(0, 2046, n)
    rq["off_p"][:8] = [0, 1, 2, 2043, 2044, 2045, 1022, 1023]
    rq["off_e"] = np.where(rq["off_p"] == 0, 2045, rq["off_p"].astype(np.int32) - 1)
    rq["off_l"] = np.where(rq["off_p"] == 2045, 0, rq["off_p"] + 1)
    rq["off_e"][8:40] = rng.integers(0, 2046, 32)      # arbitrary, unrelated arms are legal too
    rq["off_l"][8:40] = rng.integers(0, 2046, 32)
    rq["off_bits"] = rng.integers(0, 16, n)
    out = engine.track_epl(rq)
    for i in range(n):
        want = oracle.epl_explicit(oracle.ca_code(prns[rq["sv_slot"][i]]), sig[rq["ms_index"][i] - 32],
                                   int(rq["acc0"][i]), int(rq["step32"][i]), int(rq["off_e"][i]),
                                   int(rq["off_p"][i]), int(rq["off_l"][i]), int(rq["off_bits"][i]))
        assert np.array_equal(out[i], want), (i, rq[i])


def test_epl_matches_reference_trace(engine, golden):
    """Replay the reference's closed-loop trajectory open loop: given the state the reference had
    before each step, the device must reproduce its six sums (tracking.c:115-138)."""
    from oracle_lib import Oracle
    orc = Oracle()
    sig = golden["scene_signal"]
    engine.upload_signal(0, sig[:256])
    for s, prn in enumerate((5, 14)):
        engine.set_code_prn(s, prn)
        iq = golden["track_iq"][s]
        st = golden["track_state_bits"][s].view(np.float32)
        # steps where tracking ran both at k-1 and k: state before k == logged state after k-1
        ks = [k for k in range(1, 256) if iq[k].any() and iq[k - 1].any()]
        assert len(ks) > 50
        rq = np.zeros(len(ks), EPL_REQ)
        want = []
        for j, k in enumerate(ks):
            fine, foff = st[k - 1]
            e, p, l, bits = orc.epl_offsets(fine)
            rq[j] = (s, k, 0, nco_step32(np.float32(IF_HZ) + np.float32(foff)), e, p, l, bits)
        # the reference's accumulator is not logged; recover parity on |I|^2+|Q|^2-independent
        # quantities by using the oracle for acc0-dependent sums instead
        out = engine.track_epl(rq)
        for j, k in enumerate(ks):
            w = orc.epl_explicit(orc.ca_code((5, 14)[s]), sig[k], 0, int(rq["step32"][j]), int(rq["off_e"][j]),
                                 int(rq["off_p"][j]), int(rq["off_l"][j]), int(rq["off_bits"][j]))
            assert np.array_equal(out[j], w)


def test_search_windows_vs_oracle(engine, oracle):
    rng = np.random.default_rng(7)
    sig = rng.integers(0, 256, (8, 2046), dtype=np.uint8)
    sig[5] = 0                       # all-zero millisecond
    engine.upload_signal(100, sig)
    prns = [4, 11, 23, 28]
    for s, prn in enumerate(prns):
        engine.set_code_prn(s, prn)
    windows = [(0, 2046), (0, 1), (2045, 2046), (100, 107), (1000, 1500), (2016, 2046), (5, 5), (9, 3), (0, 30)]
    rq = np.zeros(len(windows) * 4, SEARCH_REQ)
    meta = []
    for i in range(rq.size):
        a0, a1 = windows[i % len(windows)]
        f = np.float32(IF_HZ + int(rng.integers(-14, 15)) * 500)
        bits = int(rng.integers(0, 16)) if i % 3 == 0 else 0
        rq[i] = (i % 4, 100 + i % 8, 0, nco_step32(f), bits, a0, a1, 0)
        meta.append((prns[i % 4], i % 8, f, bits, a0, a1))
    res = engine.search(rq)
    for i, (prn, m, f, bits, a0, a1) in enumerate(meta):
        want = oracle.search_cell(oracle.ca_code(prn), sig[m], float(f), bits, a0, a1)
        assert (res["max"][i], res["phase"][i], res["avg"][i]) == want, (i, meta[i])


def test_sweep_matches_reference_golden(engine, golden):
    """Cold-acquisition cells (acquisition.c:282-294) on the 2-satellite scene: every triple equals
    what the unmodified reference computed."""
    sig = golden["scene_signal"]
    engine.upload_signal(0, sig[:8])
    for s, prn in enumerate(golden["scene_prns"]):
        engine.set_code_prn(s, int(prn))
    step = [nco_step32(np.float32(IF_HZ - 5000 + 500 * b)) for b in range(21)]
    res = engine.sweep([0, 1, 2], step, 0, 4, 0)
    want = golden["scene_sweep"]
    assert np.array_equal(res["max"], want[..., 0])
    assert np.array_equal(res["phase"], want[..., 1])
    assert np.array_equal(res["avg"], want[..., 2])
    step3 = [nco_step32(np.float32(IF_HZ - 3000 + 500 * b)) for b in range(3)]
    res3 = engine.sweep([0, 1, 2], step3, 0, 2, 3)
    w3 = golden["scene_sweep_bits3"]
    assert np.array_equal(res3["max"], w3[..., 0]) and np.array_equal(res3["phase"], w3[..., 1])
    assert np.array_equal(res3["avg"], w3[..., 2])
    # the strong satellites are found where the scene put them
    assert res["phase"][0, 12, 0] == 1990 and res["phase"][1, 5, 0] in (101, 102)


def test_sweep_full_size_properties(engine, oracle):
    """BASELINE config 3 size (32 SV x 21 bins x 10 ms): size-independent properties - the sweep
    equals the per-cell search, and a sample of cells equals the oracle."""
    from stm32f4_sdr_gps_b200.signal_synth import config3_scene, synthesize
    scene = config3_scene(n_ms=10)
    sig = synthesize(scene)
    engine.upload_signal(0, sig)
    for prn in range(1, 33):
        engine.set_code_prn(prn - 1, prn)
    step = np.array([nco_step32(np.float32(IF_HZ - 5000 + 500 * b)) for b in range(21)], np.uint32)
    res = engine.sweep(np.arange(32), step, 0, 10, 0)
    assert res.shape == (32, 21, 10)
    rng = np.random.default_rng(9)
    idx = rng.integers(0, 32 * 21 * 10, 64)
    rq = np.zeros(idx.size, SEARCH_REQ)
    for j, c in enumerate(idx):
        s, b, m = c // 210, (c // 10) % 21, c % 10
        rq[j] = (s, m, 0, step[b], 0, 0, 2046, 0)
    single = engine.search(rq)
    flat = res.reshape(-1)
    assert np.array_equal(single["max"], flat["max"][idx]) and np.array_equal(single["phase"], flat["phase"][idx])
    for j, c in enumerate(idx[:12]):
        s, b, m = c // 210, (c // 10) % 21, c % 10
        want = oracle.search_cell(oracle.ca_code(int(s) + 1), sig[m], float(IF_HZ - 5000 + 500 * b), 0, 0, 2046)
        assert (flat["max"][c], flat["phase"][c], flat["avg"][c]) == want
    # sanity of the scene itself: the present satellites show up at their true code phase in their nearest
    # Doppler bin in a good share of the cells (the reference's 1-ms half-wave detector only fires in one
    # carrier quadrant, gps_misc.c:111-114, so not every ms is a hit)
    hits = 0
    for sat in scene.sats:
        b = int(round((sat.doppler_hz + 5000) / 500.0))
        ph = res["phase"][sat.prn - 1, b, :].astype(np.float64)
        true = sat.code_phase_samples / 8.0
        d = np.minimum(np.abs(ph - true), 2046 - np.abs(ph - true))
        hits += int((d < 2.5).sum())
    assert hits >= 30, hits


def test_ingest_iq2_adaptor(engine):
    from stm32f4_sdr_gps_b200.signal_synth import iq2_from_packed
    rng = np.random.default_rng(12)
    packed = rng.integers(0, 256, (5, 2046), dtype=np.uint8)
    engine.upload_signal_iq2(200, iq2_from_packed(packed))
    assert np.array_equal(engine.download_signal(200, 5), packed)


def test_ring_wraparound_and_errors(engine):
    rng = np.random.default_rng(13)
    packed = rng.integers(0, 256, (6, 2046), dtype=np.uint8)
    ms0 = engine.ring_ms - 3                      # wraps after three frames
    engine.upload_signal(ms0, packed)
    assert np.array_equal(engine.download_signal(ms0, 6), packed)
    with pytest.raises(GpsbError):
        engine.upload_signal(0, np.zeros((engine.ring_ms + 1, 2046), np.uint8))
    rq = np.zeros(1, EPL_REQ)
    rq["sv_slot"] = engine.max_sv
    with pytest.raises(GpsbError):
        engine.track_epl(rq)
    rq["sv_slot"], rq["off_p"] = 0, 2046
    with pytest.raises(GpsbError):
        engine.track_epl(rq)
    assert engine.track_epl(np.zeros(0, EPL_REQ)).shape == (0, 6)     # empty batch is a no-op
    assert engine.search(np.zeros(0, SEARCH_REQ)).size == 0
    rq = np.zeros(1, EPL_REQ)
    rq["sv_slot"] = 39                            # slot never given a code
    with pytest.raises(GpsbError):
        engine.track_epl(rq)


def test_dp4a_and_direct_search_agree_all_bit_shifts(engine, oracle):
    """The two wide-window implementations (include/gpsb.h GPSB_SWEEP_*) are bit-identical, for every
    sub-byte replica shift, on random data, an all-ones and an all-zero millisecond; a sample of cells is
    also checked against the oracle."""
    rng = np.random.default_rng(21)
    sig = rng.integers(0, 256, (6, 2046), dtype=np.uint8)
    sig[4] = 0xFF
    sig[5] = 0
    engine.upload_signal(40, sig)
    prns = [1, 6, 13, 17, 22, 26, 29, 30, 32]           # 9 satellites: one full tile of 8 + a tile of 1
    for s, prn in enumerate(prns):
        engine.set_code_prn(s, prn)
    step = np.array([nco_step32(np.float32(IF_HZ + d)) for d in (-4500, 0, 2250)], np.uint32)
    for bits in range(16):
        engine.set_sweep_method(1)
        a = engine.sweep(np.arange(9), step, 40, 6, bits)
        engine.set_sweep_method(0)
        b = engine.sweep(np.arange(9), step, 40, 6, bits)
        engine.set_sweep_method(1)
        assert np.array_equal(a, b), bits
        if bits in (0, 5, 15):
            for (s, k, m) in ((0, 0, 0), (8, 2, 3), (4, 1, 4), (7, 2, 5)):
                f = float(np.float32(IF_HZ + (-4500, 0, 2250)[k]))
                want = oracle.search_cell(oracle.ca_code(prns[s]), sig[m], f, bits, 0, 2046)
                assert (a["max"][s, k, m], a["phase"][s, k, m], a["avg"][s, k, m]) == want, (bits, s, k, m)
    # grouped host requests: same millisecond / NCO, different satellites and windows, mixed with narrow ones
    rq = np.zeros(12, SEARCH_REQ)
    for i in range(12):
        rq[i] = (i % 9, 40 + (i // 6), 0, step[1], 2, 0 if i % 3 else 100, 2046 if i % 4 else 1500, 0)
    rq[5]["start"], rq[5]["stop"] = 700, 730              # narrow: direct kernel
    rq[7]["start"], rq[7]["stop"] = 30, 30                # empty
    res = engine.search(rq)
    for i in range(12):
        want = oracle.search_cell(oracle.ca_code(prns[i % 9]), sig[i // 6], float(np.float32(IF_HZ)), 2,
                                  int(rq[i]["start"]), int(rq[i]["stop"]))
        if rq[i]["start"] >= rq[i]["stop"]:
            want = (0, 0, 0)
        assert (res["max"][i], res["phase"][i], res["avg"][i]) == want, i


def test_closed_loop_paths_agree(engine, oracle):
    """The three ways a small E/P/L batch can reach the GPU - staged copies, parameter-space launch with
    mapped-memory completion (k_epl_rt), resident session kernel (k_epl_session) - return identical sums,
    also when the ring frame is rewritten while the session kernel is resident."""
    rng = np.random.default_rng(31)
    sig = rng.integers(0, 256, (4, 2046), dtype=np.uint8)
    engine.upload_signal(60, sig)
    for s, prn in enumerate((8, 15, 21)):
        engine.set_code_prn(s, prn)
    rq = np.zeros(3, EPL_REQ)
    for i in range(3):
        rq[i] = (i, 60 + i, 1000 * i, 0x40000000 + 12345 * i, 500 + i, 501 + i, 502 + i, i)
    want = np.stack([oracle.epl_explicit(oracle.ca_code((8, 15, 21)[i]), sig[i], int(rq["acc0"][i]), int(rq["step32"][i]),
                                         500 + i, 501 + i, 502 + i, i) for i in range(3)])
    engine.set_realtime(False)
    assert np.array_equal(engine.track_epl(rq), want)
    engine.set_realtime(True)
    assert np.array_equal(engine.track_epl(rq), want)
    engine.session_begin(4)
    try:
        for _ in range(50):
            assert np.array_equal(engine.track_epl(rq), want)
        assert np.array_equal(engine.track_epl(rq[:1]), want[:1])      # fewer cells than slots
        sig2 = rng.integers(0, 256, (4, 2046), dtype=np.uint8)
        engine.upload_signal(60, sig2)                                  # DMA under the resident kernel
        want2 = np.stack([oracle.epl_explicit(oracle.ca_code((8, 15, 21)[i]), sig2[i], int(rq["acc0"][i]),
                                              int(rq["step32"][i]), 500 + i, 501 + i, 502 + i, i) for i in range(3)])
        assert np.array_equal(engine.track_epl(rq), want2)
        import time
        time.sleep(0.6)                                                 # idle time-out: the kernel leaves ...
        assert np.array_equal(engine.track_epl(rq), want2)              # ... and is re-launched on demand
        with pytest.raises(GpsbError):
            engine.session_begin(2)                                     # already open
    finally:
        engine.session_end()
    assert np.array_equal(engine.track_epl(rq), want2)


def test_sweep_tile_shapes_ring_wrap_and_big_batches(engine, oracle):
    """Ragged / maximum-size inputs: every satellite-tile variant of the dp4a sweep (1..9 satellites), a
    sweep whose milliseconds wrap around the end of the signal ring, an E/P/L batch larger than the
    closed-loop fast path (n > 128), and one of exactly 128."""
    rng = np.random.default_rng(55)
    ring = engine.ring_ms
    sig = rng.integers(0, 256, (3, 2046), dtype=np.uint8)
    ms0 = ring * 3 - 1                                     # frames ring-1, 0, 1
    engine.upload_signal(ms0, sig)
    prns = list(range(2, 11))
    for s, prn in enumerate(prns):
        engine.set_code_prn(s, prn)
    step = np.array([nco_step32(np.float32(IF_HZ + 750))], np.uint32)
    full = None
    for n_sv in (9, 5, 4, 3, 2, 1):
        res = engine.sweep(np.arange(n_sv), step, ms0, 3, 1)
        if full is None:
            full = res
            for (s, m) in ((0, 0), (8, 1), (3, 2)):
                want = oracle.search_cell(oracle.ca_code(prns[s]), sig[m], float(np.float32(IF_HZ + 750)), 1, 0, 2046)
                assert (res["max"][s, 0, m], res["phase"][s, 0, m], res["avg"][s, 0, m]) == want
        else:
            assert np.array_equal(res, full[:n_sv]), n_sv
    assert engine.sweep(np.zeros(0, np.uint32), step, ms0, 3, 0).size == 0
    for n in (128, 129, 700):
        rq = np.zeros(n, EPL_REQ)
        rq["sv_slot"] = rng.integers(0, 9, n)
        rq["ms_index"] = ms0 + rng.integers(0, 3, n)
        rq["acc0"] = rng.integers(0, 2**32, n, dtype=np.uint64)
        rq["step32"] = rng.integers(0, 2**32, n, dtype=np.uint64)
        rq["off_p"] = rng.integers(0, 2046, n)
        rq["off_e"] = rng.integers(0, 2046, n)
        rq["off_l"] = rng.integers(0, 2046, n)
        rq["off_bits"] = rng.integers(0, 8, n)
        out = engine.track_epl(rq)
        engine.set_realtime(False)
        assert np.array_equal(engine.track_epl(rq), out)
        engine.set_realtime(True)
        for i in rng.integers(0, n, 12):
            want = oracle.epl_explicit(oracle.ca_code(prns[rq["sv_slot"][i]]), sig[rq["ms_index"][i] - ms0],
                                       int(rq["acc0"][i]), int(rq["step32"][i]), int(rq["off_e"][i]),
                                       int(rq["off_p"][i]), int(rq["off_l"][i]), int(rq["off_bits"][i]))
            assert np.array_equal(out[i], want), (n, i)


def test_epl_batch_kernel_vs_oracle_and_cta_kernel(engine, oracle):
    """k_epl_batch (one warp per cell, frames streamed into registers, resident extended replica tables) against
    the oracle and against k_epl on the same requests: every sub-byte shift 0..15, the offsets around the period
    seam and around the odd-offset exclusions, unrelated arms, ring wrap, a batch that is not a multiple of the
    warp count; and the prompt-only form gpsb_prompt_iq."""
    rng = np.random.default_rng(4242)
    ring = engine.ring_ms
    n_ms, prns = 9, [1, 17, 32]
    sig = rng.integers(0, 256, (n_ms, 2046), dtype=np.uint8)
    ms0 = ring * 7 - 4                                       # wraps around the end of the ring
    engine.upload_signal(ms0, sig)
    for s, prn in enumerate(prns):
        engine.set_code_prn(s, prn)
    n = 2600 + 13
    rq = np.zeros(n, EPL_REQ)
    rq["sv_slot"] = rng.integers(0, 3, n)
    rq["ms_index"] = ms0 + rng.integers(0, n_ms, n)
    rq["acc0"] = rng.integers(0, 2**32, n, dtype=np.uint64)
    rq["step32"] = rng.integers(0, 2**32, n, dtype=np.uint64)
    rq["step32"][::3] &= 0x07FFFFFF                          # realistic Doppler range too
    rq["off_p"] = rng.integers(0, 2046, n)
    special = [0, 1, 2, 3, 4, 5, 2040, 2041, 2042, 2043, 2044, 2045, 1021, 1022, 1023, 1024]
    rq["off_p"][:len(special)] = special
    rq["off_e"] = np.where(rq["off_p"] == 0, 2045, rq["off_p"].astype(np.int32) - 1)
    rq["off_l"] = np.where(rq["off_p"] == 2045, 0, rq["off_p"] + 1)
    rq["off_e"][100:400] = rng.integers(0, 2046, 300)        # arbitrary, unrelated arms
    rq["off_l"][100:400] = rng.integers(0, 2046, 300)
    rq["off_bits"] = rng.integers(0, 16, n)
    engine.set_epl_batch_min(0)                              # every size through k_epl_batch
    try:
        engine.set_realtime(False)
        out = engine.track_epl(rq)
        small = engine.track_epl(rq[:5])
        prompt = engine.prompt_iq(rq)
    finally:
        engine.set_epl_batch_min(2**32 - 1)                  # never: the one-CTA-per-cell kernel
    try:
        ref = engine.track_epl(rq)
    finally:
        engine.set_epl_batch_min(512)
        engine.set_realtime(True)
    assert np.array_equal(out, ref)
    assert np.array_equal(small, ref[:5])
    assert np.array_equal(prompt, ref[:, 2:4])
    check = list(range(len(special))) + list(range(100, 130)) + [int(i) for i in rng.integers(0, n, 60)] + [n - 1]
    for i in check:
        want = oracle.epl_explicit(oracle.ca_code(prns[rq["sv_slot"][i]]), sig[rq["ms_index"][i] - ms0],
                                   int(rq["acc0"][i]), int(rq["step32"][i]), int(rq["off_e"][i]),
                                   int(rq["off_p"][i]), int(rq["off_l"][i]), int(rq["off_bits"][i]))
        assert np.array_equal(out[i], want), (i, rq[i])
    assert engine.prompt_iq(np.zeros(0, EPL_REQ)).shape == (0, 2)
    junk = rq[:700].copy()                                   # the early / late fields play no part in the prompt-only form
    junk["off_e"], junk["off_l"] = 65535, 40000
    assert np.array_equal(engine.prompt_iq(junk), ref[:700, 2:4])


def test_epl_batch_full_size_properties(oracle):
    """k_epl_batch at the size of SURVEY.md section 8(d) config 1 'batched' (10^5 cells, one per millisecond of a 205-MB
    recording), through size-independent properties: the prompt-only form equals the prompt arm of the three-arm
    form; the result of a request does not depend on where in the batch it stands (a permuted batch gives the
    permuted results); cells drawn at random equal the oracle; every sum lies in the range a popcount can produce."""
    from stm32f4_sdr_gps_b200 import Engine
    n = 100_000
    rng = np.random.default_rng(8184)
    sig = rng.integers(0, 256, (n, 2046), dtype=np.uint8)
    with Engine(device=0, max_sv=4, ring_ms=n) as eng:
        eng.set_code_prn(1, 1)
        eng.set_code_prn(2, 19)
        eng.upload_signal(0, sig)
        rq = np.zeros(n, EPL_REQ)
        rq["sv_slot"] = 1 + (np.arange(n) % 2)
        rq["ms_index"] = np.arange(n)
        rq["acc0"] = rng.integers(0, 2**32, n, dtype=np.uint64)
        rq["step32"] = nco_step32(np.float32(IF_HZ + 2000))
        rq["off_p"] = rng.integers(1, 2045, n)
        rq["off_e"], rq["off_l"] = rq["off_p"] - 1, rq["off_p"] + 1
        rq["off_bits"] = rng.integers(0, 8, n)
        epl = eng.track_epl(rq)
        prompt = eng.prompt_iq(rq)
        assert np.array_equal(prompt, epl[:, 2:4])
        perm = rng.permutation(n)
        assert np.array_equal(eng.track_epl(rq[perm]), epl[perm])
        assert epl.min() >= -8184 and epl.max() <= 8184
        chips = {1: oracle.ca_code(1), 2: oracle.ca_code(19)}
        for i in [0, n - 1] + [int(x) for x in rng.integers(0, n, 40)]:
            want = oracle.epl_explicit(chips[int(rq["sv_slot"][i])], sig[i], int(rq["acc0"][i]), int(rq["step32"][i]),
                                       int(rq["off_e"][i]), int(rq["off_p"][i]), int(rq["off_l"][i]), int(rq["off_bits"][i]))
            assert np.array_equal(epl[i], want), (i, rq[i])


def test_sweep_gather_single_rank_equals_sweep(engine, golden):
    """gpsb_sweep_gather without a communicator (one rank): the group-sharded launch (dense per-rank block) + k_unshard
    must give exactly gpsb_sweep's grid - for tiles of 8, 4 and 1 satellites and a sub-byte shift.  (Two and eight
    ranks: bench.py asserts the gathered grid equal to the one-GPU sweep in every run.)"""
    from stm32f4_sdr_gps_b200 import nco_step32
    sig = golden["scene_signal"]
    engine.upload_signal(0, sig[:12])
    for prn in range(1, 12):
        engine.set_code_prn(prn, prn)
    step = np.array([nco_step32(np.float32(4092000 - 2500 + 500 * b)) for b in range(11)], np.uint32)
    for svs, bits in ((list(range(1, 12)), 0), ([5, 14 % 11 + 1, 3], 5), ([7], 0)):
        want = engine.sweep(svs, step, 1, 10, bits)
        got = engine.sweep_gather(svs, step, 1, 10, bits)
        assert np.array_equal(got, want), (svs, bits)
    assert engine.comm_size == 1


def _config3_ref_cells(args):
    prn, bits = args
    from oracle_lib import Reference
    from stm32f4_sdr_gps_b200.signal_synth import config3_scene, synthesize
    ref = Reference()
    sig = _config3_ref_cells.sig
    chans = ref.channels(1)
    ref.channel_init(ref.channel_at(chans, 0), prn, 0)
    return ref.sweep_cells(chans, 1, sig, sig.shape[0], -5000, 500, 21, bits)[0]


def test_config3_every_cell_of_both_grids_equals_the_reference(reference):
    """BASELINE configs[2] at its stated size: 32 PRNs x 21 bins (-5000 .. +5000 Hz) x 10 ms = 6720 cells, ALL of them,
    on the reference's 2046-phase grid (sub-byte shift 0) and on one sub-byte shift of the 16368-phase grid (shift 3):
    {max, first argmax, average} equal the unmodified reference's gps_generate_prn_data2 + gps_shift_to_zero_freq +
    correlation_search (oracle/_ref), cell by cell."""
    import multiprocessing as mp
    import os
    from stm32f4_sdr_gps_b200 import Engine, nco_step32
    from stm32f4_sdr_gps_b200.signal_synth import config3_scene, synthesize
    sig = synthesize(config3_scene(n_ms=10))
    _config3_ref_cells.sig = sig
    step = np.array([nco_step32(np.float32(4092000 - 5000 + 500 * b)) for b in range(21)], np.uint32)
    with Engine(device=0, max_sv=40, ring_ms=16) as eng:
        eng.upload_signal(0, sig)
        for prn in range(1, 33):
            eng.set_code_prn(prn, prn)
        got = {bits: eng.sweep(np.arange(1, 33), step, 0, 10, bits) for bits in (0, 3)}
    with mp.get_context("fork").Pool(min(32, os.cpu_count() or 1)) as pool:
        for bits in (0, 3):
            want = np.stack(pool.map(_config3_ref_cells, [(prn, bits) for prn in range(1, 33)]))
            g = np.stack([got[bits]["max"], got[bits]["phase"], got[bits]["avg"]], axis=-1)
            bad = np.argwhere((g != want).any(axis=-1))
            assert g.shape == (32, 21, 10, 3) and len(bad) == 0, (bits, len(bad), bad[:3].tolist())
