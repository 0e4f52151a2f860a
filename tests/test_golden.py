"""Oracle vs the committed golden vectors (tests/golden/golden_l2.npz, produced from the unmodified
reference by tests/golden/make_golden.py).  CPU only; this is what pins the oracle on a box that has
neither /root/reference nor a prebuilt oracle/_ref."""
import numpy as np

IF_HZ = 4092000


def test_codes(oracle, golden):
    codes = np.unpackbits(golden["ca_codes_packed"], axis=1)[:, :1023]
    for prn in range(1, 211):
        assert np.array_equal(oracle.ca_code(prn), codes[prn - 1])


def test_replicas(oracle, golden):
    for i, prn in enumerate(golden["replica_prns"]):
        chips = oracle.ca_code(int(prn))
        for b in range(16):
            want = golden["replica_words"][i, b].view(np.uint8)
            assert np.array_equal(oracle.replica(chips, b)[:2046], want)


def test_nco(oracle, golden):
    for off, acc in zip(golden["nco_freq_offsets"], golden["nco_acc_after_511"]):
        f = np.float32(IF_HZ) + np.float32(off)
        assert (511 * oracle.nco_step32(f)) & 0xFFFFFFFF == int(acc)


def test_mixer(oracle, golden):
    sig = golden["rnd_signal"]
    for f, gi, gq in zip(golden["mix_freqs"], golden["mix_i"], golden["mix_q"]):
        di, dq, _ = oracle.mix(sig, 0, oracle.nco_step32(f))
        assert np.array_equal(di[:2044], gi) and np.array_equal(dq[:2044], gq)


def test_raw_correlator(oracle, golden):
    pad = lambda w: np.concatenate([w, np.zeros(1, np.uint16)]).view(np.uint8).copy()
    rep, di, dq = pad(golden["raw_prn"]), pad(golden["raw_i"]), pad(golden["raw_q"])
    for off in range(2046):
        assert oracle.correlation_iq(rep, di, dq, off) == tuple(golden["raw_iq"][off])
        assert oracle.correlation8(rep, di, dq, off) == golden["raw_corr8"][off]
    for (a0, a1), want in zip(golden["raw_windows"], golden["raw_search"]):
        assert oracle.correlation_search(rep, di, dq, int(a0), int(a1)) == tuple(want)


def test_simulator_kats(oracle, golden):
    chips = oracle.ca_code(1)
    for i, noise in enumerate(golden["sim_noise"]):
        sig = golden["sim_buffers"][i]
        assert oracle.search_cell(chips, sig, float(IF_HZ + 2000), 0, 0, 2046) == tuple(golden["sim_search"][i])
    assert tuple(golden["sim_search"][0]) == (7904, 100, 65)
    assert tuple(golden["sim_iq_all"][0][100]) == (32, 7904)
    sig = golden["sim_buffers"][1]
    step32 = oracle.nco_step32(np.float32(IF_HZ + 2000))
    di, dq, _ = oracle.mix(sig, 0, step32)
    for b in (0, 1, 7, 8, 15):
        rep = oracle.replica(chips, b)
        for off in list(range(0, 2046, 37)) + [1, 2045, 99, 100, 101]:
            assert oracle.correlation_iq(rep, di, dq, off) == tuple(golden["sim15_iq_bits"][b][off])


def test_scene_sweep(oracle, golden):
    sig = golden["scene_signal"]
    sweep = golden["scene_sweep"]
    rng = np.random.default_rng(1)
    cells = [(s, b, m) for s in range(3) for b in range(21) for m in range(4)]
    for idx in rng.choice(len(cells), 30, replace=False):
        s, b, m = cells[idx]
        chips = oracle.ca_code(int(golden["scene_prns"][s]))
        got = oracle.search_cell(chips, sig[m], float(IF_HZ - 5000 + 500 * b), 0, 0, 2046)
        assert got == (sweep[s, b, m, 0], sweep[s, b, m, 1], sweep[s, b, m, 2])
    sw3 = golden["scene_sweep_bits3"]
    for s in range(3):
        chips = oracle.ca_code(int(golden["scene_prns"][s]))
        got = oracle.search_cell(chips, sig[1], float(IF_HZ - 3000 + 500 * 2), 3, 0, 2046)
        assert got == tuple(sw3[s, 2, 1])
