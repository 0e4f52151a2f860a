"""The sources of the device-resident tracking loop (core/gpsb_loop_core.h, core/gpsb_epl_core.h) on the CPU.

tests/emu/loop_emu.c compiles exactly what k_track_run is made of - the loop filters with the DEVICE math
selected (lc_atanf / lc_atan2f, private rand stream, deferred SNR) and the raw-frame E/P/L correlator - and
runs the kernel's control flow.  Compared here with the host libm over the whole input domain, with the
oracle cell by cell, and with the UNMODIFIED reference closed loop as raw channel state."""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from emu_lib import load_emulator
from stm32f4_sdr_gps_b200 import Channels, load_host_library
from test_host_logic import diff_fields, host_track_ms, states_equal

IF_HZ = 4092000


def test_device_float_math_equals_host_libm_on_the_whole_domain():
    """atan2f(qp, ip)/pi (ip > 0) and atanf(qp/ip) for EVERY pair of sums in [-8184, 8184]: the fdlibm
    restatement the device runs returns the host libm's bit pattern for all 2 x 134 M inputs."""
    emu = load_emulator()
    jobs = [(k, lo, min(lo + 256, 8185)) for k in (0, 1) for lo in range(-8184, 8185, 256)]

    def run(job):
        bad = (C.c_int32 * 2)()
        return job, emu.emu_compare_float_math(job[0], job[1], job[2], bad), (bad[0], bad[1])

    with ThreadPoolExecutor(max(1, min(16, os.cpu_count() or 1))) as pool:
        results = list(pool.map(run, jobs))
    wrong = [r for r in results if r[1]]
    assert not wrong, wrong[:5]


def test_fold_half_pi_float_compare_equals_double_compare():
    """lc_fold_half_pi compares in float against 0x3FC90FDB; the reference compares in double against M_PI/2.
    Same result for every float from 0.5 up to the largest finite one, both signs (1.1 G values), plus the
    neighbourhood of zero."""
    emu = load_emulator()
    assert emu.emu_fold_mismatches(0x3f000000, 0x7f7fffff) == 0
    assert emu.emu_fold_mismatches(0x00000000, 0x00100000) == 0


def test_private_rand_stream_equals_libc_default_sequence():
    assert load_emulator().emu_rand31_mismatches(200000) == 0


def test_raw_frame_epl_cell_equals_oracle(oracle):
    """The on-the-fly correlator (no staged mixed samples) against the oracle for every work split the kernel
    may use, all sub-byte shifts, even/odd offsets, the offsets around the period seam and wrapped arms."""
    emu = load_emulator()
    rng = np.random.default_rng(2024)
    chips = oracle.ca_code(17)
    cases = []
    for off_p in (0, 1, 2, 3, 4, 5, 1022, 1023, 2042, 2043, 2044, 2045):
        off_e = off_p - 1 if off_p else 2045
        off_l = off_p + 1 if off_p < 2045 else 0
        cases.append((off_e, off_p, off_l))
    for _ in range(40):
        cases.append(tuple(int(v) for v in rng.integers(0, 2046, 3)))          # arms need not be neighbours
    for i, (off_e, off_p, off_l) in enumerate(cases):
        sig = rng.integers(0, 256, 2046, dtype=np.uint8)
        acc0, step32 = int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32))
        bits = i % 16
        want = oracle.epl_explicit(chips, sig, acc0, step32, off_e, off_p, off_l, bits)
        for nw in (1, 2, 4):
            got = np.zeros(6, np.int16)
            emu.emu_epl_cell(chips.ctypes.data, sig.ctypes.data, acc0, step32, off_e, off_p, off_l, bits, nw,
                             got.ctypes.data)
            assert np.array_equal(got, want), (off_e, off_p, off_l, bits, nw)


def _armed_pair(reference, golden, sat, prn):
    rchans = reference.channels(1)
    rch = reference.channel_at(rchans, 0)
    reference.channel_init(rch, prn, 0)
    st = reference.snapshot(rch)
    st.acq_state, st.trk_state = 9, 1
    st.found_freq_offset_hz = int(golden["track_found_freq"][sat])
    st.found_code_phase = int(golden["track_found_phase"][sat])
    reference.restore(rch, st)
    mine = Channels([prn])
    mst = mine.snapshot(0)
    mst.acq_state, mst.trk_state = 9, 1
    mst.found_freq_offset_hz, mst.found_code_phase = st.found_freq_offset_hz, st.found_code_phase
    mine.restore(0, mst)
    return rchans, rch, mine


@pytest.mark.parametrize("sat", [0, 1])
def test_emulated_device_loop_equals_reference(oracle, reference, golden, sat):
    """Pre-track on the host path, then the REST of the golden scene in one emulated device-resident run
    (loop filters with device math, raw-frame correlator, deferred SNR): I/Q sums, nav bits and the final
    channel record equal the unmodified reference run millisecond by millisecond."""
    emu, lib = load_emulator(), load_host_library()
    assert emu.emu_sizeof_channel() == lib.gpsb_host_sizeof_channel()
    sig = golden["scene_signal"]
    n_total = sig.shape[0]
    prn = (5, 14)[sat]
    chips = oracle.ca_code(prn)
    rchans, rch, mine = _armed_pair(reference, golden, sat, prn)
    lib.gpsb_host_attach(None)
    ms = 0
    while True:                                  # host path until tracking runs and a 4-ms slot starts
        st = mine.snapshot(0)
        if st.trk_state in (3, 4) and ms % 4 == 0:
            break
        reference.set_ms(ms)
        reference.lib.gps_tracking_process(rch, sig[ms].ctypes.data, ms % 4)
        host_track_ms(lib, oracle, mine.at(0), chips, sig[ms], ms, ms % 4)
        ms += 1
        assert ms < 300
    assert states_equal(mine.snapshot(0), reference.snapshot(rch))
    n_ms = n_total - ms
    want_iq, want_nav, _ = reference.track_run(rch, sig[ms:], ms, n_ms)
    aux = C.create_string_buffer(emu.emu_sizeof_aux())
    iq = np.zeros((n_ms, 6), np.int16)
    nav = np.zeros(n_ms, np.int8)
    done = C.c_uint32()
    tail = np.ascontiguousarray(sig[ms:])
    stop = emu.emu_track_run(mine.at(0), aux, tail.ctypes.data, ms, n_ms, 2, iq.ctypes.data, nav.ctypes.data,
                             C.byref(done), None)
    assert stop == 0 and done.value == n_ms
    emu.emu_resolve_snr(mine.at(0), aux)
    assert np.array_equal(iq, want_iq)
    assert np.array_equal(nav, want_nav)
    a, b = mine.snapshot(0), reference.snapshot(rch)
    assert states_equal(a, b), diff_fields(a, b)
    assert a.snr_summ_cnt < n_ms - 200                            # at least one SNR window closed (deferred log10f)
    mine.free()


def test_emulated_loop_false_lock_reseed_draws_like_libc(oracle, reference):
    """Noise-only input trips the false-lock kicker (tracking.c:300-326): the emulated device loop draws from
    its private additive-feedback generator what the reference draws from a freshly seeded libc rand()."""
    emu = load_emulator()
    rng = np.random.default_rng(77)
    n_ms = 1500
    sig = rng.integers(0, 256, (n_ms, 2046), dtype=np.uint8)
    prn = 9
    C.CDLL(None).srand(1)
    chans = reference.channels(1)
    rch = reference.channel_at(chans, 0)
    reference.channel_init(rch, prn, 0)
    st = reference.snapshot(rch)
    st.acq_state, st.trk_state, st.found_freq_offset_hz = 9, 4, 1500
    st.if_freq_offset_hz_bits = int(np.float32(1500.0).view(np.uint32))
    st.code_phase_fine_bits = int(np.float32(4000.0).view(np.uint32))
    st.pll_bad_state_cnt, st.pll_bad_state_master_cnt = 10, 80
    reference.restore(rch, st)
    mine = Channels([prn])
    from stm32f4_sdr_gps_b200 import FlatState
    mine.restore(0, FlatState.from_buffer_copy(bytes(st)))
    want_iq, want_nav, _ = reference.track_run(rch, sig, 0, n_ms)
    aux = C.create_string_buffer(emu.emu_sizeof_aux())
    iq = np.zeros((n_ms, 6), np.int16)
    nav = np.zeros(n_ms, np.int8)
    done = C.c_uint32()
    stop = emu.emu_track_run(mine.at(0), aux, sig.ctypes.data, 0, n_ms, 4, iq.ctypes.data, nav.ctypes.data,
                             C.byref(done), None)
    assert stop == 0 and done.value == n_ms
    emu.emu_resolve_snr(mine.at(0), aux)
    assert np.array_equal(iq, want_iq)
    a, b = mine.snapshot(0), reference.snapshot(rch)
    assert states_equal(a, b), diff_fields(a, b)
    mine.free()


def test_emulated_loop_hands_back_what_it_does_not_implement(oracle):
    """A channel that is not in GPS_TRACKING_RUN stops at once with nothing done (LC_STOP_STATE); an
    all-zero early+late power stops BEFORE the filters with the sums delivered (LC_STOP_DLL_NAN)."""
    emu = load_emulator()
    mine = Channels([3])
    aux = C.create_string_buffer(emu.emu_sizeof_aux())
    sig = np.zeros((4, 2046), np.uint8)
    done = C.c_uint32(99)
    before = bytes(mine.snapshot(0))
    assert emu.emu_track_run(mine.at(0), aux, sig.ctypes.data, 0, 4, 2, None, None, C.byref(done), None) == 1
    assert done.value == 0 and bytes(mine.snapshot(0)) == before
    mine.free()
