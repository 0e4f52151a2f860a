"""The slot-phase walk of the batched paths (core/gpsb_loop_core.h lc_walk_*, include/gpsb_host.h gpsb_rx_set_slot_walk).

The reference's bit synchroniser only sees a data-bit edge inside a 4-ms channel slot and only refines an edge at slot
position 2 (nav_data.c:87-138); with every millisecond processed at index = ms % 4 the alignment never moves, so three
satellites in four never deliver a subframe time stamp.  With the walk enabled a channel leaves 1..3 milliseconds out
between two slots until its edges show at position 2.  Checked here on the CPU with the sources of k_track_run
(tests/emu/loop_emu.c) against the UNMODIFIED reference driven on the walked (millisecond, slot index) schedule by the
checker's own restatement of the policy (oracle/ref_shim.c, ref_track_run_walk): four satellites whose bit edges sit at
all four alignments all end with a refined edge, and channel records, sums, nav bits and the schedule itself are equal.
The GPU leg (tests/test_gpu_loop.py) runs the kernel over the same recording."""
import ctypes as C

import numpy as np
import pytest

from stm32f4_sdr_gps_b200 import Channels, load_host_library
from stm32f4_sdr_gps_b200.signal_synth import Satellite, Scene, synthesize
from test_host_logic import diff_fields, states_equal

PRNS = (3, 12, 19, 27)
N_MS = 3600


def walk_scene(n_ms=N_MS, seed=31):
    """Four satellites, data-bit edges in receiver milliseconds 100, 101, 102, 103 (mod 20): every slot alignment."""
    rng = np.random.default_rng(seed)
    sats = []
    for k, prn in enumerate(PRNS):
        bits = rng.integers(0, 2, n_ms // 20 + 3).astype(np.uint8)
        sats.append(Satellite(prn=prn, doppler_hz=float(rng.uniform(-3000, 3000)),
                              code_phase_samples=float(rng.uniform(1500, 6000)), cn0_dbhz=50.0,
                              carrier_phase_rad=float(rng.uniform(0, 6.28)), nav_bits=bits, nav_bit_offset_ms=100 + k))
    scene = Scene(sats=sats, n_ms=n_ms, seed=seed)
    return scene, synthesize(scene)


def locked(st, sat):
    st.acq_state, st.trk_state = 9, 4
    st.found_freq_offset_hz = int(round(sat.doppler_hz / 500.0) * 500)
    st.if_freq_offset_hz_bits = int(np.float32(sat.doppler_hz).view(np.uint32))
    st.code_phase_fine_bits = int(np.float32(sat.code_phase_samples).view(np.uint32))
    return st


def reference_walk(reference, sat, sig, n_ms, enable=1, period_ms=0, cuts=()):
    from oracle_lib import RefWalk
    rchans = reference.channels(1)
    rch = reference.channel_at(rchans, 0)
    reference.channel_init(rch, sat.prn, 0)
    reference.restore(rch, locked(reference.snapshot(rch), sat))
    walk = RefWalk()
    walk.enable, walk.period_ms = enable, period_ms
    parts, at = [], 0
    for end in list(cuts) + [n_ms]:
        parts.append(reference.track_run_walk(rch, sig[at:end], at, end - at, walk))
        at = end
    iq, nav, idx = (np.concatenate([p[k] for p in parts]) for k in range(3))
    return reference.snapshot(rch), iq, nav, idx, walk


@pytest.fixture(scope="module")
def scene_and_signal():
    return walk_scene()


def test_walk_brings_every_alignment_to_a_refined_edge(reference, scene_and_signal):
    from emu_lib import load_emulator
    emu = load_emulator()
    host = load_host_library()
    scene, sig = scene_and_signal
    gaps = []
    for sat in scene.sats:
        want, iq_ref, nav_ref, idx_ref, walk = reference_walk(reference, sat, sig, N_MS)
        mine = Channels([sat.prn])
        mine.restore(0, locked(mine.snapshot(0), sat))
        aux = C.create_string_buffer(emu.emu_sizeof_aux())
        host.gpsb_host_aux_walk(aux, 1, 0)
        iq = np.zeros((N_MS, 6), np.int16)
        nav = np.zeros(N_MS, np.int8)
        done = C.c_uint32()
        # three launches, cut inside a slot, right before and inside where gaps may fall: the walk state travels in aux
        at = 0
        for end in (1203, 1810, N_MS):
            stop = emu.emu_track_run(mine.at(0), aux, sig[at:end].ctypes.data, at, end - at, 2, iq[at:].ctypes.data,
                                     nav[at:].ctypes.data, C.byref(done), None)
            assert stop == 0 and done.value == end - at
            at = end
        emu.emu_resolve_snr(mine.at(0), aux)
        got = mine.snapshot(0)
        assert states_equal(got, want), (sat.prn, diff_fields(got, want))
        assert np.array_equal(iq, iq_ref) and np.array_equal(nav, nav_ref), sat.prn
        idle = ~iq.any(axis=1)
        assert np.array_equal(idle, idx_ref == 0xFF), sat.prn           # the schedule itself
        state = (C.c_uint32 * 6)()
        host.gpsb_host_aux_walk_state(aux, C.byref(state))
        assert state[0] == walk.slot_phase and state[4] == walk.gaps_taken
        assert got.period_sync_ok_flag == 1 and got.accurate_swap_ok == 1, (sat.prn, "no refined bit edge")
        gaps.append((int(idle.sum()), int(state[4])))
        mine.free()
    # bit edges at slot positions 0 (never seen), 1, 2 (in place) and 3: idle 2, 3, 0 and 1 ms
    assert sorted(g[0] for g in gaps) == [0, 1, 2, 3], gaps


def test_walk_off_is_the_fixed_schedule(reference, scene_and_signal):
    """Default (walk off): index = ms % 4 for ever - the channel whose edges sit on a slot boundary never synchronises,
    this library and the reference alike (the behaviour tests/test_bit_sync_schedules.py pins)."""
    from emu_lib import load_emulator
    emu = load_emulator()
    scene, sig = scene_and_signal
    n_ms = 2400
    synced = []
    for sat in scene.sats:
        want, iq_ref, nav_ref, idx_ref, walk = reference_walk(reference, sat, sig, n_ms, enable=0)
        assert walk.gaps_taken == 0 and (idx_ref == np.arange(n_ms) % 4).all()
        mine = Channels([sat.prn])
        mine.restore(0, locked(mine.snapshot(0), sat))
        aux = C.create_string_buffer(emu.emu_sizeof_aux())
        done = C.c_uint32()
        assert emu.emu_track_run(mine.at(0), aux, sig.ctypes.data, 0, n_ms, 2, None, None, C.byref(done), None) == 0
        emu.emu_resolve_snr(mine.at(0), aux)
        got = mine.snapshot(0)
        assert states_equal(got, want), (sat.prn, diff_fields(got, want))
        synced.append((got.period_sync_ok_flag, got.accurate_swap_ok))
        mine.free()
    assert sorted(synced) == [(0, 0), (1, 0), (1, 0), (1, 1)], synced
