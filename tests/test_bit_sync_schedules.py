"""DESIGN.md section 10, pinned: the reference's bit synchroniser sees a data-bit edge only inside a 4-ms slot
(nav_data.c:87-138).  On the MCU's 17-ms channel schedule (main.c:134-155) the slots walk over every edge alignment;
with every millisecond processed in place (index = ms % 4, what ref_track_run, gpsb_rx_track_ms and k_track_run do)
the alignment is fixed.  One satellite whose edges fall on a slot boundary:

* every millisecond: the reference never synchronises in 6 s;
* the 17-ms schedule: the reference synchronises - and this library's host state machine, driven through the same
  calls with the oracle doing the correlations, is in the same state after every call."""
import ctypes as C

import numpy as np

from stm32f4_sdr_gps_b200 import Channels, load_host_library
from stm32f4_sdr_gps_b200.signal_synth import Satellite, Scene, synthesize
from test_host_logic import diff_fields, host_track_ms, states_equal


def test_edge_on_a_slot_boundary_needs_the_walking_schedule(oracle, reference):
    lib = load_host_library()
    lib.gpsb_host_attach(None)
    prn, doppler, code_phase, n_ms = 12, 1500.0, 5000.3, 6000
    rng = np.random.default_rng(4)
    bits = np.tile([0, 1], n_ms // 40 + 2).astype(np.uint8)      # an edge every 20 ms: the easiest case there is
    bits[rng.integers(0, bits.size, bits.size // 8)] ^= 1
    # bit edges at code epochs 100, 120, ...: they arrive in receiver milliseconds 100.3, 120.3, ... - the sign flips
    # between ms 99 and ms 100, the last and the first millisecond of two slots when index = ms % 4
    sat = Satellite(prn=prn, doppler_hz=doppler, code_phase_samples=code_phase, cn0_dbhz=50.0, nav_bits=bits,
                    nav_bit_offset_ms=100)
    sig = synthesize(Scene(sats=[sat], n_ms=n_ms, seed=12))

    def locked(st):
        st.acq_state, st.trk_state, st.found_freq_offset_hz = 9, 4, int(doppler)
        st.if_freq_offset_hz_bits = int(np.float32(doppler).view(np.uint32))
        st.code_phase_fine_bits = int(np.float32(code_phase).view(np.uint32))
        return st

    # every millisecond in place
    rchans = reference.channels(1)
    rch = reference.channel_at(rchans, 0)
    reference.channel_init(rch, prn, 0)
    reference.restore(rch, locked(reference.snapshot(rch)))
    iq, nav, _ = reference.track_run(rch, sig, 0, n_ms)
    assert np.abs(iq[1000:, 2]).mean() > 400                      # it tracks the satellite all right
    assert reference.snapshot(rch).period_sync_ok_flag == 0 and (nav >= 0).sum() == 0     # but never finds the bit edges

    # the MCU's schedule: this channel owns milliseconds 0..3 of every 17
    rchans = reference.channels(1)
    rch = reference.channel_at(rchans, 0)
    reference.channel_init(rch, prn, 0)
    reference.restore(rch, locked(reference.snapshot(rch)))
    mine = Channels([prn])
    mine.restore(0, locked(mine.snapshot(0)))
    chips = oracle.ca_code(prn)
    synced_at = None
    for ms in range(n_ms):
        slot = ms % 17
        if slot >= 4:
            continue
        reference.set_ms(ms)
        reference.lib.gps_tracking_process(rch, sig[ms].ctypes.data, slot)
        host_track_ms(lib, oracle, mine.at(0), chips, sig[ms], ms, slot)
        a, b = mine.snapshot(0), reference.snapshot(rch)
        assert states_equal(a, b), (ms, diff_fields(a, b))
        if synced_at is None and a.period_sync_ok_flag:
            synced_at = ms
    assert synced_at is not None and synced_at < 5000, synced_at
    fin = mine.snapshot(0)
    assert fin.period_sync_ok_flag == 1 and fin.trk_state == 4
    mine.free()


def test_slot_phase_moves_the_window_onto_the_edge(reference):
    """What DESIGN.md section 10 proposes for the batched paths, pinned on the CPU: with the slots of the same
    satellite started one millisecond later (index = (ms + 1) % 4), every millisecond still processed, the device
    loop's sources synchronise - and equal the reference driven with the same index sequence."""
    from emu_lib import load_emulator
    emu = load_emulator()
    prn, doppler, code_phase, n_ms = 12, 1500.0, 5000.3, 3000
    rng = np.random.default_rng(4)
    bits = np.tile([0, 1], n_ms // 40 + 2).astype(np.uint8)
    bits[rng.integers(0, bits.size, bits.size // 8)] ^= 1
    sat = Satellite(prn=prn, doppler_hz=doppler, code_phase_samples=code_phase, cn0_dbhz=50.0, nav_bits=bits,
                    nav_bit_offset_ms=100)
    sig = synthesize(Scene(sats=[sat], n_ms=n_ms, seed=12))

    def locked(st):
        st.acq_state, st.trk_state, st.found_freq_offset_hz = 9, 4, int(doppler)
        st.if_freq_offset_hz_bits = int(np.float32(doppler).view(np.uint32))
        st.code_phase_fine_bits = int(np.float32(code_phase).view(np.uint32))
        return st

    for phase, expect_sync in ((0, False), (1, True), (2, True)):
        rchans = reference.channels(1)
        rch = reference.channel_at(rchans, 0)
        reference.channel_init(rch, prn, 0)
        reference.restore(rch, locked(reference.snapshot(rch)))
        for ms in range(n_ms):
            reference.set_ms(ms)
            reference.lib.gps_tracking_process(rch, sig[ms].ctypes.data, (ms + phase) % 4)
        mine = Channels([prn])
        mine.restore(0, locked(mine.snapshot(0)))
        aux = C.create_string_buffer(emu.emu_sizeof_aux())
        done = C.c_uint32()
        stop = emu.emu_track_run_phase(mine.at(0), aux, sig.ctypes.data, 0, n_ms, 2, phase, None, None, C.byref(done), None)
        assert stop == 0 and done.value == n_ms
        emu.emu_resolve_snr(mine.at(0), aux)
        a, b = mine.snapshot(0), reference.snapshot(rch)
        assert states_equal(a, b), (phase, diff_fields(a, b))
        assert bool(a.period_sync_ok_flag) == expect_sync, phase
        if phase == 2:
            assert a.accurate_swap_ok == 1                        # the edge shows at slot position 2: it is refined too
        mine.free()
