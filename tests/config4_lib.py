"""BASELINE configs[3] (SURVEY.md section 8(d) "Config 4"): 32 PRNs searched, some present; cold sweep -> code-phase
rounds -> pre-track -> seconds of closed-loop tracking with nav bits, and the SAME done by the unmodified reference,
satellite by satellite, for a field-by-field diff.  TEST INFRASTRUCTURE (imports the checker): used by
tests/test_config4.py and by bench.py's config4 leg.

* product():   this library as a user runs it: gpsb_rx_cold_start (one sweep launch, look-ahead windows for the code
               rounds), gpsb_rx_track_run with the slot-phase walk on (pre-track on the per-ms path, then k_track_run).
* reference(): per satellite, the reference's own entry points on the same snapshots (the schedule gpsb_rx_cold_start
               reports): acquisition_start_channel + acquisition_process_channel for the 10 cells of each Doppler bin,
               acquisition_start_code_search_channel, acquisition_process_channel ..., acquisition_start_code_search3_
               channel, ..., then gps_tracking_process on the walked (millisecond, slot index) schedule
               (oracle/ref_shim.c, ref_track_run_walk).
* diff():      found-PRN set, Doppler, code phase, per-ms I/Q of all three arms, nav bits, subframe counters and the raw
               channel record of every searched satellite."""
import ctypes as C

import numpy as np

from stm32f4_sdr_gps_b200.signal_synth import Satellite, Scene, synthesize_blocks

MS_SAMPLES = 16368
SEARCHED = list(range(1, 33))
SWEEP_MS = 10
BINS = 29                    # the reference's own Doppler grid: -7000 .. +7000 Hz in steps of 500 (config.h:41-44)


def scene(n_ms, n_present=10, seed=0x5D120004):
    """32 PRNs searched, n_present in the sky: Doppler within +-5 kHz, 47 dB-Hz, random data bits and bit-edge timing."""
    rng = np.random.default_rng(seed ^ 0x4C4C)
    prns = sorted(rng.choice(np.arange(1, 33), size=n_present, replace=False).tolist())
    sats = [Satellite(prn=int(p), doppler_hz=float(rng.uniform(-4800, 4800)), code_phase_samples=float(rng.uniform(0, MS_SAMPLES)),
                      carrier_phase_rad=float(rng.uniform(0, 2 * np.pi)), cn0_dbhz=47.0,
                      nav_bit_offset_ms=int(rng.integers(0, 20))) for p in prns]
    return Scene(sats=sats, n_ms=n_ms, seed=seed)


def signal(sc):
    return synthesize_blocks(sc.sats, sc.n_ms, sc.seed)


def product(engine, sig, prns, n_track_ms, pre_ms=100, walk=True, timers=None, sweeps=3, cold_start_opts=None,
            before_start=None):
    """Returns (channels, receiver, report, per-channel logs).  The signal must already be in the engine's ring from
    frame 0 (gpsb_upload_signal).  timers: optional dict that receives wall-clock seconds per stage."""
    import time
    from stm32f4_sdr_gps_b200 import Channels, Receiver
    ch = Channels(prns)
    rx = Receiver(engine, ch)
    if walk:
        rx.set_slot_walk(True)
    if before_start:
        before_start()                                          # e.g. a barrier over the ranks
    t0 = time.perf_counter()
    rep = rx.cold_start(0, sweeps=sweeps, **(cold_start_opts or {}))
    t1 = time.perf_counter()
    t_trk = rep["ms_next"]
    # the first call covers pre-track (>= 84 ms, tracking.c:398-450; k_pretrack_run + k_track_run in one round trip),
    # the second everything after: one k_track_run launch for everybody
    iq_a, nav_a = rx.track_run(t_trk, pre_ms)
    t2 = time.perf_counter()
    iq_b, nav_b = rx.track_run(t_trk + pre_ms, n_track_ms - pre_ms)
    t3 = time.perf_counter()
    if timers is not None:
        timers.update(cold_start_s=t1 - t0, pre_track_s=t2 - t1, tracking_s=t3 - t2, total_s=t3 - t0)
    return ch, rx, rep, (np.concatenate([iq_a, iq_b]), np.concatenate([nav_a, nav_b]))


def reference(ref, sig, prn, rep, n_track_ms, walk=True, served=True):
    """One satellite through the unmodified reference on the schedule `rep` (a gpsb_cold_start_report as a dict).
    served=False: a satellite another rank goes on with - this rank stops serving it after the sweeps.
    Returns (final channel record bytes, iq, nav, state after acquisition)."""
    from oracle_lib import RefWalk
    lib = ref.lib
    # The reference keeps the best pre-track correlation of the slot in progress in two file-scope variables shared by
    # all channels (tracking.c:33-34) and never resets the phase: a satellite run in a process that ran another one before
    # would inherit them.  Every satellite starts from a clean process image, like a channel of this library does.
    for name in ("pre_track_best_corr_value", "pre_track_best_corr_phase"):
        C.c_uint16.in_dll(lib, name).value = 0
    # ... and its false-lock kicker draws from the process-wide rand() (tracking.c:316), where a channel of this library
    # has its own generator, seeded like a fresh process (core/gpsb_loop_core.h, lc_rand31_*)
    C.CDLL(None).srand(1)
    chans = ref.channels(1)
    ch = ref.channel_at(chans, 0)
    ref.channel_init(ch, prn, 0)
    ref.set_ms(rep["ms_sweep0"])
    lib.acquisition_start_channel(ch)
    for s in range(rep["n_sweeps"]):                        # every sweep on its own 10 snapshots
        first = rep["ms_sweep0"] + s * SWEEP_MS
        for b in range(BINS):                               # 10 cells per bin (acquisition.c:280-312)
            for m in range(SWEEP_MS):
                if ref.snapshot(ch).acq_state != 1:         # GPS_ACQ_FREQ_SEARCH_RUN: decided at an earlier bin
                    break
                ref.set_ms(first + SWEEP_MS - 1)
                lib.acquisition_process_channel(ch, sig[first + m].ctypes.data)
    if served and ref.snapshot(ch).acq_state == 2:          # GPS_ACQ_FREQ_SEARCH_DONE: the others are not served any more
        ref.set_ms(rep["ms_code0"])
        lib.acquisition_start_code_search_channel(ch)
        for ms in range(rep["ms_code0"], rep["ms_code12_last"] + 1):
            ref.set_ms(ms)
            lib.acquisition_process_channel(ch, sig[ms].ctypes.data)
        ref.set_ms(rep["ms_code3_first"])
        lib.acquisition_start_code_search3_channel(ch)
        for ms in range(rep["ms_code3_first"], rep["ms_last"] + 1):
            ref.set_ms(ms)
            lib.acquisition_process_channel(ch, sig[ms].ctypes.data)
    st = ref.snapshot(ch)
    acquired = st.acq_state == 9                            # GPS_ACQ_DONE
    if acquired and st.trk_state == 0:                      # gps_master.c:121-129
        st.trk_state = 1
        ref.restore(ch, st)
    after_acq = bytes(ref.snapshot(ch))
    iq = np.zeros((n_track_ms, 6), np.int16)
    nav = np.full(n_track_ms, -1, np.int8)
    if acquired:
        w = RefWalk()
        w.enable = 1 if walk else 0
        t = rep["ms_next"]
        iq, nav, _ = ref.track_run_walk(ch, sig[t:t + n_track_ms], t, n_track_ms, w)
    return bytes(ref.snapshot(ch)), iq, nav, after_acq


def diff(ch, logs, refs, prns):
    """Field-by-field comparison; returns a summary dict, raises AssertionError on the first difference."""
    iq, nav = logs
    found, summary = [], {"searched": len(prns), "acquired": [], "cells": 0, "nav_bits": 0, "subframe_words": 0}
    for i, prn in enumerate(prns):
        rec, r_iq, r_nav, _ = refs[i]
        st = ch.snapshot(i)
        if bytes(st) != rec:
            theirs = type(st).from_buffer_copy(rec)
            plain = lambda v: bytes(v) if hasattr(v, "__len__") else v
            raise AssertionError(("channel record", prn, [(n, plain(getattr(st, n)), plain(getattr(theirs, n)))
                                                          for n, _ in st._fields_
                                                          if plain(getattr(st, n)) != plain(getattr(theirs, n))][:12]))
        if st.acq_state != 9:
            assert not iq[:, i, :].any() and (nav[:, i] == -1).all(), ("a satellite that was not acquired was tracked", prn)
            continue
        found.append(prn)
        assert np.array_equal(iq[:, i, :], r_iq), ("per-ms sums", prn)
        assert np.array_equal(nav[:, i], r_nav), ("nav bits", prn)
        summary["cells"] += int(iq.shape[0])
        summary["nav_bits"] += int((r_nav >= 0).sum())
        summary["subframe_words"] += int(st.word_cnt_test)
        summary["acquired"].append({"prn": prn, "doppler_hz": int(st.found_freq_offset_hz), "code_phase": int(st.found_code_phase),
                                    "carrier_hz": float(np.uint32(st.if_freq_offset_hz_bits).view(np.float32)),
                                    "bit_edge_refined": int(st.accurate_swap_ok), "words_ok": int(st.word_cnt_test)})
    summary["found_prns"] = found
    return summary


_shared = {}


def _ref_job(i):
    from oracle_lib import Reference
    ref = _shared.get("ref")
    if ref is None:
        ref = _shared["ref"] = Reference()
    return reference(ref, _shared["sig"], _shared["prns"][i], _shared["rep"], _shared["n_track_ms"], _shared["walk"],
                     _shared["served"][i])


def reference_all(sig, prns, rep, n_track_ms, walk=True, procs=None, served=None):
    """reference() for every searched satellite, one process each on up to `procs` cores (the reference's file-scope
    scratch serves one channel at a time).  served[i] False: satellite i belongs to another rank after the sweeps."""
    import multiprocessing as mp
    import os
    _shared.update(sig=sig, prns=list(prns), rep=rep, n_track_ms=n_track_ms, walk=walk,
                   served=list(served) if served is not None else [True] * len(prns))
    _shared.pop("ref", None)
    procs = max(1, min(procs or (os.cpu_count() or 1), len(prns)))
    if procs == 1:
        return [_ref_job(i) for i in range(len(prns))]
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_ref_job, range(len(prns)), chunksize=1)
