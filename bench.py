#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 correlator engine.

Workload (BASELINE.json configs[1]): 4-satellite E/P/L tracking over 1 s of 16.368 Msps 1-bit IF
(1000 ms x 4 SV = 4000 integrate-and-dump cells, 3 arms each).  One "step" = one pass over that second.
With N GPUs every rank tracks its own 4 satellites over the same second (satellites shard, no data-path
collective): weak scaling, value = arm-samples of all ranks / max-over-ranks time.

  value        arm-samples/s of the closed loop with signal AND channel records already resident in HBM: one
               k_track_run launch per step (loop filters on the device), device-timed (CUDA events)
  e2e          same metric through the reference-facing host library (gpsb_rx_track_stream) with HOST buffers:
               the pinned signal is DMA-ed into the HBM ring in chunks while the loop launch is already tracking,
               channel records H2D, records + per-ms sums + nav bits D2H, all inside the timed region
  roofline     k_track_run against the measured HBM peak (algorithmic bytes)
  cpu_baseline the unmodified reference C (oracle/_ref) on this box's host cores, same cells
  streaming    config 5 shape: 32 channels from a host-resident stream through a ring shorter than the run
  config1_batched  400 000 cells, one per ms of an 818-MB recording, in one k_epl_batch launch (prompt arm / three
               arms), each against the HBM roofline, the reference C on a bounded sample beside it
  batch_replay the closed loop's own 4000 cells replayed open loop in one launch
  cold_acq     secondary metric: 32 SV x 21 bins x 10 ms x 2046 phases full-sky sweep (configs[2]), plus the
               16368-phase grid (8 sub-byte shifts)

`--impl reference` times the reference's own CPU path (oracle/_ref, else the oracle port) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

IF_HZ = 4092000
MS_SAMPLES = 16368
N_SV_PER_GPU = 4
N_MS = int(os.environ.get("GPSB_BENCH_NMS", "1000"))   # shrink only for profiler runs (never for a reported number)
ARMS = 3
ALL_PRNS = [5, 14, 20, 30, 1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 17, 18, 19, 21, 22, 23, 24, 25, 26, 27, 28,
            29, 31, 32]
ACQ_SV, ACQ_BINS, ACQ_MS = 32, 21, 10


# ----------------------------------------------------------------------------- workload
def make_scene(rank: int, n_ms: int):
    from stm32f4_sdr_gps_b200.signal_synth import config2_scene
    prns = ALL_PRNS[rank * N_SV_PER_GPU:(rank + 1) * N_SV_PER_GPU]
    return config2_scene(n_ms=n_ms, prns=prns, seed=0x5D120001 + rank)


def cached_signal(tag: str, scene):
    """Synthesis is pure numpy and takes seconds; cache per tag under the system temp dir."""
    from stm32f4_sdr_gps_b200.signal_synth import synthesize
    path = Path(tempfile.gettempdir()) / ("gpsb_bench_%s.npy" % tag)
    if path.exists():
        sig = np.load(path)
        if sig.shape == (scene.n_ms, 2046):
            for s in scene.sats:       # truth is filled by synthesize(); recreate the parts we need
                scene.truth[s.prn] = {}
            return sig
    sig = synthesize(scene)
    try:
        np.save(path, sig)
    except OSError:
        pass
    return sig


def truth_requests(scene, nco_step32):
    """Per-(ms, sv) cell parameters along the true trajectory of each satellite: what a locked tracking
    loop feeds the correlator (tracking.c:115-130 offset arithmetic on the true code phase)."""
    from stm32f4_sdr_gps_b200 import EPL_REQ
    n_sv = len(scene.sats)
    rq = np.zeros((scene.n_ms, n_sv), EPL_REQ)
    fo = np.zeros((scene.n_ms, n_sv), np.float32)
    fine_f = np.zeros((scene.n_ms, n_sv), np.float32)
    for s, sat in enumerate(scene.sats):
        step32 = nco_step32(np.float32(IF_HZ) + np.float32(sat.doppler_hz))
        m = np.arange(scene.n_ms, dtype=np.float64)
        tau = (sat.code_phase_samples - m * MS_SAMPLES * sat.doppler_hz / 1_575_420_000.0) % MS_SAMPLES
        fine = np.floor(tau).astype(np.int64)
        p = fine // 8
        rq["sv_slot"][:, s] = sat.prn             # satellite slot == PRN, as the host-side mirror uses them
        rq["ms_index"][:, s] = np.arange(scene.n_ms)
        rq["acc0"][:, s] = (np.arange(scene.n_ms, dtype=np.uint64) * 511 * step32) & 0xFFFFFFFF
        rq["step32"][:, s] = step32
        rq["off_p"][:, s] = p
        rq["off_e"][:, s] = np.where(p == 0, 2045, p - 1)
        rq["off_l"][:, s] = np.where(p + 1 >= 2046, 0, p + 1)
        rq["off_bits"][:, s] = fine & 7
        fo[:, s] = np.float32(sat.doppler_hz)
        fine_f[:, s] = tau.astype(np.float32)
    return rq.reshape(-1), fo.reshape(-1), fine_f.reshape(-1)


# ----------------------------------------------------------------------------- ncu evidence
def profile_traffic(name: str):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the ncu --set full capture summarised in
    profiles/<name> (tools/ncu_summary.py), or None when the file or the metrics are missing."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, 0
    try:
        for line in (REPO / "profiles" / name).read_text().splitlines():
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[2] in scale:
                total += float(f[1].replace(",", "")) * scale[f[2]]
                seen += 1
    except OSError:
        return None
    return int(total) if seen == 2 else None


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = Path(tempfile.gettempdir()) / ("gpsb_clocks_%d_%d.csv" % (os.getpid(), gpu_index))
        self.proc = None
        def die_with_parent():          # never leave a sampler behind if the bench dies (PR_SET_PDEATHSIG, SIGTERM)
            import ctypes
            ctypes.CDLL("libc.so.6").prctl(1, 15)
        try:
            self.proc = subprocess.Popen(
                ["timeout", "900", "nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=open(self.path, "w"),
                stderr=subprocess.DEVNULL, preexec_fn=die_with_parent)
            import atexit
            atexit.register(self._kill)
        except OSError:
            self.proc = None

    def _kill(self) -> None:
        if self.proc is not None and self.proc.poll() is None:
            self.proc.kill()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.path.read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            self.path.unlink()
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm (CPU)
def _ref_worker(args):
    """Closed-loop reference tracking of one satellite in its own process (reference globals are not
    re-entrant).  Returns (best seconds over `reps` timed runs, logs or None): with want_logs a
    further, untimed run returns the reference's own per-ms sums, nav bits and final channel record."""
    prn, fo_hz, fine, sig_path, n_ms, reps, so_path, want_logs = args
    import ctypes as C
    sys.path.insert(0, str(REPO / "tests"))
    from oracle_lib import Reference
    ref = Reference(so_path)
    sig = np.load(sig_path, mmap_mode="r")
    sig = np.ascontiguousarray(sig[:n_ms])
    chans = ref.channels(1)
    ch = ref.channel_at(chans, 0)
    lib = ref.lib
    lib.ref_track_time.restype = C.c_double
    lib.ref_track_time.argtypes = [C.c_void_p] * 2 + [C.c_uint32] * 2

    def arm():
        C.CDLL(None).srand(1)      # the reference's false-lock kicker draws from the process-wide rand(): a fresh process' state
        ref.channel_init(ch, prn, 0)
        st = ref.snapshot(ch)
        # start locked (GPS_ACQ_DONE / GPS_TRACKING_RUN) on the true code phase and Doppler so that all n_ms
        # steps are E/P/L integrate-and-dump steps, the same cells the GPU arm computes
        st.acq_state, st.trk_state = 9, 4
        st.found_freq_offset_hz = int(round(fo_hz / 500.0) * 500)
        st.if_freq_offset_hz_bits = int(np.float32(fo_hz).view(np.uint32))
        st.code_phase_fine_bits = int(np.float32(fine).view(np.uint32))
        ref.restore(ch, st)

    best = 1e30
    for _ in range(reps):
        arm()
        best = min(best, lib.ref_track_time(ch, sig.ctypes.data, 0, n_ms))
    logs = None
    if want_logs:
        arm()
        iq, nav, _ = ref.track_run(ch, sig, 0, n_ms)
        logs = (iq, nav, bytes(ref.snapshot(ch)))
    return best, logs


def reference_tracking(scenes_and_sigs, max_procs: int, reps: int = 3, so_path=None, want_logs: bool = False):
    """The reference C tracking every satellite of every scene over its whole recording, one process per satellite on
    up to max_procs cores AT THE SAME TIME (a channel's 1-kHz loop is serial, so one satellite cannot use more than one
    core).  Returns (seconds for the whole job, processes used, logs per satellite or None)."""
    import multiprocessing as mp
    sys.path.insert(0, str(REPO / "tests"))
    from oracle_lib import have_reference
    if not have_reference():
        return None, 0, None
    jobs, paths = [], []
    for k, (scene, sig) in enumerate(scenes_and_sigs):
        path = Path(tempfile.gettempdir()) / ("gpsb_ref_sig_%d_%d.npy" % (os.getpid(), k))
        np.save(path, sig)
        paths.append(path)
        jobs += [(s.prn, float(s.doppler_hz), float(s.code_phase_samples), str(path), scene.n_ms, reps,
                  str(so_path) if so_path else None, want_logs) for s in scene.sats]
    procs = max(1, min(max_procs, len(jobs)))
    if procs == 1:
        out = [_ref_worker(j) for j in jobs]
        wall = sum(o[0] for o in out)
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            out = pool.map(_ref_worker, jobs, chunksize=1)
        # satellites run concurrently: the job takes as long as the busiest core's share
        per = [o[0] for o in out]
        wall = max(sum(per[i::procs]) for i in range(procs))
    for path in paths:
        path.unlink(missing_ok=True)
    return wall, procs, ([o[1] for o in out] if want_logs else None)


def reference_tracking_seconds(scene, sig, max_procs: int, reps: int = 3):
    wall, procs, _ = reference_tracking([(scene, sig)], max_procs, reps)
    return wall, procs, "reference" if wall is not None else "port"


def reference_sweep_rate(acq_sig, n_sv: int = 2, n_ms: int = 2):
    """Reference C (oracle/_ref) on a bounded sample of the cold-acquisition cells: n_sv x 21 bins x n_ms full
    2046-phase searches on one core (PRN 1, 2; first n_ms milliseconds).  Returns (cells_per_second, n_cells, the
    reference's (max, phase, avg) per cell) or (None, 0, None)."""
    import ctypes as C
    sys.path.insert(0, str(REPO / "tests"))
    from oracle_lib import Reference, have_reference
    if not have_reference():
        return None, 0, None
    ref = Reference()
    lib = ref.lib
    lib.ref_sweep_time.restype = C.c_double
    lib.ref_sweep_time.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32,
                                   C.c_uint32, C.c_void_p]
    chans = ref.channels(n_sv)
    for i in range(n_sv):
        ref.channel_init(ref.channel_at(chans, i), i + 1, 0)
    out = np.zeros((n_sv, ACQ_BINS, n_ms, 3), np.uint16)
    sig = np.ascontiguousarray(acq_sig[:n_ms])
    best = min(lib.ref_sweep_time(chans, n_sv, sig.ctypes.data, n_ms, -5000, 500, ACQ_BINS, 0, out.ctypes.data)
               for _ in range(2))
    return n_sv * ACQ_BINS * n_ms / best, n_sv * ACQ_BINS * n_ms, out


def reference_prompt_rate(sig, n_ms: int = 8000):
    """Reference C (oracle/_ref), config 1 repeated over a recording: replica + stateless mixer + one
    gps_correlation_iq per millisecond on one core.  Returns (ms per second, I/Q of the sample) or (None, None)."""
    import ctypes as C
    sys.path.insert(0, str(REPO / "tests"))
    from oracle_lib import Reference, have_reference
    if not have_reference():
        return None, None
    ref = Reference()
    lib = ref.lib
    if not hasattr(lib, "ref_prompt_time"):
        return None, None
    lib.ref_prompt_time.restype = C.c_double
    lib.ref_prompt_time.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
    chans = ref.channels(1)
    ch = ref.channel_at(chans, 0)
    ref.channel_init(ch, 1, 0)
    n_ms = min(n_ms, sig.shape[0])
    data = np.ascontiguousarray(sig[:n_ms])
    out = np.zeros((n_ms, 2), np.int16)
    best = min(lib.ref_prompt_time(ch, data.ctypes.data, n_ms, float(np.float32(IF_HZ + 2000)), 100, 0, out.ctypes.data)
               for _ in range(3))
    return n_ms / best, out


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_gpus = args.gpus
    scenes = [make_scene(r, N_MS) for r in range(n_gpus)]
    pairs = [(sc, cached_signal("trk_r%d_%d" % (r, N_MS), sc)) for r, sc in enumerate(scenes)]
    cores = os.cpu_count() or 1
    times = []
    used = 1
    for w in range(args.warmup + args.steps):
        # every satellite of every rank's scene at once, one process each, on as many cores as the box has
        t, used, _ = reference_tracking(pairs, cores, reps=3)
        if t is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref/libgpsref.so missing"})
            return
        if w >= args.warmup:
            times.append(t)
    t = float(np.mean(times))
    units = n_gpus * N_SV_PER_GPU * ARMS * MS_SAMPLES * N_MS
    v = units / t
    line = {
        "impl": "reference", "metric": "correlator-samples/sec (E/P/L arms)", "value": v, "unit": "arm-samples/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 xor/popcount",
        "data": "synthetic",
        "config": {"workload": "config2: %d-SV E/P/L closed-loop tracking, 1 s @16.368 Msps 1-bit IF" % (n_gpus * N_SV_PER_GPU),
                   "n_sv": n_gpus * N_SV_PER_GPU, "n_ms": N_MS},
        "cpu_baseline": {"value": v, "unit": "arm-samples/s", "cores": used, "host_cores": cores, "kind": "reference",
                         "sample": "whole workload: unmodified reference gps_tracking_process() closed loop, one process per "
                                   "SV, all %d SV at the same time on %d of the box's %d cores (a channel's loop is "
                                   "serial: one SV cannot use more than one core)" % (n_gpus * N_SV_PER_GPU, used, cores)},
        "e2e": {"value": v, "unit": "arm-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------- own arm (GPU)
def arm_locked(channels, scene):
    """Channels start locked on the true Doppler / code phase (GPS_ACQ_DONE, GPS_TRACKING_RUN) so that every
    millisecond is an E/P/L integrate-and-dump step - the same start the reference arm gets."""
    for i, sat in enumerate(scene.sats):
        st = channels.snapshot(i)
        for name, _ in st._fields_:
            if name not in ("prn",):
                v = getattr(st, name)
                if not hasattr(v, "__len__"):
                    setattr(st, name, 0)
        st.acq_state, st.trk_state = 9, 4
        st.found_freq_offset_hz = int(round(sat.doppler_hz / 500.0) * 500)
        st.if_freq_offset_hz_bits = int(np.float32(sat.doppler_hz).view(np.uint32))
        st.code_phase_fine_bits = int(np.float32(sat.code_phase_samples).view(np.uint32))
        channels.restore(i, st)


# profiler runs only (never for a reported number): a kernel replayed by ncu cannot be fed by a concurrent copy stream
NO_STREAM = os.environ.get("GPSB_BENCH_NO_STREAM") is not None
LOOP_FIXED_SLOTS = 2          # include/gpsb.h GPSB_LOOP_FIXED_SLOTS: config 2 tracks at slot index = ms % 4 (no slot-phase walk)


def run_gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist

    from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver, nco_step32

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU implementation")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    steps, warm = args.steps, max(3, args.warmup)

    scene = make_scene(rank, N_MS)
    sig = cached_signal("trk_r%d_%d" % (rank, N_MS), scene)
    eng = Engine(device=local_rank, max_sv=211, ring_ms=N_MS + 24)
    stream = torch.cuda.Stream(device=dev)       # a real (non-default) stream shared by torch events and the engine
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def events(n):
        return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    # ---- closed-loop tracking: value (device-resident records, kernel only) and e2e (host library, host buffers)
    import ctypes as C
    channels = Channels([s.prn for s in scene.sats])
    rx = Receiver(eng, channels)
    n_ch = channels.n
    pinned_sig = torch.from_numpy(sig.copy()).pin_memory()
    eng.upload_signal(0, pinned_sig.numpy())
    ch_bytes, aux_bytes = eng.record_bytes()
    arm_locked(channels, scene)
    host_records = np.ctypeslib.as_array(C.cast(channels.base, C.POINTER(C.c_uint8)), (n_ch * ch_bytes,)).copy()
    d_pristine = torch.from_numpy(host_records).to(dev)
    d_records = torch.empty_like(d_pristine)
    d_aux = torch.zeros(n_ch * aux_bytes, dtype=torch.uint8, device=dev)
    d_iq = torch.zeros(N_MS * n_ch * 6, dtype=torch.int16, device=dev)
    d_nav = torch.zeros(N_MS * n_ch, dtype=torch.int8, device=dev)
    d_result = torch.zeros(n_ch * 24, dtype=torch.uint8, device=dev)

    def device_step():
        d_records.copy_(d_pristine)          # re-arm (not timed)
        d_aux.zero_()

    for _ in range(warm):
        device_step()
        eng.track_loop_dev_ex(n_ch, d_records.data_ptr(), d_aux.data_ptr(), 0, N_MS, d_iq.data_ptr(), d_nav.data_ptr(),
                              d_result.data_ptr(), LOOP_FIXED_SLOTS)
    barrier()
    clocks = ClockSampler(local_rank)
    launches0 = eng.launch_count
    ev = events(steps)
    for k in range(steps):
        device_step()
        flush.fill_(k)                       # evict L2 between timed iterations (not timed)
        ev[k][0].record(stream)
        eng.track_loop_dev_ex(n_ch, d_records.data_ptr(), d_aux.data_ptr(), 0, N_MS, d_iq.data_ptr(), d_nav.data_ptr(),
                              d_result.data_ptr(), LOOP_FIXED_SLOTS)
        ev[k][1].record(stream)
    barrier()
    t_dev = float(np.sum([a.elapsed_time(b) for a, b in ev])) / 1e3
    launches_value = eng.launch_count - launches0
    res = d_result.cpu().numpy().view(np.uint32).reshape(n_ch, 6)
    assert (res[:, 0] == N_MS).all() and (res[:, 1] == 0).all(), res      # every channel ran every millisecond on the device
    iq_dev = d_iq.cpu().numpy().reshape(N_MS, n_ch, 6)

    # e2e: what a user of the host library calls, host buffers in and out
    rx.set_loop_site(0)
    def e2e_call():
        if NO_STREAM:
            eng.upload_signal(0, pinned_sig.numpy())
            return rx.track_run(0, N_MS, log=True)
        return rx.track_stream(0, pinned_sig.numpy(), log=True)

    for _ in range(warm):
        arm_locked(channels, scene)
        e2e_call()
    barrier()
    ev = events(steps)
    wall_e2e = 0.0
    for k in range(steps):
        arm_locked(channels, scene)
        flush.fill_(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev[k][0].record(stream)
        # host buffer -> HBM ring in chunks WHILE the loop launch is tracking; records in/out, per-ms sums and nav bits
        # land in host arrays; all inside the timed region
        iq_log, nav_log = e2e_call()
        ev[k][1].record(stream)
        torch.cuda.synchronize()
        wall_e2e += time.perf_counter() - t0
    barrier()
    t_e2e = float(np.sum([a.elapsed_time(b) for a, b in ev])) / 1e3
    t_e2e = max(t_e2e, wall_e2e)             # the call is synchronous: never report less than its wall time
    launches = eng.launch_count - launches0
    on_device, on_host = rx.loop_stats()
    assert np.array_equal(iq_log, iq_dev), "device-resident and host-library runs disagree"
    nav_dev = d_nav.cpu().numpy().reshape(N_MS, n_ch)
    assert np.array_equal(nav_log, nav_dev), "device-resident and host-library runs disagree (nav bits)"
    final_records = [bytes(channels.snapshot(i)) for i in range(channels.n)]
    final_fine = [np.uint32(channels.snapshot(i).code_phase_fine_bits).view(np.float32) for i in range(channels.n)]

    # the same second with the loop filters on the HOST (one GPU round trip per millisecond), for comparison
    rx.set_loop_site(1)
    arm_locked(channels, scene)
    rx.track_run(0, N_MS, log=False)
    host_loop = []
    for k in range(2):
        arm_locked(channels, scene)
        t0 = time.perf_counter()
        iq_host, _ = rx.track_run(0, N_MS, log=True)
        host_loop.append(time.perf_counter() - t0)
    assert np.array_equal(iq_host, iq_dev), "host-resident and device-resident loops disagree"
    host_loop_ms = min(host_loop) * 1e3
    rx.set_loop_site(0)

    # upload-then-run (the non-overlapped form of the same call sequence), for comparison
    t_seq = []
    for k in range(5):
        arm_locked(channels, scene)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.upload_signal(0, pinned_sig.numpy())
        rx.track_run(0, N_MS, log=True)
        t_seq.append(time.perf_counter() - t0)
    upload_then_run_ms = min(t_seq) * 1e3

    # ---- config 5 on ONE GPU: 32 different satellites tracked continuously from a host-resident stream through a ring
    # SHORTER than the run (256 ms), i.e. with the producer refilling the ring behind the loop; one launch.
    from stm32f4_sdr_gps_b200.signal_synth import config2_scene, synthesize_blocks
    n_many = 32
    many_scene = config2_scene(n_ms=N_MS, prns=ALL_PRNS[:n_many], seed=0x5D120005)
    many_path = Path(tempfile.gettempdir()) / ("gpsb_bench_cfg5_%d.npy" % N_MS)
    if many_path.exists():
        many_sig = np.load(many_path)
    else:
        many_sig = synthesize_blocks(many_scene.sats, N_MS, 0x5D120005)
        if rank == 0:
            try:
                np.save(many_path, many_sig)
            except OSError:
                pass
    pinned_many = torch.from_numpy(many_sig.copy()).pin_memory()
    many = Channels([s.prn for s in many_scene.sats])
    many_blank = [many.snapshot(i) for i in range(n_many)]
    stream_eng = Engine(device=local_rank, max_sv=211, ring_ms=256)
    many_rx = Receiver(stream_eng, many)
    t_many = []
    many_iq = many_nav = None
    for k in range(0 if NO_STREAM else 5):
        for i in range(n_many):
            many.restore(i, many_blank[i])
        arm_locked(many, many_scene)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        many_iq, many_nav = many_rx.track_stream(0, pinned_many.numpy(), log=True)
        t_many.append(time.perf_counter() - t0)
    many_ms = min(t_many[1:]) * 1e3 if t_many else float("nan")
    many_dev, many_host = many_rx.loop_stats()
    many_records = [bytes(many.snapshot(i)) for i in range(n_many)]
    many_rx.close()
    stream_eng.close()

    # ---- open-loop batch replay of the same 4000 cells in ONE launch (kernel-level throughput)
    from stm32f4_sdr_gps_b200 import EPL_REQ  # noqa: F401
    rq_all, _, _ = truth_requests(scene, nco_step32)
    n_cells = rq_all.size
    d_rq = torch.from_numpy(rq_all.view(np.uint8).copy()).to(dev)
    d_out = torch.zeros(n_cells * 6, dtype=torch.int16, device=dev)
    for _ in range(warm):
        eng.track_epl_dev(n_cells, d_rq.data_ptr(), d_out.data_ptr())
    barrier()
    ev = events(steps)
    for k in range(steps):
        flush.fill_(k)
        ev[k][0].record(stream)
        eng.track_epl_dev(n_cells, d_rq.data_ptr(), d_out.data_ptr())
        ev[k][1].record(stream)
    barrier()
    batch_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))

    # ---- config 1 batched: PRN 1, Doppler +2000 Hz, byte offset 100, one cell per millisecond of a LONG recording
    # (every cell has its own 2046-byte frame, so the working set - 400 000 ms = 818 MB - exceeds L2 and the kernel is
    # fed from HBM): k_epl_batch, prompt arm only (gps_correlation_iq) and all three arms
    n_long = int(os.environ.get("GPSB_BENCH_LONG_MS", "400000"))
    long_eng = Engine(device=local_rank, max_sv=4, ring_ms=n_long)
    long_eng.set_stream(stream.cuda_stream)
    long_eng.set_code_prn(1, 1)
    rng_long = np.random.default_rng(0x5D120001 + 17 * rank)
    long_sig = rng_long.integers(0, 256, (n_long, 2046), dtype=np.uint8)
    long_eng.upload_signal(0, long_sig)
    step_2k = nco_step32(np.float32(IF_HZ + 2000))
    rq_long = np.zeros(n_long, EPL_REQ)
    rq_long["sv_slot"], rq_long["ms_index"] = 1, np.arange(n_long)
    rq_long["step32"] = step_2k                      # stateless mixer: phase 0 at the start of every millisecond
    rq_long["off_e"], rq_long["off_p"], rq_long["off_l"] = 99, 100, 101
    d_rq_long = torch.from_numpy(rq_long.view(np.uint8).copy()).to(dev)
    d_out_long = torch.zeros(n_long * 6, dtype=torch.int16, device=dev)
    long_ms = {}
    for arms, fn in ((1, long_eng.prompt_iq_dev), (3, long_eng.track_epl_dev)):
        for _ in range(warm):
            fn(n_long, d_rq_long.data_ptr(), d_out_long.data_ptr())
        barrier()
        ev = events(steps)
        for k in range(steps):
            flush.fill_(k)
            ev[k][0].record(stream)
            fn(n_long, d_rq_long.data_ptr(), d_out_long.data_ptr())
            ev[k][1].record(stream)
        barrier()
        long_ms[arms] = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        if arms == 1:
            long_prompt = d_out_long.cpu().numpy()[:2 * n_long].reshape(n_long, 2).copy()
            long_eng.set_epl_batch_kernel(1)                     # the register-staged kernel this one replaced, for comparison
            for _ in range(warm):
                fn(n_long, d_rq_long.data_ptr(), d_out_long.data_ptr())
            barrier()
            ev = events(steps)
            for k in range(steps):
                flush.fill_(k)
                ev[k][0].record(stream)
                fn(n_long, d_rq_long.data_ptr(), d_out_long.data_ptr())
                ev[k][1].record(stream)
            barrier()
            long1_reg_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
            assert np.array_equal(d_out_long.cpu().numpy()[:2 * n_long].reshape(n_long, 2), long_prompt), "the two batch kernels disagree"
            long_eng.set_epl_batch_kernel(0)
    # config 1 as the reference runs it: ONE cell.  Kernel time of a one-cell launch (events) and the latency of the
    # synchronous host call gpsb_prompt_iq (request up, launch, result down).
    ev = events(50)
    for k in range(50):
        ev[k][0].record(stream)
        long_eng.prompt_iq_dev(1, d_rq_long.data_ptr(), d_out_long.data_ptr())
        ev[k][1].record(stream)
    barrier()
    one_cell_kernel_us = float(np.median([a.elapsed_time(b) for a, b in ev][10:])) * 1e3
    t_one = []
    for k in range(60):
        t0 = time.perf_counter()
        one = long_eng.prompt_iq(rq_long[:1])
        t_one.append(time.perf_counter() - t0)
    one_cell_call_us = float(np.median(t_one[10:])) * 1e6
    assert np.array_equal(one[0], long_prompt[0])
    for _ in range(2):                      # restore the full three-arm result for the comparison below
        long_eng.track_epl_dev(n_long, d_rq_long.data_ptr(), d_out_long.data_ptr())
    long_epl = d_out_long.cpu().numpy().reshape(n_long, 6)
    assert np.array_equal(long_epl[:, 2:4], long_prompt), "prompt-only and E/P/L forms of k_epl_batch disagree"
    long_eng.close()

    # ---- cold acquisition (secondary metric, satellites sharded over ranks)
    from stm32f4_sdr_gps_b200.signal_synth import config3_scene
    acq_scene = config3_scene(n_ms=ACQ_MS)
    acq_sig = cached_signal("acq_%d" % ACQ_MS, acq_scene)
    for prn in range(1, ACQ_SV + 1):
        eng.set_code_prn(prn, prn)
    eng.upload_signal(N_MS, acq_sig)                    # frames N_MS .. N_MS+9 of the ring
    step = np.array([nco_step32(np.float32(IF_HZ - 5000 + 500 * b)) for b in range(ACQ_BINS)], np.uint32)
    # One sweep = all 32 satellites x 21 bins x 10 ms.  value: the SHARDED sweep - the 210 (bin, ms) cell groups dealt
    # round-robin over the ranks (every rank keeps whole 8-satellite tiles), ncclAllGather of the triples on the context
    # stream straight out of device memory, permutation into the (sv, bin, ms) grid - timed warmed over `steps`
    # iterations, gather INCLUDED.  Beside it the whole sweep on this rank's GPU alone (what one GPU needs).
    if world > 1:
        eng.comm_init_torch()
    all_sv = np.arange(1, ACQ_SV + 1, dtype=np.uint32)
    d_sv = torch.from_numpy(all_sv.view(np.int32)).to(dev)
    d_step = torch.from_numpy(step.view(np.int32)).to(dev)
    n_acq_cells = ACQ_SV * ACQ_BINS * ACQ_MS
    d_res = torch.zeros(n_acq_cells * 4, dtype=torch.int16, device=dev)
    acq_ms = {}
    for method, name in ((0, "direct"), (1, "dp4a")):
        eng.set_sweep_method(method)
        for _ in range(3):
            eng.sweep_dev(d_sv.data_ptr(), ACQ_SV, d_step.data_ptr(), ACQ_BINS, N_MS, ACQ_MS, 0, d_res.data_ptr())
        barrier()
        ev = events(steps)
        for k in range(steps):
            flush.fill_(k)
            ev[k][0].record(stream)
            eng.sweep_dev(d_sv.data_ptr(), ACQ_SV, d_step.data_ptr(), ACQ_BINS, N_MS, ACQ_MS, 0, d_res.data_ptr())
            ev[k][1].record(stream)
        barrier()
        acq_ms[name] = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    one_gpu_grid = d_res.cpu().numpy().view(np.uint16).reshape(ACQ_SV, ACQ_BINS, ACQ_MS, 4).copy()
    eng.set_sweep_method(1)
    for _ in range(max(3, warm)):
        d_grid_ptr = eng.sweep_gather_dev(d_sv.data_ptr(), ACQ_SV, d_step.data_ptr(), ACQ_BINS, N_MS, ACQ_MS, 0)
    barrier()
    ev = events(steps)
    for k in range(steps):
        flush.fill_(k)
        if world > 1:
            dist.barrier()                    # the ranks start a sweep together, as one job would
        ev[k][0].record(stream)
        d_grid_ptr = eng.sweep_gather_dev(d_sv.data_ptr(), ACQ_SV, d_step.data_ptr(), ACQ_BINS, N_MS, ACQ_MS, 0)
        ev[k][1].record(stream)
    barrier()
    acq_ms["gathered"] = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    grid = eng.sweep_gather(all_sv, step, N_MS, ACQ_MS).view(np.uint16).reshape(ACQ_SV, ACQ_BINS, ACQ_MS, 4)
    assert np.array_equal(grid, one_gpu_grid), "sharded + gathered sweep differs from the sweep on one GPU"
    my_sv = all_sv
    # the fine grid of SURVEY.md section 8(d) config 3: 2046 half-chip offsets x 8 sub-byte shifts = 16368 phases
    eng.set_sweep_method(1)
    ev = events(steps)
    for k in range(steps):
        flush.fill_(k)
        ev[k][0].record(stream)
        for bits in range(8):
            eng.sweep_dev(d_sv.data_ptr(), my_sv.size, d_step.data_ptr(), ACQ_BINS, N_MS, ACQ_MS, bits, d_res.data_ptr())
        ev[k][1].record(stream)
    barrier()
    acq_fine_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    eng.sweep_dev(d_sv.data_ptr(), my_sv.size, d_step.data_ptr(), ACQ_BINS, N_MS, ACQ_MS, 0, d_res.data_ptr())   # leave the bits-0 grid
    acq_fine_ms *= 1.0                    # (whole fine grid on this rank's GPU alone)
    # end to end: host signal in, all-channel Doppler votes out (upload + sweep + D2H + host chain votes)
    acq_ch = Channels([int(p) for p in my_sv])
    acq_rx = Receiver(eng, acq_ch)
    pinned_acq = torch.from_numpy(acq_sig.copy()).pin_memory()
    t_acq_e2e = []
    for k in range(5):
        for i in range(acq_ch.n):
            st = acq_ch.snapshot(i)
            st.acq_state = 0
            acq_ch.restore(i, st)
        t0 = time.perf_counter()
        eng.upload_signal(N_MS, pinned_acq.numpy())
        votes, phases = acq_rx.cold_sweep(-5000, 500, ACQ_BINS, N_MS, ACQ_MS)
        t_acq_e2e.append((time.perf_counter() - t0) * 1e3)
    acq_e2e_ms = float(np.min(t_acq_e2e))
    found = int(sum(1 for i in range(acq_ch.n) if acq_ch.snapshot(i).acq_state == 2))

    # ---- config 4 (BASELINE configs[3]): 32 PRNs searched, 10 in the sky; cold sweeps -> code-phase rounds 1..3 ->
    # pre-track -> C4_MS ms of closed-loop tracking with nav bits (slot-phase walk on).  Several GPUs: the sweep's cell
    # groups sharded + all-gathered (gpsb_sweep_gather), every rank runs the same votes, the satellites found are dealt
    # round-robin over the ranks for everything after (no further exchange).  Every output of every satellite is
    # diffed against the unmodified reference run on the same snapshots (tests/config4_lib.py).
    sys.path.insert(0, str(REPO / "tests"))
    import config4_lib as c4
    C4_MS = int(os.environ.get("GPSB_BENCH_C4_MS", "10000"))
    sc4 = c4.scene(C4_MS + 700)
    path4 = Path(tempfile.gettempdir()) / ("gpsb_bench_cfg4_%d.npy" % C4_MS)
    if rank == 0 and not path4.exists():
        tmp4 = path4.with_suffix(".tmp.npy")
        np.save(tmp4, c4.signal(sc4))
        os.replace(tmp4, path4)
    barrier()
    sig4 = np.load(path4)
    eng4 = Engine(device=local_rank, max_sv=40, ring_ms=sc4.n_ms)
    if world > 1:
        eng4.comm_init_torch()
    eng4.upload_signal(0, sig4)
    serve = dict(serve_rank=rank, serve_world=world)
    ch_w, rx_w, _, _ = c4.product(eng4, sig4, c4.SEARCHED, C4_MS, cold_start_opts=serve)      # warm-up: staging buffers at their final size, NCCL
    rx_w.close(); ch_w.free()
    tm4 = {}
    ch4, rx4, rep4, logs4 = c4.product(eng4, sig4, c4.SEARCHED, C4_MS, timers=tm4, cold_start_opts=serve, before_start=barrier)
    c4_times = [tm4["cold_start_s"], tm4["pre_track_s"], tm4["tracking_s"], tm4["total_s"]]
    mine4 = [i % world == rank for i in range(len(c4.SEARCHED))]
    t0 = time.perf_counter()
    refs4 = c4.reference_all(sig4, c4.SEARCHED, rep4, C4_MS, served=mine4, procs=max(1, (os.cpu_count() or 1) // world))
    c4_twin_s = time.perf_counter() - t0
    sum4 = c4.diff(ch4, logs4, refs4, c4.SEARCHED)          # raises on the first differing field / sum / nav bit
    sum4["acquired"] = [a for a in sum4["acquired"] if mine4[c4.SEARCHED.index(a["prn"])]]
    if world > 1:
        box = [None] * world
        dist.all_gather_object(box, (sum4, rep4, c4_twin_s))
    else:
        box = [(sum4, rep4, c4_twin_s)]
    rx4.close(); ch4.free(); eng4.close()
    clk = clocks.stop()

    # ---- max over ranks
    times = torch.tensor([t_dev, t_e2e, acq_ms["dp4a"], acq_ms["direct"], acq_e2e_ms, batch_ms, many_ms, long_ms[1], long_ms[3],
                          acq_ms["gathered"]] + c4_times, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    (t_dev, t_e2e, acq_dp4a_ms, acq_direct_ms, acq_e2e_ms, batch_ms, many_ms, long1_ms, long3_ms, acq_gathered_ms,
     c4_cold_s, c4_pre_s, c4_trk_s, c4_total_s) = [float(x) for x in times.cpu()]

    rank_has_prn12 = True        # the gathered grid holds every satellite on every rank
    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text())
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        units_step = world * N_SV_PER_GPU * ARMS * MS_SAMPLES * N_MS
        value = units_step * steps / t_dev
        e2e = units_step * steps / t_e2e
        # algorithmic bytes of one k_track_run launch: every frame once (all channels share it
        # through L2), 128 B of code + channel record and scratch in and out per SV, 12 + 1 B of logs per cell
        loop_bytes = N_MS * 2046 + N_SV_PER_GPU * (128 + 2 * (ch_bytes + aux_bytes) + 24) + N_MS * N_SV_PER_GPU * 13
        loop_kernel_ms = t_dev * 1e3 / steps
        loop_achieved = loop_bytes / (loop_kernel_ms * 1e-3) / 1e9
        batch_bytes = N_MS * 2046 + N_SV_PER_GPU * 128 + n_cells * (24 + 12)
        acq_bitmacs = ACQ_SV * ACQ_BINS * ACQ_MS * 2046 * 2 * 16368
        acq_alg_bytes = ACQ_MS * 2046 + ACQ_SV * 128 + n_acq_cells * 8
        # dp4a issue slots of the sweep: 4 correlations (I/Q x parity) x 1023 lags x 256 steps per cell
        acq_dp4a = ACQ_SV * ACQ_BINS * ACQ_MS * 4 * 1023 * 256
        sm_clk = (clk.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
        idp_peak = 148 * 64 * sm_clk                       # IDP.4A: 64 lanes/clk/SM measured (tools/ubench_int.cu)
        # ---- the reference's own C on this box's host cores, same workload, and the at-size parity of this run
        sys.path.insert(0, str(REPO / "tests"))
        from oracle_lib import best_o3_variant
        host_cores = os.cpu_count() or 1
        parity = {"config4": {"vs": "unmodified reference per satellite", "channel_records": len(c4.SEARCHED),
                              "tracking_cells": int(sum(b[0]["cells"] for b in box)), "equal": True},
                  "cold_acq_sharded_vs_one_gpu": {"cells": int(n_acq_cells), "equal": True}}
        cpu_t, cpu_cores, ref_logs = reference_tracking([(scene, sig)], host_cores, reps=3, want_logs=True)
        cpu = cpu_1core = cpu_o3 = None
        if cpu_t:
            for i, (r_iq, r_nav, r_rec) in enumerate(ref_logs):     # the reference's sums, nav bits and records vs this run's
                assert np.array_equal(r_iq, iq_dev[:, i, :]), "config 2: sums of satellite %d differ from the reference's" % i
                assert np.array_equal(r_nav, nav_dev[:, i]), "config 2: nav bits of satellite %d differ from the reference's" % i
                assert r_rec == final_records[i], "config 2: final channel record %d differs from the reference's" % i
            parity["config2_closed_loop"] = {"vs": "unmodified reference (oracle/_ref), same recording, same start",
                                             "cells": int(n_ch * N_MS), "sums": int(n_ch * N_MS * 6), "nav_rows": int(n_ch * N_MS),
                                             "channel_records": n_ch, "equal": True}
            unit_sv = N_SV_PER_GPU * ARMS * MS_SAMPLES * N_MS
            cpu = {"value": unit_sv / cpu_t, "unit": "arm-samples/s", "cores": cpu_cores, "host_cores": host_cores,
                   "kind": "reference",
                   "sample": "whole N=1 workload (4 SV x 1000 ms closed-loop gps_tracking_process), best of 3; one process per "
                             "satellite - a channel's 1-kHz loop is serial, so 4 satellites can use 4 of the box's %d cores; "
                             "32 satellites on all cores: see config5.cpu_baseline" % host_cores}
            t1, _, _ = reference_tracking([(scene, sig)], 1, reps=3)
            cpu_1core = {"value": unit_sv / t1, "unit": "arm-samples/s", "cores": 1, "kind": "reference",
                         "sample": "same workload, the four satellites one after the other on one core"}
            level, so_o3 = best_o3_variant()
            if so_o3 is not None:
                t3, c3, _ = reference_tracking([(scene, sig)], host_cores, reps=3, so_path=so_o3)
                cpu_o3 = {"value": unit_sv / t3, "unit": "arm-samples/s", "cores": c3, "kind": "reference",
                          "flags": "-O3 -march=%s -ffp-contract=off" % level,
                          "sample": "same workload, the reference compiled at -O3 for the highest x86-64 level this box's CPU "
                                    "has (the reference sources are not on the bench box, so -march=native cannot be built there)"}
        # config 5 on one GPU: all 32 satellites against the reference, and the reference on all cores beside it
        cfg5_cpu = None
        if many_iq is not None:
            t5, c5, logs5 = reference_tracking([(many_scene, many_sig)], host_cores, reps=2, want_logs=True)
            if t5:
                for i, (r_iq, r_nav, r_rec) in enumerate(logs5):
                    assert np.array_equal(r_iq, many_iq[:, i, :]), "config 5: sums of satellite %d differ from the reference's" % i
                    assert np.array_equal(r_nav, many_nav[:, i]), "config 5: nav bits of satellite %d differ from the reference's" % i
                    assert r_rec == many_records[i], "config 5: final channel record %d differs from the reference's" % i
                parity["config5_streaming_32sv"] = {"vs": "unmodified reference (oracle/_ref)", "cells": int(n_many * N_MS),
                                                    "sums": int(n_many * N_MS * 6), "nav_rows": int(n_many * N_MS),
                                                    "channel_records": n_many, "equal": True}
                cfg5_cpu = {"value": n_many * ARMS * MS_SAMPLES * N_MS / t5, "unit": "arm-samples/s", "cores": c5,
                            "host_cores": host_cores, "kind": "reference", "ms_per_s_of_signal": t5 * 1e3,
                            "sample": "whole workload: 32 SV x 1000 ms, one process per satellite on %d cores at once" % c5}
        acq_cpu_rate, acq_cpu_cells, acq_ref_cells = reference_sweep_rate(acq_sig)
        if acq_ref_cells is not None and rank_has_prn12:
            mine = grid[:acq_ref_cells.shape[0], :, :acq_ref_cells.shape[2], :3]
            assert np.array_equal(mine, acq_ref_cells), "cold acquisition: (max, phase, avg) differ from the reference's"
            parity["config3_sweep_sample"] = {"vs": "unmodified reference correlation_search", "cells": int(acq_ref_cells.size // 3),
                                              "equal": True}
        prompt_cpu_rate, prompt_cpu_iq = reference_prompt_rate(long_sig)
        if prompt_cpu_iq is not None:      # the reference's own I/Q for the head of the long recording
            assert np.array_equal(prompt_cpu_iq, long_prompt[:prompt_cpu_iq.shape[0]]), "config 1: GPU and reference C disagree"
            parity["config1_batched"] = {"vs": "unmodified reference gps_correlation_iq", "cells": int(prompt_cpu_iq.shape[0]),
                                         "equal": True}
        line = {
            "metric": "correlator-samples/sec (E/P/L arms)", "value": value, "unit": "arm-samples/s",
            "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": t_dev * 1e3 / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 xor/popcount (sums), f32/f64 (loop filters)", "data": "synthetic",
            "config": {"workload": "config2: %d-SV E/P/L closed-loop tracking, 1 s @16.368 Msps 1-bit IF (4 SV per GPU, "
                                   "every SV every ms, DLL/PLL/FLL + nav-bit logic in the loop)" % (world * N_SV_PER_GPU),
                       "n_sv": world * N_SV_PER_GPU, "n_ms": N_MS, "cells_per_step": world * n_cells,
                       "loop_site": "device (k_track_run, one launch per step)",
                       "l2": "flushed between timed iterations (256 MiB fill)", "parallelism": "sv-shard x%d" % world},
            "e2e": {"value": e2e, "unit": "arm-samples/s",
                    "h2d_bytes_per_step": int(sig.nbytes + n_ch * (ch_bytes + aux_bytes)),
                    "d2h_bytes_per_step": int(n_ch * (ch_bytes + aux_bytes + 24) + n_cells * 13),
                    "ms_per_step": t_e2e * 1e3 / steps,
                    "api": "gpsb_rx_track_stream (libgpsb_host.so): host buffers in and out, signal DMA-ed into the HBM ring in 64-ms "
                           "chunks while the loop launch is tracking"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"kernel": "k_track_run (1 launch per step, %d CTAs: one per satellite)" % n_ch, "bound": "hbm",
                         "achieved": loop_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": loop_achieved / hbm_peak,
                         "traffic": profile_traffic("k_track_run_r2.txt") if (N_MS == 1000 and n_ch == 4) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this "
                                           "launch, read from profiles/k_track_run_r2.txt",
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                         "algorithmic_bytes_per_launch": loop_bytes, "kernel_ms": loop_kernel_ms,
                         "us_per_ms_of_signal": loop_kernel_ms * 1e3 / N_MS,
                         "note": "serial by construction: millisecond t+1 is planned by the loop filters from the sums of "
                                 "millisecond t (tracking.c:92-170), so the kernel is bound by the dependent-issue latency of "
                                 "one SM per satellite, not by bandwidth; the same launch carries 1 to 148 satellites in "
                                 "the same time"},
            "cpu_baseline": cpu,
            "cpu_baseline_1core": cpu_1core,
            "cpu_baseline_o3": cpu_o3,
            "parity_checked": parity,
            "config5": {"what": "config 5 on ONE GPU: %d different satellites tracked continuously from a HOST-resident stream "
                                "through a 256-ms HBM ring (run = %d ms, the producer refills the ring behind the loop; one "
                                "k_track_run launch, %d CTAs); host buffers in, per-ms sums + nav bits + records out"
                                % (n_many, N_MS, n_many),
                        "value": n_many * ARMS * MS_SAMPLES * N_MS / (many_ms * 1e-3), "unit": "arm-samples/s",
                        "ms_per_s_of_signal": many_ms * 1e3 / N_MS, "times_real_time": N_MS / many_ms,
                        "input_msps_sustained": MS_SAMPLES * N_MS / (many_ms * 1e-3) / 1e6,
                        "target": "163.68 Msps (10x real time)",
                        "channel_ms_on_device": int(many_dev), "channel_ms_on_host_path": int(many_host),
                        "cpu_baseline": cfg5_cpu,
                        "vs_cpu_baseline": None if not cfg5_cpu else
                        n_many * ARMS * MS_SAMPLES * N_MS / (many_ms * 1e-3) / cfg5_cpu["value"]},
            "closed_loop": {"device_loop_kernel_ms": loop_kernel_ms, "host_loop_ms": host_loop_ms,
                            "upload_then_run_ms": upload_then_run_ms,
                            "host_loop_what": "same second with the loop filters on the host: one GPU round trip per ms",
                            "launches_per_step_value": launches_value / steps,
                            "channel_ms_on_device": int(on_device), "channel_ms_on_host_path": int(on_host),
                            "final_code_phase": [float(x) for x in final_fine]},
            "batch_replay": {"what": "the same %d cells replayed open loop in ONE k_epl_batch launch (signal + requests resident)" % n_cells,
                             "value": N_SV_PER_GPU * ARMS * MS_SAMPLES * N_MS / (batch_ms * 1e-3), "unit": "arm-samples/s",
                             "kernel_ms": batch_ms,
                             "roofline": {"bound": "hbm", "achieved": batch_bytes / (batch_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                          "unit": "GB/s", "frac": batch_bytes / (batch_ms * 1e-3) / 1e9 / hbm_peak,
                                          "algorithmic_bytes_per_launch": batch_bytes}},
            "config1_batched": {
                "what": "PRN 1, +2000 Hz, byte offset 100, bits 0: one prompt correlation per millisecond of a %d-ms recording "
                        "(%.0f MB, larger than L2) in ONE k_epl_batch_tma launch per rank (frames through a TMA ring in shared "
                        "memory); 'epl' = all three arms of the same cells"
                        % (n_long, n_long * 2048 / 1e6),
                "cells": n_long * world,
                "single_cell": {"what": "one cell, as the reference runs config 1", "kernel_us": one_cell_kernel_us,
                                "host_call_us": one_cell_call_us,
                                "api": "gpsb_prompt_iq(n = 1): request H2D, k_epl_batch<1> launch, result D2H, synchronous"},
                "prompt": {"kernel_ms": long1_ms, "cells_per_s": n_long * world / (long1_ms * 1e-3),
                           "arm_samples_per_s": n_long * world * MS_SAMPLES / (long1_ms * 1e-3),
                           "roofline": {"kernel": "k_epl_batch_tma<1>", "bound": "hbm",
                                        "achieved": n_long * (2046 + 24 + 4) / (long1_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                        "unit": "GB/s", "frac": n_long * (2046 + 24 + 4) / (long1_ms * 1e-3) / 1e9 / hbm_peak,
                                        "algorithmic_bytes_per_launch": n_long * (2046 + 24 + 4),
                                        "traffic": profile_traffic("k_epl_batch_tma1_r2.txt") if n_long == 400000 else None,
                                        "traffic_source": "read from profiles/k_epl_batch_tma1_r2.txt (ncu --set full of this launch)",
                                        "limit": "issue slots: ALU pipe 70 %, issue active 75 %, XU (POPC) 53 %, scoreboard stalls "
                                                 "0.4 per issue (profiles/k_epl_batch_tma1_r2.txt); the register-staged kernel it "
                                                 "replaces waited on its loads (profiles/k_epl_batch1_r1.txt)"},
                           "register_staged_kernel_ms": long1_reg_ms},
                "epl": {"kernel_ms": long3_ms, "cells_per_s": n_long * world / (long3_ms * 1e-3),
                        "arm_samples_per_s": n_long * world * ARMS * MS_SAMPLES / (long3_ms * 1e-3),
                        "roofline": {"kernel": "k_epl_batch_tma<3>", "bound": "hbm",
                                     "achieved": n_long * (2046 + 24 + 12) / (long3_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                     "unit": "GB/s", "frac": n_long * (2046 + 24 + 12) / (long3_ms * 1e-3) / 1e9 / hbm_peak,
                                     "algorithmic_bytes_per_launch": n_long * (2046 + 24 + 12),
                                     "popc_pipe_frac": n_long * 512 * 6 / (long3_ms * 1e-3) / (148 * 16 * sm_clk)}},
                "cpu_baseline": None if not prompt_cpu_rate else {
                    "value": prompt_cpu_rate, "unit": "cells/s (1 core)", "cores": 1, "kind": "reference",
                    "sample": "first 8000 ms: gps_generate_prn_data2 + gps_shift_to_zero_freq + gps_correlation_iq per ms; "
                              "I/Q identical to the GPU's"}},
            "config4": {
                "what": "BASELINE configs[3]: 32 PRNs searched (10 in the sky), %d Doppler sweep(s) of 29 bins x 10 ms -> code-phase "
                        "rounds 1..3 (look-ahead windows) -> pre-track -> %d ms closed-loop tracking with nav bits, slot-phase "
                        "walk on; %s" % (rep4["n_sweeps"], C4_MS,
                                         "one GPU" if world == 1 else "sweep cell groups sharded + ncclAllGather, satellites found "
                                         "dealt round-robin over %d ranks" % world),
                "wall_ms": c4_total_s * 1e3, "cold_start_ms": c4_cold_s * 1e3, "pre_track_ms": c4_pre_s * 1e3,
                "tracking_ms": c4_trk_s * 1e3,
                "time_to_decision_ms": c4_cold_s * 1e3, "time_to_first_tracking_ms": (c4_cold_s + c4_pre_s) * 1e3,
                "signal_ms": int(rep4["ms_next"] + C4_MS), "times_real_time": (rep4["ms_next"] + C4_MS) / (c4_total_s * 1e3),
                "launches_cold_start_rank0": rep4["launches"], "schedule_rank0": rep4,
                "found_prns": sorted(p for b in box for p in [a["prn"] for a in b[0]["acquired"]]),
                "in_the_sky": sorted(s_.prn for s_ in sc4.sats),
                "bit_edges_refined": int(sum(a["bit_edge_refined"] for b in box for a in b[0]["acquired"])),
                "parity": {"vs": "unmodified reference (oracle/_ref) per satellite on the same snapshots", "equal": True,
                           "channel_records": len(c4.SEARCHED), "tracking_cells": int(sum(b[0]["cells"] for b in box)),
                           "nav_bits": int(sum(b[0]["nav_bits"] for b in box))},
                "reference_twin_s": max(b[2] for b in box),
                "reference_twin_what": "the same work by the unmodified reference C, one process per satellite on the box's "
                                       "cores (driven from Python, dominated by the 870 Doppler cells per PRN)"},
            "cold_acq": {"metric": "full-sky 32-SV cold-acq ms", "value": acq_gathered_ms, "unit": "ms",
                         "value_what": "sweep sharded over %d rank(s) by (bin, ms) cell group + ncclAllGather of the triples on "
                                       "the context stream + permutation into the (sv, bin, ms) grid; warmed, events, max over "
                                       "ranks" % world,
                         "one_gpu_ms": acq_dp4a_ms,
                         "direct_xor_popc_ms": acq_direct_ms, "e2e_ms": acq_e2e_ms,
                         "fine_grid_16368_phases_ms": acq_fine_ms,
                         "cells": ACQ_SV * ACQ_BINS * ACQ_MS, "phases": 2046, "bit_macs": acq_bitmacs,
                         "bit_macs_per_s": acq_bitmacs / (acq_gathered_ms * 1e-3), "doppler_votes_passed_rank0": found,
                         "cpu_baseline": None if not acq_cpu_rate else {
                             "value": ACQ_SV * ACQ_BINS * ACQ_MS / acq_cpu_rate * 1e3, "unit": "ms (extrapolated, 1 core)",
                             "cores": 1, "kind": "reference",
                             "sample": "%d of the 6720 cells (2 SV x 21 bins x 2 ms), correlation_search 0..2046" % acq_cpu_cells},
                         "sharding": "(bin, ms) cell groups round-robin over %d rank(s), whole 8-satellite tiles per rank" % world,
                         "roofline": {"kernel": "k_acq_dp4a", "bound": "int-dot-product pipe (IDP.4A 64 lanes/clk/SM)",
                                      "what": "the whole sweep on one GPU (one_gpu_ms)",
                                      "achieved": acq_dp4a / (acq_dp4a_ms * 1e-3) / 1e12, "peak": idp_peak / 1e12,
                                      "unit": "T dp4a/s", "frac": acq_dp4a / (acq_dp4a_ms * 1e-3) / idp_peak,
                                      "frac_sharded": acq_dp4a / (acq_gathered_ms * 1e-3) / (idp_peak * world),
                                      "hbm_frac": acq_alg_bytes / (acq_dp4a_ms * 1e-3) / 1e9 / hbm_peak}},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    rx.close()
    acq_rx.close()
    eng.close()


_json_out = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's real stdout."""
    out = _json_out or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    # stdout carries the JSON line and nothing else: the compiled reference (oracle/_ref) printf()s its acquisition
    # decisions ("PRN=.. FINAL FREQ=..") from C, in this process and in the forked per-satellite workers.  File descriptor 1
    # is pointed at /dev/null for everybody (GPSB_BENCH_KEEP_STDOUT=1: at stderr); the JSON line goes to a private
    # duplicate of the original stdout.
    global _json_out
    sys.stdout.flush()
    _json_out = os.fdopen(os.dup(1), "w")
    if os.environ.get("GPSB_BENCH_KEEP_STDOUT"):
        os.dup2(2, 1)
    else:
        sink = os.open(os.devnull, os.O_WRONLY)
        os.dup2(sink, 1)
        os.close(sink)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpsb", choices=["gpsb", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
