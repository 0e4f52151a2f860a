#!/usr/bin/env python
"""gpsb_rx_track_stream where copies cannot run beside the loop kernel (CUDA_LAUNCH_BLOCKING=1 makes every launch
synchronous, as a profiler's kernel replay does): the call must still return the right sums, and quickly.
    CUDA_LAUNCH_BLOCKING=1 python tools/stream_blocked_probe.py"""
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver  # noqa: E402

n_ms = int(os.environ.get("PROBE_MS", "1000"))
scene = bench.make_scene(0, n_ms)
sig = bench.cached_signal("trk_r0_%d" % n_ms, scene)
out = []
for ring in (1024, 256):
    eng = Engine(device=0, max_sv=211, ring_ms=ring)
    ch = Channels([s.prn for s in scene.sats])
    rx = Receiver(eng, ch)
    for rep in range(2):
        bench.arm_locked(ch, scene)
        t0 = time.perf_counter()
        iq, nav = rx.track_stream(0, sig, log=True)
        dt = time.perf_counter() - t0
    out.append(iq)
    print("ring %4d: %.1f ms, loop stats %s" % (ring, dt * 1e3, rx.loop_stats()), flush=True)
    rx.close(); ch.free(); eng.close()
assert np.array_equal(out[0], out[1])
print("sums identical:", True)
