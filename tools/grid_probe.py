#!/usr/bin/env python
"""k_track_run: time per millisecond against the number of CTAs in the launch (channels = copies of the same four
satellites).  Diagnostic only."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver  # noqa: E402
from stm32f4_sdr_gps_b200.signal_synth import Satellite, Scene, synthesize  # noqa: E402

n_ms = 1000
rng = np.random.default_rng(5)
sats = [Satellite(prn=p, doppler_hz=float(rng.uniform(-4000, 4000)), code_phase_samples=float(rng.uniform(0, 16368)),
                  cn0_dbhz=48.0, nav_bit_offset_ms=int(rng.integers(0, 20))) for p in range(1, 5)]
scene = Scene(sats=sats, n_ms=n_ms, seed=77)
sig = synthesize(scene)
eng = Engine(device=0, max_sv=211, ring_ms=n_ms + 8)
eng.upload_signal(0, sig)
for n_ch in [int(a) for a in sys.argv[1:]] or [4, 32, 74, 148, 152, 296]:
    many = Scene(sats=[sats[i % 4] for i in range(n_ch)], n_ms=n_ms, seed=77)
    ch = Channels([s.prn for s in many.sats])
    rx = Receiver(eng, ch)
    best = 1e9
    for rep in range(4):
        bench.arm_locked(ch, many)
        t0 = time.perf_counter()
        rx.track_run(0, n_ms, log=False)
        best = min(best, time.perf_counter() - t0)
    same = all(bytes(ch.snapshot(i)) == bytes(ch.snapshot(i % 4)) for i in range(n_ch))
    print("n_ch %4d  %8.3f us per ms   copies identical: %s" % (n_ch, best / n_ms * 1e6, same), flush=True)
    rx.close(); ch.free()
eng.close()
