#!/usr/bin/env python
"""One warm-up and a few 1000-ms device-resident tracking runs of bench config 2 (for ncu captures of k_track_run)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver  # noqa: E402

n_ms = 1000
scene = bench.make_scene(0, n_ms)
sig = bench.cached_signal("trk_r0_%d" % n_ms, scene)
eng = Engine(device=0, max_sv=211, ring_ms=n_ms + 8)
eng.upload_signal(0, sig)
ch = Channels([s.prn for s in scene.sats])
rx = Receiver(eng, ch)
rx.set_loop_site(2)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    bench.arm_locked(ch, scene)
    rx.track_run(0, n_ms, log=True)
rx.close()
eng.close()
