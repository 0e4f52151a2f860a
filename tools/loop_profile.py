#!/usr/bin/env python
"""Per-phase clock64 ticks of k_track_run (GPSB_LOOP_PROFILE=1 diagnostic build): where a millisecond goes."""
import os
import sys
import time
from pathlib import Path

os.environ["GPSB_LOOP_PROFILE"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver  # noqa: E402


def main():
    n_ms = 1000
    scene = bench.make_scene(0, n_ms)
    sig = bench.cached_signal("trk_r0_%d" % n_ms, scene)
    eng = Engine(device=0, max_sv=211, ring_ms=n_ms + 8)
    eng.upload_signal(0, sig)
    ch = Channels([s.prn for s in scene.sats])
    rx = Receiver(eng, ch)
    rx.set_loop_site(2)
    for rep in range(3):
        bench.arm_locked(ch, scene)
        t0 = time.perf_counter()
        rx.track_run(0, n_ms, log=False)
        print("wall %.3f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
    rx.close()
    eng.close()


if __name__ == "__main__":
    main()
