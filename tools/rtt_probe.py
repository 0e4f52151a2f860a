#!/usr/bin/env python
"""Closed-loop timing probe: wall time per millisecond of signal for several channel counts, with the loop
filters on the device (k_track_run) and on the host (resident session kernel / per-ms launches).
Diagnostic only - bench.py is what reports numbers."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver  # noqa: E402
from stm32f4_sdr_gps_b200.signal_synth import Satellite, Scene, synthesize  # noqa: E402


def main():
    n_ms = 1000
    rng = np.random.default_rng(5)
    import os
    quick = os.environ.get('GPSB_LOOP_EXPERIMENT') is not None
    for n_ch in ((4,) if quick else (4, 32)):
        sats = [Satellite(prn=p, doppler_hz=float(rng.uniform(-4000, 4000)), code_phase_samples=float(rng.uniform(0, 16368)),
                          cn0_dbhz=48.0, nav_bit_offset_ms=int(rng.integers(0, 20))) for p in range(1, n_ch + 1)]
        scene = Scene(sats=sats, n_ms=n_ms, seed=77)
        t0 = time.time()
        sig = synthesize(scene)
        print("n_ch %d: synthesis %.1f s" % (n_ch, time.time() - t0), flush=True)
        eng = Engine(device=0, max_sv=211, ring_ms=n_ms + 8)
        eng.upload_signal(0, sig)
        ch = Channels([s.prn for s in sats])
        rx = Receiver(eng, ch)
        finals = {}
        for site, threads, name in (((2, 1, "device loop"),) if quick else ((2, 1, "device loop"), (1, 1, "host loop, 1 thread"), (1, 0, "host loop, threads"))):
            rx.set_loop_site(site)
            rx.set_threads(threads)
            best = 1e9
            for rep in range(4):
                bench.arm_locked(ch, scene)
                t0 = time.perf_counter()
                rx.track_run(0, n_ms, log=False)
                best = min(best, time.perf_counter() - t0)
            finals[name] = [bytes(ch.snapshot(i)) for i in range(n_ch)]
            print("  n_ch %3d %-22s %8.3f us per ms  (%.1fx real time)  stats %s" % (
                n_ch, name, best / n_ms * 1e6, n_ms * 1e-3 / best, rx.loop_stats()), flush=True)
        names = list(finals)
        print("  final channel records identical across paths:", all(finals[n] == finals[names[0]] for n in names), flush=True)
        rx.close()
        eng.close()


if __name__ == "__main__":
    main()
