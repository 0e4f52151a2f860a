#!/usr/bin/env python
"""Config 4 on one GPU, wall time per stage (cold start: sweeps + code rounds; pre-track; tracking).  Diagnostic only -
bench.py's config4 leg is what reports numbers.  Usage: python tools/c4_probe.py [n_track_ms]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import config4_lib as c4  # noqa: E402
from stm32f4_sdr_gps_b200 import Engine  # noqa: E402

n_track = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
sc = c4.scene(n_track + 600)
sig = c4.signal(sc)
eng = Engine(device=0, max_sv=40, ring_ms=sc.n_ms)
eng.upload_signal(0, sig)
for rep_no in range(3):
    timers = {}
    ch, rx, rep, logs = c4.product(eng, sig, c4.SEARCHED, n_track, timers=timers)
    print("cold start %.3f ms (%d launches: %d sweeps + code rounds), pre-track call %.3f ms, tracking %.3f ms, total %.3f ms; "
          "first tracking after %.3f ms" % (timers["cold_start_s"] * 1e3, rep["launches"], rep["n_sweeps"], timers["pre_track_s"] * 1e3,
                                            timers["tracking_s"] * 1e3, timers["total_s"] * 1e3,
                                            (timers["cold_start_s"] + timers["pre_track_s"]) * 1e3), flush=True)
    rx.close(); ch.free()
eng.close()
