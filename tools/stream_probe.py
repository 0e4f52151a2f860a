#!/usr/bin/env python
"""Timing probe for the streaming closed loop (gpsb_rx_track_stream): wall time of one second of signal that
starts in pinned host memory, for several chunk sizes, next to upload-then-run.  Diagnostic only."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver  # noqa: E402


def main():
    n_ms = 1000
    scene = bench.make_scene(0, n_ms)
    sig = bench.cached_signal("trk_r0_%d" % n_ms, scene)
    pinned = torch.from_numpy(sig.copy()).pin_memory().numpy()
    eng = Engine(device=0, max_sv=211, ring_ms=n_ms + 24)
    ch = Channels([s.prn for s in scene.sats])
    rx = Receiver(eng, ch)
    ref = None
    for name, chunk in (("upload + run", None), ("stream 16", 16), ("stream 32", 32), ("stream 64", 64), ("stream 128", 128),
                        ("stream 256", 256), ("stream 1000", 1000)):
        best = 1e9
        for rep in range(6):
            bench.arm_locked(ch, scene)
            t0 = time.perf_counter()
            if chunk is None:
                eng.upload_signal(0, pinned)
                iq, nav = rx.track_run(0, n_ms, log=True)
            else:
                iq, nav = rx.track_stream(0, pinned, chunk_ms=chunk, log=True)
            best = min(best, time.perf_counter() - t0)
        if ref is None:
            ref = iq
        print("%-14s %8.3f ms   identical: %s  stats %s" % (name, best * 1e3, np.array_equal(iq, ref), rx.loop_stats()), flush=True)
    rx.close()
    eng.close()


if __name__ == "__main__":
    main()
