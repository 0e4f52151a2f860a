#!/usr/bin/env python
"""Where the end-to-end time of a cold start goes: upload, sweep (kernel + copies), host votes.  Diagnostic."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver, nco_step32  # noqa: E402
from stm32f4_sdr_gps_b200.signal_synth import config3_scene  # noqa: E402

scene = config3_scene(n_ms=10)
sig = bench.cached_signal("acq_10", scene)
eng = Engine(device=0, max_sv=211, ring_ms=64)
ch = Channels(list(range(1, 33)))
rx = Receiver(eng, ch)
step = np.array([nco_step32(np.float32(4092000 - 5000 + 500 * b)) for b in range(21)], np.uint32)
for name, fn in (("upload 10 ms", lambda: eng.upload_signal(0, sig)),
                 ("gpsb_sweep (copies + kernel)", lambda: eng.sweep(np.arange(1, 33, dtype=np.uint32), step, 0, 10, 0)),
                 ("gpsb_rx_cold_sweep (sweep + votes)", None)):
    ts = []
    for k in range(8):
        if fn is None:
            for i in range(ch.n):
                st = ch.snapshot(i)
                st.acq_state = 0
                ch.restore(i, st)
        t0 = time.perf_counter()
        if fn is None:
            rx.cold_sweep(-5000, 500, 21, 0, 10)
        else:
            fn()
        ts.append(time.perf_counter() - t0)
    print("%-36s %7.1f us" % (name, min(ts) * 1e6))
rx.close()
eng.close()
