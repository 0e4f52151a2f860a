#!/usr/bin/env python
"""Where the end-to-end time of one streamed second goes: resident kernel, streaming build of the kernel with the
whole recording already behind the watermark, and the host call gpsb_rx_track_stream.  Diagnostic."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver  # noqa: E402

n_ms = 1000
scene = bench.make_scene(0, n_ms)
sig = bench.cached_signal("trk_r0_%d" % n_ms, scene)
pinned = torch.from_numpy(sig.copy()).pin_memory().numpy()
dev = torch.device("cuda", 0)
eng = Engine(device=0, max_sv=211, ring_ms=n_ms + 24)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
ch = Channels([s.prn for s in scene.sats])
rx = Receiver(eng, ch)
eng.upload_signal(0, pinned)
bench.arm_locked(ch, scene)
ch_b, aux_b = eng.record_bytes()
rec = np.ctypeslib.as_array(C.cast(ch.base, C.POINTER(C.c_uint8)), (ch.n * ch_b,)).copy()
d_pr = torch.from_numpy(rec).to(dev)
d_rec = torch.empty_like(d_pr)
d_aux = torch.zeros(ch.n * aux_b, dtype=torch.uint8, device=dev)
d_iq = torch.zeros(n_ms * ch.n * 6, dtype=torch.int16, device=dev)
d_nav = torch.zeros(n_ms * ch.n, dtype=torch.int8, device=dev)
d_res = torch.zeros(ch.n * 24, dtype=torch.uint8, device=dev)
for flags, name in ((0, "resident kernel"), (1, "streaming build, watermark already at the end"),
                    (2, "resident kernel, fixed slots (no walk code)"), (3, "streaming build, fixed slots, watermark at the end")):
    ts = []
    for k in range(8):
        d_rec.copy_(d_pr)
        d_aux.zero_()
        if flags & 1:
            eng.stream_reset(n_ms)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        eng.track_loop_dev_ex(ch.n, d_rec.data_ptr(), d_aux.data_ptr(), 0, n_ms, d_iq.data_ptr(), d_nav.data_ptr(), d_res.data_ptr(), flags)
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("%-48s %.3f ms" % (name, min(ts)))
for name, fn in (("gpsb_rx_track_run (records in/out, resident signal)", lambda: rx.track_run(0, n_ms, log=True)),
                 ("gpsb_rx_track_stream (signal from pinned host)", lambda: rx.track_stream(0, pinned, log=True))):
    ts = []
    for k in range(8):
        bench.arm_locked(ch, scene)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    print("%-48s %.3f ms" % (name, min(ts) * 1e3))
# the same two calls the way bench.py times them: L2 flushed (a 256-MB fill) before every call
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fn in (("gpsb_rx_track_run, L2 flushed before the call", lambda: rx.track_run(0, n_ms, log=True)),
                 ("gpsb_rx_track_stream, L2 flushed before the call", lambda: rx.track_stream(0, pinned, log=True))):
    ts = []
    for k in range(8):
        bench.arm_locked(ch, scene)
        flush.fill_(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    print("%-48s %.3f ms (median %.3f)" % (name, min(ts) * 1e3, float(np.median(ts)) * 1e3))
rx.close()
eng.close()
