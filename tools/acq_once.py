#!/usr/bin/env python
"""Times the full-sky sweep (32 SV x 21 bins x 10 ms x 2046 phases) for the sweep methods (diagnostic)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stm32f4_sdr_gps_b200 import Engine, nco_step32  # noqa: E402

dev = torch.device("cuda", 0)
eng = Engine(device=0, max_sv=40, ring_ms=64)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
for p in range(1, 33):
    eng.set_code_prn(p, p)
eng.upload_signal(0, np.random.default_rng(1).integers(0, 256, (10, 2046), dtype=np.uint8))
sv = torch.arange(1, 33, dtype=torch.int32, device=dev)
step = torch.tensor([nco_step32(np.float32(4092000 - 5000 + 500 * b)) for b in range(21)], dtype=torch.int64).to(torch.int32).to(dev)
res = torch.zeros(32 * 21 * 10 * 4, dtype=torch.int16, device=dev)
out = {}
for method in (1, 2):
    eng.set_sweep_method(method)
    ts = []
    for k in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        eng.sweep_dev(sv.data_ptr(), 32, step.data_ptr(), 21, 0, 10, 0, res.data_ptr())
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out[method] = res.cpu().numpy().copy()
    print("method %d: %.1f us" % (method, float(np.median(ts[3:])) * 1e3))
print("identical:", np.array_equal(out[1], out[2]))
eng.close()
