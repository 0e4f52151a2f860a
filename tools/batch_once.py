#!/usr/bin/env python
"""A few k_epl_batch launches at the size of bench.py's config1_batched leg (for ncu captures and quick timing).
Usage: python tools/batch_once.py [n_ms] [arms] [kernel: 0 = TMA ring (default), 1 = register-staged]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stm32f4_sdr_gps_b200 import EPL_REQ, Engine, nco_step32  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
arms = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
eng = Engine(device=0, max_sv=4, ring_ms=n)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
eng.set_epl_batch_kernel(kernel)
eng.set_code_prn(1, 1)
sig = np.random.default_rng(1).integers(0, 256, (n, 2046), dtype=np.uint8)
eng.upload_signal(0, sig)
rq = np.zeros(n, EPL_REQ)
rq["sv_slot"], rq["ms_index"] = 1, np.arange(n)
rq["step32"] = nco_step32(np.float32(4092000 + 2000))
rq["off_e"], rq["off_p"], rq["off_l"] = 99, 100, 101
d_rq = torch.from_numpy(rq.view(np.uint8).copy()).to(dev)
d_out = torch.zeros(n * 6, dtype=torch.int16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fn = eng.prompt_iq_dev if arms == 1 else eng.track_epl_dev
ts = []
for k in range(8):
    flush.fill_(k)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    fn(n, d_rq.data_ptr(), d_out.data_ptr())
    b.record(stream)
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
t = float(np.median(ts[2:]))
print(("k_epl_batch_tma" if kernel == 0 else "k_epl_batch") + "<%d>: %d cells %.1f us  %.2f Gcells/s  %.0f GB/s" % (arms, n, t * 1e3, n / t / 1e6, n * 2046 / t / 1e6))
eng.close()
