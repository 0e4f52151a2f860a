import torch
x = torch.randint(0, 2**31-1, (818_400_000//4,), dtype=torch.int32, device="cuda")
flush = torch.empty(256<<20, dtype=torch.uint8, device="cuda")
for fn,name in ((lambda: x.sum(), "sum int32"), (lambda: x.max(), "max int32"), (lambda: torch.bitwise_xor(x[:x.numel()//2], x[x.numel()//2:]), "xor halves (r+w)")):
    ts=[]
    for k in range(6):
        flush.fill_(k); torch.cuda.synchronize()
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); y=fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t=sorted(ts)[len(ts)//2]
    print(name, "%.1f us  %.0f GB/s (bytes read only)" % (t*1e3, x.numel()*4/t/1e6))
