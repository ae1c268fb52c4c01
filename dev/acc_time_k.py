"""dev/acc_time_k.py -- fp32 accumulation of 2^24 uniform ids as a function of K (which kernel serves which range)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=7):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
n = 1 << 24
for k in (256, 1024, 1800, 2048, 3000, 4096, 8192, 16384, 40000, 100000):
    for dist in ("uniform", "zipf"):
        idx, val = orc.accumulate_inputs(n, k, dist, 42)
        ti, tv = torch.from_numpy(idx).to(dev), torch.from_numpy(val).to(dev)
        g = torch.zeros(k, device=dev)
        print(f"K={k:6d} {dist:8s} {timed(lambda: x.accumulate(ti, tv, g)):8.1f} us")
