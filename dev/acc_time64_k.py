"""dev/acc_time64_k.py -- fp64 accumulation of 2^24 ids as a function of K."""
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
for k in (1024, 3000, 8192, 40000):
    for dist in ("uniform", "zipf"):
        idx, val = orc.accumulate_inputs(n, k, dist, 42)
        ti = torch.from_numpy(idx).to(dev); tv = torch.from_numpy(val).to(dev).double()
        g = torch.zeros(k, dtype=torch.float64, device=dev)
        print(f"fp64 K={k:6d} {dist:8s} {timed(lambda: x.accumulate(ti, tv, g)):8.1f} us")
