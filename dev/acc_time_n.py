"""dev/acc_time_n.py -- fp32 accumulation into 1024 bins as a function of n (kernel switch at n = 2^16)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=9):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
for n in (1000, 30_000, 65_535, 65_536, 200_000, 1 << 20, 1 << 22, 1 << 24, 1 << 26):
    for dist in ("uniform", "same"):
        idx, val = orc.accumulate_inputs(n, 1024, dist, 42)
        ti, tv = torch.from_numpy(idx).to(dev), torch.from_numpy(val).to(dev)
        g = torch.zeros(1024, device=dev)
        print(f"n={n:9d} {dist:8s} {timed(lambda: x.accumulate(ti, tv, g)):8.1f} us  det {timed(lambda: x.accumulate(ti, tv, g, x.FLAG_DETERMINISTIC)):8.1f} us")
