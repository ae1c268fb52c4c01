"""dev/acc_time.py -- accumulate 2^24 -> 1024: fast vs deterministic flag, three id distributions (L2 flushed)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=10):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
n = 1 << 24
for dist in ("uniform", "zipf", "same"):
    idx, val = orc.accumulate_inputs(n, 1024, dist, 42)
    ti, tv = torch.from_numpy(idx).to(dev), torch.from_numpy(val).to(dev)
    grad = torch.zeros(1024, device=dev)
    t0 = timed(lambda: x.accumulate(ti, tv, grad))
    t1 = timed(lambda: x.accumulate(ti, tv, grad, x.FLAG_DETERMINISTIC))
    print(f"{dist:8s} fast {t0:7.1f} us   deterministic {t1:7.1f} us")
