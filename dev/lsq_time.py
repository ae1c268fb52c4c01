"""dev/lsq_time.py -- event timing of xyz_lsq_grad_f64 at 1M points (L2 flushed) and 2^28."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=20, fl=True):
    ts = []
    for i in range(reps + 3):
        if fl: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
prm = torch.zeros(8, dtype=torch.float64, device=dev); prm[1] = 1.0
for n in (100_000, 1_000_000, 10_000_000):
    data = torch.from_numpy(orc.lsq_data(n, 42)).to(dev)
    print(f"lsq_grad n={n}: {timed(lambda: x.lsq_grad(data, prm)):.1f} us (L2 flushed), {timed(lambda: x.lsq_grad(data, prm), fl=False):.1f} us (warm)")
# steady state: 8 input sets (192 MB > L2) back to back
n = 1_000_000
sets = [torch.from_numpy(orc.lsq_data(n, 42 + i)).to(dev) for i in range(8)]
for s_ in sets: x.lsq_grad(s_, prm)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(6):
    for s_ in sets: x.lsq_grad(s_, prm)
b.record(); torch.cuda.synchronize()
print(f"lsq_grad n=1M back to back over 8 sets: {a.elapsed_time(b) / 48 * 1e3:.2f} us per launch")
big = torch.empty((1 << 28, 3), dtype=torch.float64, device=dev).uniform_(-5, 5)
print(f"lsq_grad n=2^28: {timed(lambda: x.lsq_grad(big, prm), reps=5, fl=False):.1f} us")
