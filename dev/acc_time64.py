"""dev/acc_time64.py -- accumulate 2^24 -> 1024 in fp64 (tagged-table kernel) and fp32 on unaligned bases."""
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
    ti = torch.from_numpy(idx).to(dev)
    tv64 = torch.from_numpy(val).to(dev).double()
    g64 = torch.zeros(1024, dtype=torch.float64, device=dev)
    t64 = timed(lambda: x.accumulate(ti, tv64, g64))
    ti_u = torch.zeros(n + 1, dtype=torch.int32, device=dev); ti_u[1:] = ti
    tv_u = torch.zeros(n + 1, dtype=torch.float32, device=dev); tv_u[1:] = torch.from_numpy(val).to(dev)
    g32 = torch.zeros(1024, device=dev)
    tu = timed(lambda: x.accumulate(ti_u[1:], tv_u[1:], g32))
    print(f"{dist:8s} fp64 {t64:7.1f} us ({12 * n / t64 / 1e6:.0f} GB/s)   fp32 unaligned {tu:7.1f} us")
