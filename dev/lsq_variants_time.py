import os, sys, torch
ROOT = "/root/repo" if os.path.exists("/root/repo/tests") else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=30, fl=True):
    ts = []
    for i in range(reps + 3):
        if fl: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
prm = torch.zeros(8, dtype=torch.float64, device=dev); prm[1] = 1.0
n = 1_000_000
sets = [torch.from_numpy(orc.lsq_data(n, 42 + i)).to(dev) for i in range(8)]
t1 = timed(lambda: x.lsq_grad(sets[0], prm))
for s_ in sets: x.lsq_grad(s_, prm)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(6):
    for s_ in sets: x.lsq_grad(s_, prm)
b.record(); torch.cuda.synchronize()
print(f"LSQ {sys.argv[1]}: 1M flushed {t1:.2f} us, back to back {a.elapsed_time(b) / 48 * 1e3:.2f} us", flush=True)
