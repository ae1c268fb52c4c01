"""dev/big_n_check.py -- element counts beyond 2^31: index arithmetic of the striped accumulation (implicit ids) and of the
least-squares kernel.  Needs ~60 GB of device memory."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
n = (1 << 31) + 1000
val = torch.ones(n, dtype=torch.float32, device=dev)
k = 1024
g = torch.zeros(k, device=dev)
x.accumulate(None, val, g)           # id = i mod k
torch.cuda.synchronize()
want = torch.full((k,), n // k, dtype=torch.float64)
want[: n % k] += 1
assert torch.equal(g.cpu().double(), want), (g[:4], want[:4])
print("accumulate implicit ids, n = 2^31 + 1000: exact")
del val
m = (1 << 31) + 7
data = torch.empty((m, 3), dtype=torch.float64, device=dev)
data[:, 0] = 1.0; data[:, 1] = 2.0; data[:, 2] = 3.0
prm = torch.zeros(8, dtype=torch.float64, device=dev); prm[:4] = torch.tensor([0.5, 1.0, 1.0, 0.25], dtype=torch.float64)
loss = torch.zeros(1, dtype=torch.float64, device=dev)
x.lsq_grad(data, prm, loss)
torch.cuda.synchronize()
# one point: u = a - x1 = -0.5, v = c - x2 = -1, r = u^2 + b v^2 + d - y = 0.25 + 1 + 0.25 - 3 = -1.5; grads 2r[2u, v^2, 2bv, 1]
r = -1.5
per = torch.tensor([2 * r * 2 * -0.5, 2 * r * 1.0, 2 * r * 2 * 1.0 * -1.0, 2 * r], dtype=torch.float64)
got = prm[4:].cpu()
assert torch.allclose(got, per * m, rtol=1e-12), (got, per * m)
assert abs(loss.item() - r * r * m) <= 1e-12 * r * r * m
print("lsq_grad, n = 2^31 + 7: matches the closed form")
