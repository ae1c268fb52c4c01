"""Multi-GPU parity (`-m gpu`, needs >= 2 GPUs; skipped on a single-GPU box): one process per GPU over NCCL.
Every configuration is run sharded (element ranges / row bands / views) with the shared-gradient all-reduce and
compared with the same work done by ONE rank through the same C ABI, and with the fp64 oracle."""
import os
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import oracle_lib as orc
from importlib import import_module
import xyz_autodiff_cuda_b200 as x
par = import_module("xyz_autodiff_cuda_b200.parallel")
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def timed(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b_.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b_) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3

# C1: element ranges, all-reduce of 4 fp64 sums + loss
data = orc.lsq_data(1_000_003, seed=5)
vals = (0.3, 1.2, -0.4, 0.1)
b, e = par.shard_range(data.shape[0], rank, world)
prm = torch.zeros(8, dtype=torch.float64, device=dev); prm[:4] = torch.tensor(vals, dtype=torch.float64)
loss = torch.zeros(1, dtype=torch.float64, device=dev)
x.lsq_grad(D(data[b:e]), prm, loss)
g = prm[4:].clone()
par.allreduce_shared_grads(g, loss)
full_g, full_l = orc.lsq_grad(data, vals)
assert np.allclose(g.cpu().numpy(), full_g, rtol=1e-10), (g, full_g)
assert abs(loss.item() - full_l) <= 1e-10 * full_l

# C1 again with the exchange INSIDE the kernel (peer mailboxes over NVLink, no NCCL call): bit-identical on all ranks
pg = par.make_peer_group(x)
ref_bits = None
for it in range(6):   # several calls: sequence numbers, both mailbox parities
    prm2 = torch.zeros(8, dtype=torch.float64, device=dev); prm2[:4] = torch.tensor(vals, dtype=torch.float64)
    loss2 = torch.zeros(1, dtype=torch.float64, device=dev)
    x.lsq_grad_allreduce(D(data[b:e]), prm2, pg, loss2)
    torch.cuda.synchronize()
    g2 = prm2[4:].clone()
    assert np.allclose(g2.cpu().numpy(), full_g, rtol=1e-10), (it, g2, full_g)
    assert abs(loss2.item() - full_l) <= 1e-10 * full_l
    same = g2.clone(); dist.broadcast(same, 0)
    assert torch.equal(same, g2), "ranks disagree bitwise"
    if ref_bits is None: ref_bits = g2.clone()
    assert torch.equal(ref_bits, g2), "run-to-run bits differ"
# timing, informational (rank 0 prints): kernel + NCCL all-reduce of the 5 sums vs. the fused kernel
dd = D(data[b:e])
buf = torch.zeros(8, dtype=torch.float64, device=dev)
def nccl_path():
    x.lsq_grad(dd, buf, None)
    dist.all_reduce(buf[4:])
def fused_path():
    x.lsq_grad_allreduce(dd, buf, pg, None)
t_nccl, t_fused = timed(nccl_path), timed(fused_path)
if rank == 0:
    print(f"TIMING lsq 1M points over {world} GPUs: kernel + NCCL all-reduce {t_nccl:.1f} us/iter, fused peer-memory kernel {t_fused:.1f} us/iter")
# a rank with no points still takes part
x.lsq_grad_allreduce(D(data[b:e]) if rank else torch.empty((0, 3), dtype=torch.float64, device=dev), prm2, pg, None)
torch.cuda.synchronize()
dist.barrier()
pg.close()

# C2: K = 1024 fp32 accumulators
idx, val = orc.accumulate_inputs(1 << 21, 1024, "zipf", seed=2)
b, e = par.shard_range(idx.size, rank, world)
grad = torch.zeros(1024, device=dev)
x.accumulate(D(idx[b:e]), D(val[b:e]), grad)
par.allreduce_shared_grads(grad)
exact = orc.accumulate_exact(idx, val, 1024)
assert (np.abs(grad.cpu().numpy() - exact) <= 1e-4 * orc.accumulate_exact(idx, np.abs(val), 1024) + 1e-30).all()
# ... and with the exchange inside the finishing kernel (peer mailboxes), odd shard boundaries included
pg2 = par.make_peer_group(x)
for n_el in (idx.size, idx.size - 3, 1000):
    b2, e2 = par.shard_range(n_el, rank, world)
    if n_el == idx.size - 3: b2, e2 = min(b2 + 1, e2), e2      # unaligned slice
    lo = dist.get_rank()
    spans = [None] * world; dist.all_gather_object(spans, (b2, e2))
    keep = np.zeros(idx.size, bool)
    for (bb, ee) in spans: keep[bb:ee] = True
    grad2 = torch.zeros(1024, device=dev)
    x.accumulate_allreduce(D(idx[b2:e2]), D(val[b2:e2]), grad2, pg2)
    torch.cuda.synchronize()
    ex = orc.accumulate_exact(idx[keep], val[keep], 1024)
    assert (np.abs(grad2.cpu().numpy() - ex) <= 1e-4 * orc.accumulate_exact(idx[keep], np.abs(val[keep]), 1024) + 1e-30).all()
    same = grad2.clone(); dist.broadcast(same, 0)
    assert torch.equal(same, grad2), "ranks disagree bitwise"
ti_, tv_ = D(idx[b:e]), D(val[b:e])
def acc_nccl():
    x.accumulate(ti_, tv_, grad); dist.all_reduce(grad)
def acc_fused():
    x.accumulate_allreduce(ti_, tv_, grad, pg2)
t_a, t_b = timed(acc_nccl), timed(acc_fused)
if rank == 0:
    print(f"TIMING accumulate 2^21 -> 1024 over {world} GPUs: kernel + NCCL all-reduce {t_a:.1f} us/iter, fused peer-memory finish {t_b:.1f} us/iter")
dist.barrier(); pg2.close()

# C3: element ranges, no collective: every rank's slice equals the oracle on that slice
J, W_, S, go = orc.covproj_inputs(100_000, seed=3)
b, e = par.shard_range(J.shape[0], rank, world)
outs = [torch.empty((e - b, k), device=dev) for k in (3, 6, 9, 6)]
x.covproj_fwd_bwd(D(J[b:e]), D(W_[b:e]), D(S[b:e]), D(go[b:e]), *outs)
ref = orc.covproj(J[b:e], W_[b:e], S[b:e], go[b:e], np.float64)
for got, want in zip(outs, ref):
    assert np.abs(got.cpu().numpy() - want).max() <= 1e-5 * np.abs(want).max()

# C3 variant B (one shared W): element ranges + the all-reduce of the 9 shared gradients inside the kernel
pg3 = par.make_peer_group(x)
W9 = W_[0].copy()
Wrep = np.broadcast_to(W9, (J.shape[0], 9)).copy()
ref_gW = orc.covproj(J, Wrep, S, go, np.float64)[2]
for it in range(3):
    o3, gJ3, gS3 = [torch.empty((e - b, k), device=dev) for k in (3, 6, 6)]
    gW3 = torch.zeros(9, device=dev)
    x.covproj_shared_w_fwd_bwd(D(J[b:e]), D(W9), D(S[b:e]), D(go[b:e]), o3, gJ3, gW3, gS3, group=pg3)
    torch.cuda.synchronize()
    assert (np.abs(gW3.cpu().numpy() - ref_gW.sum(0)) <= 1e-4 * np.abs(ref_gW).sum(0)).all(), it
    same = gW3.clone(); dist.broadcast(same, 0)
    assert torch.equal(same, gW3), "ranks disagree bitwise"
    want_slice = orc.covproj(J[b:e], Wrep[b:e], S[b:e], go[b:e], np.float64)
    assert np.abs(o3.cpu().numpy() - want_slice[0]).max() <= 1e-5 * np.abs(want_slice[0]).max()
    assert np.abs(gJ3.cpu().numpy() - want_slice[1]).max() <= 1e-5 * np.abs(want_slice[1]).max()
dist.barrier(); pg3.close()

# C4 on G GPUs: tile-aligned row bands of one image, Gaussians replicated, all-reduce of grads + loss
W, H, N = 160, 128, 300
params, target = orc.splat_scene(N, W, H, seed=21)
rg, ro, rl, tol = orc.splat_tolerance(params, target, W, H)
tp, tt = D(params), D(target)
grads = torch.zeros((N, 9), device=dev); out = torch.zeros((W * H, 3), device=dev); l = torch.zeros(1, device=dev)
par.splat_iteration_sharded(x, tp, grads, [tt], [out], l, W, H, mode="rows")
assert (np.abs(grads.cpu().numpy() - rg) <= tol).all()
assert abs(l.item() - rl) <= 1e-4 * abs(rl)
r0, r1 = par.row_bands(H, world)[rank]
o = out.cpu().numpy().reshape(H, W, 3)[r0:r1]; want = ro.reshape(H, W, 3)[r0:r1]
assert (np.abs(o - want) <= 1e-5 * np.maximum(np.abs(want), np.abs(ro).max() * 1e-3)).all()

# C5: views round-robin over ranks, all-reduce == sum over views
V = 2 * world
targets = [orc.splat_scene(1, W, H, seed=100 + v)[1] for v in range(V)]
mine = par.views_for_rank(V, rank, world)
grads.zero_(); l.zero_()
outs = [torch.zeros((W * H, 3), device=dev) for _ in mine]
par.splat_iteration_sharded(x, tp, grads, [D(targets[v]) for v in mine], outs, l, W, H, mode="views")
want_g = np.zeros((N, 9)); want_l = 0.0; want_tol = np.zeros((N, 9))
for t_ in targets:
    g_, o_, l_, tol_ = orc.splat_tolerance(params, t_, W, H)
    want_g += g_; want_l += l_; want_tol += tol_
assert (np.abs(grads.cpu().numpy() - want_g) <= want_tol).all()
assert abs(l.item() - want_l) <= 1e-4 * abs(want_l)
# replicas stay in lock step: Adam on the reduced gradients gives identical parameters on every rank
adam = torch.zeros((N, 18), device=dev)
p2 = tp.clone()
x.adam_step_individual(p2, grads, adam, 0.1, 0.01, 0.001, 0.02, 0.05, iteration=1)
ref_p = p2.clone()
dist.broadcast(ref_p, 0)
assert torch.equal(ref_p, p2)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_paths_match_single_gpu_and_oracle_over_nccl():
    n = min(torch.cuda.device_count(), 8)
    with tempfile.TemporaryDirectory() as d:
        w = os.path.join(d, "worker.py")
        open(w, "w").write(WORKER)
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                              "--master-addr", "127.0.0.1", "--master-port", "29633", w, ROOT],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-4000:]
        assert res.stdout.count("ok") >= n
        for line in res.stdout.splitlines():
            if line.startswith("TIMING"):
                print(line)
