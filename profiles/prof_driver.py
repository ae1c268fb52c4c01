"""profiles/prof_driver.py -- launches every hot kernel twice at BASELINE size, for ncu.

  ncu --set full --clock-control none --import-source on \
      -k regex:'covproj_tma|covproj_sharedw|batched_kernel|lsq_grad|accumulate|splat_forward|splat_backward|adam_kernel' -c 24 \
      -o gpurun_out/prof_rNN python profiles/prof_driver.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc  # noqa: E402  (input generators only)
import xyz_autodiff_cuda_b200 as x  # noqa: E402

dev = torch.device("cuda:0")
reps = int(os.environ.get("PROF_REPS", "2"))
which = os.environ.get("PROF_ONLY", "covproj,covproj_shared_w,batched,lsq,accumulate,splat,adam,splat_band").split(",")

if "covproj" in which:
    E = 1 << 26
    ins = [torch.empty((E, w), device=dev).uniform_(-1, 1) for w in (6, 9, 6, 3)]
    outs = [torch.empty((E, w), device=dev) for w in (3, 6, 9, 6)]
    for _ in range(reps):
        x.covproj_fwd_bwd(*ins, *outs)
    torch.cuda.synchronize()
    del ins, outs

if "covproj_shared_w" in which:
    E = 1 << 26
    ins = [torch.empty((E, w), device=dev).uniform_(-1, 1) for w in (6, 6, 3)]
    outs = [torch.empty((E, w), device=dev) for w in (3, 6, 6)]
    w9 = torch.empty(9, device=dev).uniform_(-1, 1)
    gw9 = torch.zeros(9, device=dev)
    for _ in range(reps):
        x.covproj_shared_w_fwd_bwd(ins[0], w9, ins[1], ins[2], outs[0], outs[1], gw9, outs[2])
    torch.cuda.synchronize()
    del ins, outs

if "batched" in which:
    # include/xyz_autodiff/batched.cuh through the two user graphs of tests/csrc/batched_probe.cu
    import ctypes
    so = os.path.join(ROOT, "tests", "csrc", "_build", "libxyz_batched.so")
    if os.path.exists(so):
        B = ctypes.CDLL(so)
        B.batched_lsq.restype = ctypes.c_float
        B.batched_lsq.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        B.batched_chain.restype = ctypes.c_float
        B.batched_chain.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        n = 1 << 25
        tin = torch.empty((n, 15), device=dev).uniform_(-1, 1)
        tout = torch.empty((n, 15), device=dev)
        w18 = torch.empty(18, device=dev).uniform_(-1, 1)
        gw18 = torch.zeros(18, device=dev)
        B.batched_chain(tin.data_ptr(), tout.data_ptr(), n, w18.data_ptr(), gw18.data_ptr(), reps)
        del tin, tout
        m = 1 << 27
        data = torch.empty((m, 3), dtype=torch.float64, device=dev).uniform_(-5, 5)
        v5 = torch.tensor([0.0, 1.0, 0.0, 0.0, 0.0], dtype=torch.float64, device=dev)
        g5 = torch.zeros(5, dtype=torch.float64, device=dev)
        B.batched_lsq(data.data_ptr(), m, v5.data_ptr(), g5.data_ptr(), reps)
        torch.cuda.synchronize()
        del data

if "lsq" in which:
    n = 1 << 28
    data = torch.empty((n, 3), dtype=torch.float64, device=dev).uniform_(-5, 5)
    prm = torch.zeros(8, dtype=torch.float64, device=dev)
    prm[1] = 1.0
    for _ in range(reps):
        x.lsq_grad(data, prm)
    torch.cuda.synchronize()
    del data
    data = torch.from_numpy(orc.lsq_data(1_000_000, 42)).to(dev)  # BASELINE configs[0] size: 24 MB, launch/latency bound
    for _ in range(reps):
        x.lsq_grad(data, prm)
    torch.cuda.synchronize()
    del data

if "accumulate" in which:
    n = 1 << 24
    for dist in ("uniform", "zipf", "same"):
        idx, val = orc.accumulate_inputs(n, 1024, dist, 42)
        ti, tv = torch.from_numpy(idx).to(dev), torch.from_numpy(val).to(dev)
        grad = torch.zeros(1024, device=dev)
        for _ in range(reps):
            x.accumulate(ti, tv, grad)
        torch.cuda.synchronize()

if "splat" in which or "adam" in which:
    W = H = 1024
    N = 100_000
    params, target = orc.splat_c4_scene(N, W, H, 42)
    tp, tt = torch.from_numpy(params).to(dev), torch.from_numpy(target).to(dev)
    grads = torch.zeros((N, 9), device=dev)
    img = torch.zeros((W * H, 3), device=dev)
    loss = torch.zeros(1, device=dev)
    adam = torch.zeros((N, 18), device=dev)
    for it in range(reps):
        x.zero_gradients(grads)
        loss.zero_()
        if "splat" in which:
            x.launch_gaussian_splatting(tp, grads, tt, img, loss, W, H, N)
        if "adam" in which:
            x.adam_step_individual(tp, grads, adam, 0.1, 0.01, 0.001, 0.02, 0.05, iteration=it + 1)
    torch.cuda.synchronize()
if "splat_band" in which:
    # round 2: what ONE rank of the 8-GPU row-band configuration runs per iteration: a 128-row band of the C4 scene through
    # a caller-owned workspace (band-local histograms, three-deep staging in the forward pass), then the fused
    # peer-memory optimiser step (world size 1 here: the reduce-scatter / all-gather degenerate to local traffic)
    W = H = 1024
    N = 100_000
    params, target = orc.splat_c4_scene(N, W, H, 42)
    tt = torch.from_numpy(target).to(dev)
    grp = x.PeerGroup(0, 1, lambda h: [h])
    ps = x.PeerSplat(grp, N, lambda h: [h])
    ps.params.copy_(torch.from_numpy(params).to(dev))
    img = torch.zeros((W * H, 3), device=dev)
    loss = torch.zeros(1, device=dev)
    ws = x.SplatWorkspace(W, H, N, 1_200_000, rows=(448, 576))
    for it in range(reps):
        loss.zero_()
        ws.launch(ps.params, ps.grads, tt, img, loss)
        ps.adam_step(0.1, 0.01, 0.001, 0.02, 0.05, iteration=it + 1, total_loss=loss)
    torch.cuda.synchronize()
    assert not ws.status()["overflowed"]
print("prof_driver done", x.launch_count(), "launches")
