"""xyz-autodiff-cuda_b200/host_api.py (`-m gpu`): the host-buffer entry points that bench.py times as `e2e` -- pinned host
arrays in, host results out, copies overlapped with the kernels.  Every helper must return exactly what the device-resident
call returns on the same data."""
import os
import sys
from importlib import import_module

import numpy as np
import pytest
import torch

import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x

pytestmark = pytest.mark.gpu
host_api = import_module("xyz_autodiff_cuda_b200.host_api")
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("n,chunk,depth", [(1000, 256, 2), (300_001, 1 << 16, 3), (1 << 18, 1 << 20, 4), (5, 4, 2)])
def test_covproj_host_pipeline_equals_the_device_call(n, chunk, depth):
    J, W, S, g = orc.covproj_inputs(n, seed=n)
    h_in = [torch.from_numpy(a).pin_memory() for a in (J, W, S, g)]
    h_out = [torch.full((n, k), float("nan"), dtype=torch.float32).pin_memory() for k in (3, 6, 9, 6)]
    pipe = host_api.CovprojHostPipeline(DEV, chunk_elems=chunk, depth=depth)
    h2d, d2h = pipe.run(h_in, h_out)
    torch.cuda.synchronize()
    assert h2d == n * 24 * 4 and d2h == n * 24 * 4
    d_out = [torch.empty((n, k), device=DEV) for k in (3, 6, 9, 6)]
    x.covproj_fwd_bwd(*[t.to(DEV) for t in h_in], *d_out)
    for a, b in zip(h_out, d_out):
        assert torch.equal(a, b.cpu())
    # a second run over the same buffers (slots are reused) gives the same bits
    for t in h_out:
        t.fill_(float("nan"))
    pipe.run(h_in, h_out)
    torch.cuda.synchronize()
    for a, b in zip(h_out, d_out):
        assert torch.equal(a, b.cpu())


def test_lsq_and_accumulate_host_helpers():
    data = orc.lsq_data(100_003, seed=4)
    vals = (0.3, 1.2, -0.4, 0.1)
    grad, loss = host_api.lsq_grad_host(torch.from_numpy(data).pin_memory(), vals, DEV)
    want_g, want_l = orc.lsq_grad(data, vals)
    assert np.allclose(grad.numpy(), want_g, rtol=1e-10) and abs(loss.item() - want_l) <= 1e-10 * want_l
    idx, val = orc.accumulate_inputs(200_000, 1024, "zipf", seed=8)
    got = host_api.accumulate_host(torch.from_numpy(idx).pin_memory(), torch.from_numpy(val).pin_memory(), 1024, DEV)
    exact = orc.accumulate_exact(idx, val, 1024)
    assert (np.abs(got.numpy() - exact) <= 1e-4 * orc.accumulate_exact(idx, np.abs(val), 1024) + 1e-30).all()


def test_splat_iteration_host_equals_the_device_call():
    W, H, N = 96, 64, 200
    params, target = orc.splat_scene(N, W, H, seed=5)
    tt = torch.from_numpy(target).to(DEV)
    out = torch.zeros((W * H, 3), device=DEV)
    loss, grads = host_api.splat_iteration_host(torch.from_numpy(params).pin_memory(), tt, out, W, H, DEV,
                                                flags=x.FLAG_DETERMINISTIC)
    g2 = torch.zeros((N, 9), device=DEV)
    o2 = torch.zeros((W * H, 3), device=DEV)
    l2 = torch.zeros(1, device=DEV)
    x.launch_gaussian_splatting(torch.from_numpy(params).to(DEV), g2, tt, o2, l2, W, H, N, x.FLAG_DETERMINISTIC)
    torch.cuda.synchronize()
    assert torch.equal(grads, g2.cpu()) and torch.equal(loss, l2.cpu()) and torch.equal(out, o2)


def test_splat_host_iteration_class_equals_the_device_call():
    """host_api.SplatHostIteration (pinned params in, pinned loss + gradients out, workspace launch): what bench.py times
    as the splat e2e."""
    W, H, N = 96, 64, 200
    params, target = orc.splat_scene(N, W, H, seed=5)
    tt = torch.from_numpy(target).to(DEV)
    it = host_api.SplatHostIteration(N, W, H, tt, 100 * N, DEV, flags=x.FLAG_DETERMINISTIC)
    ph = torch.from_numpy(params).pin_memory()
    g = torch.zeros((N, 9), device=DEV); o = torch.zeros((W * H, 3), device=DEV); l = torch.zeros(1, device=DEV)
    x.launch_gaussian_splatting(torch.from_numpy(params).to(DEV), g, tt, o, l, W, H, N, x.FLAG_DETERMINISTIC)
    for _ in range(2):
        loss, grads = it.run(ph)
        assert loss.item() == l.item() and torch.equal(grads, g.cpu()) and torch.equal(it.output, o)
    assert it.h2d_bytes == N * 36 and it.d2h_bytes == N * 36 + 4
