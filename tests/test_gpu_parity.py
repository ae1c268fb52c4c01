"""GPU parity tests (`-m gpu`): the CUDA path, called through the C ABI (ctypes bindings in
xyz-autodiff-cuda_b200/__init__.py), against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * integer / index work (tile rectangles, sorted tile lists, per-tile ranges): bit-exact;
  * fp32 values and per-element gradients: 1e-5 relative per element;
  * accumulated sums: 1e-4 relative -- stated against the sum of |terms| (a cancelling sum cannot be
    more accurate than that in any summation order), and bit-exact run to run in deterministic mode;
  * fp64 least squares: 1e-10 relative (reference tests use 1e-10, tests/test_parallel_gradient_accumulation.cu:94).
"""
import os

import numpy as np
import pytest
import torch

import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def rel_err(got, want, scale=None):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    scale = np.abs(want) if scale is None else scale
    return np.abs(got - want) / (scale + 1e-300)


# ---------------------------------------------------------------------------------------------------
# C3 covariance projection
# ---------------------------------------------------------------------------------------------------
def run_covproj(J, W, S, g):
    n = J.shape[0]
    outs = [torch.full((n, k), float("nan"), dtype=torch.float32, device=DEV) for k in (3, 6, 9, 6)]
    x.covproj_fwd_bwd(dev(J), dev(W), dev(S), dev(g), *outs)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in outs]


@pytest.mark.parametrize("n", [1, 3, 127, 128, 129, 1000, 128 * 148 * 3 * 2 + 77, 300_001])
def test_covproj_matches_fp64_oracle(n):
    J, W, S, g = orc.covproj_inputs(n, seed=n)
    got = run_covproj(J, W, S, g)
    want = orc.covproj(J, W, S, g, np.float64)
    for a, b, name in zip(got, want, ("out", "gJ", "gW", "gS")):
        # per-element 1e-5 relative to the row's magnitude (rows are sums of O(10) products)
        scale = np.abs(b).max(axis=1, keepdims=True)
        assert rel_err(a, b, scale).max() < 1e-5, name


def test_covproj_empty_and_unaligned():
    z = torch.empty((0, 6), dtype=torch.float32, device=DEV)
    x.covproj_fwd_bwd(z, torch.empty((0, 9), device=DEV), z, torch.empty((0, 3), device=DEV),
                      torch.empty((0, 3), device=DEV), z, torch.empty((0, 9), device=DEV), z)
    # base pointers that are only 4-byte aligned take the plain-load path
    n = 1000
    J, W, S, g = orc.covproj_inputs(n, seed=5)
    bufs = []
    for a in (J, W, S, g):
        t = torch.zeros(a.size + 1, dtype=torch.float32, device=DEV)
        t[1:] = dev(a).flatten()
        bufs.append(t[1:].view(a.shape))
    outs = [torch.zeros(n * k + 1, dtype=torch.float32, device=DEV)[1:].view(n, k) for k in (3, 6, 9, 6)]
    x.covproj_fwd_bwd(*bufs, *outs)
    want = orc.covproj(J, W, S, g, np.float64)
    for a, b in zip(outs, want):
        assert rel_err(a.cpu().numpy(), b, np.abs(b).max(axis=1, keepdims=True)).max() < 1e-5


def test_covproj_full_size_properties():
    """2^24 elements (1/4 of BASELINE's 2^26; same kernel, > L2): size-independent properties --
    the output is linear in S and in g, and the TMA path equals the plain path bit for bit."""
    n = 1 << 24
    gen = torch.Generator(device=DEV).manual_seed(7)
    J = torch.rand((n, 6), device=DEV, generator=gen) * 2 - 1
    W = torch.rand((n, 9), device=DEV, generator=gen) * 2 - 1
    S = torch.rand((n, 6), device=DEV, generator=gen) * 2 - 1
    g = torch.rand((n, 3), device=DEV, generator=gen) * 2 - 1
    o1 = [torch.empty((n, k), device=DEV) for k in (3, 6, 9, 6)]
    o2 = [torch.empty((n, k), device=DEV) for k in (3, 6, 9, 6)]
    x.covproj_fwd_bwd(J, W, S, g, *o1)
    x.covproj_fwd_bwd(J, W, S * 2, g * 4, *o2)  # exact power-of-two scalings
    assert torch.equal(o2[0], o1[0] * 2)          # out  ~ S
    assert torch.equal(o2[1], o1[1] * 8)          # gJ   ~ S g
    assert torch.equal(o2[2], o1[2] * 8)          # gW   ~ S g
    assert torch.equal(o2[3], o1[3] * 4)          # gS   ~ g
    # plain path on a shifted (unaligned) copy of a slice
    m = 100_000
    sl = [torch.zeros(m * k + 1, device=DEV)[1:].view(m, k) for k in (6, 9, 6, 3)]
    for d, s in zip(sl, (J, W, S, g)):
        d.copy_(s[:m])
    o3 = [torch.zeros(m * k + 1, device=DEV)[1:].view(m, k) for k in (3, 6, 9, 6)]
    x.covproj_fwd_bwd(*sl, *o3)
    for a, b in zip(o3, o1):
        assert torch.equal(a, b[:m])


def run_covproj_shared_w(J, W9, S, g, gW0=None):
    n = J.shape[0]
    out, gJ, gS = [torch.full((n, k), float("nan"), dtype=torch.float32, device=DEV) for k in (3, 6, 6)]
    gW = torch.zeros(9, dtype=torch.float32, device=DEV) if gW0 is None else dev(np.asarray(gW0, np.float32))
    x.covproj_shared_w_fwd_bwd(dev(J), dev(W9), dev(S), dev(g), out, gJ, gW, gS)
    torch.cuda.synchronize()
    return out.cpu().numpy(), gJ.cpu().numpy(), gW.cpu().numpy(), gS.cpu().numpy()


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 128 * 148 * 4 * 2 + 77, 300_001])
def test_covproj_shared_w_matches_fp64_oracle(n):
    """Variant B: one shared W.  Oracle = the reference's own matmul graph evaluated per element with W replicated
    (oracle.covproj, fp64); the shared gradient is the fp64 sum of its per-element adjoints.  Per-element outputs
    1e-5 relative; the accumulated gW 1e-4 relative to the sum of |terms| (BASELINE tolerance for accumulated sums)."""
    J, W, S, g = orc.covproj_inputs(n, seed=n + 1)
    W9 = W[0].copy()
    Wrep = np.broadcast_to(W9, (n, 9)).copy()
    out, gJ, gW, gS = run_covproj_shared_w(J, W9, S, g)
    w_out, w_gJ, w_gW, w_gS = orc.covproj(J, Wrep, S, g, np.float64)
    for a, b, name in ((out, w_out, "out"), (gJ, w_gJ, "gJ"), (gS, w_gS, "gS")):
        assert rel_err(a, b, np.abs(b).max(axis=1, keepdims=True)).max() < 1e-5, name
    assert (np.abs(gW - w_gW.sum(0)) <= 1e-4 * np.abs(w_gW).sum(0) + 1e-30).all()
    # per-element results are bit-identical to the per-element-W kernel fed the replicated W
    o2 = run_covproj(J, Wrep, S, g)
    assert np.array_equal(out, o2[0]) and np.array_equal(gJ, o2[1]) and np.array_equal(gS, o2[3])


def test_covproj_shared_w_accumulates_deterministically_and_unaligned():
    n = 70_003
    J, W, S, g = orc.covproj_inputs(n, seed=12)
    W9 = W[3].copy()
    a = run_covproj_shared_w(J, W9, S, g)
    b = run_covproj_shared_w(J, W9, S, g)
    assert np.array_equal(a[2], b[2])                      # fixed-order reduction: bit-identical run to run
    c = run_covproj_shared_w(J, W9, S, g, gW0=np.arange(9) * 10.0)
    assert np.allclose(c[2] - np.arange(9) * 10.0, a[2], rtol=0, atol=1e-3 * np.abs(a[2]).max())  # adds into the caller's values
    # unaligned bases: the plain-load path of the same kernel
    bufs = []
    for arr in (J, S, g):
        t = torch.zeros(arr.size + 1, dtype=torch.float32, device=DEV)
        t[1:] = dev(arr).flatten()
        bufs.append(t[1:].view(arr.shape))
    out, gJ, gS = [torch.zeros(n * k + 1, dtype=torch.float32, device=DEV)[1:].view(n, k) for k in (3, 6, 6)]
    gW = torch.zeros(9, device=DEV)
    x.covproj_shared_w_fwd_bwd(bufs[0], dev(W9), bufs[1], bufs[2], out, gJ, gW, gS)
    assert np.array_equal(out.cpu().numpy(), a[0]) and np.array_equal(gJ.cpu().numpy(), a[1]) and np.array_equal(gS.cpu().numpy(), a[3])
    Wrep = np.broadcast_to(W9, (n, 9)).copy()
    w_gW = orc.covproj(J, Wrep, S, g, np.float64)[2]
    assert (np.abs(gW.cpu().numpy() - w_gW.sum(0)) <= 1e-4 * np.abs(w_gW).sum(0)).all()
    # empty batch: nothing happens, gW untouched
    z6, z3 = torch.empty((0, 6), device=DEV), torch.empty((0, 3), device=DEV)
    gW = torch.full((9,), 1.5, device=DEV)
    x.covproj_shared_w_fwd_bwd(z6, dev(W9), z6, z3, z3, z6, gW, z6)
    assert (gW.cpu().numpy() == 1.5).all()


def test_covproj_shared_w_full_size_properties():
    """2^24 elements: out, gJ, gS equal the per-element-W kernel; gW is linear in g (exact power-of-two scaling) and
    equals the fp64 sum of the per-element kernel's gW within the accumulated-sum tolerance."""
    n = 1 << 24
    gen = torch.Generator(device=DEV).manual_seed(8)
    J = torch.rand((n, 6), device=DEV, generator=gen) * 2 - 1
    S = torch.rand((n, 6), device=DEV, generator=gen) * 2 - 1
    g = torch.rand((n, 3), device=DEV, generator=gen) * 2 - 1
    W9 = torch.rand(9, device=DEV, generator=gen) * 2 - 1
    out, gJ, gS = [torch.empty((n, k), device=DEV) for k in (3, 6, 6)]
    gW = torch.zeros(9, device=DEV)
    x.covproj_shared_w_fwd_bwd(J, W9, S, g, out, gJ, gW, gS)
    gW4 = torch.zeros(9, device=DEV)
    x.covproj_shared_w_fwd_bwd(J, W9, S, g * 4, out, gJ, gW4, gS)
    assert torch.equal(gW4, gW * 4)
    o = [torch.empty((n, k), device=DEV) for k in (3, 6, 9, 6)]
    x.covproj_fwd_bwd(J, W9.expand(n, 9).contiguous(), S, g * 4, *o)
    assert torch.equal(o[0], out) and torch.equal(o[1], gJ) and torch.equal(o[3], gS)
    ref = o[2].double().sum(0)
    assert ((gW4.double() - ref).abs() <= 1e-4 * o[2].double().abs().sum(0)).all()


# ---------------------------------------------------------------------------------------------------
# C1 least squares
# ---------------------------------------------------------------------------------------------------
def run_lsq(data, values, flags=0, grad0=None):
    prm = torch.zeros(8, dtype=torch.float64, device=DEV)
    prm[:4] = dev(np.asarray(values, np.float64))
    if grad0 is not None:
        prm[4:] = dev(np.asarray(grad0, np.float64))
    ls = torch.zeros(1, dtype=torch.float64, device=DEV)
    x.lsq_grad(dev(data), prm, ls, flags)
    torch.cuda.synchronize()
    return prm[4:].cpu().numpy(), ls.item()


LSQ_FIXED = [  # reference examples/optimization/tests/test_linear_regression_gradient.cu:161-216
    ((2.0, 3.0, 5.0), (1.0, 1.5, 0.5, 0.2)),
    ((1.0, 1.0, 2.0), (0.0, 0.0, 0.0, 0.0)),
    ((-2.0, -1.5, 3.0), (-1.0, 2.0, -0.5, -0.3)),
    ((100.0, 150.0, 500.0), (50.0, 30.0, 70.0, 20.0)),
    ((1e-3, 2e-3, 1e-3), (1e-4, 2e-4, 1e-4, 1e-5)),
]


@pytest.mark.parametrize("pt,prm", LSQ_FIXED)
def test_lsq_fixed_cases_closed_form(pt, prm):
    x1, x2, y = pt
    a, b, c, d = prm
    r = (a - x1) ** 2 + b * (c - x2) ** 2 + d - y
    want = 2 * r * np.array([2 * (a - x1), (c - x2) ** 2, 2 * b * (c - x2), 1.0])
    got, loss = run_lsq(np.array([pt]), prm)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-300)
    assert np.isclose(loss, r * r, rtol=1e-12)


def test_lsq_reference_batch_of_32():
    # test_linear_regression_gradient.cu:218-286: i = 0..31: (0.1 i, 0.2 i, 0.3 i + 1), params (.5, 1, .3, .1)
    i = np.arange(32, dtype=np.float64)
    data = np.stack([0.1 * i, 0.2 * i, 0.3 * i + 1.0], -1)
    got, _ = run_lsq(data, (0.5, 1.0, 0.3, 0.1))
    want, _ = orc.lsq_grad(data, (0.5, 1.0, 0.3, 0.1))
    assert np.allclose(got, want, rtol=1e-12)


@pytest.mark.parametrize("n", [1, 255, 256, 257, 1000, 1_000_000])
@pytest.mark.parametrize("residual_only", [False, True, 2])
def test_lsq_matches_oracle(n, residual_only):
    data = orc.lsq_data(n, seed=42)
    flags = {False: 0, True: x.FLAG_RESIDUAL_ONLY, 2: x.FLAG_LSQ_SHIPPED_GRAPH}[residual_only]
    got, loss = run_lsq(data, (0.0, 1.0, 0.0, 0.0), flags)
    want, wloss = orc.lsq_grad(data, (0.0, 1.0, 0.0, 0.0), residual_only, threads=8)
    # sums of up to 1e6 terms in a different (tree) order: 1e-10 relative to the sum of |terms| ~ |want|
    assert rel_err(got, want).max() < 1e-10
    assert abs(loss - wloss) <= 1e-10 * abs(wloss)


def test_lsq_accumulates_and_is_deterministic():
    data = orc.lsq_data(200_003, seed=1)
    g1, _ = run_lsq(data, (0.3, 1.2, -0.4, 0.1))
    g2, _ = run_lsq(data, (0.3, 1.2, -0.4, 0.1))
    assert np.array_equal(g1, g2)  # no floating-point atomics: bit-identical run to run
    g3, _ = run_lsq(data, (0.3, 1.2, -0.4, 0.1), grad0=(1.0, 2.0, 3.0, 4.0))
    assert np.allclose(g3, g1 + np.array([1.0, 2.0, 3.0, 4.0]), rtol=1e-14)
    # unaligned base pointer (8-byte aligned only) takes the plain-load path
    t = torch.zeros(data.size + 1, dtype=torch.float64, device=DEV)
    t[1:] = dev(data).flatten()
    prm = torch.zeros(8, dtype=torch.float64, device=DEV)
    prm[:4] = dev(np.array([0.3, 1.2, -0.4, 0.1]))
    x.lsq_grad(t[1:].view(-1, 3), prm)
    assert rel_err(prm[4:].cpu().numpy(), g1).max() < 1e-12


def test_lsq_sgd_update_and_select_batch():
    prm = dev(np.array([1.0, 2.0, 3.0, 4.0, 10.0, 20.0, 30.0, 40.0]))
    x.lsq_sgd_update(prm, 0.1, 8)
    want = np.array([1.0, 2.0, 3.0, 4.0]) - 0.1 * np.array([10.0, 20.0, 30.0, 40.0]) / 8
    assert np.allclose(prm[:4].cpu().numpy(), want, rtol=1e-15)
    data = orc.lsq_data(10_000, seed=3)
    batch = torch.zeros((8192, 3), dtype=torch.float64, device=DEV)
    x.lsq_select_batch(dev(data), batch, seed=12345, epoch=7)
    assert np.array_equal(batch.cpu().numpy(), orc.lsq_select_batch(data, 8192, 12345, 7))  # index work: bit-exact


def test_lsq_fused_sgd_step_equals_the_four_call_epoch():
    """xyz_lsq_sgd_step_f64 == select_batch + clear + lsq_grad + sgd_update (the reference's epoch body,
    linear_regression_sgd.cu:185-209), same sampling; sums differ only by fp64 reassociation."""
    data = dev(orc.lsq_data(100_000, seed=3))
    for flags in (0, x.FLAG_RESIDUAL_ONLY):
        pa = torch.zeros(8, dtype=torch.float64, device=DEV)
        pa[1] = 1.0
        pb = pa.clone()
        batch = torch.empty((8192, 3), dtype=torch.float64, device=DEV)
        la = torch.zeros(1, dtype=torch.float64, device=DEV)
        lb = torch.zeros(1, dtype=torch.float64, device=DEV)
        for epoch in range(5):
            x.lsq_select_batch(data, batch, 42, epoch)
            pa[4:] = 0.0
            x.lsq_grad(batch, pa, la, flags)
            x.lsq_sgd_update(pa, 1e-4, 8192)
            x.lsq_sgd_step(data, pb, 8192, 42, epoch, 1e-4, lb, flags)
            torch.cuda.synchronize()
            a, b = pa.cpu().numpy(), pb.cpu().numpy()
            assert np.allclose(a[4:], b[4:], rtol=1e-12, atol=1e-9), (epoch, a, b)
            assert np.allclose(a[:4], b[:4], rtol=1e-12, atol=1e-14)
        assert abs(la.item() - lb.item()) <= 1e-12 * abs(la.item())
        # deterministic: a second run from the same state is bit-identical
        pc = torch.zeros(8, dtype=torch.float64, device=DEV)
        pc[1] = 1.0
        for epoch in range(5):
            x.lsq_sgd_step(data, pc, 8192, 42, epoch, 1e-4, None, flags)
        torch.cuda.synchronize()
        assert torch.equal(pc, pb)


def test_lsq_many_epochs_in_one_launch_equal_single_epoch_calls():
    """xyz_lsq_sgd_run_f64 (cooperative kernel, parameters in registers, one grid barrier per epoch) against the same
    epochs run one call at a time: parameters, last gradient and accumulated loss bit for bit; several batch sizes
    (1 CTA, partial CTA, the driver's 8192, more CTAs than SMs), both losses, a second chunk continuing the first."""
    data = dev(orc.lsq_data(100_000, seed=9))
    for batch, flags in ((1, 0), (100, 0), (8192, 0), (8192, x.FLAG_RESIDUAL_ONLY), (70_000, 0)):
        epochs = 37
        lrs = [1e-4 * np.exp(-0.01 * e) for e in range(2 * epochs)]
        p1 = torch.zeros(8, dtype=torch.float64, device=DEV); p1[1] = 1.0
        l1 = torch.zeros(1, dtype=torch.float64, device=DEV)
        for e in range(2 * epochs):
            x.lsq_sgd_step(data, p1, batch, 42, e, lrs[e], l1, flags)
        p2 = torch.zeros(8, dtype=torch.float64, device=DEV); p2[1] = 1.0
        l2 = torch.zeros(1, dtype=torch.float64, device=DEV)
        x.lsq_sgd_run(data, p2, batch, 42, 0, lrs[:epochs], l2, flags)
        x.lsq_sgd_run(data, p2, batch, 42, epochs, lrs[epochs:], l2, flags)
        torch.cuda.synchronize()
        assert torch.equal(p1, p2), (batch, flags, p1, p2)
        assert torch.equal(l1, l2), (batch, flags)
    # zero epochs: nothing changes
    p3 = p2.clone()
    x.lsq_sgd_run(data, p3, 8192, 42, 0, [])
    assert torch.equal(p3, p2)


# ---------------------------------------------------------------------------------------------------
# C2 accumulation
# ---------------------------------------------------------------------------------------------------
def run_acc(idx, val, k, flags=0, dtype=torch.float32):
    grad = torch.zeros(k, dtype=dtype, device=DEV)
    x.accumulate(dev(idx) if idx is not None else None, dev(val), grad, flags)
    torch.cuda.synchronize()
    return grad.cpu().numpy()


@pytest.mark.parametrize("dist", ["uniform", "zipf", "same"])
@pytest.mark.parametrize("n", [1, 31, 128, 1000, 1 << 20])
@pytest.mark.parametrize("flags", [0, 1])
def test_accumulate_matches_exact_sums(dist, n, flags):
    k = 1024
    idx, val = orc.accumulate_inputs(n, k, dist, seed=n)
    got = run_acc(idx, val, k, flags)
    exact = orc.accumulate_exact(idx, val, k)
    abs_sum = orc.accumulate_exact(idx, np.abs(val), k)
    assert (np.abs(got - exact) <= 1e-4 * abs_sum + 1e-30).all()
    # and against the reference's own sequential fp32 order, same bar
    seq = orc.accumulate(idx, val, k)
    assert (np.abs(got - seq) <= 1e-4 * abs_sum + 1e-30).all()


def test_accumulate_reference_known_answers():
    # tests/test_parallel_gradient_accumulation.cu:61-110: 10 000 threads add 1.0, 1.0 and 3 + 0.002 tid
    n = 10_000
    tid = np.arange(n)
    idx = np.tile(np.array([0, 1, 2], np.int32), n)
    val = np.stack([np.ones(n), np.ones(n), (1.0 + tid * 0.001) + (2.0 + tid * 0.001)], -1).reshape(-1)
    got = run_acc(idx, val, 3, dtype=torch.float64)
    assert abs(got[0] - 10000.0) < 1e-10 and abs(got[1] - 10000.0) < 1e-10
    assert abs(got[2] - val[2::3].sum()) < 1e-6
    # :122-167: 100 000 threads, exact
    n = 100_000
    got = run_acc(np.zeros(n, np.int32), np.ones(n, np.float32), 1)
    assert got[0] == 100000.0
    # C2 "w" pattern at full size: 2^24 ones into one of 1024 bins is exact in fp32 (2^24 is representable)
    n = 1 << 24
    got = run_acc(np.full(n, 7, np.int32), np.ones(n, np.float32), 1024)
    assert got[7] == float(n) and got.sum() == float(n)


def test_accumulate_deterministic_is_bit_identical_and_variants():
    k = 1024
    idx, val = orc.accumulate_inputs(3_000_017, k, "zipf", seed=9)
    a = run_acc(idx, val, k, x.FLAG_DETERMINISTIC)
    b = run_acc(idx, val, k, x.FLAG_DETERMINISTIC)
    assert np.array_equal(a, b)
    # implicit ids (id = i mod K)
    got = run_acc(None, val, k)
    exact = orc.accumulate_exact(None, val, k)
    assert (np.abs(got - exact) <= 1e-4 * orc.accumulate_exact(None, np.abs(val), k)).all()
    # out-of-range ids are ignored; unaligned bases; K not a multiple of 4; large K (global-RED path)
    idx2 = idx.copy()
    idx2[::1000] = -5
    idx2[1::1000] = k + 3
    keep = (idx2 >= 0) & (idx2 < k)
    got = run_acc(idx2, val, k)
    exact = orc.accumulate_exact(idx2[keep], val[keep], k)
    assert (np.abs(got - exact) <= 1e-4 * orc.accumulate_exact(idx2[keep], np.abs(val[keep]), k) + 1e-30).all()
    ti = torch.zeros(idx.size + 1, dtype=torch.int32, device=DEV)
    ti[1:] = dev(idx)
    tv = torch.zeros(val.size + 1, dtype=torch.float32, device=DEV)
    tv[1:] = dev(val)
    grad = torch.zeros(k, device=DEV)
    x.accumulate(ti[1:], tv[1:], grad)
    assert (np.abs(grad.cpu().numpy() - orc.accumulate_exact(idx, val, k)) <= 1e-4 * orc.accumulate_exact(idx, np.abs(val), k)).all()
    # K = 1, 3 (32 warps), 1001 (odd K), 5000 (8-warp tables), 30000 (match.any tables), 100000 (global REDs)
    for kk in (1, 3, 1001, 5000, 30_000, 100_000):
        i3, v3 = orc.accumulate_inputs(200_000, kk, "uniform", seed=kk)
        got = run_acc(i3, v3, kk)
        assert (np.abs(got - orc.accumulate_exact(i3, v3, kk)) <= 1e-4 * orc.accumulate_exact(i3, np.abs(v3), kk) + 1e-30).all()
    # fp64 values through the tagged-table kernel (16 warps at K = 1024), all three id distributions
    for dist in ("uniform", "zipf", "same"):
        i4, v4 = orc.accumulate_inputs(1 << 19, k, dist, seed=11)
        v4 = v4.astype(np.float64) * (1.0 + 1e-9)
        got = run_acc(i4, v4, k, dtype=torch.float64)
        exact = np.zeros(k)
        np.add.at(exact, i4, v4)
        abs_sum = np.zeros(k)
        np.add.at(abs_sum, i4, np.abs(v4))
        assert (np.abs(got - exact) <= 1e-12 * abs_sum + 1e-300).all()
    # accumulates into the caller's values
    grad = torch.ones(k, device=DEV)
    x.accumulate(dev(idx), dev(val), grad)
    assert np.allclose(grad.cpu().numpy() - 1.0, orc.accumulate_exact(idx, val, k), atol=1e-2)


def test_accumulate_random_shapes():
    """Seeded sweep over (n, K, id distribution, flags, invalid ids, alignment): every dispatch path of xyz_accumulate_f32
    against the exact fp64 sums, tolerance 1e-4 of the sum of |terms| (BASELINE's bound for accumulated sums)."""
    rng = np.random.default_rng(20260)
    for case in range(40):
        k = int(rng.choice([1, 2, 5, 33, 128, 777, 1024, 1200, 1771, 1772, 2500, 6000]))
        n = int(rng.choice([1, 63, 4097, 65_535, 65_536, 65_537, 200_000, 524_288 + 3, 1_000_001]))
        dist = str(rng.choice(["uniform", "zipf", "same"]))
        flags = int(rng.choice([0, x.FLAG_DETERMINISTIC]))
        if k > 4096 and flags:
            flags = 0
        idx, val = orc.accumulate_inputs(n, k, dist, seed=1000 + case)
        if rng.random() < 0.3 and n > 10:
            bad = rng.integers(0, n, size=max(1, n // 50))
            idx[bad] = rng.choice([-1, -7, k, k + 100, 2**31 - 1], size=bad.size)
        keep = (idx >= 0) & (idx < k)
        shift = int(rng.random() < 0.25)
        ti = torch.zeros(n + 1, dtype=torch.int32, device=DEV)
        tv = torch.zeros(n + 1, dtype=torch.float32, device=DEV)
        ti[shift:shift + n] = dev(idx)
        tv[shift:shift + n] = dev(val)
        grad = torch.zeros(k, device=DEV)
        x.accumulate(ti[shift:shift + n], tv[shift:shift + n], grad, flags)
        got = grad.cpu().numpy()
        exact = orc.accumulate_exact(idx[keep], val[keep], k)
        abs_sum = orc.accumulate_exact(idx[keep], np.abs(val[keep]), k)
        assert (np.abs(got - exact) <= 1e-4 * abs_sum + 1e-30).all(), (case, k, n, dist, flags, shift)


def test_accumulate_striped_tables_variants():
    """The lane-striped fast path (fp32, aligned, n >= 2^16): every table count T (12, 6, 4, 3, 2 tables per SM), partial
    last unit, implicit ids with K below and above the 128-element row, invalid ids, all-invalid input, and bit-identical
    deterministic results for every id distribution."""
    def tol_ok(got, idx, val, k):
        exact = orc.accumulate_exact(idx, val, k)
        abs_sum = orc.accumulate_exact(idx, np.abs(val), k)
        return (np.abs(got - exact) <= 1e-4 * abs_sum + 1e-30).all()
    for kk, n in ((1, 70_001), (7, 65_536), (200, 131_072 + 5), (300, 100_000), (500, 99_999), (800, 1 << 17), (1024, 600_000),
                  (1025, 1 << 17), (1500, 250_001), (1771, 1 << 17), (3000, 1 << 17), (3626, 70_000), (3627, 70_000), (9000, 80_000)):
        for dist in ("uniform", "zipf", "same"):
            i, v = orc.accumulate_inputs(n, kk, dist, seed=kk + n)
            assert tol_ok(run_acc(i, v, kk), i, v, kk), (kk, n, dist)
            a = run_acc(i, v, kk, x.FLAG_DETERMINISTIC)
            b = run_acc(i, v, kk, x.FLAG_DETERMINISTIC)
            assert np.array_equal(a, b) and tol_ok(a, i, v, kk), (kk, n, dist)
        _, v = orc.accumulate_inputs(n, kk, "uniform", seed=3)
        got = run_acc(None, v, kk)
        assert (np.abs(got - orc.accumulate_exact(None, v, kk)) <= 1e-4 * orc.accumulate_exact(None, np.abs(v), kk) + 1e-30).all(), kk
    # fp64 flavour (8 copies per bin, two half-warp update phases): every table count, partial units, implicit ids,
    # deterministic bit-identity; the sums are exact to 1e-12 of the sum of |terms|
    for kk, n in ((1, 70_001), (3, 65_536), (100, 131_072 + 5), (500, 80_000), (700, 99_999), (1024, 600_000), (1500, 250_001), (1770, 1 << 17), (2500, 1 << 17), (3624, 70_000),
                  (3625, 70_000), (6000, 100_000)):
        for dist in ("uniform", "zipf", "same"):
            i, v = orc.accumulate_inputs(n, kk, dist, seed=kk + n + 1)
            v64 = v.astype(np.float64) * (1.0 + 1e-9)
            exact = np.zeros(kk); np.add.at(exact, i, v64)
            abs_sum = np.zeros(kk); np.add.at(abs_sum, i, np.abs(v64))
            got = run_acc(i, v64, kk, dtype=torch.float64)
            assert (np.abs(got - exact) <= 1e-12 * abs_sum + 1e-300).all(), (kk, n, dist)
            a = run_acc(i, v64, kk, x.FLAG_DETERMINISTIC, dtype=torch.float64)
            b = run_acc(i, v64, kk, x.FLAG_DETERMINISTIC, dtype=torch.float64)
            assert np.array_equal(a, b) and (np.abs(a - exact) <= 1e-12 * abs_sum + 1e-300).all(), (kk, n, dist)
        got = run_acc(None, v64, kk, dtype=torch.float64)
        exact = np.zeros(kk); np.add.at(exact, np.arange(n) % kk, v64)
        assert (np.abs(got - exact) <= 1e-12 * np.abs(v64).sum() + 1e-300).all(), kk
    # a short input with a bin count beyond the small kernel's tables, fixed order requested: striped passes
    i, v = orc.accumulate_inputs(5000, 12_000, "uniform", seed=2)
    a = run_acc(i, v, 12_000, x.FLAG_DETERMINISTIC)
    assert np.array_equal(a, run_acc(i, v, 12_000, x.FLAG_DETERMINISTIC)) and tol_ok(a, i, v, 12_000)
    # nothing but invalid ids: grad stays untouched
    n, k = 1 << 17, 1024
    grad = torch.full((k,), 2.5, device=DEV)
    x.accumulate(torch.full((n,), -1, dtype=torch.int32, device=DEV), torch.ones(n, device=DEV), grad)
    assert (grad.cpu().numpy() == 2.5).all()
    grad = torch.full((k,), 2.5, device=DEV)
    x.accumulate(torch.full((n,), k, dtype=torch.int32, device=DEV), torch.ones(n, device=DEV), grad, x.FLAG_DETERMINISTIC)
    assert (grad.cpu().numpy() == 2.5).all()
    # integer-valued inputs: every partial sum is exact, so the result must equal the exact sum bit for bit
    i, _ = orc.accumulate_inputs(1 << 22, k, "zipf", seed=5)
    v = (np.arange(1 << 22) % 7 - 3).astype(np.float32)
    assert np.array_equal(run_acc(i, v, k).astype(np.float64), orc.accumulate_exact(i, v, k))


# ---------------------------------------------------------------------------------------------------
# C4 splat
# ---------------------------------------------------------------------------------------------------
def run_splat(params, target, W, H, flags=0, rows=None, grads0=None, loss0=0.0):
    N = params.shape[0]
    grads = torch.zeros((N, 9), dtype=torch.float32, device=DEV) if grads0 is None else dev(grads0)
    out = torch.full((W * H, 3), float("nan"), dtype=torch.float32, device=DEV)
    loss = torch.full((1,), loss0, dtype=torch.float32, device=DEV)
    x.launch_gaussian_splatting(dev(params) if N else torch.empty((0, 9), device=DEV), grads, dev(target), out, loss,
                                W, H, N, flags, rows=rows)
    torch.cuda.synchronize()
    return grads.cpu().numpy(), out.cpu().numpy(), loss.item()


def check_splat_against_fp64(params, target, W, H, got):
    """image: 1e-5 relative per element; loss and gradients (accumulated sums): 1e-4 relative to the sum of
    |terms| (+ the terms of pairs that sit on the L1 kink, whose sign no fp32 implementation can decide)."""
    rg, ro, rl, tol = orc.splat_tolerance(params, target, W, H)
    g, o, l = got
    assert (np.abs(o - ro) <= 1e-5 * np.maximum(np.abs(ro), np.abs(ro).max() * 1e-3)).all(), "image"
    assert abs(l - rl) <= 1e-4 * abs(rl), "loss"
    assert (np.abs(g - rg) <= tol).all(), "gradients vs sum|terms|"
    return rg, ro, rl, tol


@pytest.mark.parametrize("W,H,N,seed", [(64, 48, 50, 3), (37, 21, 7, 0), (16, 16, 1, 1), (100, 70, 300, 11)])
@pytest.mark.parametrize("flags", [0, 2])  # fast-math flavour, IEEE flavour
def test_splat_small_scenes_match_fp64_oracle(W, H, N, seed, flags):
    params, target = orc.splat_scene(N, W, H, seed=seed)
    got = run_splat(params, target, W, H, flags)
    check_splat_against_fp64(params, target, W, H, got)


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref/libxyz_ref.so (reference built for the host) not present")
@pytest.mark.parametrize("case", range(12))
def test_splat_ill_conditioned_scenes_within_the_fp32_bound(case):
    """Sub-pixel to image-sized, arbitrarily rotated Gaussians, odd image sizes, centres outside the image: every flag
    combination of the CUDA path against fp64 under the fp32 conditioning bound that the reference's own fp32 kernel
    is shown to need (tests/test_oracle.py::test_fp32_conditioning_bound_holds_for_the_reference_kernel)."""
    params, target, W, H = orc.splat_hard_scene(case)
    N = params.shape[0]
    rg, ro, rl, tol_g, tol_i = orc.splat_tolerance_fp32(params, target, W, H)
    for flags in (0, x.FLAG_PRECISE_MATH, x.FLAG_DETERMINISTIC, x.FLAG_NO_CULL, x.FLAG_PRECISE_MATH | x.FLAG_DETERMINISTIC):
        g, o, l = run_splat(params, target, W, H, flags)
        assert (np.abs(o - ro) <= tol_i).all(), (case, flags, "image")
        assert abs(l - rl) <= 1e-4 * abs(rl) + tol_i.sum(), (case, flags, "loss")
        assert (np.abs(g - rg) <= tol_g).all(), (case, flags, "gradients")


def test_splat_matches_reference_kernel_on_host():
    """Against the reference's OWN kernel body compiled for the host (fp32, IEEE): the GPU's IEEE flavour differs
    only by FMA contraction and summation order."""
    W, H, N = 64, 48, 50
    params, target = orc.splat_scene(N, W, H, seed=3)
    g, o, l = run_splat(params, target, W, H, x.FLAG_PRECISE_MATH)
    rg, ro, rl, _ = orc.splat(params, target, W, H, np.float32, which="ref")
    tol = orc.splat_tolerance(params, target, W, H)[3]
    assert (np.abs(o - ro) <= 1e-5 * np.maximum(np.abs(ro), np.abs(ro).max() * 1e-3)).all()
    assert abs(l - rl) <= 1e-4 * abs(rl)
    assert (np.abs(g - rg) <= tol).all()


def test_splat_cull_is_result_preserving_and_deterministic():
    """Skipping exactly-zero pairs must not change a single bit of the image (same summation order);
    deterministic mode is bit-identical run to run and equals the atomic mode to tolerance."""
    W, H, N = 160, 128, 400
    params, target = orc.splat_scene(N, W, H, seed=21)
    for flags in (0, x.FLAG_PRECISE_MATH):
        g0, o0, l0 = run_splat(params, target, W, H, flags)
        g1, o1, l1 = run_splat(params, target, W, H, flags | x.FLAG_NO_CULL)
        assert np.array_equal(o0, o1)
        assert l0 == l1
        tol = orc.splat_tolerance(params, target, W, H)[3]
        assert (np.abs(g0 - g1) <= tol).all()
        d0 = run_splat(params, target, W, H, flags | x.FLAG_DETERMINISTIC)
        d1 = run_splat(params, target, W, H, flags | x.FLAG_DETERMINISTIC)
        assert np.array_equal(d0[0], d1[0]) and np.array_equal(d0[1], d1[1]) and d0[2] == d1[2]
        assert np.array_equal(d0[1], o0)
        assert (np.abs(d0[0] - g0) <= tol).all()


def test_splat_backward_cull_drops_only_terms_below_fp32_resolution():
    """The backward pass leaves out the list entries with d2 > 48 on all of their tile (weights below exp(-24)).  In
    deterministic mode every kept entry is computed by the same instructions with and without the cull, so the difference
    between the two IS the dropped terms: below 1e-8 of the sum of |terms| (the tail integral of the largest coefficient,
    d2 exp(-d2 / 2) beyond 48, is 1e-9 of its total) -- which can still move the rounding of an fp32 sum by its last bit,
    so one ulp of the gradient is allowed on top (the parity bar is 1e-4).  The image and the loss must not move at all,
    and the atomic mode must agree to the usual tolerance."""
    for (W, H, N, seed, small) in ((160, 128, 400, 21, True), (256, 192, 300, 5, False)):
        params, target = orc.splat_scene(N, W, H, seed=seed, small=small)
        rg, ro, rl, tol = orc.splat_tolerance(params, target, W, H)
        absg = (tol - 1e-30) / 1e-4   # an upper bound of sum|terms| per component (it also holds the kink terms)
        for flags in (0, x.FLAG_PRECISE_MATH):
            D, A = x.FLAG_DETERMINISTIC, x.FLAG_BWD_ALL_PAIRS
            g_cut, o_cut, l_cut = run_splat(params, target, W, H, flags | D)
            g_all, o_all, l_all = run_splat(params, target, W, H, flags | D | A)
            assert np.array_equal(o_cut, o_all) and l_cut == l_all
            assert (np.abs(g_cut - g_all) <= 1e-8 * absg + 2.0 ** -23 * np.abs(g_all)).all(), float((np.abs(g_cut - g_all) / absg).max())
            g_atomic = run_splat(params, target, W, H, flags)[0]
            g_atomic_all = run_splat(params, target, W, H, flags | A)[0]
            assert (np.abs(g_atomic - rg) <= tol).all() and (np.abs(g_atomic_all - rg) <= tol).all()
            assert (np.abs(g_atomic - g_all) <= tol).all()


def test_splat_backward_stats_count_the_work_items():
    """xyz_splat_last_backward_stats: with XYZ_FLAG_BWD_ALL_PAIRS every list entry is a work item; with the default cull
    fewer; every item stands for a whole tile of pairs; the classic and the workspace entry point agree."""
    W, H, N = 256, 192, 300
    params, target = orc.splat_scene(N, W, H, seed=5, small=False)
    run_splat(params, target, W, H, x.FLAG_BWD_ALL_PAIRS)
    entries = x.splat_last_stats()["entries"]
    b_all = x.splat_last_backward_stats()
    assert b_all["items"] == entries and b_all["pairs"] == 256 * entries and b_all["ctas"] >= (entries + 127) // 128
    run_splat(params, target, W, H, 0)
    b_cut = x.splat_last_backward_stats()
    assert 0 < b_cut["items"] < entries and b_cut["pairs"] == 256 * b_cut["items"]
    ws = x.SplatWorkspace(W, H, N, entries + 64, 0)
    run_splat_ws(ws, params, target, W, H)
    torch.cuda.synchronize()
    assert x.splat_last_backward_stats() == b_cut
    run_splat(params, target, W, H, x.FLAG_DETERMINISTIC)
    assert x.splat_last_backward_stats()["items"] == b_cut["items"]


def test_splat_tail_cull_stays_inside_the_stated_bound():
    """XYZ_FLAG_TAIL_CULL (opt-in) drops pairs with weight < exp(-28): the image moves by at most
    N * exp(-28) * max|sigmoid(opacity) * color| and the result still meets the fp64 tolerances."""
    W, H, N = 160, 128, 400
    params, target = orc.splat_scene(N, W, H, seed=21)
    g0, o0, l0 = run_splat(params, target, W, H, 0)
    e0 = x.splat_last_stats()["entries"]
    g1, o1, l1 = run_splat(params, target, W, H, x.FLAG_TAIL_CULL)
    e1 = x.splat_last_stats()["entries"]
    assert e1 < e0
    bound = N * np.exp(-28.0) * np.abs(params[:, 5:8]).max()
    assert np.abs(o1 - o0).max() <= bound + 1e-12
    check_splat_against_fp64(params, target, W, H, (g1, o1, l1))


def test_splat_tile_binning_is_bit_exact():
    """Integer work: tile rectangles, stably sorted tile lists and per-tile ranges equal the CPU restatement
    computed from the same per-Gaussian records."""
    W, H, N = 200, 150, 1000
    params, target = orc.splat_scene(N, W, H, seed=5)
    params[:5, 0] = [-500.0, 900.0, 100.0, 50.0, 0.0]        # far off-screen / on the border
    params[5, 2:4] = -30.0                                     # degenerate covariance (det regularised, Q15)
    params[6, 2:4] = 5.0                                       # huge Gaussian: covers the whole image
    R, D = x.FLAG_RADIX_BINNING, x.FLAG_DETERMINISTIC   # both binning paths, both payload modes
    for flags, d2max in ((0, 176.0), (x.FLAG_PRECISE_MATH, 209.0), (x.FLAG_NO_CULL, 176.0), (D, 176.0), (R, 176.0),
                         (R | x.FLAG_PRECISE_MATH, 209.0), (R | x.FLAG_NO_CULL, 176.0), (R | D, 176.0)):
        run_splat(params, target, W, H, flags)
        st = x.splat_last_stats()
        rects, ranges, ids, recs = x.splat_debug_binning(N, st["tiles"], st["entries"])
        orects, oranges, oids = orc.splat_binning(recs, W, H, d2max=d2max, no_cull=bool(flags & x.FLAG_NO_CULL))
        assert np.array_equal(rects, orects)
        assert np.array_equal(ranges, oranges)
        assert np.array_equal(ids, oids)
        assert st["entries"] == oids.size
        # the records themselves: IEEE ops on both sides; libm vs CUDA expf/sinf/cosf differ by <= 2 ulp, which
        # the covariance inverse amplifies by its condition number (a few units here)
        want = orc.splat_records(params)
        ok = np.isfinite(want).all(axis=1) & (np.abs(params[:, 2:4]) < 3).all(axis=1)
        scale = np.abs(want[:, 2:5]).max(axis=1, keepdims=True)       # ib cancels when the Gaussian is near-isotropic
        assert (np.abs(recs[ok] - want[ok]) <= 1e-5 * np.maximum(np.abs(want[ok]), scale[ok])).all()


def test_splat_row_band_prefilter_keeps_the_integer_results():
    """A launch that renders a row band skips, ahead of the records, the Gaussians that cannot reach the band (bound on the
    ellipse's reach from max(exp(scale))).  Rectangles, lists and ranges of the band must equal the CPU restatement run on
    the records of a FULL launch (the band launch does not write the records of the Gaussians it skips); degenerate,
    huge, off-screen and strongly anisotropic Gaussians take the full path."""
    W, H, N = 200, 400, 3000
    params, target = orc.splat_scene(N, W, H, seed=9)
    params[:5, 0] = [-500.0, 900.0, 100.0, 50.0, 0.0]
    params[5, 2:4] = -30.0                                     # degenerate covariance (det regularised, Q15): full path
    params[6, 2:4] = 5.0                                       # covers the whole image
    params[7, 2:4] = [-4.0, 3.0]                               # axes 1100 : 1, rotated: full path
    run_splat(params, target, W, H, 0)
    st = x.splat_last_stats()
    recs = x.splat_debug_binning(N, st["tiles"], st["entries"])[3]
    for rb, re_ in ((0, 64), (112, 240), (304, 400), (37, 41)):
        for flags in (0, x.FLAG_RADIX_BINNING, x.FLAG_DETERMINISTIC):
            run_splat(params, target, W, H, flags, rows=(rb, re_))
            st = x.splat_last_stats()
            rects, ranges, ids, _ = x.splat_debug_binning(N, st["tiles"], st["entries"])
            orects, oranges, oids = orc.splat_binning(recs, W, H, rb, re_)
            assert np.array_equal(rects, orects), (rb, re_, flags)
            assert np.array_equal(ids, oids) and st["entries"] == oids.size, (rb, re_, flags)
            nonempty = oranges[:, 1] > oranges[:, 0]
            assert np.array_equal(ranges[nonempty], oranges[nonempty]), (rb, re_, flags)
            assert (rects[:, 3] > rects[:, 1]).sum() < N // 2     # most Gaussians miss the band


def test_splat_counting_sort_binning_equals_the_radix_path():
    """The default binning (stable counting sort by tile, csrc/splat_host.cu section 2b) and the radix path give the
    same lists, so image and loss are bit-identical; deterministic gradients too (same rows, same order).  Scenes:
    many chunks of Gaussians, Gaussians taller than the 16 cached tile rows, a row band, empty tiles, a chunk with
    no entries at all."""
    R, D = x.FLAG_RADIX_BINNING, x.FLAG_DETERMINISTIC
    cases = []
    W, H, N = 512, 384, 40_000
    params, target = orc.splat_scene(N, W, H, seed=11)
    params[100:110, 2:4] = 4.0                                 # huge: every tile, more than 16 tile rows
    params[200:300, 0] = -4000.0                               # off screen: no entries
    cases.append((params, target, W, H, None))
    W2, H2, N2 = 300, 700, 3000
    p2, t2 = orc.splat_scene(N2, W2, H2, seed=12)
    p2[:, 0] = np.minimum(p2[:, 0], 120.0)                     # the right part of the image stays empty
    p2[::7, 2] = 3.0; p2[::7, 3] = 0.0; p2[::7, 4] = 0.3       # tall, thin, rotated
    cases.append((p2, t2, W2, H2, None))
    cases.append((p2, t2, W2, H2, (160, 400)))
    for params, target, W, H, band in cases:
        N = params.shape[0]
        for flags in (0, D, x.FLAG_PRECISE_MATH):
            res = []
            for extra in (0, R):
                out = run_splat(params, target, W, H, flags | extra, rows=band)
                st = x.splat_last_stats()
                rects, ranges, ids, recs = x.splat_debug_binning(N, st["tiles"], st["entries"])
                res.append((out, st["entries"], rects.copy(), ranges.copy(), ids.copy()))
            (ga, oa, la), ea, ra, rga, ida = res[0]
            (gb, ob, lb), eb, rb, rgb_, idb = res[1]
            assert ea == eb and np.array_equal(ra, rb) and np.array_equal(ida, idb)
            nonempty = rgb_[:, 1] > rgb_[:, 0]                 # the radix path leaves empty tiles at (0, 0)
            assert np.array_equal(rga[nonempty], rgb_[nonempty]) and (rga[~nonempty, 0] == rga[~nonempty, 1]).all()
            assert np.array_equal(oa, ob, equal_nan=True) and la == lb   # NaN = outside the row band, untouched
            if flags & D:
                assert np.array_equal(ga, gb)
            else:
                assert np.abs(ga - gb).max() <= 1e-4 * np.abs(gb).max()   # atomics: order differs run to run


def test_splat_counting_sort_beyond_8192_tiles_equals_the_radix_path():
    """Row bands of more than 8192 tiles (here 130 x 65 = 8450) keep the counting sort -- the kernels opt in to the SM's
    full shared memory -- and give the lists of the radix path bit for bit; the workspace entry point accepts them."""
    W, H, N = 2080, 1040, 20_000
    params, target = orc.splat_scene(N, W, H, seed=31, small=False)
    params[:50, 2:4] = 3.5                                      # a few very large Gaussians
    R, D = x.FLAG_RADIX_BINNING, x.FLAG_DETERMINISTIC
    res = []
    for extra in (0, R):
        out = run_splat(params, target, W, H, D | extra)
        st = x.splat_last_stats()
        assert st["tiles"] == 8450
        rects, ranges, ids, recs = x.splat_debug_binning(N, st["tiles"], st["entries"])
        res.append((out, st["entries"], rects.copy(), ranges.copy(), ids.copy()))
    (ga, oa, la), ea, ra, rga, ida = res[0]
    (gb, ob, lb), eb, rb, rgb_, idb = res[1]
    assert ea == eb and np.array_equal(ra, rb) and np.array_equal(ida, idb)
    nonempty = rgb_[:, 1] > rgb_[:, 0]
    assert np.array_equal(rga[nonempty], rgb_[nonempty])
    assert np.array_equal(oa, ob) and la == lb and np.array_equal(ga, gb)
    ws = x.SplatWorkspace(W, H, N, ea + 100, D)
    g, o, l = run_splat_ws(ws, params, target, W, H)
    torch.cuda.synchronize()
    assert np.array_equal(o.cpu().numpy(), oa) and l.item() == la and np.array_equal(g.cpu().numpy(), ga)
    x.shutdown()


def test_splat_async_launch_equals_the_synchronous_one_and_reports_overflow():
    """XYZ_FLAG_ASYNC: no host synchronisation from the second launch of a scene shape on; buffers sized for 1.5 x the
    last known list length.  A launch that fits gives the synchronous results; one that does not renders empty lists
    (memory-safe), the next call reports XYZ_ERR_WORKSPACE once, and the launch after that is sized afresh."""
    W, H, N = 256, 192, 5000
    params, target = orc.splat_scene(N, W, H, seed=21)
    A = x.FLAG_ASYNC
    g0, o0, l0 = run_splat(params, target, W, H)               # ordinary launch: the host learns the list length
    e0 = x.splat_last_stats()["entries"]
    for _ in range(3):
        g1, o1, l1 = run_splat(params, target, W, H, A)
    assert np.array_equal(o1, o0) and l1 == l0
    assert np.abs(g1 - g0).max() <= 1e-4 * np.abs(g0).max()    # atomics: order differs run to run
    assert x.splat_last_stats()["entries"] == e0               # fetched from the device on demand
    p2 = params.copy()
    p2[:, 0:2] += 0.37                                         # what a training step does: lists change a little
    g2, o2, l2 = run_splat(p2, target, W, H, A)
    g2s, o2s, l2s = run_splat(p2, target, W, H)
    assert np.array_equal(o2, o2s) and l2 == l2s
    run_splat(p2, target, W, H, A)
    p3 = params.copy()
    p3[:, 2:4] += 1.2                                          # every Gaussian 3.3 x larger: far more than 1.5 x the entries
    g3, o3, l3 = run_splat(p3, target, W, H, A)
    assert (o3 == 0).all() and (g3 == 0).all() and abs(l3 - np.abs(target).sum()) <= 1e-5 * np.abs(target).sum()
    with pytest.raises(RuntimeError):
        run_splat(p3, target, W, H, A)                         # reports the overflow, launches nothing
    g4, o4, l4 = run_splat(p3, target, W, H, A)                # sized from the reported length
    g4s, o4s, l4s = run_splat(p3, target, W, H)
    assert np.array_equal(o4, o4s) and l4 == l4s
    assert np.abs(g4 - g4s).max() <= 1e-4 * np.abs(g4s).max()


def test_splat_edge_cases():
    W, H = 50, 40
    target = orc.test_image(W, H)
    # no Gaussians: image 0, loss = sum |target|, nothing else touched
    g, o, l = run_splat(np.zeros((0, 9), np.float32), target, W, H)
    assert (o == 0).all() and abs(l - np.abs(target).sum()) <= 1e-5 * np.abs(target).sum()
    # loss and gradients ACCUMULATE into the caller's buffers (gaussian_splatting_training.cu:131-135)
    params, target = orc.splat_scene(20, W, H, seed=2)
    g0, o0, l0 = run_splat(params, target, W, H)
    g1, o1, l1 = run_splat(params, target, W, H, grads0=np.ones((20, 9), np.float32), loss0=5.0)
    assert np.allclose(g1 - 1.0, g0, rtol=1e-4, atol=1e-4) and abs((l1 - 5.0) - l0) <= 1e-4 * l0
    # every Gaussian off screen
    params[:, 0] = -1e4
    g, o, l = run_splat(params, target, W, H)
    assert (o == 0).all() and (g == 0).all()


def test_splat_row_bands_partition_the_image():
    """Sharding one image by row bands (multi-GPU C4): bands render disjoint rows of the same image and their
    losses / gradients add up to the full launch."""
    W, H, N = 96, 80, 200
    params, target = orc.splat_scene(N, W, H, seed=8)
    gf, of, lf = run_splat(params, target, W, H, x.FLAG_DETERMINISTIC)
    gsum = np.zeros_like(gf)
    lsum = 0.0
    img = np.zeros_like(of)
    for r0, r1 in ((0, 24), (24, 50), (50, 80)):   # deliberately not multiples of 16
        g, o, l = run_splat(params, target, W, H, x.FLAG_DETERMINISTIC, rows=(r0, r1))
        rows = np.zeros(H, bool)
        rows[r0:r1] = True
        mask = np.repeat(rows, W)
        assert np.array_equal(o[mask], of[mask])
        assert np.isnan(o[~mask]).all()             # rows outside the band are not written
        img[mask] = o[mask]
        gsum += g
        lsum += l
    assert np.array_equal(img, of)
    tol = orc.splat_tolerance(params, target, W, H)[3]
    assert (np.abs(gsum - gf) <= tol).all() and abs(lsum - lf) <= 1e-4 * lf


def test_splat_c4_distribution_reduced_size():
    """BASELINE config 4's input distribution (initialize_random, seed 42; create_test_image target) at a size
    the fp64 all-pairs oracle finishes in seconds: 128 x 128, 2000 Gaussians."""
    W, H, N = 128, 128, 2000
    params, target = orc.splat_c4_scene(N, W, H, seed=42)
    got = run_splat(params, target, W, H)
    check_splat_against_fp64(params, target, W, H, got)


def test_splat_full_size_properties():
    """BASELINE config 4 at full size (100 000 Gaussians, 1024 x 1024): no oracle can run all 1e11 pairs, so
    check size-independent properties: deterministic mode reproduces itself bit for bit, the atomic mode
    agrees with it, and a row-band of the full launch matches the same rows rendered alone."""
    W, H, N = 1024, 1024, 100_000
    params, target = orc.splat_c4_scene(N, W, H, seed=42)
    d0 = run_splat(params, target, W, H, x.FLAG_DETERMINISTIC)
    d1 = run_splat(params, target, W, H, x.FLAG_DETERMINISTIC)
    assert np.array_equal(d0[0], d1[0]) and np.array_equal(d0[1], d1[1]) and d0[2] == d1[2]
    a = run_splat(params, target, W, H, 0)
    assert np.array_equal(a[1], d0[1]) and a[2] == d0[2]
    assert np.abs(a[0] - d0[0]).max() <= 1e-4 * np.abs(d0[0]).max()
    assert np.isfinite(d0[0]).all() and np.isfinite(d0[1]).all()
    b = run_splat(params, target, W, H, x.FLAG_DETERMINISTIC, rows=(512, 528))
    assert np.array_equal(b[1][512 * W:528 * W], d0[1][512 * W:528 * W])
    st = x.splat_last_stats()
    assert st["entries"] > 0


def check_splat_sampled(params, target, W, H, got, n_pixels=512, n_gauss=64, seed=0):
    """Full-size scenes against the fp64 oracle on a SAMPLE (the all-pairs oracle would need 1e11 .. 3e12 pair
    evaluations): `n_pixels` random pixels plus one whole 16 x 16 tile summed over EVERY Gaussian (the reference's
    forward loop, gaussian_splatting_kernel.cu:33-62), and `n_gauss` random Gaussians (plus the first and the last)
    back-propagated over EVERY pixel (:73-111) with the image under test as pixel_out.
    Image: a pixel is a sequential fp32 sum of 10^3 .. 3 x 10^4 positive terms, i.e. an accumulated sum -> 1e-4 relative
    to the sum of |terms| (+ the fp32 conditioning of the exponent); the median error must stay below the per-element bar
    1e-5.  Loss: 1e-4 against the fp64 sum over the image under test.  Gradients: splat_tolerance's bound."""
    g, o, l = got
    N = params.shape[0]
    rr = np.random.default_rng(seed)
    tx, ty = int(rr.integers(0, (W + 15) // 16)), int(rr.integers(0, (H + 15) // 16))
    tile = np.stack(np.meshgrid(np.arange(tx * 16, min(W, tx * 16 + 16)), np.arange(ty * 16, min(H, ty * 16 + 16))), -1).reshape(-1, 2)
    corners = np.array([[0, 0], [W - 1, 0], [0, H - 1], [W - 1, H - 1]])
    xy = np.concatenate([np.stack([rr.integers(0, W, n_pixels), rr.integers(0, H, n_pixels)], -1), tile, corners]).astype(np.int32)
    want, cond = orc.splat_pixels(params, xy, cond=True)
    have = o.reshape(H, W, 3)[xy[:, 1], xy[:, 0]].astype(np.float64)
    err = np.abs(have - want)
    tol_i = 1e-4 * np.abs(want) + orc.FP32_EXPONENT_ULPS * 2.0 ** -24 * cond + 1e-30
    assert (err <= tol_i).all(), f"image: worst {np.max(err / tol_i):.3f} of the bound"
    assert np.median(err / (np.abs(want) + 1e-30)) <= 1e-5, "image: median relative error above the per-element bar"
    l64 = np.abs(o.astype(np.float64) - target.astype(np.float64)).sum()
    assert abs(l - l64) <= 1e-4 * l64, "loss"
    ids = np.unique(np.concatenate([rr.choice(N, min(n_gauss, N), replace=False), [0, N - 1]])).astype(np.int32)
    gs, tol = orc.splat_grads_sample(params, ids, target, o, W, H)
    assert (np.abs(g[ids] - gs) <= tol).all(), f"gradients: worst {np.max(np.abs(g[ids] - gs) / tol):.3f} of the bound"
    return float(np.max(err / (np.abs(want) + 1e-30))), float(np.max(np.abs(g[ids] - gs) / tol))


@pytest.mark.parametrize("flags", [0, 2], ids=["fast", "precise"])
def test_splat_c4_full_size_against_the_sampled_fp64_oracle(flags):
    """BASELINE configs[3] at FULL size: 100 000 Gaussians (initialize_random, seed 42), 1024 x 1024 test image."""
    W, H, N = 1024, 1024, 100_000
    params, target = orc.splat_c4_scene(N, W, H, seed=42)
    got = run_splat(params, target, W, H, flags)
    check_splat_sampled(params, target, W, H, got, seed=flags)


def test_splat_c5_size_one_view_against_the_sampled_fp64_oracle():
    """BASELINE configs[4]'s per-GPU work: 3 000 000 Gaussians, one 1024 x 1024 view (1.2e8 list entries: the
    one-chunk-per-SM binning policy for lists beyond kBinLongList, 32-bit entry offsets close to their range)."""
    W, H, N = 1024, 1024, 3_000_000
    params, target = orc.splat_c4_scene(N, W, H, seed=42)
    target = np.roll(target.reshape(H, W, 3), shift=(37, 64), axis=(0, 1)).reshape(W * H, 3).copy()  # bench.py's view 1
    got = run_splat(params, target, W, H, 0)
    st = x.splat_last_stats()
    assert st["entries"] > (40 << 20)
    check_splat_sampled(params, target, W, H, got, n_pixels=256, n_gauss=32, seed=5)
    x.shutdown()  # give the 1 GB of entry-sized scratch back


def test_covproj_full_size_against_the_oracle_on_sampled_rows():
    """BASELINE configs[2] at FULL size (2^26 elements, 12.9 GB): 100 000 random rows plus the first and the last tiles
    against the fp64 oracle (byte offsets beyond 2^31 in every array)."""
    n = 1 << 26
    gen = torch.Generator(device=DEV).manual_seed(11)
    ins = [torch.empty((n, k), device=DEV).uniform_(-1, 1, generator=gen) for k in (6, 9, 6, 3)]
    # S = A A^T + I (SPD), the distribution SURVEY 8d names and orc.covproj_inputs draws; built in slices to bound memory
    for lo in range(0, n, 1 << 24):
        A = torch.empty((1 << 24, 3, 3), device=DEV).uniform_(-1, 1, generator=gen)
        Sm = torch.bmm(A, A.transpose(1, 2)) + torch.eye(3, device=DEV)
        ins[2][lo:lo + (1 << 24)] = torch.stack([Sm[:, 0, 0], Sm[:, 0, 1], Sm[:, 0, 2], Sm[:, 1, 1], Sm[:, 1, 2], Sm[:, 2, 2]], -1)
        del A, Sm
    outs = [torch.full((n, k), float("nan"), device=DEV) for k in (3, 6, 9, 6)]
    x.covproj_fwd_bwd(*ins, *outs)
    rr = np.random.default_rng(3)
    rows = np.unique(np.concatenate([rr.integers(0, n, 100_000), np.arange(0, 300), np.arange(n - 300, n),
                                     np.arange((1 << 31) // 36 - 150, (1 << 31) // 36 + 150)]))
    ridx = torch.from_numpy(rows).to(DEV)
    J, W, S, g = [t[ridx].cpu().numpy() for t in ins]
    want = orc.covproj(J, W, S, g, np.float64)
    # 1e-5 relative per element, stated against the sum of |terms| of that element (= the same chain evaluated on the
    # absolute values of the inputs): over 10^5 rows some entries cancel to far below their row's magnitude
    scale = orc.covproj(np.abs(J), np.abs(W), np.abs(S), np.abs(g), np.float64)
    for a, b, sc, name in zip(outs, want, scale, ("out", "gJ", "gW", "gS")):
        a = a[ridx].cpu().numpy()
        assert rel_err(a, b, sc).max() < 1e-5, name
        assert np.median(rel_err(a, b, np.abs(b).max(axis=1, keepdims=True))) < 1e-6, name
    for a in outs:  # every row was written
        assert not torch.isnan(a).any()


# ---------------------------------------------------------------------------------------------------
# Caller-provided workspace (no allocation, no synchronisation, no library state) and stream re-entrancy
# ---------------------------------------------------------------------------------------------------
def run_splat_ws(ws, params, target, W, H, stream=None, flags=None):
    N = params.shape[0]
    grads = torch.zeros((N, 9), dtype=torch.float32, device=DEV)
    out = torch.full((W * H, 3), float("nan"), dtype=torch.float32, device=DEV)
    loss = torch.zeros(1, dtype=torch.float32, device=DEV)
    ws.launch(dev(params) if N else torch.empty((0, 9), device=DEV), grads, dev(target), out, loss, flags=flags, stream=stream)
    return grads, out, loss


@pytest.mark.parametrize("flags", [0, 2, 1, 3], ids=["fast", "precise", "fast-det", "precise-det"])
def test_splat_workspace_launch_equals_the_classic_launch(flags):
    """xyz_launch_gaussian_splatting_ws on a caller-owned workspace: image and loss bit-identical to the classic entry
    point, gradients bit-identical in deterministic mode (and to the atomics' tolerance otherwise); row bands; the
    header reports the list length; xyz_splat_last_stats works on it."""
    W, H, N = 256, 192, 5000
    params, target = orc.splat_scene(N, W, H, seed=21)
    g0, o0, l0 = run_splat(params, target, W, H, flags)
    e0 = x.splat_last_stats()["entries"]
    ws = x.SplatWorkspace(W, H, N, e0 + 1000, flags)
    g, o, l = run_splat_ws(ws, params, target, W, H)
    torch.cuda.synchronize()
    assert np.array_equal(o.cpu().numpy(), o0) and l.item() == l0
    if flags & x.FLAG_DETERMINISTIC:
        assert np.array_equal(g.cpu().numpy(), g0)
    else:
        assert np.abs(g.cpu().numpy() - g0).max() <= 1e-4 * np.abs(g0).max()
    st = ws.status()
    assert st == {"entries": e0, "overflows": 0, "max_entries": e0 + 1000, "overflowed": False}
    assert x.splat_last_stats()["entries"] == e0
    # a row band on its own workspace
    wsb = x.SplatWorkspace(W, H, N, e0, flags, rows=(64, 144))
    gb, ob, lb = run_splat_ws(wsb, params, target, W, H)
    g0b, o0b, l0b = run_splat(params, target, W, H, flags, rows=(64, 144))
    assert np.array_equal(ob.cpu().numpy()[64 * W:144 * W], o0b[64 * W:144 * W]) and lb.item() == l0b
    assert np.abs(gb.cpu().numpy() - g0b).max() <= 1e-4 * np.abs(g0b).max()
    assert wsb.bytes < ws.bytes


def test_splat_forward_configurations_give_identical_bits():
    """The forward pass picks its CTA shape by the number of tiles (whole tiles with 64 / 128 / 256 threads, or two
    half-tile CTAs): image AND loss must be bit-identical in every configuration (XYZ_SPLAT_FWD_THREADS is read once per
    process, so each configuration runs in its own interpreter)."""
    import hashlib
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, hashlib, numpy as np, torch\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "import oracle_lib as orc, xyz_autodiff_cuda_b200 as x\n"
        "W, H, N = 200, 150, 3000\n"
        "p, t = orc.splat_scene(N, W, H, seed=8)\n"
        "d = lambda a: torch.from_numpy(a).to('cuda:0')\n"
        "g = torch.zeros((N, 9), device='cuda:0'); o = torch.zeros((W * H, 3), device='cuda:0'); l = torch.zeros(1, device='cuda:0')\n"
        "x.launch_gaussian_splatting(d(p), g, d(t), o, l, W, H, N, x.FLAG_DETERMINISTIC)\n"
        "torch.cuda.synchronize()\n"
        "print('HASH', hashlib.sha256(o.cpu().numpy().tobytes()).hexdigest(), l.item().hex(), hashlib.sha256(g.cpu().numpy().tobytes()).hexdigest())\n")
    seen = {}
    for cfg in ("32", "64", "128", "256"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, XYZ_SPLAT_FWD_THREADS=cfg), stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:]
        seen[cfg] = [ln for ln in r.stdout.splitlines() if ln.startswith("HASH")][0]
    assert len(set(seen.values())) == 1, seen


def test_splat_workspace_overflow_is_memory_safe_and_reported():
    W, H, N = 256, 192, 5000
    params, target = orc.splat_scene(N, W, H, seed=21)
    run_splat(params, target, W, H)
    e0 = x.splat_last_stats()["entries"]
    ws = x.SplatWorkspace(W, H, N, e0 // 2)
    g, o, l = run_splat_ws(ws, params, target, W, H)
    torch.cuda.synchronize()
    st = ws.status()
    assert st["overflowed"] and st["overflows"] == 1 and st["entries"] == e0
    assert (o == 0).all() and (g == 0).all() and abs(l.item() - np.abs(target).sum()) <= 1e-5 * np.abs(target).sum()
    with pytest.raises(RuntimeError):
        x.splat_last_stats()                                   # XYZ_ERR_WORKSPACE: that launch's outputs are void
    # the same workspace keeps working for a scene that fits; the counter is sticky
    small = params.copy()
    small[:, 2:4] -= 1.5
    g2, o2, l2 = run_splat_ws(ws, small, target, W, H)
    g2s, o2s, l2s = run_splat(small, target, W, H)
    assert np.array_equal(o2.cpu().numpy(), o2s) and l2.item() == l2s
    st = ws.status()
    assert not st["overflowed"] and st["overflows"] == 1
    # wrong sizes are refused, not executed
    with pytest.raises(RuntimeError):
        lib_ws = x.SplatWorkspace(W, H, N, 1000)
        lib_ws.max_entries = 10 ** 7                           # claims more entries than the buffer was sized for
        run_splat_ws(lib_ws, params, target, W, H)
    with pytest.raises(ValueError):
        x.SplatWorkspace(W, H, N, 1000, x.FLAG_RADIX_BINNING)


def test_splat_two_streams_in_flight_equal_sequential_launches():
    """Re-entrancy (the way 8 views run on fewer GPUs): two views rendered concurrently on two streams -- once through
    two workspaces, once through the classic entry point (library scratch is per stream) -- equal the sequential
    results bit for bit (image, loss; deterministic gradients)."""
    W, H, N = 320, 256, 8000
    params, target = orc.splat_scene(N, W, H, seed=5)
    target2 = np.roll(target.reshape(H, W, 3), (13, 29), (0, 1)).reshape(W * H, 3).copy()
    F = x.FLAG_DETERMINISTIC
    ref = [run_splat(params, t, W, H, F) for t in (target, target2)]
    e0 = x.splat_last_stats()["entries"]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    tp, t1, t2 = dev(params), dev(target), dev(target2)
    torch.cuda.synchronize()
    for mode in ("workspace", "classic"):
        for rep in range(3):
            outs = []
            wss = [x.SplatWorkspace(W, H, N, e0 + 64, F) for _ in range(2)] if mode == "workspace" else [None, None]
            for st, tt, ws in ((s1, t1, wss[0]), (s2, t2, wss[1])):
                with torch.cuda.stream(st):
                    g = torch.zeros((N, 9), device=DEV); o = torch.zeros((W * H, 3), device=DEV); l = torch.zeros(1, device=DEV)
                    st.wait_stream(torch.cuda.default_stream())
                    if ws is not None:
                        ws.launch(tp, g, tt, o, l, stream=st)
                    else:
                        x.launch_gaussian_splatting(tp, g, tt, o, l, W, H, N, F, stream=st)
                    outs.append((g, o, l))
            torch.cuda.synchronize()
            for (g, o, l), (rg, ro, rl) in zip(outs, ref):
                assert np.array_equal(o.cpu().numpy(), ro) and l.item() == rl and np.array_equal(g.cpu().numpy(), rg), (mode, rep)


def test_splat_workspace_launch_captured_from_the_first_call():
    """No allocation and no synchronisation inside the workspace launch: a whole iteration (zero-grad, launch, Adam) is
    captured into a CUDA graph WITHOUT a warm-up call and replays to the eager result."""
    W, H, N = 256, 192, 5000
    params, target = orc.splat_scene(N, W, H, seed=21)
    lr = (0.5, 0.01, 0.01, 0.01, 0.02)
    def fresh():
        return (dev(params), torch.zeros((N, 9), device=DEV), torch.zeros((N, 18), device=DEV),
                torch.zeros((W * H, 3), device=DEV), torch.zeros(1, device=DEV))
    tt = dev(target)
    # eager trajectory, classic entry points
    p, g, a, o, l = fresh()
    losses = []
    for it in range(1, 4):
        l.zero_()
        x.launch_gaussian_splatting(p, g, tt, o, l, W, H, N, x.FLAG_DETERMINISTIC)
        x.adam_step_individual(p, g, a, *lr, iteration=1, zero_grads=True)   # fixed bias correction: the graph bakes it in
        losses.append(l.item())
    want_p = p.cpu().numpy()
    # the same as ONE captured graph replayed three times
    p, g, a, o, l = fresh()
    ws = x.SplatWorkspace(W, H, N, 80 * N, x.FLAG_DETERMINISTIC)
    st = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(graph, stream=st, capture_error_mode="thread_local"):
            l.zero_()
            ws.launch(p, g, tt, o, l, stream=st)
            x.adam_step_individual(p, g, a, *lr, iteration=1, zero_grads=True, stream=st)
    got = []
    for it in range(3):
        graph.replay()
        torch.cuda.synchronize()
        got.append(l.item())
    assert got == losses
    assert np.array_equal(p.cpu().numpy(), want_p)
    assert not ws.status()["overflowed"]


def test_classic_launch_refuses_to_grow_scratch_inside_a_capture():
    """Library-owned scratch cannot grow while its stream is capturing: XYZ_ERR_WORKSPACE instead of a cudaMalloc /
    synchronisation inside the capture (and instead of freeing a buffer a graph may replay on)."""
    W, H, N = 128, 96, 700
    params, target = orc.splat_scene(N, W, H, seed=2)
    tp, tt = dev(params), dev(target)
    g, o, l = torch.zeros((N, 9), device=DEV), torch.zeros((W * H, 3), device=DEV), torch.zeros(1, device=DEV)
    st = torch.cuda.Stream()      # a fresh stream: its arenas are empty
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(graph, stream=st, capture_error_mode="thread_local"):
            code = x.lib().xyz_launch_gaussian_splatting(tp.data_ptr(), g.data_ptr(), tt.data_ptr(), o.data_ptr(), l.data_ptr(),
                                                         W, H, N, st.cuda_stream, x.FLAG_ASYNC)
    torch.cuda.synchronize()
    assert code == -2      # XYZ_ERR_WORKSPACE, nothing was enqueued
    # outside a capture the same stream works, and a later capture on the warmed-up stream is accepted
    x.launch_gaussian_splatting(tp, g, tt, o, l, W, H, N, 0, stream=st)
    x.launch_gaussian_splatting(tp, g, tt, o, l, W, H, N, x.FLAG_ASYNC, stream=st)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(graph, stream=st, capture_error_mode="thread_local"):
            code = x.lib().xyz_launch_gaussian_splatting(tp.data_ptr(), g.data_ptr(), tt.data_ptr(), o.data_ptr(), l.data_ptr(),
                                                         W, H, N, st.cuda_stream, x.FLAG_ASYNC)
    assert code == 0
    graph.replay()
    torch.cuda.synchronize()


def test_comm_and_peer_optimiser_step_on_one_rank():
    """World size 1 of the multi-GPU entry points (the N-rank cases run in tests/test_gpu_multi.py): the library's NCCL
    communicator initialises, the sharded and the fused peer-memory optimiser steps equal adam_step_individual + zero-grad."""
    N = 10_001
    rng = np.random.default_rng(4)
    p0 = rng.normal(size=(N, 9)).astype(np.float32)
    g0 = rng.normal(size=(N, 9)).astype(np.float32)
    a0 = np.abs(rng.normal(size=(N, 18))).astype(np.float32) * 0.1
    lr = (0.1, 0.01, 0.001, 0.02, 0.05)
    p, g, a = dev(p0), dev(g0), dev(a0)
    x.adam_step_individual(p, g, a, *lr, iteration=4)
    comm = x.Comm(0, 1, lambda b: b)
    p1, g1, a1, l1 = dev(p0), dev(g0), dev(a0), torch.full((1,), 2.5, device=DEV)
    comm.allreduce_grads(g1)
    assert torch.equal(g1, dev(g0))
    comm.adam_step_individual_sharded(p1, g1, a1, *lr, iteration=4, total_loss=l1)
    assert torch.equal(p1, p) and torch.equal(a1, a) and (g1 == 0).all() and l1.item() == 2.5
    comm.destroy()
    grp = x.PeerGroup(0, 1, lambda h: [h])
    ps = x.PeerSplat(grp, N, lambda h: [h])
    ps.params.copy_(dev(p0)); ps.grads.copy_(dev(g0)); ps.adam.copy_(dev(a0))
    for it in range(3):                                        # several calls: sequence numbers, ticket reset
        if it:
            ps.params.copy_(dev(p0)); ps.grads.copy_(dev(g0)); ps.adam.copy_(dev(a0))
        l2 = torch.full((1,), 2.5, device=DEV)
        ps.adam_step(*lr, iteration=4, total_loss=l2)
        torch.cuda.synchronize()
        assert l2.item() == 2.5 and (ps.grads == 0).all()
        assert np.allclose(ps.params.cpu().numpy(), p.cpu().numpy(), rtol=1e-6, atol=1e-7)
        assert np.allclose(ps.adam.cpu().numpy(), a.cpu().numpy(), rtol=1e-6, atol=1e-7)
    ps.close(); grp.close()


def test_python_bindings_reject_mismatched_sizes_and_devices():
    """Wrong element counts are Python errors, not out-of-bounds kernels."""
    N, W, H = 100, 32, 32
    params, target = orc.splat_scene(N, W, H, seed=1)
    tp, tt = dev(params), dev(target)
    g, o, l = torch.zeros((N, 9), device=DEV), torch.zeros((W * H, 3), device=DEV), torch.zeros(1, device=DEV)
    with pytest.raises(ValueError):
        x.launch_gaussian_splatting(tp, g, tt, o, l, W, H, N + 1)
    with pytest.raises(ValueError):
        x.launch_gaussian_splatting(tp, g, tt, o[:-1], l, W, H, N)
    with pytest.raises(ValueError):
        x.adam_step_individual(tp, g, torch.zeros((N, 17), device=DEV), 0.1, 0.1, 0.1, 0.1, 0.1)
    with pytest.raises(ValueError):
        x.covproj_fwd_bwd(*[torch.zeros((10, k), device=DEV) for k in (6, 9, 6, 3)],
                          *[torch.zeros((10, k), device=DEV) for k in (3, 6, 8, 6)])
    with pytest.raises(ValueError):
        x.lsq_grad(torch.zeros((10, 3), dtype=torch.float64, device=DEV), torch.zeros(7, dtype=torch.float64, device=DEV))
    if torch.cuda.device_count() > 1:
        with pytest.raises(ValueError):
            x.zero_gradients(torch.zeros((N, 9), device="cuda:1"))


# ---------------------------------------------------------------------------------------------------
# Adam / zero-grad
# ---------------------------------------------------------------------------------------------------
def test_zero_gradients_and_adam_match_oracle():
    rng = np.random.default_rng(4)
    for n in (1, 7, 1000, 100_003):
        grads = dev(rng.normal(size=(n, 9)).astype(np.float32))
        x.zero_gradients(grads)
        assert (grads == 0).all()
    n = 10_001
    p = rng.normal(size=(n, 9)).astype(np.float32)
    g = rng.normal(size=(n, 9)).astype(np.float32)
    a = np.abs(rng.normal(size=(n, 18))).astype(np.float32) * 0.1
    lr = (0.1, 0.01, 0.001, 0.02, 0.05)
    for it in (1, 2, 50):
        tp, ta = dev(p), dev(a)
        x.adam_step_individual(tp, dev(g), ta, *lr, 0.9, 0.999, 1e-8, it)
        wp, wa = orc.adam_step_individual(p, g, a, lr, 0.9, 0.999, 1e-8, it)
        # FMA contraction on the device side: 1 ulp of the TERMS (|beta m| ~ 0.1, |g| ~ 1), results may cancel
        assert np.allclose(ta.cpu().numpy(), wa, rtol=1e-5, atol=2e-7)
        assert np.allclose(tp.cpu().numpy(), wp, rtol=1e-5, atol=1e-6)
    # fused zero-grad flavour: same update, gradients cleared in the same pass
    tp, ta, tg = dev(p), dev(a), dev(g)
    tp2, ta2 = dev(p), dev(a)
    x.adam_step_individual(tp, tg, ta, *lr, 0.9, 0.999, 1e-8, 7, zero_grads=True)
    x.adam_step_individual(tp2, dev(g), ta2, *lr, 0.9, 0.999, 1e-8, 7)
    assert torch.equal(tp, tp2) and torch.equal(ta, ta2) and (tg == 0).all()
    tp, ta = dev(p), dev(a)
    x.adam_step(tp, dev(g), ta, 0.01, 0.9, 0.999, 1e-8, 3)
    wp, wa = orc.adam_step_individual(p, g, a, (0.01,) * 5, 0.9, 0.999, 1e-8, 3)
    assert np.allclose(tp.cpu().numpy(), wp, rtol=1e-5, atol=1e-6)


REF_CUDA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libxyz_ref_cuda.so")


@pytest.mark.skipif(not os.path.exists(REF_CUDA), reason="oracle/_ref/libxyz_ref_cuda.so (reference CUDA kernels for sm_100a) not present")
def test_against_the_reference_cuda_kernels_on_the_same_gpu():
    """north_star: results must match the reference's own CUDA path on the same inputs.  The reference's kernels are
    built unmodified for sm_100a (fast-math flags for the splat kernel, as its CMake does) and run on this GPU."""
    import ctypes
    R = ctypes.CDLL(REF_CUDA)
    vp, ll = ctypes.c_void_p, ctypes.c_longlong
    R.refcuda_splat.argtypes = [vp] * 5 + [ctypes.c_int] * 3
    R.refcuda_lsq.argtypes = [vp, ll, vp]
    R.refcuda_accumulate.argtypes = [vp, vp, ll, vp]
    R.refcuda_covproj.argtypes = [vp] * 8 + [ll]
    # splat (fast-math flavour on both sides)
    for (W, H, N, seed) in [(64, 48, 50, 3), (100, 70, 300, 11)]:
        params, target = orc.splat_scene(N, W, H, seed=seed)
        g, o, l = run_splat(params, target, W, H, 0)
        tp, tt = dev(params), dev(target)
        rg = torch.zeros((N, 9), device=DEV)
        ro = torch.zeros((W * H, 3), device=DEV)
        rl = torch.zeros(1, device=DEV)
        assert R.refcuda_splat(tp.data_ptr(), rg.data_ptr(), tt.data_ptr(), ro.data_ptr(), rl.data_ptr(), W, H, N) == 0
        torch.cuda.synchronize()
        ro, rg, rl = ro.cpu().numpy(), rg.cpu().numpy(), rl.item()
        tol = orc.splat_tolerance(params, target, W, H)[3]
        assert (np.abs(o - ro) <= 1e-5 * np.maximum(np.abs(ro), np.abs(ro).max() * 1e-3)).all()
        assert abs(l - rl) <= 1e-4 * abs(rl)
        assert (np.abs(g - rg) <= 2 * tol).all()   # both sides are fp32 sums in different orders
    # covariance chain
    n = 10_000
    J, W9, S, gg = orc.covproj_inputs(n, seed=3)
    got = run_covproj(J, W9, S, gg)
    ins = [dev(a) for a in (J, W9, S, gg)]
    outs = [torch.zeros((n, k), device=DEV) for k in (3, 6, 9, 6)]
    assert R.refcuda_covproj(*[t.data_ptr() for t in ins], *[t.data_ptr() for t in outs], n) == 0
    torch.cuda.synchronize()
    for a, b in zip(got, outs):
        b = b.cpu().numpy()
        assert rel_err(a, b, np.abs(b).max(axis=1, keepdims=True)).max() < 1e-5
    # least squares
    data = orc.lsq_data(100_000, seed=6)
    got, _ = run_lsq(data, (0.3, 1.2, -0.4, 0.1))
    prm = dev(np.array([0.3, 1.2, -0.4, 0.1, 0, 0, 0, 0], np.float64))
    td = dev(data)
    assert R.refcuda_lsq(td.data_ptr(), data.shape[0], prm.data_ptr()) == 0
    torch.cuda.synchronize()
    assert rel_err(got, prm[4:].cpu().numpy()).max() < 1e-10
    # accumulation
    idx, val = orc.accumulate_inputs(1 << 20, 1024, "zipf", seed=4)
    got = run_acc(idx, val, 1024)
    ti, tv = dev(idx), dev(val)
    grad = torch.zeros(1024, device=DEV)
    assert R.refcuda_accumulate(ti.data_ptr(), tv.data_ptr(), idx.size, grad.data_ptr()) == 0
    torch.cuda.synchronize()
    assert (np.abs(got - grad.cpu().numpy()) <= 2e-4 * orc.accumulate_exact(idx, np.abs(val), 1024) + 1e-30).all()
    # the SHIPPED example's own kernel (parallel_gradient_computation_kernel, linear_regression_sgd.cu:86-123, compiled
    # unmodified) against XYZ_FLAG_LSQ_SHIPPED_GRAPH, and one update_parameters_kernel step against xyz_lsq_sgd_update_f64
    if hasattr(R, "refcuda_lsq_shipped"):
        R.refcuda_lsq_shipped.argtypes = [vp, ll, vp]
        R.refcuda_lsq_update.argtypes = [vp, ctypes.c_double, ctypes.c_int]
        data = orc.lsq_data(8192, seed=6)
        td = dev(data)
        ours = torch.zeros(8, dtype=torch.float64, device=DEV)
        ours[:4] = dev(np.array([0.2, 1.1, -0.3, 0.05]))
        theirs = ours.clone()
        x.lsq_grad(td, ours, flags=x.FLAG_LSQ_SHIPPED_GRAPH)
        assert R.refcuda_lsq_shipped(td.data_ptr(), 8192, theirs.data_ptr()) == 0
        torch.cuda.synchronize()
        assert rel_err(ours.cpu().numpy()[4:], theirs.cpu().numpy()[4:]).max() < 1e-10
        x.lsq_sgd_update(ours, 1e-4, 8192)
        assert R.refcuda_lsq_update(theirs.data_ptr(), 1e-4, 8192) == 0
        torch.cuda.synchronize()
        assert rel_err(ours.cpu().numpy()[:4], theirs.cpu().numpy()[:4]).max() < 1e-12


REF_CUDA_OURHDR = os.path.join(os.path.dirname(REF_CUDA), "libxyz_ref_cuda_ourhdr.so")


@pytest.mark.skipif(not os.path.exists(REF_CUDA_OURHDR), reason="oracle/_ref/libxyz_ref_cuda_ourhdr.so not built")
def test_reference_kernel_sources_on_this_repos_headers():
    """Drop-in check of the header library at kernel level: the reference's splat kernel file (unmodified, with its own
    custom Logics) and the reference-style least-squares / accumulation / matmul-chain kernels of
    oracle/ref_cuda_driver.cu, compiled against include/xyz_autodiff of THIS repo (oracle/Makefile), must give the
    results of the same sources compiled against the reference's headers: values bit-identical (same user code, same
    flags), atomically accumulated sums within the accumulated-sum tolerance (the warp-aggregated add_grad changes the
    summation order)."""
    import ctypes
    vp, ll = ctypes.c_void_p, ctypes.c_longlong
    libs = []
    for path in (REF_CUDA, REF_CUDA_OURHDR):
        R = ctypes.CDLL(path)
        R.refcuda_splat.argtypes = [vp] * 5 + [ctypes.c_int] * 3
        R.refcuda_lsq.argtypes = [vp, ll, vp]
        R.refcuda_accumulate.argtypes = [vp, vp, ll, vp]
        R.refcuda_covproj.argtypes = [vp] * 8 + [ll]
        libs.append(R)
    for (W, H, N, seed) in [(64, 48, 50, 3), (100, 70, 300, 11)]:
        params, target = orc.splat_scene(N, W, H, seed=seed)
        tp, tt = dev(params), dev(target)
        res = []
        for R in libs:
            rg = torch.zeros((N, 9), device=DEV)
            ro = torch.zeros((W * H, 3), device=DEV)
            rl = torch.zeros(1, device=DEV)
            assert R.refcuda_splat(tp.data_ptr(), rg.data_ptr(), tt.data_ptr(), ro.data_ptr(), rl.data_ptr(), W, H, N) == 0
            torch.cuda.synchronize()
            res.append((ro.cpu().numpy(), rg.cpu().numpy(), rl.item()))
        assert np.array_equal(res[0][0], res[1][0]), "image"
        tol = orc.splat_tolerance(params, target, W, H)[3]
        assert (np.abs(res[0][1] - res[1][1]) <= 2 * tol).all()
        assert abs(res[0][2] - res[1][2]) <= 1e-4 * abs(res[0][2])
    n = 10_000
    J, W9, S, gg = orc.covproj_inputs(n, seed=3)
    ins = [dev(a) for a in (J, W9, S, gg)]
    outs = []
    for R in libs:
        o = [torch.zeros((n, k), device=DEV) for k in (3, 6, 9, 6)]
        assert R.refcuda_covproj(*[t.data_ptr() for t in ins], *[t.data_ptr() for t in o], n) == 0
        torch.cuda.synchronize()
        outs.append([t.cpu().numpy() for t in o])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    data = orc.lsq_data(100_000, seed=6)
    td = dev(data)
    grads = []
    for R in libs:
        prm = dev(np.array([0.3, 1.2, -0.4, 0.1, 0, 0, 0, 0], np.float64))
        assert R.refcuda_lsq(td.data_ptr(), data.shape[0], prm.data_ptr()) == 0
        torch.cuda.synchronize()
        grads.append(prm[4:].cpu().numpy())
    assert rel_err(grads[1], grads[0]).max() < 1e-10
    idx, val = orc.accumulate_inputs(1 << 20, 1024, "zipf", seed=4)
    ti, tv = dev(idx), dev(val)
    exact = orc.accumulate_exact(idx, val, 1024)
    abs_sum = orc.accumulate_exact(idx, np.abs(val), 1024)
    for R in libs:
        grad = torch.zeros(1024, device=DEV)
        assert R.refcuda_accumulate(ti.data_ptr(), tv.data_ptr(), idx.size, grad.data_ptr()) == 0
        torch.cuda.synchronize()
        assert (np.abs(grad.cpu().numpy() - exact) <= 1e-4 * abs_sum + 1e-30).all()


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError):
        x.covproj_fwd_bwd(*[torch.zeros((4, k)) for k in (6, 9, 6, 3, 3, 6, 9, 6)])
    assert x.launch_count() > 0
