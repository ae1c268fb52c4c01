"""include/xyz_autodiff/batched.cuh (`-m gpu`): user graphs written with the public op:: API, evaluated for a whole
batch by batched::for_each (TMA-staged tiles, register-resident shared-parameter adjoints, fixed-order reduction).

Checked against the oracle (the reference's own graph code compiled for the host, fp64) and against the hand-written
kernels of the C ABI that implement the same two graphs (xyz_lsq_grad_f64, xyz_covproj_shared_w_fwd_bwd_f32).
tests/csrc/batched_probe.cu holds the two graphs."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "csrc", "_build", "libxyz_batched.so")
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "csrc"), "_build/libxyz_batched.so"], check=True,
                       stdout=subprocess.DEVNULL)
    L = ctypes.CDLL(SO)
    L.batched_lsq.restype = ctypes.c_float
    L.batched_lsq.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    L.batched_chain.restype = ctypes.c_float
    L.batched_chain.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return L


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("n", [1, 127, 128, 1000, 1_000_000])
def test_batched_least_squares_graph(lib, n):
    data = orc.lsq_data(n, seed=n)
    vals = (0.3, 1.2, -0.4, 0.1)
    d = dev(data)
    v5 = dev(np.array(vals + (0.0,), np.float64))
    g5 = torch.zeros(5, dtype=torch.float64, device=DEV)
    assert lib.batched_lsq(d.data_ptr(), n, v5.data_ptr(), g5.data_ptr(), 1) >= 0
    torch.cuda.synchronize()
    want_g, want_l = orc.lsq_grad(data, vals)
    got = g5.cpu().numpy()
    scale = np.abs(want_g).max()
    assert (np.abs(got[:4] - want_g) <= 1e-10 * scale).all(), (got, want_g)
    assert abs(got[4] - want_l) <= 1e-10 * abs(want_l)
    # the hand-written kernel of the C ABI computes the same graph
    prm = torch.zeros(8, dtype=torch.float64, device=DEV)
    prm[:4] = dev(np.array(vals))
    loss = torch.zeros(1, dtype=torch.float64, device=DEV)
    x.lsq_grad(d, prm, loss)
    assert np.allclose(prm[4:].cpu().numpy(), got[:4], rtol=1e-10, atol=1e-10 * scale)
    # fixed-order reduction: bit-identical run to run, and it accumulates into the caller's gradients
    g5b = torch.zeros(5, dtype=torch.float64, device=DEV)
    lib.batched_lsq(d.data_ptr(), n, v5.data_ptr(), g5b.data_ptr(), 1)
    assert torch.equal(g5, g5b)
    lib.batched_lsq(d.data_ptr(), n, v5.data_ptr(), g5b.data_ptr(), 1)
    assert np.allclose(g5b.cpu().numpy(), 2 * got, rtol=1e-14)


def pack_chain(J, S, g):
    return np.ascontiguousarray(np.concatenate([J, S, g], axis=1), np.float32)   # rows of ChainIn: J6 S6 g3


@pytest.mark.parametrize("n", [1, 128, 129, 70_003, 1 << 20])
def test_batched_matrix_chain_with_shared_w(lib, n):
    J, W, S, g = orc.covproj_inputs(n, seed=n + 7)
    W9 = W[0].copy()
    w18 = dev(np.concatenate([W9, W9.reshape(3, 3).T.reshape(-1)]).astype(np.float32))
    gw18 = torch.zeros(18, dtype=torch.float32, device=DEV)
    tin = dev(pack_chain(J, S, g))
    tout = torch.full((n, 15), float("nan"), dtype=torch.float32, device=DEV)
    assert lib.batched_chain(tin.data_ptr(), tout.data_ptr(), n, w18.data_ptr(), gw18.data_ptr(), 1) >= 0
    torch.cuda.synchronize()
    out = tout.cpu().numpy()
    Wrep = np.broadcast_to(W9, (n, 9)).copy()
    w_out, w_gJ, w_gW, w_gS = orc.covproj(J, Wrep, S, g, np.float64)
    # per element: 1e-5 relative to the magnitude of the sums that produce it.  The row maximum is that magnitude
    # except for the few rows (a handful per million) whose every entry is a cancellation; for those the bound is the
    # sum of |terms|, obtained by running the same graph on |inputs| (all terms positive, nothing cancels).
    a_out, a_gJ, _, a_gS = orc.covproj(np.abs(J), np.abs(Wrep), np.abs(S), np.abs(g), np.float64)
    for a, b, mag, name in ((out[:, 0:3], w_out, a_out, "out"), (out[:, 3:9], w_gJ, a_gJ, "gJ"), (out[:, 9:15], w_gS, a_gS, "gS")):
        scale = np.maximum(np.abs(b).max(axis=1, keepdims=True), 0.1 * mag.max(axis=1, keepdims=True))
        assert (np.abs(a - b) <= 1e-5 * scale).all(), name
    gw = gw18.cpu().numpy().astype(np.float64)
    gW = gw[:9] + gw[9:].reshape(3, 3).T.reshape(-1)          # adjoint of W = its own + the transposed leaf's
    assert (np.abs(gW - w_gW.sum(0)) <= 1e-4 * np.abs(w_gW).sum(0) + 1e-30).all()
    # the hand-written kernel: same per-element values to 1e-5, same accumulated gradient to 1e-4
    o, gJ, gS = [torch.empty((n, k), device=DEV) for k in (3, 6, 6)]
    gW9 = torch.zeros(9, device=DEV)
    x.covproj_shared_w_fwd_bwd(dev(J), dev(W9), dev(S), dev(g), o, gJ, gW9, gS)
    assert (np.abs(o.cpu().numpy() - out[:, 0:3]) <= 1e-5 * np.abs(w_out).max(axis=1, keepdims=True)).all()
    assert (np.abs(gW9.cpu().numpy() - gW) <= 1e-4 * np.abs(w_gW).sum(0) + 1e-30).all()


@pytest.mark.parametrize("n", [129, 70_003, 1 << 20])
def test_batched_chain_as_one_ternary_node(lib, n):
    """op::covariance_projection (one TernaryOperation node) through batched::for_each: per-element values and adjoints
    bit-identical to the hand-written kernel (same arithmetic in the same order), accumulated dW within 1e-4 of the
    sum of |terms| of the fp64 oracle."""
    lib.batched_chain_ternary.restype = ctypes.c_float
    lib.batched_chain_ternary.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    J, W, S, g = orc.covproj_inputs(n, seed=n + 11)
    W9 = W[2].copy()
    tin = dev(pack_chain(J, S, g))
    tout = torch.full((n, 15), float("nan"), dtype=torch.float32, device=DEV)
    gw9 = torch.zeros(9, dtype=torch.float32, device=DEV)
    assert lib.batched_chain_ternary(tin.data_ptr(), tout.data_ptr(), n, dev(W9).data_ptr(), gw9.data_ptr(), 1) >= 0
    torch.cuda.synchronize()
    out = tout.cpu().numpy()
    o, gJ, gS = [torch.empty((n, k), device=DEV) for k in (3, 6, 6)]
    gW9 = torch.zeros(9, device=DEV)
    x.covproj_shared_w_fwd_bwd(dev(J), dev(W9), dev(S), dev(g), o, gJ, gW9, gS)
    assert np.array_equal(out[:, 0:3], o.cpu().numpy())
    assert np.array_equal(out[:, 3:9], gJ.cpu().numpy())
    assert np.array_equal(out[:, 9:15], gS.cpu().numpy())
    Wrep = np.broadcast_to(W9, (n, 9)).copy()
    w_gW = orc.covproj(J, Wrep, S, g, np.float64)[2]
    assert (np.abs(gw9.cpu().numpy() - w_gW.sum(0)) <= 1e-4 * np.abs(w_gW).sum(0) + 1e-30).all()


def test_batched_unaligned_base_takes_the_plain_path(lib):
    n = 5000
    J, W, S, g = orc.covproj_inputs(n, seed=3)
    W9 = W[1].copy()
    w18 = dev(np.concatenate([W9, W9.reshape(3, 3).T.reshape(-1)]).astype(np.float32))
    packed = pack_chain(J, S, g)
    res = []
    for shift in (0, 1):   # 4-byte shifted copies of input and output
        tin = torch.zeros(n * 15 + 1, dtype=torch.float32, device=DEV)
        tin[shift:shift + n * 15] = dev(packed).flatten()
        tout = torch.zeros(n * 15 + 1, dtype=torch.float32, device=DEV)
        gw18 = torch.zeros(18, dtype=torch.float32, device=DEV)
        assert lib.batched_chain(tin[shift:].data_ptr(), tout[shift:].data_ptr(), n, w18.data_ptr(), gw18.data_ptr(), 1) >= 0
        torch.cuda.synchronize()
        res.append((tout[shift:shift + n * 15].cpu().numpy(), gw18.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0])               # per-element results: same arithmetic on both paths
    assert np.allclose(res[0][1], res[1][1], rtol=1e-4, atol=1e-4 * np.abs(res[0][1]).max())


def test_batched_throughput_is_reported(lib, capsys):
    """Informational: the generic template path against the hand-written kernels at the BASELINE sizes."""
    n = 1 << 25
    tin = torch.empty((n, 15), dtype=torch.float32, device=DEV).uniform_(-1, 1)
    tout = torch.empty((n, 15), dtype=torch.float32, device=DEV)
    w18 = torch.empty(18, dtype=torch.float32, device=DEV).uniform_(-1, 1)
    gw18 = torch.zeros(18, dtype=torch.float32, device=DEV)
    lib.batched_chain(tin.data_ptr(), tout.data_ptr(), n, w18.data_ptr(), gw18.data_ptr(), 2)
    ms = lib.batched_chain(tin.data_ptr(), tout.data_ptr(), n, w18.data_ptr(), gw18.data_ptr(), 5)
    assert ms > 0
    chain_gbs = 120 * n / ms / 1e6
    del tin, tout
    m = 1 << 27
    data = torch.empty((m, 3), dtype=torch.float64, device=DEV).uniform_(-5, 5)
    v5 = torch.tensor([0.0, 1.0, 0.0, 0.0, 0.0], dtype=torch.float64, device=DEV)
    g5 = torch.zeros(5, dtype=torch.float64, device=DEV)
    lib.batched_lsq(data.data_ptr(), m, v5.data_ptr(), g5.data_ptr(), 2)
    ms2 = lib.batched_lsq(data.data_ptr(), m, v5.data_ptr(), g5.data_ptr(), 5)
    assert ms2 > 0
    with capsys.disabled():
        print(f"\nBATCHED for_each: matrix chain (op::matmul graph, shared W) 2^25 elements {ms:.3f} ms = {chain_gbs:.0f} GB/s; "
              f"least squares (op:: graph, fp64) 2^27 points {ms2:.3f} ms = {24 * m / ms2 / 1e6:.0f} GB/s")
    assert chain_gbs > 2000     # a template path that falls off the HBM roofline by 3x would be a regression
