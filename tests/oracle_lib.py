"""tests/oracle_lib.py -- TEST INFRASTRUCTURE: numpy bindings of the two CPU checkers and the seeded
synthetic-input generators shared by tests/, __graft_entry__.smoke() and bench.py's cpu legs.

  port : oracle/_build/libxyz_oracle.so  (oracle/xyz_oracle.cpp, the restatement; always available)
  ref  : oracle/_ref/libxyz_ref.so       (the reference's own headers + splat kernel body compiled
                                          for the host; built only where /root/reference exists,
                                          travels prebuilt to the GPU box)
Nothing under xyz-autodiff-cuda_b200/ imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_PATH = os.path.join(ROOT, "oracle", "_build", "libxyz_oracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libxyz_ref.so")

_libs = {}
_ll = ctypes.c_longlong


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


def load(which: str = "port") -> ctypes.CDLL:
    if which not in _libs:
        path = PORT_PATH if which == "port" else REF_PATH
        if which == "port" and not os.path.exists(path):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libxyz_oracle.so"], check=True,
                           stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(path)
        pre = "orc" if which == "port" else "ref"
        getattr(L, pre + "_kind").restype = ctypes.c_char_p
        _libs[which] = L
    return _libs[which]


def _fn(which, name):
    return getattr(load(which), ("orc_" if which == "port" else "ref_") + name)


# ---- op ids (tests/csrc/api_eval.inc) ---------------------------------------------------------------
OPS = ["EXP", "SIN", "COS", "SIGMOID", "SQUARED", "NEG", "L1", "L2", "SUM", "ADD_C", "SUB_C", "MUL_C", "DIV_C",
       "CONST_ADD", "CONST_SUB", "SYM_INV", "QUAT", "BROADCAST3", "ADD", "SUB", "MUL", "DIV", "MATMUL", "COV_GEN",
       "MAT_TO_COV3", "SCALE_ROT_COV3", "MAHALANOBIS", "MAHALANOBIS_CENTER"]
OP = {n: i for i, n in enumerate(OPS)}


def eval_op(which, op, in1, in2=None, cst=0.0, gout=None, aux=0, dtype=np.float64, fn=None):
    """Returns (out, gin1, gin2) of one op evaluated by the chosen checker (or by `fn`, a ctypes function with
    the same signature, e.g. the public-header probes in tests/csrc)."""
    in1 = np.ascontiguousarray(in1, dtype)
    in2 = np.ascontiguousarray(in2 if in2 is not None else np.zeros(1), dtype)
    gout = np.ascontiguousarray(gout, dtype)
    out = np.zeros(16, dtype)
    gin1 = np.zeros(16, dtype)
    gin2 = np.zeros(16, dtype)
    nout = ctypes.c_int(0)
    f = fn if fn is not None else _fn(which, "eval_op_f64" if dtype == np.float64 else "eval_op_f32")
    c = ctypes.c_double(cst) if dtype == np.float64 else ctypes.c_float(cst)
    rc = f(OP[op], aux, _p(in1), in1.size, _p(in2), in2.size, c, _p(gout), _p(out), ctypes.byref(nout), _p(gin1),
           _p(gin2))
    assert rc == 0, f"eval_op {op} rc={rc}"
    return out[:nout.value].copy(), gin1[:in1.size].copy(), gin2[:in2.size].copy()


# ---- whole-path functions ---------------------------------------------------------------------------
def splat(params, target, W, H, dtype=np.float64, which="port", threads=None, absgrads=None, kinkgrads=None):
    """All-pairs reference semantics.  Returns (grads, output, loss, margin); fp64 only: `absgrads`
    (N x 9 float64, +=) receives the sum of |per-pair gradient terms|, `kinkgrads` the part of it that
    comes from pairs sitting on the L1 kink (sign ambiguous in fp32)."""
    threads = threads or os.cpu_count() or 1
    N = params.shape[0]
    p = np.ascontiguousarray(params, dtype)
    t = np.ascontiguousarray(target, dtype)
    g = np.zeros((N, 9), dtype)
    o = np.zeros((W * H, 3), dtype)
    l = np.zeros(1, dtype)
    m = np.zeros(1, np.float64)
    if dtype == np.float64:
        assert which == "port", "the reference kernel is float-only"
        _fn(which, "splat_f64")(_p(p), _p(g), _p(t), _p(o), _p(l), W, H, N, threads, _p(m), _p(absgrads), _p(kinkgrads))
    else:
        _fn(which, "splat_f32")(_p(p), _p(g), _p(t), _p(o), _p(l), W, H, N, threads)
        m[0] = np.nan
    return g, o, float(l[0]), float(m[0])


def splat_tolerance(params, target, W, H, rtol=1e-4):
    """fp64 ground truth + the per-component tolerance for fp32 gradients:
    rtol * sum|terms| + 2 * (terms of kink-ambiguous pairs) + 1e-5 * (largest sum|terms| of that Gaussian).
    The last term covers components whose per-pair chain rule cancels EXACTLY in the oracle (e.g. d/d(rotation)
    of an isotropic Gaussian is 0 pair by pair) while an implementation that applies the chain rule to
    accumulated sums is left with rounding residue of the non-cancelled intermediates.
    Returns (grads, output, loss, tol)."""
    n = params.shape[0]
    absg = np.zeros((n, 9))
    kink = np.zeros((n, 9))
    g, o, l, _ = splat(params, target, W, H, np.float64, absgrads=absg, kinkgrads=kink)
    return g, o, l, rtol * absg + 2.0 * kink + 1e-5 * absg.max(axis=1, keepdims=True) + 1e-30


FP32_EXPONENT_ULPS = 4.0   # units of 2^-24 per unit of orc's exponent sensitivity (the reference's own fp32 kernel stays below 0.5 of this bound)


def splat_tolerance_fp32(params, target, W, H, rtol=1e-4):
    """splat_tolerance plus the conditioning of the exponent in fp32 (orc_splat_f64_cond): returns
    (grads, output, loss, tol_grads (N x 9), tol_image (P x 3)).  For well-conditioned scenes (the BASELINE
    distributions) the extra terms are far below the 1e-4 / 1e-5 bounds; for sub-pixel or strongly anisotropic
    Gaussians they are what fp32 itself -- the reference's kernel included -- can deliver."""
    n = params.shape[0]
    p = np.ascontiguousarray(params, np.float64)
    t = np.ascontiguousarray(target, np.float64)
    g = np.zeros((n, 9)); o = np.zeros((W * H, 3)); l = np.zeros(1); m = np.zeros(1)
    absg = np.zeros((n, 9)); kink = np.zeros((n, 9)); condg = np.zeros((n, 9)); condi = np.zeros((W * H, 3))
    load("port").orc_splat_f64_cond(_p(p), _p(g), _p(t), _p(o), _p(l), W, H, n, os.cpu_count() or 1, _p(m), _p(absg), _p(kink),
                                    _p(condg), _p(condi))
    eps = FP32_EXPONENT_ULPS * 2.0 ** -24
    tol_g = rtol * absg + 2.0 * kink + 1e-5 * absg.max(axis=1, keepdims=True) + eps * condg + 1e-30
    tol_i = 1e-5 * np.maximum(np.abs(o), np.abs(o).max() * 1e-3) + eps * condi + 1e-30
    return g, o, float(l[0]), tol_g, tol_i


def splat_pixels(params, xy, cond=False, threads=None):
    """fp64, EVERY Gaussian summed into the n sampled pixels xy (n x 2: x, y) in ascending index (the forward loop of the
    reference kernel).  Returns out (n x 3) [, condimage (n x 3)]: the sampled rows of splat()'s output."""
    threads = threads or os.cpu_count() or 1
    p = np.ascontiguousarray(params, np.float64)
    q = np.ascontiguousarray(xy, np.int32)
    n = q.shape[0]
    out = np.zeros((n, 3)); ci = np.zeros((n, 3)) if cond else None
    load("port").orc_splat_pixels_f64(_p(p), p.shape[0], _p(q), n, _p(out), _p(ci), threads)
    return (out, ci) if cond else out


def splat_grads_sample(params, ids, target, image, W, H, rtol=1e-4, threads=None):
    """fp64 gradients of the m sampled Gaussians `ids` over EVERY pixel, with `image` (P x 3, the image under test) as the
    finished pixel_out of the backward loop.  Returns (grads (m x 9), tol (m x 9)) with splat_tolerance's bound."""
    threads = threads or os.cpu_count() or 1
    p = np.ascontiguousarray(params, np.float64)
    t = np.ascontiguousarray(target, np.float64)
    im = np.ascontiguousarray(image, np.float64)
    ii = np.ascontiguousarray(ids, np.int32)
    m = ii.shape[0]
    g = np.zeros((m, 9)); absg = np.zeros((m, 9)); kink = np.zeros((m, 9))
    load("port").orc_splat_grads_sample_f64(_p(p), _p(ii), m, _p(t), _p(im), W, H, _p(g), _p(absg), _p(kink), threads)
    return g, rtol * absg + 2.0 * kink + 1e-5 * absg.max(axis=1, keepdims=True) + 1e-30


def splat_hard_scene(case: int):
    """Seeded ILL-CONDITIONED scenes: odd image sizes, Gaussians from sub-pixel to image-sized (log-scale in [-1, 2.3]),
    arbitrary rotations, some centres far outside the image.  Returns (params, target, W, H)."""
    rr = np.random.default_rng(4242 + case)
    W, H = int(rr.integers(1, 130)), int(rr.integers(1, 100))
    N = int(rr.choice([1, 2, 17, 64, 200, 400]))
    params, target = splat_scene(N, W, H, seed=1000 + case, small=bool(case % 2))
    params[:, 2:4] = rr.uniform(-1.0, 2.3, (N, 2)).astype(np.float32)
    if case % 3 == 0:
        params[: max(1, N // 8), 0:2] += 500.0
    return params, target, W, H


def lsq_grad(data, values, residual_only=False, which="port", threads=1):
    """Returns (grad[4], loss_sum).  residual_only: False / True (root = residual of the test graph) / 2 (the graph the
    shipped example builds, linear_regression_sgd.cu:103-122)."""
    d = np.ascontiguousarray(data, np.float64)
    prm = np.concatenate([np.asarray(values, np.float64), np.zeros(4)])
    ls = np.zeros(1)
    _fn(which, "lsq_grad_f64")(_p(d), _ll(d.shape[0]), _p(prm), _p(ls), int(residual_only), threads)
    return prm[4:].copy(), float(ls[0])


def accumulate(idx, val, k, which="port", threads=1):
    val = np.ascontiguousarray(val)
    grad = np.zeros(k, val.dtype)
    name = "accumulate_f32" if val.dtype == np.float32 else "accumulate_f64"
    ip = _p(np.ascontiguousarray(idx, np.int32)) if idx is not None else None
    _fn(which, name)(ip, _p(val), _ll(val.size), _p(grad), k, threads)
    return grad


def accumulate_exact(idx, val, k):
    grad = np.zeros(k, np.float64)
    ip = _p(np.ascontiguousarray(idx, np.int32)) if idx is not None else None
    load("port").orc_accumulate_f32_exact(ip, _p(np.ascontiguousarray(val, np.float32)), _ll(val.size), _p(grad), k)
    return grad


def covproj(J, W, S, g, dtype=np.float64, which="port", threads=None):
    threads = threads or os.cpu_count() or 1
    n = J.shape[0]
    ins = [np.ascontiguousarray(a, dtype) for a in (J, W, S, g)]
    outs = [np.zeros((n, k), dtype) for k in (3, 6, 9, 6)]
    _fn(which, "covproj_f64" if dtype == np.float64 else "covproj_f32")(*[_p(a) for a in ins], *[_p(a) for a in outs],
                                                                       _ll(n), threads)
    return outs


def adam_step_individual(params, grads, adam, lr, beta1, beta2, eps, iteration):
    p = np.ascontiguousarray(params, np.float32).copy()
    a = np.ascontiguousarray(adam, np.float32).copy()
    g = np.ascontiguousarray(grads, np.float32)
    lr = np.asarray(lr, np.float32)
    load("port").orc_adam_step_individual(_p(p), _p(g), _p(a), p.shape[0], _p(lr), ctypes.c_float(beta1),
                                          ctypes.c_float(beta2), ctypes.c_float(eps), iteration)
    return p, a


def lsq_select_batch(data, batch_size, seed, epoch):
    d = np.ascontiguousarray(data, np.float64)
    out = np.zeros((batch_size, 3), np.float64)
    load("port").orc_lsq_select_batch(_p(d), _ll(d.shape[0]), _p(out), _ll(batch_size), ctypes.c_uint64(seed),
                                      ctypes.c_uint64(epoch))
    return out


def splat_records(params):
    p = np.ascontiguousarray(params, np.float32)
    rec = np.zeros((p.shape[0], 12), np.float32)
    load("port").orc_splat_records(_p(p), p.shape[0], _p(rec))
    return rec


def splat_binning(records, W, H, row_begin=0, row_end=None, d2max=176.0, no_cull=False):
    """Integer oracle: returns (rects (N,4), tile_ranges (T,2), sorted_ids)."""
    row_end = H if row_end is None else row_end
    rec = np.ascontiguousarray(records, np.float32)
    N = rec.shape[0]
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    rects = np.zeros((max(N, 1), 4), np.int32)
    ranges = np.zeros((ntiles, 2), np.int32)
    L = load("port")
    L.orc_splat_binning.restype = _ll
    total = L.orc_splat_binning(_p(rec), N, W, H, row_begin, row_end, ctypes.c_float(d2max), int(no_cull), _p(rects),
                                _p(ranges), None, _ll(0))
    ids = np.zeros(max(total, 1), np.int32)
    L.orc_splat_binning(_p(rec), N, W, H, row_begin, row_end, ctypes.c_float(d2max), int(no_cull), _p(rects),
                        _p(ranges), _p(ids), _ll(total))
    return rects[:N], ranges, ids[:total]


def splat_binning_counting(records, W, H, row_begin=0, row_end=None, d2max=176.0, no_cull=False, ctas_total=592,
                           bwd_chunk=128):
    """Sequential emulation of the counting-sort binning kernels (index arithmetic transcribed from
    csrc/splat_host.cu section 2b).  Returns (tile_ranges, sorted_ids, chunk_info: the backward work records as the binning leaves them) or raises if a tile slot was written
    twice / outside its tile's range / not at all."""
    row_end = H if row_end is None else row_end
    rec = np.ascontiguousarray(records, np.float32)
    N = rec.shape[0]
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    ranges = np.zeros((ntiles, 2), np.int32)
    L = load("port")
    fn = L.orc_splat_binning_counting
    fn.restype = _ll
    total = fn(_p(rec), N, W, H, row_begin, row_end, ctypes.c_float(d2max), int(no_cull), int(ctas_total), int(bwd_chunk),
               _p(ranges), None, _ll(0), None, 0)
    ids = np.full(max(total, 1), -1, np.int32)
    n_info = total // bwd_chunk + ntiles
    info = np.full((max(n_info, 1), 4), -7, np.int32)
    got = fn(_p(rec), N, W, H, row_begin, row_end, ctypes.c_float(d2max), int(no_cull), int(ctas_total), int(bwd_chunk),
             _p(ranges), _p(ids), _ll(total), _p(info), int(n_info))
    if got != total:
        raise AssertionError("counting-sort emulation: a slot was written twice, out of range, or left empty")
    return ranges, ids[:total], info[:n_info]


def kat(which, name, *args, n=64, dtype=np.float64):
    res = np.zeros(n, dtype)
    k = _fn(which, "kat_" + name)(*args, _p(res))
    return res[:k].copy()


# ---- seeded synthetic inputs ----------------------------------------------------
def std_mt19937_u32(seed: int, n: int) -> np.ndarray:
    """The raw 32-bit stream of std::mt19937(seed) (numpy's legacy RandomState seeds identically)."""
    rs = np.random.RandomState(seed)
    return np.frombuffer(rs.bytes(4 * n), dtype="<u4").copy()


def _std_uniform_float(u32: np.ndarray, lo: float, hi: float) -> np.ndarray:
    """libstdc++ std::uniform_real_distribution<float>(lo, hi) driven by one mt19937 draw each."""
    r = (u32.astype(np.float32) * np.float32(1.0 / 4294967296.0)).astype(np.float32)
    r = np.where(r >= np.float32(1.0), np.nextafter(np.float32(1.0), np.float32(0.0)), r)
    return (np.float32(hi - lo) * r + np.float32(lo)).astype(np.float32)


def splat_c4_scene(N: int, W: int, H: int, seed: int = 42):
    """GaussianCollection::initialize_random (reference gaussian_parameters.cu:27-66) driven by
    std::mt19937(seed), and create_test_image (image_utils.cpp:59-75) as the target."""
    u = std_mt19937_u32(seed, 8 * N).reshape(N, 8)  # draw order per Gaussian: x, y, s0, s1, r, g, b, opacity
    p = np.zeros((N, 9), np.float32)
    p[:, 0] = _std_uniform_float(u[:, 0], 0.0, float(W))
    p[:, 1] = _std_uniform_float(u[:, 1], 0.0, float(H))
    p[:, 2] = np.maximum(np.float32(1.0), _std_uniform_float(u[:, 2], 0.0, 2.0))
    p[:, 3] = np.maximum(np.float32(1.0), _std_uniform_float(u[:, 3], 0.0, 2.0))
    p[:, 4] = 0.0
    p[:, 5] = _std_uniform_float(u[:, 4], 0.1, 0.2)
    p[:, 6] = _std_uniform_float(u[:, 5], 0.1, 0.2)
    p[:, 7] = _std_uniform_float(u[:, 6], 0.1, 0.2)
    p[:, 8] = _std_uniform_float(u[:, 7], 0.05, 0.1)
    return p, test_image(W, H)


def test_image(W: int, H: int) -> np.ndarray:
    x = (np.arange(W, dtype=np.float32) / np.float32(W))[None, :].repeat(H, 0)
    y = (np.arange(H, dtype=np.float32) / np.float32(H))[:, None].repeat(W, 1)
    b = np.float32(0.5) * (x + y)
    return np.stack([x, y, b], -1).reshape(W * H, 3).astype(np.float32)


test_image.__test__ = False  # not a pytest test


def splat_scene(N: int, W: int, H: int, seed: int = 0, small: bool = True):
    """A generic random scene with rotations, mixed signs of (out - target) and small Gaussians."""
    rng = np.random.default_rng(seed)
    p = np.zeros((N, 9), np.float32)
    p[:, 0] = rng.uniform(-2, W + 2, N)
    p[:, 1] = rng.uniform(-2, H + 2, N)
    p[:, 2:4] = rng.uniform(0.0, 1.4 if small else 2.0, (N, 2))
    p[:, 4] = rng.uniform(-np.pi, np.pi, N)
    p[:, 5:8] = rng.uniform(0.05, 0.9, (N, 3))
    p[:, 8] = rng.uniform(-2, 2, N)
    target = rng.uniform(0, 1.5, (W * H, 3)).astype(np.float32)
    return p, target


def covproj_inputs(n: int, seed: int = 42):
    rng = np.random.default_rng(seed)
    J = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
    W = rng.uniform(-1, 1, (n, 9)).astype(np.float32)
    A = rng.uniform(-1, 1, (n, 3, 3)).astype(np.float32)
    Sm = A @ A.transpose(0, 2, 1) + np.eye(3, dtype=np.float32)  # SPD
    S = np.stack([Sm[:, 0, 0], Sm[:, 0, 1], Sm[:, 0, 2], Sm[:, 1, 1], Sm[:, 1, 2], Sm[:, 2, 2]], -1).astype(np.float32)
    g = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    return J, W, np.ascontiguousarray(S), g


def lsq_data(n: int, seed: int = 42):
    """x1, x2 ~ U(-5, 5); y = (x1-2.5)^2 + 1.8 (x2+1.2)^2 + 0.7 + N(0, 0.5^2)
    (reference linear_regression_sgd.cu:17-20, 46-58; numpy RNG, not std::normal_distribution)."""
    rng = np.random.default_rng(seed)
    x1 = rng.uniform(-5, 5, n)
    x2 = rng.uniform(-5, 5, n)
    y = (x1 - 2.5) ** 2 + 1.8 * (x2 + 1.2) ** 2 + 0.7 + rng.normal(0, 0.5, n)
    return np.ascontiguousarray(np.stack([x1, x2, y], -1))


def accumulate_inputs(n: int, k: int, dist: str = "uniform", seed: int = 42):
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        idx = rng.integers(0, k, n, dtype=np.int32)
    elif dist == "zipf":
        w = 1.0 / np.arange(1, k + 1) ** 1.2
        idx = rng.choice(k, n, p=w / w.sum()).astype(np.int32)
    elif dist == "same":
        idx = np.full(n, k // 3, np.int32)
    else:
        raise ValueError(dist)
    val = rng.uniform(-1, 1, n).astype(np.float32)
    return idx, val
