"""Public header set (include/xyz_autodiff/, examples/.../operations/) against the REFERENCE's own headers.

tests/csrc/api_eval.inc is ONE source text written against the public API; it is compiled
  * against the reference headers through oracle/shim   -> oracle/_ref/libxyz_ref.so            (ref_*)
  * against this repo's headers by g++ for the host     -> tests/csrc/_build/libxyz_api_host.so (mine_*)
  * against this repo's headers by nvcc inside kernels  -> tests/csrc/_build/libxyz_api_cuda.so (cuda_*, -m gpu)
so "drop-in" is checked by construction (the same user code compiles against both) and by value.
Where the reference tree is absent (GPU box) the prebuilt _ref library travels with the snapshot, and the
plain-C++ restatement (oracle port, orc_*) is always available as a second checker.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "tests", "csrc", "_build", "libxyz_api_host.so")
CUDA_SO = os.path.join(ROOT, "tests", "csrc", "_build", "libxyz_api_cuda.so")


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "csrc")], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(HOST_SO):
        _build()
    return ctypes.CDLL(HOST_SO)


@pytest.fixture(scope="module")
def cuda():
    if not os.path.exists(CUDA_SO):
        _build()
    return ctypes.CDLL(CUDA_SO)


def _checkers():
    return ["port", "ref"] if orc.have_ref() else ["port"]


UNARY = ["EXP", "SIN", "COS", "SIGMOID", "SQUARED", "NEG", "L1", "L2", "SUM"]
CONST = ["ADD_C", "SUB_C", "MUL_C", "DIV_C"]
BINARY = ["ADD", "SUB", "MUL", "DIV"]
MATMUL_SHAPES = [222, 333, 444, 233, 232, 332, 423, 331, 134]


def op_cases(rng):
    """(op, in1, in2, cst, gout, aux) over the whole op table of tests/csrc/api_eval.inc."""
    cases = []
    for n in (1, 2, 3, 4, 6, 9):
        for op in UNARY:
            cases.append((op, rng.uniform(-2, 2, n), None, 0.0, rng.uniform(-1, 1, 16), 0))
        for op in CONST:
            cases.append((op, rng.uniform(-2, 2, n), None, float(rng.uniform(0.5, 3)), rng.uniform(-1, 1, 16), 0))
        for op in ("CONST_ADD", "CONST_SUB"):
            cases.append((op, rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), 0.0, rng.uniform(-1, 1, 16), 0))
    for n in (1, 2, 3, 4, 9):
        for op in BINARY:
            b = rng.uniform(0.5, 2, n) * rng.choice([-1, 1], n)
            cases.append((op, rng.uniform(-2, 2, n), b, 0.0, rng.uniform(-1, 1, 16), 0))
    for shp in MATMUL_SHAPES:
        a, b, c = shp // 100, (shp // 10) % 10, shp % 10
        cases.append(("MATMUL", rng.uniform(-1, 1, a * b), rng.uniform(-1, 1, b * c), 0.0, rng.uniform(-1, 1, 16), shp))
    cases.append(("SYM_INV", np.array([2.0, 0.3, 1.5]), None, 0.0, rng.uniform(-1, 1, 16), 0))
    cases.append(("SYM_INV", np.array([1e-5, 0.0, 1e-5]), None, 0.0, rng.uniform(-1, 1, 16), 0))  # |det| < 1e-8 (Q15)
    cases.append(("QUAT", rng.uniform(-1, 1, 4), None, 0.0, rng.uniform(-1, 1, 16), 0))
    cases.append(("BROADCAST3", rng.uniform(-1, 1, 1), None, 0.0, rng.uniform(-1, 1, 16), 0))
    cases.append(("COV_GEN", rng.uniform(0.5, 2, 2), rng.uniform(-3, 3, 1), 0.0, rng.uniform(-1, 1, 16), 0))
    cases.append(("MAT_TO_COV3", rng.uniform(-2, 2, 4), None, 0.0, rng.uniform(-1, 1, 16), 0))
    cases.append(("SCALE_ROT_COV3", rng.uniform(0.5, 2, 2), rng.uniform(-3, 3, 1), 0.0, rng.uniform(-1, 1, 16), 0))
    cases.append(("MAHALANOBIS", rng.uniform(-2, 2, 2), rng.uniform(0.1, 3, 3), 0.0, rng.uniform(-1, 1, 16), 0))
    cases.append(("MAHALANOBIS_CENTER", rng.uniform(0.1, 3, 2), rng.uniform(0.1, 3, 3), 1.7, rng.uniform(-1, 1, 16), 0))
    # kinks: l1 at exactly 0 (subgradient 0), l2 of the zero vector (no adjoint)
    cases.append(("L1", np.array([0.0, -1.0, 2.0]), None, 0.0, rng.uniform(-1, 1, 16), 0))
    cases.append(("L2", np.zeros(3), None, 0.0, rng.uniform(-1, 1, 16), 0))
    return cases


def compare_ops(fn64, fn32, tol64, tol32):
    n = 0
    for which in _checkers():
        rng = np.random.default_rng(123)
        for op, in1, in2, cst, gout, aux in op_cases(rng):
            for dtype, fn, tol in ((np.float64, fn64, tol64), (np.float32, fn32, tol32)):
                want = orc.eval_op(which, op, in1, in2, cst, gout, aux, dtype)
                got = orc.eval_op(which, op, in1, in2, cst, gout, aux, dtype, fn=fn)
                for a, b, what in zip(got, want, ("value", "grad1", "grad2")):
                    scale = np.maximum(np.abs(b), np.abs(b).max() if b.size else 1.0) + 1e-30
                    assert (np.abs(a - b) <= tol * scale).all(), (which, op, aux, dtype.__name__, what, a, b)
                n += 1
    return n


def test_host_headers_match_reference_ops(host):
    """Every op of the public header set, host-compiled, fp64 and fp32, against the reference's Logic structs
    (and the port): same formulas in the same order -> agreement to a few ulp."""
    n = compare_ops(host.mine_eval_op_f64, host.mine_eval_op_f32, 1e-14, 1e-6)
    assert n > 300


def test_host_known_answers(host):
    """The reference tests' known answers on the host path."""
    res = np.zeros(64)
    k = host.mine_kat_dag(P(res))
    assert list(res[:k]) == [6.0, 4.0, 9.0, 4.0, 12.0, 7.0]                # tests/test_dag_backward.cu:80,130,171
    k = host.mine_kat_shared_subgraph(P(res))
    assert res[1] == 0.0 and abs(res[3] - 2 * np.exp(0.5)) < 1e-12         # Q3: deep-shared stops, shallow works
    k = host.mine_kat_broadcast(P(res))
    assert list(res[:5]) == [3.5] * 4 + [10.0] and list(res[5:14]) == [-2.25] * 8 + [4.0]
    assert list(res[14:21]) == [3.0, 5.0, 7.0, 3.0, 1.0, 1.0, 1.0]         # tests/operation/unary/test_broadcast.cu
    host.mine_kat_chain.argtypes = [ctypes.c_double] * 4 + [ctypes.c_void_p]
    k = host.mine_kat_chain(1.5, -0.7, 2.25, 0.3, P(res))                  # f = x z + y: dx = g z, dy = g, dz = g x
    assert np.allclose(res[:4], [1.5 * 2.25 - 0.7, 0.3 * 2.25, 0.3, 0.3 * 1.5], rtol=1e-15)
    fres = np.zeros(64, np.float32)
    k = host.mine_kat_matrices(P(fres))
    assert list(fres[:12]) == [1, 2, 6, 8, 15, 18, 1, 4, 9, 4, 10, 18]      # tests/test_diagonal_matrix.cu:145,167
    assert list(fres[12:21]) == [1, 2, 3, 4, 5, 6, 2, 3, 5]                 # tests/test_symmetric_matrix.cu
    assert list(fres[21:27]) == [1, 4, 2, 5, 3, 6]                          # tests/test_matrix_transpose.cu
    # accum::slice views of one RegisterLeaf: same gradients as the VariableRef graph, accumulated over 3 evaluations
    host.mine_kat_leaf_slices.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 3 + [ctypes.c_int, ctypes.c_void_p]
    prm = np.array([1.0, 1.5, 0.5, 0.2])
    sl = np.zeros(9)
    assert host.mine_kat_leaf_slices(P(prm), 2.0, 3.0, 5.0, 3, P(sl)) == 9
    one = np.zeros(9)
    host.mine_kat_lsq_point.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4 + [ctypes.c_void_p]
    host.mine_kat_lsq_point(P(prm), 2.0, 3.0, 5.0, 1e-6, P(one))
    assert sl[0] == one[0] and np.allclose(sl[1:5], 3 * one[1:5], rtol=1e-15)
    assert sl[5] == sl[1] and sl[6] == 0.0 and sl[7] == 0.0 and sl[8] == sl[4]
    host.mine_kat_ternary.argtypes = [ctypes.c_double] * 3 + [ctypes.c_void_p]
    k = host.mine_kat_ternary(1.5, -2.0, 0.25, P(res))
    assert np.allclose(res[:4], [1.5 * -2.0 + 0.25, -2.0, 1.5, 1.0]) and np.allclose(res[4:7], res[1:4], atol=1e-8)


def _all_kats(run):
    out = {}
    out["dag"] = run("dag")
    out["shared_subgraph"] = run("shared_subgraph")
    out["broadcast"] = run("broadcast")
    out["chain"] = run("chain", 1.5, -0.7, 2.25, 0.3)
    out["operators"] = run("operators", np.array([0.5, -1.5]), np.array([2.0, 0.25]), np.array([0.1, 0.2]),
                           np.array([1.5, -3.0]))
    out["lsq_point"] = run("lsq_point", np.array([1.0, 1.5, 0.5, 0.2]), 2.0, 3.0, 5.0, 1e-6)
    out["splat_pair"] = run("splat_pair", np.array([0.5, 0.3, 0.2, 0.4, 0.3, 0.8, 0.4, 0.16, 0.1, 1.0, 0.7]))
    out["math_f64"] = run("math_f64", 0.73)
    return out


def _kat_runner(lib, prefix):
    def run(name, *args):
        res = np.zeros(64)
        cargs = [P(a) if isinstance(a, np.ndarray) else ctypes.c_double(a) for a in args]
        k = getattr(lib, f"{prefix}_kat_{name}")(*cargs, P(res))
        return res[:k].copy()
    return run


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref/libxyz_ref.so not present")
def test_host_graphs_match_reference_graphs(host):
    """Whole graphs (DAG protocol, operator sugar, least-squares point analytic + numerical, the splat pair
    graph, math dispatcher, matrix views): this repo's headers vs the reference's headers, same source text."""
    mine = _all_kats(_kat_runner(host, "mine"))
    ref = _all_kats(_kat_runner(orc.load("ref"), "ref"))
    for name in mine:
        assert mine[name].size == ref[name].size and mine[name].size > 0, name
        assert np.allclose(mine[name], ref[name], rtol=1e-13, atol=1e-15), name
    # analytic vs numerical on the least-squares point: the reference's acceptance rule min(abs, rel) <= 1e-5
    a, n = mine["lsq_point"][1:5], mine["lsq_point"][5:9]
    assert (np.minimum(np.abs(a - n), np.abs(a - n) / (np.abs(a) + 1e-15)) <= 1e-5).all()
    fa, fb = np.zeros(64, np.float32), np.zeros(64, np.float32)
    ka = host.mine_kat_matrices(P(fa))
    kb = orc.load("ref").ref_kat_matrices(P(fb))
    assert ka == kb and np.array_equal(fa[:ka], fb[:kb])
    host.mine_kat_math_f32.argtypes = [ctypes.c_float, ctypes.c_void_p]
    orc.load("ref").ref_kat_math_f32.argtypes = [ctypes.c_float, ctypes.c_void_p]
    ka = host.mine_kat_math_f32(0.73, P(fa))
    kb = orc.load("ref").ref_kat_math_f32(0.73, P(fb))
    assert ka == kb == 20 and np.allclose(fa[:ka], fb[:kb], rtol=1e-6)


def _broadcast_logic_expected(x0, up, s0, v):
    return [x0] * 4 + [sum(up), 5.0, 5.0] + [s0 * t for t in v] + [sum(v)] + [s0] * 3


def test_materialising_broadcast_value_and_gradient(host):
    """xyz_autodiff::broadcast<N> (BroadcastLogic, reference operations/unary/broadcast_logic.cuh:11-46)
    -- values, summed adjoint, run(), run_numerical() and use inside a graph -- against the closed form and, where the
    reference is built, against the reference's own header evaluating the same source text."""
    vp, dbl = ctypes.c_void_p, ctypes.c_double
    host.mine_kat_broadcast_logic.argtypes = [dbl, vp, dbl, vp, vp]
    up, v = np.array([1.0, -2.5, 0.25, 4.0]), np.array([1.0, 3.0, 5.0])
    res = np.zeros(64)
    k = host.mine_kat_broadcast_logic(3.5, P(up), 2.0, P(v), P(res))
    want = _broadcast_logic_expected(3.5, list(up), 2.0, list(v))
    assert k == len(want) == 14
    assert np.allclose(res[:k], want, rtol=0, atol=1e-9) and list(res[:6]) == want[:6]  # only run_numerical is inexact
    if orc.have_ref():
        R = orc.load("ref")
        R.ref_kat_broadcast_logic.argtypes = [dbl, vp, dbl, vp, vp]
        ref = np.zeros(64)
        assert R.ref_kat_broadcast_logic(3.5, P(up), 2.0, P(v), P(ref)) == k
        assert np.array_equal(ref[:k], res[:k])


# ---- device side ----------------------------------------------------------------------------------------
def test_ternary_covariance_projection_node(host):
    """op::covariance_projection (one ternary node, not in the reference) against the oracle's composition of the
    reference's own op::matmul nodes (tree with transposed leaves): fp64 to 1e-12, fp32 to 1e-5 of the row magnitude; and
    the node's analytic adjoints against its own central differences (run_numerical)."""
    vp = ctypes.c_void_p
    host.mine_covproj_ternary_f64.argtypes = [vp] * 8
    host.mine_covproj_ternary_f32.argtypes = [vp] * 8
    host.mine_covproj_ternary_numerical.argtypes = [vp] * 4
    n = 200
    J, W, S, g = orc.covproj_inputs(n, seed=31)
    for which in _checkers():
        for dtype, fn, tol in ((np.float64, host.mine_covproj_ternary_f64, 1e-12), (np.float32, host.mine_covproj_ternary_f32, 1e-5)):
            want = orc.covproj(J, W, S, g, np.float64, which=which)
            for e in range(n):
                ins = [np.ascontiguousarray(a[e], dtype) for a in (J, W, S, g)]
                outs = [np.zeros(k, dtype) for k in (3, 6, 9, 6)]
                assert fn(*[P(a) for a in ins], *[P(a) for a in outs]) == 0
                for a, b in zip(outs, want):
                    assert (np.abs(a - b[e]) <= tol * (np.abs(b[e]).max() + 1e-30)).all(), (which, dtype.__name__, e)
    res = np.zeros(42)
    Jd, Wd, Sd = (np.ascontiguousarray(a[0], np.float64) for a in (J, W, S))
    assert host.mine_covproj_ternary_numerical(P(Jd), P(Wd), P(Sd), P(res)) == 42
    assert (np.abs(res[:21] - res[21:]) <= 1e-5 * np.maximum(1.0, np.abs(res[:21]))).all()


VARIABLE_KNOWN = ([1, 2, 3, 4, 0, 2, 4, 6] + [0, 0, 0] + [10, 11, 12, 0, 3, 6]          # tests/test_variable.cu:91-181
                  + [7, -3, 0, 2.5] + [4, 10, 4, 10])


def test_variable_leaves_known_answers(host):
    """VariableRef (external buffers, data() / grad(), zero_grad, add_grad) and Variable (own
    buffers, ref() view) -- the reference tests' known answers (tests/test_variable.cu) and bit for bit the reference's
    header on the same source text."""
    res = np.zeros(64)
    k = host.mine_kat_variable(P(res))
    assert k == 25 and list(res[:k]) == VARIABLE_KNOWN
    if orc.have_ref():
        rb = np.zeros(64)
        assert orc.load("ref").ref_kat_variable(P(rb)) == k and np.array_equal(res[:k], rb[:k])


def test_matrix_types_more_known_answers(host):
    """SURVEY 8 row a20, second helping: DenseMatrix as a Variable (flat access, gradients start at 0, zero_grad keeps the
    values, (i, j) == flat, data() / grad()), transpose twice, and the transpose of a diagonal view -- the known answers
    of tests/test_dense_matrix.cu:99-185 and tests/test_matrix_transpose.cu:111-177; host = the reference's headers."""
    fa = np.zeros(64, np.float32)
    k = host.mine_kat_matrices2(P(fa))
    want = ([1, 2, 3, 4, 5, 6] + [0] * 6 + [1, 2, 3, 2, 4, 6] + [0] * 6 + [1, 1, 1]
            + [1, 4, 2, 5, 3, 6, 1, 6] + [1, 2, 3, 0, 0])
    assert k == 40 and list(fa[:k]) == want
    if orc.have_ref():
        fb = np.zeros(64, np.float32)
        assert orc.load("ref").ref_kat_matrices2(P(fb)) == k and np.array_equal(fa[:k], fb[:k])
    # diagonal / symmetric views the way the reference's tests use them (default template argument, make_* helpers,
    # flat-index storage, symmetric transpose): tests/test_diagonal_matrix.cu:105-187, tests/test_symmetric_matrix.cu:118-192
    k3 = host.mine_kat_matrices3(P(fa))
    want3 = ([1, 0, 0, 0, 2, 0, 0, 0, 3] + [1, 2, 6, 8, 15, 18] + [1, 4, 9, 4, 10, 18] + [1, 2, 2, 4, 3, 6, 2, 5]
             + [1, 2, 2, 4, 5, 5] + [1, 2, 3, 4, 5, 6, 2, 3, 5])
    assert k3 == 44 and list(fa[:k3]) == want3
    if orc.have_ref():
        fb = np.zeros(64, np.float32)
        assert orc.load("ref").ref_kat_matrices3(P(fb)) == k3 and np.array_equal(fa[:k3], fb[:k3])


NETWORK_CASES = [  # (a, b, c, d, x1, x2, y_target), delta, tolerance: tests/operation/base/test_linear_regression_network.cu:186-325
    ((1.0, 1.5, 0.5, 0.2, 2.0, 1.5, 3.0), 1e-6, 1e-5),               # SimpleLinearRegressionNetwork
    ((0.8, 1.2, -0.3, 0.5, 1.0, 0.8, 2.5), 1e-6, 1e-5),              # RegularizedLinearRegressionNetwork
    ((1.0, 0.0, 0.0, 0.0, 2.0, 0.0, 0.0), 1e-7, 1e-5),               # SubtractSquareOperation
    ((0.5, 0.8, 0.3, 0.2, 1.0, 0.5, 1.5), 1e-6, 1e-5),               # DuplicateAddNetwork
    ((10.0, 15.0, -8.0, 5.0, 20.0, -10.0, 100.0), 1e-5, 1e-3),       # StressTestWithLargeValues
    ((0.001, 0.002, -0.001, 0.0001, 0.01, -0.01, 0.001), 1e-8, 1e-5),  # EdgeCaseNearZero
]


def test_reference_network_graphs_analytic_and_numerical(host):
    """The networks of the reference's network-gradient tests, written with the operator sugar exactly as there: the
    duplicate-add graph (both leaves feed two nodes: gradients 2 and 2), (a - x1)^2, and the least-squares loss --
    closed form, analytic against numerical under the reference's acceptance rule min(abs, rel) <= tolerance, and this
    repo's headers against the reference's on the same source text."""
    vp, dbl = ctypes.c_void_p, ctypes.c_double
    host.mine_kat_networks.argtypes = [vp, dbl, vp]
    ref = orc.load("ref") if orc.have_ref() else None
    if ref is not None:
        ref.ref_kat_networks.argtypes = [vp, dbl, vp]
    for prm, delta, tol in NETWORK_CASES:
        p, res = np.array(prm), np.zeros(32)
        assert host.mine_kat_networks(P(p), delta, P(res)) == 17
        a, b, c, d, x1, x2, y = prm
        assert res[0] == 2 * (a + b) and list(res[1:3]) == [2.0, 2.0]
        assert np.isclose(res[5], (a - x1) ** 2, rtol=1e-15) and np.isclose(res[6], 2 * (a - x1), rtol=1e-15)
        r = (a - x1) ** 2 + b * (c - x2) ** 2 + d - y
        want = 2 * r * np.array([2 * (a - x1), (c - x2) ** 2, 2 * b * (c - x2), 1.0])
        assert np.isclose(res[8], r * r, rtol=1e-12) and np.allclose(res[9:13], want, rtol=1e-12, atol=1e-300)
        for ana, num in ((res[1:3], res[3:5]), (res[6:7], res[7:8]), (res[9:13], res[13:17])):
            err = np.abs(ana - num)
            assert (np.minimum(err, err / (np.abs(ana) + 1e-15)) <= tol).all(), (prm, ana, num)
        if ref is not None:
            rb = np.zeros(32)
            assert ref.ref_kat_networks(P(p), delta, P(rb)) == 17
            assert np.allclose(res[:17], rb[:17], rtol=1e-13, atol=1e-300)


CONST_ARRAY_KNOWN = ([1, 2, 3, 3] + [10, 20, 30] + [100, 200, 300]                        # tests/test_const_array.cu:167-217
                     + [10, 20, 30, 15, 35, 55, 10, 20, 30, 20, 50, 80]                   # :247-305
                     + [150, 250, 350, 100, 200, 300, 160, 260, 360]                      # :307-333
                     + [10, 20, 30, 40, 4, 1, 2, 3, 4, 5, 5])                             # :219-245


def test_const_array_known_answers(host):
    """ConstArray<T, N> (reference const_array.cuh:7-223) -- element access, array constructor, copy,
    compound and binary operators between arrays and with any ConstArrayLike type on either side: the reference tests'
    own known answers (tests/test_const_array.cu), and bit for bit the reference's header on the same source text."""
    fa = np.zeros(64, np.float32)
    ka = host.mine_kat_const_array(P(fa))
    assert ka == 60 and list(fa[:42]) == CONST_ARRAY_KNOWN
    a1, a2 = np.array([1.5, -2.0, 8.0], np.float32), np.array([4.0, 0.25, -3.0], np.float32)
    half = np.float32(0.5)
    want = np.concatenate([a1 * a2, a1 * a2 / half, a1 * a2, a1 / a2, half * a1, half / a2])
    assert np.array_equal(fa[42:60], want.astype(np.float32))
    if orc.have_ref():
        fb = np.zeros(64, np.float32)
        kb = orc.load("ref").ref_kat_const_array(P(fb))
        assert kb == ka and np.array_equal(fa[:ka], fb[:kb])


@pytest.mark.gpu
def test_device_headers_match_reference_ops(cuda):
    """The same op table evaluated INSIDE a kernel on the B200 against the reference's host-compiled Logic
    structs: fp64 to 1e-12 (FMA contraction + libdevice transcendentals), fp32 to 1e-5 relative."""
    n = compare_ops(cuda.cuda_eval_op_f64, cuda.cuda_eval_op_f32, 1e-12, 1e-5)
    assert n > 300


@pytest.mark.gpu
def test_device_known_answers(cuda, host):
    res, fres, inp, want = np.zeros(64), np.zeros(64, np.float32), np.zeros(16), np.zeros(64)
    k = cuda.cuda_kat(0, None, P(res), None)
    assert list(res[:k]) == [6.0, 4.0, 9.0, 4.0, 12.0, 7.0]
    k = cuda.cuda_kat(1, None, P(res), None)
    assert res[1] == 0.0 and abs(res[3] - 2 * np.exp(0.5)) < 1e-12
    k = cuda.cuda_kat(2, None, P(res), None)
    assert list(res[:5]) == [3.5] * 4 + [10.0] and list(res[14:21]) == [3.0, 5.0, 7.0, 3.0, 1.0, 1.0, 1.0]
    k = cuda.cuda_kat(9, None, None, P(fres))
    assert list(fres[:12]) == [1, 2, 6, 8, 15, 18, 1, 4, 9, 4, 10, 18] and list(fres[21:27]) == [1, 4, 2, 5, 3, 6]
    # graphs with inputs: device vs this repo's host path (which test_host_graphs pins to the reference)
    inp[:11] = [0.5, 0.3, 0.2, 0.4, 0.3, 0.8, 0.4, 0.16, 0.1, 1.0, 0.7]
    k = cuda.cuda_kat(6, P(inp), P(res), None)
    host.mine_kat_splat_pair(P(inp[:11].copy()), P(want))
    assert k == 13 and np.allclose(res[:k], want[:k], rtol=1e-12)
    inp[:8] = [1.0, 1.5, 0.5, 0.2, 2.0, 3.0, 5.0, 1e-6]
    k = cuda.cuda_kat(5, P(inp), P(res), None)
    r = (1.0 - 2.0) ** 2 + 1.5 * (0.5 - 3.0) ** 2 + 0.2 - 5.0
    assert np.allclose(res[1:5], 2 * r * np.array([2 * (1.0 - 2.0), (0.5 - 3.0) ** 2, 2 * 1.5 * (0.5 - 3.0), 1.0]),
                       rtol=1e-12)
    assert np.allclose(res[5:9], res[1:5], rtol=1e-5)                       # numerical backward on the device
    inp[0] = 0.73
    k = cuda.cuda_kat(7, P(inp), P(res), None)
    host.mine_kat_math_f64.argtypes = [ctypes.c_double, ctypes.c_void_p]
    host.mine_kat_math_f64(0.73, P(want))
    assert k == 20 and np.allclose(res[:k], want[:k], rtol=1e-13)
    # a15b on the device: the materialising broadcast
    inp[:9] = [3.5, 1.0, -2.5, 0.25, 4.0, 2.0, 1.0, 3.0, 5.0]
    k = cuda.cuda_kat(10, P(inp), P(res), None)
    want14 = _broadcast_logic_expected(3.5, [1.0, -2.5, 0.25, 4.0], 2.0, [1.0, 3.0, 5.0])
    assert k == 14 and np.allclose(res[:k], want14, rtol=0, atol=1e-9) and list(res[:6]) == want14[:6]
    # a19 on the device: ConstArray (the reference's device kernels, tests/test_const_array.cu:85-165) = the host path
    k = cuda.cuda_kat(11, None, None, P(fres))
    fwant = np.zeros(64, np.float32)
    assert k == host.mine_kat_const_array(P(fwant)) == 60
    assert list(fres[:42]) == CONST_ARRAY_KNOWN and np.array_equal(fres[:k], fwant[:k])
    k = cuda.cuda_kat(12, None, P(res), None)                               # a1 / a2 on the device (atomic add_grad)
    assert k == 25 and list(res[:k]) == VARIABLE_KNOWN


@pytest.mark.gpu
def test_device_concurrent_accumulation_known_answers(cuda):
    """The reference's concurrency tests (Appendix D): VariableRef::add_grad from 10 000 / 100 000 threads onto 3
    fp64 addresses, and onto __shared__ memory; then accumulate.cuh's on-chip pre-reduction on the same sums."""
    g = np.zeros(3)
    cuda.cuda_global_accumulation.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
    assert cuda.cuda_global_accumulation(10_000, 256, P(g)) == 0
    tid = np.arange(10_000)
    assert abs(g[0] - 10000.0) < 1e-10 and abs(g[1] - 10000.0) < 1e-10      # test_parallel_gradient_accumulation.cu:94-101
    assert abs(g[2] - ((1.0 + tid * 0.001) + (2.0 + tid * 0.001)).sum()) < 1e-6
    assert cuda.cuda_global_accumulation(100_000, 512, P(g)) == 0
    assert g[0] == 100000.0 and g[1] == 100000.0                             # :162-167
    out = np.zeros(8, np.float32)
    assert cuda.cuda_shared_accumulation(1, 128, P(out)) == 0
    assert out[0] == 128.0 and out[1] == 256.0                               # test_shared_memory_atomic.cu:102-109
    assert cuda.cuda_shared_accumulation(4, 64, P(out)) == 0
    assert list(out) == [64.0, 128.0] * 4                                    # :179-186
    # RegisterLeaf + block_accumulate: the least-squares graph through the PUBLIC headers, one RED per CTA
    data = orc.lsq_data(100_000, seed=3)
    values = np.array([0.3, 1.2, -0.4, 0.1])
    grads = np.zeros(4)
    cuda.cuda_lsq_register_leaf.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p]
    assert cuda.cuda_lsq_register_leaf(P(data), 100_000, P(values), P(grads)) == 0
    want, _ = orc.lsq_grad(data, values)
    assert (np.abs(grads - want) <= 1e-10 * np.abs(want)).all()
    vals = np.random.default_rng(0).uniform(-1, 1, 100_000).astype(np.float32)
    tgt = np.zeros(8, np.float32)
    assert cuda.cuda_block_accumulate_f32(P(vals), vals.size, P(tgt)) == 0
    assert np.allclose(tgt, vals.astype(np.float64).sum() * np.arange(1, 9), rtol=1e-4, atol=1e-2)


@pytest.mark.gpu
def test_gradient_testers_rehosted_without_gtest():
    """include/xyz_autodiff/testing.cuh (the reference's Unary/Binary/NetworkGradientTester API, gtest-free, one launch
    for all random cases) over the op table and the splat example's custom Logics: every Logic passes at the
    reference's 1e-5 rule, a deliberately wrong Logic and a forbidden tolerance are rejected."""
    path = os.path.join(os.path.dirname(__file__), "csrc", "_build", "libxyz_testers.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(os.path.dirname(__file__), "csrc")], check=True, stdout=subprocess.DEVNULL)
    L = ctypes.CDLL(path)
    buf = ctypes.create_string_buffer(1 << 16)
    code = L.tester_run_all(buf, len(buf))
    text = buf.value.decode()
    assert code // 1000 == 0, text
    assert code % 1000 >= 27, text
    assert "FAIL" not in text
