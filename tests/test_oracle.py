"""Pins the CPU oracle port (oracle/xyz_oracle.cpp) before anything else trusts it:
  (1) against the committed golden vectors generated from the reference itself (tests/golden/);
  (2) live against oracle/_ref/libxyz_ref.so (the reference's own code on the host) where present;
  (3) against the reference tests' known answers.
Both libraries are built with -O2 -ffp-contract=off, so fp32 and fp64 comparisons are BIT-EXACT when both
run single-threaded."""
import os

import numpy as np
import pytest

import oracle_lib as orc
from test_api_headers import op_cases

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
need_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref/libxyz_ref.so not present")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_port_splat_equals_golden_bit_for_bit(tag):
    W, H, N, _ = GOLD[f"splat_{tag}_shape"]
    g, o, l, _ = orc.splat(GOLD[f"splat_{tag}_params"], GOLD[f"splat_{tag}_target"], int(W), int(H), np.float32, threads=1)
    assert np.array_equal(g, GOLD[f"splat_{tag}_grads"])
    assert np.array_equal(o, GOLD[f"splat_{tag}_output"])
    assert np.float32(l) == GOLD[f"splat_{tag}_loss"]


def test_port_splat_c4_distribution_equals_golden():
    params, target = orc.splat_c4_scene(64, 48, 32, seed=42)  # std::mt19937(42) emulation is part of what is pinned
    assert np.array_equal(params, GOLD["splat_c4_params"]) and np.array_equal(target, GOLD["splat_c4_target"])
    g, o, l, _ = orc.splat(params, target, 48, 32, np.float32, threads=1)
    assert np.array_equal(g, GOLD["splat_c4_grads"]) and np.array_equal(o, GOLD["splat_c4_output"])
    assert np.float32(l) == GOLD["splat_c4_loss"]
    assert orc.std_mt19937_u32(42, 2).tolist() == [1608637542, 3421126067]  # first outputs of std::mt19937(42)


def test_port_lsq_accumulate_covproj_equal_golden():
    for ro in (0, 1):
        g, l = orc.lsq_grad(GOLD["lsq_data"], (0.0, 1.0, 0.0, 0.0), bool(ro), threads=1)
        assert np.array_equal(g, GOLD[f"lsq_grad_{ro}"]) and l == GOLD[f"lsq_loss_{ro}"]
    assert np.array_equal(orc.accumulate(GOLD["acc_idx"], GOLD["acc_val"], 64), GOLD["acc_grad_f32"])
    assert np.array_equal(orc.accumulate(GOLD["acc_idx"], GOLD["acc_val"].astype(np.float64), 64), GOLD["acc_grad_f64"])
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        outs = orc.covproj(GOLD["cov_J"], GOLD["cov_W"], GOLD["cov_S"], GOLD["cov_g"], dt, threads=1)
        for name, a in zip(("out", "gJ", "gW", "gS"), outs):
            assert np.array_equal(a, GOLD[f"cov_{tag}_{name}"]), (tag, name)


def test_port_op_table_equals_golden():
    rng = np.random.default_rng(123)
    rows = []
    for op, in1, in2, cst, gout, aux in op_cases(rng):
        o, g1, g2 = orc.eval_op("port", op, in1, in2, cst, gout, aux, np.float64)
        rows.append(np.concatenate([o, g1, g2]))
    assert np.array_equal(np.array([r.size for r in rows]), GOLD["op_table_rows"])
    got, want = np.concatenate(rows), GOLD["op_table_f64"]
    # (g*c)*v vs g*(c*v) in the quaternion rule is the only reassociation: 1 ulp
    assert np.allclose(got, want, rtol=1e-15, atol=1e-300)
    assert (got == want).mean() > 0.99


def test_threaded_oracle_matches_single_thread():
    params, target = orc.splat_scene(30, 80, 64, seed=4)
    g1, o1, l1, _ = orc.splat(params, target, 80, 64, np.float64, threads=1)
    g8, o8, l8, _ = orc.splat(params, target, 80, 64, np.float64, threads=8)
    assert np.array_equal(o1, o8) and np.allclose(g1, g8, rtol=1e-12) and abs(l1 - l8) <= 1e-12 * l1
    data = orc.lsq_data(10_000, seed=2)
    a, _ = orc.lsq_grad(data, (0.3, 1.2, -0.4, 0.1), threads=1)
    b, _ = orc.lsq_grad(data, (0.3, 1.2, -0.4, 0.1), threads=8)
    assert np.allclose(a, b, rtol=1e-12)


def test_known_answers_from_reference_tests():
    # tests/test_dag_backward.cu:80,130,171 ; tests/operation/unary/test_broadcast.cu ; closed-form least squares
    assert GOLD["kat_dag"].tolist() == [6.0, 4.0, 9.0, 4.0, 12.0, 7.0]
    assert GOLD["kat_shared_subgraph"][1] == 0.0
    assert GOLD["kat_broadcast"][4] == 10.0 and GOLD["kat_broadcast"][13] == 4.0
    for (x1, x2, y), (a, b, c, d) in [((2.0, 3.0, 5.0), (1.0, 1.5, 0.5, 0.2)), ((-2.0, -1.5, 3.0), (-1.0, 2.0, -0.5, -0.3)),
                                      ((100.0, 150.0, 500.0), (50.0, 30.0, 70.0, 20.0))]:
        r = (a - x1) ** 2 + b * (c - x2) ** 2 + d - y
        want = 2 * r * np.array([2 * (a - x1), (c - x2) ** 2, 2 * b * (c - x2), 1.0])
        got, loss = orc.lsq_grad(np.array([[x1, x2, y]]), (a, b, c, d))
        assert np.allclose(got, want, rtol=1e-13) and np.isclose(loss, r * r, rtol=1e-13)
    # accumulation: 10 000 x (+1, +1, 3 + 0.002 tid), tests/test_parallel_gradient_accumulation.cu:94-110
    tid = np.arange(10_000)
    idx = np.tile(np.array([0, 1, 2], np.int32), 10_000)
    val = np.stack([np.ones(10_000), np.ones(10_000), (1.0 + tid * 0.001) + (2.0 + tid * 0.001)], -1).reshape(-1)
    got = orc.accumulate(idx, val, 3)
    assert got[0] == 10000.0 and got[1] == 10000.0 and abs(got[2] - val[2::3].sum()) < 1e-6
    # splat quirks Q1/Q2: a far-away Gaussian is NOT culled (all-pairs), and the L1 argument is wc + out - target
    p = np.array([[5.0, 5.0, 0.0, 0.0, 0.0, 0.5, 0.5, 0.5, 0.0]], np.float32)
    t = np.full((16 * 16, 3), 10.0, np.float32)                       # target > out everywhere -> sign is -1
    g, o, l, _ = orc.splat(p, t, 16, 16, np.float64)
    assert (g[0, 5:8] < 0).all() and o.max() < 0.26


@need_ref
def test_port_equals_reference_live_bit_for_bit():
    params, target = orc.splat_scene(40, 70, 50, seed=9)
    a = orc.splat(params, target, 70, 50, np.float32, which="ref", threads=1)
    b = orc.splat(params, target, 70, 50, np.float32, which="port", threads=1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    data = orc.lsq_data(5000, seed=8)
    for ro in (False, True, 2):  # 2 = the shipped example's graph, built from the reference's own op:: factories in _ref
        assert np.array_equal(orc.lsq_grad(data, (0.1, 0.9, -0.2, 0.3), ro, "ref")[0],
                              orc.lsq_grad(data, (0.1, 0.9, -0.2, 0.3), ro, "port")[0])
    J, W, S, g = orc.covproj_inputs(3000, seed=1)
    for dt in (np.float32, np.float64):
        for x, y in zip(orc.covproj(J, W, S, g, dt, "ref", 1), orc.covproj(J, W, S, g, dt, "port", 1)):
            assert np.array_equal(x, y)
    idx, val = orc.accumulate_inputs(20_000, 1024, "zipf", seed=3)
    assert np.array_equal(orc.accumulate(idx, val, 1024, "ref"), orc.accumulate(idx, val, 1024, "port"))


def test_tile_binning_restatement_properties():
    """The integer oracle (this repo's own; 'parity unpinned' by the reference, which has no binning): lists are
    ascending per tile, ranges partition the list, and every pair with a non-zero fp32 weight is inside its
    Gaussian's rectangle -- the property that makes the cull result-preserving."""
    W, H, N = 200, 150, 300
    params, _ = orc.splat_scene(N, W, H, seed=5)
    rec = orc.splat_records(params)
    rects, ranges, ids = orc.splat_binning(rec, W, H, d2max=176.0)
    assert ranges[0, 0] == 0 and ranges[-1, 1] == ids.size and (ranges[1:, 0] == ranges[:-1, 1]).all()
    for t in range(ranges.shape[0]):
        seg = ids[ranges[t, 0]:ranges[t, 1]]
        assert (np.diff(seg) > 0).all()
    ys, xs = np.mgrid[0:H, 0:W]
    for g in range(0, N, 7):
        cx, cy, ia, ib, ic = rec[g, :5]
        dx, dy = xs.astype(np.float32) - cx, ys.astype(np.float32) - cy
        d2 = ia * dx * dx + 2 * ib * dx * dy + ic * dy * dy
        nz = np.exp(-(0.5 * d2.astype(np.float64))) > 2.0 ** -126      # FTZ threshold of the fast-math flavour
        tx0, ty0, tx1, ty1 = rects[g]
        inside = (xs >= tx0 * 16) & (xs < tx1 * 16) & (ys >= ty0 * 16) & (ys < ty1 * 16)
        assert not (nz & ~inside).any()


@pytest.mark.parametrize("seed,small", [(5, True), (6, False), (7, True)])
def test_tile_row_spans_keep_every_nonzero_pair(seed, small):
    """Second cull level (per tile row, the span the ellipse can reach): every (pixel, Gaussian) pair whose fp32
    weight is not exactly zero lies in a tile that lists the Gaussian; the lists are a strict subset of the
    rectangles for anisotropic / rotated Gaussians; d2max = 56 (XYZ_FLAG_TAIL_CULL) keeps every pair above e^-28."""
    W, H, N = 230, 170, 160
    params, _ = orc.splat_scene(N, W, H, seed=seed, small=small)
    params[:, 2] += 0.4                                   # anisotropic
    rec = orc.splat_records(params)
    tiles_x = (W + 15) // 16
    ys, xs = np.mgrid[0:H, 0:W]
    tile_of = (ys // 16) * tiles_x + xs // 16
    for d2max, thresh in ((176.0, 2.0 ** -126), (56.0, np.exp(-28.0))):
        rects, ranges, ids = orc.splat_binning(rec, W, H, d2max=d2max)
        member = np.zeros((ranges.shape[0], N), bool)
        for t in range(ranges.shape[0]):
            member[t, ids[ranges[t, 0]:ranges[t, 1]]] = True
        rect_tiles = ((rects[:, 2] - rects[:, 0]) * (rects[:, 3] - rects[:, 1])).sum()
        assert ids.size < rect_tiles
        for g in range(N):
            cx, cy, ia, ib, ic = rec[g, :5].astype(np.float64)
            dx, dy = xs - cx, ys - cy
            d2 = ia * dx * dx + 2 * ib * dx * dy + ic * dy * dy
            nz = np.exp(-0.5 * d2) > thresh
            assert member[tile_of[nz], g].all(), (g, d2max)


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref/libxyz_ref.so not built")
@pytest.mark.parametrize("case", range(6))
def test_fp32_conditioning_bound_holds_for_the_reference_kernel(case):
    """Ill-conditioned scenes (sub-pixel / strongly anisotropic Gaussians): the reference's OWN kernel body, evaluated in
    fp32 on the host, misses the plain 1e-5 / 1e-4 tolerances against fp64 by up to 14x -- the exponent amplifies input
    rounding by its term magnitudes and by the conditioning of the 2x2 inverse.  splat_tolerance_fp32 adds exactly that
    sensitivity (computed in fp64 by the oracle); the reference's fp32 results must lie inside it, which makes it the bar
    any fp32 implementation is held to on such scenes (tests/test_gpu_parity.py applies it to the CUDA path)."""
    params, target, W, H = orc.splat_hard_scene(case)
    g, o, l, tol_g, tol_i = orc.splat_tolerance_fp32(params, target, W, H)
    g32, o32, l32, _ = orc.splat(params, target, W, H, np.float32, which="ref")
    assert (np.abs(o32 - o) <= tol_i).all()
    assert (np.abs(g32 - g) <= tol_g).all()
    assert abs(l32 - l) <= 1e-4 * abs(l) + tol_i.sum()


def test_counting_sort_binning_index_arithmetic_over_shapes():
    """The counting-sort binning of the CUDA path (chunks, column prefixes, warp-owned bands of tile rows, lane-owned
    tile classes, groups of 32 candidates, sub-groups of 8 hits) restated sequentially with the kernels' own index
    expressions: equal to the plain stable binning for image shapes from one tile to more than 64 tile rows, row bands,
    few and many chunks, Gaussians taller than the 16 stored tile rows, empty scenes."""
    rng = np.random.default_rng(5)
    shapes = [(16, 16), (40, 23), (50, 40), (200, 150), (130, 300), (64, 1100), (300, 700), (1024, 90), (333, 515)]
    for W, H in shapes:
        for N in (1, 33, 700):
            params, _ = orc.splat_scene(N, W, H, seed=int(rng.integers(1 << 30)))
            params[::9, 2:4] = 3.0                              # some cover the whole image (more than 16 tile rows)
            params[::13, 0] = -4000.0                           # some are off screen
            rec = orc.splat_records(params)
            bands = [(0, H)]
            if H >= 48:
                bands.append((16 * (H // 48), min(H, 16 * (H // 48) + 16 * max(1, H // 40))))
            if H >= 30:
                bands.append((5, H - 3))                        # not tile-aligned
            for rb, re_ in bands:
                nc = N == 33 and (rb, re_) == (0, H)            # one no-cull case per shape: every tile, every Gaussian
                rects, ranges, ids = orc.splat_binning(rec, W, H, rb, re_, no_cull=nc)
                for ctas in (1, 5, 592):
                    cr, ci, info = orc.splat_binning_counting(rec, W, H, rb, re_, no_cull=nc, ctas_total=ctas)
                    assert np.array_equal(cr, ranges) and np.array_equal(ci, ids), (W, H, N, rb, re_, ctas)
                    # the backward work records: ceil(len / 128) set aside per tile (left for the forward pass to fill
                    # in: untouched here), every record beyond them = -1
                    n_kept = sum((e - b + 127) // 128 for b, e in ranges)
                    assert (info[n_kept:] == -1).all() and (info[:n_kept] == -7).all()



def test_sampled_oracle_equals_the_all_pairs_oracle():
    """orc_splat_pixels_f64 / orc_splat_grads_sample_f64 (the checkers for scenes too large for the all-pairs oracle)
    against the all-pairs fp64 oracle on a scene it can run: sampled image rows are bit-identical (same operations in
    the same order); sampled gradients agree to fp64 reassociation (pixel order instead of tile order)."""
    W, H, N = 70, 45, 300
    params, target = orc.splat_scene(N, W, H, seed=11)
    g, o, l, tol = orc.splat_tolerance(params, target, W, H)
    rr = np.random.default_rng(5)
    xy = np.stack([rr.integers(0, W, 200), rr.integers(0, H, 200)], -1).astype(np.int32)
    out, cond = orc.splat_pixels(params, xy, cond=True)
    assert np.array_equal(out, o.reshape(H, W, 3)[xy[:, 1], xy[:, 0]])
    assert (cond >= 0).all() and np.isfinite(cond).all()
    ids = rr.choice(N, 40, replace=False).astype(np.int32)
    gs, tols = orc.splat_grads_sample(params, ids, target, o, W, H)
    assert np.abs(gs - g[ids]).max() <= 1e-12 * max(1.0, np.abs(g).max())
    assert np.allclose(tols, tol[ids], rtol=1e-9, atol=1e-30)


def test_lsq_shipped_example_graph_closed_form():
    """XYZ_FLAG_LSQ_SHIPPED_GRAPH's oracle: r = (a - x1) + b (c - x2)^2 + d - y, dr/d(a, b, c, d) = (1, (c - x2)^2,
    2 b (c - x2), 1) -- the graph parallel_gradient_computation_kernel builds (linear_regression_sgd.cu:103-122)."""
    data = orc.lsq_data(1000, seed=4)
    a, b, c, d = 0.1, 0.9, -0.2, 0.3
    g, l = orc.lsq_grad(data, (a, b, c, d), 2)
    v = c - data[:, 1]
    want = np.array([len(data), (v * v).sum(), (2 * b * v).sum(), len(data)])
    assert np.allclose(g, want, rtol=1e-12)
    assert np.isclose(l, ((a - data[:, 0]) + b * v * v + d - data[:, 2]).sum(), rtol=1e-12)


def test_conic_min_over_rect_closed_form():
    """The backward cull (csrc/splat_kernels.cuh: conic_min_over_rect) keeps a list entry when the minimum of
    q = ia dx^2 + 2 ib dx dy + ic dy^2 over the tile's pixel rectangle is within the bound.  The closed form used there --
    0 if the centre lies inside, else the smallest of the four edge minima with the parabola's vertex clamped to the edge
    -- restated here in numpy: it must be a lower bound of q over the rectangle's pixels (so no pixel within the bound is
    ever dropped) and equal the minimum over a dense sampling of the rectangle."""
    rng = np.random.default_rng(11)

    def closed_form(ia, ib, ic, x0, x1, y0, y1):
        if x0 <= 0 <= x1 and y0 <= 0 <= y1:
            return 0.0
        q = lambda dx, dy: ia * dx * dx + 2 * ib * dx * dy + ic * dy * dy
        kx, ky = -ib / ia, -ib / ic
        return min(q(min(max(kx * y0, x0), x1), y0), q(min(max(kx * y1, x0), x1), y1),
                   q(x0, min(max(ky * x0, y0), y1)), q(x1, min(max(ky * x1, y0), y1)))

    for _ in range(400):
        s0, s1, th = np.exp(rng.uniform(-1, 2.3)), np.exp(rng.uniform(-1, 2.3)), rng.uniform(-np.pi, np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]) @ np.diag([s0, s1])
        inv = np.linalg.inv(R @ R.T)
        ia, ib, ic = inv[0, 0], inv[0, 1], inv[1, 1]
        cx, cy = rng.uniform(-40, 60, 2)
        x0, y0 = -cx, -cy                                   # a 16 x 16 tile at the origin, relative to the centre
        m = closed_form(ia, ib, ic, x0, x0 + 15, y0, y0 + 15)
        px, py = np.meshgrid(x0 + np.arange(16), y0 + np.arange(16))
        q_pixels = ia * px * px + 2 * ib * px * py + ic * py * py
        assert m <= q_pixels.min() * (1 + 1e-12) + 1e-12
        fx, fy = np.meshgrid(np.linspace(x0, x0 + 15, 301), np.linspace(y0, y0 + 15, 301))
        q_dense = ia * fx * fx + 2 * ib * fx * fy + ic * fy * fy
        assert m <= q_dense.min() + 1e-9 and abs(m - q_dense.min()) <= 2e-2 * max(q_dense.min(), 1.0)

