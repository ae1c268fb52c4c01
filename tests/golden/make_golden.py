"""tests/golden/make_golden.py -- regenerates tests/golden/reference_vectors.npz from the REFERENCE itself.

Run where /root/reference exists (after `make -C oracle`): every array below is produced by
oracle/_ref/libxyz_ref.so, i.e. by the reference's own headers and its own splat kernel body compiled for
the host (oracle/ref_driver.cpp), single-threaded, -O2 -ffp-contract=off.  The fixture is what pins the
oracle port (tests/test_oracle.py) on machines where the reference tree is absent.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as orc  # noqa: E402
from test_api_headers import op_cases  # noqa: E402


def main():
    assert orc.have_ref(), "build oracle/_ref first (make -C oracle)"
    out = {}
    # splat: two small scenes, fp32, the reference kernel body
    for tag, (W, H, N, seed) in {"a": (37, 21, 7, 0), "b": (64, 48, 50, 3)}.items():
        params, target = orc.splat_scene(N, W, H, seed=seed)
        g, o, l, _ = orc.splat(params, target, W, H, np.float32, which="ref", threads=1)
        out[f"splat_{tag}_shape"] = np.array([W, H, N, seed])
        out[f"splat_{tag}_params"], out[f"splat_{tag}_target"] = params, target
        out[f"splat_{tag}_grads"], out[f"splat_{tag}_output"], out[f"splat_{tag}_loss"] = g, o, np.float32(l)
    # config-4 distribution (initialize_random + create_test_image), reduced
    params, target = orc.splat_c4_scene(64, 48, 32, seed=42)
    g, o, l, _ = orc.splat(params, target, 48, 32, np.float32, which="ref", threads=1)
    out["splat_c4_params"], out["splat_c4_target"] = params, target
    out["splat_c4_grads"], out["splat_c4_output"], out["splat_c4_loss"] = g, o, np.float32(l)
    # least squares
    data = orc.lsq_data(1000, seed=42)
    for ro in (0, 1):
        g, l = orc.lsq_grad(data, (0.0, 1.0, 0.0, 0.0), bool(ro), which="ref", threads=1)
        out[f"lsq_grad_{ro}"], out[f"lsq_loss_{ro}"] = g, np.float64(l)
    out["lsq_data"] = data
    # accumulation
    idx, val = orc.accumulate_inputs(5000, 64, "zipf", seed=1)
    out["acc_idx"], out["acc_val"] = idx, val
    out["acc_grad_f32"] = orc.accumulate(idx, val, 64, which="ref")
    out["acc_grad_f64"] = orc.accumulate(idx, val.astype(np.float64), 64, which="ref")
    # covariance projection (tree of reference op::matmul nodes)
    J, W9, S, g = orc.covproj_inputs(200, seed=7)
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        outs = orc.covproj(J, W9, S, g, dt, which="ref", threads=1)
        for name, a in zip(("out", "gJ", "gW", "gS"), outs):
            out[f"cov_{tag}_{name}"] = a
    out["cov_J"], out["cov_W"], out["cov_S"], out["cov_g"] = J, W9, S, g
    # the op table, fp64: value + both input adjoints, flattened
    rng = np.random.default_rng(123)
    rows = []
    for op, in1, in2, cst, gout, aux in op_cases(rng):
        o, g1, g2 = orc.eval_op("ref", op, in1, in2, cst, gout, aux, np.float64)
        rows.append(np.concatenate([o, g1, g2]))
    out["op_table_f64"] = np.concatenate(rows)
    out["op_table_rows"] = np.array([r.size for r in rows])
    # known-answer graphs
    for name in ("dag", "shared_subgraph", "broadcast"):
        out[f"kat_{name}"] = orc.kat("ref", name)
    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
