"""dev/bwd_cull_sweep.py -- the backward cull's bound D (XYZ_SPLAT_BWD_D2, read once per process: one subprocess per value):
backward time and pairs at C4, and what the cull drops -- the deterministic-mode gradients with the bound minus those
with XYZ_FLAG_BWD_ALL_PAIRS (every kept entry is computed by the same instructions, so the difference IS the dropped
terms), relative to the fp64 sum of |terms| of 66 sampled Gaussians (oracle, every pixel)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import numpy as np, torch
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    import xyz_autodiff_cuda_b200 as x
    dev = torch.device("cuda:0")
    W = H = 1024; N = 100_000
    params, target = orc.splat_c4_scene(N, W, H, 42)
    tp, tt = torch.from_numpy(params).to(dev), torch.from_numpy(target).to(dev)
    g = torch.zeros((N, 9), device=dev); o = torch.zeros((W * H, 3), device=dev); l = torch.zeros(1, device=dev)
    acc = {}
    for i in range(13):
        x.launch_gaussian_splatting(tp, g, tt, o, l, W, H, N, x.FLAG_TIMING)
        t = x.splat_last_timing()
        if i >= 3:
            for k, v in t.items(): acc[k] = acc.get(k, 0.0) + v / 10
    pairs = x.splat_last_backward_stats()["pairs"]
    res = {}
    for name, fl in (("cull", 0), ("all", x.FLAG_BWD_ALL_PAIRS)):
        g.zero_(); l.zero_()
        x.launch_gaussian_splatting(tp, g, tt, o, l, W, H, N, x.FLAG_DETERMINISTIC | fl)
        torch.cuda.synchronize()
        res[name] = g.cpu().numpy().astype(np.float64)
    rng = np.random.default_rng(7)
    ids = np.unique(np.concatenate([rng.choice(N, 64, replace=False), [0, N - 1]])).astype(np.int32)
    _, tol = orc.splat_grads_sample(params, ids, target, o.cpu().numpy(), W, H)   # tol >= 1e-4 * sum|terms|
    rel = np.abs(res["cull"][ids] - res["all"][ids]) / (tol / 1e-4)
    print(f"BWDSWEEP D={os.environ.get('XYZ_SPLAT_BWD_D2', 'default'):>7s} backward={acc['backward_us']:7.1f} us "
          f"forward={acc['forward_loss_us']:7.1f} us pairs={pairs:.4g} dropped/sum|terms|: max {rel.max():.2e} "
          f"per component {' '.join(f'{v:.1e}' for v in rel.max(axis=0))}", flush=True)
else:
    for d in ("176", "96", "80", "64", "56", "48", "40", "32"):
        subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=dict(os.environ, XYZ_SPLAT_BWD_D2=d))
