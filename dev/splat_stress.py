"""dev/splat_stress.py -- seeded sweep of small random splat scenes (odd image sizes, large and tiny Gaussians, all flag
combinations) against the fp64 oracle with the tolerances of tests/test_gpu_parity.py.  Prints the failures."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rng = np.random.default_rng(777)
bad = 0
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for case in range(cases):
    W, H = int(rng.integers(1, 130)), int(rng.integers(1, 100))
    N = int(rng.choice([1, 2, 17, 64, 200, 400]))
    params, target = orc.splat_scene(N, W, H, seed=1000 + case, small=bool(case % 2))
    if case % 5 == 0:
        params[:, 2:4] = rng.uniform(-1.0, 2.3, (N, 2)).astype(np.float32)   # from sub-pixel to image-sized
    if case % 7 == 0:
        params[: max(1, N // 8), 0:2] += 500.0                               # some Gaussians far outside the image
    rg, ro, rl, tol = orc.splat_tolerance(params, target, W, H)
    for flags in (0, x.FLAG_PRECISE_MATH, x.FLAG_DETERMINISTIC, x.FLAG_NO_CULL, x.FLAG_PRECISE_MATH | x.FLAG_DETERMINISTIC):
        grads = torch.zeros((N, 9), device=dev); out = torch.full((W * H, 3), float("nan"), device=dev); loss = torch.zeros(1, device=dev)
        x.launch_gaussian_splatting(D(params), grads, D(target), out, loss, W, H, N, flags)
        torch.cuda.synchronize()
        g, o, l = grads.cpu().numpy(), out.cpu().numpy(), loss.item()
        ok_img = (np.abs(o - ro) <= 1e-5 * np.maximum(np.abs(ro), np.abs(ro).max() * 1e-3)).all()
        ok_loss = abs(l - rl) <= 1e-4 * abs(rl) + 1e-30
        ok_g = (np.abs(g - rg) <= tol).all()
        if not (ok_img and ok_loss and ok_g):
            bad += 1
            if flags == x.FLAG_PRECISE_MATH and orc.have_ref():
                # how far is the REFERENCE's own fp32 evaluation (its kernel body compiled for the host) from fp64 here?
                g32, o32, l32, _ = orc.splat(params, target, W, H, np.float32, which="ref")
                img_scale = np.maximum(np.abs(ro), np.abs(ro).max() * 1e-3)
                print(f"     reference fp32 vs fp64: max|dg|/tol {np.max(np.abs(g32 - rg) / (tol + 1e-30)):.3g}, "
                      f"image rel {np.max(np.abs(o32 - ro) / img_scale):.3g};  ours vs fp64: image rel {np.max(np.abs(o - ro) / img_scale):.3g};"
                      f"  ours vs reference fp32: max|dg|/tol {np.max(np.abs(g - g32) / (tol + 1e-30)):.3g}, image rel {np.max(np.abs(o - o32) / img_scale):.3g}")
            print(f"FAIL case {case} W={W} H={H} N={N} flags={flags}: image {ok_img} loss {ok_loss} grads {ok_g} "
                  f"max|dg|/tol {np.max(np.abs(g - rg) / (tol + 1e-30)):.3g}")
print(f"splat_stress: {cases} scenes x 5 flag sets, {bad} failures")
