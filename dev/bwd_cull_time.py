"""dev/bwd_cull_time.py -- stage times of the C4 scene (full image, 128-row band) with the backward cull and with
XYZ_FLAG_BWD_ALL_PAIRS, and the size of what the cull drops (deterministic mode: the difference IS the dropped terms)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
W = H = 1024; N = 100_000
params, target = orc.splat_c4_scene(N, W, H, 42)
tp, tt = torch.from_numpy(params).to(dev), torch.from_numpy(target).to(dev)
g = torch.zeros((N, 9), device=dev); o = torch.zeros((W * H, 3), device=dev); l = torch.zeros(1, device=dev)
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for name, fl in (("cull", 0), ("all_pairs", x.FLAG_BWD_ALL_PAIRS)):
    for rows in ((0, 1024), (448, 576)):
        acc = {}
        for i in range(13):
            x.launch_gaussian_splatting(tp, g, tt, o, l, W, H, N, x.FLAG_TIMING | fl, rows=rows)
            t = x.splat_last_timing()
            if i >= 3:
                for k, v in t.items(): acc[k] = acc.get(k, 0.0) + v / 10
        print(f"BWDCULL {tag} {name:9s} rows={rows[1] - rows[0]:5d} " + " ".join(f"{k[:-3]}={v:7.1f}" for k, v in acc.items()), flush=True)
res = {}
for name, fl in (("cull", 0), ("all_pairs", x.FLAG_BWD_ALL_PAIRS)):
    g.zero_(); l.zero_()
    x.launch_gaussian_splatting(tp, g, tt, o, l, W, H, N, x.FLAG_DETERMINISTIC | fl)
    torch.cuda.synchronize()
    res[name] = g.cpu().numpy().astype(np.float64)
d = np.abs(res["cull"] - res["all_pairs"])
scale = np.abs(res["all_pairs"]).max(axis=0)
print(f"BWDCULL {tag} dropped terms: max |diff| per component / max |grad| of that component =",
      " ".join(f"{v:.2e}" for v in d.max(axis=0) / scale), flush=True)
