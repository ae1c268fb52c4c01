"""dev/fwd_threads_sweep.py -- forward / backward stage times of a row band of the C4 scene for the forward kernel's
threads-per-tile configurations (XYZ_SPLAT_FWD_THREADS is read once per process: run once per setting)."""
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
for rows in ((0, 1024), (0, 512), (256, 512), (384, 512), (448, 512)):
    acc = {}
    for i in range(13):
        x.launch_gaussian_splatting(tp, g, tt, o, l, W, H, N, x.FLAG_TIMING, rows=rows)
        t = x.splat_last_timing()
        if i >= 3:
            for k, v in t.items(): acc[k] = acc.get(k, 0.0) + v / 10
    print(f"FWDSWEEP threads={os.environ.get('XYZ_SPLAT_FWD_THREADS', 'auto')} rows={rows[1] - rows[0]:5d} "
          + " ".join(f"{k[:-3]}={v:7.1f}" for k, v in acc.items()), flush=True)
