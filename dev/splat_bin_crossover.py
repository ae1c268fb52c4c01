"""dev/splat_bin_crossover.py -- iteration time of the two binning paths over N (1024x1024, the C4 distribution)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
W = H = 1024
for N in [int(a) for a in sys.argv[1:]] or [25_000, 100_000, 200_000, 400_000, 800_000, 1_600_000]:
    params, target = orc.splat_c4_scene(N, W, H, 42)
    tp, tt = torch.from_numpy(params).to(dev), torch.from_numpy(target).to(dev)
    grads = torch.zeros((N, 9), device=dev); img = torch.zeros((W * H, 3), device=dev); loss = torch.zeros(1, device=dev)
    res = {}
    for fl, nm in ((x.FLAG_RADIX_BINNING, "radix"), (0, "default")):
        ts = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); x.zero_gradients(grads); loss.zero_()
            x.launch_gaussian_splatting(tp, grads, tt, img, loss, W, H, N, fl); b.record(); torch.cuda.synchronize()
            if i >= 3: ts.append(a.elapsed_time(b))
        ts.sort(); res[nm] = ts[len(ts) // 2]
    e = x.splat_last_stats()["entries"]
    print(f"N {N:8d} entries {e:10d}: radix {res['radix']:9.4f} ms, default {res['default']:9.4f} ms, diff {1e3 * (res['default'] - res['radix']):+8.1f} us")
