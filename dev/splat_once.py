"""dev/splat_once.py -- three C4 splat iterations (for an ncu launch list: dev/launch_breakdown.py reads the last)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
W = H = 1024; N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
fl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
params, target = orc.splat_c4_scene(N, W, H, 42)
tp, tt = torch.from_numpy(params).to(dev), torch.from_numpy(target).to(dev)
grads = torch.zeros((N, 9), device=dev); img = torch.zeros((W * H, 3), device=dev); loss = torch.zeros(1, device=dev)
for _ in range(3):
    x.zero_gradients(grads); loss.zero_(); x.launch_gaussian_splatting(tp, grads, tt, img, loss, W, H, N, fl)
torch.cuda.synchronize()
