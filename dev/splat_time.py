"""dev/splat_time.py -- C4 splat iteration time (zero_grad + loss reset + launch), median of 9."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
W = H = 1024; N = 100_000
params, target = orc.splat_c4_scene(N, W, H, 42)
tp, tt = torch.from_numpy(params).to(dev), torch.from_numpy(target).to(dev)
grads = torch.zeros((N, 9), device=dev); img = torch.zeros((W * H, 3), device=dev); loss = torch.zeros(1, device=dev)
def it(fl=0):
    x.zero_gradients(grads); loss.zero_(); x.launch_gaussian_splatting(tp, grads, tt, img, loss, W, H, N, fl)
def med(fl):
    ts = []
    for i in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); it(fl); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort()
    return ts
for fl, nm in ((x.FLAG_RADIX_BINNING, "radix binning"), (x.FLAG_DETERMINISTIC, "deterministic"),
               (x.FLAG_DETERMINISTIC | x.FLAG_RADIX_BINNING, "deterministic, radix binning")):
    ts = med(fl)
    print(f"   {nm}: median {ts[len(ts)//2]:.4f} ms, min {ts[0]:.4f} ms")
ts = med(0)
import hashlib
h = hashlib.sha1(img.cpu().numpy().tobytes()).hexdigest()[:16]
gh = float(grads.abs().sum().item())
print(f"{sys.argv[1] if len(sys.argv) > 1 else ''} C4 splat: median {ts[len(ts)//2]:.4f} ms, min {ts[0]:.4f} ms, loss {loss.item():.9g}, image sha1 {h}, sum|grads| {gh:.9g}")
for fl, nm in ((x.FLAG_PRECISE_MATH, "precise"),):
    x.zero_gradients(grads); loss.zero_(); x.launch_gaussian_splatting(tp, grads, tt, img, loss, W, H, N, fl)
    torch.cuda.synchronize()
    print(f"   {nm}: loss {loss.item():.9g}, image sha1 {hashlib.sha1(img.cpu().numpy().tobytes()).hexdigest()[:16]}")
