"""dev/splat_async_time.py -- C4 iteration: ordinary launch, XYZ_FLAG_ASYNC, and a CUDA graph of the whole iteration."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
W = H = 1024; N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
params, target = orc.splat_c4_scene(N, W, H, 42)
tp, tt = torch.from_numpy(params).to(dev), torch.from_numpy(target).to(dev)
grads = torch.zeros((N, 9), device=dev); img = torch.zeros((W * H, 3), device=dev); loss = torch.zeros(1, device=dev)
def it(fl):
    x.zero_gradients(grads); loss.zero_(); x.launch_gaussian_splatting(tp, grads, tt, img, loss, W, H, N, fl)
def med(f, n=12):
    ts = []
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
print(f"ordinary       : {med(lambda: it(0)):.4f} ms  loss {loss.item():.9g}")
print(f"XYZ_FLAG_ASYNC : {med(lambda: it(x.FLAG_ASYNC)):.4f} ms  loss {loss.item():.9g}")
def ten(): [it(x.FLAG_ASYNC) for _ in range(10)]
print(f"XYZ_FLAG_ASYNC, 10 iterations back to back: {med(ten) / 10:.4f} ms/iter")
def ten0(): [it(0) for _ in range(10)]
print(f"ordinary, 10 iterations back to back      : {med(ten0) / 10:.4f} ms/iter")
try:
    it(0); it(x.FLAG_ASYNC); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        it(x.FLAG_ASYNC); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            it(x.FLAG_ASYNC)
    torch.cuda.synchronize()
    print(f"CUDA graph of one iteration: {med(g.replay):.4f} ms  loss {loss.item():.9g}")
    def ten_g(): [g.replay() for _ in range(10)]
    print(f"CUDA graph, 10 replays back to back: {med(ten_g) / 10:.4f} ms/iter")
except Exception as e:  # noqa: BLE001
    print("graph capture failed:", repr(e)[:300])
