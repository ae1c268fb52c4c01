"""dev/zero_copy_probe.py -- covproj with the kernel reading / writing PINNED HOST memory directly (TMA over PCIe) against
the staged H2D / kernel / D2H pipeline of host_api.CovprojHostPipeline."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xyz_autodiff_cuda_b200 as x
from importlib import import_module
host_api = import_module("xyz_autodiff_cuda_b200.host_api")
dev = torch.device("cuda:0")
n = 1 << 24
h_in = [torch.empty((n, w), dtype=torch.float32).uniform_(-1, 1).pin_memory() for w in (6, 9, 6, 3)]
h_out = [torch.zeros((n, w), dtype=torch.float32).pin_memory() for w in (3, 6, 9, 6)]
h_out2 = [torch.zeros((n, w), dtype=torch.float32).pin_memory() for w in (3, 6, 9, 6)]
L = x.lib()
st = torch.cuda.current_stream()
def zero_copy():
    rc = L.xyz_covproj_fwd_bwd_f32(*[t.data_ptr() for t in h_in], *[t.data_ptr() for t in h_out], n, st.cuda_stream, 0)
    assert rc == 0, rc
pipe = host_api.CovprojHostPipeline(dev)
def staged():
    pipe.run(h_in, h_out2)
def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
tz = timed(zero_copy); ts = timed(staged)
print(f"zero-copy kernel: {tz:.2f} ms = {n / tz / 1e3:.3e} evals/s, {96 * n / tz / 1e6:.1f} GB/s each way")
print(f"staged pipeline : {ts:.2f} ms = {n / ts / 1e3:.3e} evals/s, {96 * n / ts / 1e6:.1f} GB/s each way")
print("identical results:", all(torch.equal(a, b) for a, b in zip(h_out, h_out2)))
