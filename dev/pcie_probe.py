"""dev/pcie_probe.py -- pinned H2D / D2H bandwidth alone and concurrently (bounds the e2e figure of bench.py)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 28  # 1 GiB of float32
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device=dev)
d_b = torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3
gb = n * 4 / 1e9
def h2d(): d_a.copy_(h_in, non_blocking=True)
def d2h(): h_out.copy_(d_b, non_blocking=True)
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both_chunked(chunk=1 << 24):
    for o in range(0, n, chunk):
        with torch.cuda.stream(s1): d_a[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
        with torch.cuda.stream(s2): h_out[o:o + chunk].copy_(d_b[o:o + chunk], non_blocking=True)
print(f"H2D alone {gb / t(h2d):.1f} GB/s, D2H alone {gb / t(d2h):.1f} GB/s")
tb = t(both)
print(f"concurrent: {gb / tb:.1f} GB/s each direction ({2 * gb / tb:.1f} GB/s total)")
tb = t(both_chunked)
print(f"concurrent, 64 MiB chunks: {gb / tb:.1f} GB/s each direction")
