"""dev/sanitize_driver.py -- one modest-size call of every kernel added late in round 1, for compute-sanitizer:
  compute-sanitizer --tool memcheck|racecheck python dev/sanitize_driver.py
(striped accumulation incl. deterministic rows and implicit ids, shared-W chain, multi-epoch SGD, batched::for_each)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x
dev = torch.device("cuda:0")
D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
# striped accumulation: T = 3 (K = 1024), T = 12 (K = 5), T = 2 (K = 1700); partial last unit; deterministic; implicit ids
for k, n in ((1024, (1 << 17) + 77), (5, 1 << 16), (1700, 70_001)):
    for dist in ("uniform", "same"):
        idx, val = orc.accumulate_inputs(n, k, dist, seed=k)
        for flags in (0, x.FLAG_DETERMINISTIC):
            g = torch.zeros(k, device=dev)
            x.accumulate(D(idx), D(val), g, flags)
            ex = orc.accumulate_exact(idx, val, k)
            assert np.allclose(g.cpu().numpy(), ex, rtol=0, atol=1e-4 * np.abs(val).sum()), (k, dist, flags)
    g = torch.zeros(k, device=dev)
    x.accumulate(None, D(val), g)
# fp64 striped flavour (T = 3, 12, 2), deterministic rows, implicit ids; and a shard misaligned by one element
for k, n in ((1024, (1 << 16) + 300), (7, 1 << 16), (1700, 70_001)):
    idx, val = orc.accumulate_inputs(n, k, "zipf", seed=k + 1)
    v64 = val.astype(np.float64)
    for flags in (0, x.FLAG_DETERMINISTIC):
        g = torch.zeros(k, dtype=torch.float64, device=dev)
        x.accumulate(D(idx), D(v64), g, flags)
        ex = np.zeros(k); np.add.at(ex, idx, v64)
        assert np.allclose(g.cpu().numpy(), ex, rtol=0, atol=1e-9 * np.abs(v64).sum()), (k, flags)
    g = torch.zeros(k, dtype=torch.float64, device=dev)
    x.accumulate(None, D(v64), g)
idx, val = orc.accumulate_inputs((1 << 17) + 9, 1024, "uniform", seed=77)
ti = torch.zeros(idx.size + 1, dtype=torch.int32, device=dev); ti[1:] = D(idx)
tv = torch.zeros(val.size + 1, dtype=torch.float32, device=dev); tv[1:] = D(val)
g = torch.zeros(1024, device=dev)
x.accumulate(ti[1:], tv[1:], g)
assert np.allclose(g.cpu().numpy(), orc.accumulate_exact(idx, val, 1024), rtol=0, atol=1e-4 * np.abs(val).sum())
# shared-W chain: TMA path + tail, and the plain path
J, W, S, gg = orc.covproj_inputs(5000 + 13, seed=1)
for sl in (slice(None), slice(1, 900)):
    n = J[sl].shape[0]
    o, gJ, gS = [torch.empty((n, w), device=dev) for w in (3, 6, 6)]
    gW = torch.zeros(9, device=dev)
    x.covproj_shared_w_fwd_bwd(D(J[sl]), D(W[0]), D(S[sl]), D(gg[sl]), o, gJ, gW, gS)
# multi-epoch SGD (cooperative launch)
data = D(orc.lsq_data(20_000, seed=2))
p = torch.zeros(8, dtype=torch.float64, device=dev); p[1] = 1.0
x.lsq_sgd_run(data, p, 4096, 7, 0, [1e-4] * 9)
# batched::for_each through the probe library
so = os.path.join(ROOT, "tests", "csrc", "_build", "libxyz_batched.so")
B = ctypes.CDLL(so)
B.batched_chain.restype = ctypes.c_float
B.batched_chain.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
B.batched_lsq.restype = ctypes.c_float
B.batched_lsq.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
n = 40_000 + 5
tin = torch.rand((n, 15), device=dev); tout = torch.empty((n, 15), device=dev)
w18 = torch.rand(18, device=dev); gw = torch.zeros(18, device=dev)
assert B.batched_chain(tin.data_ptr(), tout.data_ptr(), n, w18.data_ptr(), gw.data_ptr(), 1) >= 0
pts = torch.rand((n, 3), dtype=torch.float64, device=dev)
v5 = torch.tensor([0.0, 1.0, 0.0, 0.0, 0.0], dtype=torch.float64, device=dev); g5 = torch.zeros(5, dtype=torch.float64, device=dev)
assert B.batched_lsq(pts.data_ptr(), n, v5.data_ptr(), g5.data_ptr(), 1) >= 0
torch.cuda.synchronize()
print("sanitize_driver ok")
# splat: counting-sort binning (default) and the radix path, atomics and deterministic rows; Gaussians taller than the
# 16 cached tile rows, a batch with more spans than register slots, a row band, > 4096 tiles (opt-in shared memory)
if os.environ.get("SANITIZE_SPLAT", "1") != "0":
    for (Ws, Hs, Ns, band) in ((320, 400, 1500, None), (320, 400, 1500, (100, 300)), (2048, 1040, 600, None)):
        params, target = orc.splat_scene(Ns, Ws, Hs, seed=3)
        params[10:60, 2:4] = 3.5                      # 50 consecutive huge Gaussians: > 512 spans in one batch
        params[100:140, 0] = -5000.0                  # off screen
        tp, tt = D(params), D(target)
        for flags in (0, x.FLAG_DETERMINISTIC, x.FLAG_RADIX_BINNING, x.FLAG_PRECISE_MATH):
            grads = torch.zeros((Ns, 9), device=dev); img = torch.zeros((Ws * Hs, 3), device=dev)
            loss = torch.zeros(1, device=dev)
            x.launch_gaussian_splatting(tp, grads, tt, img, loss, Ws, Hs, Ns, flags, rows=band)
            torch.cuda.synchronize()
            assert np.isfinite(grads.cpu().numpy()).all() and np.isfinite(loss.item())
print("sanitize_driver: ok")
# round 2: workspace launches (band-local histograms, deterministic offsets without the library scan, every forward CTA
# shape is reached by the image sizes above / below), two workspaces in flight on two streams, empty band, N = 0, an
# overflowing launch, more than 8192 tiles through the opt-in shared memory, the fused peer-memory optimiser step
if os.environ.get("SANITIZE_R02", "1") != "0":
    for (Ws, Hs, Ns, band) in ((320, 400, 1500, None), (320, 400, 1500, (96, 304)), (2080, 1040, 800, None)):
        params, target = orc.splat_scene(Ns, Ws, Hs, seed=4)
        params[10:40, 2:4] = 3.0
        tp, tt = D(params), D(target)
        for flags in (0, x.FLAG_DETERMINISTIC, x.FLAG_PRECISE_MATH | x.FLAG_DETERMINISTIC):
            ws = x.SplatWorkspace(Ws, Hs, Ns, 400 * Ns, flags, rows=band)
            grads = torch.zeros((Ns, 9), device=dev); img = torch.zeros((Ws * Hs, 3), device=dev); loss = torch.zeros(1, device=dev)
            ws.launch(tp, grads, tt, img, loss)
            assert not ws.status()["overflowed"]
            assert np.isfinite(grads.cpu().numpy()).all() and np.isfinite(loss.item())
    Ws, Hs, Ns = 320, 400, 1500
    params, target = orc.splat_scene(Ns, Ws, Hs, seed=4)
    tp, tt = D(params), D(target)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    keep = []
    for st in (s1, s2):
        ws = x.SplatWorkspace(Ws, Hs, Ns, 300 * Ns, x.FLAG_DETERMINISTIC)
        grads = torch.zeros((Ns, 9), device=dev); img = torch.zeros((Ws * Hs, 3), device=dev); loss = torch.zeros(1, device=dev)
        st.wait_stream(torch.cuda.current_stream())   # the zero fills above ran on the current stream
        with torch.cuda.stream(st):
            ws.launch(tp, grads, tt, img, loss, stream=st)
        keep.append((ws, grads, img, loss))
    torch.cuda.synchronize()
    assert torch.equal(keep[0][1], keep[1][1]) and torch.equal(keep[0][2], keep[1][2])
    for rows, n in (((64, 64), Ns), (None, 0)):
        ws = x.SplatWorkspace(Ws, Hs, n, 1000, 0, rows=rows)
        grads = torch.zeros((max(n, 1), 9), device=dev); img = torch.zeros((Ws * Hs, 3), device=dev); loss = torch.zeros(1, device=dev)
        ws.launch(tp[:n] if n else torch.empty((0, 9), device=dev), grads[:n] if n else torch.empty((0, 9), device=dev), tt, img, loss)
        assert ws.status()["entries"] == 0 and (grads == 0).all()
    ws = x.SplatWorkspace(Ws, Hs, Ns, 500, 0)        # far too small: empty lists, memory-safe
    grads = torch.zeros((Ns, 9), device=dev); img = torch.ones((Ws * Hs, 3), device=dev); loss = torch.zeros(1, device=dev)
    ws.launch(tp, grads, tt, img, loss)
    assert ws.status()["overflowed"] and (img == 0).all() and (grads == 0).all()
    grp = x.PeerGroup(0, 1, lambda h: [h])
    for n in (3, 4097, 50_001):
        ps = x.PeerSplat(grp, n, lambda h: [h])
        ps.params.normal_(); ps.grads.normal_()
        l = torch.ones(1, device=dev)
        for it in (1, 0, 0):
            ps.adam_step(0.1, 0.01, 0.001, 0.02, 0.05, iteration=it, total_loss=l)
        torch.cuda.synchronize()
        assert (ps.grads == 0).all() and np.isfinite(ps.params.cpu().numpy()).all() and l.item() == 1.0
        ps.close()
    grp.close()
    print("sanitize_driver: round-2 paths ok")
