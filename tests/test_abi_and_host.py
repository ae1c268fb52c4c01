"""CPU-side checks (`-m "not gpu"`): the C-ABI library loads and exports exactly what include/xyz_b200.h declares,
the header is valid C with the reference's struct layouts, the product has no CPU fallback, and the multi-GPU host
logic (sharding + shared-gradient all-reduce) works at world_size 2 over gloo."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

import oracle_lib as orc
import xyz_autodiff_cuda_b200 as x

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "xyz_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"XYZ_API\s+[\w\s\*]+?\b(xyz_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = x.lib()  # raises if the library was not built
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/xyz_b200.h but not exported by libxyz_b200.so"
    assert sorted(x.EXPORTS) == names
    out = subprocess.run(["nm", "-D", "--defined-only", x.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    exported = sorted(set(re.findall(r" T (xyz_\w+)", out)))
    assert exported == names, "the library exports symbols the header does not declare (or vice versa)"
    assert b"sm_100a" in L.xyz_b200_version()


def test_header_is_plain_c_with_reference_layouts():
    src = r'''
#include "xyz_b200.h"
_Static_assert(sizeof(xyz_gaussian_params) == 36, "GaussianParams is 9 floats");
_Static_assert(sizeof(xyz_gaussian_grads) == 36, "GaussianGrads is 9 floats");
_Static_assert(sizeof(xyz_adam_state) == 72, "AdamState is 18 floats");
_Static_assert(sizeof(xyz_data_point) == 24, "DataPoint is 3 doubles");
_Static_assert(sizeof(xyz_lsq_parameters) == 64, "Parameters is value[4] + grad[4]");
int main(void) { return XYZ_FLAG_DETERMINISTIC + XYZ_FLAG_PRECISE_MATH == 3 ? 0 : 1; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o",
                        os.path.join(d, "t")], check=True)
        assert subprocess.run([os.path.join(d, "t")]).returncode == 0


def test_no_cpu_fallback_and_argument_checks():
    with pytest.raises(RuntimeError):
        x.lsq_grad(torch.zeros((4, 3), dtype=torch.float64), torch.zeros(8, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        x.accumulate(torch.zeros(4, dtype=torch.int32), torch.zeros(4), torch.zeros(2))
    with pytest.raises(RuntimeError):
        x.launch_gaussian_splatting(torch.zeros((1, 9)), torch.zeros((1, 9)), torch.zeros((256, 3)), torch.zeros((256, 3)),
                                    torch.zeros(1), 16, 16, 1)
    # the product package never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "xyz-autodiff-cuda_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle_lib" not in text and "libxyz_oracle" not in text and "libxyz_ref" not in text, f


def test_sharding_helpers():
    from importlib import import_module
    par = import_module("xyz_autodiff_cuda_b200.parallel")
    for n in (0, 1, 7, 1000, (1 << 26) + 3):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
    for h in (16, 100, 1024, 1030):
        for world in (1, 2, 4, 8):
            bands = par.row_bands(h, world)
            assert bands[0][0] == 0 and bands[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            assert all(b % 16 == 0 for b, _ in bands)
    assert par.views_for_rank(8, 1, 4) == [1, 5] and sum(len(par.views_for_rank(8, r, 3)) for r in range(3)) == 8


def test_balanced_row_bands_partition_is_valid_and_minimises_the_largest_band():
    """parallel.balanced_row_bands against a dynamic-programming optimum of the same contiguous partition problem."""
    import random
    from importlib import import_module
    par = import_module("xyz_autodiff_cuda_b200.parallel")

    def optimum(cost, world):
        n, pre = len(cost), [0.0]
        for c in cost:
            pre.append(pre[-1] + c)
        dp = [[float("inf")] * (n + 1) for _ in range(world + 1)]
        dp[0][0] = 0.0
        for k in range(1, world + 1):
            for i in range(n + 1):
                dp[k][i] = min([dp[k - 1][i]] + [max(dp[k - 1][j], pre[i] - pre[j]) for j in range(i)])
        return dp[world][n]

    rng = random.Random(3)
    for trial in range(400):
        n, world = rng.randint(1, 20), rng.choice([1, 2, 3, 4, 8])
        h = n * 16 - rng.randint(0, 15)
        kind = trial % 4
        cost = ([rng.randint(0, 5) for _ in range(n)], [rng.random() * 100 for _ in range(n)],
                [rng.choice([0, 0, 1000, 3]) for _ in range(n)], [0] * n)[kind]
        b = par.balanced_row_bands(cost, h, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == h
        assert all(p[1] == q[0] for p, q in zip(b, b[1:])) and all(p[0] <= p[1] for p in b)
        assert all(p[0] % 16 == 0 or p[0] == h for p in b)           # tile aligned (empty bands sit at the end)
        largest = max(sum(cost[p[0] // 16:(p[1] + 15) // 16]) for p in b)
        assert largest <= optimum(cost, world) * (1 + 1e-6) + 1e-9, (cost, world, b)
    with pytest.raises(ValueError):
        par.balanced_row_bands([1, 2, 3], 1024, 2)                    # one cost per tile row expected


def test_splat_workspace_size_query_is_host_arithmetic():
    """xyz_splat_workspace_bytes needs no GPU: sizes grow with the list bound, the deterministic flag adds its rows per
    entry, a row band needs less than the whole image, and what the workspace launch cannot do asks for 0 bytes."""
    L = x.lib()
    f = L.xyz_splat_workspace_bytes
    full = f(1024, 1024, 100_000, 0, 1024, 4_200_000, 0)
    assert full > 0 and full % 256 == 0
    more = f(1024, 1024, 100_000, 0, 1024, 8_400_000, 0)
    per_entry = (more - full) / 4_200_000
    assert 8.0 <= per_entry <= 8.5                                    # 4 B list + 4 B backward item (+ work records)
    det = f(1024, 1024, 100_000, 0, 1024, 4_200_000, x.FLAG_DETERMINISTIC)
    assert 39.9 <= (det - full) / 4_200_000 <= 40.3                   # + 4 B ids + 36 B gradient row per entry
    band = f(1024, 1024, 100_000, 0, 128, 4_200_000, 0)
    assert 0 < band < full
    assert f(1024, 1024, 100_000, 512, 512, 100, 0) > 0                # an empty band is a valid (empty) launch
    assert f(1024, 1024, 0, 0, 1024, 0, 0) > 0
    assert f(1024, 1024, 100_000, 0, 1024, 4_200_000, x.FLAG_RADIX_BINNING) == 0
    assert f(16384, 16384, 10, 0, 16384, 100, 0) == 0                 # more than 57 344 tiles in the band
    assert f(0, 1024, 10, 0, 1024, 10, 0) == 0 and f(1024, 1024, 10, 0, 1025, 10, 0) == 0
    assert f(1024, 1024, 10, 8, 4, 10, 0) == 0 and f(1024, 1024, 10, 0, 1024, -1, 0) == 0
    assert f(1024, 1024, 10, 0, 1024, 1 << 31, 0) == 0


def test_entry_points_reject_bad_arguments_before_any_cuda_call():
    """Argument errors are reported as XYZ_ERR_INVALID_ARGUMENT (-1) / XYZ_ERR_NOT_INITIALISED (-3) by host code alone."""
    L = x.lib()
    null = ctypes.c_void_p(0)
    four = (ctypes.c_longlong * 4)()
    assert L.xyz_launch_gaussian_splatting_ws(null, null, null, null, null, 16, 16, 0, 0, 16, null, 0, 0, null, 0) == -1
    assert L.xyz_splat_workspace_init(null, 4096, null) == -1
    assert L.xyz_splat_workspace_status(null, null, four) == -1
    assert L.xyz_splat_last_stats(None) == -3
    assert L.xyz_comm_rank(null) == -1 and L.xyz_comm_world(null) == -1
    assert L.xyz_comm_destroy(null) == 0
    assert L.xyz_allreduce_grads(null, null, 9, null) == -1
    assert L.xyz_allreduce_f64(null, null, 1, null) == -1
    lr = (ctypes.c_float * 5)(1e-3, 1e-3, 1e-3, 1e-3, 1e-3)
    assert L.xyz_adam_step_individual_sharded(null, null, null, null, 10, lr, 0.9, 0.999, 1e-8, 1, null, null) == -1
    assert L.xyz_comm_init_rank(null, None, 0, 1) == -1
    assert L.xyz_comm_init_all(null, 0, None) == -1
    assert L.xyz_comm_unique_id(None) == -1


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import oracle_lib as orc
from importlib import import_module
import xyz_autodiff_cuda_b200
par = import_module("xyz_autodiff_cuda_b200.parallel")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# C1: each rank owns an element range; only the 4 shared-parameter gradients (+ loss) are exchanged
data = orc.lsq_data(10_001, seed=5)
b, e = par.shard_range(data.shape[0], rank, world)
g, l = orc.lsq_grad(data[b:e], (0.3, 1.2, -0.4, 0.1))      # stand-in for the CUDA kernel on a CPU-only box
tg, tl = torch.from_numpy(g.copy()), torch.tensor([l], dtype=torch.float64)
par.allreduce_shared_grads(tg, tl)
full_g, full_l = orc.lsq_grad(data, (0.3, 1.2, -0.4, 0.1))
assert np.allclose(tg.numpy(), full_g, rtol=1e-12) and abs(tl.item() - full_l) <= 1e-12 * full_l
# C2: K shared fp32 accumulators
idx, val = orc.accumulate_inputs(50_000, 1024, "zipf", seed=2)
b, e = par.shard_range(idx.size, rank, world)
t = torch.from_numpy(orc.accumulate(idx[b:e], val[b:e], 1024))
par.allreduce_shared_grads(t)
exact = orc.accumulate_exact(idx, val, 1024)
assert (np.abs(t.numpy() - exact) <= 1e-4 * orc.accumulate_exact(idx, np.abs(val), 1024) + 1e-30).all()
# C5: one view per rank, Gaussians replicated, gradient all-reduce == sum over views
W, H, N = 48, 32, 20
params, _ = orc.splat_scene(N, W, H, seed=7)
targets = [orc.splat_scene(1, W, H, seed=100 + v)[1] for v in range(world)]
g, o, l, _ = orc.splat(params, targets[rank], W, H, np.float64)
tg = torch.from_numpy(g.copy()); tl = torch.tensor([l], dtype=torch.float64)
par.allreduce_shared_grads(tg, tl)
want = sum(orc.splat(params, t_, W, H, np.float64)[0] for t_ in targets)
assert np.allclose(tg.numpy(), want, rtol=1e-12, atol=1e-12)
# C4 on G ranks: ONE image in tile-aligned row bands, Gaussians replicated; a band is rendered as its own image with the
# centres moved up by the band's first row (fp64: the move costs nothing); gradients and loss add up to the whole image's
W, H, N = 40, 48, 25
params, target = orc.splat_scene(N, W, H, seed=11)
params[:, 1] = np.round(params[:, 1] * 64.0) / 64.0      # so that cy - row_begin is exact in the fp32 parameter array
r0, r1 = par.row_bands(H, world)[rank]
shifted = params.copy(); shifted[:, 1] -= r0
g, o, l, _ = orc.splat(shifted, target.reshape(H, W, 3)[r0:r1].reshape(-1, 3).copy(), W, r1 - r0, np.float64)
tg = torch.from_numpy(g.copy()); tl = torch.tensor([l], dtype=torch.float64)
par.allreduce_shared_grads(tg, tl)
fg, fo, fl, _ = orc.splat(params, target, W, H, np.float64)
assert np.allclose(o, fo.reshape(H, W, 3)[r0:r1].reshape(-1, 3), rtol=1e-9, atol=1e-12)
assert np.allclose(tg.numpy(), fg, rtol=1e-9, atol=1e-12) and abs(tl.item() - fl) <= 1e-9 * abs(fl)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_world_size_2_gloo_shared_gradient_allreduce():
    with tempfile.TemporaryDirectory() as d:
        w = os.path.join(d, "worker.py")
        open(w, "w").write(WORKER)
        import socket
        with socket.socket() as sock:  # a port nobody holds right now (a fixed one may be taken by a parallel run)
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", str(port), w, ROOT],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
        assert res.returncode == 0, res.stdout[-3000:]
        assert res.stdout.count("ok") >= 2
