#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the xyz-autodiff-cuda hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload covproj|splat_c5|splat_c4_rows]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W        (N > 1)

Headline workload (BASELINE.json configs[2], the largest single-GPU "gradient evals/s" configuration and the
only one whose working set (12.9 GB) exceeds L2): batched covariance projection S' = (J W) S (J W)^T,
forward + reverse, 2^26 elements, fp32.  A step = one pass of the kernel over the batch.
  value     evals/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same through the host-buffer pipeline (pinned host arrays -> H2D -> kernel -> D2H)
  roofline  192 algorithmic bytes per eval / kernel time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own op::matmul graph compiled for the host (oracle/_ref), all host
            cores, on a bounded sample
The other BASELINE configs (least squares 1M, accumulation 16M->1K, splat 100K Gaussians 1024^2) are timed
briefly at N=1 and reported under "also" in the same JSON line.
--impl reference times the reference's CPU path alone (rank 0 only).
--workload splat_c5        BASELINE configs[4]: 3M Gaussians, 8 views sharded over the ranks, NCCL all-reduce of the gradients.
--workload splat_c4_rows   BASELINE configs[3] with ONE image split into row bands over the ranks (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fwd+bwd gradient evals/sec"
UNIT = "evals/s"
COVPROJ_BYTES = 192  # per eval: in J6 W9 S6 g3, out out3 gJ6 gW9 gS6, fp32
FULL_E = 1 << 26


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_covproj(sample_elems, repeats=1):
    """The reference's matmul graph on the host cores (oracle/_ref if present, else the port)."""
    import numpy as np
    import oracle_lib as orc
    which = "ref" if orc.have_ref() else "port"
    cores = os.cpu_count() or 1
    J, W, S, g = orc.covproj_inputs(sample_elems, seed=42)
    orc.covproj(J[:1024], W[:1024], S[:1024], g[:1024], np.float32, which=which, threads=cores)  # warm the library
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.covproj(J, W, S, g, np.float32, which=which, threads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": sample_elems / best, "unit": UNIT, "cores": cores,
            "kind": "reference" if which == "ref" else "port",
            "sample": f"{sample_elems} of the {FULL_E} elements (seed 42), fp32, std::thread x {cores}, "
                      f"{'reference op::matmul graph compiled for the host (oracle/_ref)' if which == 'ref' else 'oracle port'}"}, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 1 << 23
    times = []
    base = None
    for i in range(args.warmup + args.steps):
        base, dt = cpu_covproj(sample)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = sample / (ms / 1000.0)
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "covproj_fwd_bwd 3x3 chain S'=(JW)S(JW)^T, 2^26 elems (BASELINE configs[2]); "
                                   "each step = a bounded sample of 2^23 elements on the host cores"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def time_kernel(fn, steps, warmup, stream):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record(stream)
    for i in range(steps):
        fn()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return ev[0].elapsed_time(ev[steps]), per


def also_workloads(dev, peak_gbs):
    """Brief device-resident timings of the other BASELINE configs (N=1 only)."""
    import numpy as np
    import torch
    import oracle_lib as orc
    import xyz_autodiff_cuda_b200 as x
    out = {}
    st = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def timed(fn, reps=10, flush_l2=True):
        ts = []
        for i in range(reps + 3):
            if flush_l2:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            fn()
            b.record(st)
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    def timed_rotating(make_call, n_sets, rounds=6):
        """Back-to-back launches over n_sets DIFFERENT input sets (together larger than L2, so every launch still reads
        from HBM) between one pair of events: the steady-state cost per launch, without the event / launch gap that a
        single ~20 us launch between two events carries."""
        calls = [make_call(i) for i in range(n_sets)]
        for c in calls:
            c()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(rounds):
            for c in calls:
                c()
        b.record(st)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / (rounds * n_sets)

    # C1: least squares, 1M residuals fp64 (24 MB)
    n = 1_000_000
    data = torch.from_numpy(orc.lsq_data(n, 42)).to(dev)
    prm = torch.zeros(8, dtype=torch.float64, device=dev)
    prm[1] = 1.0
    ms = timed(lambda: x.lsq_grad(data, prm))
    out["c1_lsq_1M_f64"] = {"evals_per_s": n / (ms / 1e3), "ms": ms, "gbs": 24 * n / (ms / 1e3) / 1e9,
                            "hbm_frac": 24 * n / (ms / 1e3) / 1e9 / peak_gbs, "l2": "flushed between iterations"}
    sets = [data] + [data.clone() for _ in range(7)]   # 8 x 24 MB = 192 MB > L2
    ms_rot = timed_rotating(lambda i: (lambda: x.lsq_grad(sets[i], prm)), len(sets))
    out["c1_lsq_1M_f64"].update({"ms_back_to_back": ms_rot, "hbm_frac_back_to_back": 24 * n / (ms_rot / 1e3) / 1e9 / peak_gbs,
                                 "back_to_back": "8 input sets (192 MB > L2) launched back to back, 48 launches between two events"})
    del sets
    n2 = 1 << 28
    data2 = torch.empty((n2, 3), dtype=torch.float64, device=dev).uniform_(-5, 5)
    ms = timed(lambda: x.lsq_grad(data2, prm), reps=5, flush_l2=False)
    out["c1_lsq_2^28_f64"] = {"evals_per_s": n2 / (ms / 1e3), "ms": ms, "gbs": 24 * n2 / (ms / 1e3) / 1e9,
                              "hbm_frac": 24 * n2 / (ms / 1e3) / 1e9 / peak_gbs, "l2": "6.4 GB input > L2"}
    del data2
    # C2: accumulation 2^24 -> 1024
    n = 1 << 24
    for dist in ("uniform", "zipf", "same"):
        idx, val = orc.accumulate_inputs(n, 1024, dist, 42)
        ti, tv = torch.from_numpy(idx).to(dev), torch.from_numpy(val).to(dev)
        grad = torch.zeros(1024, device=dev)
        ms = timed(lambda: x.accumulate(ti, tv, grad))
        out[f"c2_accumulate_2^24_{dist}"] = {"elems_per_s": n / (ms / 1e3), "ms": ms, "gbs": 8 * n / (ms / 1e3) / 1e9,
                                             "hbm_frac": 8 * n / (ms / 1e3) / 1e9 / peak_gbs,
                                             "l2": "flushed between iterations"}
        if dist == "uniform":
            pairs = [(ti, tv)] + [(ti.clone(), tv.clone()) for _ in range(7)]   # 8 x 134 MB > L2
            ms_rot = timed_rotating(lambda i: (lambda: x.accumulate(pairs[i][0], pairs[i][1], grad)), len(pairs))
            out[f"c2_accumulate_2^24_{dist}"].update({
                "ms_back_to_back": ms_rot, "hbm_frac_back_to_back": 8 * n / (ms_rot / 1e3) / 1e9 / peak_gbs,
                "back_to_back": "8 input sets (1.07 GB > L2) launched back to back, 48 launches between two events"})
            del pairs
    # C3 variant B: one shared W, per-element adjoints of W accumulated into 9 gradients; 2^26 elements (8 GB > L2)
    n3 = 1 << 26
    ins3 = [torch.empty((n3, w), device=dev).uniform_(-1, 1) for w in (6, 6, 3)]
    outs3 = [torch.empty((n3, w), device=dev) for w in (3, 6, 6)]
    w9 = torch.empty(9, device=dev).uniform_(-1, 1)
    gw9 = torch.zeros(9, device=dev)
    ms = timed(lambda: x.covproj_shared_w_fwd_bwd(ins3[0], w9, ins3[1], ins3[2], outs3[0], outs3[1], gw9, outs3[2]),
               reps=5, flush_l2=False)
    out["c3b_covproj_shared_w_2^26"] = {"evals_per_s": n3 / (ms / 1e3), "ms": ms, "bytes_per_eval": 120,
                                        "gbs": 120 * n3 / (ms / 1e3) / 1e9, "hbm_frac": 120 * n3 / (ms / 1e3) / 1e9 / peak_gbs,
                                        "l2": "8.05 GB per step > L2"}
    del ins3, outs3
    torch.cuda.empty_cache()
    # the same two graphs written by a USER with the public op:: API and run through include/xyz_autodiff/batched.cuh
    # (tests/csrc/batched_probe.cu; the probe library is built by __graft_entry__.build())
    try:
        import ctypes
        so = os.path.join(ROOT, "tests", "csrc", "_build", "libxyz_batched.so")
        if os.path.exists(so):
            B = ctypes.CDLL(so)
            B.batched_chain.restype = ctypes.c_float
            B.batched_chain.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
            B.batched_lsq.restype = ctypes.c_float
            B.batched_lsq.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
            nb = 1 << 26
            tin = torch.empty((nb, 15), device=dev).uniform_(-1, 1)
            tout = torch.empty((nb, 15), device=dev)
            w18 = torch.empty(18, device=dev).uniform_(-1, 1)
            gw18 = torch.zeros(18, device=dev)
            B.batched_chain(tin.data_ptr(), tout.data_ptr(), nb, w18.data_ptr(), gw18.data_ptr(), 2)
            ms = B.batched_chain(tin.data_ptr(), tout.data_ptr(), nb, w18.data_ptr(), gw18.data_ptr(), 5)
            out["batched_for_each_matmul_chain_shared_w_2^26"] = {
                "evals_per_s": nb / (ms / 1e3), "ms": ms, "gbs": 120 * nb / (ms / 1e3) / 1e9,
                "hbm_frac": 120 * nb / (ms / 1e3) / 1e9 / peak_gbs,
                "what": "user graph of op::matmul nodes through batched::for_each (include/xyz_autodiff/batched.cuh)"}
            del tin, tout
            torch.cuda.empty_cache()
            mb = 1 << 28
            dpts = torch.empty((mb, 3), dtype=torch.float64, device=dev).uniform_(-5, 5)
            v5 = torch.tensor([0.0, 1.0, 0.0, 0.0, 0.0], dtype=torch.float64, device=dev)
            g5 = torch.zeros(5, dtype=torch.float64, device=dev)
            B.batched_lsq(dpts.data_ptr(), mb, v5.data_ptr(), g5.data_ptr(), 2)
            ms = B.batched_lsq(dpts.data_ptr(), mb, v5.data_ptr(), g5.data_ptr(), 5)
            out["batched_for_each_least_squares_2^28_f64"] = {
                "evals_per_s": mb / (ms / 1e3), "ms": ms, "gbs": 24 * mb / (ms / 1e3) / 1e9,
                "hbm_frac": 24 * mb / (ms / 1e3) / 1e9 / peak_gbs,
                "what": "user graph of op:: nodes (fp64) through batched::for_each"}
            del dpts
            torch.cuda.empty_cache()
    except Exception as e:
        out["batched_for_each"] = {"error": repr(e)}
    # C4: splat 100K Gaussians, 1024^2
    W = H = 1024
    N = 100_000
    params, target = orc.splat_c4_scene(N, W, H, 42)
    tp, tt = torch.from_numpy(params).to(dev), torch.from_numpy(target).to(dev)
    grads = torch.zeros((N, 9), device=dev)
    img = torch.zeros((W * H, 3), device=dev)
    loss = torch.zeros(1, device=dev)

    def splat_iter():
        x.zero_gradients(grads)
        loss.zero_()
        x.launch_gaussian_splatting(tp, grads, tt, img, loss, W, H, N)

    ms = timed(splat_iter, reps=5, flush_l2=False)
    stats = x.splat_last_stats()
    out["c4_splat_100K_1024x1024"] = {"ms_per_iter": ms, "tile_list_entries": stats["entries"],
                                      "pairs_per_pass": stats["pairs_per_pass"],
                                      "pair_evals_per_s": 2 * stats["pairs_per_pass"] / (ms / 1e3),
                                      "reference_pairs_per_pass": N * W * H,
                                      "iteration": "zero_grad + loss reset + launch (fwd + bwd), fast-math flavour"}
    def splat_iter_tail():
        x.zero_gradients(grads)
        loss.zero_()
        x.launch_gaussian_splatting(tp, grads, tt, img, loss, W, H, N, x.FLAG_TAIL_CULL)

    ms_tail = timed(splat_iter_tail, reps=5, flush_l2=False)
    out["c4_splat_100K_1024x1024_tail_cull_opt_in"] = {
        "ms_per_iter": ms_tail, "tile_list_entries": x.splat_last_stats()["entries"],
        "note": "XYZ_FLAG_TAIL_CULL: NOT the parity path -- also skips pairs with weight < exp(-28) (bounded error, "
                "include/xyz_b200.h); the headline c4 number above is the result-preserving default"}
    out["reference_cuda_same_b200"] = reference_cuda(dev, timed, x, tp, tt, W, H, N, ms)
    try:
        out["reference_cpu_host_path"] = cpu_reference_other_configs()
    except Exception as e:
        out["reference_cpu_host_path"] = {"error": repr(e)}
    return out


def cpu_reference_other_configs():
    """The reference's host-compiled path (oracle/_ref, else the port) on all host cores for the configs that are
    not the headline: C1 at full size, C2 on a 2^22-element sample, C4 on a reduced shape scaled by the pair count
    (the reference evaluates every (pixel, Gaussian) pair, so its cost is exactly proportional to pairs)."""
    import numpy as np
    import oracle_lib as orc
    which = "ref" if orc.have_ref() else "port"
    cores = os.cpu_count() or 1
    out = {"kind": "reference" if which == "ref" else "port", "cores": cores}

    def best(fn, reps=2):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    data = orc.lsq_data(1_000_000, 42)
    t = best(lambda: orc.lsq_grad(data, (0.0, 1.0, 0.0, 0.0), which=which, threads=cores))
    out["c1_lsq_1M_f64"] = {"evals_per_s": data.shape[0] / t, "ms": t * 1e3, "sample": "all 10^6 points"}
    n = 1 << 22
    idx, val = orc.accumulate_inputs(n, 1024, "uniform", 42)
    t = best(lambda: orc.accumulate(idx, val, 1024, which=which, threads=cores))
    out["c2_accumulate_uniform"] = {"elems_per_s": n / t, "ms_scaled_to_2^24": t * 1e3 * 4, "sample": "2^22 of 2^24 elements"}
    W = H = 256
    N = 1000
    params, target = orc.splat_c4_scene(N, W, H, 42)
    t = best(lambda: orc.splat(params, target, W, H, np.float32, which=which, threads=cores), reps=1)
    pairs = N * W * H
    full_pairs = 100_000 * 1024 * 1024
    out["c4_splat"] = {"pairs_per_s_both_passes": pairs / t, "ms_at_sample": t * 1e3,
                       "ms_per_iter_scaled_to_100K_1024x1024": t * 1e3 * full_pairs / pairs,
                       "sample": f"{N} Gaussians on {W}x{H} ({pairs} pairs), scaled by pair count"}
    return out


def reference_cuda(dev, timed, x, tp, tt, W, H, N, ours_splat_ms):
    """The reference's own CUDA kernels built unmodified for sm_100a (oracle/_ref/libxyz_ref_cuda.so), timed on
    the same GPU with the same CUDA-event harness.  Splat is run at a reduced Gaussian count and scaled by N
    (its cost is exactly linear in N: every pixel visits every Gaussian)."""
    import ctypes
    import numpy as np
    import torch
    import oracle_lib as orc
    path = os.path.join(ROOT, "oracle", "_ref", "libxyz_ref_cuda.so")
    if not os.path.exists(path):
        return {"unavailable": "oracle/_ref/libxyz_ref_cuda.so not built (needs /root/reference at build time)"}
    R = ctypes.CDLL(path)
    vp, ll = ctypes.c_void_p, ctypes.c_longlong
    R.refcuda_splat.argtypes = [vp] * 5 + [ctypes.c_int] * 3
    R.refcuda_lsq.argtypes = [vp, ll, vp]
    R.refcuda_accumulate.argtypes = [vp, vp, ll, vp]
    R.refcuda_covproj.argtypes = [vp] * 8 + [ll]
    res = {}
    # splat, reduced N
    n_ref = 256
    grads = torch.zeros((n_ref, 9), device=dev)
    img = torch.zeros((W * H, 3), device=dev)
    loss = torch.zeros(1, device=dev)
    sub = tp[:n_ref].contiguous()
    ms = timed(lambda: R.refcuda_splat(sub.data_ptr(), grads.data_ptr(), tt.data_ptr(), img.data_ptr(), loss.data_ptr(),
                                       W, H, n_ref), reps=3, flush_l2=False)
    scaled = ms * N / n_ref
    res["c4_splat"] = {"ms_at_N": {str(n_ref): ms}, "ms_per_iter_scaled_to_100K": scaled,
                       "speedup_of_this_repo": scaled / ours_splat_ms,
                       "note": "reference kernel cost is linear in N (all pairs, 9 atomics per pair)"}
    # the same reference sources compiled against THIS repo's headers (warp-aggregated add_grad): drop-in speed-up of
    # unmodified user kernels
    path2 = os.path.join(ROOT, "oracle", "_ref", "libxyz_ref_cuda_ourhdr.so")
    R2 = None
    if os.path.exists(path2):
        R2 = ctypes.CDLL(path2)
        R2.refcuda_splat.argtypes = [vp] * 5 + [ctypes.c_int] * 3
        R2.refcuda_lsq.argtypes = [vp, ll, vp]
        R2.refcuda_accumulate.argtypes = [vp, vp, ll, vp]
        ms2 = timed(lambda: R2.refcuda_splat(sub.data_ptr(), grads.data_ptr(), tt.data_ptr(), img.data_ptr(), loss.data_ptr(),
                                             W, H, n_ref), reps=3, flush_l2=False)
        res["c4_splat"]["reference_source_on_this_repos_headers_ms_at_N"] = {str(n_ref): ms2}
        res["c4_splat"]["header_drop_in_speedup"] = ms / ms2
    # covproj 2^24
    n = 1 << 24
    ins = [torch.empty((n, w), device=dev).uniform_(-1, 1) for w in (6, 9, 6, 3)]
    outs = [torch.empty((n, w), device=dev) for w in (3, 6, 9, 6)]
    ms_ref = timed(lambda: R.refcuda_covproj(*[t.data_ptr() for t in ins], *[t.data_ptr() for t in outs], n), reps=5,
                   flush_l2=False)
    ms_our = timed(lambda: x.covproj_fwd_bwd(*ins, *outs), reps=5, flush_l2=False)
    res["c3_covproj_2^24"] = {"reference_ms": ms_ref, "this_repo_ms": ms_our, "speedup_of_this_repo": ms_ref / ms_our}
    del ins, outs
    # least squares 1M
    n = 1_000_000
    data = torch.from_numpy(orc.lsq_data(n, 42)).to(dev)
    prm = torch.zeros(8, dtype=torch.float64, device=dev)
    prm[1] = 1.0
    ms_ref = timed(lambda: R.refcuda_lsq(data.data_ptr(), n, prm.data_ptr()), reps=5)
    ms_our = timed(lambda: x.lsq_grad(data, prm), reps=5)
    res["c1_lsq_1M"] = {"reference_ms": ms_ref, "this_repo_ms": ms_our, "speedup_of_this_repo": ms_ref / ms_our}
    if R2 is not None:
        res["c1_lsq_1M"]["reference_source_on_this_repos_headers_ms"] = timed(
            lambda: R2.refcuda_lsq(data.data_ptr(), n, prm.data_ptr()), reps=5)
    # accumulation 2^24 -> 1024, uniform ids
    n = 1 << 24
    idx, val = orc.accumulate_inputs(n, 1024, "uniform", 42)
    ti, tv = torch.from_numpy(idx).to(dev), torch.from_numpy(val).to(dev)
    grad = torch.zeros(1024, device=dev)
    ms_ref = timed(lambda: R.refcuda_accumulate(ti.data_ptr(), tv.data_ptr(), n, grad.data_ptr()), reps=5)
    ms_our = timed(lambda: x.accumulate(ti, tv, grad), reps=5)
    res["c2_accumulate_2^24_uniform"] = {"reference_ms": ms_ref, "this_repo_ms": ms_our,
                                         "speedup_of_this_repo": ms_ref / ms_our}
    if R2 is not None:
        res["c2_accumulate_2^24_uniform"]["reference_source_on_this_repos_headers_ms"] = timed(
            lambda: R2.refcuda_accumulate(ti.data_ptr(), tv.data_ptr(), n, grad.data_ptr()), reps=5)
    return res


def run_splat_c5(args):
    """BASELINE configs[4]: N Gaussians replicated, V views (targets) sharded round-robin over the ranks; one
    iteration = zero_grad + this rank's views (forward + backward each) + NCCL all-reduce of the N x 9 gradient
    buffer and the loss + Adam on every replica.  Strong scaling: V is fixed, ranks share it.
    The reference has one target image only; view v = the reference's test image rolled by 64 v pixels in x and
    37 v in y (our definition)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle_lib as orc
    import xyz_autodiff_cuda_b200 as x
    from importlib import import_module
    par = import_module("xyz_autodiff_cuda_b200.parallel")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = H = 1024
    N, V = args.gaussians, args.views
    params, target = orc.splat_c4_scene(N, W, H, 42)
    tp = torch.from_numpy(params).to(dev)
    base = torch.from_numpy(target).to(dev).reshape(H, W, 3)
    mine = par.views_for_rank(V, rank, world)
    targets = [torch.roll(base, shifts=(37 * v, 64 * v), dims=(0, 1)).reshape(W * H, 3).contiguous() for v in mine]
    outs = [torch.zeros((W * H, 3), device=dev) for _ in mine]
    grads = torch.zeros((N, 9), device=dev)
    loss = torch.zeros(1, device=dev)
    adam = torch.zeros((N, 18), device=dev)
    st = torch.cuda.current_stream()
    ar_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def step(it):
        x.zero_gradients(grads)
        loss.zero_()
        for t, o in zip(targets, outs):
            x.launch_gaussian_splatting(tp, grads, t, o, loss, W, H, N)
        ar_ev[0].record(st)
        par.allreduce_shared_grads(grads, loss)
        ar_ev[1].record(st)
        x.adam_step_individual(tp, grads, adam, 0.1, 0.01, 0.001, 0.02, 0.05, iteration=it + 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    x.reset_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record(st)
    for i in range(args.steps):
        step(i)
    b.record(st)
    barrier()
    launches = x.launch_count()
    t = torch.tensor([a.elapsed_time(b), ar_ev[0].elapsed_time(ar_ev[1])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t[0].item() / args.steps
    if rank == 0:
        stats = x.splat_last_stats()
        print(json.dumps({
            "metric": "splat fwd+bwd ms/iter", "value": ms, "unit": "ms/iter", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"mini-gaussian-splatting {N} Gaussians, {V} views of 1024x1024 sharded over "
                                   f"{world} GPUs, NCCL all-reduce of N x 9 grads (BASELINE configs[4])",
                       "views_per_gpu": len(mine), "allreduce_bytes": N * 36 + 4,
                       "allreduce_ms_last_iter_max_over_ranks": t[1].item(),
                       "tile_list_entries_per_view": stats["entries"], "l2": "per-iteration working set "
                       f"{(stats['entries'] * 16 + N * 160) / 1e6:.0f} MB"},
            "gpu_launches": int(launches)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



def run_splat_c4_rows(args):
    """BASELINE configs[3] on G GPUs (north_star: partition "by image tile"): ONE 1024 x 1024 target, Gaussians replicated,
    every rank renders and back-propagates a tile-aligned band of rows, then the N x 9 gradients + loss are
    all-reduced (NCCL, in place) and Adam runs on every replica.  Strong scaling."""
    import torch
    import torch.distributed as dist
    import oracle_lib as orc
    import xyz_autodiff_cuda_b200 as x
    from importlib import import_module
    par = import_module("xyz_autodiff_cuda_b200.parallel")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = H = 1024
    N = args.gaussians if args.gaussians != 3_000_000 else 100_000
    params, target = orc.splat_c4_scene(N, W, H, 42)
    tp, tt = torch.from_numpy(params).to(dev), torch.from_numpy(target).to(dev)
    out = torch.zeros((W * H, 3), device=dev)
    grads = torch.zeros((N, 9), device=dev)
    loss = torch.zeros(1, device=dev)
    adam = torch.zeros((N, 18), device=dev)
    st = torch.cuda.current_stream()

    def step(it):
        x.zero_gradients(grads)
        loss.zero_()
        par.splat_iteration_sharded(x, tp, grads, [tt], [out], loss, W, H, mode="rows")
        x.adam_step_individual(tp, grads, adam, 0.1, 0.01, 0.001, 0.02, 0.05, iteration=it + 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    x.reset_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record(st)
    for i in range(args.steps):
        step(i)
    b.record(st)
    barrier()
    launches = x.launch_count()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t[0].item() / args.steps
    if rank == 0:
        print(json.dumps({
            "metric": "splat fwd+bwd ms/iter", "value": ms, "unit": "ms/iter", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"mini-gaussian-splatting {N} Gaussians, one 1024x1024 image split into {world} row bands, "
                                   f"NCCL all-reduce of N x 9 grads + Adam on every replica (BASELINE configs[3] on {world} GPUs)",
                       "rows_per_gpu": H // world, "allreduce_bytes": N * 36 + 4,
                       "l2": "per-iteration working set (entries + records + rest tiles) streamed once; Adam moves the Gaussians every step"},
            "gpu_launches": int(launches)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ours(args):
    import torch
    import torch.distributed as dist
    import xyz_autodiff_cuda_b200 as x
    from importlib import import_module
    host_api = import_module("xyz_autodiff_cuda_b200.host_api")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    x.lib()
    peak_gbs, peak_src = measured_peak_gbs()

    # weak scaling: every rank owns E elements; the path has no data-path collective (per-element gradients)
    E = args.elems
    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    ins = [torch.empty((E, w), dtype=torch.float32, device=dev) for w in (6, 9, 6, 3)]
    for t in ins:
        t.uniform_(-1.0, 1.0, generator=gen)
    outs = [torch.empty((E, w), dtype=torch.float32, device=dev) for w in (3, 6, 9, 6)]
    st = torch.cuda.current_stream()

    def step():
        x.covproj_fwd_bwd(*ins, *outs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    x.reset_launch_count()
    barrier()
    total_ms, per = time_kernel(step, args.steps, 0, st)
    launches = x.launch_count()
    barrier()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    ms_per_step = total_ms / args.steps
    value = world * E / (ms_per_step / 1e3)
    kernel_ms = sum(per) / len(per)  # one launch per step: the kernel's average launch duration on this rank
    achieved = COVPROJ_BYTES * E / (kernel_ms / 1e3) / 1e9

    # e2e: pinned host arrays -> chunked H2D / kernel / D2H pipeline, same metric
    e2e_elems = min(E, args.e2e_elems)
    del outs
    h_in = [torch.empty((e2e_elems, w), dtype=torch.float32).pin_memory() for w in (6, 9, 6, 3)]
    for h, d in zip(h_in, ins):
        h.copy_(d[:e2e_elems])
    h_out = [torch.empty((e2e_elems, w), dtype=torch.float32).pin_memory() for w in (3, 6, 9, 6)]
    pipe = host_api.CovprojHostPipeline(dev)
    h2d = d2h = 0
    for _ in range(2):
        h2d, d2h = pipe.run(h_in, h_out)
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(e2e_steps):
        pipe.run(h_in, h_out)
    b.record(st)
    barrier()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_elems * e2e_steps / (t.item() / 1e3)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "covproj_fwd_bwd 3x3 chain S'=(JW)S(JW)^T fwd+bwd (BASELINE configs[2])",
                       "elems_per_gpu": E, "bytes_per_eval": COVPROJ_BYTES, "l2": "inputs larger than L2 "
                       f"({COVPROJ_BYTES * E / 1e9:.1f} GB per step per GPU)", "parallelism": f"dp{world} by element range, no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "traffic": None, "peak_source": peak_src, "kernel": "covproj_tma_kernel",
                         "kernel_ms": kernel_ms,
                         # also against the nominal HBM3e figure (not the roofline denominator)
                         "frac_of_nominal_8000_gbs": achieved / 8000.0},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "elems_per_step": e2e_elems, "steps": e2e_steps,
                    "api": "host_api.CovprojHostPipeline (pinned host -> H2D -> xyz_covproj_fwd_bwd_f32 -> D2H, 3 streams)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1:
            base, _ = cpu_covproj(1 << 23)
            line["cpu_baseline"] = base
            if not args.no_also:
                del ins, h_in, h_out, pipe
                torch.cuda.empty_cache()
                try:
                    line["also"] = also_workloads(dev, peak_gbs)
                except Exception as e:  # the headline must still be printed
                    line["also"] = {"error": repr(e)}
        traffic_file = os.path.join(ROOT, "profiles", "covproj_traffic.json")
        if os.path.exists(traffic_file):
            try:
                with open(traffic_file) as f:
                    tr = json.load(f)
                line["roofline"]["traffic"] = tr["dram_bytes_per_eval"] * E
                line["roofline"]["traffic_source"] = tr.get("source")
            except Exception:
                pass
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elems", type=int, default=FULL_E, help="elements per GPU (default 2^26, BASELINE configs[2])")
    ap.add_argument("--e2e-elems", type=int, default=1 << 25, help="elements per e2e step (host pinned memory bound)")
    ap.add_argument("--no-also", action="store_true", help="skip the brief timings of the other configs")
    ap.add_argument("--workload", default="covproj", choices=["covproj", "splat_c5", "splat_c4_rows"],
                    help="covproj = the headline (BASELINE configs[2]); splat_c5 = configs[4], views sharded over the ranks; "
                         "splat_c4_rows = configs[3] with ONE image split into row bands over the ranks")
    ap.add_argument("--gaussians", type=int, default=3_000_000, help="splat_c5: number of Gaussians")
    ap.add_argument("--views", type=int, default=8, help="splat_c5: number of views (targets)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "splat_c5":
        run_splat_c5(args)
    elif args.workload == "splat_c4_rows":
        run_splat_c4_rows(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
