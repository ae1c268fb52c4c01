"""Multi-GPU parity (`-m gpu`, needs >= 2 GPUs; skipped on a single-GPU box): one process per GPU, launched with
torchrun.  One pytest test per configuration (tests/multi_worker.py holds the cases), so a failure in one does not hide the
others.  Every case runs SHARDED (element ranges / row bands / views) with its exchange of the shared-parameter gradients
and is compared with the same work done by ONE rank through the same C ABI, and with the fp64 oracle.
The worker's output (incl. its TIMING lines) is appended to gpurun_out/multi_gpu_<n>ranks.log when that directory exists."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multi_worker.py")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")]

CASES = ["c1", "c2", "c3", "c3b", "c4_rows", "c5_views", "comm_nccl", "peer_adam", "splat_peer_iteration"]


def run_case(case, port):
    n = min(torch.cuda.device_count(), 8)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, ROOT, case],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"multi_gpu_{n}ranks.log"), "a") as f:
            f.write(f"==== case {case}, {n} ranks, rc {res.returncode}\n")
            f.write("\n".join(l for l in res.stdout.splitlines() if l.startswith(("CASE", "TIMING")) or res.returncode) + "\n")
    assert res.returncode == 0, res.stdout[-4000:]
    assert res.stdout.count(f"CASE {case} rank") == n, res.stdout[-2000:]
    for line in res.stdout.splitlines():
        if line.startswith("TIMING"):
            print(line)


@pytest.mark.parametrize("case", CASES)
def test_sharded_case_matches_single_rank_and_oracle(case):
    run_case(case, 29633 + CASES.index(case))
