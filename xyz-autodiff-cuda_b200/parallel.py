"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed (NCCL on the GPU box, gloo in
the CPU tests).  SURVEY.md section 8e: every configuration shards by independent units, and the only exchange
is an all-reduce of the SHARED-parameter gradients:

  C1 least squares / C2 accumulation   contiguous element ranges per rank; all-reduce of 4 fp64 / K fp32 sums
  C3 covariance projection             contiguous element ranges per rank; NO collective (per-element gradients)
  C4 one image on G GPUs               tile-aligned row bands per rank, Gaussians replicated; all-reduce of the
                                       N x 9 gradient buffer and of the scalar loss
  C5 V views on G GPUs                 views round-robin per rank, Gaussians replicated; same all-reduce

The collectives run in place on the gradient tensors, on the current stream, right behind the kernels that
produced them (no host synchronisation in between).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist

TILE = 16


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of n independent elements for `rank`."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return (n * rank) // world, (n * (rank + 1)) // world


def row_bands(height: int, world: int) -> List[Tuple[int, int]]:
    """Tile-aligned row bands [row_begin, row_end) that partition an image of `height` rows over `world` ranks.
    Bands are multiples of 16 rows (the last one takes the remainder); ranks beyond the tile-row count get empty
    bands."""
    tiles_y = (height + TILE - 1) // TILE
    bands = []
    for r in range(world):
        t0, t1 = (tiles_y * r) // world, (tiles_y * (r + 1)) // world
        bands.append((min(height, t0 * TILE), min(height, t1 * TILE)))
    return bands


def views_for_rank(num_views: int, rank: int, world: int) -> List[int]:
    return list(range(rank, num_views, world))


def allreduce_shared_grads(*tensors: torch.Tensor) -> None:
    """In-place sum over ranks of the shared-parameter gradients (and loss scalars).  No-op without a process
    group or with a single rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def splat_iteration_sharded(x, params, grads, targets, outputs, loss, width, height, flags=0, mode="views"):
    """One forward+backward of the splat path on this rank's share, then the gradient all-reduce.
    mode == "views": `targets` / `outputs` are this rank's lists of per-view images (C5);
    mode == "rows" : one image, this rank renders its row band of targets[0] into outputs[0] (C4 on G GPUs).
    grads / loss must be zeroed by the caller (reference semantics)."""
    n = params.shape[0]
    if mode == "views":
        for tgt, out in zip(targets, outputs):
            x.launch_gaussian_splatting(params, grads, tgt, out, loss, width, height, n, flags)
    elif mode == "rows":
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        r0, r1 = row_bands(height, world)[rank]
        if r1 > r0:
            x.launch_gaussian_splatting(params, grads, targets[0], outputs[0], loss, width, height, n, flags, rows=(r0, r1))
    else:
        raise ValueError(mode)
    allreduce_shared_grads(grads, loss)


def make_peer_group(x):
    """PeerGroup over the ranks of the default process group (all on one NVSwitch box): the 64-byte CUDA IPC handles
    travel through one all_gather_object; everything after that is NVLink loads/stores issued by the kernels."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0

    def exchange(handle: bytes):
        if world == 1:
            return [handle]
        out = [None] * world
        dist.all_gather_object(out, handle)
        return out

    group = x.PeerGroup(rank, world, exchange)
    if world > 1:
        dist.barrier()  # every rank has opened every mailbox before the first kernel stores into them
    return group
