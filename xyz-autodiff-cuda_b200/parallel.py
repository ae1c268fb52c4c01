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


def balanced_row_bands(tile_row_cost, height: int, world: int) -> List[Tuple[int, int]]:
    """Tile-aligned row bands with EQUALISED cost instead of equal height: `tile_row_cost[r]` is the work of tile row r
    (e.g. its number of (tile, Gaussian) list entries, from one launch over the whole image).  Gaussians near the image
    border reach fewer tiles, so equal-height bands give the border ranks ~30 % less work than the middle ones and every
    iteration waits for the slowest rank.  Contiguous partition minimising the largest band (binary search on the bound);
    every rank computes the same bands from the same costs."""
    cost = [float(c) for c in tile_row_cost]
    n = len(cost)
    if n != (height + TILE - 1) // TILE:
        raise ValueError("one cost per tile row expected")
    if world >= n:
        cuts = list(range(n)) + [n] * (world - n + 1)
    else:
        def parts_needed(bound):
            parts, acc = 1, 0.0
            for c in cost:
                if acc + c > bound and acc > 0.0:
                    parts, acc = parts + 1, 0.0
                acc += c
            return parts
        lo, hi = max(cost + [0.0]), sum(cost) + 1.0
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            if parts_needed(mid) <= world:
                hi = mid
            else:
                lo = mid
        cuts, acc = [0], 0.0
        for r, c in enumerate(cost):
            remaining_rows, remaining_parts = n - r, world - (len(cuts) - 1)
            if r > cuts[-1] and (acc + c > hi or remaining_rows < remaining_parts) and len(cuts) < world:
                cuts.append(r)
                acc = 0.0
            acc += c
        while len(cuts) < world:   # fewer parts were needed than ranks: split the tail rows off one by one
            cuts.append(min(n, cuts[-1] + 1) if cuts[-1] < n else n)
        cuts = sorted(cuts) + [n]
    return [(min(height, cuts[r] * TILE), min(height, cuts[r + 1] * TILE)) for r in range(world)]


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


class ShardedSplatTrainer:
    """One rank of the reference's training loop (GaussianSplattingTrainer::train, gaussian_splatting_training.cu:127-158:
    zero gradients, reset the loss, launch, Adam) sharded over the GPUs of one box -- by VIEW (C5: `targets` = this rank's
    views, whole image each) or by ROW BAND of one image (C4 on G GPUs: `rows` = this rank's band).  Gaussians are
    replicated; per iteration the ranks exchange the N x 9 gradients (+ the loss):

      exchange = "nccl"          xyz_allreduce_grads (the library's own NCCL communicator) + Adam on every replica
                                 (the form BASELINE configs[4] names)
      exchange = "nccl_sharded"  xyz_adam_step_individual_sharded: reduce-scatter, Adam on the rank's range, all-gather
      exchange = "peer"          xyz_adam_step_individual_peer: ONE kernel over NVLink peer memory, no NCCL call; with
                                 `graph=True` the whole iteration is one captured CUDA graph (no host work per iteration)

    Every launch goes through a caller-owned workspace (no allocation / synchronisation inside the iteration)."""

    LR = (0.1, 0.01, 0.001, 0.02, 0.05)

    def __init__(self, x, params, targets, width, height, rank=0, world=1, rows=None, exchange="nccl", comm=None,
                 group=None, gather=None, flags=0, lr=None, max_entries=None):
        self.x, self.w, self.h, self.rank, self.world = x, width, height, rank, world
        self.exchange, self.comm, self.flags = exchange, comm, flags
        self.lr = tuple(lr) if lr is not None else self.LR
        self.targets = list(targets)
        dev = torch.device("cuda", torch.cuda.current_device())
        n = params.shape[0]
        self.n = n
        self.rows = rows
        self.ps = None
        if exchange == "peer":
            self.ps = x.PeerSplat(group, n, gather)
            self.params, self.grads, self.adam = self.ps.params, self.ps.grads, self.ps.adam
            self.params.copy_(params)
        else:
            if world > 1 and comm is None:
                raise ValueError("exchange over NCCL needs a Comm")
            self.params = params.clone()
            self.grads = torch.zeros((n, 9), dtype=torch.float32, device=dev)
            self.adam = torch.zeros((n, 18), dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.outputs = [torch.zeros((width * height, 3), dtype=torch.float32, device=dev) for _ in self.targets]
        self.empty = rows is not None and rows[1] <= rows[0]
        self.ws = None
        if self.targets and not self.empty:
            if max_entries is None:  # learn the list length of this scene once (one ordinary launch), + 30 % head-room
                scratch_g = torch.zeros((n, 9), dtype=torch.float32, device=dev)
                x.launch_gaussian_splatting(self.params, scratch_g, self.targets[0], self.outputs[0], self.loss, width, height,
                                            n, flags, rows=rows)
                max_entries = int(x.splat_last_stats()["entries"] * 1.3) + 4096
                self.loss.zero_()
            self.max_entries = max_entries
            self.ws = x.SplatWorkspace(width, height, n, max_entries, flags, rows=rows)
        self.graph = None
        self.stream = None

    def _enqueue(self, it, stream=None, flags=None):
        x = self.x
        if stream is None:
            self.loss.zero_()
        else:
            with torch.cuda.stream(stream):
                self.loss.zero_()
        if self.ws is not None:
            for tgt, out in zip(self.targets, self.outputs):
                self.ws.launch(self.params, self.grads, tgt, out, self.loss, flags=flags, stream=stream)
        if self.exchange == "peer":
            self.ps.adam_step(*self.lr, iteration=it, total_loss=self.loss, stream=stream)
        elif self.exchange == "nccl_sharded" and self.world > 1:
            self.comm.adam_step_individual_sharded(self.params, self.grads, self.adam, *self.lr, iteration=max(it, 1),
                                                   total_loss=self.loss, stream=stream)
        else:
            if self.world > 1:
                self.comm.allreduce_grads(self.grads, stream=stream)
                self.comm.allreduce_grads(self.loss, stream=stream)
            x.adam_step_individual(self.params, self.grads, self.adam, *self.lr, iteration=max(it, 1), stream=stream,
                                   zero_grads=True)

    def iteration(self, it, flags=None):
        """Enqueue iteration `it` (1-based) on the current stream; nothing waits for the GPU."""
        self._enqueue(it, None, flags)

    def capture(self):
        """Capture one iteration as a CUDA graph (exchange == "peer": the optimiser step counts its iterations on the
        device, so every replay is the next Adam step)."""
        if self.exchange != "peer":
            raise ValueError("graph capture needs the device-side step counter of the peer exchange")
        self.stream = torch.cuda.Stream()
        self.graph = torch.cuda.CUDAGraph()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="thread_local"):
                self._enqueue(0, self.stream)
        torch.cuda.current_stream().wait_stream(self.stream)

    def replay(self):
        self.graph.replay()

    def close(self):
        self.graph = None
        if self.ps is not None:
            self.ps.close()
            self.ps = None
