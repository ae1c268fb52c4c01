"""xyz-autodiff-cuda_b200 -- host-side mirror (Python) of the C ABI in include/xyz_b200.h.

The product is libxyz_b200.so (hand-written sm_100a kernels, xyz-autodiff-cuda_b200/csrc/).  This
module only binds it with ctypes and passes device pointers of torch tensors; torch is plumbing
(device memory, streams, torch.distributed), not the compute path.  There is NO CPU fallback: if
the library is missing or a kernel launch fails, a RuntimeError is raised.

Names follow the reference's own host API where it has one:
  launch_gaussian_splatting   examples/mini-gaussian-splatting/gaussian_splatting_kernel.cuh:38-47
  zero_gradients / adam_step_individual / adam_step
                              examples/mini-gaussian-splatting/gaussian_parameters.h:78-91
  GaussianParams layout       examples/mini-gaussian-splatting/gaussian_parameters.h:12-41
  DataPoint / Parameters      examples/optimization/linear_regression_sgd.cu:31-39
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XYZ_B200_LIB") or os.path.join(_HERE, "lib", "libxyz_b200.so")  # override: kernel variants (dev/)

FLAG_DETERMINISTIC = 1
FLAG_PRECISE_MATH = 2
FLAG_RESIDUAL_ONLY = 4
FLAG_NO_CULL = 8
FLAG_IMPLICIT_IDS = 16
FLAG_TAIL_CULL = 32
FLAG_RADIX_BINNING = 64
FLAG_ASYNC = 128
FLAG_LSQ_SHIPPED_GRAPH = 256
FLAG_TIMING = 512
FLAG_BWD_ALL_PAIRS = 1024

GAUSSIAN_FLOATS = 9   # center[2] scale[2] rotation[1] color[3] opacity[1]
ADAM_FLOATS = 18
EXPORTS = [
    "xyz_b200_version", "xyz_b200_shutdown", "xyz_b200_launch_count", "xyz_b200_reset_launch_count",
    "xyz_lsq_grad_f64", "xyz_lsq_sgd_update_f64", "xyz_lsq_select_batch", "xyz_lsq_sgd_step_f64", "xyz_lsq_sgd_run_f64",
    "xyz_peer_mailbox_bytes", "xyz_peer_mailbox_create", "xyz_peer_mailbox_open", "xyz_peer_mailbox_close",
    "xyz_peer_mailbox_destroy", "xyz_lsq_grad_f64_allreduce", "xyz_accumulate_f32_allreduce",
    "xyz_accumulate_f32", "xyz_accumulate_f64", "xyz_covproj_fwd_bwd_f32", "xyz_covproj_shared_w_fwd_bwd_f32",
    "xyz_covproj_shared_w_fwd_bwd_f32_allreduce",
    "xyz_launch_gaussian_splatting", "xyz_launch_gaussian_splatting_rows", "xyz_splat_last_stats",
    "xyz_splat_last_backward_stats",
    "xyz_splat_debug_binning", "xyz_zero_gradients", "xyz_adam_step_individual", "xyz_adam_step",
    "xyz_adam_step_individual_zero_grads",
    "xyz_splat_workspace_bytes", "xyz_splat_workspace_init", "xyz_launch_gaussian_splatting_ws", "xyz_splat_workspace_status",
    "xyz_comm_unique_id", "xyz_comm_init_rank", "xyz_comm_init", "xyz_comm_init_all", "xyz_comm_destroy", "xyz_comm_rank",
    "xyz_comm_world", "xyz_comm_group_start", "xyz_comm_group_end", "xyz_allreduce_grads", "xyz_allreduce_f64",
    "xyz_adam_step_individual_sharded", "xyz_peer_alloc", "xyz_adam_step_individual_peer", "xyz_splat_last_timing",
]

_lib = None
_vp, _ll, _i = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
PEER_MAX_WORLD = 8


class PeerGroupStruct(ctypes.Structure):
    """xyz_peer_group (include/xyz_b200.h)."""
    _fields_ = [("mailbox", ctypes.c_void_p * PEER_MAX_WORLD), ("rank", ctypes.c_int), ("world", ctypes.c_int)]


class PeerSplatBuffersStruct(ctypes.Structure):
    """xyz_peer_splat_buffers (include/xyz_b200.h)."""
    _fields_ = [("params", ctypes.c_void_p * PEER_MAX_WORLD), ("grads", ctypes.c_void_p * PEER_MAX_WORLD)]



def lib() -> ctypes.CDLL:
    """Load libxyz_b200.so (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  xyz-autodiff-cuda_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.xyz_b200_version.restype = ctypes.c_char_p
        L.xyz_b200_launch_count.restype = ctypes.c_uint64
        L.xyz_b200_reset_launch_count.restype = None
        L.xyz_lsq_grad_f64.argtypes = [_vp, _ll, _vp, _vp, _vp, _i]
        L.xyz_lsq_sgd_update_f64.argtypes = [_vp, ctypes.c_double, _ll, _vp]
        L.xyz_lsq_select_batch.argtypes = [_vp, _ll, _vp, _ll, ctypes.c_uint64, ctypes.c_uint64, _vp]
        L.xyz_lsq_sgd_step_f64.argtypes = [_vp, _ll, _vp, _ll, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_double, _vp, _vp, _i]
        L.xyz_lsq_sgd_run_f64.argtypes = [_vp, _ll, _vp, _ll, ctypes.c_uint64, ctypes.c_uint64, _i, ctypes.POINTER(ctypes.c_double),
                                          _vp, _vp, _i]
        L.xyz_peer_mailbox_bytes.restype = ctypes.c_size_t
        L.xyz_peer_mailbox_create.argtypes = [ctypes.POINTER(_vp), ctypes.c_char_p]
        L.xyz_peer_mailbox_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(_vp)]
        L.xyz_peer_mailbox_close.argtypes = [_vp]
        L.xyz_peer_mailbox_destroy.argtypes = [_vp]
        L.xyz_lsq_grad_f64_allreduce.argtypes = [_vp, _ll, _vp, _vp, ctypes.POINTER(PeerGroupStruct), ctypes.c_ulonglong,
                                                 _vp, _i]
        L.xyz_accumulate_f32_allreduce.argtypes = [_vp, _vp, _ll, _vp, _i, ctypes.POINTER(PeerGroupStruct),
                                                   ctypes.c_ulonglong, _vp, _i]
        L.xyz_accumulate_f32.argtypes = [_vp, _vp, _ll, _vp, _i, _vp, _i]
        L.xyz_accumulate_f64.argtypes = [_vp, _vp, _ll, _vp, _i, _vp, _i]
        L.xyz_covproj_fwd_bwd_f32.argtypes = [_vp] * 8 + [_ll, _vp, _i]
        L.xyz_covproj_shared_w_fwd_bwd_f32.argtypes = [_vp] * 8 + [_ll, _vp, _i]
        L.xyz_covproj_shared_w_fwd_bwd_f32_allreduce.argtypes = [_vp] * 8 + [_ll, ctypes.POINTER(PeerGroupStruct),
                                                                 ctypes.c_ulonglong, _vp, _i]
        L.xyz_launch_gaussian_splatting.argtypes = [_vp] * 5 + [_i, _i, _i, _vp, _i]
        L.xyz_launch_gaussian_splatting_rows.argtypes = [_vp] * 5 + [_i, _i, _i, _i, _i, _vp, _i]
        L.xyz_splat_last_stats.argtypes = [_vp]
        L.xyz_splat_last_backward_stats.argtypes = [_vp]
        L.xyz_splat_debug_binning.argtypes = [_vp, _vp, _vp, _vp]
        L.xyz_zero_gradients.argtypes = [_vp, _i, _vp]
        L.xyz_adam_step_individual.argtypes = [_vp, _vp, _vp, _i, _vp, ctypes.c_float, ctypes.c_float,
                                               ctypes.c_float, _i, _vp]
        L.xyz_adam_step_individual_zero_grads.argtypes = L.xyz_adam_step_individual.argtypes
        L.xyz_adam_step.argtypes = [_vp, _vp, _vp, _i, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                    ctypes.c_float, _i, _vp]
        L.xyz_splat_workspace_bytes.restype = ctypes.c_size_t
        L.xyz_splat_workspace_bytes.argtypes = [_i, _i, _i, _i, _i, _ll, _i]
        L.xyz_splat_workspace_init.argtypes = [_vp, ctypes.c_size_t, _vp]
        L.xyz_launch_gaussian_splatting_ws.argtypes = [_vp] * 5 + [_i, _i, _i, _i, _i, _vp, ctypes.c_size_t, _ll, _vp, _i]
        L.xyz_splat_workspace_status.argtypes = [_vp, _vp, _vp]
        L.xyz_splat_last_timing.argtypes = [_vp]
        L.xyz_comm_unique_id.argtypes = [ctypes.c_char_p]
        L.xyz_comm_init_rank.argtypes = [ctypes.POINTER(_vp), ctypes.c_char_p, _i, _i]
        L.xyz_comm_init.argtypes = [ctypes.POINTER(_vp), _vp, _i, _i]
        L.xyz_comm_init_all.argtypes = [ctypes.POINTER(_vp), _i, _vp]
        L.xyz_comm_destroy.argtypes = [_vp]
        L.xyz_comm_rank.argtypes = [_vp]
        L.xyz_comm_world.argtypes = [_vp]
        L.xyz_allreduce_grads.argtypes = [_vp, _vp, _ll, _vp]
        L.xyz_allreduce_f64.argtypes = [_vp, _vp, _ll, _vp]
        L.xyz_adam_step_individual_sharded.argtypes = [_vp, _vp, _vp, _vp, _i, _vp, ctypes.c_float, ctypes.c_float,
                                                       ctypes.c_float, _i, _vp, _vp]
        L.xyz_peer_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(_vp), ctypes.c_char_p]
        L.xyz_adam_step_individual_peer.argtypes = [ctypes.POINTER(PeerGroupStruct), ctypes.POINTER(PeerSplatBuffersStruct),
                                                    _vp, _i, _vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, _i, _vp,
                                                    _vp]
        _lib = L
    return _lib


def _check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"{what} failed with code {code} "
                           f"({'XYZ_ERR' if code < 0 else 'cudaError_t'}); no fallback path exists")


def _dev(t: torch.Tensor, dtype: torch.dtype, what: str, numel: Optional[int] = None) -> int:
    """Device pointer of a contiguous CUDA tensor of `dtype` that lives on the CURRENT device (the library's scratch and
    launches go to the current device) and, if given, holds at least `numel` elements: a mismatch is a Python error here,
    not an out-of-bounds access in a kernel."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (the hot path has no CPU implementation)")
    if t.dtype != dtype:
        raise TypeError(f"{what}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{what}: tensor must be contiguous")
    if t.device.index != torch.cuda.current_device():
        raise ValueError(f"{what}: tensor is on {t.device} but the current device is cuda:{torch.cuda.current_device()} "
                         "(wrap the call in torch.cuda.device(...))")
    if numel is not None and t.numel() < numel:
        raise ValueError(f"{what}: needs at least {numel} elements, got {t.numel()}")
    return t.data_ptr()


def _stream(stream: Optional[torch.cuda.Stream]) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def version() -> str:
    return lib().xyz_b200_version().decode()


def launch_count() -> int:
    return int(lib().xyz_b200_launch_count())


def reset_launch_count() -> None:
    lib().xyz_b200_reset_launch_count()


def shutdown() -> None:
    _check(lib().xyz_b200_shutdown(), "xyz_b200_shutdown")


# ---- C1 ------------------------------------------------------------------------------------------
def lsq_grad(data: torch.Tensor, params: torch.Tensor, loss_sum: Optional[torch.Tensor] = None, flags: int = 0,
             stream=None) -> None:
    """data: (E, 3) float64 DataPoint{x1, x2, y}; params: (8,) float64 = value[4] + grad[4] (grad += )."""
    n = data.shape[0]
    _check(lib().xyz_lsq_grad_f64(_dev(data, torch.float64, "data", 3 * n), n, _dev(params, torch.float64, "params", 8),
                                  _dev(loss_sum, torch.float64, "loss_sum", 1) if loss_sum is not None else None,
                                  _stream(stream), flags), "xyz_lsq_grad_f64")


class PeerGroup:
    """This rank's mailbox + the opened mailboxes of the other ranks of one NVSwitch box (xyz_peer_* in
    include/xyz_b200.h).  `exchange(handle_bytes) -> list of every rank's handle` is the host channel (e.g. an
    all-gather over torch.distributed); the kernels themselves talk over NVLink peer memory only."""

    def __init__(self, rank: int, world: int, exchange):
        if not 1 <= world <= PEER_MAX_WORLD:
            raise ValueError("world must be 1..8")
        L = lib()
        self.rank, self.world, self.seq = rank, world, 0
        self._local = _vp()
        handle = ctypes.create_string_buffer(64)
        _check(L.xyz_peer_mailbox_create(ctypes.byref(self._local), handle), "xyz_peer_mailbox_create")
        handles = exchange(handle.raw)
        self.struct = PeerGroupStruct()
        self.struct.rank, self.struct.world = rank, world
        self._opened = []
        for r in range(world):
            if r == rank:
                self.struct.mailbox[r] = self._local.value
            else:
                ptr = _vp()
                _check(L.xyz_peer_mailbox_open(ctypes.create_string_buffer(handles[r], 64), ctypes.byref(ptr)),
                       "xyz_peer_mailbox_open")
                self._opened.append(ptr)
                self.struct.mailbox[r] = ptr.value

    def next_seq(self) -> int:
        self.seq += 1
        return self.seq

    def close(self) -> None:
        L = lib()
        for ptr in self._opened:
            L.xyz_peer_mailbox_close(ptr)
        self._opened = []
        if self._local:
            L.xyz_peer_mailbox_destroy(self._local)
            self._local = _vp()


def lsq_grad_allreduce(data: torch.Tensor, params: torch.Tensor, group: PeerGroup,
                       loss_sum: Optional[torch.Tensor] = None, flags: int = 0, stream=None) -> None:
    """lsq_grad on this rank's points, with the sum over all ranks of the 4 gradients (+ loss) done inside the
    kernel over NVLink peer memory: params.grad += global sum on every rank (bit-identical)."""
    n = data.shape[0]
    dp = _dev(data, torch.float64, "data", 3 * n) if n > 0 else None
    _check(lib().xyz_lsq_grad_f64_allreduce(dp, n, _dev(params, torch.float64, "params", 8),
                                            _dev(loss_sum, torch.float64, "loss_sum", 1) if loss_sum is not None else None,
                                            ctypes.byref(group.struct), group.next_seq(), _stream(stream), flags),
           "xyz_lsq_grad_f64_allreduce")


def lsq_sgd_update(params: torch.Tensor, lr: float, batch: int, stream=None) -> None:
    _check(lib().xyz_lsq_sgd_update_f64(_dev(params, torch.float64, "params", 8), lr, batch, _stream(stream)),
           "xyz_lsq_sgd_update_f64")


def lsq_select_batch(data: torch.Tensor, batch: torch.Tensor, seed: int, epoch: int, stream=None) -> None:
    _check(lib().xyz_lsq_select_batch(_dev(data, torch.float64, "data", 3 * data.shape[0]), data.shape[0],
                                      _dev(batch, torch.float64, "batch", 3 * batch.shape[0]), batch.shape[0], seed, epoch,
                                      _stream(stream)), "xyz_lsq_select_batch")


def lsq_sgd_step(data: torch.Tensor, params: torch.Tensor, batch: int, seed: int, epoch: int, lr: float,
                 loss_sum: Optional[torch.Tensor] = None, flags: int = 0, stream=None) -> None:
    """One SGD epoch in one launch: sample the batch (same hash as lsq_select_batch), params.grad = batch gradient,
    params.value -= lr * grad / batch."""
    _check(lib().xyz_lsq_sgd_step_f64(_dev(data, torch.float64, "data", 3 * data.shape[0]), data.shape[0],
                                      _dev(params, torch.float64, "params", 8), batch, seed, epoch, lr,
                                      _dev(loss_sum, torch.float64, "loss_sum", 1) if loss_sum is not None else None,
                                      _stream(stream), flags), "xyz_lsq_sgd_step_f64")


def lsq_sgd_run(data: torch.Tensor, params: torch.Tensor, batch: int, seed: int, epoch_begin: int, learning_rates,
                loss_sum: Optional[torch.Tensor] = None, flags: int = 0, stream=None) -> None:
    """len(learning_rates) SGD epochs (epoch_begin, epoch_begin + 1, ...) in ONE cooperative launch; the final state is
    bit-identical to calling lsq_sgd_step once per epoch."""
    lrs = [float(v) for v in learning_rates]
    arr = (ctypes.c_double * len(lrs))(*lrs)
    _check(lib().xyz_lsq_sgd_run_f64(_dev(data, torch.float64, "data", 3 * data.shape[0]), data.shape[0],
                                     _dev(params, torch.float64, "params", 8), batch, seed, epoch_begin, len(lrs), arr,
                                     _dev(loss_sum, torch.float64, "loss_sum", 1) if loss_sum is not None else None,
                                     _stream(stream), flags), "xyz_lsq_sgd_run_f64")


# ---- C2 ------------------------------------------------------------------------------------------
def accumulate(idx: Optional[torch.Tensor], val: torch.Tensor, grad: torch.Tensor, flags: int = 0, stream=None) -> None:
    """grad[idx[i]] += val[i]  (VariableRef::add_grad over K shared parameters)."""
    n, k = val.numel(), grad.numel()
    if idx is None:
        flags |= FLAG_IMPLICIT_IDS
    ip = _dev(idx, torch.int32, "idx", n) if idx is not None else None
    if val.dtype == torch.float32:
        code = lib().xyz_accumulate_f32(ip, _dev(val, torch.float32, "val"), n, _dev(grad, torch.float32, "grad"), k,
                                        _stream(stream), flags)
    else:
        code = lib().xyz_accumulate_f64(ip, _dev(val, torch.float64, "val"), n, _dev(grad, torch.float64, "grad"), k,
                                        _stream(stream), flags)
    _check(code, "xyz_accumulate")


def accumulate_allreduce(idx: Optional[torch.Tensor], val: torch.Tensor, grad: torch.Tensor, group: "PeerGroup",
                         flags: int = 0, stream=None) -> None:
    """accumulate on this rank's elements; grad[b] += the sum over ALL ranks, exchanged over NVLink peer memory by
    the finishing kernel (fp32, K <= 4096)."""
    n, k = val.numel(), grad.numel()
    if idx is None:
        flags |= FLAG_IMPLICIT_IDS
    ip = _dev(idx, torch.int32, "idx", n) if (idx is not None and n > 0) else None
    vp = _dev(val, torch.float32, "val") if n > 0 else None
    _check(lib().xyz_accumulate_f32_allreduce(ip, vp, n, _dev(grad, torch.float32, "grad"), k,
                                              ctypes.byref(group.struct), group.next_seq(), _stream(stream), flags),
           "xyz_accumulate_f32_allreduce")


# ---- C3 ------------------------------------------------------------------------------------------
def covproj_fwd_bwd(J, W, S, g, out, gJ, gW, gS, flags: int = 0, stream=None) -> None:
    """Per element: out = packed (J W) S (J W)^T and the adjoints of J (6), W (9), S (6) given g (3)."""
    n = J.shape[0]
    f = torch.float32
    _check(lib().xyz_covproj_fwd_bwd_f32(_dev(J, f, "J", 6 * n), _dev(W, f, "W", 9 * n), _dev(S, f, "S", 6 * n),
                                         _dev(g, f, "g", 3 * n), _dev(out, f, "out", 3 * n), _dev(gJ, f, "gJ", 6 * n),
                                         _dev(gW, f, "gW", 9 * n), _dev(gS, f, "gS", 6 * n),
                                         n, _stream(stream), flags), "xyz_covproj_fwd_bwd_f32")


def covproj_shared_w_fwd_bwd(J, W9, S, g, out, gJ, gW9, gS, flags: int = 0, stream=None, group: "Optional[PeerGroup]" = None) -> None:
    """Variant with ONE shared 3x3 W (9 floats): out / gJ / gS per element (overwritten), gW9 += the sum over the elements
    of the per-element adjoints of W (fixed-order reduction).  With `group` the elements are this rank's shard and
    gW9 += the sum over ALL ranks (exchanged inside the kernel over NVLink mailboxes)."""
    n = J.shape[0]
    f = torch.float32
    ptr = lambda t, nm, w: _dev(t, f, nm, w * n) if n > 0 else None  # noqa: E731
    args = [ptr(J, "J", 6), _dev(W9, f, "W9", 9), ptr(S, "S", 6), ptr(g, "g", 3), ptr(out, "out", 3), ptr(gJ, "gJ", 6),
            _dev(gW9, f, "gW9", 9), ptr(gS, "gS", 6), n]
    if W9.numel() != 9 or gW9.numel() != 9:
        raise ValueError("W9 and gW9 must hold 9 floats")
    if group is None:
        _check(lib().xyz_covproj_shared_w_fwd_bwd_f32(*args, _stream(stream), flags), "xyz_covproj_shared_w_fwd_bwd_f32")
    else:
        _check(lib().xyz_covproj_shared_w_fwd_bwd_f32_allreduce(*args, ctypes.byref(group.struct), group.next_seq(),
                                                                _stream(stream), flags),
               "xyz_covproj_shared_w_fwd_bwd_f32_allreduce")


# ---- C4 / C5 ---------------------------------------------------------------------------------------
def _splat_args(gaussians, gradients, target_image, output_image, total_loss, image_width, image_height, num_gaussians):
    f = torch.float32
    if image_width <= 0 or image_height <= 0 or num_gaussians < 0:
        raise ValueError("image_width / image_height must be positive, num_gaussians >= 0")
    npix = image_width * image_height
    gp = _dev(gaussians, f, "gaussians", GAUSSIAN_FLOATS * num_gaussians) if num_gaussians else None
    gg = _dev(gradients, f, "gradients", GAUSSIAN_FLOATS * num_gaussians) if num_gaussians else None
    return [gp, gg, _dev(target_image, f, "target_image", 3 * npix), _dev(output_image, f, "output_image", 3 * npix),
            _dev(total_loss, f, "total_loss", 1), image_width, image_height, num_gaussians]


def launch_gaussian_splatting(gaussians, gradients, target_image, output_image, total_loss, image_width: int,
                              image_height: int, num_gaussians: int, flags: int = 0, stream=None,
                              rows: Optional[tuple] = None) -> None:
    """Same argument order as the reference's launch_gaussian_splatting; tensors are (N, 9) / (P, 3) float32."""
    args = _splat_args(gaussians, gradients, target_image, output_image, total_loss, image_width, image_height,
                       num_gaussians)
    if rows is None:
        code = lib().xyz_launch_gaussian_splatting(*args, _stream(stream), flags)
    else:
        code = lib().xyz_launch_gaussian_splatting_rows(*args, int(rows[0]), int(rows[1]), _stream(stream), flags)
    _check(code, "xyz_launch_gaussian_splatting")


class SplatWorkspace:
    """Caller-owned workspace of xyz_launch_gaussian_splatting_ws: sized once for (image, N, row band, max_entries,
    flags); launches on it never allocate or synchronise and keep no library state, so several workspaces can be in
    flight on different streams, and a launch can be captured into a CUDA graph from the first call."""

    def __init__(self, image_width: int, image_height: int, num_gaussians: int, max_entries: int, flags: int = 0,
                 rows: Optional[tuple] = None, device=None):
        self.shape = (image_width, image_height, num_gaussians)
        self.rows = (0, image_height) if rows is None else (int(rows[0]), int(rows[1]))
        self.max_entries, self.flags = int(max_entries), int(flags)
        self.bytes = int(lib().xyz_splat_workspace_bytes(image_width, image_height, num_gaussians, self.rows[0], self.rows[1],
                                                         self.max_entries, self.flags))
        if self.bytes == 0:
            raise ValueError("xyz_splat_workspace_bytes: unsupported shape / flags (radix binning, more than 57344 tiles "
                             "in the row band, or invalid sizes)")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.buffer = torch.zeros(self.bytes + 256, dtype=torch.uint8, device=dev)
        self.ptr = (self.buffer.data_ptr() + 255) // 256 * 256  # header zeroed by torch.zeros

    def launch(self, gaussians, gradients, target_image, output_image, total_loss, flags: Optional[int] = None,
               stream=None) -> None:
        w, h, n = self.shape
        args = _splat_args(gaussians, gradients, target_image, output_image, total_loss, w, h, n)
        _check(lib().xyz_launch_gaussian_splatting_ws(*args, self.rows[0], self.rows[1], self.ptr, self.bytes,
                                                      self.max_entries, _stream(stream),
                                                      self.flags if flags is None else flags),
               "xyz_launch_gaussian_splatting_ws")

    def status(self, stream=None) -> dict:
        buf = (ctypes.c_longlong * 4)()
        _check(lib().xyz_splat_workspace_status(self.ptr, _stream(stream), buf), "xyz_splat_workspace_status")
        return {"entries": buf[0], "overflows": buf[1], "max_entries": buf[2], "overflowed": bool(buf[3])}


def splat_last_stats() -> dict:
    buf = (ctypes.c_longlong * 4)()
    _check(lib().xyz_splat_last_stats(buf), "xyz_splat_last_stats")
    return {"entries": buf[0], "tiles": buf[1], "longest_tile_list": buf[2], "pairs_per_pass": buf[3]}


def splat_last_backward_stats() -> dict:
    """What the backward pass of this thread's most recent launch worked on (reads its work records back)."""
    buf = (ctypes.c_longlong * 3)()
    _check(lib().xyz_splat_last_backward_stats(buf), "xyz_splat_last_backward_stats")
    return {"items": buf[0], "pairs": buf[1], "ctas": buf[2]}


def splat_last_timing() -> dict:
    """Stage times (us) of this thread's most recent launch made with FLAG_TIMING."""
    buf = (ctypes.c_float * 6)()
    _check(lib().xyz_splat_last_timing(buf), "xyz_splat_last_timing")
    return dict(zip(("preprocess_hist_us", "scans_us", "scatter_us", "forward_loss_us", "backward_us", "total_us"),
                    (float(v) for v in buf)))


def splat_debug_binning(num_gaussians: int, num_tiles: int, entries: int):
    import numpy as np
    rects = np.zeros((max(num_gaussians, 1), 4), np.int32)
    ranges = np.zeros((num_tiles, 2), np.int32)
    ids = np.zeros(max(entries, 1), np.int32)
    recs = np.zeros((max(num_gaussians, 1), 12), np.float32)
    _check(lib().xyz_splat_debug_binning(rects.ctypes.data, ranges.ctypes.data, ids.ctypes.data, recs.ctypes.data),
           "xyz_splat_debug_binning")
    return rects[:num_gaussians], ranges, ids[:entries], recs[:num_gaussians]


def zero_gradients(gradients: torch.Tensor, stream=None) -> None:
    n = gradients.shape[0]
    _check(lib().xyz_zero_gradients(_dev(gradients, torch.float32, "gradients", GAUSSIAN_FLOATS * n), n, _stream(stream)),
           "xyz_zero_gradients")


def adam_step_individual(params, grads, adam, lr_center, lr_scale, lr_rotation, lr_color, lr_opacity, beta1=0.9,
                         beta2=0.999, epsilon=1e-8, iteration=1, stream=None, zero_grads=False) -> None:
    """zero_grads=True: the gradients are cleared in the same pass (no zero_gradients call next iteration)."""
    f = torch.float32
    lr = (ctypes.c_float * 5)(lr_center, lr_scale, lr_rotation, lr_color, lr_opacity)
    fn = lib().xyz_adam_step_individual_zero_grads if zero_grads else lib().xyz_adam_step_individual
    n = params.shape[0]
    _check(fn(_dev(params, f, "params", GAUSSIAN_FLOATS * n), _dev(grads, f, "grads", GAUSSIAN_FLOATS * n),
              _dev(adam, f, "adam", ADAM_FLOATS * n), n, lr, beta1, beta2, epsilon, iteration, _stream(stream)),
           "xyz_adam_step_individual")


def adam_step(params, grads, adam, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8, iteration=1, stream=None) -> None:
    f = torch.float32
    n = params.shape[0]
    _check(lib().xyz_adam_step(_dev(params, f, "params", GAUSSIAN_FLOATS * n), _dev(grads, f, "grads", GAUSSIAN_FLOATS * n),
                               _dev(adam, f, "adam", ADAM_FLOATS * n), n, learning_rate, beta1, beta2, epsilon, iteration,
                               _stream(stream)),
           "xyz_adam_step")


# ---- multi-GPU exchange of the splat gradients ---------------------------------------------------------------------
class Comm:
    """xyz_comm (include/xyz_b200.h): the library's own NCCL communicator, one rank per process.
    `exchange(id_bytes_or_None) -> id_bytes` is the host channel that carries rank 0's 128-byte NCCL id to every rank
    (e.g. a broadcast over torch.distributed); NCCL itself is bound by the library at run time."""

    def __init__(self, rank: int, world: int, exchange):
        L = lib()
        self.rank, self.world = rank, world
        ident = ctypes.create_string_buffer(128)
        if rank == 0:
            _check(L.xyz_comm_unique_id(ident), "xyz_comm_unique_id")
        raw = exchange(ident.raw if rank == 0 else None)
        self._c = _vp()
        _check(L.xyz_comm_init_rank(ctypes.byref(self._c), ctypes.create_string_buffer(raw, 128), rank, world),
               "xyz_comm_init_rank")

    def allreduce_grads(self, grads: torch.Tensor, stream=None) -> None:
        """grads = sum over ranks, fp32, in place, on the stream that produced them."""
        _check(lib().xyz_allreduce_grads(self._c, _dev(grads, torch.float32, "grads"), grads.numel(), _stream(stream)),
               "xyz_allreduce_grads")

    def allreduce_f64(self, values: torch.Tensor, stream=None) -> None:
        _check(lib().xyz_allreduce_f64(self._c, _dev(values, torch.float64, "values"), values.numel(), _stream(stream)),
               "xyz_allreduce_f64")

    def adam_step_individual_sharded(self, params, grads, adam, lr_center, lr_scale, lr_rotation, lr_color, lr_opacity,
                                     beta1=0.9, beta2=0.999, epsilon=1e-8, iteration=1, total_loss=None, stream=None) -> None:
        """reduce-scatter(grads) -> Adam on this rank's Gaussian range -> zero grads -> all-gather(params)."""
        f = torch.float32
        n = params.shape[0]
        lr = (ctypes.c_float * 5)(lr_center, lr_scale, lr_rotation, lr_color, lr_opacity)
        _check(lib().xyz_adam_step_individual_sharded(
            self._c, _dev(params, f, "params", GAUSSIAN_FLOATS * n), _dev(grads, f, "grads", GAUSSIAN_FLOATS * n),
            _dev(adam, f, "adam", ADAM_FLOATS * n), n, lr, beta1, beta2, epsilon, iteration,
            _dev(total_loss, f, "total_loss", 1) if total_loss is not None else None, _stream(stream)),
            "xyz_adam_step_individual_sharded")

    def destroy(self) -> None:
        if self._c:
            lib().xyz_comm_destroy(self._c)
            self._c = _vp()


class _RawCudaArray:
    """A library-allocated device buffer seen through __cuda_array_interface__ (so torch can alias it without a copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


def peer_alloc_tensor(shape, dtype=torch.float32):
    """(tensor, ipc_handle, raw_ptr): a zero-filled device tensor in memory allocated by xyz_peer_alloc on the current
    device, i.e. memory other ranks can map with their xyz_peer_mailbox_open."""
    n = 1
    for d in shape:
        n *= int(d)
    item = torch.empty((), dtype=dtype).element_size()
    ptr = _vp()
    handle = ctypes.create_string_buffer(64)
    _check(lib().xyz_peer_alloc(max(1, n) * item, ctypes.byref(ptr), handle), "xyz_peer_alloc")
    typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4"}[dtype]
    t = torch.as_tensor(_RawCudaArray(ptr.value, shape, typestr), device=torch.device("cuda", torch.cuda.current_device()))
    return t, handle.raw, ptr.value


class PeerSplat:
    """The buffers of xyz_adam_step_individual_peer for one rank: params and grads (N, 9) live in library-allocated,
    IPC-exported memory and are mapped into every other rank; adam (N, 18) is local.  `group` is a PeerGroup (mailboxes),
    `exchange(obj) -> list of every rank's obj` the host channel for the IPC handles."""

    def __init__(self, group: "PeerGroup", num_gaussians: int, exchange):
        L = lib()
        self.group, self.n = group, num_gaussians
        self.params, hp, pp = peer_alloc_tensor((num_gaussians, GAUSSIAN_FLOATS))
        self.grads, hg, pg = peer_alloc_tensor((num_gaussians, GAUSSIAN_FLOATS))
        self._own = (pp, pg)
        self.adam = torch.zeros((num_gaussians, ADAM_FLOATS), dtype=torch.float32,
                                device=torch.device("cuda", torch.cuda.current_device()))
        handles = exchange((hp, hg))
        self.struct = PeerSplatBuffersStruct()
        self._opened = []
        for r in range(group.world):
            if r == group.rank:
                self.struct.params[r], self.struct.grads[r] = pp, pg
            else:
                for which, field in ((0, self.struct.params), (1, self.struct.grads)):
                    ptr = _vp()
                    _check(L.xyz_peer_mailbox_open(ctypes.create_string_buffer(handles[r][which], 64), ctypes.byref(ptr)),
                           "xyz_peer_mailbox_open")
                    self._opened.append(ptr)
                    field[r] = ptr.value

    def adam_step(self, lr_center, lr_scale, lr_rotation, lr_color, lr_opacity, beta1=0.9, beta2=0.999, epsilon=1e-8,
                  iteration=1, total_loss=None, stream=None) -> None:
        """Fused reduce-scatter + Adam + all-gather + zero-grad (+ loss all-reduce) over NVLink peer memory: one launch.
        iteration=0: the step number is counted on the device (for replayed CUDA graphs)."""
        lr = (ctypes.c_float * 5)(lr_center, lr_scale, lr_rotation, lr_color, lr_opacity)
        _check(lib().xyz_adam_step_individual_peer(
            ctypes.byref(self.group.struct), ctypes.byref(self.struct),
            _dev(self.adam, torch.float32, "adam", ADAM_FLOATS * self.n), self.n,
            lr, beta1, beta2, epsilon, iteration,
            _dev(total_loss, torch.float32, "total_loss", 1) if total_loss is not None else None, _stream(stream)),
            "xyz_adam_step_individual_peer")

    def close(self) -> None:
        L = lib()
        torch.cuda.synchronize()
        for ptr in self._opened:
            L.xyz_peer_mailbox_close(ptr)
        self._opened = []
        self.params = self.grads = None
        for ptr in self._own:
            L.xyz_peer_mailbox_destroy(ptr)
        self._own = ()
