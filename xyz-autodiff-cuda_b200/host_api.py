"""Host-buffer entry points: the call a user of the reference makes when the data lives in host memory.

The reference's drivers keep their inputs on the host (std::vector<DataPoint>, GaussianCollection::host_params,
ImageData) and copy them to the device around the kernel launches
(examples/optimization/linear_regression_sgd.cu:164-168, examples/mini-gaussian-splatting/
gaussian_parameters.cu:68-116).  These helpers do the same around the C-ABI kernels, with pinned
staging and a chunked three-stage pipeline (H2D copy / kernel / D2H copy on separate streams) so the
PCIe transfers overlap the kernels.  They are what bench.py times for the `e2e` figure.
"""
from __future__ import annotations

import torch

import xyz_autodiff_cuda_b200 as x


class CovprojHostPipeline:
    """out, gJ, gW, gS (host, pinned) = covproj_fwd_bwd(J, W, S, g (host, pinned)), chunk by chunk."""

    WIDTHS_IN = (6, 9, 6, 3)
    WIDTHS_OUT = (3, 6, 9, 6)

    def __init__(self, device: torch.device, chunk_elems: int = 1 << 20, depth: int = 4):
        self.device = device
        self.chunk = chunk_elems
        self.depth = depth
        self.s_in = torch.cuda.Stream(device)
        self.s_k = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        f = torch.float32
        self.d_in = [[torch.empty((chunk_elems, w), dtype=f, device=device) for w in self.WIDTHS_IN] for _ in range(depth)]
        self.d_out = [[torch.empty((chunk_elems, w), dtype=f, device=device) for w in self.WIDTHS_OUT] for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]      # H2D of slot done
        self.ev_k = [torch.cuda.Event() for _ in range(depth)]       # kernel of slot done
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]     # D2H of slot done

    def run(self, host_in, host_out) -> tuple:
        """host_in: 4 pinned tensors (n, 6|9|6|3); host_out: 4 pinned tensors (n, 3|6|9|6).
        Returns (h2d_bytes, d2h_bytes).  Leaves the work queued; call torch.cuda.synchronize() to finish."""
        n = host_in[0].shape[0]
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_k, self.s_out):
            s.wait_stream(cur)
        h2d = d2h = 0
        for c, e0 in enumerate(range(0, n, self.chunk)):
            e1 = min(n, e0 + self.chunk)
            m = e1 - e0
            slot = c % self.depth
            with torch.cuda.stream(self.s_in):
                if c >= self.depth:
                    self.s_in.wait_event(self.ev_k[slot])    # kernel that read this slot's inputs is done
                for d, h in zip(self.d_in[slot], host_in):
                    d[:m].copy_(h[e0:e1], non_blocking=True)
                    h2d += m * d.shape[1] * 4
                self.ev_in[slot].record(self.s_in)
            with torch.cuda.stream(self.s_k):
                self.s_k.wait_event(self.ev_in[slot])
                if c >= self.depth:
                    self.s_k.wait_event(self.ev_out[slot])   # D2H that read this slot's outputs is done
                x.covproj_fwd_bwd(*[d[:m] for d in self.d_in[slot]], *[d[:m] for d in self.d_out[slot]], stream=self.s_k)
                self.ev_k[slot].record(self.s_k)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_k[slot])
                for d, h in zip(self.d_out[slot], host_out):
                    h[e0:e1].copy_(d[:m], non_blocking=True)
                    d2h += m * d.shape[1] * 4
                self.ev_out[slot].record(self.s_out)
        for s in (self.s_in, self.s_k, self.s_out):
            cur.wait_stream(s)
        return h2d, d2h


def lsq_grad_host(data_host: torch.Tensor, values, device, flags: int = 0):
    """One gradient evaluation from a pinned host DataPoint array: H2D, kernel, D2H of {grad[4], loss}."""
    d = data_host.to(device, non_blocking=True)
    prm = torch.zeros(8, dtype=torch.float64, device=device)
    prm[:4] = torch.as_tensor(values, dtype=torch.float64).to(device)
    loss = torch.zeros(1, dtype=torch.float64, device=device)
    x.lsq_grad(d, prm, loss, flags)
    return prm[4:].cpu(), loss.cpu()


def accumulate_host(idx_host: torch.Tensor, val_host: torch.Tensor, k: int, device, flags: int = 0):
    idx = idx_host.to(device, non_blocking=True)
    val = val_host.to(device, non_blocking=True)
    grad = torch.zeros(k, dtype=val.dtype, device=device)
    x.accumulate(idx, val, grad, flags)
    return grad.cpu()


def splat_iteration_host(params_host: torch.Tensor, target_dev: torch.Tensor, output_dev: torch.Tensor, width: int,
                         height: int, device, flags: int = 0):
    """One iteration of the reference training loop's device work (gaussian_splatting_training.cu:127-151) with
    the parameters coming from pinned host memory: upload params, zero grads, launch, read back loss and
    gradients.  Returns (loss (1,) host, grads (N, 9) host)."""
    n = params_host.shape[0]
    params = params_host.to(device, non_blocking=True)
    grads = torch.empty((n, 9), dtype=torch.float32, device=device)
    x.zero_gradients(grads)
    loss = torch.zeros(1, dtype=torch.float32, device=device)
    x.launch_gaussian_splatting(params, grads, target_dev, output_dev, loss, width, height, n, flags)
    return loss.cpu(), grads.cpu()


class SplatHostIteration:
    """The device work of one iteration of the reference's training loop (gaussian_splatting_training.cu:127-151:
    zero gradients, reset the loss, launch_gaussian_splatting, read the loss back) for a caller whose Gaussians live in
    HOST memory: the parameters come from a pinned host array every iteration, the loss and the gradients go back to
    pinned host arrays.  Everything is queued on one stream through a caller-owned workspace
    (xyz_launch_gaussian_splatting_ws: no allocation, no synchronisation inside); `run` returns after ONE stream
    synchronisation.  This is what bench.py times as the splat `e2e`."""

    def __init__(self, num_gaussians: int, width: int, height: int, target_dev: torch.Tensor, max_entries: int,
                 device, flags: int = 0):
        self.n, self.w, self.h, self.device = num_gaussians, width, height, device
        self.target = target_dev
        f = torch.float32
        self.params = torch.empty((num_gaussians, 9), dtype=f, device=device)
        self.grads = torch.empty((num_gaussians, 9), dtype=f, device=device)
        self.output = torch.empty((width * height, 3), dtype=f, device=device)
        self.loss = torch.zeros(1, dtype=f, device=device)
        self.grads_host = torch.empty((num_gaussians, 9), dtype=f).pin_memory()
        self.loss_host = torch.zeros(1, dtype=f).pin_memory()
        self.ws = x.SplatWorkspace(width, height, num_gaussians, max_entries, flags, device=device)
        self.h2d_bytes = num_gaussians * 36
        self.d2h_bytes = num_gaussians * 36 + 4

    def run(self, params_host: torch.Tensor, stream=None):
        """params_host: pinned (N, 9) float32.  Returns (loss_host (1,), grads_host (N, 9)) -- pinned, reused per call."""
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.stream(st):
            self.params.copy_(params_host, non_blocking=True)
            x.zero_gradients(self.grads, stream=st)
            self.loss.zero_()
            self.ws.launch(self.params, self.grads, self.target, self.output, self.loss, stream=st)
            self.loss_host.copy_(self.loss, non_blocking=True)
            self.grads_host.copy_(self.grads, non_blocking=True)
        st.synchronize()
        return self.loss_host, self.grads_host
