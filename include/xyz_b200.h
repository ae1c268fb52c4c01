/* include/xyz_b200.h -- C ABI of libxyz_b200.so, the B200 (sm_100a) drop-in for the data-parallel
 * hot path of xyz-autodiff-cuda.
 *
 * The reference has NO extern "C" surface: its only host-callable entry
 * point on this path is the C++ function launch_gaussian_splatting(...)
 * (reference: examples/mini-gaussian-splatting/gaussian_splatting_kernel.cuh:38-47) plus the
 * GaussianCollection::{zero_gradients_gpu, adam_step_gpu, adam_step_gpu_individual} methods
 * (examples/mini-gaussian-splatting/gaussian_parameters.h:78-91) and file-local kernels
 * launched from main() (examples/optimization/linear_regression_sgd.cu:86-134).  Each entry
 * point below cites the reference interface it replaces.  A C++ shim with the reference's
 * exact launch_gaussian_splatting signature lives in include/xyz_b200_compat.hpp.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what
 *     the reference uses);
 *   - the return value is a cudaError_t as int (0 = success) or a negative XYZ_ERR_* code;
 *     nothing throws; nothing prints;
 *   - accumulation semantics are the reference's: gradient and loss outputs are ADDED to, the
 *     caller zeroes them (gaussian_splatting_training.cu:131-135, linear_regression_sgd.cu:201);
 *   - calls are asynchronous on `stream` except where stated.
 */
#ifndef XYZ_B200_H_
#define XYZ_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define XYZ_API __attribute__((visibility("default")))
#else
#define XYZ_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- data layouts (bit-identical to the reference's structs) --------------------------- */

/* gaussian_parameters.h:12-18 (GaussianParams) and :35-41 (GaussianGrads): 9 floats, 36 B. */
typedef struct xyz_gaussian_params {
    float center[2];
    float scale[2];   /* log-scale: sigma = exp(scale) */
    float rotation[1];
    float color[3];
    float opacity[1]; /* logit: alpha = sigmoid(opacity) */
} xyz_gaussian_params;
typedef xyz_gaussian_params xyz_gaussian_grads;

/* gaussian_parameters.h:21-32 (AdamState): 18 floats, 72 B. */
typedef struct xyz_adam_state {
    float m_center[2], v_center[2];
    float m_scale[2], v_scale[2];
    float m_rotation[1], v_rotation[1];
    float m_color[3], v_color[3];
    float m_opacity[1], v_opacity[1];
} xyz_adam_state;

/* linear_regression_sgd.cu:31-33 (DataPoint) and :36-39 (Parameters). */
typedef struct xyz_data_point {
    double x1, x2, y;
} xyz_data_point;
typedef struct xyz_lsq_parameters {
    double value[4]; /* a, b, c, d */
    double grad[4];
} xyz_lsq_parameters;

/* ---- flags ------------------------------------------------------------------------------ */
enum {
    XYZ_FLAG_DETERMINISTIC = 1, /* fixed-order reductions: results are bit-identical run to run */
    XYZ_FLAG_PRECISE_MATH  = 2, /* splat: IEEE expf/sinf/cosf/div (the reference's test builds);
                                   default is the training app's fast-math/FTZ flavour
                                   (examples/mini-gaussian-splatting/CMakeLists.txt:22-31) */
    XYZ_FLAG_RESIDUAL_ONLY = 4, /* lsq: the SAME graph (test_linear_regression_gradient.cu:52-71) differentiated at
                                   the residual y_pred - y instead of at its square (root y_diff.run()).  NOT the shipped
                                   example's graph: see XYZ_FLAG_LSQ_SHIPPED_GRAPH */
    XYZ_FLAG_NO_CULL       = 8, /* splat: evaluate every (pixel, Gaussian) pair like the reference
                                   kernel does; default skips pairs whose weight is exactly 0 */
    XYZ_FLAG_IMPLICIT_IDS  = 16, /* accumulate: idx == NULL means id = i mod K */
    XYZ_FLAG_TAIL_CULL     = 32, /* splat, opt-in, NOT result-preserving: also skip pairs with d2 > 56, i.e. with a
                                    Gaussian weight below exp(-28) = 6.9e-13 (the default only skips weights that
                                    are exactly 0.0f).  Every pixel changes by at most
                                    N * 6.9e-13 * max|sigmoid(opacity) * color|; ~3x fewer pairs are evaluated. */
    XYZ_FLAG_RADIX_BINNING = 64, /* splat: build the tile lists with (tile, Gaussian) keys + a stable radix sort -- the
                                    path row bands of more than 57 344 tiles (14.7 Mpixel) take anyway -- instead of the default stable
                                    counting sort by tile.  Same lists bit for bit; for tests and comparisons. */
    XYZ_FLAG_LSQ_SHIPPED_GRAPH = 256, /* lsq: the graph parallel_gradient_computation_kernel really builds
                                   (linear_regression_sgd.cu:103-122): combined_terms adds the UN-squared
                                   x1_term = a - x1 (x1_term2 is dead code there) and loss.run() differentiates the
                                   residual:  r = (a - x1) + b (c - x2)^2 + d - y,  dr/d(a, b, c, d) = (1, (c - x2)^2,
                                   2 b (c - x2), 1).  Reproduces the shipped example's SGD trajectory. */
    XYZ_FLAG_TIMING        = 512, /* splat: record CUDA events between the stages of this launch; read them with
                                    xyz_splat_last_timing.  A measuring aid (the events cost a little themselves). */
    XYZ_FLAG_ASYNC         = 128, /* splat, opt-in: no host synchronisation inside the launch (and so capturable in a
                                    CUDA graph once the scratch has its size).  Takes effect from the second launch of a
                                    scene shape (N, width, height, rows) on a host thread, with the counting-sort
                                    binning and without XYZ_FLAG_DETERMINISTIC; otherwise the launch is an ordinary one.
                                    The entry-sized buffers and the backward grid are sized for 1.5 x the most recent
                                    list length the host has seen.  If a launch turns out longer, it renders EMPTY
                                    lists (image 0, loss = sum |target|, no gradients; memory-safe) and the next
                                    xyz_launch_gaussian_splatting* / xyz_splat_last_stats call on that thread returns
                                    XYZ_ERR_WORKSPACE once, without launching anything, and the buffers are sized afresh
                                    from the reported length: repeat the iteration.  Results of a launch that fits are
                                    the same as without the flag. */
    XYZ_FLAG_BWD_ALL_PAIRS = 1024, /* splat: the backward pass visits every pair of the tile lists, like the forward
                                    pass.  By default it leaves out the (tile, Gaussian) entries with d2 > 48 on all of
                                    the tile: every gradient term of such a pair carries the factor
                                    exp(-d2 / 2) < exp(-24) = 3.8e-11; what is left out of a gradient sum is below 1e-9
                                    of the sum of its terms' magnitudes (1/60 of an fp32 epsilon: at most the last bit of
                                    the fp32 sum moves; the bar for atomically accumulated sums is 1e-4), and 65 % of the
                                    list entries drop out at 100 K Gaussians x 1024^2.
                                    Image and loss are not affected (the forward pass always renders the full lists).
                                    XYZ_FLAG_NO_CULL implies this flag. */
};

enum {
    XYZ_ERR_INVALID_ARGUMENT = -1,
    XYZ_ERR_WORKSPACE        = -2,
    XYZ_ERR_NOT_INITIALISED  = -3,
    XYZ_ERR_COMM             = -4  /* NCCL is missing (dlopen of libnccl.so.2 failed) or returned an error */
};

/* ---- library ----------------------------------------------------------------------------- */
XYZ_API const char* xyz_b200_version(void);
/* Frees the library-owned scratch buffers of the current device (splat tile lists,
 * reduction partials).  Synchronises the device. */
XYZ_API int xyz_b200_shutdown(void);
/* Number of kernels the library has launched since load / since the last reset (host counter). */
XYZ_API uint64_t xyz_b200_launch_count(void);
XYZ_API void xyz_b200_reset_launch_count(void);

/* ---- C1: batched least-squares forward+reverse (fp64) --------------------------------------
 * Replaces compute_gradient_kernel<Analytical>
 * (examples/optimization/tests/test_linear_regression_gradient.cu:33-79) and, with
 * XYZ_FLAG_LSQ_SHIPPED_GRAPH, parallel_gradient_computation_kernel
 * (examples/optimization/linear_regression_sgd.cu:86-123).  For every data point builds
 * r = (a-x1)^2 + b(c-x2)^2 + d - y, loss = r^2, and ADDS d(loss)/d(a,b,c,d) into params->grad
 * (XYZ_FLAG_RESIDUAL_ONLY: d(r) instead; XYZ_FLAG_LSQ_SHIPPED_GRAPH: the shipped example's r, see the flag).
 * loss_sum (optional, may be NULL) receives += sum of per-point root values.            */
XYZ_API int xyz_lsq_grad_f64(const xyz_data_point* data, long long n_points, xyz_lsq_parameters* params,
                     double* loss_sum, void* stream, int flags);
/* update_parameters_kernel (linear_regression_sgd.cu:126-134): value -= lr * grad / batch. */
XYZ_API int xyz_lsq_sgd_update_f64(xyz_lsq_parameters* params, double learning_rate, long long batch_size,
                           void* stream);
/* select_batch_kernel (linear_regression_sgd.cu:68-81) with a counter-based RNG instead of a
 * per-thread curand_init: batch[i] = data[hash(seed, epoch, i) mod n_total].               */
XYZ_API int xyz_lsq_select_batch(const xyz_data_point* data, long long n_total, xyz_data_point* batch,
                         long long batch_size, uint64_t seed, uint64_t epoch, void* stream);

/* One whole SGD epoch of the reference driver (linear_regression_sgd.cu:185-209: select_batch_kernel, cudaMemset of
 * the gradients, parallel_gradient_computation_kernel, update_parameters_kernel) in ONE launch: sample i of the
 * batch is data[hash(seed, epoch, i) mod n_total] (the sampling of xyz_lsq_select_batch), params->grad is
 * OVERWRITTEN with the batch gradient, then value -= learning_rate * grad / batch_size.  loss_sum (optional)
 * receives += the batch loss.  Deterministic.  Flags: XYZ_FLAG_RESIDUAL_ONLY, XYZ_FLAG_LSQ_SHIPPED_GRAPH.                        */
XYZ_API int xyz_lsq_sgd_step_f64(const xyz_data_point* data, long long n_total, xyz_lsq_parameters* params,
                         long long batch_size, uint64_t seed, uint64_t epoch, double learning_rate,
                         double* loss_sum, void* stream, int flags);

/* n_epochs epochs of xyz_lsq_sgd_step_f64 (epochs epoch_begin .. epoch_begin + n_epochs - 1, learning rate
 * learning_rates_host[e] for the e-th of them; the array is read before the call returns) in ONE cooperative launch:
 * parameters stay in registers, one grid barrier per epoch.  The final params (value and grad) and *loss_sum are
 * bit-identical to n_epochs single-epoch calls.  Falls back to one launch per epoch when the grid cannot be
 * co-resident.  The driver loop of linear_regression_sgd.cu:185-229 between two progress prints.            */
XYZ_API int xyz_lsq_sgd_run_f64(const xyz_data_point* data, long long n_total, xyz_lsq_parameters* params,
                        long long batch_size, uint64_t seed, uint64_t epoch_begin, int n_epochs,
                        const double* learning_rates_host, double* loss_sum, void* stream, int flags);

/* ---- fused reduce + all-reduce over NVLink peer memory (one process per GPU, one box) -----------------------
 * The reference is single-GPU.  When the elements of C1 are sharded over the GPUs of an NVSwitch box the only
 * exchange is the sum of the 4 shared-parameter gradients (+ loss); xyz_lsq_grad_f64_allreduce does it inside the
 * gradient kernel: its last CTA stores the partial row into every rank's MAILBOX (device memory exported with CUDA
 * IPC), publishes a sequence number with st.release.sys, waits for all ranks and adds the rows in rank order --
 * bit-identical on every rank, no NCCL call, no second launch.
 *   setup (once):  xyz_peer_mailbox_create -> exchange the 64-byte handles by any host channel ->
 *                  xyz_peer_mailbox_open for every other rank -> fill xyz_peer_group (mailbox[rank] = own pointer)
 *   per call:      seq must be 1, 2, 3, ... identically on every rank (two calls may be in flight).            */
#define XYZ_PEER_MAX_WORLD 8
#define XYZ_PEER_SLOT_DOUBLES 8
#define XYZ_PEER_VEC_FLOATS 4096
typedef struct xyz_peer_group {
    void* mailbox[XYZ_PEER_MAX_WORLD]; /* device pointers, index = rank; entries >= world are ignored */
    int rank;
    int world;
} xyz_peer_group;
XYZ_API size_t xyz_peer_mailbox_bytes(void);
XYZ_API int xyz_peer_mailbox_create(void** local_ptr, unsigned char ipc_handle_out[64]);
XYZ_API int xyz_peer_mailbox_open(const unsigned char ipc_handle[64], void** peer_ptr);
XYZ_API int xyz_peer_mailbox_close(void* peer_ptr);
XYZ_API int xyz_peer_mailbox_destroy(void* local_ptr);
/* xyz_lsq_grad_f64 on this rank's points + the all-reduce: params->grad += sum over ALL ranks, *loss_sum likewise. */
XYZ_API int xyz_lsq_grad_f64_allreduce(const xyz_data_point* data, long long n_points, xyz_lsq_parameters* params,
                               double* loss_sum, const xyz_peer_group* group, unsigned long long seq,
                               void* stream, int flags);

/* xyz_accumulate_f32 on this rank's elements + the all-reduce of the K sums (K <= XYZ_PEER_VEC_FLOATS), declared
 * below next to xyz_accumulate_f32. */

/* ---- C2: accumulation of per-element gradients into K shared parameters (fp32) ---------------
 * Replaces the VariableRef::add_grad pattern (include/xyz_autodiff/variable.cuh:48-50) as
 * exercised by tests/test_parallel_gradient_accumulation.cu:25-49: grad[idx[i]] += val[i].   */
/* Out-of-range ids are ignored.  idx == NULL with XYZ_FLAG_IMPLICIT_IDS means id = i mod k.  XYZ_FLAG_DETERMINISTIC
 * gives bit-identical results run to run (K <= 16384 on 16-byte aligned arrays; larger K has no fixed-order path and
 * returns XYZ_ERR_INVALID_ARGUMENT).
 * Cost regimes on B200 for 2^24 elements (any id distribution unless noted): K <= 1810 one pass, 31 us; K <= 3626 one
 * pass, 57 us; K <= 16384 one pass per 1810 bins (K = 8192: 163 us); beyond that warp-aggregated global atomics
 * (uniform ids 160 us; heavily skewed ids up to ~2 ms).  16-byte aligned arrays (or slices of aligned arrays with the
 * same offset in idx and val) take the vector paths.                                                               */
XYZ_API int xyz_accumulate_f32(const int32_t* idx, const float* val, long long n, float* grad, int k,
                       void* stream, int flags);
/* Multi-GPU: the elements are sharded over the ranks of an xyz_peer_group, grad[b] += sum over ALL ranks.  The CTAs
 * write partial rows, and ONE finishing kernel adds them, stores the row into every rank's mailbox over NVLink,
 * waits for the other ranks and adds the rows in rank order (bit-identical on every rank; no NCCL call). */
XYZ_API int xyz_accumulate_f32_allreduce(const int32_t* idx, const float* val, long long n, float* grad, int k,
                                 const xyz_peer_group* group, unsigned long long seq, void* stream, int flags);
/* fp64 flavour used by the reference's own accumulation tests (3 addresses, double). */
XYZ_API int xyz_accumulate_f64(const int32_t* idx, const double* val, long long n, double* grad, int k,
                       void* stream, int flags);

/* ---- C3: batched covariance projection S' = (J W) S (J W)^T, forward + reverse (fp32) ----------
 * The composition of op::matmul<2,3,3>, op::matmul<2,3,3>, op::matmul<2,3,2>
 * (include/xyz_autodiff/operations/binary/matmul_logic.cuh:13-81) over a packed symmetric
 * 3x3 (include/xyz_autodiff/symmetric_matrix_view.cuh:24-29) as one kernel.
 * Per element e (all row-major, contiguous per array):
 *   J[e]: 2x3 (6)   W[e]: 3x3 (9)   S[e]: packed upper-triangular 3x3 (6: 00 01 02 11 12 22)
 *   g[e]: upstream adjoint of the packed 2x2 output (3: 00 01 11)
 * writes out[e] (3), gJ[e] (6), gW[e] (9), gS[e] (6).  Outputs are OVERWRITTEN (per-element
 * gradients have a single writer, nothing to accumulate into).                               */
XYZ_API int xyz_covproj_fwd_bwd_f32(const float* J, const float* W, const float* S, const float* g,
                            float* out, float* gJ, float* gW, float* gS, long long n,
                            void* stream, int flags);

/* Variant with ONE shared W (BASELINE configs[2] "variant B"): W9 is a single 3x3 parameter used by
 * every element, gW9 its 9 gradient accumulators.  out / gJ / gS are per element and OVERWRITTEN as above; the
 * per-element adjoints of W are summed and ADDED into gW9 (the VariableRef::add_grad accumulation of
 * include/xyz_autodiff/variable.cuh:48-50, done with a fixed-order reduction instead of 9 atomics per element:
 * bit-identical run to run).  120 algorithmic bytes per element.                                              */
XYZ_API int xyz_covproj_shared_w_fwd_bwd_f32(const float* J, const float* W9, const float* S, const float* g,
                                     float* out, float* gJ, float* gW9, float* gS, long long n,
                                     void* stream, int flags);
/* Multi-GPU: elements sharded over the ranks of an xyz_peer_group; gW9 += the sum over ALL ranks (exchanged by the
 * kernel's last CTA over NVLink mailboxes, rank-ordered, bit-identical on every rank).  n may be 0 on a rank.  */
XYZ_API int xyz_covproj_shared_w_fwd_bwd_f32_allreduce(const float* J, const float* W9, const float* S, const float* g,
                                               float* out, float* gJ, float* gW9, float* gS, long long n,
                                               const xyz_peer_group* group, unsigned long long seq,
                                               void* stream, int flags);

/* ---- C4/C5: mini-gaussian-splatting -------------------------------------------------------------
 * Replaces launch_gaussian_splatting (gaussian_splatting_kernel.cuh:38-47 /
 * gaussian_splatting_kernel.cu:114-149): renders `output` (overwritten), ADDS the L1 loss into
 * *total_loss and ADDS d(loss)/d(params) into `gradients`.  target/output are P x 3 floats
 * (PixelOutput = ConstArray<float,3>).
 *
 * Two families of entry points:
 *   xyz_launch_gaussian_splatting[_rows]   the reference's call shape: no workspace argument.  Scratch is library-owned,
 *       one arena per (device, stream) -- launches on different streams never share buffers -- and grows on demand
 *       (8 bytes per (tile, Gaussian) list entry, 40 more with XYZ_FLAG_DETERMINISTIC); the call synchronises `stream`
 *       once (the list length is data dependent) unless XYZ_FLAG_ASYNC applies.  While `stream` is being captured into a
 *       CUDA graph the scratch cannot grow (XYZ_ERR_WORKSPACE: run the iteration once outside the capture first); a
 *       buffer a capture has seen is kept alive until xyz_b200_shutdown, so replays stay valid.
 *   xyz_launch_gaussian_splatting_ws       caller-provided workspace (xyz_splat_workspace_bytes): never allocates, never
 *       synchronises, keeps no library state -- re-entrant across streams and host threads (one workspace per launch in
 *       flight), capturable from the first call.
 * The per-tile lists come from a stable counting sort by tile (row bands of up to 57 344 tiles = 14.7 Mpixel) or a
 * stable radix sort of (tile, Gaussian) keys (library sort; classic entry points only); the choice never changes a
 * result bit.  Environment (read once, tuning only): XYZ_SPLAT_BIN_CTAS_PER_SM = CTAs per SM of the counting-sort
 * kernels (default 4, 1 for predicted lists beyond 4e7 entries).                                        */
XYZ_API int xyz_launch_gaussian_splatting(const xyz_gaussian_params* gaussians, xyz_gaussian_grads* gradients,
                                  const float* target_image, float* output_image, float* total_loss,
                                  int image_width, int image_height, int num_gaussians,
                                  void* stream, int flags);
/* Row-band variant for sharding one image across GPUs: only pixel rows [row_begin, row_end)
 * are rendered / contribute loss and gradients.  Buffers are full-image sized. */
XYZ_API int xyz_launch_gaussian_splatting_rows(const xyz_gaussian_params* gaussians, xyz_gaussian_grads* gradients,
                                       const float* target_image, float* output_image, float* total_loss,
                                       int image_width, int image_height, int num_gaussians,
                                       int row_begin, int row_end, void* stream, int flags);
/* Workspace variant.  `workspace` is device memory of at least
 *   xyz_splat_workspace_bytes(width, height, N, row_begin, row_end, max_entries, flags)
 * bytes, 256-byte aligned, whose first 256 bytes were zeroed once (xyz_splat_workspace_init); the flags must be the ones
 * the size was asked for (XYZ_FLAG_RADIX_BINNING and row bands of more than 57 344 tiles are not available here:
 * XYZ_ERR_INVALID_ARGUMENT; the size query returns 0).  max_entries bounds the number of (tile, Gaussian) list entries
 * -- at BASELINE's distributions about 42 per Gaussian for a 1024^2 image.  If a scene needs more, that launch renders
 * EMPTY lists (image 0, loss = sum |target|, no gradients; memory-safe) and counts an overflow in the workspace header;
 * nothing is reported by the call itself (it does not wait for the GPU): ask xyz_splat_workspace_status when convenient.
 * Header (the first 32 bytes, device memory, 4 x uint64): {list length of the most recent launch, number of launches that
 * did not fit (sticky), internal, max_entries of the most recent launch}.                                              */
XYZ_API size_t xyz_splat_workspace_bytes(int image_width, int image_height, int num_gaussians, int row_begin, int row_end,
                                 long long max_entries, int flags);
XYZ_API int xyz_splat_workspace_init(void* workspace, size_t workspace_bytes, void* stream);
XYZ_API int xyz_launch_gaussian_splatting_ws(const xyz_gaussian_params* gaussians, xyz_gaussian_grads* gradients,
                                     const float* target_image, float* output_image, float* total_loss,
                                     int image_width, int image_height, int num_gaussians, int row_begin, int row_end,
                                     void* workspace, size_t workspace_bytes, long long max_entries,
                                     void* stream, int flags);
/* Synchronises `stream`, then status_host = {list length of the most recent launch on this workspace, overflow count
 * (sticky), its max_entries, 1 if that launch overflowed (its outputs are void) else 0}. */
XYZ_API int xyz_splat_workspace_status(const void* workspace, void* stream, long long status_host[4]);

/* Statistics of the most recent splat launch of this host thread (host values; synchronises that launch's stream if
 * its list length is not known yet): stats[0] = (tile, Gaussian) list entries, stats[1] = tiles, stats[2] = longest tile
 * list, stats[3] = pixel-Gaussian pairs evaluated per pass.  XYZ_ERR_NOT_INITIALISED if there is none or its scratch has
 * been reallocated or freed since; XYZ_ERR_WORKSPACE if it overflowed. */
XYZ_API int xyz_splat_last_stats(long long stats_host[4]);
/* What the backward pass of that launch worked on (waits for the launch and reads its work records back):
 * stats[0] = work items (list entries the backward pass keeps, see XYZ_FLAG_BWD_ALL_PAIRS), stats[1] = pixel-Gaussian
 * pairs it evaluated, stats[2] = backward CTAs with work.  Same error codes as xyz_splat_last_stats. */
XYZ_API int xyz_splat_last_backward_stats(long long stats_host[3]);
/* Stage times of this host thread's most recent splat launch made with XYZ_FLAG_TIMING, in microseconds (waits for that
 * launch): {per-Gaussian records + tile histograms, column / tile scans, list scatter, forward + loss, backward, total}. */
XYZ_API int xyz_splat_last_timing(float stage_us_host[6]);
/* Copies the integer tile-binning results of the most recent splat launch to host buffers (for the
 * bit-exact integer parity tests): per-Gaussian tile rectangles (N x 4 int32: tx0, ty0, tx1, ty1,
 * half-open), per-tile [begin, end) ranges (tiles x 2 int32; an empty tile has begin == end --
 * its running offset from the counting sort, (0, 0) from the radix path), the sorted Gaussian ids
 * (entries int32) and the per-Gaussian float records the rectangles were derived from
 * (N x 12 float: cx, cy, ia, ib, ic, sigmoid(opacity), r, g, b, and three floats that are left untouched).  Tiles outside
 * the launch's row band are reported as (0, 0); a launch that renders a row band does not write the records of
 * Gaussians that cannot reach the band (their rectangles are empty).  Any pointer may be NULL.  Synchronises the
 * launch's stream. */
XYZ_API int xyz_splat_debug_binning(int32_t* rects_host, int32_t* tile_ranges_host, int32_t* sorted_ids_host,
                            float* records_host);

/* zero_gradients_kernel (gaussian_parameters.cu:227-257). */
XYZ_API int xyz_zero_gradients(xyz_gaussian_grads* gradients, int num_gaussians, void* stream);
/* adam_step_individual_kernel (gaussian_parameters.cu:260-320, host wrapper :352-386); lr =
 * {center, scale, rotation, color, opacity}.  No clamps, like the reference's GPU kernel. */
XYZ_API int xyz_adam_step_individual(xyz_gaussian_params* params, const xyz_gaussian_grads* grads,
                             xyz_adam_state* adam, int num_gaussians, const float lr_host[5],
                             float beta1, float beta2, float epsilon, int iteration, void* stream);
/* The same step with zero_gradients_kernel fused in: every gradient is cleared right after it has been consumed, so
 * the training loop needs no xyz_zero_gradients call between iterations (one pass over the gradients less). */
XYZ_API int xyz_adam_step_individual_zero_grads(xyz_gaussian_params* params, xyz_gaussian_grads* grads,
                                        xyz_adam_state* adam, int num_gaussians, const float lr_host[5],
                                        float beta1, float beta2, float epsilon, int iteration, void* stream);
/* adam_step_kernel (gaussian_parameters.cu:173-224, host wrapper :322-350): single rate. */
XYZ_API int xyz_adam_step(xyz_gaussian_params* params, const xyz_gaussian_grads* grads, xyz_adam_state* adam,
                  int num_gaussians, float learning_rate, float beta1, float beta2, float epsilon,
                  int iteration, void* stream);

/* ---- multi-GPU: exchange of the shared-parameter gradients of C4/C5 (N x 9 floats, up to 108 MB) ---------------------
 * The reference is single-GPU (cudaSetDevice(0), gaussian_splatting_training.cu:209).  When views (C5) or row bands (C4)
 * are sharded over the GPUs of one NVSwitch box, every rank holds a partial gradient of ALL Gaussians.
 *
 * (1) NCCL (xyz_comm_init / xyz_allreduce_grads).  libnccl.so.2 is bound at run time (dlopen; the copy already
 *     loaded into the process, e.g. PyTorch's, is preferred), so the library itself has no link-time dependency.
 *     Either adopt a communicator the application already has (xyz_comm_init), or create one: rank 0 calls
 *     xyz_comm_unique_id, the 128 bytes travel by any host channel, every rank calls xyz_comm_init_rank with its device
 *     current; a single process driving several GPUs uses xyz_comm_init_all.  Collectives are enqueued on `stream`
 *     right behind the kernels that produced the gradients (no host synchronisation), in place.                       */
typedef struct xyz_comm xyz_comm;
XYZ_API int xyz_comm_unique_id(unsigned char id_out[128]);
XYZ_API int xyz_comm_init_rank(xyz_comm** comm_out, const unsigned char id[128], int rank, int world);
XYZ_API int xyz_comm_init(xyz_comm** comm_out, void* nccl_comm /* ncclComm_t, stays owned by the caller */, int rank, int world);
XYZ_API int xyz_comm_init_all(xyz_comm** comms_out /* ndev */, int ndev, const int* devices /* NULL: 0 .. ndev-1 */);
XYZ_API int xyz_comm_destroy(xyz_comm* comm);
XYZ_API int xyz_comm_rank(const xyz_comm* comm);
XYZ_API int xyz_comm_world(const xyz_comm* comm);
/* Bracket calls on SEVERAL communicators issued by ONE host thread (ncclGroupStart / ncclGroupEnd). */
XYZ_API int xyz_comm_group_start(void);
XYZ_API int xyz_comm_group_end(void);
/* grads[i] = sum over ranks, fp32, in place (ncclAllReduce).  n = number of floats (N x 9 for the gradient buffer). */
XYZ_API int xyz_allreduce_grads(xyz_comm* comm, float* grads, long long n, void* stream);
XYZ_API int xyz_allreduce_f64(xyz_comm* comm, double* values, long long n, void* stream);
/* The optimiser step of a sharded iteration without redundant work: reduce-scatter of the gradients by Gaussian range,
 * adam_step_individual (+ fused zero-grad of the WHOLE local gradient buffer) on this rank's range only, all-gather of the
 * updated parameters -- the same bytes on the wire as the all-reduce, Adam at 1 / world of its cost.  AdamState rows
 * outside the rank's range are not touched.  *total_loss (optional) is all-reduced too. */
XYZ_API int xyz_adam_step_individual_sharded(xyz_comm* comm, xyz_gaussian_params* params, xyz_gaussian_grads* grads,
                                     xyz_adam_state* adam, int num_gaussians, const float lr_host[5], float beta1,
                                     float beta2, float epsilon, int iteration, float* total_loss, void* stream);

/* (2) The same step as ONE kernel over NVLink peer memory, no NCCL call: every rank's parameter and gradient buffers are
 *     mapped into every other rank (CUDA IPC between processes: xyz_peer_alloc + xyz_peer_mailbox_open; plain peer access
 *     inside one process).  The kernel waits until every rank's gradients are complete (sequence flags in the
 *     mailboxes, st.release.sys / ld.acquire.sys), then the owner of a Gaussian range LOADS that range of every rank's
 *     gradient buffer over NVLink and adds the rows in rank order, applies adam_step_individual, STORES the new
 *     parameters into every rank's parameter buffer and zeroes the range in every rank's gradient buffer; a second
 *     flag round makes sure all remote stores have landed before any rank's next kernel runs.  Parameters are therefore
 *     bit-identical on all ranks by construction; the loss is summed in rank order on every rank.
 *     The exchange keeps its sequence numbers in the mailboxes (device memory; separate from the sequence space of the
 *     other fused exchanges), so the call is the same every time and a captured CUDA graph of a whole iteration can be
 *     replayed.  iteration >= 1 is the reference's Adam step number (bias correction); iteration == 0 means "count the
 *     steps on the device" (first call = step 1) -- what a replayed graph needs.  Every rank must make the same calls. */
typedef struct xyz_peer_splat_buffers {
    xyz_gaussian_params* params[XYZ_PEER_MAX_WORLD]; /* index = rank; [rank] = this rank's own buffer */
    xyz_gaussian_grads* grads[XYZ_PEER_MAX_WORLD];
} xyz_peer_splat_buffers;
/* cudaMalloc + zero-fill + IPC handle, for buffers other ranks will map (open / close / destroy: xyz_peer_mailbox_*). */
XYZ_API int xyz_peer_alloc(size_t bytes, void** local_ptr, unsigned char ipc_handle_out[64]);
XYZ_API int xyz_adam_step_individual_peer(const xyz_peer_group* group, const xyz_peer_splat_buffers* buffers,
                                  xyz_adam_state* adam, int num_gaussians, const float lr_host[5], float beta1,
                                  float beta2, float epsilon, int iteration, float* total_loss, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XYZ_B200_H_ */
