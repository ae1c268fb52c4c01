// oracle/xyz_oracle.cpp -- TEST INFRASTRUCTURE: the CPU restatement ("port") of the reference's
// algorithm for the hot path.  It is NOT product code: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library built from this file
// (oracle/_build/libxyz_oracle.so).  The product (libxyz_b200.so) never links or calls it.
//
// Pinning: every function here is checked (tests/test_oracle.py) against
//   (1) oracle/_ref/libxyz_ref.so -- the reference's OWN headers and splat kernel body compiled
//       for the host (oracle/ref_driver.cpp), bit-for-bit in fp64 and fp32 (both are built with
//       -ffp-contract=off), wherever /root/reference exists;
//   (2) the fixtures under tests/golden/ that were generated from (1) by
//       tests/golden/make_golden.py (committed, so the check also runs where the reference is
//       absent);
//   (3) the reference tests' known answers.
// The tile-binning functions (orc_splat_binning*) restate THIS repo's integer work -- the
// reference has none -- and are pinned only by construction:
// "parity unpinned" applies to them and to nothing else in this file.
//
// Plain C++17, no dependencies.  All citations are relative to /root/reference.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

template <class F>
void parallel_ranges(long long n, int threads, F&& fn) {
    threads = std::max(1, threads);
    if (threads == 1 || n < threads) {
        fn(0, 0LL, n);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        long long lo = n * t / threads, hi = n * (t + 1) / threads;
        pool.emplace_back([&fn, t, lo, hi] { fn(t, lo, hi); });
    }
    for (auto& th : pool) th.join();
}

// math dispatcher, host side: include/xyz_autodiff/operations/math.cuh:14-298 routes to std::.
template <class T> inline T m_exp(T x) { return std::exp(x); }
template <class T> inline T m_sigmoid(T x) { return T(1) / (T(1) + std::exp(-x)); }  // math.cuh:200-204

// ---------------------------------------------------------------------------------------------
// Per-op formulas.  out = forward value(s); gin += backward.
// Same op ids as tests/csrc/api_eval.inc.
// ---------------------------------------------------------------------------------------------
enum OpId : int {
    OP_EXP = 0, OP_SIN, OP_COS, OP_SIGMOID, OP_SQUARED, OP_NEG, OP_L1, OP_L2, OP_SUM,
    OP_ADD_C, OP_SUB_C, OP_MUL_C, OP_DIV_C, OP_CONST_ADD, OP_CONST_SUB,
    OP_SYM_INV, OP_QUAT, OP_BROADCAST3,
    OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_MATMUL,
    OP_COV_GEN, OP_MAT_TO_COV3, OP_SCALE_ROT_COV3, OP_MAHALANOBIS, OP_MAHALANOBIS_CENTER,
    OP_COUNT
};

// operations/unary/sym_matrix2_inv_logic.cuh:21-40
template <class T>
inline void sym_inv_fwd(const T* in, T* out) {
    const T a = in[0], b = in[1], c = in[2];
    T det = a * c - b * b;
    if (std::abs(det) < T(1e-8)) det = T(1e-8);
    const T inv_det = T(1) / det;
    out[0] = c * inv_det;
    out[1] = -b * inv_det;
    out[2] = a * inv_det;
}
// operations/unary/sym_matrix2_inv_logic.cuh:43-77
template <class T>
inline void sym_inv_bwd(const T* in, const T* g, T* gin) {
    const T a = in[0], b = in[1], c = in[2];
    T det = a * c - b * b;
    if (std::abs(det) < T(1e-8)) det = T(1e-8);
    const T inv_det = T(1) / det;
    const T inv_det2 = inv_det * inv_det;
    const T da_da = -c * c * inv_det2;
    const T da_db = T(2) * c * b * inv_det2;
    const T da_dc = inv_det - a * c * inv_det2;
    const T db_da = b * c * inv_det2;
    const T db_db = -inv_det - T(2) * b * b * inv_det2;
    const T db_dc = a * b * inv_det2;
    const T dc_da = inv_det - a * c * inv_det2;
    const T dc_db = T(2) * a * b * inv_det2;
    const T dc_dc = -a * a * inv_det2;
    gin[0] += g[0] * da_da + g[1] * db_da + g[2] * dc_da;
    gin[1] += g[0] * da_db + g[1] * db_db + g[2] * dc_db;
    gin[2] += g[0] * da_dc + g[1] * db_dc + g[2] * dc_dc;
}

// examples/mini-gaussian-splatting/operations/covariance_generation.cuh:154-172
template <class T>
inline void scale_rot_cov_fwd(const T* s, T theta, T* out) {
    const T c = std::cos(theta), sn = std::sin(theta);
    const T m00 = s[0] * c, m01 = -s[1] * sn, m10 = s[0] * sn, m11 = s[1] * c;
    out[0] = m00 * m00 + m01 * m01;
    out[1] = m00 * m10 + m01 * m11;
    out[2] = m10 * m10 + m11 * m11;
}
// covariance_generation.cuh:175-211
template <class T>
inline void scale_rot_cov_bwd(const T* s, T theta, const T* g, T* gs, T* gtheta) {
    const T c = std::cos(theta), sn = std::sin(theta);
    const T m00 = s[0] * c, m01 = -s[1] * sn, m10 = s[0] * sn, m11 = s[1] * c;
    const T g00 = g[0] * T(2) * m00 + g[1] * m10;
    const T g01 = g[0] * T(2) * m01 + g[1] * m11;
    const T g10 = g[1] * m00 + g[2] * T(2) * m10;
    const T g11 = g[1] * m01 + g[2] * T(2) * m11;
    gs[0] += g00 * c + g10 * sn;
    gs[1] += g01 * (-sn) + g11 * c;
    *gtheta += g00 * (-s[0] * sn) + g01 * (-s[1] * c) + g10 * (s[0] * c) + g11 * (-s[1] * sn);
}

// operations/unary/to_rotation_matrix_logic.cuh:17-46 / :49-113
template <class T>
inline void quat_fwd(const T* q, T* o) {
    const T x = q[0], y = q[1], z = q[2], w = q[3];
    o[0] = T(1) - T(2) * (y * y + z * z);
    o[1] = T(2) * (x * y - z * w);
    o[2] = T(2) * (x * z + y * w);
    o[3] = T(2) * (x * y + z * w);
    o[4] = T(1) - T(2) * (x * x + z * z);
    o[5] = T(2) * (y * z - x * w);
    o[6] = T(2) * (x * z - y * w);
    o[7] = T(2) * (y * z + x * w);
    o[8] = T(1) - T(2) * (x * x + y * y);
}
template <class T>
inline void quat_bwd(const T* q, const T* g, T* gin) {
    const T x = q[0], y = q[1], z = q[2], w = q[3];
    // coefficient tables: d r_k / d {x,y,z,w}, in the reference's accumulation order (k = 0..8)
    const T cx[9] = {T(0), T(2) * y, T(2) * z, T(2) * y, T(-4) * x, T(-2) * w, T(2) * z, T(2) * w, T(-4) * x};
    const T cy[9] = {T(-4) * y, T(2) * x, T(2) * w, T(2) * x, T(0), T(2) * z, T(-2) * w, T(2) * z, T(-4) * y};
    const T cz[9] = {T(-4) * z, T(-2) * w, T(2) * x, T(2) * w, T(-4) * z, T(2) * y, T(2) * x, T(2) * y, T(0)};
    const T cw[9] = {T(0), T(-2) * z, T(2) * y, T(2) * z, T(0), T(-2) * x, T(-2) * y, T(2) * x, T(0)};
    // the reference writes `grad += output.grad(k) * T(c) * v` = (g*c)*v; for the exact-zero rows
    // it writes `output.grad(k) * T(0)`.  (g*c)*v and g*(c*v) agree to 1 ulp; tests use 1e-12.
    T gx = 0, gy = 0, gz = 0, gw = 0;
    for (int k = 0; k < 9; ++k) {
        gx += g[k] * cx[k];
        gy += g[k] * cy[k];
        gz += g[k] * cz[k];
        gw += g[k] * cw[k];
    }
    gin[0] += gx;
    gin[1] += gy;
    gin[2] += gz;
    gin[3] += gw;
}

template <class T>
int eval_op(int op, int aux, const T* in1, int n1, const T* in2, int n2, T cst, const T* gout, T* out, int* nout,
            T* gin1, T* gin2) {
    (void)n2;
    auto zero = [](T* p, int n) { for (int i = 0; i < n; ++i) p[i] = T(0); };
    switch (op) {
        case OP_EXP:  // unary/exp_logic.cuh:17-36 (backward recomputes exp)
            for (int i = 0; i < n1; ++i) { out[i] = m_exp(in1[i]); gin1[i] = gout[i] * m_exp(in1[i]); }
            *nout = n1; return 0;
        case OP_SIN:  // unary/sin_logic.cuh:17-36
            for (int i = 0; i < n1; ++i) { out[i] = std::sin(in1[i]); gin1[i] = gout[i] * std::cos(in1[i]); }
            *nout = n1; return 0;
        case OP_COS:  // unary/cos_logic.cuh:17-36
            for (int i = 0; i < n1; ++i) { out[i] = std::cos(in1[i]); gin1[i] = gout[i] * (-std::sin(in1[i])); }
            *nout = n1; return 0;
        case OP_SIGMOID:  // unary/sigmoid_logic.cuh:17-36
            for (int i = 0; i < n1; ++i) {
                const T s = m_sigmoid(in1[i]);
                out[i] = s;
                gin1[i] = gout[i] * (s * (T(1) - s));
            }
            *nout = n1; return 0;
        case OP_SQUARED:  // unary/squared_logic.cuh:17-33 (`g * 2.0 * x`, Q5)
            for (int i = 0; i < n1; ++i) { out[i] = in1[i] * in1[i]; gin1[i] = T(double(gout[i]) * 2.0 * double(in1[i])); }
            *nout = n1; return 0;
        case OP_NEG:  // unary/neg_logic.cuh:17-34
            for (int i = 0; i < n1; ++i) { out[i] = -in1[i]; gin1[i] = -gout[i]; }
            *nout = n1; return 0;
        case OP_L1: {  // unary/l1_norm_logic.cuh:17-37
            T s = 0;
            for (int i = 0; i < n1; ++i) s += std::abs(in1[i]);
            out[0] = s;
            for (int i = 0; i < n1; ++i) gin1[i] = gout[0] * (in1[i] > T(0) ? T(1) : (in1[i] < T(0) ? T(-1) : T(0)));
            *nout = 1; return 0;
        }
        case OP_L2: {  // unary/l2_norm_logic.cuh:17-40
            T s = 0;
            for (int i = 0; i < n1; ++i) s += in1[i] * in1[i];
            const T nrm = std::sqrt(s);
            out[0] = nrm;
            zero(gin1, n1);
            if (nrm > T(1e-8)) for (int i = 0; i < n1; ++i) gin1[i] = gout[0] * in1[i] / nrm;
            *nout = 1; return 0;
        }
        case OP_SUM: {  // unary/sum_logic.cuh:17-36
            T s = 0;
            for (int i = 0; i < n1; ++i) s += in1[i];
            out[0] = s;
            for (int i = 0; i < n1; ++i) gin1[i] = gout[0];
            *nout = 1; return 0;
        }
        case OP_ADD_C:  // unary/add_constant_logic.cuh:11-48
            for (int i = 0; i < n1; ++i) { out[i] = in1[i] + cst; gin1[i] = gout[i]; }
            *nout = n1; return 0;
        case OP_SUB_C:  // unary/sub_constant_logic.cuh:11-48
            for (int i = 0; i < n1; ++i) { out[i] = in1[i] - cst; gin1[i] = gout[i]; }
            *nout = n1; return 0;
        case OP_MUL_C:  // unary/mul_constant_logic.cuh:11-48
            for (int i = 0; i < n1; ++i) { out[i] = in1[i] * cst; gin1[i] = gout[i] * cst; }
            *nout = n1; return 0;
        case OP_DIV_C:  // unary/div_constant_logic.cuh:11-48
            for (int i = 0; i < n1; ++i) { out[i] = in1[i] / cst; gin1[i] = gout[i] / cst; }
            *nout = n1; return 0;
        case OP_CONST_ADD:  // unary/const_array_add_logic.cuh:14-48
            for (int i = 0; i < n1; ++i) { out[i] = in1[i] + in2[i]; gin1[i] = gout[i]; }
            *nout = n1; return 0;
        case OP_CONST_SUB:  // unary/const_array_sub_logic.cuh:14-48
            for (int i = 0; i < n1; ++i) { out[i] = in1[i] - in2[i]; gin1[i] = gout[i]; }
            *nout = n1; return 0;
        case OP_SYM_INV:
            sym_inv_fwd(in1, out);
            zero(gin1, 3);
            sym_inv_bwd(in1, gout, gin1);
            *nout = 3; return 0;
        case OP_QUAT:
            quat_fwd(in1, out);
            zero(gin1, 4);
            quat_bwd(in1, gout, gin1);
            *nout = 9; return 0;
        case OP_BROADCAST3:  // unary/broadcast.cuh:28-46: a view; adjoints sum into the scalar
            for (int j = 0; j < 3; ++j) out[j] = in1[0];
            gin1[0] = T(0);
            for (int j = 0; j < 3; ++j) gin1[0] += gout[j];
            *nout = 3; return 0;
        case OP_ADD:  // binary/add_logic.cuh:11-50
            for (int i = 0; i < n1; ++i) { out[i] = in1[i] + in2[i]; gin1[i] = gout[i]; gin2[i] = gout[i]; }
            *nout = n1; return 0;
        case OP_SUB:  // binary/sub_logic.cuh:11-52
            for (int i = 0; i < n1; ++i) { out[i] = in1[i] - in2[i]; gin1[i] = gout[i]; gin2[i] = -gout[i]; }
            *nout = n1; return 0;
        case OP_MUL:  // binary/mul_logic.cuh:11-50
            for (int i = 0; i < n1; ++i) { out[i] = in1[i] * in2[i]; gin1[i] = gout[i] * in2[i]; gin2[i] = gout[i] * in1[i]; }
            *nout = n1; return 0;
        case OP_DIV:  // binary/div_logic.cuh:11-50
            for (int i = 0; i < n1; ++i) {
                out[i] = in1[i] / in2[i];
                gin1[i] = gout[i] / in2[i];
                gin2[i] = -gout[i] * in1[i] / (in2[i] * in2[i]);
            }
            *nout = n1; return 0;
        case OP_MATMUL: {  // binary/matmul_logic.cuh:33-69 (row-major; adds in i -> j -> k order)
            const int a = aux / 100, b = (aux / 10) % 10, c = aux % 10;
            for (int i = 0; i < a; ++i)
                for (int j = 0; j < c; ++j) {
                    T s = 0;
                    for (int k = 0; k < b; ++k) s += in1[i * b + k] * in2[k * c + j];
                    out[i * c + j] = s;
                }
            zero(gin1, a * b);
            zero(gin2, b * c);
            for (int i = 0; i < a; ++i)
                for (int j = 0; j < c; ++j) {
                    const T g = gout[i * c + j];
                    for (int k = 0; k < b; ++k) gin1[i * b + k] += g * in2[k * c + j];
                    for (int k = 0; k < b; ++k) gin2[k * c + j] += g * in1[i * b + k];
                }
            *nout = a * c; return 0;
        }
        case OP_COV_GEN: {  // covariance_generation.cuh:28-74: M = R(theta) diag(s)
            const T c = std::cos(in2[0]), sn = std::sin(in2[0]);
            out[0] = in1[0] * c; out[1] = -in1[1] * sn; out[2] = in1[0] * sn; out[3] = in1[1] * c;
            gin1[0] = gout[0] * c + gout[2] * sn;
            gin1[1] = gout[1] * (-sn) + gout[3] * c;
            gin2[0] = gout[0] * (-in1[0] * sn) + gout[1] * (-in1[1] * c) + gout[2] * (in1[0] * c) + gout[3] * (-in1[1] * sn);
            *nout = 4; return 0;
        }
        case OP_MAT_TO_COV3: {  // covariance_generation.cuh:94-135: Sigma = M M^T packed
            const T m00 = in1[0], m01 = in1[1], m10 = in1[2], m11 = in1[3];
            out[0] = m00 * m00 + m01 * m01;
            out[1] = m00 * m10 + m01 * m11;
            out[2] = m10 * m10 + m11 * m11;
            gin1[0] = gout[0] * T(2) * m00 + gout[1] * m10;
            gin1[1] = gout[0] * T(2) * m01 + gout[1] * m11;
            gin1[2] = gout[1] * m00 + gout[2] * T(2) * m10;
            gin1[3] = gout[1] * m01 + gout[2] * T(2) * m11;
            *nout = 3; return 0;
        }
        case OP_SCALE_ROT_COV3:
            scale_rot_cov_fwd(in1, in2[0], out);
            zero(gin1, 2);
            gin2[0] = T(0);
            scale_rot_cov_bwd(in1, in2[0], gout, gin1, gin2);
            *nout = 3; return 0;
        case OP_MAHALANOBIS: {  // mahalanobis_distance.cuh:30-63
            const T dx = in1[0], dy = in1[1], a = in2[0], b = in2[1], c = in2[2], g = gout[0];
            out[0] = a * dx * dx + T(2) * b * dx * dy + c * dy * dy;
            gin1[0] = g * (T(2) * a * dx + T(2) * b * dy);
            gin1[1] = g * (T(2) * b * dx + T(2) * c * dy);
            gin2[0] = g * dx * dx;
            gin2[1] = g * T(2) * dx * dy;
            gin2[2] = g * dy * dy;
            *nout = 1; return 0;
        }
        case OP_MAHALANOBIS_CENTER: {  // mahalanobis_distance.cuh:88-119 (query point in cst, gout[1])
            const T dx = cst - in1[0], dy = gout[1] - in1[1], a = in2[0], b = in2[1], c = in2[2], g = gout[0];
            out[0] = a * dx * dx + T(2) * b * dx * dy + c * dy * dy;
            gin1[0] = -(g * (T(2) * a * dx + T(2) * b * dy));
            gin1[1] = -(g * (T(2) * b * dx + T(2) * c * dy));
            gin2[0] = g * dx * dx;
            gin2[1] = g * T(2) * dx * dy;
            gin2[2] = g * dy * dy;
            *nout = 1; return 0;
        }
        default: return -1;
    }
}

// ---------------------------------------------------------------------------------------------
// Splat: all-pairs fused render + L1 loss + backward, statement order of
// examples/mini-gaussian-splatting/gaussian_splatting_kernel.cu:19-111, including
//   Q1 (the cull at :48-50 / :90-92 reads an un-forwarded node: never true) and
//   Q2 (rest_sum += weighted_color at :102-103 adds zeros).
// Gaussian g: {center[2], scale[2], rotation[1], color[3], opacity[1]} (gaussian_parameters.h:12-18).
// ---------------------------------------------------------------------------------------------
template <class T>
struct PairFwd {
    T es[2], cov[3], inv[3], dx, dy, d2, e, so, w, wc[3];
};

template <class T>
inline void pair_forward(const T* gp, T px, T py, PairFwd<T>& f) {
    f.es[0] = m_exp(gp[2]);                                    // op::exp(scale)                         kernel.cu:44
    f.es[1] = m_exp(gp[3]);
    scale_rot_cov_fwd(f.es, gp[4], f.cov);                     // scale_rotation_to_covariance_3param     :45
    sym_inv_fwd(f.cov, f.inv);                                 // sym_matrix2_inv                         :46
    f.dx = px - gp[0];                                         // mahalanobis_distance_with_center        :47
    f.dy = py - gp[1];
    f.d2 = f.inv[0] * f.dx * f.dx + T(2) * f.inv[1] * f.dx * f.dy + f.inv[2] * f.dy * f.dy;
    const T sd = f.d2 * T(0.5);                                // * 0.5f (mul_constant)                   :51
    const T ns = -sd;                                          // op::neg                                 :52
    f.e = m_exp(ns);                                           // op::exp                                 :53
    f.so = m_sigmoid(gp[8]);                                   // op::sigmoid(opacity)                    :54
    f.w = f.e * f.so;                                          // gaussian_value * sig_opacity            :55
    for (int i = 0; i < 3; ++i) f.wc[i] = gp[5 + i] * f.w;     // color * broadcast<3>(weighted_gauss)    :56-57
}

// How many units of relative rounding error of the INPUTS the exponent d2 / 2 of one pair amplifies: the size of the
// three products whose signed sum it is, times the conditioning of the 2x2 inverse they are built from
// (det = A C - B^2 cancels for strongly anisotropic, rotated covariances: every entry of Sigma^-1 inherits the
// relative error (|A C| + B^2) / |det| ulp).
template <class T>
T exponent_sensitivity(const PairFwd<T>& f) {
    const T mag = T(0.5) * (std::abs(f.inv[0]) * f.dx * f.dx + T(2) * std::abs(f.inv[1] * f.dx * f.dy) +
                            std::abs(f.inv[2]) * f.dy * f.dy);
    const T det = f.cov[0] * f.cov[2] - f.cov[1] * f.cov[1];
    const T det_cond = (std::abs(f.cov[0] * f.cov[2]) + f.cov[1] * f.cov[1]) / std::max(std::abs(det), T(1e-30));
    return mag * (T(1) + det_cond);
}

// The backward of ONE (pixel, Gaussian) pair: l1_loss.run() of gaussian_splatting_kernel.cu:84-110 with `pix` as the
// finished pixel_out.  gg (9) += ; absgrads / kinkgrads / condgrads (9 each, optional) as described at splat_pixel.
template <class T>
inline void pair_backward(const T* gp, T* gg, const T* tgt, const T* pix, T px, T py, T* margin, T* absgrads, T* kinkgrads,
                          T* condgrads) {
    PairFwd<T> f;
    T before[9];
    if (absgrads) for (int k = 0; k < 9; ++k) before[k] = gg[k];
    pair_forward(gp, px, py, f);                            // l1_loss.run() -> forward
    T sgn[3];
    bool near_kink = false;
    for (int i = 0; i < 3; ++i) {
        T rest = tgt[i] - pix[i];                           // rest_sum = target - pixel_out            :101
        rest += T(0);                                       // += un-forwarded weighted_color (Q2)      :102
        const T cd = f.wc[i] - rest;                        // color_diff = weighted_color - rest_sum   :106
        sgn[i] = cd > T(0) ? T(1) : (cd < T(0) ? T(-1) : T(0));  // l1 backward, seed 1.0
        if (margin && f.w > T(1e-10) && std::abs(cd) < *margin) *margin = std::abs(cd);
        // sign(cd) is numerically ambiguous in fp32 when |cd| is within rounding distance of 0
        // (rest = tgt - out carries the ABSOLUTE rounding error of the fp32 image, ~1e-6 * |out|)
        if (std::abs(cd) <= T(2e-5) * std::max(std::abs(f.wc[i]), std::max(std::abs(tgt[i]), std::abs(pix[i]))))
            near_kink = true;
    }
    // mul backward (binary/mul_logic.cuh:33-41): color.grad += g*bc ; bc(=weighted_gauss).grad += g*color
    T g_w = T(0);
    for (int i = 0; i < 3; ++i) {
        gg[5 + i] += sgn[i] * f.w;
        g_w += sgn[i] * gp[5 + i];
    }
    // weighted_gauss = gaussian_value * sig_opacity
    const T g_e = g_w * f.so;
    const T g_so = g_w * f.e;
    // gaussian_value chain first (input1), then sigmoid (input2): operation.cuh:235-249
    const T g_ns = g_e * m_exp(-(f.d2 * T(0.5)));           // exp backward recomputes exp(input)
    const T g_sd = -g_ns;                                   // neg backward
    const T g_d2 = g_sd * T(0.5);                           // mul_constant backward
    // mahalanobis_distance.cuh:100-119
    const T a = f.inv[0], b = f.inv[1], c = f.inv[2];
    const T grad_dx = g_d2 * (T(2) * a * f.dx + T(2) * b * f.dy);
    const T grad_dy = g_d2 * (T(2) * b * f.dx + T(2) * c * f.dy);
    gg[0] += -grad_dx;
    gg[1] += -grad_dy;
    const T g_inv[3] = {g_d2 * f.dx * f.dx, g_d2 * T(2) * f.dx * f.dy, g_d2 * f.dy * f.dy};
    T g_cov[3] = {T(0), T(0), T(0)};
    sym_inv_bwd(f.cov, g_inv, g_cov);
    T g_es[2] = {T(0), T(0)};
    scale_rot_cov_bwd(f.es, gp[4], g_cov, g_es, &gg[4]);    // rotation leaf: direct accumulate
    gg[2] += g_es[0] * m_exp(gp[2]);                        // exp backward (recomputed)
    gg[3] += g_es[1] * m_exp(gp[3]);
    const T s = m_sigmoid(gp[8]);                           // sigmoid backward (recomputed)
    gg[8] += g_so * (s * (T(1) - s));
    if (absgrads)  // sum of |per-pair term|: the scale the 1e-4 tolerance on accumulated sums is stated against
        for (int k = 0; k < 9; ++k) {
            const T term = std::abs(gg[k] - before[k]);
            absgrads[k] += term;
            // upper bound of what a flipped sign can change: |g_w| <= sum |color_i| instead of |sum s_i color_i|
            if (kinkgrads && near_kink) kinkgrads[k] += term + std::abs(f.w) + T(1e-30);
            // every term is proportional to exp(-d2 / 2); an fp32 evaluation of d2 / 2 = sum of three products of
            // rounded factors carries an ABSOLUTE error of a few ulp of their magnitudes, i.e. the term a RELATIVE
            // error of (a few 2^-24) x mag.  condgrads = sum |term| x mag is that sensitivity.
            if (condgrads) {
                const T mag = exponent_sensitivity(f);
                condgrads[k] += term * mag;
            }
        }
}

// One pixel of the reference kernel.  grads += ; *loss += ; out[3] written.  `margin` (optional)
// tracks min |color_diff_i| over all pairs whose weight is not negligible (> 1e-10): the distance of the
// closest L1 kink (tests use it to make sure a sign cannot flip within fp32 noise).
template <class T>
void splat_pixel(const T* params, T* grads, const T* tgt, T* out, T* loss, int px_i, int py_i, int N, T* margin,
                 T* absgrads, T* kinkgrads, T* condgrads = nullptr, T* condimg = nullptr) {
    const T px = static_cast<T>(px_i), py = static_cast<T>(py_i);
    T pix[3] = {T(0), T(0), T(0)};
    PairFwd<T> f;
    for (int g = 0; g < N; ++g) {                               // kernel.cu:33-62
        pair_forward(params + 9 * g, px, py, f);
        for (int i = 0; i < 3; ++i) pix[i] += f.wc[i];
        // conditioning of the exponent: the size of the three terms whose signed sum is d2 / 2 (see condgrads below)
        if (condimg) {
            const T mag = exponent_sensitivity(f);
            for (int i = 0; i < 3; ++i) condimg[i] += std::abs(f.wc[i]) * mag;
        }
    }
    for (int i = 0; i < 3; ++i) out[i] = pix[i];                // :63
    for (int i = 0; i < 3; ++i) *loss += std::abs(pix[i] - tgt[i]);  // :68-70
    for (int g = 0; g < N; ++g)                                 // :73-111
        pair_backward(params + 9 * g, grads + 9 * g, tgt, pix, px, py, margin, absgrads ? absgrads + 9 * g : nullptr,
                      (absgrads && kinkgrads) ? kinkgrads + 9 * g : nullptr,
                      (absgrads && condgrads) ? condgrads + 9 * g : nullptr);
}

// Pixel order = the reference launch's block order run sequentially (16x16 tiles, row-major
// blocks, threadIdx.y outer / threadIdx.x inner) so that a single-threaded run accumulates in
// the same order as oracle/ref_driver.cpp.
template <class T>
int splat_all_pairs(const T* params, T* grads, const T* target, T* output, T* loss, int W, int H, int N, int threads,
                    T* margin_out, T* absgrads_out, T* kinkgrads_out, T* condgrads_out = nullptr, T* condimg_out = nullptr) {
    constexpr int TS = 16;  // gaussian_splatting_kernel.cuh:21
    const int bx = (W + TS - 1) / TS, by = (H + TS - 1) / TS;
    const long long nblocks = 1LL * bx * by;
    threads = static_cast<int>(std::max<long long>(1, std::min<long long>(threads, nblocks)));
    std::vector<std::vector<T>> priv_g(threads), priv_a(threads), priv_k(threads), priv_c(threads);
    std::vector<T> priv_l(threads, T(0)), priv_m(threads, T(1e30));
    parallel_ranges(nblocks, threads, [&](int t, long long lo, long long hi) {
        priv_g[t].assign(static_cast<size_t>(N) * 9, T(0));
        if (absgrads_out) priv_a[t].assign(static_cast<size_t>(N) * 9, T(0));
        if (kinkgrads_out) priv_k[t].assign(static_cast<size_t>(N) * 9, T(0));
        if (condgrads_out) priv_c[t].assign(static_cast<size_t>(N) * 9, T(0));
        for (long long b = lo; b < hi; ++b) {
            const int bxi = static_cast<int>(b % bx), byi = static_cast<int>(b / bx);
            for (int ty = 0; ty < TS; ++ty)
                for (int tx = 0; tx < TS; ++tx) {
                    const int x = bxi * TS + tx, y = byi * TS + ty;
                    if (x >= W || y >= H) continue;
                    const size_t p = static_cast<size_t>(y) * W + x;
                    splat_pixel<T>(params, priv_g[t].data(), target + 3 * p, output + 3 * p, &priv_l[t], x, y, N,
                                   margin_out ? &priv_m[t] : nullptr, absgrads_out ? priv_a[t].data() : nullptr,
                                   kinkgrads_out ? priv_k[t].data() : nullptr,
                                   (condgrads_out && absgrads_out) ? priv_c[t].data() : nullptr,
                                   condimg_out ? condimg_out + 3 * p : nullptr);  // one writer per pixel
                }
        }
    });
    T m = T(1e30);
    for (int t = 0; t < threads; ++t) {
        for (size_t i = 0; i < static_cast<size_t>(N) * 9; ++i) grads[i] += priv_g[t][i];
        if (absgrads_out)
            for (size_t i = 0; i < static_cast<size_t>(N) * 9; ++i) absgrads_out[i] += priv_a[t][i];
        if (kinkgrads_out)
            for (size_t i = 0; i < static_cast<size_t>(N) * 9; ++i) kinkgrads_out[i] += priv_k[t][i];
        if (condgrads_out && absgrads_out)
            for (size_t i = 0; i < static_cast<size_t>(N) * 9; ++i) condgrads_out[i] += priv_c[t][i];
        *loss += priv_l[t];
        m = std::min(m, priv_m[t]);
    }
    if (margin_out) *margin_out = m;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// SAMPLED all-pairs checks for scenes too large to run splat_all_pairs on (100 K Gaussians x 1024^2 is 2 x 10^11 pairs;
// the reference kernel needs ~50 s of a B200 for it).  Same statements as splat_pixel, restricted to
//   (a) a set of PIXELS, every Gaussian: the forward loop gaussian_splatting_kernel.cu:33-62 (ascending Gaussian index);
//   (b) a set of GAUSSIANS, every pixel: the backward loop :73-111 with a GIVEN image as pixel_out (the image under test,
//       which (a) checks separately) -- the gradient of one Gaussian is a sum over pixels of pair terms that depend on the
//       other Gaussians only through pixel_out.
// The per-Gaussian part of pair_forward does not depend on the pixel; (a) evaluates it once per Gaussian through the
// same functions, so every value is bit-identical to pair_forward's (this file is compiled with -ffp-contract=off).
// ---------------------------------------------------------------------------------------------
template <class T>
struct GaussFwd {
    T cx, cy, cov[3], inv[3], so, col[3];
};

template <class T>
int splat_pixels_sample(const T* params, int N, const int* xy, int n, T* out, T* condimg, int threads) {
    std::vector<GaussFwd<T>> pre(static_cast<size_t>(N));
    parallel_ranges(N, threads, [&](int, long long lo, long long hi) {
        for (long long g = lo; g < hi; ++g) {
            const T* gp = params + 9 * g;
            GaussFwd<T>& q = pre[g];
            T es[2] = {m_exp(gp[2]), m_exp(gp[3])};               // kernel.cu:44
            scale_rot_cov_fwd(es, gp[4], q.cov);                  // :45
            sym_inv_fwd(q.cov, q.inv);                            // :46
            q.cx = gp[0];
            q.cy = gp[1];
            q.so = m_sigmoid(gp[8]);                              // :54
            for (int i = 0; i < 3; ++i) q.col[i] = gp[5 + i];
        }
    });
    constexpr int kBlock = 2048;  // Gaussians per pass over a thread's pixels (keeps them in cache)
    parallel_ranges(n, threads, [&](int, long long lo, long long hi) {
        for (long long p = lo; p < hi; ++p)
            for (int i = 0; i < 3; ++i) {
                out[3 * p + i] = T(0);
                if (condimg) condimg[3 * p + i] = T(0);
            }
        for (int g0 = 0; g0 < N; g0 += kBlock) {
            const int g1 = std::min(N, g0 + kBlock);
            for (long long p = lo; p < hi; ++p) {
                const T px = static_cast<T>(xy[2 * p]), py = static_cast<T>(xy[2 * p + 1]);
                T* o = out + 3 * p;
                for (int g = g0; g < g1; ++g) {                   // ascending g per pixel, like :33-62
                    const GaussFwd<T>& q = pre[g];
                    PairFwd<T> f;
                    f.dx = px - q.cx;                             // :47
                    f.dy = py - q.cy;
                    f.d2 = q.inv[0] * f.dx * f.dx + T(2) * q.inv[1] * f.dx * f.dy + q.inv[2] * f.dy * f.dy;
                    const T sd = f.d2 * T(0.5);                   // :51
                    const T ns = -sd;                             // :52
                    f.e = m_exp(ns);                              // :53
                    f.w = f.e * q.so;                             // :55
                    for (int i = 0; i < 3; ++i) {
                        f.wc[i] = q.col[i] * f.w;                 // :56-57
                        o[i] += f.wc[i];                          // :61
                    }
                    if (condimg) {
                        for (int i = 0; i < 3; ++i) {
                            f.inv[i] = q.inv[i];
                            f.cov[i] = q.cov[i];
                        }
                        const T mag = exponent_sensitivity(f);
                        for (int i = 0; i < 3; ++i) condimg[3 * p + i] += std::abs(f.wc[i]) * mag;
                    }
                }
            }
        }
    });
    return 0;
}

template <class T>
int splat_grads_sample(const T* params, const int* ids, int m, const T* target, const T* image, int W, int H, T* grads,
                       T* absgrads, T* kinkgrads, int threads) {
    parallel_ranges(m, threads, [&](int, long long lo, long long hi) {
        for (long long s = lo; s < hi; ++s) {
            const T* gp = params + 9 * static_cast<size_t>(ids[s]);
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const size_t p = static_cast<size_t>(y) * W + x;
                    pair_backward<T>(gp, grads + 9 * s, target + 3 * p, image + 3 * p, static_cast<T>(x), static_cast<T>(y),
                                     nullptr, absgrads ? absgrads + 9 * s : nullptr,
                                     (absgrads && kinkgrads) ? kinkgrads + 9 * s : nullptr, nullptr);
                }
        }
    });
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Covariance projection S' = (J W) S (J W)^T, forward + reverse, closed form (fp64 or fp32).
// Restates the tree of op::matmul<2,3,3>, <2,3,3>, <3,3,2>, <2,3,2> nodes of
// oracle/ref_driver.cpp::covproj_one (binary/matmul_logic.cuh:33-69 semantics) with the two
// transposed leaves folded back:  G = [[g0, g1], [0, g2]] is the dense adjoint of the 2x2 product.
// ---------------------------------------------------------------------------------------------
template <class T>
inline void covproj_one(const T* J, const T* W, const T* S, const T* g, T* out, T* gJ, T* gW, T* gS) {
    const T Sf[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
    T Tm[6], U[6];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) {
            T s = 0;
            for (int k = 0; k < 3; ++k) s += J[i * 3 + k] * W[k * 3 + j];
            Tm[i * 3 + j] = s;
        }
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) {
            T s = 0;
            for (int k = 0; k < 3; ++k) s += Tm[i * 3 + k] * Sf[k * 3 + j];
            U[i * 3 + j] = s;
        }
    T P[4];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            T s = 0;
            for (int k = 0; k < 3; ++k) s += U[i * 3 + k] * Tm[j * 3 + k];
            P[i * 2 + j] = s;
        }
    out[0] = P[0];
    out[1] = P[1];
    out[2] = P[3];
    const T G[4] = {g[0], g[1], T(0), g[2]};
    // dU = G T ; dT(b) = G^T U  (through T^T) ; dT(a) = dU S^T ; dS = T^T dU
    T dU[6], dTa[6], dTb[6], dSf[9];
    for (int i = 0; i < 2; ++i)
        for (int k = 0; k < 3; ++k) {
            T s = 0;
            for (int j = 0; j < 2; ++j) s += G[i * 2 + j] * Tm[j * 3 + k];
            dU[i * 3 + k] = s;
        }
    for (int j = 0; j < 2; ++j)
        for (int k = 0; k < 3; ++k) {
            T s = 0;
            for (int i = 0; i < 2; ++i) s += G[i * 2 + j] * U[i * 3 + k];
            dTb[j * 3 + k] = s;
        }
    for (int i = 0; i < 2; ++i)
        for (int k = 0; k < 3; ++k) {
            T s = 0;
            for (int j = 0; j < 3; ++j) s += dU[i * 3 + j] * Sf[k * 3 + j];
            dTa[i * 3 + k] = s;
        }
    for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j) {
            T s = 0;
            for (int i = 0; i < 2; ++i) s += Tm[i * 3 + k] * dU[i * 3 + j];
            dSf[k * 3 + j] = s;
        }
    // each path is pushed through J W separately and summed last, like the tree does
    for (int i = 0; i < 2; ++i)
        for (int k = 0; k < 3; ++k) {
            T sa = 0, sb = 0;
            for (int j = 0; j < 3; ++j) {
                sa += dTa[i * 3 + j] * W[k * 3 + j];
                sb += dTb[i * 3 + j] * W[k * 3 + j];
            }
            gJ[i * 3 + k] = sa + sb;
        }
    for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j) {
            T sa = 0, sb = 0;
            for (int i = 0; i < 2; ++i) {
                sa += J[i * 3 + k] * dTa[i * 3 + j];
                sb += J[i * 3 + k] * dTb[i * 3 + j];
            }
            gW[k * 3 + j] = sa + sb;
        }
    gS[0] = dSf[0];
    gS[1] = dSf[1] + dSf[3];
    gS[2] = dSf[2] + dSf[6];
    gS[3] = dSf[4];
    gS[4] = dSf[5] + dSf[7];
    gS[5] = dSf[8];
}

// ---------------------------------------------------------------------------------------------
// Tile binning (THIS repo's integer work; no reference counterpart).  Mirrors
// xyz-autodiff-cuda_b200/csrc/splat_common.cuh::gaussian_tile_rect -- every float op below is a single
// IEEE-754 binary32 operation (this file is compiled with -ffp-contract=off; the CUDA side uses
// __fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn), so identical record floats give identical integers.
// rec = {cx, cy, ia, ib, ic, ...}; rect = {tx0, ty0, tx1, ty1} half-open, in 16-pixel tiles.
// ---------------------------------------------------------------------------------------------
constexpr int kTile = 16;

inline void tile_rect(const float* rec, int W, int H, int row_begin, int row_end, float d2max, int no_cull, int32_t* rect) {
    const int tiles_x = (W + kTile - 1) / kTile;
    const int ty_lo = row_begin / kTile, ty_hi = (row_end + kTile - 1) / kTile;
    int tx0 = 0, tx1 = tiles_x, ty0 = ty_lo, ty1 = ty_hi;
    const float cx = rec[0], cy = rec[1], ia = rec[2], ib = rec[3], ic = rec[4];
    const float det = ia * ic - ib * ib;
    const bool ok = !no_cull && det > 0.0f && ia > 0.0f && ic > 0.0f && std::isfinite(det) && std::isfinite(cx) &&
                    std::isfinite(cy) && std::isfinite(ia) && std::isfinite(ic);
    if (ok) {
        const float hx = std::sqrt(d2max * ic / det) * 1.001f + 1.0f;
        const float hy = std::sqrt(d2max * ia / det) * 1.001f + 1.0f;
        if (std::isfinite(hx) && std::isfinite(hy)) {
            const float x_lo = std::floor(cx - hx), x_hi = std::ceil(cx + hx);
            const float y_lo = std::floor(cy - hy), y_hi = std::ceil(cy + hy);
            if (x_hi < 0.0f || y_hi < static_cast<float>(row_begin) || x_lo > static_cast<float>(W - 1) ||
                y_lo > static_cast<float>(row_end - 1)) {
                tx0 = tx1 = ty0 = ty1 = 0;
            } else {
                const int xi0 = static_cast<int>(std::max(x_lo, 0.0f));
                const int xi1 = static_cast<int>(std::min(x_hi, static_cast<float>(W - 1)));
                const int yi0 = static_cast<int>(std::max(y_lo, static_cast<float>(row_begin)));
                const int yi1 = static_cast<int>(std::min(y_hi, static_cast<float>(row_end - 1)));
                tx0 = xi0 / kTile;
                tx1 = xi1 / kTile + 1;
                ty0 = yi0 / kTile;
                ty1 = yi1 / kTile + 1;
            }
        }
    }
    rect[0] = tx0;
    rect[1] = ty0;
    rect[2] = tx1;
    rect[3] = ty1;
}

// Second level of the cull: the span of tiles of row ty that the ellipse d2 <= d2max can reach (mirrors
// csrc/splat_host.cu::span_coef / span_edge / tile_row_span operation for operation).
struct SpanCoef {
    float cx, cy, k, a, b, hxv, hyv_m, dyR;
    int ok;
};
inline SpanCoef span_coef(const float* rec, float d2max, int no_cull) {
    SpanCoef c;
    const float cx = rec[0], cy = rec[1], ia = rec[2], ib = rec[3], ic = rec[4];
    c.cx = cx;
    c.cy = cy;
    const float det = ia * ic - ib * ib;
    c.ok = !no_cull && det > 0.0f && ia > 0.0f && ic > 0.0f && std::isfinite(det) && std::isfinite(cx) &&
           std::isfinite(cy) && std::isfinite(ia) && std::isfinite(ic);
    c.k = ib / ia;
    c.a = d2max / ia;
    c.b = det / (ia * ia);
    c.hxv = std::sqrt(d2max * ic / det);
    c.hyv_m = std::sqrt(d2max * ia / det) * 1.001f + 1.0f;
    c.dyR = -(ib / ic) * c.hxv;
    if (!(std::isfinite(c.k) && std::isfinite(c.a) && std::isfinite(c.b) && std::isfinite(c.hxv) && std::isfinite(c.hyv_m) &&
          std::isfinite(c.dyR)))
        c.ok = 0;
    return c;
}
inline float span_edge(const SpanCoef& c, float dy, float sign) {
    const float rad = std::max(c.a - c.b * (dy * dy), 0.0f);
    return (-c.k) * dy + sign * std::sqrt(rad);
}
inline void tile_row_span(const SpanCoef& c, const int32_t* r, int ty, int W, int row_begin, int row_end, int* s0, int* s1) {
    *s0 = r[0];
    *s1 = r[2];
    if (!c.ok) return;
    const int y0 = std::max(ty * kTile, row_begin), y1 = std::min(ty * kTile + kTile - 1, row_end - 1);
    const float lo = std::max(static_cast<float>(y0) - c.cy, -c.hyv_m);
    const float hi = std::min(static_cast<float>(y1) - c.cy, c.hyv_m);
    if (lo > hi) {
        *s0 = *s1 = 0;
        return;
    }
    const float fmx = (lo <= c.dyR && c.dyR <= hi) ? c.hxv : std::max(span_edge(c, lo, 1.0f), span_edge(c, hi, 1.0f));
    const float gmn = (lo <= -c.dyR && -c.dyR <= hi) ? -c.hxv : std::min(span_edge(c, lo, -1.0f), span_edge(c, hi, -1.0f));
    const float xr = std::ceil(c.cx + (fmx + (std::fabs(fmx) * 0.001f + 1.0f)));
    const float xl = std::floor(c.cx - (-gmn + (std::fabs(gmn) * 0.001f + 1.0f)));
    if (!(std::isfinite(xr) && std::isfinite(xl))) return;
    if (xr < 0.0f || xl > static_cast<float>(W - 1)) {
        *s0 = *s1 = 0;
        return;
    }
    const int xi0 = static_cast<int>(std::max(xl, 0.0f));
    const int xi1 = static_cast<int>(std::min(xr, static_cast<float>(W - 1)));
    *s0 = std::max(xi0 / kTile, r[0]);
    *s1 = std::min(xi1 / kTile + 1, r[2]);
    if (*s1 < *s0) *s1 = *s0;
}

}  // namespace

extern "C" {

const char* orc_kind() { return "port"; }

int orc_eval_op_f64(int op, int aux, const double* in1, int n1, const double* in2, int n2, double cst,
                    const double* gout, double* out, int* nout, double* gin1, double* gin2) {
    return eval_op<double>(op, aux, in1, n1, in2, n2, cst, gout, out, nout, gin1, gin2);
}
int orc_eval_op_f32(int op, int aux, const float* in1, int n1, const float* in2, int n2, float cst, const float* gout,
                    float* out, int* nout, float* gin1, float* gin2) {
    return eval_op<float>(op, aux, in1, n1, in2, n2, cst, gout, out, nout, gin1, gin2);
}

int orc_splat_f32(const float* params, float* grads, const float* target, float* output, float* loss, int W, int H,
                  int N, int threads) {
    return splat_all_pairs<float>(params, grads, target, output, loss, W, H, N, threads, nullptr, nullptr, nullptr);
}
// fp64 ground truth of the same statements; *margin = min |color_diff| over all pairs and channels;
// absgrads (N x 9, optional, +=) = sum over pairs of |per-pair gradient term|; kinkgrads (N x 9, optional,
// needs absgrads, +=) = the same sum restricted to pairs whose L1 argument is within fp32 rounding
// distance of 0 (their sign, hence their term, is ambiguous for ANY fp32 implementation).
int orc_splat_f64(const double* params, double* grads, const double* target, double* output, double* loss, int W,
                  int H, int N, int threads, double* margin, double* absgrads, double* kinkgrads) {
    return splat_all_pairs<double>(params, grads, target, output, loss, W, H, N, threads, margin, absgrads, kinkgrads);
}
// As orc_splat_f64, plus the fp32 CONDITIONING of the scene: condgrads (N x 9, +=, needs absgrads) = sum over pairs of
// |term| x mag and condimage (P x 3, +=) = sum over Gaussians of |contribution| x mag, where mag = (|a| dx^2 +
// 2 |b dx dy| + |c| dy^2) / 2 is the size of the products whose signed sum is the exponent.  An fp32 evaluation of the
// exponent is off by a few 2^-24 x mag, so (a few 2^-24) x cond* bounds what ANY fp32 implementation -- the
// reference's included -- can promise on scenes with sub-pixel or strongly anisotropic Gaussians.
int orc_splat_f64_cond(const double* params, double* grads, const double* target, double* output, double* loss, int W,
                       int H, int N, int threads, double* margin, double* absgrads, double* kinkgrads, double* condgrads,
                       double* condimage) {
    return splat_all_pairs<double>(params, grads, target, output, loss, W, H, N, threads, margin, absgrads, kinkgrads,
                                   condgrads, condimage);
}

// Sampled checks for full-size scenes (see splat_pixels_sample / splat_grads_sample above).
// xy: n x 2 int32 pixel coordinates; out / condimage: n x 3.
int orc_splat_pixels_f64(const double* params, int N, const int* xy, int n, double* out, double* condimage, int threads) {
    return splat_pixels_sample<double>(params, N, xy, n, out, condimage, threads);
}
// ids: m Gaussian indices; target / image: full P x 3; grads / absgrads / kinkgrads: m x 9 (+=).
int orc_splat_grads_sample_f64(const double* params, const int* ids, int m, const double* target, const double* image,
                               int W, int H, double* grads, double* absgrads, double* kinkgrads, int threads) {
    return splat_grads_sample<double>(params, ids, m, target, image, W, H, grads, absgrads, kinkgrads, threads);
}

// Least squares: examples/optimization/tests/test_linear_regression_gradient.cu:52-78 (squared loss); residual_only = 1:
// the same graph differentiated at the residual; residual_only = 2: the graph AND root the shipped example runs
// (linear_regression_sgd.cu:103-122).
// params = {value[4], grad[4]}; grad += ; *loss_sum += root values.
int orc_lsq_grad_f64(const double* data, long long n, double* params, double* loss_sum, int residual_only, int threads) {
    threads = std::max(1, threads);
    std::vector<double> priv(static_cast<size_t>(threads) * 5, 0.0);
    const double a = params[0], b = params[1], c = params[2], d = params[3];
    parallel_ranges(n, threads, [&](int t, long long lo, long long hi) {
        double* g = &priv[static_cast<size_t>(t) * 5];
        for (long long i = lo; i < hi; ++i) {
            const double x1 = data[3 * i], x2 = data[3 * i + 1], yt = data[3 * i + 2];
            const double u = a - x1;            // sub_constant(a, x1)
            const double u2 = u * u;            // squared
            const double v = c - x2;            // sub_constant(c, x2)
            const double v2 = v * v;            // squared
            const double bv2 = b * v2;          // mul(b, v2)
            if (residual_only == 2) {
                // the graph the SHIPPED example builds (linear_regression_sgd.cu:103-122): combined_terms = x1_term +
                // x2_term with the UN-squared x1_term = a - x1 (x1_term2 is dead), root = the residual (loss.run())
                const double r2 = ((u + bv2) + d) - yt;
                g[4] += r2;
                g[0] += 1.0;                    // sub_constant backward: a += seed
                g[1] += 1.0 * v2;
                g[2] += 1.0 * b * 2.0 * v;
                g[3] += 1.0;
                continue;
            }
            const double comb = u2 + bv2;       // add
            const double ypred = comb + d;      // add(combined, d)
            const double r = ypred - yt;        // sub_constant(y_pred, y)
            double seed;                        // adjoint arriving at y_pred
            if (residual_only) {
                seed = 1.0;
                g[4] += r;
            } else {
                seed = 1.0 * 2.0 * r;           // squared backward: g * 2.0 * x
                g[4] += r * r;
            }
            // add(combined, d): combined += seed ; d += seed.  add(x1_term, x2_term): both += seed.
            // x1_term = squared(u): u += seed*2.0*u -> a.  x2_term = mul(b, v2): b += seed*v2 ; v2 += seed*b.
            // v2 = squared(v): v += (seed*b)*2.0*v -> c.
            // leaf accumulation order inside one run(): a, b, c, d each get exactly one add.
            g[0] += seed * 2.0 * u;
            g[1] += seed * v2;
            g[2] += seed * b * 2.0 * v;
            g[3] += seed;
        }
    });
    for (int t = 0; t < threads; ++t) {
        for (int k = 0; k < 4; ++k) params[4 + k] += priv[static_cast<size_t>(t) * 5 + k];
        if (loss_sum) *loss_sum += priv[static_cast<size_t>(t) * 5 + 4];
    }
    return 0;
}

// update_parameters_kernel, linear_regression_sgd.cu:126-134
int orc_lsq_sgd_update_f64(double* params, double lr, long long batch) {
    for (int i = 0; i < 4; ++i) params[i] -= lr * params[4 + i] / static_cast<double>(batch);
    return 0;
}

// Counter-based replacement of select_batch_kernel (linear_regression_sgd.cu:68-81): THIS repo's
// definition (splitmix64 of seed, epoch, slot), restated from csrc/lsq_kernels.cu.
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
int orc_lsq_select_batch(const double* data, long long n_total, double* batch, long long batch_size, uint64_t seed,
                         uint64_t epoch) {
    for (long long i = 0; i < batch_size; ++i) {
        const uint64_t h = splitmix64(splitmix64(seed ^ (epoch * 0xD1B54A32D192ED03ull)) + static_cast<uint64_t>(i));
        const long long j = static_cast<long long>(h % static_cast<uint64_t>(n_total));
        for (int k = 0; k < 3; ++k) batch[3 * i + k] = data[3 * j + k];
    }
    return 0;
}

// Accumulation: VariableRef::add_grad (variable.cuh:48-50) in element order
// (tests/test_parallel_gradient_accumulation.cu:32-43).
int orc_accumulate_f32(const int* idx, const float* val, long long n, float* grad, int k, int threads) {
    threads = std::max(1, threads);
    std::vector<std::vector<float>> priv(threads);
    parallel_ranges(n, threads, [&](int t, long long lo, long long hi) {
        priv[t].assign(k, 0.f);
        for (long long i = lo; i < hi; ++i) priv[t][idx ? idx[i] : static_cast<int>(i % k)] += val[i];
    });
    for (int t = 0; t < threads; ++t)
        if (!priv[t].empty())
            for (int j = 0; j < k; ++j) grad[j] += priv[t][j];
    return 0;
}
int orc_accumulate_f64(const int* idx, const double* val, long long n, double* grad, int k, int threads) {
    threads = std::max(1, threads);
    std::vector<std::vector<double>> priv(threads);
    parallel_ranges(n, threads, [&](int t, long long lo, long long hi) {
        priv[t].assign(k, 0.0);
        for (long long i = lo; i < hi; ++i) priv[t][idx ? idx[i] : static_cast<int>(i % k)] += val[i];
    });
    for (int t = 0; t < threads; ++t)
        if (!priv[t].empty())
            for (int j = 0; j < k; ++j) grad[j] += priv[t][j];
    return 0;
}
// Exact (fp64-accumulated) sums of fp32 values: the ground truth the 1e-4 tolerance is stated against.
int orc_accumulate_f32_exact(const int* idx, const float* val, long long n, double* grad, int k) {
    for (long long i = 0; i < n; ++i) grad[idx ? idx[i] : static_cast<int>(i % k)] += static_cast<double>(val[i]);
    return 0;
}

int orc_covproj_f32(const float* J, const float* W, const float* S, const float* g, float* out, float* gJ, float* gW,
                    float* gS, long long n, int threads) {
    parallel_ranges(n, threads, [&](int, long long lo, long long hi) {
        for (long long e = lo; e < hi; ++e)
            covproj_one<float>(J + 6 * e, W + 9 * e, S + 6 * e, g + 3 * e, out + 3 * e, gJ + 6 * e, gW + 9 * e, gS + 6 * e);
    });
    return 0;
}
int orc_covproj_f64(const double* J, const double* W, const double* S, const double* g, double* out, double* gJ,
                    double* gW, double* gS, long long n, int threads) {
    parallel_ranges(n, threads, [&](int, long long lo, long long hi) {
        for (long long e = lo; e < hi; ++e)
            covproj_one<double>(J + 6 * e, W + 9 * e, S + 6 * e, g + 3 * e, out + 3 * e, gJ + 6 * e, gW + 9 * e, gS + 6 * e);
    });
    return 0;
}

// zero_gradients_kernel, gaussian_parameters.cu:227-241
int orc_zero_gradients(float* grads, int n) {
    std::memset(grads, 0, static_cast<size_t>(n) * 9 * sizeof(float));
    return 0;
}

// adam_step_individual_kernel, gaussian_parameters.cu:260-320 with the host wrapper's
// beta^t = powf(beta, iteration) (:357-358).  lr = {center, scale, rotation, color, opacity}.
// AdamState layout (gaussian_parameters.h:21-32): per group m[k] then v[k].
int orc_adam_step_individual(float* params, const float* grads, float* adam, int n, const float* lr, float beta1,
                             float beta2, float eps, int iteration) {
    const float b1t = std::pow(beta1, static_cast<float>(iteration));
    const float b2t = std::pow(beta2, static_cast<float>(iteration));
    static const int group_off[5] = {0, 2, 4, 5, 8}, group_len[5] = {2, 2, 1, 3, 1}, adam_off[5] = {0, 4, 8, 10, 16};
    for (int i = 0; i < n; ++i) {
        float* p = params + 9 * i;
        const float* g = grads + 9 * i;
        float* st = adam + 18 * i;
        for (int grp = 0; grp < 5; ++grp) {
            const float lrc = lr[grp] * std::sqrt(1.0f - b2t) / (1.0f - b1t);
            for (int j = 0; j < group_len[grp]; ++j) {
                float& m = st[adam_off[grp] + j];
                float& v = st[adam_off[grp] + group_len[grp] + j];
                const float gr = g[group_off[grp] + j];
                m = beta1 * m + (1.0f - beta1) * gr;
                v = beta2 * v + (1.0f - beta2) * gr * gr;
                p[group_off[grp] + j] -= lrc * m / (std::sqrt(v) + eps);
            }
        }
    }
    return 0;
}

// Per-Gaussian records as the CUDA preprocess kernel defines them (csrc/splat_common.cuh):
// rec[12] = {cx, cy, ia, ib, ic, sigmoid(opacity), r, g, b, 0, 0, 0}, fp32, IEEE ops, libm
// transcendentals (the GPU's differ in the last ulp: compare records with a tolerance, and feed
// the GPU's own records to orc_splat_binning for the bit-exact integer comparison).
int orc_splat_records(const float* params, int N, float* rec) {
    for (int g = 0; g < N; ++g) {
        const float* gp = params + 9 * g;
        float es[2] = {std::exp(gp[2]), std::exp(gp[3])}, cov[3], inv[3];
        scale_rot_cov_fwd(es, gp[4], cov);
        sym_inv_fwd(cov, inv);
        float* r = rec + 12 * g;
        r[0] = gp[0]; r[1] = gp[1]; r[2] = inv[0]; r[3] = inv[1]; r[4] = inv[2];
        r[5] = m_sigmoid(gp[8]);
        r[6] = gp[5]; r[7] = gp[6]; r[8] = gp[7];
        r[9] = r[10] = r[11] = 0.f;
    }
    return 0;
}

// Integer work: rectangles -> (tile, gaussian) keys in Gaussian order -> stable sort by tile ->
// per-tile [begin, end).  Returns the number of list entries; sorted_ids may be NULL (count only).
long long orc_splat_binning(const float* rec, int N, int W, int H, int row_begin, int row_end, float d2max,
                            int no_cull, int32_t* rects, int32_t* tile_ranges, int32_t* sorted_ids, long long capacity) {
    const int tiles_x = (W + kTile - 1) / kTile, tiles_y = (H + kTile - 1) / kTile;
    const int ntiles = tiles_x * tiles_y;
    std::vector<long long> count(ntiles + 1, 0);
    std::vector<int32_t> rl(static_cast<size_t>(N) * 4);
    for (int g = 0; g < N; ++g) {
        tile_rect(rec + 12 * g, W, H, row_begin, row_end, d2max, no_cull, &rl[4 * g]);
        const SpanCoef sc = span_coef(rec + 12 * g, d2max, no_cull);
        for (int ty = rl[4 * g + 1]; ty < rl[4 * g + 3]; ++ty) {
            int s0, s1;
            tile_row_span(sc, &rl[4 * g], ty, W, row_begin, row_end, &s0, &s1);
            for (int tx = s0; tx < s1; ++tx) count[ty * tiles_x + tx + 1]++;
        }
    }
    if (rects) std::memcpy(rects, rl.data(), rl.size() * sizeof(int32_t));
    for (int t = 0; t < ntiles; ++t) count[t + 1] += count[t];
    const long long total = count[ntiles];
    if (tile_ranges)
        for (int t = 0; t < ntiles; ++t) {
            tile_ranges[2 * t] = static_cast<int32_t>(count[t]);
            tile_ranges[2 * t + 1] = static_cast<int32_t>(count[t + 1]);
        }
    if (sorted_ids && total <= capacity) {
        std::vector<long long> cur(count.begin(), count.end() - 1);
        for (int g = 0; g < N; ++g) {  // ascending g => each tile list is ascending (stable)
            const SpanCoef sc = span_coef(rec + 12 * g, d2max, no_cull);
            for (int ty = rl[4 * g + 1]; ty < rl[4 * g + 3]; ++ty) {
                int s0, s1;
                tile_row_span(sc, &rl[4 * g], ty, W, row_begin, row_end, &s0, &s1);
                for (int tx = s0; tx < s1; ++tx) sorted_ids[cur[ty * tiles_x + tx]++] = g;
            }
        }
    }
    return total;
}


// Sequential emulation of THIS repo's counting-sort binning (xyz-autodiff-cuda_b200/csrc/splat_host.cu section 2b):
// same chunking, per-chunk histograms, column scan, tile scan, and the scatter kernel's ownership rule -- warp w of a
// chunk's CTA owns a band of tile rows, lane (rl, xl) of the warp owns the tiles (S0 + rl, x = xl mod XP), Gaussians in
// groups of 32 candidates, hits in sub-groups of 8 -- with every index expression transcribed from the kernel.  Lanes
// and warps run one after the other here; since a tile has exactly one owner lane, any interleaving gives the same
// lists.  Purpose: the index arithmetic can be swept over image shapes / row bands / chunk counts on the CPU
// (tests/test_oracle.py); the CUDA kernels themselves are held to orc_splat_binning on the GPU.
// Outputs: tile_ranges (tiles x 2), sorted_ids (capacity), backward work records chunk_info (chunk_info_size x 4: the
// binning only marks the records beyond those set aside for the tiles with -1).
// Returns the list length, or -1 if an owner rule is violated (a slot written twice / out of its tile's range).
long long orc_splat_binning_counting(const float* rec, int N, int W, int H, int row_begin, int row_end, float d2max,
                                     int no_cull, int ctas_total, int bwd_chunk, int32_t* tile_ranges,
                                     int32_t* sorted_ids, long long capacity, int32_t* chunk_info, int chunk_info_size) {
    constexpr int kSpanRowsE = 16, kBatch = 32, kWarps = 8;
    const int tiles_x = (W + kTile - 1) / kTile, tiles_y = (H + kTile - 1) / kTile, n_tiles = tiles_x * tiles_y;
    const int ng = N > 0 ? N : 1;
    const int chunk_size = ((ng + ctas_total - 1) / ctas_total + kBatch - 1) / kBatch * kBatch;
    const int n_chunks = (ng + chunk_size - 1) / chunk_size;
    // preprocess: rectangles, stored spans of the first 16 rows, tiles per Gaussian, per-chunk histogram
    std::vector<int32_t> rects(static_cast<size_t>(ng) * 4, 0), spans(static_cast<size_t>(ng) * kSpanRowsE * 2, 0);
    std::vector<unsigned int> touched(ng, 0), hist(static_cast<size_t>(n_chunks) * n_tiles, 0);
    for (int g = 0; g < N; ++g) {
        int32_t* r = &rects[4 * g];
        tile_rect(rec + 12 * g, W, H, row_begin, row_end, d2max, no_cull, r);
        const SpanCoef sc = span_coef(rec + 12 * g, d2max, no_cull);
        unsigned int cnt = 0;
        for (int ty = r[1]; ty < r[3]; ++ty) {
            int s0, s1;
            tile_row_span(sc, r, ty, W, row_begin, row_end, &s0, &s1);
            cnt += static_cast<unsigned int>(std::max(s1 - s0, 0));
            if (ty - r[1] < kSpanRowsE) {
                spans[(static_cast<size_t>(g) * kSpanRowsE + (ty - r[1])) * 2] = s0;
                spans[(static_cast<size_t>(g) * kSpanRowsE + (ty - r[1])) * 2 + 1] = std::max(s1, s0);
            }
            for (int tx = s0; tx < s1; ++tx) hist[static_cast<size_t>(g / chunk_size) * n_tiles + ty * tiles_x + tx]++;
        }
        touched[g] = cnt;
    }
    // column scan (exclusive prefix over the chunks, per tile) + tile scan
    std::vector<unsigned int> tile_total(n_tiles, 0);
    for (int t = 0; t < n_tiles; ++t) {
        unsigned int run = 0;
        for (int c = 0; c < n_chunks; ++c) {
            const unsigned int x = hist[static_cast<size_t>(c) * n_tiles + t];
            hist[static_cast<size_t>(c) * n_tiles + t] = run;
            run += x;
        }
        tile_total[t] = run;
    }
    std::vector<long long> begin(n_tiles + 1, 0);
    std::vector<int> chunk_offsets(n_tiles + 1, 0);
    for (int t = 0; t < n_tiles; ++t) {
        begin[t + 1] = begin[t] + tile_total[t];
        // backward work records set aside per tile: ceil(len / chunk) (the forward pass fills them in)
        chunk_offsets[t + 1] = chunk_offsets[t] + static_cast<int>((tile_total[t] + bwd_chunk - 1) / bwd_chunk);
        if (tile_ranges) {
            tile_ranges[2 * t] = static_cast<int32_t>(begin[t]);
            tile_ranges[2 * t + 1] = static_cast<int32_t>(begin[t + 1]);
        }
    }
    const long long total = begin[n_tiles];
    if (!sorted_ids || total > capacity) return total;
    std::vector<char> written(static_cast<size_t>(total > 0 ? total : 1), 0);
    bool ok = true;
    // scatter
    for (int block = 0; block < n_chunks; ++block) {
        const int g_begin = block * chunk_size, g_end = std::min(g_begin + chunk_size, N);
        if (chunk_info) {  // the work records beyond the tiles' own: none (distributed over the CTAs and threads like in the kernel)
            for (int tid = 0; tid < 256; ++tid) {
                for (int c = chunk_offsets[n_tiles] + block * 256 + tid; c < chunk_info_size; c += n_chunks * 256) {
                    int32_t* ci = chunk_info + 4 * static_cast<size_t>(c);
                    ci[0] = ci[1] = ci[2] = ci[3] = -1;
                }
            }
        }
        std::vector<unsigned int> s_next(n_tiles);
        for (int t = 0; t < n_tiles; ++t) s_next[t] = static_cast<unsigned int>(begin[t]) + hist[static_cast<size_t>(block) * n_tiles + t];
        const int ty_lo = row_begin / kTile, ty_hi = (row_end + kTile - 1) / kTile;
        const int band = (ty_hi - ty_lo + kWarps - 1) / kWarps;
        for (int warp = 0; warp < kWarps; ++warp) {
            const int R0 = ty_lo + warp * band, R1 = std::min(R0 + band, ty_hi);
            if (R0 >= R1) continue;
            const int rl_bits = band >= 8 ? 3 : (band >= 4 ? 2 : (band >= 2 ? 1 : 0));
            const int RL = 1 << rl_bits, XP = 32 >> rl_bits;
            for (int S0 = R0; S0 < R1; S0 += RL) {
                const int S1 = std::min(S0 + RL, R1);
                for (int gb = g_begin; gb < g_end; gb += 32) {
                    int ry[32], rw[32];
                    int hit_lane[32];
                    int n_hits = 0;
                    for (int lane = 0; lane < 32; ++lane) {
                        const int g = gb + lane;
                        ry[lane] = rw[lane] = 0;
                        if (g < g_end && touched[g] != 0u) { ry[lane] = rects[4 * g + 1]; rw[lane] = rects[4 * g + 3]; }
                        if (ry[lane] < S1 && rw[lane] > S0 && rw[lane] > ry[lane]) hit_lane[n_hits++] = lane;  // rank -> lane
                    }
                    for (int base = 0; base < n_hits; base += 8) {
                        const int n_sub = std::min(8, n_hits - base);
                        int st_x[32][2], st_y[32][2];  // what every lane staged
                        for (int lane = 0; lane < 32; ++lane) {
                            const int rl = lane >> (5 - rl_bits), ty = S0 + rl;
                            for (int q = 0; q < 2; ++q) {
                                st_x[lane][q] = st_y[lane][q] = 0;
                                const int h = 4 * q + (lane & 3);
                                if (h >= n_sub) continue;
                                const int src = hit_lane[base + h];
                                if (!(ty < S1 && ty >= ry[src] && ty < rw[src])) continue;
                                const int gg = gb + src, j = ty - ry[src];
                                int s0, s1;
                                if (j < kSpanRowsE) {
                                    s0 = spans[(static_cast<size_t>(gg) * kSpanRowsE + j) * 2];
                                    s1 = spans[(static_cast<size_t>(gg) * kSpanRowsE + j) * 2 + 1];
                                } else {
                                    const SpanCoef sc = span_coef(rec + 12 * gg, d2max, no_cull);
                                    tile_row_span(sc, &rects[4 * gg], ty, W, row_begin, row_end, &s0, &s1);
                                }
                                st_x[lane][q] = s0;
                                st_y[lane][q] = s1;
                            }
                        }
                        for (int i = 0; i < n_sub; ++i) {
                            const int gg = gb + hit_lane[base + i];
                            for (int lane = 0; lane < 32; ++lane) {
                                const int rl = lane >> (5 - rl_bits), xl = lane & (XP - 1), ty = S0 + rl;
                                const int from = (rl << (5 - rl_bits)) | (i & 3);
                                const int sx = st_x[from][i >> 2], sy = st_y[from][i >> 2];
                                for (int x = sx + ((xl - sx) & (XP - 1)); x < sy; x += XP) {
                                    const int tile = ty * tiles_x + x;
                                    const unsigned int pos = s_next[tile]++;
                                    if (pos >= static_cast<unsigned int>(begin[tile + 1]) || pos < static_cast<unsigned int>(begin[tile]) || written[pos]) {
                                        ok = false;
                                    } else {
                                        written[pos] = 1;
                                        sorted_ids[pos] = gg;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    for (long long e = 0; e < total; ++e) ok = ok && written[e];
    return ok ? total : -1;
}

}  // extern "C"
