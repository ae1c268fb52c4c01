// tests/csrc/api_probe_host.cpp -- TEST INFRASTRUCTURE: THIS repo's include/xyz_autodiff compiled by a
// plain host C++20 compiler (no CUDA, no shim) and driven by tests/csrc/api_eval.inc -- the same text
// oracle/ref_driver.cpp compiles against the REFERENCE headers.  Exports mine_* with the ref_* signatures.
#include "api_headers.inc"

#define API_FN inline
#include "api_eval.inc"

// op::covariance_projection (the ternary node of operations/ternary/covariance_projection_logic.cuh) for one element:
// out3 = packed S', adjoints of J (6), W (9), S (6) given the adjoint g3 of S'
template <typename T>
static int covproj_ternary(const T* J, const T* W, const T* S, const T* g, T* out, T* gJ, T* gW, T* gS) {
    using namespace xyz_autodiff;
    Variable<6, T> vJ(J), vS(S);
    Variable<9, T> vW(W);
    auto P = op::covariance_projection(vJ, vW, vS);
    static_assert(OperationNode<decltype(P)> && DifferentiableVariableConcept<decltype(P)>);
    P.forward();
    for (int i = 0; i < 3; ++i) out[i] = P[i];
    P.zero_grad();
    for (int i = 0; i < 3; ++i) P.add_grad(i, g[i]);
    P.backward();
    for (int i = 0; i < 6; ++i) gJ[i] = vJ.grad(i);
    for (int i = 0; i < 9; ++i) gW[i] = vW.grad(i);
    for (int i = 0; i < 6; ++i) gS[i] = vS.grad(i);
    // run_numerical on a second copy: central differences of the same node, summed over the three outputs
    return 0;
}

extern "C" {
int mine_eval_op_f64(int op, int aux, const double* in1, int n1, const double* in2, int n2, double cst,
                     const double* gout, double* out, int* nout, double* gin1, double* gin2) {
    return api_eval::eval_op<double>(op, aux, in1, n1, in2, n2, cst, gout, out, nout, gin1, gin2);
}
int mine_eval_op_f32(int op, int aux, const float* in1, int n1, const float* in2, int n2, float cst, const float* gout,
                     float* out, int* nout, float* gin1, float* gin2) {
    return api_eval::eval_op<float>(op, aux, in1, n1, in2, n2, cst, gout, out, nout, gin1, gin2);
}
int mine_kat_dag(double* res) { double scratch[8]; return api_eval::kat_dag(res, scratch); }
int mine_kat_shared_subgraph(double* res) { return api_eval::kat_shared_subgraph(res); }
int mine_kat_broadcast(double* res) { return api_eval::kat_broadcast(res); }
int mine_kat_broadcast_logic(double x0, const double* up4, double s0, const double* v3, double* res) {
    return api_eval::kat_broadcast_logic(x0, up4, s0, v3, res);
}
int mine_kat_chain(double x, double y, double z, double up, double* res) { return api_eval::kat_chain(x, y, z, up, res); }
int mine_kat_operators(const double* a, const double* b, const double* c, const double* d, double* res) {
    return api_eval::kat_operators(a, b, c, d, res);
}
int mine_kat_lsq_point(const double* p, double x1, double x2, double yt, double delta, double* res) {
    double scratch[8];
    return api_eval::kat_lsq_point(p, x1, x2, yt, delta, res, scratch);
}
int mine_kat_splat_pair(const double* in, double* res) { return api_eval::kat_splat_pair(in, res); }
int mine_kat_math_f64(double x, double* res) { return api_eval::kat_math<double>(x, res); }
int mine_kat_math_f32(float x, float* res) { return api_eval::kat_math<float>(x, res); }
int mine_kat_matrices(float* res) { return api_eval::kat_matrices(res); }
int mine_kat_const_array(float* res) { return api_eval::kat_const_array(res); }
int mine_kat_matrices3(float* res) { float scratch[12]; return api_eval::kat_matrices3(res, scratch); }
int mine_kat_networks(const double* p, double delta, double* res) {
    double scratch[8];
    return api_eval::kat_networks(p, delta, res, scratch);
}
int mine_kat_matrices2(float* res) { return api_eval::kat_matrices2(res); }
int mine_kat_variable(double* res) { double scratch[8]; return api_eval::kat_variable(res, scratch); }

int mine_covproj_ternary_f64(const double* J, const double* W, const double* S, const double* g, double* out, double* gJ,
                             double* gW, double* gS) {
    return covproj_ternary<double>(J, W, S, g, out, gJ, gW, gS);
}
int mine_covproj_ternary_f32(const float* J, const float* W, const float* S, const float* g, float* out, float* gJ, float* gW,
                             float* gS) {
    return covproj_ternary<float>(J, W, S, g, out, gJ, gW, gS);
}
// the node's analytic adjoints (all outputs seeded with 1, run()) against its own central differences (run_numerical)
int mine_covproj_ternary_numerical(const double* J, const double* W, const double* S, double* res42) {
    using namespace xyz_autodiff;
    for (int pass = 0; pass < 2; ++pass) {
        Variable<6, double> vJ(J), vS(S);
        Variable<9, double> vW(W);
        auto P = op::covariance_projection(vJ, vW, vS);
        if (pass == 0) P.run(); else P.run_numerical(1e-6);
        double* r = res42 + 21 * pass;
        for (int i = 0; i < 6; ++i) r[i] = vJ.grad(i);
        for (int i = 0; i < 9; ++i) r[6 + i] = vW.grad(i);
        for (int i = 0; i < 6; ++i) r[15 + i] = vS.grad(i);
    }
    return 42;
}

// accum::slice: the least-squares graph on four one-component views of ONE RegisterLeaf<4> (what batched::for_each
// hands to a user graph); res = {loss, da, db, dc, dd} accumulated over `reps` evaluations of the same point
int mine_kat_leaf_slices(const double* p, double x1, double x2, double yt, int reps, double* res) {
    using namespace xyz_autodiff;
    accum::RegisterLeaf<4, double> shared(p);
    static_assert(DifferentiableVariableConcept<accum::LeafSlice<1, 2, accum::RegisterLeaf<4, double>>>);
    double loss_value = 0.0;
    for (int r = 0; r < reps; ++r) {
        auto a = accum::slice<0, 1>(shared);
        auto b = accum::slice<1, 1>(shared);
        auto c = accum::slice<2, 1>(shared);
        auto d = accum::slice<3, 1>(shared);
        auto x1_minus_a = op::sub_constant(a, x1);
        auto x1_term = op::squared(x1_minus_a);
        auto x2_minus_c = op::sub_constant(c, x2);
        auto x2_squared = op::squared(x2_minus_c);
        auto x2_term = op::mul(b, x2_squared);
        auto combined = op::add(x1_term, x2_term);
        auto y_pred = op::add(combined, d);
        auto y_diff = op::sub_constant(y_pred, yt);
        auto loss = op::squared(y_diff);
        loss.run();
        loss_value = loss[0];
    }
    res[0] = loss_value;
    for (int i = 0; i < 4; ++i) res[1 + i] = shared.grad(i);
    auto bc = accum::slice<1, 2>(shared);
    bc.zero_grad();                      // zeroes components 1 and 2 of the parent only
    for (int i = 0; i < 4; ++i) res[5 + i] = shared.grad(i);
    return 9;
}

// ternary node + RegisterLeaf on the host: y = a*b + c ; returns {y, da, db, dc}
int mine_kat_ternary(double a0, double b0, double c0, double* res) {
    using namespace xyz_autodiff;
    Variable<1, double> a(a0), b(b0), c(c0);
    api_static_checks::Tern node(api_static_checks::TernaryProbeLogic{}, a, b, c);
    node.run();
    res[0] = node[0];
    res[1] = a.grad(0);
    res[2] = b.grad(0);
    res[3] = c.grad(0);
    // same graph evaluated numerically
    Variable<1, double> a2(a0), b2(b0), c2(c0);
    api_static_checks::Tern node2(api_static_checks::TernaryProbeLogic{}, a2, b2, c2);
    node2.run_numerical(1e-6);
    res[4] = a2.grad(0);
    res[5] = b2.grad(0);
    res[6] = c2.grad(0);
    return 7;
}
}
