// tests/csrc/tester_probe.cu -- TEST INFRASTRUCTURE: drives the gtest-free gradient testers of
// include/xyz_autodiff/testing.cuh the way the reference's tests/operation/**/test_*_gradient.cu files drive theirs
// (one tester per Logic), plus a deliberately wrong Logic, the forbidden-tolerance rule and a network functor in
// the style of tests/operation/base/test_linear_regression_network.cu:41-82.
#include "api_headers.inc"

#include <xyz_autodiff/testing.cuh>

#include <cstdio>
#include <cstring>
#include <sstream>

using namespace xyz_autodiff;
namespace xt = xyz_autodiff::testing;

namespace {

template <std::size_t N>
using Ref = VariableRef<N, double>;

// wrong on purpose: forward x^2, backward claims d/dx = 3x
template <std::size_t Dim>
struct WrongSquareLogic {
    static constexpr std::size_t outputDim = Dim;
    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        for (std::size_t i = 0; i < Dim; ++i) y[i] = x[i] * x[i];
    }
    template <typename Output, typename Input>
    XYZ_HD void backward(const Output& y, Input& x) const {
        for (std::size_t i = 0; i < Dim; ++i) x.add_grad(i, y.grad(i) * 3.0 * x[i]);
    }
};

struct LsqParams {
    double a, b, c, d, x1, x2, y_target;
};

// (a - x1)^2 + b (c - x2)^2 + d - y, squared: the reference's regression network
struct LsqNetwork {
    template <xt::GradientTag tag>
    __device__ void operator()(LsqParams* value, LsqParams* diff, double delta) const {
        VariableRef<1, double> a(&value->a, &diff->a), b(&value->b, &diff->b), c(&value->c, &diff->c), d(&value->d, &diff->d);
        auto u = a - value->x1;
        auto u2 = op::squared(u);
        auto v = c - value->x2;
        auto v2 = op::squared(v);
        auto bv2 = b * v2;
        auto s = u2 + bv2;
        auto pred = s + d;
        auto r = pred - value->y_target;
        auto loss = op::squared(r);
        if constexpr (tag == xt::GradientTag::Analytical) loss.run(); else loss.run_numerical(delta);
    }
};

struct Tally {
    int failed = 0, run = 0;
    std::ostringstream log;
    void add(const xt::GradientReport& r, bool expect_pass = true) {
        ++run;
        const bool ok = (r.passed() == expect_pass);
        if (!ok) ++failed;
        log << (ok ? "ok   " : "FAIL ") << r.name << ": " << r.num_tests << " cases, failures " << r.num_failures
            << ", max error " << r.max_error << (r.message.empty() ? "" : " | ") << r.message << "\n";
    }
};

}  // namespace

extern "C" int tester_run_all(char* text, int cap) {
    Tally t;
    // unary Logics (tests/operation/unary/*_gradient.cu)
    t.add(xt::UnaryGradientTester<op::ExpLogic<3>, 3, 3>::test("ExpLogic<3>"));
    t.add(xt::UnaryGradientTester<op::SinLogic<3>, 3, 3>::test("SinLogic<3>"));
    t.add(xt::UnaryGradientTester<op::CosLogic<3>, 3, 3>::test("CosLogic<3>"));
    t.add(xt::UnaryGradientTester<op::SigmoidLogic<4>, 4, 4>::test("SigmoidLogic<4>"));
    t.add(xt::UnaryGradientTester<op::SquaredLogic<2>, 2, 2>::test("SquaredLogic<2>"));
    t.add(xt::UnaryGradientTester<op::NegLogic<5>, 5, 5>::test("NegLogic<5>"));
    t.add(xt::UnaryGradientTester<op::L1NormLogic<4>, 4, 1>::test("L1NormLogic<4>"));
    t.add(xt::UnaryGradientTester<op::L2NormLogic<4>, 4, 1>::test("L2NormLogic<4>"));
    t.add(xt::UnaryGradientTester<op::SumLogic<6>, 6, 1>::test("SumLogic<6>"));
    t.add(xt::UnaryGradientTester<op::QuaternionToRotationMatrixLogic<4>, 4, 9>::test("QuaternionToRotationMatrixLogic"));
    // sym_matrix2_inv needs a well-conditioned input range (tests/operation/unary/..._gradient.cu use custom ranges)
    t.add(xt::UnaryGradientTester<op::SymMatrix2InvLogic<3>, 3, 3>::run("SymMatrix2InvLogic [2.5,3]x..", 200, 1e-5, 1e-6, 2.5, 3.0));
    // parameterised Logics through the `logic` argument
    t.add(xt::UnaryGradientTester<op::MulConstantLogic<Ref<3>>, 3, 3>::test_custom("MulConstantLogic(2.5)", 50, 1e-5, 1e-5, -2.0, 2.0,
                                                                                  op::MulConstantLogic<Ref<3>>(2.5)));
    t.add(xt::UnaryGradientTester<op::DivConstantLogic<Ref<2>>, 2, 2>::test_custom("DivConstantLogic(-0.7)", 50, 1e-5, 1e-5, -2.0, 2.0,
                                                                                  op::DivConstantLogic<Ref<2>>(-0.7)));
    // binary Logics (tests/operation/binary/*_gradient.cu)
    t.add(xt::BinaryGradientTester<op::AddLogic<Ref<3>, Ref<3>>, 3, 3, 3>::test("AddLogic"));
    t.add(xt::BinaryGradientTester<op::SubLogic<Ref<3>, Ref<3>>, 3, 3, 3>::test("SubLogic"));
    t.add(xt::BinaryGradientTester<op::MulLogic<Ref<3>, Ref<3>>, 3, 3, 3>::test("MulLogic"));
    t.add(xt::BinaryGradientTester<op::DivLogic<Ref<2>, Ref<2>>, 2, 2, 2>::run("DivLogic, |x| in [0.5, 2]", 100, 1e-5, 1e-6, 0.5, 2.0));
    t.add(xt::BinaryGradientTester<op::MatMulLogic<2, 3, 3, Ref<6>, Ref<9>>, 6, 9, 6>::test("MatMulLogic<2,3,3>"));
    t.add(xt::BinaryGradientTester<op::MatMulLogic<4, 2, 3, Ref<8>, Ref<6>>, 8, 6, 12>::test("MatMulLogic<4,2,3>"));
    // the splat example's custom Logics (examples/mini-gaussian-splatting/tests/*.cu)
    t.add(xt::BinaryGradientTester<op::ScaleRotationToCovariance3ParamLogic<Ref<2>, Ref<1>>, 2, 1, 3>::test("ScaleRotationToCovariance3Param"));
    t.add(xt::BinaryGradientTester<op::CovarianceMatrixGenerationLogic<Ref<2>, Ref<1>>, 2, 1, 4>::test("CovarianceMatrixGeneration"));
    t.add(xt::UnaryGradientTester<op::MatrixToCovariance3ParamLogic<Ref<4>>, 4, 3>::test("MatrixToCovariance3Param"));
    t.add(xt::BinaryGradientTester<op::MahalanobisDistanceLogic<Ref<2>, Ref<3>>, 2, 3, 1>::test("MahalanobisDistance"));
    // the testers must FAIL on a wrong Logic and on a forbidden tolerance (Q10)
    t.add(xt::UnaryGradientTester<WrongSquareLogic<2>, 2, 2>::test("WrongSquareLogic (must fail)"), /*expect_pass=*/false);
    t.add(xt::UnaryGradientTester<op::ExpLogic<1>, 1, 1>::test_custom("tolerance 1e-3 (forbidden, must fail)", 10, 1e-3, 1e-5), false);
    // whole-network tester
    {
        LsqParams p{0.5, 1.0, 0.3, 0.1, 2.0, 3.0, 5.0};
        auto [ok, err, msg] = xt::NetworkGradientTester<LsqParams, LsqNetwork>::test_single_case(LsqNetwork{}, p, 1e-5, 1e-7);
        xt::GradientReport r;
        r.name = "NetworkGradientTester single case";
        r.num_tests = 1;
        r.num_failures = ok ? 0 : 1;
        r.max_error = err;
        r.message = msg;
        t.add(r);
        t.add(xt::NetworkGradientTester<LsqParams, LsqNetwork>::test_random_cases(LsqNetwork{}, "NetworkGradientTester 100 random cases",
                                                                                  100, 1e-5, 1e-7, -1.5, 1.5));
    }
    const std::string s = t.log.str();
    if (text && cap > 0) {
        std::strncpy(text, s.c_str(), static_cast<size_t>(cap) - 1);
        text[cap - 1] = 0;
    }
    return t.failed * 1000 + t.run;
}
