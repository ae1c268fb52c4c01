// xyz_autodiff/detail/pointwise.cuh -- the loop skeletons shared by every component-wise Logic.
//
// A component-wise operation is fully described by a RULE: the scalar function and its local partial
// derivative(s).  The skeletons below supply everything else once -- the fixed-trip loops (unrolled into registers),
// the order in which operand adjoints are issued (first operand, then second: the order the reference's Logics use, so
// atomically accumulated leaves see the same sequence), and the node factory.  The per-operation headers under
// operations/unary and operations/binary only name a rule and keep the reference's public spelling
// (AddLogic<Input1, Input2>, op::add, ...).
#pragma once

#include <cstddef>

#include "../operations/operation.cuh"

namespace xyz_autodiff {
namespace detail {

// y_i = f(x_i); the derivative is recomputed from the INPUT in the reverse pass (nothing is cached between passes).
// Rule: static T value(T x); static T pullback(T x, T g)  -- returns the term added to the operand's adjoint.
template <std::size_t N, typename Rule>
struct PointwiseMap {
    static constexpr std::size_t outputDim = N;

    template <typename Result, typename Operand>
    XYZ_HD void forward(Result& y, const Operand& x) const {
        using S = typename Operand::value_type;
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) {
            const S xi = x[i];
            y[i] = Rule::template value<S>(xi);
        }
    }
    template <typename Result, typename Operand>
    XYZ_HD void backward(const Result& y, Operand& x) const {
        using S = typename Operand::value_type;
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) {
            const S xi = x[i];
            const S gi = y.grad(i);
            x.add_grad(i, Rule::template pullback<S>(xi, gi));
        }
    }
};

// y_i = f(x_i, c) with a scalar constant c stored in the Logic.
// Rule: static T value(T x, T c); static T pullback(T g, T c).
template <typename Operand, typename Rule>
struct PointwiseWithScalar {
    using T = typename Operand::value_type;
    static constexpr std::size_t Dim = Operand::size;
    static constexpr std::size_t outputDim = Dim;
    using Output = Variable<Dim, T>;

    T constant_c;

    XYZ_HD explicit PointwiseWithScalar(T c) : constant_c(c) {}

    XYZ_HD void forward(Output& y, const Operand& x) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) y[i] = Rule::template value<T>(x[i], constant_c);
    }
    XYZ_HD void backward(const Output& y, Operand& x) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) x.add_grad(i, Rule::template pullback<T>(y.grad(i), constant_c));
    }
};

// y_i = f(x_i, c_i) with c any indexable constant array held BY REFERENCE (it must outlive the node); the operand's
// adjoint receives the upstream adjoint unchanged (f = x + c or x - c).
template <std::size_t N, typename Constants, typename Rule>
struct PointwiseWithArray {
    static constexpr std::size_t outputDim = N;

    const Constants& constant_array;

    XYZ_HD explicit PointwiseWithArray(const Constants& c) : constant_array(c) {}

    template <typename Result, typename Operand>
    XYZ_HD void forward(Result& y, const Operand& x) const {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) y[i] = Rule::value(x[i], constant_array[i]);
    }
    template <typename Result, typename Operand>
    XYZ_HD void backward(const Result& y, Operand& x) const {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) x.add_grad(i, y.grad(i));
    }
};

// y_i = f(a_i, b_i).  Rule: static T value(T a, T b); static void pullback(T a, T b, T g, T& to_a, T& to_b).
template <typename Left, typename Right, typename Rule>
struct PointwisePair {
    using T = typename Left::value_type;
    static constexpr std::size_t Dim = Left::size;
    static constexpr std::size_t outputDim = Dim;
    using Output = Variable<Dim, T>;

    XYZ_HD void forward(Output& y, const Left& a, const Right& b) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) y[i] = Rule::template value<T>(a[i], b[i]);
    }
    XYZ_HD void backward(const Output& y, Left& a, Right& b) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) {
            T to_a, to_b;
            Rule::template pullback<T>(a[i], b[i], y.grad(i), to_a, to_b);
            a.add_grad(i, to_a);
            b.add_grad(i, to_b);
        }
    }
};

// scalar = finish(sum_i term(x_i)).  Rule: static T term(T x); static T finish(T sum);
// static bool has_adjoint(T result) -- false where the reduction is not differentiable (no adjoint is issued at all);
// static T pullback(T x, T result, T g).
template <std::size_t N, typename Rule>
struct FoldToScalar {
    static constexpr std::size_t outputDim = 1;

    template <typename Result, typename Operand>
    XYZ_HD void forward(Result& y, const Operand& x) const {
        using S = typename Operand::value_type;
        S running = S(0);
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) running += Rule::template term<S>(x[i]);
        y[0] = Rule::template finish<S>(running);
    }
    template <typename Result, typename Operand>
    XYZ_HD void backward(const Result& y, Operand& x) const {
        using S = typename Operand::value_type;
        const S upstream = y.grad(0);
        const S result = y[0];
        if (!Rule::template has_adjoint<S>(result)) return;
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) x.add_grad(i, Rule::template pullback<S>(x[i], result, upstream));
    }
};

// node factories: the Logic's outputDim sizes the node
template <typename Logic, typename Operand, typename... LogicArgs>
XYZ_HD auto make_unary_node(Operand& x, LogicArgs&&... args) {
    return UnaryOperation<Logic::outputDim, Logic, Operand>(Logic(static_cast<LogicArgs&&>(args)...), x);
}
template <typename Logic, typename Operand1, typename Operand2>
XYZ_HD auto make_binary_node(Operand1& a, Operand2& b) {
    return BinaryOperation<Logic::outputDim, Logic, Operand1, Operand2>(Logic{}, a, b);
}

}  // namespace detail
}  // namespace xyz_autodiff
