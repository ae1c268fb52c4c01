// xyz_autodiff/variable.cuh -- the two leaf/value types of the graph.
//
//   VariableRef<N, T>  non-owning view of N values + N adjoints living in global or shared memory
//                      (or host memory); add_grad is the cross-thread accumulation point
//                      (reference include/xyz_autodiff/variable.cuh:12-59).
//   Variable<N, T>     owning, thread-private storage (registers on the device); add_grad is a
//                      plain += (reference include/xyz_autodiff/variable.cuh:62-190).
//
// Both are fully usable on the host.  For leaves whose adjoints should be pre-reduced on chip
// before they reach memory (warp shuffle -> shared memory -> one vector RED per CTA) see
// xyz_autodiff/accumulate.cuh.
#pragma once

#include "concept/variable.cuh"
#include "detail/config.cuh"

namespace xyz_autodiff {

template <std::size_t N, typename T>
    requires FloatingPointConcept<T>
class VariableRef {
public:
    using value_type = T;
    static constexpr std::size_t size = N;

    XYZ_HD constexpr VariableRef(T* values, T* adjoints) : values_(values), adjoints_(adjoints) {}

    XYZ_HD T* data() const noexcept { return values_; }
    XYZ_HD T* grad() const noexcept { return adjoints_; }
    XYZ_HD constexpr T& operator[](std::size_t i) const noexcept { return values_[i]; }
    XYZ_HD const T& grad(std::size_t i) const noexcept { return adjoints_[i]; }

    // Thread-safe: many threads may hold refs onto the same parameter.
    XYZ_HD void add_grad(std::size_t i, T value) const noexcept { detail::accumulate(adjoints_ + i, value); }

    XYZ_HD void zero_grad() const noexcept {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) adjoints_[i] = T(0);
    }

private:
    T* const values_;
    T* const adjoints_;
};

template <std::size_t N, typename T>
    requires FloatingPointConcept<T>
class Variable {
public:
    using value_type = T;
    static constexpr std::size_t size = N;

    XYZ_HD Variable() { fill(T(0)); }
    XYZ_HD Variable(const T& initial_value) { fill(initial_value); }
    XYZ_HD Variable(const T* values) {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) {
            v_[i] = values[i];
            g_[i] = T(0);
        }
    }
    Variable(const Variable&) = default;
    Variable(Variable&&) noexcept = default;
    Variable& operator=(const Variable&) = default;
    Variable& operator=(Variable&&) noexcept = default;

    XYZ_HD T* data() noexcept { return v_; }
    XYZ_HD const T* data() const noexcept { return v_; }
    XYZ_HD T* grad() noexcept { return g_; }
    XYZ_HD const T* grad() const noexcept { return g_; }
    XYZ_HD T& operator[](std::size_t i) noexcept { return v_[i]; }
    XYZ_HD const T& operator[](std::size_t i) const noexcept { return v_[i]; }
    XYZ_HD const T& grad(std::size_t i) const noexcept { return g_[i]; }

    // Thread-private storage: no atomics.
    XYZ_HD void add_grad(std::size_t i, T value) noexcept { g_[i] += value; }

    XYZ_HD void zero_grad() noexcept {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) g_[i] = T(0);
    }

    XYZ_HD VariableRef<N, T> ref() noexcept { return VariableRef<N, T>(v_, g_); }
    XYZ_HD VariableRef<N, T> ref() const noexcept {
        return VariableRef<N, T>(const_cast<T*>(v_), const_cast<T*>(g_));
    }

private:
    XYZ_HD void fill(T value) {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) {
            v_[i] = value;
            g_[i] = T(0);
        }
    }
    T v_[N];
    T g_[N];
};

}  // namespace xyz_autodiff
