// xyz_autodiff/symmetric_matrix_view.cuh -- an N(N+1)/2-vector variable seen as a symmetric N x N
// matrix; storage is the upper triangle, row by row: (0,0) (0,1) ... (0,N-1) (1,1) ... (N-1,N-1).
// Contract of reference include/xyz_autodiff/symmetric_matrix_view.cuh:11-107 (index rule :24-29).
#pragma once

#include <type_traits>

#include "concept/matrix.cuh"
#include "concept/variable.cuh"
#include "detail/config.cuh"

namespace xyz_autodiff {

template <typename T, std::size_t N, DifferentiableVariableConcept Storage>
class SymmetricMatrixView {
public:
    using value_type = T;
    static constexpr std::size_t rows = N;
    static constexpr std::size_t cols = N;
    static constexpr std::size_t size = N * N;
    static constexpr std::size_t storage_size = N * (N + 1) / 2;

    XYZ_HD explicit SymmetricMatrixView(Storage& packed) : packed_(packed) {
        static_assert(Storage::size == storage_size, "storage must hold N(N+1)/2 entries");
        static_assert(std::is_same_v<typename Storage::value_type, T>, "value types must agree");
    }
    SymmetricMatrixView(const SymmetricMatrixView&) = default;

    // packed position of (r, c): rows above r hold N + (N-1) + ... entries
    XYZ_HD static constexpr std::size_t packed_index(std::size_t r, std::size_t c) {
        const std::size_t lo = r <= c ? r : c, hi = r <= c ? c : r;
        return lo * N - lo * (lo - 1) / 2 + (hi - lo);
    }

    XYZ_HD T& operator()(std::size_t r, std::size_t c) { return packed_[packed_index(r, c)]; }
    XYZ_HD const T& operator()(std::size_t r, std::size_t c) const { return packed_[packed_index(r, c)]; }
    XYZ_HD SymmetricMatrixView& transpose() { return *this; }
    XYZ_HD const SymmetricMatrixView& transpose() const { return *this; }

    // variable interface over the packed storage
    XYZ_HD T* data() { return packed_.data(); }
    XYZ_HD const T* data() const { return packed_.data(); }
    XYZ_HD T* grad() { return packed_.grad(); }
    XYZ_HD const T* grad() const { return packed_.grad(); }
    XYZ_HD T& operator[](std::size_t i) { return packed_[i]; }
    XYZ_HD const T& operator[](std::size_t i) const { return packed_[i]; }
    XYZ_HD const T& grad(std::size_t i) const { return packed_.grad(i); }
    XYZ_HD void add_grad(std::size_t i, T value) { packed_.add_grad(i, value); }
    XYZ_HD void zero_grad() { packed_.zero_grad(); }

    XYZ_HD auto& underlying_variable() { return packed_; }
    XYZ_HD const auto& underlying_variable() const { return packed_; }

private:
    Storage& packed_;
};

template <std::size_t N, DifferentiableVariableConcept Storage>
XYZ_HD auto make_symmetric_matrix_view(Storage& packed) {
    return SymmetricMatrixView<typename Storage::value_type, N, Storage>(packed);
}

}  // namespace xyz_autodiff
