// xyz_autodiff/detail/variable_facade.cuh -- the variable interface of a matrix VIEW, forwarded to the variable it
// wraps.  A view holds a reference, so forwarding is const-callable: constness of the view is shallow, like the
// reference's views (include/xyz_autodiff/diagonal_matrix_view.cuh:33-62, symmetric_matrix_view.cuh:60-90).
#pragma once

#include <cstddef>

#include "config.cuh"

namespace xyz_autodiff::detail {

// Derived must offer `stored()`, callable on a const view, returning the wrapped variable by mutable reference.
template <typename Derived>
class VariableFacade {
public:
    XYZ_HD decltype(auto) data() const { return wrapped().data(); }
    XYZ_HD decltype(auto) grad() const { return wrapped().grad(); }
    XYZ_HD decltype(auto) operator[](std::size_t i) const { return wrapped()[i]; }
    XYZ_HD decltype(auto) grad(std::size_t i) const { return wrapped().grad(i); }
    template <typename S>
    XYZ_HD void add_grad(std::size_t i, S term) const {
        wrapped().add_grad(i, term);
    }
    XYZ_HD void zero_grad() const { wrapped().zero_grad(); }

private:
    XYZ_HD decltype(auto) wrapped() const { return static_cast<const Derived&>(*this).stored(); }
};

}  // namespace xyz_autodiff::detail
