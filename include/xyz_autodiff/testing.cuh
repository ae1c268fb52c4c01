// xyz_autodiff/testing.cuh -- gradient-verification tools for user-written Logics and networks
// (reference include/xyz_autodiff/testing.cuh:1-8).  Unlike the reference's, these need no gtest (they return a
// GradientReport; with gtest included first they also raise ADD_FAILURE) and run every random case in ONE launch.
#pragma once

#include "testing/unary_gradient_tester.cuh"
#include "testing/binary_gradient_tester.cuh"
#include "testing/network_gradient_tester.cuh"
