// StaggeredTensorMacros.h -- the loop macros of the reference's host API (include/StaggeredTensorMacros.h:6-103), for
// drivers that fill or scan tensors point by point on the host (test/full_test.cpp:144, src/Norms.cpp).  Same macro names,
// parameters and index ranges; the loop variables are i, j, k and `CODE` / `function(args)` may use them.  Element access
// goes through the host copy of the tensor, which the host layer refreshes from the device when it is stale.
#ifndef STAGGERED_TENSOR_MACROS_H
#define STAGGERED_TENSOR_MACROS_H

#include <array>
#include <cstddef>

// All points of `tensor`, or its interior points (one layer stripped on every side) when include_border is false.
#define STAGGERED_TENSOR_ITERATE_OVER_ALL_POINTS(tensor, include_border, CODE)                  \
  {                                                                                             \
    const std::array<size_t, 3> &mif_ext_ = tensor.sizes();                                     \
    const size_t mif_lo_ = (include_border) ? 0 : 1, mif_cut_ = (include_border) ? 0 : 1;       \
    for (size_t k = mif_lo_; k + mif_cut_ < mif_ext_[2]; k++)                                   \
      for (size_t j = mif_lo_; j + mif_cut_ < mif_ext_[1]; j++)                                 \
        for (size_t i = mif_lo_; i + mif_cut_ < mif_ext_[0]; i++) {                             \
          CODE                                                                                  \
        }                                                                                       \
  }

// The points this rank owns: physical borders included, ghost layers towards neighbouring ranks and periodic ghost
// layers excluded.
#define STAGGERED_TENSOR_ITERATE_OVER_ALL_OWNER_POINTS(tensor, CODE)                                                   \
  {                                                                                                                    \
    const std::array<size_t, 3> &mif_ext_ = tensor.sizes();                                                            \
    const auto &mif_c_ = tensor.constants;                                                                             \
    const size_t mif_x0_ = mif_c_.periodic_bc[0] ? 1 : 0, mif_x1_ = mif_ext_[0] - (mif_c_.periodic_bc[0] ? 1 : 0);      \
    const size_t mif_y0_ = (mif_c_.prev_proc_y != -1 || mif_c_.periodic_bc[1]) ? 1 : 0;                                 \
    const size_t mif_y1_ = mif_ext_[1] - ((mif_c_.next_proc_y != -1 || mif_c_.periodic_bc[1]) ? 1 : 0);                 \
    const size_t mif_z0_ = (mif_c_.prev_proc_z != -1 || mif_c_.periodic_bc[2]) ? 1 : 0;                                 \
    const size_t mif_z1_ = mif_ext_[2] - ((mif_c_.next_proc_z != -1 || mif_c_.periodic_bc[2]) ? 1 : 0);                 \
    for (size_t k = mif_z0_; k < mif_z1_; k++)                                                                         \
      for (size_t j = mif_y0_; j < mif_y1_; j++)                                                                       \
        for (size_t i = mif_x0_; i < mif_x1_; i++) {                                                                   \
          CODE                                                                                                         \
        }                                                                                                              \
  }

#define STAGGERED_TENSOR_FUNCTION_ON_ALL_POINTS(tensor, function, include_border, args...) \
  STAGGERED_TENSOR_ITERATE_OVER_ALL_POINTS(tensor, include_border, tensor(i, j, k) = function(args);)

#define VELOCITY_TENSOR_SET_FOR_ALL_POINTS(velocity, f_u, f_v, f_w, include_border, args...) \
  {                                                                                          \
    STAGGERED_TENSOR_FUNCTION_ON_ALL_POINTS(velocity.u, f_u, include_border, args)           \
    STAGGERED_TENSOR_FUNCTION_ON_ALL_POINTS(velocity.v, f_v, include_border, args)           \
    STAGGERED_TENSOR_FUNCTION_ON_ALL_POINTS(velocity.w, f_w, include_border, args)           \
  }

#endif  // STAGGERED_TENSOR_MACROS_H
