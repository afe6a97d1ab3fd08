// VectorFunction.h -- drop-in for include/VectorFunction.h:16-54: bundles of three std::function components.
#ifndef VECTOR_FUNCTION_H
#define VECTOR_FUNCTION_H

#include <array>
#include <functional>

#include "Real.h"

namespace mif {

class VectorFunction {
public:
  using Component = std::function<Real(Real, Real, Real)>;
  VectorFunction(const Component f_u, const Component f_v, const Component f_w);
  VectorFunction(const VectorFunction &other);

  const Component f_u, f_v, f_w;
  const std::array<const Component *, 3> components;

  VectorFunction operator+(const VectorFunction &other) const;
  VectorFunction operator*(const Real scalar) const;
};

class TimeVectorFunction {
public:
  using Component = std::function<Real(Real, Real, Real, Real)>;
  TimeVectorFunction(const Component f_u, const Component f_v, const Component f_w);
  TimeVectorFunction(const TimeVectorFunction &other);

  const Component f_u, f_v, f_w;
  const std::array<const Component *, 3> components;

  // The field frozen at `time` (src/VectorFunction.cpp:46-50).
  VectorFunction set_time(Real time) const;
  // f(time_2) - f(time_1): "second minus first", exactly as src/VectorFunction.cpp:52-60.
  VectorFunction get_difference_over_time(Real time_1, Real time_2) const;
};

}  // namespace mif

#endif  // VECTOR_FUNCTION_H
