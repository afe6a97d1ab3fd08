// VelocityTensor.h -- drop-in for include/VelocityTensor.h:10-113.
#ifndef VELOCITY_TENSOR_H
#define VELOCITY_TENSOR_H

#include "StaggeredTensor.h"
#include "VectorFunction.h"

namespace mif {

// The three velocity components live on the faces of the pressure cells: u is shifted half a cell backwards in
// x, v in y, w in z (include/VelocityTensor.h:10-86).  The coordinate shift itself is handled by
// StaggeredTensor::coordinate from the staggering direction.
class UTensor : public StaggeredTensor {
public:
  explicit UTensor(const Constants &constants) : StaggeredTensor(constants, StaggeringDirection::x) {}
};
class VTensor : public StaggeredTensor {
public:
  explicit VTensor(const Constants &constants) : StaggeredTensor(constants, StaggeringDirection::y) {}
};
class WTensor : public StaggeredTensor {
public:
  explicit WTensor(const Constants &constants) : StaggeredTensor(constants, StaggeringDirection::z) {}
};

class VelocityTensor {
public:
  UTensor u;
  VTensor v;
  WTensor w;
  std::array<StaggeredTensor *, 3> components;
  const Constants &constants;

  explicit VelocityTensor(const Constants &constants);
  VelocityTensor(const VelocityTensor &) = delete;

  void swap_data(VelocityTensor &other);
  void set(const VectorFunction &f, bool include_border);
  // Dirichlet faces (+ periodic ghosts) on the GPU (src/VelocityTensor.cpp:36-233).
  void apply_bc(const VectorFunction &exact_velocity);
};

}  // namespace mif

#endif  // VELOCITY_TENSOR_H
