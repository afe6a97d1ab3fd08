// PressureTensor.h -- drop-in for include/PressureTensor.h:16-36: the ghost-free scratch buffer the reference
// hands to its Poisson solver.  libmifgpu transforms in place inside the ghosted delta-p tensor, so this
// buffer is not touched by the GPU path; it is kept so that call sites compile and can still use the copies.
#ifndef PRESSURE_TENSOR_H
#define PRESSURE_TENSOR_H

#include "PressureSolverStructures.h"
#include "StaggeredTensor.h"

namespace mif {

class PressureTensor : public Tensor<Real, 1U, int> {
public:
  PressureSolverStructures &structures;
  const int max_size;

  explicit PressureTensor(PressureSolverStructures &structures);
  PressureTensor(const PressureTensor &) = delete;

  void copy_from_staggered(const StaggeredTensor &other);
  void copy_to_staggered(StaggeredTensor &other, int base_tag) const;
  void print_inline() const;
};

}  // namespace mif

#endif  // PRESSURE_TENSOR_H
