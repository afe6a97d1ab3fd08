// StaggeredTensor.h -- drop-in for include/StaggeredTensor.h:30-142.  The host array keeps the reference's
// ghosted x-fastest layout and element access; in addition every tensor owns a device-resident twin
// (mifgpu_tensor).  During the time loop the device copy is authoritative; the host copy is refreshed
// lazily the first time it is read after a device-side update (before norms, VTK output, element access).
#ifndef STAGGERED_TENSOR_H
#define STAGGERED_TENSOR_H

#include <functional>

#include "Constants.h"
#include "Tensor.h"
#ifdef MIF_REFERENCE_SOURCE_COMPAT
#include <mpi.h>  // host/compat/mpi.h: the reference's StaggeredTensor.h:7 exposes MPI to its includers (test/velocity_test.cpp)
#endif

struct mifgpu_tensor;

namespace mif {

enum StaggeringDirection { x, y, z, none };

class StaggeredTensor : public Tensor<Real, 3U, size_t> {
public:
  StaggeredTensor(const Constants &constants, const StaggeringDirection &staggering);
  StaggeredTensor(const StaggeredTensor &) = delete;
  ~StaggeredTensor() override;

  const Constants &constants;
  StaggeringDirection staggering;

  // Element access in the reference layout.  Reading pulls the field back from the GPU if the device copy
  // is newer; the non-const overload also marks the host copy as the newer one.
  Real &operator()(size_t i, size_t j, size_t k) {
    host_for_write();
    return Tensor::operator()(i, j, k);
  }
  const Real &operator()(size_t i, size_t j, size_t k) const {
    host_for_read();
    return Tensor::operator()(i, j, k);
  }
  Real &operator()(size_t i) {
    host_for_write();
    return Tensor::operator()(i);
  }
  const Real &operator()(size_t i) const {
    host_for_read();
    return Tensor::operator()(i);
  }
  void *raw_data() {
    host_for_write();
    return Tensor::raw_data();
  }
  void swap_data(StaggeredTensor &other);

  // Halo exchange entry points of the reference (src/StaggeredTensor.cpp:60-165).  The exchange itself
  // happens inside libmifgpu on the device, so on the host these only keep the call sequence valid.
  void send_mpi_data(int base_tag);
  void receive_mpi_data(int base_tag);
  void recompute_mpi_addressing();
  // Periodic ghost copy on the host array (src/StaggeredTensor.cpp:221-257).
  void apply_periodic_bc();

  // Analytic functions sampled at grid indices.  "unstaggered" is the pressure point of index (i, j, k):
  // min + h * (base + index) (include/StaggeredTensor.h:113-128); the plain variant subtracts half a cell in
  // the staggering direction of this tensor (include/VelocityTensor.h:16-22,40-46,64-70).
  Real coordinate(int direction, int index, bool staggered) const {
    const Real lo = direction == 0 ? constants.min_x_global : (direction == 1 ? constants.min_y_global : constants.min_z_global);
    const Real h = direction == 0 ? constants.dx : (direction == 1 ? constants.dy : constants.dz);
    const int base = direction == 0 ? constants.base_i : (direction == 1 ? constants.base_j : constants.base_k);
    const Real point = lo + h * (base + index);
    if (!staggered || static_cast<int>(staggering) != direction) return point;
    return point - (direction == 0 ? constants.dx_over_2 : (direction == 1 ? constants.dy_over_2 : constants.dz_over_2));
  }
  Real evaluate_function_at_index_unstaggered(Real time, int i, int j, int k,
                                              const std::function<Real(Real, Real, Real, Real)> &f) const {
    return f(time, coordinate(0, i, false), coordinate(1, j, false), coordinate(2, k, false));
  }
  Real evaluate_function_at_index_unstaggered(int i, int j, int k, const std::function<Real(Real, Real, Real)> &f) const {
    return f(coordinate(0, i, false), coordinate(1, j, false), coordinate(2, k, false));
  }
  virtual Real evaluate_function_at_index(Real time, int i, int j, int k,
                                          const std::function<Real(Real, Real, Real, Real)> &f) const {
    return f(time, coordinate(0, i, true), coordinate(1, j, true), coordinate(2, k, true));
  }
  virtual Real evaluate_function_at_index(int i, int j, int k, const std::function<Real(Real, Real, Real)> &f) const {
    return f(coordinate(0, i, true), coordinate(1, j, true), coordinate(2, k, true));
  }

  void print() const;
  void print(const std::function<bool(Real)> &filter) const;
  void print_inline() const;

  // Fill from an analytic function at the staggered coordinates (src/StaggeredTensor.cpp:205-219).
  void set(const std::function<Real(Real, Real, Real)> &f, bool include_border);

  // ---- device twin (not part of the reference interface) ------------------------------------------
  mifgpu_tensor *device() const;    // device tensor, up to date with the host copy
  void device_was_written() const;  // libmifgpu changed the device copy: the host copy is stale
  void sync_host() const { host_for_read(); }
  // Output path (src/VTKDatExport.cpp reads three planes and a few lines, not whole fields): when the newest copy
  // is on the device, fetch_box brings only the index box [lo, hi) into the host array -- the host copy as a whole
  // stays stale -- and peek reads the host array without synchronising.  A box that leaves the tensor (the
  // reference's own out-of-row reads next to periodic / last planes) falls back to refreshing the whole field.
  void fetch_box(const std::array<int, 3> &lo, const std::array<int, 3> &hi) const;
  const Real &peek(size_t i, size_t j, size_t k) const { return Tensor::operator()(i, j, k); }

private:
  void host_for_read() const;
  void host_for_write() {
    host_for_read();
    device_valid_ = false;
  }
  mutable mifgpu_tensor *device_ = nullptr;
  mutable bool host_valid_ = true;
  mutable bool device_valid_ = false;
};

}  // namespace mif

#endif  // STAGGERED_TENSOR_H
