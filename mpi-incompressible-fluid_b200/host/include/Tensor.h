// Tensor.h -- the part of the reference's mif::Tensor (include/Tensor.h:42-238) that the projection path
// and its drivers use: an owning, zero-initialised, x-fastest array of up to three dimensions.
#ifndef MPI_INCOMPRESSIBLE_FLUID_TENSOR_H
#define MPI_INCOMPRESSIBLE_FLUID_TENSOR_H

#include <array>
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "Real.h"

namespace mif {

template <typename Type = Real, uint8_t SpaceDim = 3, typename DimensionsType = std::size_t>
class Tensor {
  static_assert(SpaceDim >= 1 && SpaceDim <= 3, "1 to 3 space dimensions are supported");

public:
  explicit Tensor(const std::array<DimensionsType, SpaceDim> &in_dimensions) : dimensions_(in_dimensions) {
    size_t total = 1;
    for (DimensionsType d : in_dimensions) total *= static_cast<size_t>(d);
    data_.assign(total, static_cast<Type>(0));
  }
  Tensor(const Tensor &) = delete;
  Tensor(Tensor &&) = default;
  Tensor &operator=(Tensor &&) = default;
  virtual ~Tensor() = default;

  // idx = i + j*sx + k*sx*sy (include/Tensor.h:232-238)
  Type &operator()(DimensionsType i) { return data_[static_cast<size_t>(i)]; }
  const Type &operator()(DimensionsType i) const { return data_[static_cast<size_t>(i)]; }
  Type &operator()(DimensionsType i, DimensionsType j) { return data_[offset(i, j)]; }
  const Type &operator()(DimensionsType i, DimensionsType j) const { return data_[offset(i, j)]; }
  Type &operator()(DimensionsType i, DimensionsType j, DimensionsType k) { return data_[offset(i, j, k)]; }
  const Type &operator()(DimensionsType i, DimensionsType j, DimensionsType k) const { return data_[offset(i, j, k)]; }

  const std::array<DimensionsType, SpaceDim> &sizes() const { return dimensions_; }
  size_t size() const { return data_.size(); }
  void *raw_data() { return data_.data(); }
  const void *raw_data() const { return data_.data(); }
  void swap_data(Tensor &other) { data_.swap(other.data_); }

protected:
  size_t offset(DimensionsType i, DimensionsType j) const {
    return static_cast<size_t>(i) + static_cast<size_t>(j) * static_cast<size_t>(dimensions_[0]);
  }
  size_t offset(DimensionsType i, DimensionsType j, DimensionsType k) const {
    return static_cast<size_t>(i) +
           static_cast<size_t>(dimensions_[0]) * (static_cast<size_t>(j) + static_cast<size_t>(dimensions_[1]) * static_cast<size_t>(k));
  }
  std::array<DimensionsType, SpaceDim> dimensions_;
  std::vector<Type> data_;
};

}  // namespace mif

#endif  // MPI_INCOMPRESSIBLE_FLUID_TENSOR_H
