// VTKDatExport.cpp -- output of the reference driver (solution.vtk, profile*.dat, full.vtk) for the host layer.
// One process holds the whole domain, so the MPI-IO offsets and gathers of the reference (src/VTKDatExport.cpp)
// reduce to sequential writes; the byte layout of the files is the reference's.
#include "VTKDatExport.h"

#include "../../../include/mifgpu.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace mif {

namespace {

// Index of the pressure plane at or just below `pos`, and the interpolation weight of that plane.  The reference
// returns the weight through a float (std::tuple<size_t, float>, src/VTKDatExport.cpp:44-51); kept, because the
// weight enters the profile values.
struct PlaneIndex {
  int index;
  Real weight;
};
PlaneIndex locate(Real pos, Real min_pos_global, Real delta) {
  const Real float_index = (pos - min_pos_global) / delta;
  const Real below = std::floor(float_index);
  const float weight = static_cast<float>(1.0 - (float_index - below));
  return {static_cast<int>(static_cast<size_t>(below)), static_cast<Real>(weight)};
}

// Which local indices are written: all owned pressure points; in a periodic direction the range starts on the
// ghost plane and the last rank covers one more point (the reference's COMPUTE_INDEXING, src/VTKDatExport.cpp:89-113).
struct WriteRange {
  int lo[3], hi[3], global_lo[3];
};
WriteRange write_range(const Constants &c) {
  WriteRange r;
  const int count[3] = {static_cast<int>(c.Nx_global),
                        static_cast<int>(c.Ny_owner + ((c.y_rank == c.Py - 1 && c.periodic_bc[1]) ? 1 : 0)),
                        static_cast<int>(c.Nz_owner + ((c.z_rank == c.Pz - 1 && c.periodic_bc[2]) ? 1 : 0))};
  r.lo[0] = c.periodic_bc[0] ? 1 : 0;
  r.lo[1] = c.prev_proc_y == -1 ? 0 : 1;
  r.lo[2] = c.prev_proc_z == -1 ? 0 : 1;
  const int base[3] = {c.base_i, c.base_j, c.base_k};
  for (int d = 0; d < 3; d++) {
    r.hi[d] = r.lo[d] + count[d];
    r.global_lo[d] = base[d] + r.lo[d];
  }
  return r;
}

// The writers read through peek(): the values they need were brought to the host by prefetch() (planes / lines
// gathered on the device) or by a whole-field refresh, never by a per-element synchronisation.
Real u_at_point(const VelocityTensor &v, int i, int j, int k) { return (v.u.peek(i, j, k) + v.u.peek(i + 1, j, k)) / 2; }
Real v_at_point(const VelocityTensor &v, int i, int j, int k) { return (v.v.peek(i, j, k) + v.v.peek(i, j + 1, k)) / 2; }
Real w_at_point(const VelocityTensor &v, int i, int j, int k) { return (v.w.peek(i, j, k) + v.w.peek(i, j, k + 1)) / 2; }

// Makes everything the writers read for the pressure points of the index box [lo, hi) available on the host.
// reach_all = false: the cell-centre averages above read index + 1 along the staggering direction of the tensor
// only; reach_all = true: the profile interpolation of writeDat reads index + 1 in every direction.
void prefetch(const StaggeredTensor &t, std::array<int, 3> lo, std::array<int, 3> hi, bool reach_all) {
  for (int d = 0; d < 3; d++)
    if (reach_all || static_cast<int>(t.staggering) == d) hi[d] += 1;
  t.fetch_box(lo, hi);
}
void prefetch(const VelocityTensor &velocity, const StaggeredTensor &pressure, const std::array<int, 3> &lo,
              const std::array<int, 3> &hi, bool reach_all) {
  prefetch(velocity.u, lo, hi, reach_all);
  prefetch(velocity.v, lo, hi, reach_all);
  prefetch(velocity.w, lo, hi, reach_all);
  prefetch(pressure, lo, hi, reach_all);
}

const char *const kTypeName = (sizeof(Real) == 8) ? "double" : "float";  // src/VTKDatExport.cpp:123

void append_big_endian(std::string &out, const std::vector<Real> &values) {
  for (Real value : values) {  // sizeof(Real) bytes per value, as the reference writes them (src/VTKDatExport.cpp:123)
    unsigned char raw[sizeof(Real)];
    std::memcpy(raw, &value, sizeof(Real));
    char bytes[sizeof(Real)];
    for (size_t b = 0; b < sizeof(Real); b++) bytes[b] = static_cast<char>(raw[sizeof(Real) - 1 - b]);  // little-endian host
    out.append(bytes, sizeof(Real));
  }
}

void write_file(const std::string &filename, const std::string &content) {
  FILE *file = std::fopen(filename.c_str(), "wb");
  if (!file) throw std::runtime_error("cannot open " + filename);
  std::fwrite(content.data(), 1, content.size(), file);
  std::fclose(file);
}

// MPI_Allgather of the counts + the MPI-IO offsets / MPI_Gatherv of the reference's writers (src/VTKDatExport.cpp:
// 54-69,219-311,519-554): on several ranks every local array travels to rank 0, which receives the ranks' parts one
// after the other in rank order -- the order the reference's file offsets produce.  Returns false on the ranks that do
// not write.
bool gather_to_root(const Constants &c, std::vector<Real> &values) {
  if (c.P == 1) return true;
  std::vector<uint64_t> counts(c.P, 0);
  // first call: sizes only, so that rank 0 can allocate
  std::vector<double> sizes(c.P, 0.0);
  sizes[c.rank] = static_cast<double>(values.size());
  if (mifgpu_allreduce(c.gpu(), sizes.data(), c.P, 0) != MIFGPU_OK) throw std::runtime_error(std::string("mifgpu_allreduce: ") + mifgpu_last_error());
  size_t total = 0;
  for (double n : sizes) total += static_cast<size_t>(n);
  // mifgpu_gather moves doubles in both builds of the library: a float build widens its values for the trip
  std::vector<double> mine(values.begin(), values.end()), all(c.rank == 0 ? total : 0);
  if (mifgpu_gather(c.gpu(), mine.data(), mine.size(), all.data(), counts.data()) != MIFGPU_OK)
    throw std::runtime_error(std::string("mifgpu_gather: ") + mifgpu_last_error());
  if (c.rank != 0) return false;
  values.assign(all.begin(), all.end());
  return true;
}

}  // namespace

void writeVTK(const std::string &filename, const VelocityTensor &velocity, const StaggeredTensor &pressure) {
  const Constants &c = velocity.constants;
  const WriteRange r = write_range(c);
  std::vector<Real> xyz, su, sv, sw, sp;
  auto add_point = [&](int i, int j, int k) {
    xyz.push_back(c.min_x_global + (c.base_i + i) * c.dx);
    xyz.push_back(c.min_y_global + (c.base_j + j) * c.dy);
    xyz.push_back(c.min_z_global + (c.base_k + k) * c.dz);
    su.push_back(u_at_point(velocity, i, j, k));
    sv.push_back(v_at_point(velocity, i, j, k));
    sw.push_back(w_at_point(velocity, i, j, k));
    sp.push_back(pressure.peek(i, j, k));
  };
  // Planes in the reference's order: z = 0 (i outer, j inner), x = 0 (j outer, k inner), y = 0 (i outer, k inner).
  {
    const int k_global = locate(0.0, c.min_z_global, c.dz).index;
    if (k_global >= r.global_lo[2] && k_global < r.global_lo[2] + (r.hi[2] - r.lo[2])) {
      const int k = k_global - r.global_lo[2] + r.lo[2];
      prefetch(velocity, pressure, {r.lo[0], r.lo[1], k}, {r.hi[0], r.hi[1], k + 1}, false);
      for (int i = r.lo[0]; i < r.hi[0]; i++)
        for (int j = r.lo[1]; j < r.hi[1]; j++) add_point(i, j, k);
    }
  }
  {
    const int i = locate(0.0, c.min_x_global, c.dx).index - r.global_lo[0] + r.lo[0];
    prefetch(velocity, pressure, {i, r.lo[1], r.lo[2]}, {i + 1, r.hi[1], r.hi[2]}, false);
    for (int j = r.lo[1]; j < r.hi[1]; j++)
      for (int k = r.lo[2]; k < r.hi[2]; k++) add_point(i, j, k);
  }
  {
    const int j_global = locate(0.0, c.min_y_global, c.dy).index;
    if (j_global >= r.global_lo[1] && j_global < r.global_lo[1] + (r.hi[1] - r.lo[1])) {
      const int j = j_global - r.global_lo[1] + r.lo[1];
      prefetch(velocity, pressure, {r.lo[0], j, r.lo[2]}, {r.hi[0], j + 1, r.hi[2]}, false);
      for (int i = r.lo[0]; i < r.hi[0]; i++)
        for (int k = r.lo[2]; k < r.hi[2]; k++) add_point(i, j, k);
    }
  }
  bool writer = true;
  for (std::vector<Real> *part : {&xyz, &su, &sv, &sw, &sp}) writer = gather_to_root(c, *part);
  if (!writer) return;
  const int points = static_cast<int>(su.size());
  char text[256];
  std::string out;
  std::snprintf(text, sizeof(text), "# vtk DataFile Version 2.0\nvtk output\nBINARY\nDATASET UNSTRUCTURED_GRID \nPOINTS %d %s\n",
                points, kTypeName);
  out += text;
  append_big_endian(out, xyz);
  std::snprintf(text, sizeof(text), "\nPOINT_DATA %d\nSCALARS u %s 1\nLOOKUP_TABLE default\n", points, kTypeName);
  out += text;
  append_big_endian(out, su);
  const char *names[3] = {"v", "w", "p"};
  const std::vector<Real> *fields[3] = {&sv, &sw, &sp};
  for (int f = 0; f < 3; f++) {
    std::snprintf(text, sizeof(text), "\nSCALARS %s %s 1\nLOOKUP_TABLE default\n", names[f], kTypeName);
    out += text;
    append_big_endian(out, *fields[f]);
  }
  write_file(filename, out);
}

void writeDat(const std::string &filename, const VelocityTensor &velocity, const StaggeredTensor &pressure,
              const int direction, const Real x, const Real y, const Real z) {
  if (direction < 0 || direction > 2) throw std::invalid_argument("writeDat: direction must be 0, 1 or 2");
  const Constants &c = velocity.constants;
  const WriteRange r = write_range(c);
  const PlaneIndex at[3] = {locate(x, c.min_x_global, c.dx), locate(y, c.min_y_global, c.dy), locate(z, c.min_z_global, c.dz)};
  const Real precision = 1e-6;
  bool aligned[3];
  for (int d = 0; d < 3; d++) aligned[d] = std::abs(at[d].weight - 1.0) < precision;
  const Real wi = at[0].weight, wj = at[1].weight, wk = at[2].weight;
  auto U = [&](int i, int j, int k) { return velocity.u.peek(i, j, k); };
  auto V = [&](int i, int j, int k) { return velocity.v.peek(i, j, k); };
  auto W = [&](int i, int j, int k) { return velocity.w.peek(i, j, k); };
  auto P = [&](int i, int j, int k) { return pressure.peek(i, j, k); };

  std::vector<Real> coordinate, su, sv, sw, sp;
  auto inside = [&](int d) { return at[d].index >= r.global_lo[d] && at[d].index < r.global_lo[d] + (r.hi[d] - r.lo[d]); };
  auto local = [&](int d) { return at[d].index - r.global_lo[d] + r.lo[d]; };
  // The expressions below are the reference's (src/VTKDatExport.cpp:392-515), including its use of velocity.v in
  // the second term of the unaligned-u branches of the y and z profiles.
  if (direction == 0 && inside(1) && inside(2)) {
    const int j = local(1), k = local(2);
    prefetch(velocity, pressure, {r.lo[0], j, k}, {r.hi[0], j + 1, k + 1}, true);
    for (int i = r.lo[0]; i < r.hi[0]; i++) {
      coordinate.push_back(c.min_x_global + (c.base_i + i) * c.dx);
      su.push_back(wj * wk * (U(i, j, k) + U(i + 1, j, k)) / 2 + (1 - wj) * wk * (U(i, j + 1, k) + U(i + 1, j + 1, k)) / 2 +
                   wj * (1 - wk) * (U(i, j, k + 1) + U(i + 1, j, k + 1)) / 2 +
                   (1 - wj) * (1 - wk) * (U(i, j + 1, k + 1) + U(i + 1, j + 1, k + 1)) / 2);
      if (aligned[1]) sv.push_back(wk * (V(i, j, k) + V(i, j + 1, k)) / 2 + (1 - wk) * (V(i, j, k + 1) + V(i, j + 1, k + 1)) / 2);
      else sv.push_back(wk * V(i, j + 1, k) + (1 - wk) * V(i, j + 1, k + 1));
      if (aligned[2]) sw.push_back(wj * (W(i, j, k) + W(i, j, k + 1)) / 2 + (1 - wj) * (W(i, j + 1, k) + W(i, j + 1, k + 1)) / 2);
      else sw.push_back(wj * W(i, j, k + 1) + (1 - wj) * W(i, j + 1, k + 1));
      sp.push_back(wj * wk * P(i, j, k) + (1 - wj) * wk * P(i, j + 1, k) + wj * (1 - wk) * P(i, j, k + 1) +
                   (1 - wj) * (1 - wk) * P(i, j + 1, k + 1));
    }
  } else if (direction == 1 && inside(0) && inside(2)) {
    const int i = local(0), k = local(2);
    prefetch(velocity, pressure, {i, r.lo[1], k}, {i + 1, r.hi[1], k + 1}, true);
    for (int j = r.lo[1]; j < r.hi[1]; j++) {
      coordinate.push_back(c.min_y_global + (c.base_j + j) * c.dy);
      if (aligned[0]) su.push_back(wk * (U(i, j, k) + U(i + 1, j, k)) / 2 + (1 - wk) * (U(i, j, k + 1) + U(i + 1, j, k + 1)) / 2);
      else su.push_back(wk * U(i + 1, j, k) + (1 - wk) * V(i + 1, j, k + 1));
      sv.push_back(wi * wk * (V(i, j, k) + V(i, j + 1, k)) / 2 + (1 - wi) * wk * (V(i + 1, j, k) + V(i + 1, j + 1, k)) / 2 +
                   wi * (1 - wk) * (V(i, j, k + 1) + V(i, j + 1, k + 1)) / 2 +
                   (1 - wi) * (1 - wk) * (V(i + 1, j, k + 1) + V(i + 1, j + 1, k + 1)) / 2);
      if (aligned[2]) sw.push_back(wi * (W(i, j, k) + W(i, j, k + 1)) / 2 + (1 - wi) * (W(i + 1, j, k) + W(i + 1, j, k + 1)) / 2);
      else sw.push_back(wi * W(i, j, k + 1) + (1 - wi) * W(i + 1, j, k + 1));
      sp.push_back(wi * wk * P(i, j, k) + (1 - wi) * wk * P(i + 1, j, k) + wi * (1 - wk) * P(i, j, k + 1) +
                   (1 - wi) * (1 - wk) * P(i + 1, j, k + 1));
    }
  } else if (direction == 2 && inside(0) && inside(1)) {
    const int i = local(0), j = local(1);
    prefetch(velocity, pressure, {i, j, r.lo[2]}, {i + 1, j + 1, r.hi[2]}, true);
    for (int k = r.lo[2]; k < r.hi[2]; k++) {
      coordinate.push_back(c.min_z_global + (c.base_k + k) * c.dz);
      if (aligned[0]) su.push_back(wj * (U(i, j, k) + U(i + 1, j, k)) / 2 + (1 - wj) * (U(i, j + 1, k) + U(i + 1, j + 1, k)) / 2);
      else su.push_back(wj * U(i + 1, j, k) + (1 - wj) * V(i + 1, j + 1, k));
      if (aligned[1]) sv.push_back(wi * (V(i, j, k) + V(i, j + 1, k)) / 2 + (1 - wi) * (V(i + 1, j, k) + V(i + 1, j + 1, k)) / 2);
      else sv.push_back(wi * V(i, j + 1, k) + (1 - wi) * V(i + 1, j + 1, k));
      sw.push_back(wi * wj * (W(i, j, k) + W(i, j, k + 1)) / 2 + (1 - wi) * wj * (W(i + 1, j, k) + W(i + 1, j, k + 1)) / 2 +
                   wi * (1 - wj) * (W(i, j + 1, k) + W(i, j + 1, k + 1)) / 2 +
                   (1 - wi) * (1 - wj) * (W(i + 1, j + 1, k) + W(i + 1, j + 1, k + 1)) / 2);
      sp.push_back(wi * wj * P(i, j, k) + (1 - wi) * wj * P(i + 1, j, k) + wi * (1 - wj) * P(i, j + 1, k) +
                   (1 - wi) * (1 - wj) * P(i + 1, j + 1, k));
    }
  }
  bool writer = true;
  for (std::vector<Real> *part : {&coordinate, &su, &sv, &sw, &sp}) writer = gather_to_root(c, *part);
  if (!writer) return;
  // rows ordered by coordinate (stable, like the reference's insertion sort)
  std::vector<size_t> order(coordinate.size());
  std::iota(order.begin(), order.end(), size_t{0});
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return coordinate[a] < coordinate[b]; });
  FILE *file = std::fopen(filename.c_str(), "w");
  if (!file) throw std::runtime_error("cannot open " + filename);
  for (size_t n : order) {
    const Real px = direction == 0 ? coordinate[n] : x, py = direction == 1 ? coordinate[n] : y, pz = direction == 2 ? coordinate[n] : z;
    std::fprintf(file, "%.8f %.8f %.8f %.8e %.8e %.8e %.8e\n", px, py, pz, su[n], sv[n], sw[n], sp[n]);
  }
  std::fclose(file);
}

void writeVTKFullMesh(const std::string &filename, const mif::VelocityTensor &velocity, const StaggeredTensor &pressure) {
  const Constants &c = velocity.constants;
  if (c.Py * c.Pz != 1) throw std::invalid_argument("writeVTKFullMesh needs a single rank");
  const WriteRange r = write_range(c);
  velocity.u.sync_host();  // every point is written: whole fields
  velocity.v.sync_host();
  velocity.w.sync_host();
  pressure.sync_host();
  const int nx = r.hi[0] - r.lo[0], ny = r.hi[1] - r.lo[1], nz = r.hi[2] - r.lo[2];
  std::ofstream out(filename);
  out << "# vtk DataFile Version 3.0\npressure mesh solution\nASCII\nDATASET STRUCTURED_POINTS\n"
      << "DIMENSIONS " << nx << ' ' << ny << ' ' << nz << '\n'
      << "ORIGIN " << c.min_x_global << " " << c.min_y_global << " " << c.min_z_global << "\n"
      << "SPACING " << c.dx << ' ' << c.dy << ' ' << c.dz << '\n'
      << "POINT_DATA " << nx * ny * nz << '\n';
  auto scalar = [&](const char *name, auto &&value) {
    out << "SCALARS " << name << " double 1\nLOOKUP_TABLE default\n";
    for (int k = r.lo[2]; k < r.hi[2]; ++k)
      for (int j = r.lo[1]; j < r.hi[1]; ++j)
        for (int i = r.lo[0]; i < r.hi[0]; ++i) out << value(i, j, k) << ' ';
  };
  scalar("u", [&](int i, int j, int k) { return u_at_point(velocity, i, j, k); });
  scalar("v", [&](int i, int j, int k) { return v_at_point(velocity, i, j, k); });
  scalar("w", [&](int i, int j, int k) { return w_at_point(velocity, i, j, k); });
  scalar("|u|", [&](int i, int j, int k) {
    const Real ux = u_at_point(velocity, i, j, k), uy = v_at_point(velocity, i, j, k), uz = w_at_point(velocity, i, j, k);
    return std::sqrt(ux * ux + uy * uy + uz * uz);
  });
  scalar("p", [&](int i, int j, int k) { return pressure.peek(i, j, k); });
}

}  // namespace mif
