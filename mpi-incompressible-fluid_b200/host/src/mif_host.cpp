// mif_host.cpp -- the C++ host layer that gives libmifgpu (include/mifgpu.h) the reference's own class and
// function names (namespace mif), so that the reference's drivers and tests compile against it unchanged in
// meaning: Constants, StaggeredTensor, VelocityTensor, PressureTensor, PressureSolverStructures, timestep,
// timestep_nhn, solve_pressure_equation_*, adjust_pressure, the norms and parse_input_file.
// Everything numerical on the time-step path runs on the GPU through the C ABI; this file only moves data and
// translates std::function boundary data into what the ABI accepts.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <cctype>
#include <initializer_list>
#include <string>
#include <vector>
#include <thread>
#include <chrono>
#include <cstdio>
#include <unistd.h>

#include "../../../include/mifgpu.h"
#include "Constants.h"
#include "InputParser.h"
#include "Launch.h"
#include "Manufactured.h"
#include "Norms.h"
#include "PressureEquation.h"
#include "TestCaseBoundaries.h"
#include "ManufacturedVelocity.h"
#include "Timestep.h"
#include "TimestepVelocity.h"

namespace mif {

namespace {

[[noreturn]] void throw_gpu_error(const char *what) {
  throw std::runtime_error(std::string(what) + ": " + mifgpu_last_error());
}
void check(int rc, const char *what) {
  if (rc != MIFGPU_OK) throw_gpu_error(what);
}

// Rendezvous file of the n-th context of this job: $MIF_RENDEZVOUS_DIR (default /tmp) / mif_comm_<job>_<n>, where
// <job> is MIF_JOB_ID (scripts/mifrun), else what the launcher at hand provides, else the parent process id (the
// ranks of one mpirun / shell loop share their parent).
std::string rendezvous_file(int n) {
  const char *dir = std::getenv("MIF_RENDEZVOUS_DIR");
  std::string job;
  for (const char *name : {"MIF_JOB_ID", "PMIX_NAMESPACE", "OMPI_MCA_ess_base_jobid", "SLURM_STEP_ID", "TORCHELASTIC_RUN_ID"})
    if (const char *value = std::getenv(name)) {
      job = value;
      if (std::string(name) == "SLURM_STEP_ID" && std::getenv("SLURM_JOB_ID")) job = std::string(std::getenv("SLURM_JOB_ID")) + "." + job;
      break;
    }
  if (job.empty()) job = "ppid" + std::to_string(static_cast<long>(getppid()));
  for (char &c : job)
    if (!(std::isalnum(static_cast<unsigned char>(c)) || c == '.' || c == '-' || c == '_')) c = '_';
  return std::string(dir ? dir : "/tmp") + "/mif_comm_" + job + "_" + std::to_string(n);
}

size_t owners(size_t points, int parts, int index) {
  return points / parts + (static_cast<size_t>(index) < points % parts ? 1 : 0);
}
// Number of local unstaggered points along a decomposed direction, ghosts included (src/Constants.cpp:81-82).
size_t local_extent(size_t n_global, bool periodic, int parts, int index, size_t owner) {
  if (parts == 1) return periodic ? n_global + 1 : n_global;
  const bool at_wall = !periodic && (index == 0 || index == parts - 1);
  return at_wall ? owner + 1 : owner + 2;
}
int base_index(size_t points, bool periodic, int parts, int index) {
  const size_t first_owned = points / parts * index + std::min(static_cast<size_t>(index), points % parts);
  return static_cast<int>(first_owned) - ((index > 0 || periodic) ? 1 : 0);
}

}  // namespace

// ---- Constants (src/Constants.cpp:58-120) ----------------------------------------------------------
Constants::Constants(size_t Nx_global_, size_t Ny_global_, size_t Nz_global_, Real x_size_, Real y_size_global_,
                     Real z_size_global_, Real min_x_global_, Real min_y_global_, Real min_z_global_, Real Re_,
                     Real final_time_, unsigned int num_time_steps_, int Py_, int Pz_, int rank_,
                     const std::array<bool, 3> &periodic_bc_)
    : Nx_global(Nx_global_), Ny_global(Ny_global_), Nz_global(Nz_global_),
      x_size(x_size_), y_size_global(y_size_global_), z_size_global(z_size_global_),
      min_x_global(min_x_global_), min_y_global(min_y_global_), min_z_global(min_z_global_),
      Re(Re_), final_time(final_time_), num_time_steps(num_time_steps_), periodic_bc(periodic_bc_),
      Py(Py_), Pz(Pz_), rank(rank_), y_rank(rank_ / Pz_), z_rank(rank_ % Pz_),
      dt(final_time_ / num_time_steps_),
      Nx_domains(Nx_global_ - 1), Ny_domains_global(Ny_global_ - 1), Nz_domains_global(Nz_global_ - 1),
      dx(x_size_ / Nx_domains), dy(y_size_global_ / Ny_domains_global), dz(z_size_global_ / Nz_domains_global),
      one_over_2_dx(1 / (2 * dx)), one_over_2_dy(1 / (2 * dy)), one_over_2_dz(1 / (2 * dz)),
      one_over_8_dx(1 / (8 * dx)), one_over_8_dy(1 / (8 * dy)), one_over_8_dz(1 / (8 * dz)),
      one_over_dx2_Re(1 / (Re_ * dx * dx)), one_over_dy2_Re(1 / (Re_ * dy * dy)), one_over_dz2_Re(1 / (Re_ * dz * dz)),
      dx_over_2(dx / 2), dy_over_2(dy / 2), dz_over_2(dz / 2),
      one_over_dx(1 / dx), one_over_dy(1 / dy), one_over_dz(1 / dz),
      P(Py_ * Pz_),
      Ny_owner(owners(Ny_global_ - periodic_bc_[1], Py_, y_rank)),
      Nz_owner(owners(Nz_global_ - periodic_bc_[2], Pz_, z_rank)),
      Nx(periodic_bc_[0] ? Nx_global_ + 1 : Nx_global_),
      Ny(local_extent(Ny_global_, periodic_bc_[1], Py_, y_rank, Ny_owner)),
      Nz(local_extent(Nz_global_, periodic_bc_[2], Pz_, z_rank, Nz_owner)),
      Nx_staggered(periodic_bc_[0] ? Nx : Nx + 1),
      Ny_staggered(y_rank == Py_ - 1 ? Ny + 1 : Ny),
      Nz_staggered(z_rank == Pz_ - 1 ? Nz + 1 : Nz),
      base_i(periodic_bc_[0] ? -1 : 0),
      base_j(base_index(Ny_global_ - periodic_bc_[1], periodic_bc_[1], Py_, y_rank)),
      base_k(base_index(Nz_global_ - periodic_bc_[2], periodic_bc_[2], Pz_, z_rank)),
      prev_proc_y(y_rank == 0 ? ((Py_ > 1 && periodic_bc_[1]) ? rank_ + (Py_ - 1) * Pz_ : -1) : rank_ - Pz_),
      next_proc_y(y_rank == Py_ - 1 ? ((Py_ > 1 && periodic_bc_[1]) ? rank_ - (Py_ - 1) * Pz_ : -1) : rank_ + Pz_),
      prev_proc_z(z_rank == 0 ? ((Pz_ > 1 && periodic_bc_[2]) ? rank_ + Pz_ - 1 : -1) : rank_ - 1),
      next_proc_z(z_rank == Pz_ - 1 ? ((Pz_ > 1 && periodic_bc_[2]) ? rank_ - (Pz_ - 1) : -1) : rank_ + 1) {}

Constants::~Constants() {
  if (gpu_ctx_) mifgpu_destroy(gpu_ctx_);
}

mifgpu_ctx *Constants::gpu() const {
  if (gpu_ctx_) return gpu_ctx_;
  mifgpu_params p{};
  p.Nx_global = Nx_global; p.Ny_global = Ny_global; p.Nz_global = Nz_global;
  p.x_size = x_size; p.y_size_global = y_size_global; p.z_size_global = z_size_global;
  p.min_x_global = min_x_global; p.min_y_global = min_y_global; p.min_z_global = min_z_global;
  p.Re = Re; p.final_time = final_time; p.num_time_steps = num_time_steps;
  p.Py = Py; p.Pz = Pz; p.rank = rank;
  for (int d = 0; d < 3; d++) p.periodic_bc[d] = periodic_bc[d];
  const char *device = std::getenv("MIFGPU_DEVICE");
  p.device = device ? std::atoi(device) : (P > 1 ? launch_local_rank() : 0);
  if (P == 1) {
    check(mifgpu_create(&p, &gpu_ctx_), "mifgpu_create");
    return gpu_ctx_;
  }
  // Several ranks (one process per GPU): rank 0 publishes a fresh communicator id in a rendezvous file, the others wait
  // for it; creating the context is collective, so once it returns on rank 0 every rank has read the file.
  static int contexts_created = 0;  // all ranks create their contexts in the same order
  const std::string path = rendezvous_file(contexts_created++);
  char id[MIFGPU_UNIQUE_ID_BYTES];
  if (rank == 0) {
    check(mifgpu_comm_unique_id(id), "mifgpu_comm_unique_id");
    const std::string tmp = path + ".tmp";
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) throw std::runtime_error("cannot write " + tmp);
    std::fclose(f);
    if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("cannot publish " + path);
  } else {
    FILE *f = nullptr;
    for (int waited_ms = 0; !(f = std::fopen(path.c_str(), "rb")); waited_ms += 5) {
      if (waited_ms > 120000) throw std::runtime_error("no communicator id from rank 0 after 120 s (" + path + ")");
      std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    const size_t got = std::fread(id, 1, sizeof(id), f);
    std::fclose(f);
    if (got != sizeof(id)) throw std::runtime_error("short communicator id in " + path);
  }
  const int rc = mifgpu_create_distributed(&p, id, &gpu_ctx_);
  if (rank == 0) std::remove(path.c_str());
  check(rc, "mifgpu_create_distributed");
  return gpu_ctx_;
}

// ---- Launch (replaces MPI_Init / MPI_Comm_rank / MPI_Comm_size) ----------------------------------------
namespace {
int env_int(std::initializer_list<const char *> names, int fallback) {
  for (const char *name : names)
    if (const char *value = std::getenv(name)) return std::atoi(value);
  return fallback;
}
}  // namespace
int launch_rank() { return env_int({"MIF_RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID", "RANK"}, 0); }
int launch_size() { return env_int({"MIF_WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS", "WORLD_SIZE"}, 1); }
int launch_local_rank() {
  return env_int({"MIF_LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID", "LOCAL_RANK"}, launch_rank());
}

// ---- StaggeredTensor --------------------------------------------------------------------------------
namespace {
std::array<size_t, 3> extents_of(const Constants &c, StaggeringDirection s) {
  return {s == StaggeringDirection::x ? c.Nx_staggered : c.Nx, s == StaggeringDirection::y ? c.Ny_staggered : c.Ny,
          s == StaggeringDirection::z ? c.Nz_staggered : c.Nz};
}
}  // namespace

StaggeredTensor::StaggeredTensor(const Constants &constants_, const StaggeringDirection &staggering_)
    : Tensor(extents_of(constants_, staggering_)), constants(constants_), staggering(staggering_) {}

StaggeredTensor::~StaggeredTensor() {
  if (device_) mifgpu_tensor_destroy(device_);
}

mifgpu_tensor *StaggeredTensor::device() const {
  if (!device_) check(mifgpu_tensor_create(constants.gpu(), static_cast<int>(staggering), &device_), "mifgpu_tensor_create");
  if (!device_valid_) {
    check(mifgpu_tensor_upload(device_, data_.data()), "mifgpu_tensor_upload");
    device_valid_ = true;
  }
  return device_;
}

void StaggeredTensor::device_was_written() const {
  device_valid_ = true;
  host_valid_ = false;
}

void StaggeredTensor::host_for_read() const {
  if (host_valid_) return;
  check(mifgpu_tensor_download(device_, const_cast<Real *>(data_.data())), "mifgpu_tensor_download");
  host_valid_ = true;
}

void StaggeredTensor::fetch_box(const std::array<int, 3> &lo, const std::array<int, 3> &hi) const {
  if (host_valid_) return;
  static const bool whole_fields = std::getenv("MIF_EXPORT_WHOLE_FIELDS") != nullptr;  // A/B switch
  const auto &s = sizes();
  bool inside = !whole_fields;
  for (int d = 0; d < 3; d++)
    if (lo[d] < 0 || hi[d] <= lo[d] || hi[d] > static_cast<int>(s[d])) inside = false;
  if (!inside) {
    host_for_read();
    return;
  }
  const size_t bx = hi[0] - lo[0], by = hi[1] - lo[1], bz = hi[2] - lo[2];
  std::vector<Real> box(bx * by * bz);
  const int32_t lo32[3] = {lo[0], lo[1], lo[2]}, hi32[3] = {hi[0], hi[1], hi[2]};
  check(mifgpu_tensor_download_box(device_, lo32, hi32, box.data()), "mifgpu_tensor_download_box");
  Real *host = const_cast<Real *>(data_.data());
  for (size_t k = 0; k < bz; k++)
    for (size_t j = 0; j < by; j++)
      std::memcpy(host + offset(lo[0], lo[1] + j, lo[2] + k), box.data() + (k * by + j) * bx, bx * sizeof(Real));
}

void StaggeredTensor::swap_data(StaggeredTensor &other) {
  Tensor::swap_data(other);
  std::swap(host_valid_, other.host_valid_);
  std::swap(device_valid_, other.device_valid_);
  std::swap(device_, other.device_);
}

void StaggeredTensor::send_mpi_data(int) {}
void StaggeredTensor::receive_mpi_data(int) {}
void StaggeredTensor::recompute_mpi_addressing() {}

void StaggeredTensor::apply_periodic_bc() {
  const auto &s = sizes();
  host_for_write();
  auto at = [&](size_t i, size_t j, size_t k) -> Real & { return Tensor::operator()(i, j, k); };
  if (constants.periodic_bc[0])
    for (size_t k = 0; k < s[2]; k++)
      for (size_t j = 0; j < s[1]; j++) {
        at(0, j, k) = at(s[0] - 2, j, k);
        at(s[0] - 1, j, k) = at(1, j, k);
      }
  if (constants.periodic_bc[1] && constants.Py == 1)
    for (size_t k = 0; k < s[2]; k++)
      for (size_t i = 0; i < s[0]; i++) {
        at(i, 0, k) = at(i, s[1] - 2, k);
        at(i, s[1] - 1, k) = at(i, 1, k);
      }
  if (constants.periodic_bc[2] && constants.Pz == 1)
    for (size_t j = 0; j < s[1]; j++)
      for (size_t i = 0; i < s[0]; i++) {
        at(i, j, 0) = at(i, j, s[2] - 2);
        at(i, j, s[2] - 1) = at(i, j, 1);
      }
}

void StaggeredTensor::set(const std::function<Real(Real, Real, Real)> &f, bool include_border) {
  const auto &s = sizes();
  const size_t lo = include_border ? 0 : 1, trim = include_border ? 0 : 1;
  host_for_write();
  for (size_t k = lo; k < s[2] - trim; k++)
    for (size_t j = lo; j < s[1] - trim; j++)
      for (size_t i = lo; i < s[0] - trim; i++) Tensor::operator()(i, j, k) = evaluate_function_at_index(i, j, k, f);
}

void StaggeredTensor::print() const {
  const auto &s = sizes();
  host_for_read();
  for (size_t k = 0; k < s[2]; k++) {
    for (size_t j = 0; j < s[1]; j++) {
      for (size_t i = 0; i < s[0]; i++) std::cout << Tensor::operator()(i, j, k) << " ";
      std::cout << std::endl;
    }
    std::cout << std::endl;
  }
  std::cout << std::endl;
}

void StaggeredTensor::print(const std::function<bool(Real)> &filter) const {
  const auto &s = sizes();
  host_for_read();
  for (size_t k = 0; k < s[2]; k++)
    for (size_t j = 0; j < s[1]; j++)
      for (size_t i = 0; i < s[0]; i++) {
        const Real value = Tensor::operator()(i, j, k);
        if (filter(value)) std::cout << "(" << i << "," << j << "," << k << "): " << value << std::endl;
      }
  std::cout << std::endl;
}

void StaggeredTensor::print_inline() const {
  host_for_read();
  for (size_t i = 0; i < size(); i++) std::cout << Tensor::operator()(i) << " ";
  std::cout << std::endl;
}

// ---- VectorFunction (src/VectorFunction.cpp) ----------------------------------------------------------
VectorFunction::VectorFunction(const Component f_u_, const Component f_v_, const Component f_w_)
    : f_u(f_u_), f_v(f_v_), f_w(f_w_), components{&this->f_u, &this->f_v, &this->f_w} {}
VectorFunction::VectorFunction(const VectorFunction &other)
    : f_u(other.f_u), f_v(other.f_v), f_w(other.f_w), components{&this->f_u, &this->f_v, &this->f_w} {}

VectorFunction VectorFunction::operator*(const Real scalar) const {
  auto scaled = [scalar](const Component &f) { return Component([scalar, f](Real x, Real y, Real z) { return scalar * f(x, y, z); }); };
  return VectorFunction(scaled(f_u), scaled(f_v), scaled(f_w));
}
VectorFunction VectorFunction::operator+(const VectorFunction &other) const {
  auto sum = [](const Component &f, const Component &g) {
    return Component([f, g](Real x, Real y, Real z) { return f(x, y, z) + g(x, y, z); });
  };
  return VectorFunction(sum(f_u, other.f_u), sum(f_v, other.f_v), sum(f_w, other.f_w));
}

TimeVectorFunction::TimeVectorFunction(const Component f_u_, const Component f_v_, const Component f_w_)
    : f_u(f_u_), f_v(f_v_), f_w(f_w_), components{&this->f_u, &this->f_v, &this->f_w} {}
TimeVectorFunction::TimeVectorFunction(const TimeVectorFunction &other)
    : f_u(other.f_u), f_v(other.f_v), f_w(other.f_w), components{&this->f_u, &this->f_v, &this->f_w} {}

VectorFunction TimeVectorFunction::set_time(Real time) const {
  auto frozen = [time](const Component &f) {
    return VectorFunction::Component([time, f](Real x, Real y, Real z) { return f(time, x, y, z); });
  };
  return VectorFunction(frozen(f_u), frozen(f_v), frozen(f_w));
}
VectorFunction TimeVectorFunction::get_difference_over_time(Real time_1, Real time_2) const {
  auto diff = [time_1, time_2](const Component &f) {
    return VectorFunction::Component([time_1, time_2, f](Real x, Real y, Real z) { return f(time_2, x, y, z) - f(time_1, x, y, z); });
  };
  return VectorFunction(diff(f_u), diff(f_v), diff(f_w));
}

// ---- boundary data for the C ABI ----------------------------------------------------------------------
namespace {

// The generated exact solutions are double functions whatever Real is (include/Manufactured.h), the lid-driven cases
// of TestCaseBoundaries.h are Real functions: the wrapped pointer is looked up with the type it was wrapped with.
typedef double (*TimeFn)(double, double, double, double);

template <class Fn>
bool is_function(const TimeVectorFunction::Component &f, Fn fn) {
  const Fn *target = f.template target<Fn>();
  return target && *target == fn;
}

// The analytic families libmifgpu evaluates on the device are recognised by the address of the functions the
// caller wrapped into the TimeVectorFunction; anything else is served through the host callback.
int detect_kind(const TimeVectorFunction &f) {
  if (is_function(f.f_u, u_exact) && is_function(f.f_v, v_exact) && is_function(f.f_w, w_exact)) return MIFGPU_BC_ETHIER_STEINMAN;
  if (is_function(f.f_u, exact_u_t1) && is_function(f.f_v, exact_v_t1) && is_function(f.f_w, exact_w_t1)) return MIFGPU_BC_TEST_CASE_1;
  if (is_function(f.f_u, exact_u_t2) && is_function(f.f_v, exact_v_t2) && is_function(f.f_w, exact_w_t2)) return MIFGPU_BC_TEST_CASE_2;
  if (is_function(f.f_u, u_exact_v_test) && is_function(f.f_v, v_exact_v_test) && is_function(f.f_w, w_exact_v_test))
    return MIFGPU_BC_VELOCITY_TEST;
  return MIFGPU_BC_HOST_CALLBACK;
}

struct FaceSource {
  const VelocityTensor *shape;                    // gives extents and staggered coordinates
  const TimeVectorFunction *velocity = nullptr;   // which = 0, time dependent
  const VectorFunction *velocity_fixed = nullptr; // which = 0, already frozen in time (VelocityTensor::apply_bc)
  const TimeVectorFunction *gradient = nullptr;   // which = 1 inside timestep_nhn
  const VectorFunction *gradient_fixed = nullptr; // which = 1 for solve_pressure_equation_non_homogeneous_neumann
};

// One face of one velocity component, i.e. what VelocityTensor::apply_bc stores there
// (src/VelocityTensor.cpp:47-217): the analytic value at the staggered point for tangential components; for the
// wall-normal component the wall value plus/minus half a cell of the tangential divergence of the analytic field.
void fill_velocity_face(const VelocityTensor &vt, const VectorFunction &f, int component, int face, mifgpu_real *values) {
  const Constants &c = vt.constants;
  const StaggeredTensor &tensor = *vt.components[component];
  const auto &s = tensor.sizes();
  const int dir = 2 - face / 2;
  const bool upper = face & 1;
  const size_t na = dir == 0 ? s[1] : s[0], nb = dir == 2 ? s[1] : s[2];
  const size_t wall[3] = {upper ? c.Nx - 1 : 0, upper ? c.Ny - 1 : 0, upper ? c.Nz - 1 : 0};
  const Real half[3] = {c.dx_over_2, c.dy_over_2, c.dz_over_2};
  const Real inv[3] = {c.one_over_dx, c.one_over_dy, c.one_over_dz};
  const int t1 = dir == 0 ? 1 : 0, t2 = dir == 2 ? 1 : 2;  // the two tangential directions, in the reference's order
  for (size_t b = 0; b < nb; b++)
    for (size_t a = 0; a < na; a++) {
      int idx[3];
      idx[dir] = static_cast<int>(wall[dir]);
      idx[t1] = static_cast<int>(a);
      idx[t2] = static_cast<int>(b);
      Real value;
      if (component == dir) {
        const Real at_wall = tensor.evaluate_function_at_index_unstaggered(idx[0], idx[1], idx[2], *f.components[dir]);
        Real divergence = 0;
        for (int t : {t1, t2}) {
          int next[3] = {idx[0], idx[1], idx[2]};
          next[t] += 1;
          const StaggeredTensor &other = *vt.components[t];
          divergence += (other.evaluate_function_at_index(next[0], next[1], next[2], *f.components[t]) -
                         other.evaluate_function_at_index(idx[0], idx[1], idx[2], *f.components[t])) * inv[t];
        }
        value = upper ? at_wall - half[dir] * divergence : at_wall + half[dir] * divergence;
      } else {
        value = tensor.evaluate_function_at_index(idx[0], idx[1], idx[2], *f.components[component]);
      }
      values[a + b * na] = value;
    }
}

// Neumann data on one face of the pressure tensor (src/PressureEquation.cpp:15-55): g_n at unstaggered points.
void fill_gradient_face(const Constants &c, const VectorFunction &g, int face, mifgpu_real *values) {
  const int dir = 2 - face / 2;
  const bool upper = face & 1;
  const size_t n[3] = {c.Nx, c.Ny, c.Nz};
  const size_t na = dir == 0 ? n[1] : n[0], nb = dir == 2 ? n[1] : n[2];
  const int t1 = dir == 0 ? 1 : 0, t2 = dir == 2 ? 1 : 2;
  const Real lo[3] = {c.min_x_global, c.min_y_global, c.min_z_global}, h[3] = {c.dx, c.dy, c.dz};
  const int base[3] = {c.base_i, c.base_j, c.base_k};
  for (size_t b = 0; b < nb; b++)
    for (size_t a = 0; a < na; a++) {
      int idx[3];
      idx[dir] = upper ? static_cast<int>(n[dir]) - 1 : 0;
      idx[t1] = static_cast<int>(a);
      idx[t2] = static_cast<int>(b);
      values[a + b * na] = (*g.components[dir])(lo[0] + h[0] * (base[0] + idx[0]), lo[1] + h[1] * (base[1] + idx[1]),
                                                lo[2] + h[2] * (base[2] + idx[2]));
    }
}

void face_callback(void *user, int which, double time, double time_prev, int component, int face, mifgpu_real *values) {
  const FaceSource &src = *static_cast<const FaceSource *>(user);
  if (which == 0) {
    if (src.velocity_fixed) fill_velocity_face(*src.shape, *src.velocity_fixed, component, face, values);
    else fill_velocity_face(*src.shape, src.velocity->set_time(time), component, face, values);
  } else if (src.gradient_fixed) {
    fill_gradient_face(src.shape->constants, *src.gradient_fixed, face, values);
  } else {
    // PRESSURE_EQUATION_true in src/Timestep.cpp:89-93 passes get_difference_over_time(new_time, prev_time).
    fill_gradient_face(src.shape->constants, src.gradient->get_difference_over_time(time, time_prev), face, values);
  }
}

void triple(const VelocityTensor &v, mifgpu_tensor *out[3]) {
  for (int c = 0; c < 3; c++) out[c] = v.components[c]->device();
}
void written(const VelocityTensor &v) {
  for (int c = 0; c < 3; c++) v.components[c]->device_was_written();
}

void run_timestep(VelocityTensor &velocity, VelocityTensor &velocity_buffer, VelocityTensor &velocity_buffer_2,
                  const TimeVectorFunction &exact_velocity, const TimeVectorFunction *exact_pressure_gradient, Real t_n,
                  StaggeredTensor &pressure, StaggeredTensor &pressure_buffer) {
  FaceSource source{&velocity};
  source.velocity = &exact_velocity;
  source.gradient = exact_pressure_gradient;
  mifgpu_bc bc{};
  bc.kind = detect_kind(exact_velocity);
  bc.Re = velocity.constants.Re;
  bc.callback = face_callback;
  bc.user = &source;
  mifgpu_tensor *v[3], *vb[3], *vb2[3];
  triple(velocity, v);
  triple(velocity_buffer, vb);
  triple(velocity_buffer_2, vb2);
  check(mifgpu_timestep(velocity.constants.gpu(), v, vb, vb2, &bc, t_n, pressure.device(), pressure_buffer.device(),
                        exact_pressure_gradient != nullptr),
        "mifgpu_timestep");
  written(velocity);
  written(velocity_buffer);
  written(velocity_buffer_2);
  pressure.device_was_written();
  pressure_buffer.device_was_written();
}

}  // namespace

// ---- VelocityTensor -----------------------------------------------------------------------------------
VelocityTensor::VelocityTensor(const Constants &constants_)
    : u(constants_), v(constants_), w(constants_), components({&this->u, &this->v, &this->w}), constants(constants_) {}

void VelocityTensor::swap_data(VelocityTensor &other) {
  for (size_t c = 0; c < 3U; c++) components[c]->swap_data(*other.components[c]);
}

void VelocityTensor::set(const VectorFunction &f, bool include_border) {
  for (size_t c = 0; c < 3U; c++) components[c]->set(*f.components[c], include_border);
}

void VelocityTensor::apply_bc(const VectorFunction &exact_velocity) {
  FaceSource source{this};
  source.velocity_fixed = &exact_velocity;
  mifgpu_bc bc{};
  bc.kind = MIFGPU_BC_HOST_CALLBACK;  // the function is already frozen in time: nothing to recognise
  bc.Re = constants.Re;
  bc.callback = face_callback;
  bc.user = &source;
  mifgpu_tensor *v3[3];
  triple(*this, v3);
  check(mifgpu_apply_bc(constants.gpu(), v3, &bc, 0.0), "mifgpu_apply_bc");
  written(*this);
}

// ---- Poisson solver objects ---------------------------------------------------------------------------
PressureSolverStructures::PressureSolverStructures(const Constants &constants_)
    : constants(constants_),
      periodic_bc{constants_.periodic_bc[0], constants_.periodic_bc[1], constants_.periodic_bc[2]},
      Nx_points(static_cast<int>(constants_.Nx_global) - (constants_.periodic_bc[0] ? 1 : 0)),
      Ny_points(static_cast<int>(constants_.Ny_global) - (constants_.periodic_bc[1] ? 1 : 0)),
      Nz_points(static_cast<int>(constants_.Nz_global) - (constants_.periodic_bc[2] ? 1 : 0)) {
  constants.gpu();  // builds the transform tables and eigenvalues now, like the reference's constructor
  const int ny = static_cast<int>(constants.Ny_owner), nz = static_cast<int>(constants.Nz_owner);
  xSize[0] = Nx_points; xSize[1] = ny; xSize[2] = nz;
  ySize[0] = Nx_points; ySize[1] = Ny_points; ySize[2] = nz;
  zSize[0] = Nx_points; zSize[1] = ny; zSize[2] = Nz_points;
}

PressureTensor::PressureTensor(PressureSolverStructures &structures_)
    : Tensor({structures_.Nx_points * static_cast<int>(structures_.constants.Ny_owner) * static_cast<int>(structures_.constants.Nz_owner)}),
      structures(structures_),
      max_size(structures_.Nx_points * static_cast<int>(structures_.constants.Ny_owner) * static_cast<int>(structures_.constants.Nz_owner)) {}

namespace {
// Owner points: everything except ghosts of other ranks and periodic images (include/StaggeredTensorMacros.h:41-83).
template <class F>
void for_each_owner_point(const StaggeredTensor &t, F &&body) {
  const Constants &c = t.constants;
  const auto &s = t.sizes();
  const size_t lo[3] = {c.periodic_bc[0] ? 1u : 0u, (c.prev_proc_y != -1 || c.periodic_bc[1]) ? 1u : 0u,
                        (c.prev_proc_z != -1 || c.periodic_bc[2]) ? 1u : 0u};
  const size_t hi[3] = {c.periodic_bc[0] ? s[0] - 1 : s[0], (c.next_proc_y != -1 || c.periodic_bc[1]) ? s[1] - 1 : s[1],
                        (c.next_proc_z != -1 || c.periodic_bc[2]) ? s[2] - 1 : s[2]};
  for (size_t k = lo[2]; k < hi[2]; k++)
    for (size_t j = lo[1]; j < hi[1]; j++)
      for (size_t i = lo[0]; i < hi[0]; i++) body(i, j, k);
}
}  // namespace

void PressureTensor::copy_from_staggered(const StaggeredTensor &other) {
  int index = 0;
  for_each_owner_point(other, [&](size_t i, size_t j, size_t k) { (*this)(index++) = other(i, j, k); });
}

void PressureTensor::copy_to_staggered(StaggeredTensor &other, int) const {
  int index = 0;
  for_each_owner_point(other, [&](size_t i, size_t j, size_t k) { other(i, j, k) = (*this)(index++); });
  other.apply_periodic_bc();
}

void PressureTensor::print_inline() const {
  for (int i = 0; i < max_size; i++) std::cout << (*this)(i) << " ";
  std::cout << std::endl;
}

void solve_pressure_equation_homogeneous_periodic(StaggeredTensor &pressure, PressureTensor &, const VelocityTensor &velocity, Real dt) {
  mifgpu_tensor *v[3];
  triple(velocity, v);
  check(mifgpu_solve_pressure(velocity.constants.gpu(), pressure.device(), v, dt, nullptr, 0.0), "mifgpu_solve_pressure");
  pressure.device_was_written();
}

void solve_pressure_equation_non_homogeneous_neumann(StaggeredTensor &pressure, PressureTensor &, const VelocityTensor &velocity,
                                                     const VectorFunction &exact_pressure_gradient, Real dt) {
  FaceSource source{&velocity};
  source.gradient_fixed = &exact_pressure_gradient;
  mifgpu_bc bc{};
  bc.kind = MIFGPU_BC_HOST_CALLBACK;
  bc.Re = velocity.constants.Re;
  bc.callback = face_callback;
  bc.user = &source;
  mifgpu_tensor *v[3];
  triple(velocity, v);
  check(mifgpu_solve_pressure(velocity.constants.gpu(), pressure.device(), v, dt, &bc, 0.0), "mifgpu_solve_pressure");
  pressure.device_was_written();
}

void adjust_pressure(StaggeredTensor &pressure, const std::function<Real(Real, Real, Real)> &exact_pressure) {
  const Constants &c = pressure.constants;
  Real difference = 0;
  for_each_owner_point(pressure, [&](size_t i, size_t j, size_t k) {
    difference += pressure.evaluate_function_at_index(i, j, k, exact_pressure) - pressure(i, j, k);
  });
  if (c.P > 1) {
    // rank 0 adds the ranks' sums in rank order and sends the result back (src/PressureEquation.cpp:298-333)
    std::vector<double> all(c.P, 0.0);
    std::vector<uint64_t> counts(c.P, 0);
    const double mine = difference;
    check(mifgpu_gather(c.gpu(), &mine, 1, all.data(), counts.data()), "mifgpu_gather");
    double total = 0.0;
    if (c.rank == 0)
      for (int r = 0; r < c.P; r++) total += all[r];
    check(mifgpu_allreduce(c.gpu(), &total, 1, 0), "mifgpu_allreduce");  // only rank 0 contributes: a broadcast
    difference = total;
  }
  const size_t nx = c.Nx_global - (c.periodic_bc[0] ? 1 : 0), ny = c.Ny_global - (c.periodic_bc[1] ? 1 : 0),
               nz = c.Nz_global - (c.periodic_bc[2] ? 1 : 0);
  difference /= static_cast<Real>(nx * ny * nz);
  for (size_t k = 0; k < c.Nz; k++)
    for (size_t j = 0; j < c.Ny; j++)
      for (size_t i = 0; i < c.Nx; i++) pressure(i, j, k) += difference;
}

// ---- time integrator ----------------------------------------------------------------------------------
void timestep(VelocityTensor &velocity, VelocityTensor &velocity_buffer, VelocityTensor &velocity_buffer_2,
              const TimeVectorFunction &exact_velocity, Real t_n, StaggeredTensor &pressure, StaggeredTensor &pressure_buffer,
              PressureTensor &) {
  run_timestep(velocity, velocity_buffer, velocity_buffer_2, exact_velocity, nullptr, t_n, pressure, pressure_buffer);
}

void timestep_nhn(VelocityTensor &velocity, VelocityTensor &velocity_buffer, VelocityTensor &velocity_buffer_2,
                  const TimeVectorFunction &exact_velocity, const TimeVectorFunction &exact_pressure_gradient, Real t_n,
                  StaggeredTensor &pressure, StaggeredTensor &pressure_buffer, PressureTensor &) {
  run_timestep(velocity, velocity_buffer, velocity_buffer_2, exact_velocity, &exact_pressure_gradient, t_n, pressure,
               pressure_buffer);
}

// src/TimestepVelocity.cpp:58-90.  The forcing is the reference's forcing_{x,y,z} (the only one its
// calculate_momentum_rhs_with_forcing_* can add), evaluated on the device; the boundary data is any TimeVectorFunction.
void timestep_velocity(VelocityTensor &velocity, VelocityTensor &velocity_buffer, VelocityTensor &rhs_buffer,
                       const TimeVectorFunction &exact_velocity, Real t_n) {
  FaceSource source{&velocity};
  source.velocity = &exact_velocity;
  mifgpu_bc bc{};
  bc.kind = detect_kind(exact_velocity);
  bc.Re = Reynolds;  // the global read by forcing_x/y/z in the reference (include/ManufacturedVelocity.h)
  bc.callback = face_callback;
  bc.user = &source;
  mifgpu_tensor *v[3], *vb[3], *rb[3];
  triple(velocity, v);
  triple(velocity_buffer, vb);
  triple(rhs_buffer, rb);
  check(mifgpu_timestep_velocity(velocity.constants.gpu(), v, vb, rb, &bc, t_n), "mifgpu_timestep_velocity");
  written(velocity);
  written(velocity_buffer);
  written(rhs_buffer);
}

// ---- norms (src/Norms.cpp) ----------------------------------------------------------------------------
namespace {
template <class Reduce>
Real velocity_error(const VelocityTensor &velocity, const TimeVectorFunction &exact, Real time, Reduce reduce) {
  const Constants &c = velocity.constants;
  Real acc = 0;
  // Components are averaged to the pressure points; wall points carry no error (Dirichlet data) and are skipped.
  for (size_t k = 1; k + 1 < c.Nz; k++) {
    const Real z = c.min_z_global + (c.base_k + static_cast<int>(k)) * c.dz;
    for (size_t j = 1; j + 1 < c.Ny; j++) {
      const Real y = c.min_y_global + (c.base_j + static_cast<int>(j)) * c.dy;
      for (size_t i = 1; i + 1 < c.Nx; i++) {
        const Real x = c.min_x_global + (c.base_i + static_cast<int>(i)) * c.dx;
        const Real eu = exact.f_u(time, x, y, z) - (velocity.u(i, j, k) + velocity.u(i + 1, j, k)) / 2.0;
        const Real ev = exact.f_v(time, x, y, z) - (velocity.v(i, j, k) + velocity.v(i, j + 1, k)) / 2.0;
        const Real ew = exact.f_w(time, x, y, z) - (velocity.w(i, j, k) + velocity.w(i, j, k + 1)) / 2.0;
        acc = reduce(acc, eu, ev, ew);
      }
    }
  }
  return acc;
}
template <class Reduce>
Real pressure_error(const StaggeredTensor &pressure, const std::function<Real(Real, Real, Real, Real)> &exact, Real time, Reduce reduce) {
  const Constants &c = pressure.constants;
  Real acc = 0;
  for_each_owner_point(pressure, [&](size_t i, size_t j, size_t k) {
    const Real x = c.min_x_global + (c.base_i + static_cast<int>(i)) * c.dx;
    const Real y = c.min_y_global + (c.base_j + static_cast<int>(j)) * c.dy;
    const Real z = c.min_z_global + (c.base_k + static_cast<int>(k)) * c.dz;
    acc = reduce(acc, exact(time, x, y, z) - pressure(i, j, k));
  });
  return acc;
}
Real cell_volume(const Constants &c) { return c.dx * c.dy * c.dz; }

// Analytic families that libmifgpu evaluates on the device take the device path: the tensors stay in HBM and only a
// few kB of per-CTA partial sums come back (mifgpu_velocity_error_norms / mifgpu_pressure_error_norms).  Anything
// else is reduced on the host after a download, literally as in src/Norms.cpp.
bool device_velocity_norms(const VelocityTensor &velocity, const TimeVectorFunction &exact, Real time, double norms[3]) {
  const int kind = detect_kind(exact);
  if (kind == MIFGPU_BC_HOST_CALLBACK) return false;
  mifgpu_bc bc{};
  bc.kind = kind;
  bc.Re = velocity.constants.Re;
  mifgpu_tensor *v[3];
  triple(velocity, v);
  check(mifgpu_velocity_error_norms(velocity.constants.gpu(), v, &bc, time, norms), "mifgpu_velocity_error_norms");
  return true;
}
bool device_pressure_norms(const StaggeredTensor &pressure, const std::function<Real(Real, Real, Real, Real)> &exact, Real time,
                           double norms[3]) {
  const TimeFn *target = exact.target<TimeFn>();
  if (!target || *target != p_exact) return false;
  mifgpu_bc bc{};
  bc.kind = MIFGPU_BC_ETHIER_STEINMAN;
  bc.Re = pressure.constants.Re;
  check(mifgpu_pressure_error_norms(pressure.constants.gpu(), pressure.device(), &bc, time, norms), "mifgpu_pressure_error_norms");
  return true;
}
}  // namespace

// Extension of the reference interface: adjust_pressure with the time-dependent exact pressure passed unfrozen, so
// that p_exact can be recognised and the whole operation stays on the device (the reference's signature, which takes
// a lambda frozen in time, is served above on the host).
void adjust_pressure(StaggeredTensor &pressure, const std::function<Real(Real, Real, Real, Real)> &exact_pressure, Real time) {
  const TimeFn *target = exact_pressure.target<TimeFn>();
  if (target && *target == p_exact) {
    mifgpu_bc bc{};
    bc.kind = MIFGPU_BC_ETHIER_STEINMAN;
    bc.Re = pressure.constants.Re;
    check(mifgpu_adjust_pressure(pressure.constants.gpu(), pressure.device(), &bc, time), "mifgpu_adjust_pressure");
    pressure.device_was_written();
    return;
  }
  adjust_pressure(pressure, [&exact_pressure, time](Real x, Real y, Real z) { return exact_pressure(time, x, y, z); });
}

Real ErrorL1Norm(const VelocityTensor &velocity, const TimeVectorFunction &exact_velocity, Real time) {
  double n[3];
  if (device_velocity_norms(velocity, exact_velocity, time, n)) return n[0];
  return velocity_error(velocity, exact_velocity, time, [](Real s, Real a, Real b, Real c) { return s + std::sqrt(a * a + b * b + c * c); }) *
         cell_volume(velocity.constants);
}
Real ErrorL2Norm(const VelocityTensor &velocity, const TimeVectorFunction &exact_velocity, Real time) {
  double n[3];
  if (device_velocity_norms(velocity, exact_velocity, time, n)) return n[1];
  return std::sqrt(velocity_error(velocity, exact_velocity, time, [](Real s, Real a, Real b, Real c) { return s + a * a + b * b + c * c; }) *
                   cell_volume(velocity.constants));
}
Real ErrorLInfNorm(const VelocityTensor &velocity, const TimeVectorFunction &exact_velocity, Real time) {
  double n[3];
  if (device_velocity_norms(velocity, exact_velocity, time, n)) return n[2];
  return velocity_error(velocity, exact_velocity, time,
                        [](Real s, Real a, Real b, Real c) { return std::max({s, std::abs(a), std::abs(b), std::abs(c)}); });
}
Real ErrorL1Norm(const StaggeredTensor &pressure, const std::function<Real(Real, Real, Real, Real)> &exact_pressure, Real time) {
  double n[3];
  if (device_pressure_norms(pressure, exact_pressure, time, n)) return n[0];
  return pressure_error(pressure, exact_pressure, time, [](Real s, Real e) { return s + std::abs(e); }) * cell_volume(pressure.constants);
}
Real ErrorL2Norm(const StaggeredTensor &pressure, const std::function<Real(Real, Real, Real, Real)> &exact_pressure, Real time) {
  double n[3];
  if (device_pressure_norms(pressure, exact_pressure, time, n)) return n[1];
  return std::sqrt(pressure_error(pressure, exact_pressure, time, [](Real s, Real e) { return s + e * e; }) * cell_volume(pressure.constants));
}
Real ErrorLInfNorm(const StaggeredTensor &pressure, const std::function<Real(Real, Real, Real, Real)> &exact_pressure, Real time) {
  double n[3];
  if (device_pressure_norms(pressure, exact_pressure, time, n)) return n[2];
  return pressure_error(pressure, exact_pressure, time, [](Real s, Real e) { return std::max(s, std::abs(e)); });
}
// One process drives the whole domain, so the rank-0 gather of the reference degenerates to the identity.
// src/Norms.cpp:120-162: the local errors travel to rank 0, which folds them in rank order with the reference's
// reduction (so the result has the reference's rounding); every other rank gets -1.
namespace {
template <typename Reduction>
Real accumulate_error(Real local_error, const Constants &constants, Reduction reduction) {
  if (constants.P == 1) return local_error;
  std::vector<double> all(constants.P, 0.0);
  std::vector<uint64_t> counts(constants.P, 0);
  const double mine = local_error;
  check(mifgpu_gather(constants.gpu(), &mine, 1, all.data(), counts.data()), "mifgpu_gather");
  if (constants.rank != 0) return -1;
  Real global_error = local_error;
  for (int r = 1; r < constants.P; r++) global_error = reduction(global_error, static_cast<Real>(all[r]));
  return global_error;
}
}  // namespace
Real accumulate_error_mpi_l1(Real local_error, const Constants &constants) {
  return accumulate_error(local_error, constants, [](Real global, Real other) { return global + other; });
}
Real accumulate_error_mpi_l2(Real local_error, const Constants &constants) {
  return accumulate_error(local_error, constants, [](Real global, Real other) { return std::sqrt(global * global + other * other); });
}
Real accumulate_error_mpi_linf(Real local_error, const Constants &constants) {
  return accumulate_error(local_error, constants, [](Real global, Real other) { return std::max(global, other); });
}

// ---- input file (src/InputParser.cpp) -----------------------------------------------------------------
void parse_input_file(const std::string &filename, size_t &Nx_global, size_t &Ny_global, size_t &Nz_global, Real &dt,
                      unsigned int &num_time_steps, int &Py, int &Pz, bool &test_case_2) {
  std::ifstream file(filename);
  if (!file.is_open()) throw std::runtime_error("Error opening input file: " + filename);
  const char *const names[8] = {"Nx", "Ny", "Nz", "dt", "Nt", "Py", "Pz", "test_case_2"};
  bool seen[8] = {};
  const std::string blanks = " \t\n\r\f\v";
  std::string line;
  while (std::getline(file, line)) {
    const size_t colon = line.find(':');
    if (colon == std::string::npos) continue;  // lines without a colon are ignored
    std::string key = line.substr(0, colon), value = line.substr(colon + 1);
    key.erase(key.find_last_not_of(blanks) + 1);          // trailing blanks of the key
    value.erase(0, value.find_first_not_of(blanks));      // leading blanks of the value
    int which = -1;
    for (int n = 0; n < 8; n++)
      if (key == names[n] && !seen[n]) which = n;
    if (which < 0) throw std::runtime_error("Unknown or duplicate key: " + key);
    seen[which] = true;
    switch (which) {
      case 0: Nx_global = std::stoul(value); break;
      case 1: Ny_global = std::stoul(value); break;
      case 2: Nz_global = std::stoul(value); break;
      case 3: dt = std::stod(value); break;
      case 4: num_time_steps = static_cast<unsigned int>(std::stoul(value)); break;
      case 5: Py = std::stoi(value); break;
      case 6: Pz = std::stoi(value); break;
      default: test_case_2 = (value == "true"); break;
    }
  }
  for (bool s : seen)
    if (!s) throw std::runtime_error("Missing one or more inputs. Required inputs: Nx, Ny, Nz, dt, Nt, Py, Pz, test_case_2.");
}

}  // namespace mif
