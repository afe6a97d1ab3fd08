// oracle/ref_dump.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Driver around the UNMODIFIED reference library (oracle/_ref/libmif_ref.a, compiled from
// /root/reference/src + deps/2Decomp_C by oracle/Makefile) that records raw FP64 fields, because the
// reference's own tests only print error norms (test/full_test.cpp:171-175), which are far too
// coarse for a 1e-11 parity contract.  The set-ups replicate the reference's mains:
//
//   full  N steps Pz out [nhn|hn [Ny Nz [pxyz]]]  test/full_test.cpp:36-76,118-126 (Ethier-Steinman; all walls unless the
//                                     optional mask, e.g. 010, makes directions periodic -- a parity set-up for the
//                                     periodic code paths on several ranks, not a physical one)
//   lid   Nx Ny Nz dt steps tc2 Pz out  src/main.cpp:121-156               (test case 1 / 2)
//   ptest kind Nx Ny Nz Pz out        test/pressure_test_{hn,mixed,nhn}.cpp:20-59 (any grid, kind = hn|mixed|nhn)
//   vtest N steps Pz out [mixed]      test/velocity_test{,_mixed}.cpp
//
// Every rank writes <out>/<field>_r<rank>.f64 (raw doubles, reference layout i + j*sx + k*sx*sy)
// and appends "<field> <rank> <sx> <sy> <sz>" lines to <out>/manifest_r<rank>.txt.
#include <mpi.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "Manufactured.h"
#include "ManufacturedPressure.h"
#include "ManufacturedVelocity.h"
#include "Norms.h"
#include "PressureEquation.h"
#include "TestCaseBoundaries.h"
#include "VTKDatExport.h"
#include "Timestep.h"

// include/TimestepVelocity.h shares its include guard with include/Timestep.h (both use
// TIMESTEP_VELOCITY_H), so the two cannot be included together; declare the one function we need.
namespace mif {
void timestep_velocity(VelocityTensor &velocity, VelocityTensor &velocity_buffer, VelocityTensor &rhs_buffer,
                       const TimeVectorFunction &exact_velocity, Real t_n);
}

double Reynolds;

using namespace mif;

// thread_local: under the MPI shim every rank is a thread of this process.
static thread_local std::string g_out;
static thread_local int g_rank = 0;

static void dump(const std::string &name, StaggeredTensor &t) {
  const auto &s = t.sizes();
  const std::string path = g_out + "/" + name + "_r" + std::to_string(g_rank) + ".f64";
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) { std::perror(path.c_str()); std::exit(2); }
  std::fwrite(t.raw_data(), sizeof(Real), s[0] * s[1] * s[2], f);
  std::fclose(f);
  const std::string mpath = g_out + "/manifest_r" + std::to_string(g_rank) + ".txt";
  FILE *m = std::fopen(mpath.c_str(), "a");
  std::fprintf(m, "%s %d %zu %zu %zu\n", name.c_str(), g_rank, s[0], s[1], s[2]);
  std::fclose(m);
}

static void dump_state(const std::string &suffix, VelocityTensor &vel, StaggeredTensor *p) {
  dump("u" + suffix, vel.u);
  dump("v" + suffix, vel.v);
  dump("w" + suffix, vel.w);
  if (p) dump("p" + suffix, *p);
}

static void reset_manifest() {
  const std::string mpath = g_out + "/manifest_r" + std::to_string(g_rank) + ".txt";
  std::remove(mpath.c_str());
}

static int run_full(int argc, char **argv, int size) {
  const size_t N = std::atol(argv[2]);
  const unsigned steps = std::atoi(argv[3]);
  const int Pz = std::atoi(argv[4]);
  g_out = argv[5];
  const bool nhn = argc > 6 && std::strcmp(argv[6], "nhn") == 0;
  // optional anisotropic grid (same physics as test/full_test.cpp, which is cubic): ... [nhn|hn] Ny Nz
  const size_t Ny = argc > 8 ? std::atol(argv[7]) : N, Nz = argc > 8 ? std::atol(argv[8]) : N;
  reset_manifest();
  const int Py = size / Pz;
  constexpr Real Re = 1e3;
  const char *mask = argc > 9 ? argv[9] : "000";
  const std::array<bool, 3> periodic{mask[0] == '1', mask[1] == '1', mask[2] == '1'};
  const Constants constants(N, Ny, Nz, 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, Re, 1e-4, steps, Py, Pz, g_rank, periodic);
  PressureSolverStructures structures(constants);
  Reynolds = Re;
  VelocityTensor velocity(constants), velocity_buffer(constants), velocity_buffer_2(constants);
  StaggeredTensor pressure(constants, StaggeringDirection::none);
  StaggeredTensor pressure_buffer(constants, StaggeringDirection::none);
  PressureTensor solver_buffer(structures);
  TimeVectorFunction exact_velocity(u_exact, v_exact, w_exact);
  TimeVectorFunction exact_pressure_gradient(dp_dx_exact, dp_dy_exact, dp_dz_exact);
  velocity.set(exact_velocity.set_time(0.0), true);
  const std::function<Real(Real, Real, Real)> p0 = [](Real x, Real y, Real z) { return p_exact(0.0, x, y, z); };
  pressure.set(p0, true);
  dump_state("_s0", velocity, &pressure);
  for (unsigned step = 0; step < steps; step++) {
    const Real t = step * constants.dt;
    if (nhn)
      timestep_nhn(velocity, velocity_buffer, velocity_buffer_2, exact_velocity, exact_pressure_gradient, t,
                   pressure, pressure_buffer, solver_buffer);
    else
      timestep(velocity, velocity_buffer, velocity_buffer_2, exact_velocity, t, pressure, pressure_buffer,
               solver_buffer);
    dump_state("_s" + std::to_string(step + 1), velocity, &pressure);
  }
  dump("dp_last", pressure_buffer);
  return 0;
}

static int run_lid(int argc, char **argv, int size) {
  const size_t Nx = std::atol(argv[2]), Ny = std::atol(argv[3]), Nz = std::atol(argv[4]);
  const Real dt = std::atof(argv[5]);
  const unsigned steps = std::atoi(argv[6]);
  const bool tc2 = std::atoi(argv[7]) != 0;
  const int Pz = std::atoi(argv[8]);
  g_out = argv[9];
  (void)argc;
  reset_manifest();
  const int Py = size / Pz;
  constexpr Real Re = 1e3;
  const std::array<bool, 3> periodic{false, false, tc2};
  const Constants constants(Nx, Ny, Nz, 1.0, 1.0, tc2 ? 1.0 : 2.0, tc2 ? -0.5 : 0.0, tc2 ? -0.5 : 0.0,
                            tc2 ? -0.5 : -1.0, Re, dt * steps, steps, Py, Pz, g_rank, periodic);
  PressureSolverStructures structures(constants);
  Reynolds = Re;
  VelocityTensor velocity(constants), velocity_buffer(constants), velocity_buffer_2(constants);
  StaggeredTensor pressure(constants, StaggeringDirection::none);
  StaggeredTensor pressure_buffer(constants, StaggeringDirection::none);
  PressureTensor solver_buffer(structures);
  TimeVectorFunction exact_velocity(tc2 ? exact_u_t2 : exact_u_t1, tc2 ? exact_v_t2 : exact_v_t1,
                                    tc2 ? exact_w_t2 : exact_w_t1);
  velocity.set(exact_velocity.set_time(0.0), true);
  pressure.set(tc2 ? exact_p_initial_t2 : exact_p_initial_t1, true);
  dump_state("_s0", velocity, &pressure);
  for (unsigned step = 0; step < steps; step++) {
    timestep(velocity, velocity_buffer, velocity_buffer_2, exact_velocity, step * constants.dt, pressure,
             pressure_buffer, solver_buffer);
    dump_state("_s" + std::to_string(step + 1), velocity, &pressure);
  }
  return 0;
}

static int run_ptest(int argc, char **argv, int size) {
  const std::string kind = argv[2];
  const size_t Nx = std::atol(argv[3]), Ny = std::atol(argv[4]), Nz = std::atol(argv[5]);
  const int Pz = std::atoi(argv[6]);
  g_out = argv[7];
  (void)argc;
  reset_manifest();
  const int Py = size / Pz;
  const bool nhn = kind == "nhn";
  const Real lo = nhn ? -M_PI / 2.0 : 0.0;
  const Real len = nhn ? M_PI / 2.0 : 2 * M_PI;
  const std::array<bool, 3> periodic{false, false, kind == "mixed"};
  constexpr Real time = 1.0;
  const Constants constants(Nx, Ny, Nz, len, len, len, lo, lo, lo, 1.0, 1.0, 1, Py, Pz, g_rank, periodic);
  PressureSolverStructures structures(constants);
  VelocityTensor velocity(constants);
  PressureTensor solver_buffer(structures);
  StaggeredTensor pressure(constants, StaggeringDirection::none);
  TimeVectorFunction exact_velocity(u_exact_p_test, v_exact_p_test, w_exact_p_test);
  velocity.set(exact_velocity.set_time(time), true);
  dump_state("_in", velocity, nullptr);
  if (nhn) {
    TimeVectorFunction g_t(dp_dx_exact_p_test, dp_dy_exact_p_test, dp_dz_exact_p_test);
    solve_pressure_equation_non_homogeneous_neumann(pressure, solver_buffer, velocity, g_t.set_time(time),
                                                    constants.dt);
  } else {
    solve_pressure_equation_homogeneous_periodic(pressure, solver_buffer, velocity, constants.dt);
  }
  dump("p_out", pressure);
  adjust_pressure(pressure, [](Real x, Real y, Real z) { return p_exact_p_test(time, x, y, z); });
  const Real l1 = accumulate_error_mpi_l1(ErrorL1Norm(pressure, p_exact_p_test, time), constants);
  const Real l2 = accumulate_error_mpi_l2(ErrorL2Norm(pressure, p_exact_p_test, time), constants);
  const Real li = accumulate_error_mpi_linf(ErrorLInfNorm(pressure, p_exact_p_test, time), constants);
  if (g_rank == 0) std::printf("Errors: %.17g %.17g %.17g\n", l1, l2, li);
  return 0;
}

static int run_vtest(int argc, char **argv, int size) {
  const size_t N = std::atol(argv[2]);
  const unsigned steps = std::atoi(argv[3]);
  const int Pz = std::atoi(argv[4]);
  g_out = argv[5];
  const bool mixed = argc > 6 && std::strcmp(argv[6], "mixed") == 0;
  reset_manifest();
  const int Py = size / Pz;
  constexpr Real Re = 1e4;
  const std::array<bool, 3> periodic{mixed, mixed, false};
  const Real len = mixed ? 2 * M_PI : 1.0;
  const Constants constants(N, N, N, len, len, 1.0, 0.0, 0.0, 0.0, Re, 1e-4, steps, Py, Pz, g_rank, periodic);
  Reynolds = Re;
  VelocityTensor velocity(constants), velocity_buffer(constants), rhs_buffer(constants);
  TimeVectorFunction exact_velocity(u_exact_v_test, v_exact_v_test, w_exact_v_test);
  velocity.set(exact_velocity.set_time(0.0), true);
  dump_state("_s0", velocity, nullptr);
  for (unsigned step = 0; step < steps; step++) {
    timestep_velocity(velocity, velocity_buffer, rhs_buffer, exact_velocity, step * constants.dt);
    dump_state("_s" + std::to_string(step + 1), velocity, nullptr);
  }
  return 0;
}

// export N periodic_z out: the reference's output writers (src/VTKDatExport.cpp) on fields set analytically, the
// counterpart of mpi-incompressible-fluid_b200/host/apps/export_test.cpp.
static int run_export(int argc, char **argv) {
  (void)argc;
  const size_t N = std::atol(argv[2]);
  const bool periodic_z = std::atoi(argv[3]) != 0;
  const std::string out = argv[4];
  const std::array<bool, 3> periodic{false, false, periodic_z};
  const Constants constants(N, N + 2, N + 1, 1.0, 1.0, 2.0, -0.25, -0.5, -1.0, 1.0, 1.0, 1, 1, 1, 0, periodic);
  VelocityTensor velocity(constants);
  StaggeredTensor pressure(constants, StaggeringDirection::none);
  velocity.u.set([](Real x, Real y, Real z) { return 0.5 * x + 0.25 * y * z - 2.0 * z; }, true);
  velocity.v.set([](Real x, Real y, Real z) { return x * y - 0.125 * z + 1.0; }, true);
  velocity.w.set([](Real x, Real y, Real z) { return 4.0 * x - y + 0.5 * z * z; }, true);
  pressure.set([](Real x, Real y, Real z) { return x * x - 0.5 * y + 0.25 * z; }, true);
  writeVTK(out + "/solution.vtk", velocity, pressure);
  writeVTKFullMesh(out + "/full.vtk", velocity, pressure);
  writeDat(out + "/profile_y.dat", velocity, pressure, 1, 0.25, 0.0, 0.0);
  writeDat(out + "/profile_x.dat", velocity, pressure, 0, 0.0, 0.0, 0.0);
  writeDat(out + "/profile_z.dat", velocity, pressure, 2, 0.25, 0.0, 0.0);
  return 0;
}

int main(int argc, char *argv[]) {
  int size;
  MPI_Init(&argc, &argv);
  MPI_Comm_rank(MPI_COMM_WORLD, &g_rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  if (argc < 2) {
    std::fprintf(stderr, "usage: ref_dump full|lid|ptest|vtest ...\n");
    return 1;
  }
  const std::string mode = argv[1];
  int rc = 1;
  if (mode == "full" && argc >= 6) rc = run_full(argc, argv, size);
  else if (mode == "lid" && argc >= 10) rc = run_lid(argc, argv, size);
  else if (mode == "ptest" && argc >= 8) rc = run_ptest(argc, argv, size);
  else if (mode == "vtest" && argc >= 6) rc = run_vtest(argc, argv, size);
  else if (mode == "export" && argc >= 5) rc = run_export(argc, argv);
  else std::fprintf(stderr, "ref_dump: bad arguments\n");
  MPI_Finalize();
  return rc;
}
