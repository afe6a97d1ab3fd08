// pressure_test -- the reference's stand-alone Poisson tests (test/pressure_test_hn.cpp, _mixed.cpp, _nhn.cpp) on the
// GPU path: one solve on an N x 3N x 5N grid with p = t cos x cos y cos z, t = 1.
//   usage: pressure_test hn|mixed|nhn N [Pz]      (several ranks: scripts/mifrun -n P pressure_test hn N Pz, Py = P / Pz)
// Prints "Errors: L1 L2 Linf" like the reference.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>

#include "Launch.h"
#include "ManufacturedPressure.h"
#include "Norms.h"
#include "PressureEquation.h"

double Reynolds = 1.0;

int main(int argc, char *argv[]) {
  using namespace mif;
  if (argc < 3) {
    std::cerr << "usage: pressure_test hn|mixed|nhn N" << std::endl;
    return 1;
  }
  const std::string kind = argv[1];
  const size_t N = std::atol(argv[2]);
  const bool nhn = kind == "nhn";
  const Real lo = nhn ? -M_PI / 2.0 : 0.0, length = nhn ? M_PI / 2.0 : 2 * M_PI;
  constexpr Real time = 1.0;
  // test/pressure_test_*.cpp:16-36: rank and size from the launcher, Pz from the command line, Py = size / Pz
  const int rank = launch_rank(), size = launch_size();
  const int Pz = argc > 3 ? std::atoi(argv[3]) : 1;
  const int Py = Pz > 0 ? size / Pz : 0;
  if (Pz < 1 || Py < 1 || Py * Pz != size) {
    if (rank == 0) std::cerr << "pressure_test: Pz must divide the number of processes" << std::endl;
    return 1;
  }
  const Constants constants(N, 3 * N, 5 * N, length, length, length, lo, lo, lo, 1.0, 1.0, 1, Py, Pz, rank,
                            {false, false, kind == "mixed"});
  PressureSolverStructures structures(constants);
  VelocityTensor velocity(constants);
  PressureTensor pressure_solver_buffer(structures);
  StaggeredTensor pressure(constants, StaggeringDirection::none);

  TimeVectorFunction exact_velocity(u_exact_p_test, v_exact_p_test, w_exact_p_test);
  velocity.set(exact_velocity.set_time(time), true);

  const auto before = std::chrono::high_resolution_clock::now();
  if (nhn) {
    TimeVectorFunction gradient(dp_dx_exact_p_test, dp_dy_exact_p_test, dp_dz_exact_p_test);
    solve_pressure_equation_non_homogeneous_neumann(pressure, pressure_solver_buffer, velocity, gradient.set_time(time), constants.dt);
  } else {
    solve_pressure_equation_homogeneous_periodic(pressure, pressure_solver_buffer, velocity, constants.dt);
  }
  pressure.sync_host();
  const Real seconds = std::chrono::duration<Real>(std::chrono::high_resolution_clock::now() - before).count();

  adjust_pressure(pressure, [](Real x, Real y, Real z) { return p_exact_p_test(1.0, x, y, z); });
  const Real l1 = accumulate_error_mpi_l1(ErrorL1Norm(pressure, p_exact_p_test, time), constants);
  const Real l2 = accumulate_error_mpi_l2(ErrorL2Norm(pressure, p_exact_p_test, time), constants);
  const Real linf = accumulate_error_mpi_linf(ErrorLInfNorm(pressure, p_exact_p_test, time), constants);
  if (rank == 0) {
    std::cout << "Time: " << seconds << "s " << seconds / N / (3 * N) / (5 * N) << std::endl;
    std::cout << "Errors: " << l1 << " " << l2 << " " << linf << std::endl;
  }
  return 0;
}
