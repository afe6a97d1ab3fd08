// velocity_test -- the reference's velocity-only convergence tests (test/velocity_test.cpp, test/velocity_test_mixed.cpp)
// on the GPU path: manufactured field of ManufacturedVelocity.h with its forcing, no pressure, Re = 1e4, T = 1e-4.
//   usage: velocity_test N steps [Pz] [mixed]      mixed: periodic x and y on [0, 2 pi]^2 x [0, 1]
//   (several ranks: scripts/mifrun -n P velocity_test N steps Pz, Py = P / Pz)
// Prints the same three numbers: velocity L1 L2 Linf (rank 0).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "Launch.h"
#include "ManufacturedVelocity.h"
#include "Norms.h"
#include "TimestepVelocity.h"

double Reynolds;

int main(int argc, char *argv[]) {
  using namespace mif;
  if (argc < 3) {
    std::cerr << "usage: velocity_test N steps [mixed]" << std::endl;
    return 1;
  }
  const size_t N = std::atol(argv[1]);
  const unsigned int steps = std::atoi(argv[2]);
  const bool mixed = std::strcmp(argv[argc - 1], "mixed") == 0;
  // test/velocity_test.cpp: rank and size from the launcher, Pz from the command line, Py = size / Pz
  const int rank = launch_rank(), size = launch_size();
  const int Pz = (argc > 3 && std::strcmp(argv[3], "mixed") != 0) ? std::atoi(argv[3]) : 1;
  const int Py = Pz > 0 ? size / Pz : 0;
  if (Pz < 1 || Py < 1 || Py * Pz != size) {
    if (rank == 0) std::cerr << "velocity_test: Pz must divide the number of processes" << std::endl;
    return 1;
  }
  constexpr Real Re = 1e4, final_time = 1e-4;
  const Real len = mixed ? 2 * M_PI : 1.0;
  const Constants constants(N, N, N, len, len, 1.0, 0.0, 0.0, 0.0, Re, final_time, steps, Py, Pz, rank, {mixed, mixed, false});
  Reynolds = Re;

  VelocityTensor velocity(constants), velocity_buffer(constants), rhs_buffer(constants);
  TimeVectorFunction exact_velocity(u_exact_v_test, v_exact_v_test, w_exact_v_test);
  velocity.set(exact_velocity.set_time(0.0), true);
  for (unsigned int step = 0; step < steps; step++)
    timestep_velocity(velocity, velocity_buffer, rhs_buffer, exact_velocity, step * constants.dt);

  const Real l1 = accumulate_error_mpi_l1(ErrorL1Norm(velocity, exact_velocity, final_time), constants);
  const Real l2 = accumulate_error_mpi_l2(ErrorL2Norm(velocity, exact_velocity, final_time), constants);
  const Real linf = accumulate_error_mpi_linf(ErrorLInfNorm(velocity, exact_velocity, final_time), constants);
  if (rank == 0) std::cout << l1 << " " << l2 << " " << linf << std::endl;
  return 0;
}
