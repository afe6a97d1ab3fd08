// full_test -- the reference's velocity + pressure convergence test (test/full_test.cpp) on the GPU path:
// Ethier-Steinman solution on [0,1]x[0,1]x[-1,1], all walls, Re = 1e3, T = 1e-4.
//   usage: full_test N steps [Pz] [nhn]        (several ranks: scripts/mifrun -n P full_test N steps Pz, Py = P / Pz)
// Prints the same nine numbers: velocity L1 L2 Linf, pressure L1 L2 Linf, pressure-gradient L1 L2 Linf (rank 0).
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "Launch.h"
#include "Manufactured.h"
#include "Norms.h"
#include "PressureEquation.h"
#include "Timestep.h"

double Reynolds;

int main(int argc, char *argv[]) {
  using namespace mif;
  if (argc < 3) {
    std::cerr << "usage: full_test N steps [nhn]" << std::endl;
    return 1;
  }
  const size_t N = std::atol(argv[1]);
  const unsigned int steps = std::atoi(argv[2]);
  const bool nhn = std::strcmp(argv[argc - 1], "nhn") == 0;
  // test/full_test.cpp:24-28,55-59: rank and size from the launcher, Pz from the command line, Py = size / Pz
  const int rank = launch_rank(), size = launch_size();
  const int Pz = (argc > 3 && std::strcmp(argv[3], "nhn") != 0) ? std::atoi(argv[3]) : 1;
  const int Py = Pz > 0 ? size / Pz : 0;
  if (Pz < 1 || Py < 1 || Py * Pz != size) {
    if (rank == 0) std::cerr << "full_test: Pz must divide the number of processes" << std::endl;
    return 1;
  }
  constexpr Real Re = 1e3, final_time = 1e-4;
  const Constants constants(N, N, N, 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, Re, final_time, steps, Py, Pz, rank, {false, false, false});
  PressureSolverStructures structures(constants);
  Reynolds = Re;

  VelocityTensor velocity(constants), velocity_buffer(constants), velocity_buffer_2(constants);
  StaggeredTensor pressure(constants, StaggeringDirection::none), pressure_buffer(constants, StaggeringDirection::none);
  PressureTensor pressure_solver_buffer(structures);

  TimeVectorFunction exact_velocity(u_exact, v_exact, w_exact);
  TimeVectorFunction exact_pressure_gradient(dp_dx_exact, dp_dy_exact, dp_dz_exact);
  velocity.set(exact_velocity.set_time(0.0), true);
  pressure.set([](Real x, Real y, Real z) { return p_exact(0.0, x, y, z); }, true);

  for (unsigned int step = 0; step < steps; step++) {
    const Real t = step * constants.dt;
    if (nhn) timestep_nhn(velocity, velocity_buffer, velocity_buffer_2, exact_velocity, exact_pressure_gradient, t, pressure, pressure_buffer, pressure_solver_buffer);
    else timestep(velocity, velocity_buffer, velocity_buffer_2, exact_velocity, t, pressure, pressure_buffer, pressure_solver_buffer);
  }

  // Discrete pressure gradient on the interior staggered points, exact data on the faces.
  VelocityTensor pressure_gradient(constants);
  for (int c = 0; c < 3; c++) {
    StaggeredTensor &g = *pressure_gradient.components[c];
    const auto &s = g.sizes();
    const Real inv_h = c == 0 ? constants.one_over_dx : (c == 1 ? constants.one_over_dy : constants.one_over_dz);
    for (size_t k = 1; k + 1 < s[2]; k++)
      for (size_t j = 1; j + 1 < s[1]; j++)
        for (size_t i = 1; i + 1 < s[0]; i++)
          g(i, j, k) = (pressure(i, j, k) - pressure(i - (c == 0), j - (c == 1), k - (c == 2))) * inv_h;
  }
  pressure_gradient.apply_bc(exact_pressure_gradient.set_time(final_time));

  adjust_pressure(pressure, [](Real x, Real y, Real z) { return p_exact(1e-4, x, y, z); });

  // test/full_test.cpp:150-176: local norms, folded on rank 0
  const Real errors[9] = {
      accumulate_error_mpi_l1(ErrorL1Norm(velocity, exact_velocity, final_time), constants),
      accumulate_error_mpi_l2(ErrorL2Norm(velocity, exact_velocity, final_time), constants),
      accumulate_error_mpi_linf(ErrorLInfNorm(velocity, exact_velocity, final_time), constants),
      accumulate_error_mpi_l1(ErrorL1Norm(pressure, p_exact, final_time), constants),
      accumulate_error_mpi_l2(ErrorL2Norm(pressure, p_exact, final_time), constants),
      accumulate_error_mpi_linf(ErrorLInfNorm(pressure, p_exact, final_time), constants),
      accumulate_error_mpi_l1(ErrorL1Norm(pressure_gradient, exact_pressure_gradient, final_time), constants),
      accumulate_error_mpi_l2(ErrorL2Norm(pressure_gradient, exact_pressure_gradient, final_time), constants),
      accumulate_error_mpi_linf(ErrorLInfNorm(pressure_gradient, exact_pressure_gradient, final_time), constants)};
  if (rank == 0) {
    for (int e = 0; e < 9; e++) std::cout << errors[e] << (e < 8 ? " " : "");
    std::cout << std::endl;
  }
  return 0;
}
