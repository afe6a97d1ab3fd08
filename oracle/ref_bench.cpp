// oracle/ref_bench.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU timing of the UNMODIFIED reference library (oracle/_ref/libmif_ref.a) for bench.py's
// `cpu_baseline` / `--impl reference` legs.  Set-ups mirror the reference's own mains:
//
//   step    Nx Ny Nz steps warmup Pz    src/main.cpp:121-156 test case 1 (config 3 of BASELINE.json):
//                                       full mif::timestep (3 RK stages incl. 3 Poisson solves)
//   poisson Nx Ny Nz solves warmup Pz   test/pressure_test_mixed.cpp:31-60 (config 2): one
//                                       solve_pressure_equation_homogeneous_periodic per "solve"
//
// Ranks are threads of this process (MIF_SHIM_NP, oracle/shim/mpi.h).  Prints one JSON line on rank 0.
// Every number is "reference stencils/transposes + in-repo FFT (oracle/fft_cpu.h)": FFTW and MPI are
// not installed in this image.
#include <mpi.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "ManufacturedPressure.h"
#include "PressureEquation.h"
#include "TestCaseBoundaries.h"
#include "Timestep.h"

double Reynolds;

using namespace mif;

int main(int argc, char *argv[]) {
  int rank, size;
  MPI_Init(&argc, &argv);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  if (argc < 8) {
    if (rank == 0) std::fprintf(stderr, "usage: ref_bench step|poisson Nx Ny Nz iters warmup Pz\n");
    return 1;
  }
  const std::string mode = argv[1];
  const size_t Nx = std::atol(argv[2]), Ny = std::atol(argv[3]), Nz = std::atol(argv[4]);
  const int iters = std::atoi(argv[5]), warmup = std::atoi(argv[6]), Pz = std::atoi(argv[7]);
  const int Py = size / Pz;
  double seconds = 0.0;
  if (mode == "step") {
    constexpr Real Re = 1e3;
    const Real dt = 1e-3;
    const std::array<bool, 3> periodic{false, false, false};
    const Constants constants(Nx, Ny, Nz, 1.0, 1.0, 2.0, 0.0, 0.0, -1.0, Re, dt * (iters + warmup),
                              iters + warmup, Py, Pz, rank, periodic);
    PressureSolverStructures structures(constants);
    Reynolds = Re;
    VelocityTensor velocity(constants), velocity_buffer(constants), velocity_buffer_2(constants);
    StaggeredTensor pressure(constants, StaggeringDirection::none);
    StaggeredTensor pressure_buffer(constants, StaggeringDirection::none);
    PressureTensor solver_buffer(structures);
    TimeVectorFunction exact_velocity(exact_u_t1, exact_v_t1, exact_w_t1);
    velocity.set(exact_velocity.set_time(0.0), true);
    pressure.set(exact_p_initial_t1, true);
    for (int s = 0; s < warmup; s++)
      timestep(velocity, velocity_buffer, velocity_buffer_2, exact_velocity, s * constants.dt, pressure,
               pressure_buffer, solver_buffer);
    MPI_Barrier(MPI_COMM_WORLD);
    const double t0 = MPI_Wtime();
    for (int s = warmup; s < warmup + iters; s++)
      timestep(velocity, velocity_buffer, velocity_buffer_2, exact_velocity, s * constants.dt, pressure,
               pressure_buffer, solver_buffer);
    MPI_Barrier(MPI_COMM_WORLD);
    seconds = MPI_Wtime() - t0;
  } else {
    const std::array<bool, 3> periodic{false, false, true};
    const Constants constants(Nx, Ny, Nz, 2 * M_PI, 2 * M_PI, 2 * M_PI, 0.0, 0.0, 0.0, 1.0, 1.0, 1, Py, Pz, rank,
                              periodic);
    PressureSolverStructures structures(constants);
    VelocityTensor velocity(constants);
    PressureTensor solver_buffer(structures);
    StaggeredTensor pressure(constants, StaggeringDirection::none);
    TimeVectorFunction exact_velocity(u_exact_p_test, v_exact_p_test, w_exact_p_test);
    velocity.set(exact_velocity.set_time(1.0), true);
    for (int s = 0; s < warmup; s++)
      solve_pressure_equation_homogeneous_periodic(pressure, solver_buffer, velocity, constants.dt);
    MPI_Barrier(MPI_COMM_WORLD);
    const double t0 = MPI_Wtime();
    for (int s = 0; s < iters; s++)
      solve_pressure_equation_homogeneous_periodic(pressure, solver_buffer, velocity, constants.dt);
    MPI_Barrier(MPI_COMM_WORLD);
    seconds = MPI_Wtime() - t0;
  }
  if (rank == 0) {
    const double cells = double(Nx - 1) * double(Ny - 1) * double(Nz - 1);
    std::printf("{\"mode\": \"%s\", \"Nx\": %zu, \"Ny\": %zu, \"Nz\": %zu, \"iters\": %d, \"warmup\": %d, "
                "\"ranks\": %d, \"Py\": %d, \"Pz\": %d, \"seconds\": %.6f, \"cells\": %.0f, "
                "\"cell_iters_per_s\": %.6e}\n",
                mode.c_str(), Nx, Ny, Nz, iters, warmup, size, Py, Pz, seconds, cells,
                cells * iters / seconds);
  }
  MPI_Finalize();
  return 0;
}
