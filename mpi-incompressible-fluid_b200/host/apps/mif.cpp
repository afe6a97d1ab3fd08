// mif -- the reference's driver (src/main.cpp) on the GPU path: `mif <input file>` with the reference's input format
// (Nt, dt, Nx, Ny, Nz, Py, Pz, test_case_2), Re = 1e3, test case 1 on [0,1]x[0,1]x[-1,1] (all walls) or test case 2 on
// [-0.5,0.5]^3 (z periodic); writes solution.vtk and profile1.dat, profile2.dat (and profile3.dat for case 2).
// One process drives one GPU; `scripts/mifrun -n P mif input.txt` starts Py * Pz = P of them (the reference's mpirun).
#include <cstdio>
#include <iostream>

#include "InputParser.h"
#include "Launch.h"
#include "PressureEquation.h"
#include "TestCaseBoundaries.h"
#include "Timestep.h"
#include "VTKDatExport.h"

double Reynolds;

int main(int argc, char *argv[]) {
  using namespace mif;
  if (argc != 2) {
    std::cerr << "Usage: ./mif [input parameter file]" << std::endl;
    return 1;
  }
  const int rank = launch_rank();  // MPI_Comm_rank (src/main.cpp:52-56): one process per GPU, started by scripts/mifrun
  size_t Nx_global = 0, Ny_global = 0, Nz_global = 0;
  Real dt = 0;
  unsigned int num_time_steps = 0;
  int Py = 0, Pz = 0;
  bool test_case_2 = false;
  try {
    parse_input_file(argv[1], Nx_global, Ny_global, Nz_global, dt, num_time_steps, Py, Pz, test_case_2);
    if (Pz < 1 || Py < 1) {
      std::cerr << "The number of processors in each direction must be at least 1." << std::endl;
      return 0;
    }
    if (Pz * Py != launch_size()) {
      if (rank == 0)
        std::cerr << "The number of precessors in the input file do not match with the ones provided to mpirun." << std::endl;
      return 0;
    }
  } catch (const std::exception &ex) {
    std::cerr << "Error parsing input file: " << ex.what() << std::endl;
    return 0;
  }

  constexpr Real Re = 1e3;
  const Constants constants(Nx_global, Ny_global, Nz_global, 1.0, 1.0, test_case_2 ? 1.0 : 2.0, test_case_2 ? -0.5 : 0.0,
                            test_case_2 ? -0.5 : 0.0, test_case_2 ? -0.5 : -1.0, Re, dt * num_time_steps, num_time_steps, Py, Pz,
                            rank, {false, false, test_case_2});
  PressureSolverStructures structures(constants);
  Reynolds = Re;

  VelocityTensor velocity(constants), velocity_buffer(constants), velocity_buffer_2(constants);
  StaggeredTensor pressure(constants, StaggeringDirection::none), pressure_buffer(constants, StaggeringDirection::none);
  PressureTensor pressure_solver_buffer(structures);

  TimeVectorFunction exact_velocity(test_case_2 ? exact_u_t2 : exact_u_t1, test_case_2 ? exact_v_t2 : exact_v_t1,
                                    test_case_2 ? exact_w_t2 : exact_w_t1);
  velocity.set(exact_velocity.set_time(0.0), true);
  pressure.set(test_case_2 ? exact_p_initial_t2 : exact_p_initial_t1, true);

  for (unsigned int time_step = 0; time_step < num_time_steps; time_step++)
    timestep(velocity, velocity_buffer, velocity_buffer_2, exact_velocity, time_step * constants.dt, pressure, pressure_buffer,
             pressure_solver_buffer);

  if (rank == 0)
    for (const char *old : {"profile1.dat", "profile2.dat", "profile3.dat", "solution.vtk"}) std::remove(old);
  writeVTK("solution.vtk", velocity, pressure);
  if (!test_case_2) {
    writeDat("profile1.dat", velocity, pressure, 1, 0.5, 0.5, 0);
    writeDat("profile2.dat", velocity, pressure, 0, 0.5, 0.5, 0);
  } else {
    writeDat("profile1.dat", velocity, pressure, 1, 0, 0, 0);
    writeDat("profile2.dat", velocity, pressure, 0, 0, 0, 0);
    writeDat("profile3.dat", velocity, pressure, 2, 0, 0, 0);
  }
  return 0;
}
