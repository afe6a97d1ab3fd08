// PressureSolverStructures.h -- drop-in for include/PressureSolverStructures.h:14-48.  In the reference this
// object owns the FFTW plans, the 2Decomp pencil decomposition and the 3-D eigenvalue table; here all of that
// lives inside the libmifgpu context of the Constants object (transform tables, three 1-D eigenvalue arrays),
// so this class only records the transform lengths and keeps the context alive.
#ifndef PRESSURE_SOLVER_STRUCTURES_H
#define PRESSURE_SOLVER_STRUCTURES_H

#include "Constants.h"

// Sources written against the reference get <mpi.h>, <iostream>, <string> and `using namespace std` through this header
// (include/PressureSolverStructures.h:9 pulls in deps/2Decomp_C/C2Decomp.hpp:4-14) and some rely on it
// (test/pressure_test_mixed.cpp uses MPI_Init, std::cout and an unqualified chrono:: with none of them included).  The
// build of the unchanged reference drivers (host/Makefile, bin/ref_*) defines MIF_REFERENCE_SOURCE_COMPAT to get the same
// environment; the host layer's own code does not.
#ifdef MIF_REFERENCE_SOURCE_COMPAT
#include <mpi.h>  // host/compat/mpi.h
#include <math.h>
#include <memory.h>

#include <cstdlib>
#include <iostream>
#include <string>
using namespace ::std;
#endif

namespace mif {

class PressureSolverStructures {
public:
  const Constants &constants;
  bool periodic_bc[3];
  const int Nx_points;  // transform length along x: N_global, or N_global - 1 if periodic
  const int Ny_points;
  const int Nz_points;
  // Local pencil sizes (2Decomp's xSize / ySize / zSize, deps/2Decomp_C/C2Decomp.hpp:81-83); on one rank all
  // three equal the global transform lengths.
  int xSize[3], ySize[3], zSize[3];

  explicit PressureSolverStructures(const Constants &constants);
  PressureSolverStructures(const PressureSolverStructures &) = delete;
};

}  // namespace mif

#endif  // PRESSURE_SOLVER_STRUCTURES_H
