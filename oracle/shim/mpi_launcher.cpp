// oracle/shim/mpi_launcher.cpp -- TEST INFRASTRUCTURE ONLY.
// `main` of every executable built against the MPI shim.  The reference's own `main` functions
// (test/full_test.cpp, test/pressure_test_*.cpp, src/main.cpp, ...) are compiled with
// their `main` symbol renamed to mifshim_user_main and run here once per rank, each rank on its own thread
// (MIF_SHIM_NP=<ranks>, default 1) -- the stand-in for `mpirun -n <ranks>`.  The rename is done on the
// object file (objcopy --redefine-sym), so the symbol has C linkage like `main` itself.
#include <cstdlib>

#include "mpi.h"

extern "C" int mifshim_user_main(int argc, char *argv[]);

int main(int argc, char *argv[]) {
  const char *np_env = std::getenv("MIF_SHIM_NP");
  const int np = np_env ? std::atoi(np_env) : 1;
  return mifshim_run(np > 0 ? np : 1, mifshim_user_main, argc, argv);
}
