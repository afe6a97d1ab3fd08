// export_test -- writes solution.vtk, full.vtk and three profiles from fields set analytically on the host (no GPU
// call is made), so the output writers and the input parser can be checked on a machine without a GPU.
//   usage: export_test N periodic_z outdir     |     export_test parse <input file>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "InputParser.h"
#include "VTKDatExport.h"

double Reynolds = 1.0;

int main(int argc, char *argv[]) {
  using namespace mif;
  if (argc == 3 && std::strcmp(argv[1], "parse") == 0) {
    size_t nx = 0, ny = 0, nz = 0;
    Real dt = 0;
    unsigned int nt = 0;
    int py = 0, pz = 0;
    bool tc2 = false;
    try {
      parse_input_file(argv[2], nx, ny, nz, dt, nt, py, pz, tc2);
    } catch (const std::exception &ex) {
      std::cout << "error: " << ex.what() << std::endl;
      return 0;
    }
    std::cout << nx << " " << ny << " " << nz << " " << dt << " " << nt << " " << py << " " << pz << " " << (tc2 ? 1 : 0) << std::endl;
    return 0;
  }
  if (argc != 4) return 1;
  const size_t N = std::atol(argv[1]);
  const bool periodic_z = std::atoi(argv[2]) != 0;
  const std::string out = argv[3];
  const Constants constants(N, N + 2, N + 1, 1.0, 1.0, 2.0, -0.25, -0.5, -1.0, 1.0, 1.0, 1, 1, 1, 0, {false, false, periodic_z});
  VelocityTensor velocity(constants);
  StaggeredTensor pressure(constants, StaggeringDirection::none);
  // quadratic fields with dyadic coefficients: every value is computed exactly the same way by any compiler
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
