// VTKDatExport.h -- drop-in for include/VTKDatExport.h:10-27: the output entry points of the reference driver.
// Host-side post-processing of the downloaded fields; file formats are those of src/VTKDatExport.cpp.
#ifndef VTK_EXPORT_H
#define VTK_EXPORT_H

#include <string>

#include "VelocityTensor.h"

namespace mif {

// Legacy binary VTK (big-endian doubles, DATASET UNSTRUCTURED_GRID): the planes z = 0, x = 0, y = 0 as points with
// the scalars u, v, w (averaged to the pressure points) and p (src/VTKDatExport.cpp:115-312).
void writeVTK(const std::string &filename, const VelocityTensor &velocity, const StaggeredTensor &pressure);

// Text profile along `direction` (0 = x, 1 = y, 2 = z) through the point (x, y, z), one row per pressure point:
// "x y z u v w p" formatted "%.8f %.8f %.8f %.8e %.8e %.8e %.8e" (src/VTKDatExport.cpp:342-583).
void writeDat(const std::string &filename, const VelocityTensor &velocity, const StaggeredTensor &pressure,
              const int direction, const Real x, const Real y, const Real z);

// ASCII STRUCTURED_POINTS dump of the whole pressure mesh with u, v, w, |u|, p (src/VTKDatExport.cpp:586-669).
void writeVTKFullMesh(const std::string &filename, const mif::VelocityTensor &velocity, const StaggeredTensor &pressure);

}  // namespace mif

#endif  // VTK_EXPORT_H
