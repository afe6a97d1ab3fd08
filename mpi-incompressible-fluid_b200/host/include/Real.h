// Real.h -- mirrors include/Real.h of the reference: `Real` is the floating-point type of every field.
// libmifgpu is an FP64 implementation (the reference's default and only tested build, USE_DOUBLE=1,
// CMakeLists.txt:29 / Makefile:29); a USE_DOUBLE=0 build is not provided.
#ifndef REAL_H
#define REAL_H

#if defined(USE_DOUBLE) && !USE_DOUBLE
#error "libmifgpu implements the FP64 build of mpi-incompressible-fluid only (USE_DOUBLE=1)"
#endif

#ifndef Real
#define Real double
#endif

#endif  // REAL_H
