// Real.h -- mirrors include/Real.h of the reference (:9-17): `Real` is the floating-point type of every field.
// USE_DOUBLE=1 (the reference's default, CMakeLists.txt:29 / Makefile:29) binds libmifgpu.so; USE_DOUBLE=0 makes Real a
// float and binds libmifgpu_f32.so, the float build of the same device code (MIFGPU_FP32 selects the float field type
// in ../../include/mifgpu.h).
#ifndef REAL_H
#define REAL_H

#if defined(USE_DOUBLE) && !USE_DOUBLE
#undef Real
#define Real float
#ifndef MIFGPU_FP32
#define MIFGPU_FP32 1
#endif
#else
#ifndef Real
#define Real double
#endif
#endif

#endif  // REAL_H
