// TEST INFRASTRUCTURE ONLY: the constants of CUDA's math_constants.h that the kernels use.
#ifndef MIF_SIMT_EMU_MATH_CONSTANTS_H
#define MIF_SIMT_EMU_MATH_CONSTANTS_H
#include <math.h>
#define CUDART_PI 3.1415926535897931e+0
#define CUDART_PI_F 3.141592654F
#define CUDART_INF (__builtin_inf())
#define CUDART_NAN (__builtin_nan(""))
#endif
