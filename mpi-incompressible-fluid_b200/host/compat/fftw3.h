// fftw3.h -- placeholder for sources written against the reference, which include <fftw3.h> although their own code never
// calls FFTW (test/full_test.cpp:2): the transforms of this implementation run on the GPU inside libmifgpu.
#ifndef MIF_COMPAT_FFTW3_H
#define MIF_COMPAT_FFTW3_H
#endif
