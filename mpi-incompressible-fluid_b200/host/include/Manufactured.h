// Manufactured.h -- the exact Navier-Stokes solution of the full test (Ethier & Steinman, a = pi/4, d = pi/2),
// same entry points as the reference's generated artifacts (include/Manufactured.h:6-13,
// generators/manufsol.py:31-72).  Written by hand in src/Manufactured.cpp instead of sympy code generation.
#ifndef MANUFACTURED_H
#define MANUFACTURED_H

extern double Reynolds;  // defined by each executable, as in the reference (src/main.cpp:11)

double u_exact(double t, double x, double y, double z);
double v_exact(double t, double x, double y, double z);
double w_exact(double t, double x, double y, double z);
double p_exact(double t, double x, double y, double z);
double dp_dx_exact(double t, double x, double y, double z);
double dp_dy_exact(double t, double x, double y, double z);
double dp_dz_exact(double t, double x, double y, double z);

#endif  // MANUFACTURED_H
