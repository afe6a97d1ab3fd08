// ManufacturedVelocity.h -- manufactured velocity field and momentum forcing of the velocity-only tests, same entry
// points as the reference's generated artifacts (include/ManufacturedVelocity.h, generators/manufsol_velocity.py:55-82):
// u = sin x cos y sin z sin t, v = cos x sin y sin z sin t, w = 2 cos x cos y cos z sin t and
// forcing = d_t c + (u . grad) c - lap(c) / Reynolds.  Written by hand in src/Manufactured.cpp.
#ifndef MANUFACTURED_VELOCITY_H
#define MANUFACTURED_VELOCITY_H

extern double Reynolds;  // defined by each executable, as in the reference

double u_exact_v_test(double t, double x, double y, double z);
double v_exact_v_test(double t, double x, double y, double z);
double w_exact_v_test(double t, double x, double y, double z);
double forcing_x(double t, double x, double y, double z);
double forcing_y(double t, double x, double y, double z);
double forcing_z(double t, double x, double y, double z);

#endif  // MANUFACTURED_VELOCITY_H
