// ManufacturedPressure.h -- manufactured pair of the stand-alone Poisson tests (include/ManufacturedPressure.h,
// generators/manufsol_pressure.py:36-39): p = t cos x cos y cos z and a velocity with div(u) = lap(p).
#ifndef MANUFACTURED_PRESSURE_H
#define MANUFACTURED_PRESSURE_H

double u_exact_p_test(double t, double x, double y, double z);
double v_exact_p_test(double t, double x, double y, double z);
double w_exact_p_test(double t, double x, double y, double z);
double p_exact_p_test(double t, double x, double y, double z);
double dp_dx_exact_p_test(double t, double x, double y, double z);
double dp_dy_exact_p_test(double t, double x, double y, double z);
double dp_dz_exact_p_test(double t, double x, double y, double z);

#endif  // MANUFACTURED_PRESSURE_H
