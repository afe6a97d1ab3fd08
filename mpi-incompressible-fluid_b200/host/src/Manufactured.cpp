// Manufactured.cpp -- closed-form manufactured solutions behind the reference's entry points
// (include/Manufactured.h, include/ManufacturedPressure.h).  The reference generates these functions with
// sympy (generators/manufsol.py, generators/manufsol_pressure.py); here they are written out by hand.
#include <cmath>

#include "Manufactured.h"
#include "ManufacturedPressure.h"

// `Reynolds` is defined by the executables that use it (src/main.cpp:11, test/full_test.cpp:13); the reference's pressure
// tests never define it because they link only the pressure artefact.  A weak definition here keeps those drivers
// linkable against the single host library; any executable's own definition takes precedence.
double Reynolds __attribute__((weak)) = 0.0;

namespace {
const double kA = M_PI / 4.0;  // Ethier-Steinman parameters a and d (generators/manufsol.py:31-32)
const double kD = M_PI / 2.0;

double decay(double t) { return std::exp(-kD * kD * t / Reynolds); }

// One velocity component; the other two follow by the cyclic permutation (x, y, z) -> (y, z, x).
double es_velocity(double t, double x, double y, double z) {
  return -kA * (std::exp(kA * x) * std::sin(kA * y + kD * z) + std::exp(kA * z) * std::cos(kA * x + kD * y)) * decay(t);
}

double es_pressure_dx(double t, double x, double y, double z) {
  const double a = kA, d = kD;
  const double s = 2 * a * std::exp(2 * a * x) +
                   2 * (a * std::cos(a * x + d * y) * std::cos(a * z + d * x) - d * std::sin(a * x + d * y) * std::sin(a * z + d * x)) * std::exp(a * (y + z)) +
                   2 * std::sin(a * y + d * z) * (a * std::cos(a * x + d * y) - a * std::sin(a * x + d * y)) * std::exp(a * (z + x)) +
                   2 * std::cos(a * y + d * z) * (d * std::cos(a * z + d * x) + a * std::sin(a * z + d * x)) * std::exp(a * (x + y));
  return -a * a / 2 * s * decay(t) * decay(t);
}
}  // namespace

double u_exact(double t, double x, double y, double z) { return es_velocity(t, x, y, z); }
double v_exact(double t, double x, double y, double z) { return es_velocity(t, y, z, x); }
double w_exact(double t, double x, double y, double z) { return es_velocity(t, z, x, y); }

double p_exact(double t, double x, double y, double z) {
  const double a = kA, d = kD;
  const double s = std::exp(2 * a * x) + std::exp(2 * a * y) + std::exp(2 * a * z) +
                   2 * std::sin(a * x + d * y) * std::cos(a * z + d * x) * std::exp(a * (y + z)) +
                   2 * std::sin(a * y + d * z) * std::cos(a * x + d * y) * std::exp(a * (z + x)) +
                   2 * std::sin(a * z + d * x) * std::cos(a * y + d * z) * std::exp(a * (x + y));
  return -a * a / 2 * s * decay(t) * decay(t);
}
// p is invariant under the same cyclic permutation, so its three partial derivatives are one function.
double dp_dx_exact(double t, double x, double y, double z) { return es_pressure_dx(t, x, y, z); }
double dp_dy_exact(double t, double x, double y, double z) { return es_pressure_dx(t, y, z, x); }
double dp_dz_exact(double t, double x, double y, double z) { return es_pressure_dx(t, z, x, y); }

// Poisson test pair: p = t cos x cos y cos z, velocity = grad p (so div u = lap p).
double u_exact_p_test(double t, double x, double y, double z) { return -std::sin(x) * std::cos(y) * std::cos(z) * t; }
double v_exact_p_test(double t, double x, double y, double z) { return -std::cos(x) * std::sin(y) * std::cos(z) * t; }
double w_exact_p_test(double t, double x, double y, double z) { return -std::cos(x) * std::cos(y) * std::sin(z) * t; }
double p_exact_p_test(double t, double x, double y, double z) { return std::cos(x) * std::cos(y) * std::cos(z) * t; }
double dp_dx_exact_p_test(double t, double x, double y, double z) { return u_exact_p_test(t, x, y, z); }
double dp_dy_exact_p_test(double t, double x, double y, double z) { return v_exact_p_test(t, x, y, z); }
double dp_dz_exact_p_test(double t, double x, double y, double z) { return w_exact_p_test(t, x, y, z); }

// Velocity-only tests (generators/manufsol_velocity.py:55-59) and the forcing that makes the field an exact solution
// of the momentum equation without pressure: f = d_t c + (u . grad) c - lap(c) / Re with lap(c) = -3 c.
#include "ManufacturedVelocity.h"
double u_exact_v_test(double t, double x, double y, double z) { return std::sin(x) * std::cos(y) * std::sin(z) * std::sin(t); }
double v_exact_v_test(double t, double x, double y, double z) { return std::cos(x) * std::sin(y) * std::sin(z) * std::sin(t); }
double w_exact_v_test(double t, double x, double y, double z) { return 2 * std::cos(x) * std::cos(y) * std::cos(z) * std::sin(t); }
double forcing_x(double t, double x, double y, double z) {
  const double st = std::sin(t), sx = std::sin(x), cx = std::cos(x), sy = std::sin(y), cy = std::cos(y), sz = std::sin(z);
  return sx * cy * sz * std::cos(t) + st * st * sx * cx * (2.0 - 2.0 * sy * sy - sz * sz) + 3.0 * sx * cy * sz * st / Reynolds;
}
double forcing_y(double t, double x, double y, double z) {
  const double st = std::sin(t), sx = std::sin(x), cx = std::cos(x), sy = std::sin(y), cy = std::cos(y), sz = std::sin(z);
  return cx * sy * sz * std::cos(t) + st * st * sy * cy * (2.0 - 2.0 * sx * sx - sz * sz) + 3.0 * cx * sy * sz * st / Reynolds;
}
double forcing_z(double t, double x, double y, double z) {
  const double st = std::sin(t), sx = std::sin(x), cx = std::cos(x), sy = std::sin(y), cy = std::cos(y), sz = std::sin(z), cz = std::cos(z);
  return 2.0 * cx * cy * cz * std::cos(t) + 2.0 * st * st * sz * cz * (sx * sx + sy * sy - 2.0) + 6.0 * cx * cy * cz * st / Reynolds;
}
