// TestCaseBoundaries.h -- the two lid-type test cases of the reference driver (include/TestCaseBoundaries.h):
// zero velocity everywhere except v = 1 on the wall x = 1 (case 1) or x = -0.5 (case 2), zero initial pressure.
// These exact functions are recognised by address in mif::timestep and evaluated on the device.
#ifndef TEST_CASE_BOUNDARIES_H
#define TEST_CASE_BOUNDARIES_H

#include <cmath>

#include "Real.h"

namespace mif {

constexpr Real exact_solution_precision = 1e-12;

inline Real lid_profile(Real x, Real wall) { return std::abs(x - wall) < exact_solution_precision ? 1.0 : 0.0; }

inline Real exact_u_t1(Real, Real, Real, Real) { return 0.0; }
inline Real exact_v_t1(Real, Real x, Real, Real) { return lid_profile(x, 1.0); }
inline Real exact_w_t1(Real, Real, Real, Real) { return 0.0; }
inline Real exact_p_initial_t1(Real, Real, Real) { return 0.0; }

inline Real exact_u_t2(Real, Real, Real, Real) { return 0.0; }
inline Real exact_v_t2(Real, Real x, Real, Real) { return lid_profile(x, -0.5); }
inline Real exact_w_t2(Real, Real, Real, Real) { return 0.0; }
inline Real exact_p_initial_t2(Real, Real, Real) { return 0.0; }

}  // namespace mif

#endif  // TEST_CASE_BOUNDARIES_H
