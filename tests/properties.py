"""Size-independent properties of the hot path, written in numpy and used two ways:

* on the CPU (`tests/test_properties_cpu.py`) against the oracle at small sizes -- this validates the CHECKERS;
* on the GPU (`tests/test_gpu_zz_full_size.py`) against libmifgpu at BASELINE.json's full sizes (513^3 points), where
  the oracle would need minutes to hours per call.

All arrays use the reference layout viewed as numpy (sz, sy, sx), ghosts included.  `periodic` is (x, y, z).
The properties follow from the reference's definitions:

* solve_pressure_equation (src/PressureEquation.cpp:65-264) diagonalises the 7-point Laplacian whose boundary rows are
  the even (mirror) extension in Neumann directions and the wrap-around in periodic directions
  (eigenvalues src/PressureSolverStructures.cpp:5-11), and zeroes mode (0,0,0) (:161-163).  Hence
  L p = rhs - <rhs> on every owner point, with <.> the DCT-I / DFT mode-0 mean (end points weigh 1/2 in Neumann
  directions), and <p> = 0.
* the solve is linear in the velocity and commutes with shifts along a periodic direction;
* test case 1 (include/TestCaseBoundaries.h:17-35; domain z in [-1, 1]) is mirror symmetric in z: u, v, p even, w odd.
"""
import numpy as np


def owner_slices(shape_p, periodic):
    """Owner points of the pressure tensor on one rank: everything in a non-periodic direction, 1..s-2 in a periodic one
    (include/StaggeredTensorMacros.h:41-83 for a single rank).  Returned in numpy order (z, y, x)."""
    sz, sy, sx = shape_p
    per_x, per_y, per_z = periodic
    return (slice(1, sz - 1) if per_z else slice(0, sz), slice(1, sy - 1) if per_y else slice(0, sy),
            slice(1, sx - 1) if per_x else slice(0, sx))


def _shifted(sl, by):
    return slice(sl.start + by, sl.stop + by)


def poisson_rhs(u, v, w, dt, h, shape_p, periodic):
    """div(velocity) / dt on the owner points (include/VelocityDivergence.h:9-20, src/PressureEquation.cpp:59-61)."""
    kz, jy, ix = owner_slices(shape_p, periodic)
    rhs = (u[kz, jy, _shifted(ix, 1)] - u[kz, jy, ix]) / h[0]
    rhs += (v[kz, _shifted(jy, 1), ix] - v[kz, jy, ix]) / h[1]
    rhs += (w[_shifted(kz, 1), jy, ix] - w[kz, jy, ix]) / h[2]
    rhs /= dt
    return rhs


def laplacian(p_own, h, periodic):
    """The operator the spectral solve inverts, applied to the owner block (numpy axes 0, 1, 2 = z, y, x)."""
    out = np.zeros_like(p_own)
    for axis, (hd, per) in enumerate(zip((h[2], h[1], h[0]), (periodic[2], periodic[1], periodic[0]))):
        q = np.moveaxis(p_own, axis, 0)
        o = np.moveaxis(out, axis, 0)
        inv = 1.0 / (hd * hd)
        if per:
            o += (np.roll(q, 1, axis=0) + np.roll(q, -1, axis=0) - 2.0 * q) * inv
        else:
            o[1:-1] += (q[2:] + q[:-2] - 2.0 * q[1:-1]) * inv
            o[0] += 2.0 * (q[1] - q[0]) * inv
            o[-1] += 2.0 * (q[-2] - q[-1]) * inv
    return out


def mode0_mean(a_own, periodic):
    """Mean with the weights of transform mode 0: 1/2 at both ends of a Neumann direction, uniform in a periodic one."""
    m = a_own
    for axis, per in zip((0, 1, 2), (periodic[2], periodic[1], periodic[0])):
        n = m.shape[0]
        if per:
            m = m.sum(axis=0) / n
        else:
            m = (m.sum(axis=0) - 0.5 * (m[0] + m[-1])) / (n - 1)
    return float(m)


def poisson_defects(p, u, v, w, dt, h, periodic):
    """(residual, gauge): max |L p - (rhs - <rhs>)| / max |rhs|  and  |<p>| / max |p| of one pressure solve."""
    own = owner_slices(p.shape, periodic)
    rhs = poisson_rhs(u, v, w, dt, h, p.shape, periodic)
    p_own = p[own]
    res = laplacian(p_own, h, periodic)
    res -= rhs
    res += mode0_mean(rhs, periodic)
    scale = float(np.max(np.abs(rhs)))
    return float(np.max(np.abs(res))) / scale, abs(mode0_mean(p_own, periodic)) / float(np.max(np.abs(p_own)))


def rel_diff(a, b):
    return float(np.max(np.abs(a - b))) / max(float(np.max(np.abs(b))), 1e-300)


def periodic_z_field(base, sz, shift=0):
    """Tensor with sz planes built from `base` (n planes, the period): plane k holds base[(k - 1 + shift) mod n], so
    the ghost planes are true periodic images and a shift is a pure re-indexing."""
    n = base.shape[0]
    return np.ascontiguousarray(base[(np.arange(sz) - 1 + shift) % n])


def z_mirror_defects(u, v, w, p):
    """Test case 1 is symmetric under z -> -z: u, v, p are even, w (staggered in z, index k <-> sz-1-k) is odd.
    Returns the four relative asymmetries, each against the field's own maximum (w, u: against the velocity maximum,
    because they vanish identically at the start of the lid-driven run)."""
    vmax = max(float(np.max(np.abs(a))) for a in (u, v, w))
    out = []
    for a, sign in ((u, 1.0), (v, 1.0), (w, -1.0)):
        out.append(float(np.max(np.abs(a - sign * a[::-1]))) / max(vmax, 1e-300))
    out.append(float(np.max(np.abs(p - p[::-1]))) / max(float(np.max(np.abs(p))), 1e-300))
    return out
