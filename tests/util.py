"""Shared helpers for the parity tests.

Tolerances (BASELINE.json north_star): max abs error <= 1e-4 of the basis range for filter outputs, angle error
<= 1e-3 rad where strength exceeds a threshold.  Angles are compared on the circle (period pi for the dominant
orientation, 2 pi for phase)."""
import numpy as np

BASIS_RTOL = 1e-4      # of the basis range
ANGLE_TOL = 1e-3       # rad


def synth(seed, rows, cols):
    """SURVEY section 8d synthetic input: fp32 uniform[0,255) from default_rng(seed)."""
    return np.random.default_rng(seed).uniform(0, 255, (rows, cols)).astype(np.float32)


def basis_range(planes):
    return float(max(p.max() for p in planes) - min(p.min() for p in planes))


def assert_close_range(got, want, rng, name="", rtol=BASIS_RTOL):
    assert got.shape == want.shape, (name, got.shape, want.shape)
    err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) if got.size else 0.0
    assert err <= rtol * rng, f"{name}: max abs err {err:.3e} > {rtol:g} x range {rng:.3e}"
    return err


def circ_diff(a, b, period):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64)) % period
    return np.minimum(d, period - d)


def assert_angle_close(got, want, weight, period, name="", tol=ANGLE_TOL, thresh_frac=1e-3):
    """Compare angles where `weight` (strength / magnitude) exceeds thresh_frac of its max."""
    m = weight > thresh_frac * float(weight.max()) if weight.size and weight.max() > 0 else np.zeros_like(weight, bool)
    if not m.any():
        return 0.0
    err = float(circ_diff(got[m], want[m], period).max())
    assert err <= tol, f"{name}: max angle err {err:.3e} rad > {tol:g} over {int(m.sum())} px"
    return err
