"""Shared helpers for the parity tests.

Tolerances (BASELINE.json north_star): max abs error <= 1e-4 of the basis range for filter outputs, angle error
<= 1e-3 rad where strength exceeds a threshold.  Angles are compared on the circle (period pi for the dominant
orientation, 2 pi for phase)."""
import numpy as np

BASIS_RTOL = 1e-4      # of the basis range
ANGLE_TOL = 1e-3       # rad

# Every comparison made through the helpers below is recorded here; tests/conftest.py writes the per-quantity maxima to
# gpurun_out/parity_achieved.json at the end of a GPU session (copied to profiles/parity.json by hand).
RECORDS = []


def _record(kind, name, err, tol, scale, n):
    RECORDS.append({"kind": kind, "name": name, "err": float(err), "tol": float(tol), "scale": float(scale), "n": int(n)})


def synth(seed, rows, cols):
    """SURVEY section 8d synthetic input: fp32 uniform[0,255) from default_rng(seed)."""
    return np.random.default_rng(seed).uniform(0, 255, (rows, cols)).astype(np.float32)


def basis_range(planes):
    return float(max(p.max() for p in planes) - min(p.min() for p in planes))


def own_range(want, basis_rng):
    """Range of a QUADRATIC quantity (c1..c3, strength, energy) for its own tolerance: max - min of the oracle's plane.
    Degenerate planes (constant images: strength == 0 everywhere) fall back to rtol x basis_range^2 so that rounding
    residue of the filters (taps do not sum to exactly zero) is not compared at zero tolerance."""
    w = np.asarray(want, dtype=np.float64)
    ptp = float(w.max() - w.min()) if w.size else 0.0
    return max(ptp, BASIS_RTOL * basis_rng * basis_rng)


def assert_close_range(got, want, rng, name="", rtol=BASIS_RTOL):
    """max abs error <= rtol x rng.  rng = ("own", basis_range): the quantity's OWN range (see own_range)."""
    assert got.shape == want.shape, (name, got.shape, want.shape)
    if isinstance(rng, tuple):
        rng = own_range(want, rng[1])
    err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) if got.size else 0.0
    _record("range", name, err, rtol * rng, rng, got.size)
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
    _record("angle", name, err, tol, period, int(m.sum()))
    assert err <= tol, f"{name}: max angle err {err:.3e} rad > {tol:g} over {int(m.sum())} px"
    return err
