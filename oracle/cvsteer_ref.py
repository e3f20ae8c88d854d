"""CPU oracle for the cvsteer hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The shipped path (cvsteer_b200/, include/) never does.

What it is
----------
A line-by-line restatement, in Python over ``cv2`` 4.13.0, of the reference's glue
code.  The reference (headupinclouds/cvsteer) delegates *all* arithmetic to OpenCV
(``sepFilter2D``, ``cartToPolar``, ``polarToCart``, ``patchNaNs``, Mat algebra); its
C++ cannot be compiled here (no OpenCV C++ headers, no gtest, Hunter needs network),
but Python cv2 exposes those same OpenCV primitives.  So this oracle drives the
reference's own third-party arithmetic library with a transliteration of the ~150
lines of reference glue:

    cvsteer/SteerableFilters.cpp:33-51        create(), wrap()
    cvsteer/SteerableFiltersG2.cpp:35-212     G2/H2 taps, setup, steer*, phase, find*
    cvsteer/SteerableFiltersG4.cpp:34-122     G4/H4 taps, setup, steer*

Parity pin: tests/test_oracle_golden.py reproduces the reference's only test,
TEST(cvsteer, basic) (test/test.cpp:70-103), on the reference's bundled fixture
images (decoded copies under tests/golden/), plus the frozen known-answer numbers of
SURVEY.md App. C.  Float-level quantities (basis planes, C1..C3, theta, phase) and all
of G4 are NOT pinned by any reference test; for those the pin is cv2 4.13.0 itself.

The pyramid (``pyr_down``) has no reference counterpart; it is defined as cv2.pyrDown.
"""
from __future__ import annotations

import math

import cv2
import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# Tap functions.  The reference writes e.g.
#     static float G21(float x) { return 0.9213 * (2.0 * x * x - 1.0) * exp(-x * x); }
# `x` is float, the literals are double, so the polynomial part is evaluated in double
# on a float-valued x; `exp(-x * x)` has a float argument and resolves to the float
# overload (std::exp(float) via <cmath>, SteerableFilters.h:32-33), whose float result
# is promoted to double for the final product; the return narrows to float.
# --------------------------------------------------------------------------------------
def _expf_neg_sq(x: np.float32) -> float:
    xx = F32(x) * F32(x)  # float * float
    # expf: glibc's expf is correctly rounded in practice; numpy's SIMD float32 exp is not always,
    # so round the double result instead of calling np.exp on a float32.
    return float(F32(math.exp(float(F32(-xx)))))


def _f(v: float) -> np.float32:
    return F32(v)


# cvsteer/SteerableFiltersG2.cpp:35-42
def G21(x): xd = float(x); return _f(0.9213 * (2.0 * xd * xd - 1.0) * _expf_neg_sq(x))
def G22(x): return _f(_expf_neg_sq(x))
def G23(x): xd = float(x); return _f(math.sqrt(1.8430) * xd * _expf_neg_sq(x))
def H21(x):  # `x * x * x` is a float product in C++; `-2.254 * x` is double
    xf = F32(x)
    return _f(0.9780 * (-2.254 * float(xf) + float(xf * xf * xf)) * _expf_neg_sq(x))
def H22(x): return _f(_expf_neg_sq(x))
def H23(x): return _f(float(F32(x) * F32(_expf_neg_sq(x))))  # float * float
def H24(x): xf = F32(x); return _f(0.9780 * (-0.7515 + float(xf * xf)) * _expf_neg_sq(x))


# cvsteer/SteerableFiltersG4.cpp:34-45.  Note `3.0f * x * x` is float arithmetic that is
# then promoted; x*x*x*x is float arithmetic.  At the 1e-8 relative level this does not
# matter for any tolerance used downstream; we follow the C++ typing anyway.
def G41(x):
    xf = F32(x)
    t3 = float(F32(3.0) * xf * xf)
    x4 = float(xf * xf * xf * xf)
    return _f(1.246 * (0.75 - t3 + x4) * _expf_neg_sq(x))
def G42(x): return _f(_expf_neg_sq(x))
def G43(x):
    xf = F32(x)
    return _f((-1.5 * float(xf) + float(xf * xf * xf)) * _expf_neg_sq(x))
def G44(x): return _f(1.246 * float(x) * _expf_neg_sq(x))
def G45(x):
    xf = F32(x)
    return _f(math.sqrt(1.246) * (float(xf * xf) - 0.5) * _expf_neg_sq(x))
def H41(x):
    xf = F32(x)
    xd = float(xf)
    x5 = float(xf * xf * xf * xf * xf)           # float product
    # `7.501 * x * x * x` associates as ((7.501*x)*x)*x: all double
    return _f(0.3975 * (7.189 * xd - 7.501 * xd * xd * xd + x5) * _expf_neg_sq(x))
def H42(x): return _f(_expf_neg_sq(x))
def H43(x):
    xf = F32(x)
    xd = float(xf)
    x4 = float(xf * xf * xf * xf)                # float product
    return _f(0.3975 * (1.438 - 4.501 * xd * xd + x4) * _expf_neg_sq(x))  # (4.501*x)*x in double
def H44(x): return _f(float(F32(x) * F32(_expf_neg_sq(x))))
def H45(x):
    xf = F32(x)
    return _f(0.3975 * (float(xf * xf * xf) - 2.225 * float(xf)) * _expf_neg_sq(x))
def H46(x):
    xf = F32(x)
    return _f((float(xf * xf) - 0.6638) * _expf_neg_sq(x))


def create(width: int, spacing: float, f) -> np.ndarray:
    """cvsteer/SteerableFilters.cpp:33-42 -- sample f at float(i)*spacing, i=-w..w."""
    sp = F32(spacing)
    k = np.empty((1, 2 * width + 1), F32)
    for i in range(-width, width + 1):
        k[0, i + width] = f(F32(F32(i) * sp))
    return k


def wrap(angle: np.ndarray) -> np.ndarray:
    """cvsteer/SteerableFilters.cpp:46-51 -- (pi, 2pi) -> (-pi, 0); strict '>'.

    `-M_PI - (M_PI - angle)` is a MatExpr with double scalars applied to a float Mat:
    OpenCV evaluates alpha*src + beta in one pass with double scalars and saturates to
    float, i.e. float(angle - 2*pi) computed in double.
    """
    out = angle.copy()
    tmp = (-math.pi - (math.pi - angle.astype(np.float64))).astype(F32)
    # `Mat1f > double`: cv::compare converts the scalar to the array depth, so the test is
    # angle > float(M_PI) in float [probe: cv2.compare(float32(pi), math.pi, CMP_GT) == 0].
    m = angle > F32(math.pi)
    np.copyto(out, tmp, where=m)  # wrapped.copyTo(output, mask)
    return out


def _sep(img: np.ndarray, kx: np.ndarray, ky: np.ndarray) -> np.ndarray:
    # cv::sepFilter2D(image, dst, CV_32FC1, kernelX, kernelY.t()): default anchor, delta 0,
    # BORDER_DEFAULT == BORDER_REFLECT_101.
    return cv2.sepFilter2D(img, cv2.CV_32F, kx.reshape(1, -1), ky.reshape(-1, 1))


def _polar_to_cart(theta: np.ndarray):
    # cv::polarToCart(cv::Mat(), theta, ct, st): magnitude omitted => cos, sin
    ct, st = cv2.polarToCart(None, np.ascontiguousarray(theta, F32))
    return ct, st


def compute_magnitude_and_phase(g: np.ndarray, h: np.ndarray):
    """cvsteer/SteerableFiltersG2.cpp:107-112."""
    mag, phase = cv2.cartToPolar(np.ascontiguousarray(g, F32), np.ascontiguousarray(h, F32))
    phase = wrap(phase)
    phase = cv2.patchNaNs(phase, 0.0)
    return mag, phase


def phase_weights(phase: np.ndarray, phi: float, signum: bool, k: float = 2.0) -> np.ndarray:
    """cvsteer/SteerableFiltersG2.cpp:179-186 (k is unused there too)."""
    phi = F32(phi)  # the parameter is `float phi`
    if signum:
        err = np.abs(phase - phi)
    else:
        err = np.abs(np.abs(phase) - F32(abs(phi)))
    err = err.astype(F32)
    # cv::min(error, 2.0*M_PI - error): the MatExpr (double scalar - float Mat) -> float Mat
    alt = (2.0 * math.pi - err.astype(np.float64)).astype(F32)
    err = np.minimum(err, alt)
    ct, _ = _polar_to_cart(err)
    lam = ct * ct
    lam[np.abs(err) > F32(math.pi / 2)] = 0  # scalar compared at array depth (float)
    return lam.astype(F32)


def _phase_edge(e, phase, phi, signum, k):
    lam = phase_weights(phase, phi, signum, k)
    return (e * lam).astype(F32)


def find_edges(e, phase, k=2.0):       # G2.cpp:201-204
    return _phase_edge(e, phase, math.pi / 2, False, k)
def find_dark_lines(e, phase, k=2.0):  # G2.cpp:205-208
    return _phase_edge(e, phase, 0.0, True, k)
def find_bright_lines(e, phase, k=2.0):  # G2.cpp:209-212
    return _phase_edge(e, phase, math.pi, True, k)


def _scale(alpha: float, m: np.ndarray) -> np.ndarray:
    """MatExpr `double * Mat1f` materialised as float: float(double(alpha)*m).  Every coefficient of G2.cpp:93-95 has at most
    5 significant bits, so double(alpha)*m is exact in double and rounds to the same float as the float product: the float32
    multiply below is bit-identical to the float64 round trip (tests/test_oracle_golden.py checks it) and spares the CPU
    baseline two conversions per term.  Any other coefficient takes the literal route."""
    a32 = F32(alpha)
    if float(a32) == float(alpha) and (np.frexp(float(alpha))[0] * 32.0).is_integer():
        return a32 * m
    return (alpha * m.astype(np.float64)).astype(F32)


class SteerableFiltersG2:
    """fa::SteerableFiltersG2 (cvsteer/SteerableFiltersG2.h:35-67, .cpp:44-212)."""

    PLANES = ("g2a", "g2b", "g2c", "h2a", "h2b", "h2c", "h2d")

    def __init__(self, image: np.ndarray, width: int = 4, spacing: float = 0.67):
        self.g1 = create(width, spacing, G21)
        self.g2 = create(width, spacing, G22)
        self.g3 = create(width, spacing, G23)
        self.h1 = create(width, spacing, H21)
        self.h2 = create(width, spacing, H22)
        self.h3 = create(width, spacing, H23)
        self.h4 = create(width, spacing, H24)
        self.setup(image)

    def setup(self, image: np.ndarray) -> None:
        # implicit Mat(8UC1) -> Mat1f conversion: convertTo without scaling
        img = np.ascontiguousarray(image, F32)
        self.g2a = _sep(img, self.g1, self.g2)
        self.g2b = _sep(img, self.g3, self.g3)
        self.g2c = _sep(img, self.g2, self.g1)
        self.h2a = _sep(img, self.h1, self.h2)
        self.h2b = _sep(img, self.h4, self.h3)
        self.h2c = _sep(img, self.h3, self.h4)
        self.h2d = _sep(img, self.h2, self.h1)
        a, b, c = self.g2a, self.g2b, self.g2c
        ha, hb, hc, hd = self.h2a, self.h2b, self.h2c, self.h2d
        g2aa, g2ab, g2ac, g2bb, g2bc, g2cc = a * a, a * b, a * c, b * b, b * c, c * c
        h2aa, h2ab, h2ac, h2ad = ha * ha, ha * hb, ha * hc, ha * hd
        h2bb, h2bc, h2bd = hb * hb, hb * hc, hb * hd
        h2cc, h2cd, h2dd = hc * hc, hc * hd, hd * hd
        # G2.cpp:93-95.  MatExpr evaluation: each `scalar * (Mat op Mat)` becomes a float
        # Mat; sums are float adds.  (Evaluation order differences are <=1 ulp per term.)
        s = _scale
        self.c1 = (s(0.5, g2bb) + s(0.25, g2ac) + s(0.375, g2aa + g2cc) + s(0.3125, h2aa + h2dd)
                   + s(0.5625, h2bb + h2cc) + s(0.375, h2ac + h2bd)).astype(F32)
        self.c2 = (s(0.5, g2aa - g2cc) + s(0.46875, h2aa - h2dd) + s(0.28125, h2bb - h2cc)
                   + s(0.1875, h2ac - h2bd)).astype(F32)
        self.c3 = ((-g2ab) - g2bc - s(0.9375, h2cd + h2ab) - s(1.6875, h2bc) - s(0.1875, h2ad)).astype(F32)
        strength, theta = cv2.cartToPolar(self.c2, self.c3)
        theta = wrap(theta)
        self.theta = (theta * F32(0.5)).astype(F32)
        self.strength = strength

    # getters (G2.h:40-41)
    def get_dominant_orientation_angle(self): return self.theta
    def get_dominant_orientation_strength(self): return self.strength

    # --- steering ---------------------------------------------------------------------
    def steer_point(self, p, theta: float, full: bool = False):
        """G2.cpp:115-134; p = (x, y) like cv::Point."""
        x, y = p
        th = F32(theta)
        ct = F32(math.cos(th)); st = F32(math.sin(th))  # std::cos(float) -> float
        ct2 = F32(ct * ct); ct3 = F32(ct2 * ct); st2 = F32(st * st); st3 = F32(st2 * st)
        ga, gb, gc = ct2, F32(-2.0 * float(ct) * float(st)), st2
        ha, hb, hc, hd = ct3, F32(-3.0 * float(ct2) * float(st)), F32(3.0 * float(ct) * float(st2)), F32(-st3)
        g2 = F32(ga * self.g2a[y, x] + gb * self.g2b[y, x] + gc * self.g2c[y, x])
        h2 = F32(ha * self.h2a[y, x] + hb * self.h2b[y, x] + hc * self.h2c[y, x] + hd * self.h2d[y, x])
        if not full:
            return g2, h2
        phase = F32(math.atan2(h2, g2))
        magnitude = F32(math.sqrt(float(h2 * h2 + g2 * g2)))
        c2t = F32(math.cos(float(th) * 2.0)); s2t = F32(math.sin(float(th) * 2.0))
        e = F32(self.c1[y, x] + c2t * self.c2[y, x] + s2t * self.c3[y, x])
        return g2, h2, e, magnitude, phase

    def steer_scalar(self, theta: float):
        """G2.cpp:137-145."""
        th = F32(theta)
        ct = F32(math.cos(th)); st = F32(math.sin(th))
        ct2 = F32(ct * ct); ct3 = F32(ct2 * ct); st2 = F32(st * st); st3 = F32(st2 * st)
        ga, gb, gc = ct2, F32(-2.0 * float(ct) * float(st)), st2
        ha, hb, hc, hd = ct3, F32(-3.0 * float(ct2) * float(st)), F32(3.0 * float(ct) * float(st2)), F32(-st3)
        g2 = (ga * self.g2a + gb * self.g2b + gc * self.g2c).astype(F32)
        h2 = (ha * self.h2a + hb * self.h2b + hc * self.h2c + hd * self.h2d).astype(F32)
        return g2, h2

    def steer_map(self, theta: np.ndarray):
        """G2.cpp:147-155."""
        ct, st = _polar_to_cart(theta)
        ct2 = ct * ct; ct3 = ct2 * ct; st2 = st * st; st3 = st2 * st
        g2 = (ct2 * self.g2a + _scale(-2.0, ct * st * self.g2b) + st2 * self.g2c).astype(F32)
        h2 = (ct3 * self.h2a + _scale(-3.0, ct2 * st * self.h2b) + _scale(3.0, ct * st2 * self.h2c)
              + (-st3 * self.h2d)).astype(F32)
        return g2, h2

    def steer_scalar_full(self, theta: float):
        """G2.cpp:157-165 -> (g2, h2, e, magnitude, phase)."""
        g2, h2 = self.steer_scalar(theta)
        mag, phase = compute_magnitude_and_phase(g2, h2)
        th = F32(theta)
        c2t = F32(math.cos(float(th) * 2.0)); s2t = F32(math.sin(float(th) * 2.0))
        e = (self.c1 + c2t * self.c2 + s2t * self.c3).astype(F32)
        return g2, h2, e, mag, phase

    def steer_map_full(self, theta: np.ndarray):
        """G2.cpp:167-177 -> (g2, h2, e, magnitude, phase)."""
        g2, h2 = self.steer_map(theta)
        mag, phase = compute_magnitude_and_phase(g2, h2)
        c2t, s2t = _polar_to_cart(_scale(2.0, theta))
        e = (self.c1 + self.c2 * c2t + self.c3 * s2t).astype(F32)
        return g2, h2, e, mag, phase

    computeMagnitudeAndPhase = staticmethod(compute_magnitude_and_phase)
    phaseWeights = staticmethod(phase_weights)
    findEdges = staticmethod(find_edges)
    findDarkLines = staticmethod(find_dark_lines)
    findBrightLines = staticmethod(find_bright_lines)


class SteerableFiltersG4:
    """fa::SteerableFiltersG4 (cvsteer/SteerableFiltersG4.h:35-56, .cpp:47-122)."""

    PLANES = ("g4a", "g4b", "g4c", "g4d", "g4e", "h4a", "h4b", "h4c", "h4d", "h4e", "h4f")

    def __init__(self, image: np.ndarray, width: int = 6, spacing: float = 0.5):
        self.g1 = create(width, spacing, G41)
        self.g2 = create(width, spacing, G42)
        self.g3 = create(width, spacing, G43)
        self.g4 = create(width, spacing, G44)
        self.g5 = create(width, spacing, G45)
        self.h1 = create(width, spacing, H41)
        self.h2 = create(width, spacing, H42)
        self.h3 = create(width, spacing, H43)
        self.h4 = create(width, spacing, H44)
        self.h5 = create(width, spacing, H45)
        self.h6 = create(width, spacing, H46)
        self.setup(image)

    def setup(self, image: np.ndarray) -> None:
        img = np.ascontiguousarray(image, F32)
        self.g4a = _sep(img, self.g1, self.g2)
        self.g4b = _sep(img, self.g3, self.g4)
        self.g4c = _sep(img, self.g5, self.g5)
        self.g4d = _sep(img, self.g4, self.g3)
        self.g4e = _sep(img, self.g2, self.g1)
        self.h4a = _sep(img, self.h1, self.h2)
        self.h4b = _sep(img, self.h3, self.h4)
        self.h4c = _sep(img, self.h5, self.h6)
        self.h4d = _sep(img, self.h6, self.h5)
        self.h4e = _sep(img, self.h4, self.h3)
        self.h4f = _sep(img, self.h2, self.h1)

    def steer_map(self, theta: np.ndarray):
        """G4.cpp:92-112."""
        ct, st = _polar_to_cart(theta)
        ct2 = ct * ct; ct3 = ct2 * ct; ct4 = ct3 * ct; ct5 = ct4 * ct
        st2 = st * st; st3 = st2 * st; st4 = st3 * st; st5 = st4 * st
        s = _scale
        ga, gb, gc, gd, ge = ct4, s(-4.0, ct3 * st), s(6.0, ct2 * st2), s(-4.0, ct * st3), st4
        ha, hb, hc, hd, he, hf = (ct5, s(-5.0, ct4 * st), s(10.0, ct3 * st2), s(-10.0, ct2 * st3),
                                  s(5.0, ct * st4), -st5)
        g4 = (ga * self.g4a + gb * self.g4b + gc * self.g4c + gd * self.g4d + ge * self.g4e).astype(F32)
        h4 = (ha * self.h4a + hb * self.h4b + hc * self.h4c + hd * self.h4d + he * self.h4e
              + hf * self.h4f).astype(F32)
        return g4, h4

    def steer_scalar(self, theta: float):
        """G4.cpp:114-122."""
        th = F32(theta)
        ct = F32(math.cos(th)); st = F32(math.sin(th))
        ct2 = F32(ct * ct); ct3 = F32(ct2 * ct); ct4 = F32(ct3 * ct); ct5 = F32(ct4 * ct)
        st2 = F32(st * st); st3 = F32(st2 * st); st4 = F32(st3 * st); st5 = F32(st4 * st)
        d = float
        ga, gb, gc, gd, ge = (ct4, F32(-4.0 * d(ct3) * d(st)), F32(6.0 * d(ct2) * d(st2)),
                              F32(-4.0 * d(ct) * d(st3)), st4)
        ha, hb, hc, hd, he, hf = (ct5, F32(F32(-5.0) * ct4 * st), F32(10.0 * d(ct3) * d(st2)),
                                  F32(-10.0 * d(ct2) * d(st3)), F32(5.0 * d(ct) * d(st4)), F32(-st5))
        g4 = (ga * self.g4a + gb * self.g4b + gc * self.g4c + gd * self.g4d + ge * self.g4e).astype(F32)
        h4 = (ha * self.h4a + hb * self.h4b + hc * self.h4c + hd * self.h4d + he * self.h4e
              + hf * self.h4f).astype(F32)
        return g4, h4

    def steer_map_full(self, theta: np.ndarray):
        """Config 4 ("G4/H4 steering + phase").  The reference's G4
        computeMagnitudeAndPhase is an empty body (G4.cpp:88-90); the definition used here
        is the G2 class's (G2.cpp:107-112) applied to (g4, h4) -- see SURVEY.md section 8 a7."""
        g4, h4 = self.steer_map(theta)
        mag, phase = compute_magnitude_and_phase(g4, h4)
        return g4, h4, mag, phase


def pyr_down(img: np.ndarray) -> np.ndarray:
    """Pyramid level l -> l+1.  No reference counterpart; DEFINED as cv2.pyrDown:
    separable [1 4 6 4 1]/16, BORDER_REFLECT_101, even samples, size ((W+1)//2,(H+1)//2)."""
    return cv2.pyrDown(np.ascontiguousarray(img, F32))


def pyramid(img: np.ndarray, levels: int):
    out = [np.ascontiguousarray(img, F32)]
    for _ in range(levels - 1):
        out.append(pyr_down(out[-1]))
    return out


def normalize_minmax_u8(x: np.ndarray) -> np.ndarray:
    """cv::normalize(x, dst, 0, 255, cv::NORM_MINMAX, CV_8UC1) (test/test.cpp:93-95)."""
    return cv2.normalize(x, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8UC1)


def g2_orientation(image: np.ndarray, width: int = 4, spacing: float = 0.67):
    """Mode M1 of SURVEY.md section 8d: (theta_d, strength, energy at theta_d).
    Energy at theta_d via the reference's own formula (G2.cpp:174-176)."""
    f = SteerableFiltersG2(image, width, spacing)
    c2t, s2t = _polar_to_cart(_scale(2.0, f.theta))
    e = (f.c1 + f.c2 * c2t + f.c3 * s2t).astype(F32)
    return f.theta, f.strength, e


def g2_full(image: np.ndarray, width: int = 4, spacing: float = 0.67):
    """What both reference callers run (example/steer.cpp:86-87, test/test.cpp:85-86):
    construct, then steer(getDominantOrientationAngle(), g2,h2,e,magnitude,phase)."""
    f = SteerableFiltersG2(image, width, spacing)
    g2, h2, e, mag, phase = f.steer_map_full(f.theta)
    return f, (g2, h2, e, mag, phase)
