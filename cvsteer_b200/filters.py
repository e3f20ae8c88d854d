"""Host-array mirror of the reference's class surface, over the C ABI.

Same names, argument meaning and error behaviour as fa::SteerableFiltersG2 / fa::SteerableFiltersG4
(reference cvsteer/SteerableFiltersG2.h:35-67, SteerableFiltersG4.h:35-56), with numpy arrays standing in for
cv::Mat1f, so that the parity tests read like the reference's own test (test/test.cpp:85-90).  All pixel
arithmetic happens in libcvsteer_b200.so on the GPU; numpy is used for buffers only.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi


def _f32c(a, name="image"):
    a = np.asarray(a)
    if a.ndim != 2 or a.size == 0:
        # the reference lets OpenCV throw on empty / non-2D input; we raise the ABI's INVALID_ARG
        raise capi.CvsError(capi.ERR_INVALID_ARG, f"{name} must be a non-empty 2-D array")
    return np.ascontiguousarray(a, np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def make_taps_g2(which: int, width: int = 4, spacing: float = 0.67) -> np.ndarray:
    """SteerableFilters::create with G21..H24 (which: 0..6 = g1,g2,g3,h1,h2,h3,h4)."""
    out = np.empty(2 * width + 1, np.float32)
    capi.check(capi.lib().cvs_g2_make_taps(which, width, spacing, out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


def make_taps_g4(which: int, width: int = 6, spacing: float = 0.5) -> np.ndarray:
    """SteerableFilters::create with G41..H46 (which: 0..10 = g1..g5,h1..h6)."""
    out = np.empty(2 * width + 1, np.float32)
    capi.check(capi.lib().cvs_g4_make_taps(which, width, spacing, out.ctypes.data_as(C.POINTER(C.c_float))))
    return out


class _Base:
    _prefix = ""
    _nstate = 0

    def __init__(self, image, width, spacing, device=0):
        self._lib = capi.lib()
        self._h = C.c_void_p()
        capi.check(getattr(self._lib, f"cvs_{self._prefix}_create")(C.byref(self._h), device, width, spacing))
        self.width, self.spacing = width, spacing
        self.rows = self.cols = 0
        self._cache = {}
        self.setup(image)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            getattr(self._lib, f"cvs_{self._prefix}_destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setup(self, image):
        """setup(const cv::Mat1f&): G2.cpp:60 / G4.cpp:67.  uint8 input takes the 8-bit upload path (the implicit
        Mat(8UC1) -> Mat1f conversion both reference callers rely on)."""
        a = np.asarray(image)
        self._cache = {}
        if a.dtype == np.uint8 and self._prefix == "g2" and a.ndim == 2 and a.size:
            a = np.ascontiguousarray(a)
            capi.check(self._lib.cvs_g2_setup_host_u8(self._h, _ptr(a), a.shape[0], a.shape[1], a.strides[0]))
        else:
            a = _f32c(a)
            capi.check(getattr(self._lib, f"cvs_{self._prefix}_setup_host")(self._h, _ptr(a), a.shape[0], a.shape[1],
                                                                            a.strides[0]))
        self.rows, self.cols = a.shape

    def plane(self, idx: int) -> np.ndarray:
        """Lazily downloaded host mirror of one protected member (m_g2a ... m_orientationStrength)."""
        if idx not in self._cache:
            out = np.empty((self.rows, self.cols), np.float32)
            capi.check(getattr(self._lib, f"cvs_{self._prefix}_get_plane_host")(self._h, idx, _ptr(out), out.strides[0]))
            self._cache[idx] = out
        return self._cache[idx]

    def _new(self, n):
        return [np.empty((self.rows, self.cols), np.float32) for _ in range(n)]


class SteerableFiltersG2(_Base):
    """fa::SteerableFiltersG2."""
    _prefix = "g2"

    def __init__(self, image, width: int = 4, spacing: float = 0.67, device: int = 0):
        super().__init__(image, width, spacing, device)

    # getters, G2.h:40-41
    def getDominantOrientationAngle(self):
        return self.plane(capi.THETA)

    def getDominantOrientationStrength(self):
        return self.plane(capi.STRENGTH)

    def __getattr__(self, name):  # m_g2a .. m_c3 style access for tests: f.g2a, f.c1, f.theta ...
        if name in capi.G2_PLANE_NAMES[:12]:
            return self.plane(capi.G2_PLANE_NAMES.index(name))
        raise AttributeError(name)

    def steer(self, theta, full: bool = True, point=None):
        """All steer() overloads of G2.h:44-53.
        theta float -> scalar angle; 2-D array -> per-pixel map; None -> the dominant-orientation map kept on
        the device (steer(getDominantOrientationAngle(), ...) without a round trip).
        point=(x, y) selects the single-pixel overloads.  full=False returns (g2, h2) only."""
        if point is not None:
            out = (C.c_float * 5)()
            capi.check(self._lib.cvs_g2_steer_point(self._h, int(point[0]), int(point[1]), float(theta), out))
            v = tuple(np.float32(x) for x in out)
            return v if full else v[:2]
        outs = self._new(5 if full else 2)
        ptrs = [_ptr(o) for o in outs] + [None] * (5 - len(outs))
        step = outs[0].strides[0]
        if theta is None or isinstance(theta, np.ndarray):
            if theta is None:
                tp, ts = None, 0
            else:
                th = _f32c(theta, "theta")
                if th.shape != (self.rows, self.cols):
                    raise capi.CvsError(capi.ERR_SIZE_MISMATCH, f"theta {th.shape} vs image {(self.rows, self.cols)}")
                tp, ts = _ptr(th), th.strides[0]
            capi.check(self._lib.cvs_g2_steer_map_host(self._h, tp, ts, *ptrs, step))
        else:
            capi.check(self._lib.cvs_g2_steer_scalar_host(self._h, float(theta), *ptrs, step))
        return tuple(outs)

    def computeMagnitudeAndPhase(self, g2, h2):
        g2, h2 = _f32c(g2, "g2"), _f32c(h2, "h2")
        mag, ph = np.empty_like(g2), np.empty_like(g2)
        capi.check(self._lib.cvs_magnitude_phase_host(0, _ptr(g2), _ptr(h2), g2.strides[0], _ptr(mag), _ptr(ph),
                                                      mag.strides[0], g2.shape[0], g2.shape[1]))
        return mag, ph

    @staticmethod
    def phaseWeights(phase, phi: float, signum: bool, k: float = 2.0):
        phase = _f32c(phase, "phase")
        lam = np.empty_like(phase)
        capi.check(capi.lib().cvs_phase_weights_host(0, _ptr(phase), phase.strides[0], _ptr(lam), lam.strides[0],
                                                     phase.shape[0], phase.shape[1], phi, int(bool(signum)), k))
        return lam

    @staticmethod
    def _find(kind, e, phase, k):
        e, phase = _f32c(e, "e"), _f32c(phase, "phase")
        out = np.empty_like(e)
        capi.check(capi.lib().cvs_find_host(0, kind, _ptr(e), _ptr(phase), e.strides[0], _ptr(out), out.strides[0],
                                            e.shape[0], e.shape[1], k))
        return out

    def findEdges(self, e, phase, k: float = 2.0):
        return self._find(0, e, phase, k)

    def findDarkLines(self, e, phase, k: float = 2.0):
        return self._find(1, e, phase, k)

    def findBrightLines(self, e, phase, k: float = 2.0):
        return self._find(2, e, phase, k)


class SteerableFiltersG4(_Base):
    """fa::SteerableFiltersG4."""
    _prefix = "g4"

    def __init__(self, image, width: int = 6, spacing: float = 0.5, device: int = 0):
        super().__init__(image, width, spacing, device)

    def __getattr__(self, name):
        if name in capi.G4_PLANE_NAMES[:11]:
            return self.plane(capi.G4_PLANE_NAMES.index(name))
        raise AttributeError(name)

    # Extension (the reference declares these getters but never assigns the members, G4.h:40-41,55): lowest-order
    # Fourier terms of G4(theta)^2 + H4(theta)^2, the definition the reference uses for G2 -- see CVS_G4_THETA.
    def getDominantOrientationAngle(self):
        return self.plane(capi.G4_THETA)

    def getDominantOrientationStrength(self):
        return self.plane(capi.G4_STRENGTH)

    def steer(self, theta, with_phase: bool = False):
        """steer(float | Mat1f theta, g4, h4), G4.h:45-46; with_phase adds (magnitude, phase) per the G2 definition.
        theta=None steers at the handle's own dominant-orientation map (extension)."""
        outs = self._new(4 if with_phase else 2)
        ptrs = [_ptr(o) for o in outs] + [None] * (4 - len(outs))
        step = outs[0].strides[0]
        if theta is None:
            capi.check(self._lib.cvs_g4_steer_map_host(self._h, None, 0, *ptrs, step))
        elif isinstance(theta, np.ndarray):
            th = _f32c(theta, "theta")
            if th.shape != (self.rows, self.cols):
                raise capi.CvsError(capi.ERR_SIZE_MISMATCH, f"theta {th.shape} vs image {(self.rows, self.cols)}")
            capi.check(self._lib.cvs_g4_steer_map_host(self._h, _ptr(th), th.strides[0], *ptrs, step))
        else:
            capi.check(self._lib.cvs_g4_steer_scalar_host(self._h, float(theta), *ptrs, step))
        return tuple(outs)
