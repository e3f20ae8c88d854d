"""The G4 orientation analysis has no reference implementation to compare with.  What pins it:
(1) the derivation -- lowest-order Fourier terms of the oriented energy as quadratic forms of the basis planes -- applied
to the G2/H2 steering weights reproduces EXACTLY the constants the reference hard-codes (SteerableFiltersG2.cpp:93-95);
(2) the generated G4 forms in taps_baked.inc equal the same derivation for the G4/H4 weights (G4.cpp:116-121).
CPU only.  tests/test_g4_gpu.py checks the kernel against a brute-force projection of the oracle's steered energy."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 64
TH = np.arange(N) * 2 * np.pi / N
C, S = np.cos(TH), np.sin(TH)


def forms(w):
    """w: [planes, N] steering weights -> (Q1, Q2, Q3) with C_k = b^T Q_k b."""
    q1 = np.einsum("it,jt->ij", w, w) / N
    q2 = 2 * np.einsum("it,jt,t->ij", w, w, np.cos(2 * TH)) / N
    q3 = 2 * np.einsum("it,jt,t->ij", w, w, np.sin(2 * TH)) / N
    return q1, q2, q3


G2W = np.stack([C * C, -2 * C * S, S * S])                                     # G2.cpp:141
H2W = np.stack([C ** 3, -3 * C * C * S, 3 * C * S * S, -S ** 3])               # G2.cpp:142
G4W = np.stack([C ** 4, -4 * C ** 3 * S, 6 * C * C * S * S, -4 * C * S ** 3, S ** 4])                      # G4.cpp:118
H4W = np.stack([C ** 5, -5 * C ** 4 * S, 10 * C ** 3 * S * S, -10 * C * C * S ** 3, 5 * C * S ** 4, -S ** 5])  # G4.cpp:119


def test_derivation_reproduces_reference_g2_constants():
    q1, q2, q3 = forms(G2W)
    a, b, c = 0, 1, 2
    # m_c1 = 0.5 bb + 0.25 ac + 0.375 (aa + cc) + ...      (G2.cpp:93)
    assert np.allclose([q1[b, b], 2 * q1[a, c], q1[a, a], q1[c, c]], [0.5, 0.25, 0.375, 0.375], atol=1e-14)
    # m_c2 = 0.5 (aa - cc) + ...                           (G2.cpp:94)
    assert np.allclose([q2[a, a], q2[c, c], q2[b, b], q2[a, c]], [0.5, -0.5, 0, 0], atol=1e-14)
    # m_c3 = -ab - bc - ...                                (G2.cpp:95)
    assert np.allclose([2 * q3[a, b], 2 * q3[b, c], q3[a, a], 2 * q3[a, c]], [-1, -1, 0, 0], atol=1e-14)
    q1, q2, q3 = forms(H2W)
    a, b, c, d = 0, 1, 2, 3
    assert np.allclose([q1[a, a], q1[d, d], q1[b, b], q1[c, c], 2 * q1[a, c], 2 * q1[b, d]],
                       [0.3125, 0.3125, 0.5625, 0.5625, 0.375, 0.375], atol=1e-14)
    assert np.allclose([q2[a, a], q2[d, d], q2[b, b], q2[c, c], 2 * q2[a, c], 2 * q2[b, d]],
                       [0.46875, -0.46875, 0.28125, -0.28125, 0.1875, -0.1875], atol=1e-14)
    assert np.allclose([2 * q3[c, d], 2 * q3[a, b], 2 * q3[b, c], 2 * q3[a, d]], [-0.9375, -0.9375, -1.6875, -0.1875], atol=1e-14)
    assert abs(q1[a, b]) < 1e-14 and abs(q3[a, a]) < 1e-14


def packed(q):
    n = q.shape[0]
    return [q[i, j] * (1 if i == j else 2) for i in range(n) for j in range(i, n)]


def test_generated_g4_forms_match_derivation():
    txt = open(os.path.join(ROOT, "cvsteer_b200", "csrc", "taps_baked.inc")).read()
    body = txt[txt.index("#define CVS_G4_ORIENT_FORMS"):]
    rows = re.findall(r"\{([^{}]+)\}", body)
    got = np.array([[float.fromhex(v.strip().rstrip("f")) for v in r.split(",")] for r in rows[:3]])
    assert got.shape == (3, 36)
    g, h = forms(G4W), forms(H4W)
    for k in range(3):
        want = np.array(packed(g[k]) + packed(h[k]))
        assert np.allclose(got[k], want, atol=1e-7), k
