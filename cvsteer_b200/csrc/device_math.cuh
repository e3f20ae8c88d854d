// Point-wise math shared by every kernel: the orientation analysis, steering weights, magnitude /
// phase and phase-weight maps of the reference, with OpenCV-compatible atan2.
//
// Reference lines restated here (cvsteer/SteerableFiltersG2.cpp unless noted):
//   products + C1,C2,C3 ............ :70-95
//   dominant angle + strength ...... :97-99  (+ SteerableFilters::wrap, SteerableFilters.cpp:46-51)
//   steering weights ............... :140-144 (G2), SteerableFiltersG4.cpp:116-121 (G4)
//   magnitude / phase .............. :107-112
//   oriented energy ................ :163-164, :174-176
//   phase weights / find* .......... :179-212
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace cvs {
namespace dev {

// BORDER_REFLECT_101 index folding, iterated so that images smaller than the filter radius behave like
// cv::borderInterpolate (a size-1 dimension maps every index to 0).
__host__ __device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}

// Single-instruction SFU approximations (MUFU.RCP / MUFU.SQRT), ~1 ulp; used by the fused kernels where the
// IEEE sequences (8-10 instructions each) would cost more issue slots than the whole steering step.
__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// cv::pyrDown's 5-tap [1 4 6 4 1] in OpenCV's float order (6*c + 4*(l1+r1) + l2 + r2), spelled with explicit FMAs so
// that every kernel that builds a pyramid level (stand-alone or fused into the basis kernel) rounds identically.
__device__ __forceinline__ float pyr_tap5(float v0, float v1, float v2, float v3, float v4)
{
    return fmaf(v1 + v3, 4.f, v2 * 6.f) + v0 + v4;
}

// sin/cos for a steering angle.  |x| <= pi/2 (every dominant-orientation angle) takes two short minimax polynomials in
// x^2 (abs error 1.4e-7 in fp32, the same order as cv::polarToCart's own error); anything else takes libdevice's
// accurate sincosf.  The branch is warp-uniform in practice (angle maps are either all theta_d or all arbitrary).
__device__ __forceinline__ void sincos_steer(float x, float* s, float* c)
{
    if (fabsf(x) <= 1.5709f) {
        const float t = x * x;
        float ps = fmaf(t, 2.605136842248612e-06f, -0.00019809033256024122f);
        ps = fmaf(ps, t, 0.008333050645887852f);
        ps = fmaf(ps, t, -0.16666658222675323f);
        ps = fmaf(ps, t, 1.0f);
        float pc = fmaf(t, -2.6050781798403477e-07f, 2.4760118321864866e-05f);
        pc = fmaf(pc, t, -0.0013888360699638724f);
        pc = fmaf(pc, t, 0.04166663438081741f);
        pc = fmaf(pc, t, -0.5f);
        pc = fmaf(pc, t, 1.0f);
        *s = ps * x;
        *c = pc;
    } else {
        sincosf(x, s, c);
    }
}

// cv::cartToPolar's angle (hal::fastAtan32f, in radians, range [0, 2pi)): 7th-order odd polynomial in
// min/max, evaluated with FMAs as OpenCV's SIMD path does.  With FAST = false (IEEE division) this is
// bit-identical to cv2 4.13.0 on 2M random points; FAST = true replaces the division by MUFU.RCP (<= 2 ulp on
// the ratio, i.e. ~1e-7 rad).
// NANS = false skips forcing NaN inputs through (fmaxf/fminf drop NaNs); callers that need cv::patchNaNs semantics
// test the inputs themselves (one unordered compare).
template <bool FAST = false, bool NANS = true>
__device__ __forceinline__ float cv_atan2(float y, float x)
{
    constexpr float kDeg = 57.29577951308232f;  // (float)(180/CV_PI)
    constexpr float P1 = 0.9997878412794807f * kDeg;
    constexpr float P3 = -0.3258083974640975f * kDeg;
    constexpr float P5 = 0.1555786518463281f * kDeg;
    constexpr float P7 = -0.04432655554792128f * kDeg;
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float den = mx + 2.220446049250313e-16f;  // + (float)DBL_EPSILON: (0,0) -> 0
    const float c = FAST ? mn * fast_rcp(den) : __fdiv_rn(mn, den);
    const float c2 = c * c;
    float a = fmaf(fmaf(fmaf(P7, c2, P5), c2, P3), c2, P1) * c;
    a = (ay > ax) ? 90.f - a : a;
    a = (x < 0.f) ? 180.f - a : a;
    a = (y < 0.f) ? 360.f - a : a;
    // NaN inputs: fmaxf/fminf drop NaNs, so force the NaN through as cv does (c = NaN there)
    if (NANS) a = (x != x || y != y) ? __int_as_float(0x7fc00000) : a;
    return a * 0.017453292519943295f;  // (float)(CV_PI/180)
}

// Fast-path variant for the fused kernels: the same polynomial and quadrant logic as cv_atan2, but with the degree ->
// radian factor, SteerableFilters::wrap (angles above pi come down by 2 pi) and an optional final scale folded into the
// constants: returns SCALE * wrap(atan2(y, x)) with SCALE = 1 (phase) or 0.5 (dominant orientation).  Three instructions
// shorter; differs from the exact sequence by rounding only (~1e-7 rad).
template <bool HALF>
__device__ __forceinline__ float cv_atan2_wrapped_fast(float y, float x)
{
    constexpr float S = HALF ? 0.5f : 1.0f;
    constexpr float P1 = 0.9997878412794807f * S, P3 = -0.3258083974640975f * S, P5 = 0.1555786518463281f * S,
                    P7 = -0.04432655554792128f * S;
    constexpr float kHalfPi = 1.5707963267948966f * S, kPi = 3.14159265358979f * S, kTwoPi = 6.283185307179586f * S;
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float c = mn * fast_rcp(mx + 2.220446049250313e-16f);
    const float c2 = c * c;
    float a = fmaf(fmaf(fmaf(P7, c2, P5), c2, P3), c2, P1) * c;
    a = (ay > ax) ? kHalfPi - a : a;
    a = (x < 0.f) ? kPi - a : a;
    // y < 0: the unwrapped angle is 2 pi - a, which is above pi and wraps to -a  (y == -0 stays at +a, like 360 - 0 = 360 -> 0)
    a = (y < 0.f) ? ((a == 0.f) ? 0.f : -a) : a;
    (void)kTwoPi;
    return a;
}

// cv::cartToPolar's magnitude: sqrt(x*x + y*y) with the inner sum fused; IEEE sqrt, or MUFU.SQRT when FAST.
template <bool FAST = false>
__device__ __forceinline__ float cv_magnitude(float x, float y)
{
    const float q = fmaf(x, x, y * y);
    return FAST ? fast_sqrt(q) : __fsqrt_rn(q);
}

// SteerableFilters::wrap: a > float(pi) -> float(double(a) - 2pi).  float(2pi) = 2pi + 1.7484555e-7, the
// first subtraction is exact (Sterbenz), the correction restores the double-precision result.
__device__ __forceinline__ float wrap_pi(float a)
{
    return (a > 3.14159274101257324f) ? (a - 6.28318548202514648f) + 1.7484555e-7f : a;
}

struct Orientation {
    float c1, c2, c3, strength, theta;
};

// G2.cpp:70-99 on the 7 basis values of one pixel.
template <bool FAST = false>
__device__ __forceinline__ Orientation orientation_g2(float a, float b, float c, float ha, float hb, float hc,
                                                     float hd)
{
    Orientation o;
    const float aa = a * a, cc = c * c, haa = ha * ha, hdd = hd * hd, hbb = hb * hb, hcc = hc * hc;
    const float hac = ha * hc, hbd = hb * hd;
    float c1 = 0.5f * (b * b);
    c1 = fmaf(0.25f, a * c, c1);
    c1 = fmaf(0.375f, aa + cc, c1);
    c1 = fmaf(0.3125f, haa + hdd, c1);
    c1 = fmaf(0.5625f, hbb + hcc, c1);
    c1 = fmaf(0.375f, hac + hbd, c1);
    float c2 = 0.5f * (aa - cc);
    c2 = fmaf(0.46875f, haa - hdd, c2);
    c2 = fmaf(0.28125f, hbb - hcc, c2);
    c2 = fmaf(0.1875f, hac - hbd, c2);
    float c3 = -(a * b) - b * c;
    c3 = fmaf(-0.9375f, fmaf(hc, hd, ha * hb), c3);
    c3 = fmaf(-1.6875f, hb * hc, c3);
    c3 = fmaf(-0.1875f, ha * hd, c3);
    o.c1 = c1;
    o.c2 = c2;
    o.c3 = c3;
    o.strength = cv_magnitude<FAST>(c2, c3);
    if (FAST) o.theta = cv_atan2_wrapped_fast<true>(c3, c2);
    else o.theta = 0.5f * wrap_pi(cv_atan2<false, true>(c3, c2));
    return o;
}

// G2.cpp:140-144 / :151-154
__device__ __forceinline__ void steer_g2(float ct, float st, float a, float b, float c, float ha, float hb,
                                         float hc, float hd, float& g2, float& h2)
{
    const float ct2 = ct * ct, st2 = st * st, cs = ct * st;
    g2 = fmaf(ct2, a, fmaf(-2.f * cs, b, st2 * c));
    h2 = fmaf(ct2 * ct, ha, fmaf(-3.f * ct2 * st, hb, fmaf(3.f * ct * st2, hc, -(st2 * st) * hd)));
}

// G4.cpp:97-111 / :116-121
__device__ __forceinline__ void steer_g4(float ct, float st, const float* g /*5*/, const float* h /*6*/, float& g4,
                                         float& h4)
{
    const float ct2 = ct * ct, ct3 = ct2 * ct, ct4 = ct3 * ct, ct5 = ct4 * ct;
    const float st2 = st * st, st3 = st2 * st, st4 = st3 * st, st5 = st4 * st;
    g4 = fmaf(ct4, g[0], fmaf(-4.f * ct3 * st, g[1], fmaf(6.f * ct2 * st2, g[2], fmaf(-4.f * ct * st3, g[3], st4 * g[4]))));
    h4 = fmaf(ct5, h[0],
              fmaf(-5.f * ct4 * st, h[1],
                   fmaf(10.f * ct3 * st2, h[2], fmaf(-10.f * ct2 * st3, h[3], fmaf(5.f * ct * st4, h[4], -st5 * h[5])))));
}

// The same sums when the binomial factors (1 -4 6 -4 1 / 1 -5 10 -10 5 -1) are already folded into the basis values
// (the steer-only kernels scale the column taps at compile time): 8 + 17 instructions instead of 33.
__device__ __forceinline__ void steer_g4_prescaled(float ct, float st, const float* g /*5*/, const float* h /*6*/, float& g4, float& h4)
{
    const float c2 = ct * ct, s2 = st * st, cs = ct * st;
    const float c4 = c2 * c2, c3s = c2 * cs, c2s2 = cs * cs, cs3 = cs * s2, s4 = s2 * s2;
    g4 = fmaf(c4, g[0], fmaf(c3s, g[1], fmaf(c2s2, g[2], fmaf(cs3, g[3], s4 * g[4]))));
    h4 = fmaf(c4 * ct, h[0], fmaf(c4 * st, h[1], fmaf(c3s * st, h[2], fmaf(c2s2 * st, h[3], fmaf(cs3 * st, h[4], (s4 * st) * h[5])))));
}

// G2.cpp:107-112
template <bool FAST = false>
__device__ __forceinline__ void magnitude_phase(float g, float h, float& mag, float& phase)
{
    mag = cv_magnitude<FAST>(g, h);
    const float p = FAST ? cv_atan2_wrapped_fast<false>(h, g) : wrap_pi(cv_atan2<false, false>(h, g));
    // cv::patchNaNs: the reference's angle is NaN iff an input is NaN.  One unordered compare of the two inputs
    // (setp.nan is true when either operand is NaN) selects 0.
    asm("{\n\t.reg .pred q;\n\tsetp.nan.f32 q, %1, %2;\n\tselp.f32 %0, 0f00000000, %3, q;\n\t}" : "=f"(phase) : "f"(g), "f"(h), "f"(p));
}

// G2.cpp:179-186: lambda = cos^2(err) gated at pi/2.
template <bool FAST = false>
__device__ __forceinline__ float phase_weight(float phase, float phi, bool signum)
{
    float err = signum ? fabsf(phase - phi) : fabsf(fabsf(phase) - fabsf(phi));
    const float alt = (6.28318548202514648f - err) - 1.7484555e-7f;  // float(2*M_PI - err)
    err = fminf(err, alt);
    const float ct = FAST ? __cosf(err) : cosf(err);  // err in [0, pi]
    return (fabsf(err) > 1.57079637050628662f) ? 0.f : ct * ct;
}

}  // namespace dev
}  // namespace cvs
