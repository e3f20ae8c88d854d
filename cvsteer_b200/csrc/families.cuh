// Filter families for the marching kernel: which unique 1-D tap sets exist, which (row set, column set)
// pair makes each basis plane, and the fused point-wise epilogue.
//
// G2/H2 pairs: reference cvsteer/SteerableFiltersG2.cpp:62-68; G4/H4 pairs: SteerableFiltersG4.cpp:69-80.
// cv::sepFilter2D(image, dst, CV_32F, kernelX, kernelY): kernelX runs along x (the row pass), kernelY along y
// (the column pass); correlation, no flip.  g2 == h2 == exp(-x^2) in both families, so one row pass is shared.
#pragma once
#include "../../include/cvsteer_c.h"
#include "march.cuh"

namespace cvs {

struct G2Fam {
    static constexpr int R = 4, NSETS = 6, NROW = 6, NBASIS = 7, BH = 64, MIN_CTAS = 4;
    // unique tap sets: 0 g1, 1 g2(=h2), 2 g3, 3 h1, 4 h3, 5 h4   (index into TapTable::t)
    // map from the API's 7 tap sets (g1,g2,g3,h1,h2,h3,h4) to unique sets
    __host__ __device__ static constexpr int unique_of(int api) { constexpr int t[7] = {0, 1, 2, 3, 1, 4, 5}; return t[api]; }
    __host__ __device__ static constexpr bool set_odd(int s) { return s == 2 || s == 3 || s == 4; }
    // row-filtered planes: one per unique set
    __host__ __device__ static constexpr int row_set(int p) { return p; }
    __host__ __device__ static constexpr bool row_odd(int p) { return set_odd(p); }
    // basis planes (CVS_G2A..CVS_H2D):  row(kernelX) set, column(kernelY) set
    //   G2a (g1,g2)  G2b (g3,g3)  G2c (g2,g1)  H2a (h1,h2)  H2b (h4,h3)  H2c (h3,h4)  H2d (h2,h1)
    __host__ __device__ static constexpr int basis_row(int q) { constexpr int t[7] = {0, 2, 1, 3, 5, 4, 1}; return t[q]; }
    __host__ __device__ static constexpr int basis_set(int q) { constexpr int t[7] = {1, 2, 0, 1, 4, 5, 3}; return t[q]; }
    __host__ __device__ static constexpr bool basis_odd(int q) { return set_odd(basis_set(q)); }

    static constexpr unsigned kNeedsOrient = 0x000FFF80u;   // anything beyond the 7 basis planes
    static constexpr unsigned kNeedsSteer = CVS_BIT(CVS_G2T) | CVS_BIT(CVS_H2T) | CVS_BIT(CVS_MAG) | CVS_BIT(CVS_PHASE) |
                                            CVS_BIT(CVS_EDGES) | CVS_BIT(CVS_DARK) | CVS_BIT(CVS_BRIGHT);

    // Fused epilogue on the 7 basis values of one pixel.  MASK != 0: compile-time plane set, steering at the
    // in-kernel dominant angle.  MASK == 0: run-time mask and steer source.
    template <unsigned MASK>
    __device__ __forceinline__ static void epilogue(const float (&b)[NBASIS], const MarchArgs& a, long long row_off, int x)
    {
        const unsigned m = MASK ? MASK : a.mask;
        const long long off = row_off + 4ll * x;
        auto put = [&](int p, float v) { *reinterpret_cast<float*>(reinterpret_cast<char*>(a.out[p]) + off) = v; };
#pragma unroll
        for (int q = 0; q < NBASIS; ++q)
            if (m & (1u << q)) put(q, b[q]);
        if (!(m & kNeedsOrient)) return;

        const int src = MASK ? (int)CVS_STEER_DOMINANT : a.steer_source;
        dev::Orientation o;
        const bool need_orient = (m & (CVS_BIT(CVS_C1) | CVS_BIT(CVS_C2) | CVS_BIT(CVS_C3) | CVS_BIT(CVS_THETA) |
                                       CVS_BIT(CVS_STRENGTH) | CVS_BIT(CVS_E))) || src == CVS_STEER_DOMINANT;
        if (need_orient) {
            o = dev::orientation_g2(b[0], b[1], b[2], b[3], b[4], b[5], b[6]);
            if (m & CVS_BIT(CVS_C1)) put(CVS_C1, o.c1);
            if (m & CVS_BIT(CVS_C2)) put(CVS_C2, o.c2);
            if (m & CVS_BIT(CVS_C3)) put(CVS_C3, o.c3);
            if (m & CVS_BIT(CVS_THETA)) put(CVS_THETA, o.theta);
            if (m & CVS_BIT(CVS_STRENGTH)) put(CVS_STRENGTH, o.strength);
        }
        if (!(m & (kNeedsSteer | CVS_BIT(CVS_E)))) return;

        if (src == CVS_STEER_DOMINANT && !(m & kNeedsSteer)) {
            // E(theta_d) = c1 + c2 cos(2 theta_d) + c3 sin(2 theta_d) = c1 + strength (G2.cpp:174-176 at theta_d)
            put(CVS_E, o.c1 + o.strength);
            return;
        }
        float ct, st;
        if (src == CVS_STEER_SCALAR) {
            ct = a.cos_t;
            st = a.sin_t;
        } else {
            const float th = (src == CVS_STEER_DOMINANT)
                                 ? o.theta
                                 : *reinterpret_cast<const float*>(reinterpret_cast<const char*>(a.theta_map) + off);
            sincosf(th, &st, &ct);
        }
        if (m & CVS_BIT(CVS_E)) {
            // cos 2t = c^2 - s^2, sin 2t = 2cs  (reference: polarToCart(2*theta), G2.cpp:175)
            put(CVS_E, fmaf(o.c2, fmaf(ct, ct, -st * st), fmaf(o.c3, 2.f * ct * st, o.c1)));
        }
        if (!(m & kNeedsSteer)) return;
        float g2, h2;
        dev::steer_g2(ct, st, b[0], b[1], b[2], b[3], b[4], b[5], b[6], g2, h2);
        if (m & CVS_BIT(CVS_G2T)) put(CVS_G2T, g2);
        if (m & CVS_BIT(CVS_H2T)) put(CVS_H2T, h2);
        if (m & (kNeedsSteer & ~(CVS_BIT(CVS_G2T) | CVS_BIT(CVS_H2T)))) {
            float mag, ph;
            dev::magnitude_phase(g2, h2, mag, ph);
            if (m & CVS_BIT(CVS_MAG)) put(CVS_MAG, mag);
            if (m & CVS_BIT(CVS_PHASE)) put(CVS_PHASE, ph);
            // find*(magnitude, phase): both reference callers feed magnitude (example/steer.cpp:88-90)
            if (m & CVS_BIT(CVS_EDGES)) put(CVS_EDGES, mag * dev::phase_weight(ph, 1.57079637050628662f, false));
            if (m & CVS_BIT(CVS_DARK)) put(CVS_DARK, mag * dev::phase_weight(ph, 0.f, true));
            if (m & CVS_BIT(CVS_BRIGHT)) put(CVS_BRIGHT, mag * dev::phase_weight(ph, 3.14159274101257324f, true));
        }
    }
};

struct G4Fam {
    static constexpr int R = 6, NSETS = 10, NROW = 10, NBASIS = 11, BH = 64, MIN_CTAS = 2;
    // unique tap sets: 0 g1, 1 g2(=h2), 2 g3, 3 g4, 4 g5, 5 h1, 6 h3, 7 h4, 8 h5, 9 h6
    // API order g1..g5,h1..h6
    __host__ __device__ static constexpr int unique_of(int api) { constexpr int t[11] = {0, 1, 2, 3, 4, 5, 1, 6, 7, 8, 9}; return t[api]; }
    __host__ __device__ static constexpr bool set_odd(int s) { return s == 2 || s == 3 || s == 5 || s == 7 || s == 8; }
    __host__ __device__ static constexpr int row_set(int p) { return p; }
    __host__ __device__ static constexpr bool row_odd(int p) { return set_odd(p); }
    //   G4a (g1,g2) G4b (g3,g4) G4c (g5,g5) G4d (g4,g3) G4e (g2,g1)
    //   H4a (h1,h2) H4b (h3,h4) H4c (h5,h6) H4d (h6,h5) H4e (h4,h3) H4f (h2,h1)
    __host__ __device__ static constexpr int basis_row(int q) { constexpr int t[11] = {0, 2, 4, 3, 1, 5, 6, 8, 9, 7, 1}; return t[q]; }
    __host__ __device__ static constexpr int basis_set(int q) { constexpr int t[11] = {1, 3, 4, 2, 0, 1, 7, 9, 8, 6, 5}; return t[q]; }
    __host__ __device__ static constexpr bool basis_odd(int q) { return set_odd(basis_set(q)); }

    static constexpr unsigned kNeedsSteer = CVS_G4_MASK_STEER;

    template <unsigned MASK>
    __device__ __forceinline__ static void epilogue(const float (&b)[NBASIS], const MarchArgs& a, long long row_off, int x)
    {
        const unsigned m = MASK ? MASK : a.mask;
        const long long off = row_off + 4ll * x;
        auto put = [&](int p, float v) { *reinterpret_cast<float*>(reinterpret_cast<char*>(a.out[p]) + off) = v; };
#pragma unroll
        for (int q = 0; q < NBASIS; ++q)
            if (m & (1u << q)) put(q, b[q]);
        if (!(m & kNeedsSteer)) return;
        float ct, st;
        if (a.steer_source == CVS_STEER_SCALAR) {
            ct = a.cos_t;
            st = a.sin_t;
        } else {
            const float th = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(a.theta_map) + off);
            sincosf(th, &st, &ct);
        }
        float g4, h4;
        dev::steer_g4(ct, st, &b[0], &b[5], g4, h4);
        if (m & CVS_BIT(CVS_G4T)) put(CVS_G4T, g4);
        if (m & CVS_BIT(CVS_H4T)) put(CVS_H4T, h4);
        if (m & (CVS_BIT(CVS_MAG4) | CVS_BIT(CVS_PHASE4))) {
            float mag, ph;
            dev::magnitude_phase(g4, h4, mag, ph);
            if (m & CVS_BIT(CVS_MAG4)) put(CVS_MAG4, mag);
            if (m & CVS_BIT(CVS_PHASE4)) put(CVS_PHASE4, ph);
        }
    }
};

}  // namespace cvs
