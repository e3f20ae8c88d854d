// Host-side tap generation for the Freeman-Adelson x-y separable basis filters.
// Restates SteerableFilters::create (reference cvsteer/SteerableFilters.cpp:33-42) and the tap
// functions G21..G23,H21..H24 (cvsteer/SteerableFiltersG2.cpp:35-42), G41..G45,H41..H46
// (cvsteer/SteerableFiltersG4.cpp:34-45): Tables III/IV/VI of Freeman & Adelson, PAMI 13(9) 1991,
// sampled at x = float(i)*spacing, i = -width..width.
#pragma once

namespace cvs {

enum { G2_NUM_TAPSETS = 7, G4_NUM_TAPSETS = 11, MAX_WIDTH = 32, MAX_TAPS = 2 * MAX_WIDTH + 1 };

// index order: G2 family g1,g2,g3,h1,h2,h3,h4;  G4 family g1..g5,h1..h6
void make_taps_g2(int which, int width, float spacing, float* dst);
void make_taps_g4(int which, int width, float spacing, float* dst);

// dst[i] = src[i] * (num[k] / den[k]) for the k with the largest |den[k]|, evaluated in double and rounded once.
// Used where two tap functions are the same function up to a constant (g3 = sqrt(1.843) h3 in the G2 family,
// g4 = 1.246 h4 in the G4 family), so that one row pass can serve both.
void scale_taps_by_ratio(const float* src, const float* num, const float* den, int n, float* dst);

// parity of each tap set: 0 = even (f[-i]==f[i]), 1 = odd (f[-i]==-f[i], f[0]==0); exact in fp32.
extern const int kG2TapOdd[G2_NUM_TAPSETS];
extern const int kG4TapOdd[G4_NUM_TAPSETS];

}  // namespace cvs
