#include "taps.h"

#include <cmath>

namespace cvs {
namespace {

// Every tap is  poly(x) * exp(-x^2).  The reference's expressions mix `float x` with double literals:
// the polynomial is evaluated in double (except sub-terms written with float literals / pure float
// products), std::exp(-x*x) takes a float argument and so is the float overload, and the product is
// narrowed to float on return.  The lambdas below keep exactly that typing so the taps are bit-equal
// to the reference's on the same libm.
typedef float (*TapFn)(float);

inline float gauss(float x) { return std::exp(-x * x); }  // float overload, as in the reference

// --- second derivative of Gaussian and its Hilbert transform (Table III) -- G2.cpp:35-42
const TapFn kG2[G2_NUM_TAPSETS] = {
    [](float x) -> float { return 0.9213 * (2.0 * x * x - 1.0) * gauss(x); },            // g1  (G21)
    [](float x) -> float { return gauss(x); },                                             // g2  (G22)
    [](float x) -> float { return std::sqrt(1.8430) * x * gauss(x); },                     // g3  (G23)
    [](float x) -> float { return 0.9780 * (-2.254 * x + x * x * x) * gauss(x); },         // h1  (H21)
    [](float x) -> float { return gauss(x); },                                             // h2  (H22)
    [](float x) -> float { return x * gauss(x); },                                         // h3  (H23)
    [](float x) -> float { return 0.9780 * (-0.7515 + x * x) * gauss(x); },                // h4  (H24)
};

// --- fourth derivative of Gaussian and its Hilbert transform (Table VI) -- G4.cpp:34-45
const TapFn kG4[G4_NUM_TAPSETS] = {
    [](float x) -> float { return 1.246 * (0.75 - 3.0f * x * x + x * x * x * x) * gauss(x); },                  // g1
    [](float x) -> float { return gauss(x); },                                                                  // g2
    [](float x) -> float { return (-1.5 * x + x * x * x) * gauss(x); },                                         // g3
    [](float x) -> float { return 1.246 * x * gauss(x); },                                                      // g4
    [](float x) -> float { return std::sqrt(1.246) * (x * x - 0.5) * gauss(x); },                               // g5
    [](float x) -> float { return 0.3975 * (7.189 * x - 7.501 * x * x * x + x * x * x * x * x) * gauss(x); },   // h1
    [](float x) -> float { return gauss(x); },                                                                  // h2
    [](float x) -> float { return 0.3975 * (1.438 - 4.501 * x * x + x * x * x * x) * gauss(x); },               // h3
    [](float x) -> float { return x * gauss(x); },                                                              // h4
    [](float x) -> float { return 0.3975 * (x * x * x - 2.225 * x) * gauss(x); },                               // h5
    [](float x) -> float { return (x * x - 0.6638) * gauss(x); },                                               // h6
};

void sample(TapFn f, int width, float spacing, float* dst)
{
    for (int i = -width; i <= width; ++i) dst[i + width] = f(float(i) * spacing);
}

}  // namespace

const int kG2TapOdd[G2_NUM_TAPSETS] = {0, 0, 1, 1, 0, 1, 0};
const int kG4TapOdd[G4_NUM_TAPSETS] = {0, 0, 1, 1, 0, 1, 0, 0, 1, 1, 0};

void scale_taps_by_ratio(const float* src, const float* num, const float* den, int n, float* dst)
{
    int k = 0;
    for (int i = 1; i < n; ++i)
        if (std::fabs(den[i]) > std::fabs(den[k])) k = i;
    const double ratio = den[k] != 0.f ? (double)num[k] / (double)den[k] : 0.0;
    for (int i = 0; i < n; ++i) dst[i] = (float)((double)src[i] * ratio);
}

void make_taps_g2(int which, int width, float spacing, float* dst) { sample(kG2[which], width, spacing, dst); }
void make_taps_g4(int which, int width, float spacing, float* dst) { sample(kG4[which], width, spacing, dst); }

}  // namespace cvs
