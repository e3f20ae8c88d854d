// fa::SteerableFilters -- abstract base of the steerable-filter classes.  Same surface as the reference
// (cvsteer/SteerableFilters.h:41-50): pure-virtual setup()/steer(), protected static create()/wrap().
// In this implementation create() calls the host tap generator of libcvsteer_b200 (cvs_g2_make_taps /
// cvs_g4_make_taps restate SteerableFilters.cpp:33-42 with the reference's tap functions); wrap() runs on the host
// for the small helper Mats that still use it (the per-image wrap is fused into the CUDA kernels).
#ifndef CVSTEER_B200_STEERABLEFILTERS_H_
#define CVSTEER_B200_STEERABLEFILTERS_H_

#define _USE_MATH_DEFINES
#include <cmath>
#include <math.h>
#include <stdint.h>

#include <cvsteer/cvsteer.h>

#if defined(__has_include)
#if __has_include(<opencv2/core/core.hpp>) && !defined(CVSTEER_FORCE_SHIM)
#include <opencv2/core/core.hpp>
#define CVSTEER_HAVE_OPENCV 1
#endif
#endif
#ifndef CVSTEER_HAVE_OPENCV
#include <cvsteer/cv_shim.h>
#endif

struct cvs_g2;
struct cvs_g4;

_STEER_BEGIN

namespace detail
{
// throws cv::Exception (a std::runtime_error with the shim) when a C-ABI call returned an error status
void check(int status, const char* what);
}

class SteerableFilters
{
public:
    virtual void setup(const cv::Mat1f& image) = 0;
    virtual void steer(float theta, cv::Mat1f& g2, cv::Mat1f& h2) = 0;
    virtual ~SteerableFilters() {}  // the reference omits this (Appendix D of SURVEY.md); harmless to add

protected:
    typedef float (*KernelType)(float x);
    // reference signature kept for subclass code; samples f at float(i)*spacing, i = -width..width
    static cv::Mat1f create(int width, float spacing, KernelType f);
    static void wrap(const cv::Mat1f& angle, cv::Mat1f& output);
};

_STEER_END

#endif
