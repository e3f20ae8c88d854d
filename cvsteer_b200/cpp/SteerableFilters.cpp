// fa::SteerableFilters helpers (reference cvsteer/SteerableFilters.cpp:33-51) for the drop-in classes.
#include <cvsteer/SteerableFilters.h>

#include <string>

#include "cvsteer_c.h"

_STEER_BEGIN

// Tap sampling stays on the host, as in the reference: a (2*width+1)-point table is not GPU work.
cv::Mat1f SteerableFilters::create(int width, float spacing, KernelType f)
{
    cv::Mat1f kernel(1, width * 2 + 1);
    float* k = kernel.ptr(0);
    for (int i = 0; i <= 2 * width; ++i) k[i] = f(float(i - width) * spacing);
    return kernel;
}

// Host version for small helper Mats; whole-image wraps happen inside the CUDA kernels (dev::wrap_pi).
// Same arithmetic: compare against float(M_PI), replace by float(double(angle) - 2 pi).
void SteerableFilters::wrap(const cv::Mat1f& angle, cv::Mat1f& output)
{
    cv::Mat1f out(angle.rows, angle.cols);
    const float pi_f = static_cast<float>(M_PI);
    for (int r = 0; r < angle.rows; ++r) {
        const float* a = angle.ptr(r);
        float* o = out.ptr(r);
        for (int c = 0; c < angle.cols; ++c) o[c] = (a[c] > pi_f) ? static_cast<float>(-M_PI - (M_PI - static_cast<double>(a[c]))) : a[c];
    }
    output = out;
}

void detail::check(int status, const char* what)
{
    if (status == CVS_OK) return;
    const std::string msg = std::string(what) + ": " + cvs_last_error();
#ifdef CVSTEER_HAVE_OPENCV
    throw cv::Exception(status, msg, what, __FILE__, __LINE__);
#else
    throw cv::Exception(msg);
#endif
}

_STEER_END
