// fa::SteerableFiltersG4 over the C ABI of libcvsteer_b200 (drop-in for reference cvsteer/SteerableFiltersG4.cpp).
#include <cvsteer/SteerableFiltersG4.h>

#include <cstring>

#include "cvsteer_c.h"

_STEER_BEGIN

namespace {
const int kDevice = 0;
cv::Mat1f taps(int which, int width, float spacing)
{
    cv::Mat1f k(1, 2 * width + 1);
    cvs_g4_make_taps(which, width, spacing, k.ptr(0));
    return k;
}
void copyRows(const cv::Mat1f& src, cv::Mat1f& dst)
{
    for (int r = 0; r < src.rows; ++r) std::memcpy(dst.ptr(r), src.ptr(r), sizeof(float) * (size_t)src.cols);
}
}  // namespace

// G4.cpp:47-65
SteerableFiltersG4::SteerableFiltersG4(const cv::Mat1f& image, int width, float spacing) : m_handle(nullptr), m_rows(0), m_cols(0)
{
    detail::check(cvs_g4_create(&m_handle, kDevice, width, spacing), "SteerableFiltersG4::SteerableFiltersG4");
    cv::Mat1f* t[11] = {&m_g1, &m_g2, &m_g3, &m_g4, &m_g5, &m_h1, &m_h2, &m_h3, &m_h4, &m_h5, &m_h6};
    for (int i = 0; i < 11; ++i) *t[i] = taps(i, width, spacing);
    try {
        setup(image);
    } catch (...) {
        cvs_g4_destroy(m_handle);
        throw;
    }
}

SteerableFiltersG4::~SteerableFiltersG4() { cvs_g4_destroy(m_handle); }

// G4.cpp:67-81 -- 11x sepFilter2D as one fused kernel; planes stay on the device
void SteerableFiltersG4::setup(const cv::Mat1f& image)
{
    detail::check(cvs_g4_setup_host(m_handle, image.ptr(0), image.rows, image.cols, (size_t)image.step), "SteerableFiltersG4::setup");
    m_rows = image.rows, m_cols = image.cols;
    cv::Mat1f* p[11] = {&m_g4a, &m_g4b, &m_g4c, &m_g4d, &m_g4e, &m_h4a, &m_h4b, &m_h4c, &m_h4d, &m_h4e, &m_h4f};
    for (int i = 0; i < 11; ++i) *p[i] = cv::Mat1f();  // mirrors are stale until syncHostMirrors()
    m_theta = cv::Mat1f();
    m_orientationStrength = cv::Mat1f();
#ifdef CVSTEER_EAGER_HOST_MIRRORS
    syncHostMirrors();  // strict drop-in for subclasses that read m_g4a ... m_h4f directly (11 plane downloads per image)
#endif
}

void SteerableFiltersG4::syncHostMirrors() const
{
    cv::Mat1f* p[11] = {&m_g4a, &m_g4b, &m_g4c, &m_g4d, &m_g4e, &m_h4a, &m_h4b, &m_h4c, &m_h4d, &m_h4e, &m_h4f};
    for (int i = 0; i < 11; ++i) {
        p[i]->create(m_rows, m_cols);
        detail::check(cvs_g4_get_plane_host(m_handle, i, p[i]->ptr(0), (size_t)p[i]->step), "SteerableFiltersG4: download plane");
    }
}

// G4.cpp:92-112
void SteerableFiltersG4::steer(const cv::Mat1f& theta, cv::Mat1f& g4, cv::Mat1f& h4)
{
    if (theta.rows != m_rows || theta.cols != m_cols) detail::check(CVS_ERR_SIZE_MISMATCH, "SteerableFiltersG4::steer: theta size differs from the image");
    g4.create(m_rows, m_cols);
    h4.create(m_rows, m_cols);
    const bool same = g4.step == h4.step;  // one step for both outputs in the C ABI; differing ROI steps go through a dense copy
    cv::Mat1f t4 = same ? h4 : cv::Mat1f(m_rows, m_cols), tg = same ? g4 : cv::Mat1f(m_rows, m_cols);
    detail::check(cvs_g4_steer_map_host(m_handle, theta.ptr(0), (size_t)theta.step, tg.ptr(0), t4.ptr(0), nullptr, nullptr, (size_t)tg.step),
          "SteerableFiltersG4::steer(Mat1f)");
    if (!same) copyRows(tg, g4), copyRows(t4, h4);
}

// G4.cpp:114-122
void SteerableFiltersG4::steer(float theta, cv::Mat1f& g4, cv::Mat1f& h4)
{
    g4.create(m_rows, m_cols);
    h4.create(m_rows, m_cols);
    const bool same = g4.step == h4.step;
    cv::Mat1f t4 = same ? h4 : cv::Mat1f(m_rows, m_cols), tg = same ? g4 : cv::Mat1f(m_rows, m_cols);
    detail::check(cvs_g4_steer_scalar_host(m_handle, theta, tg.ptr(0), t4.ptr(0), nullptr, nullptr, (size_t)tg.step), "SteerableFiltersG4::steer(float)");
    if (!same) copyRows(tg, g4), copyRows(t4, h4);
}

void SteerableFiltersG4::computeDominantOrientation()
{
    m_theta.create(m_rows, m_cols);
    m_orientationStrength.create(m_rows, m_cols);
    detail::check(cvs_g4_get_plane_host(m_handle, CVS_G4_THETA, m_theta.ptr(0), (size_t)m_theta.step), "SteerableFiltersG4: theta");
    detail::check(cvs_g4_get_plane_host(m_handle, CVS_G4_STRENGTH, m_orientationStrength.ptr(0), (size_t)m_orientationStrength.step),
                  "SteerableFiltersG4: strength");
}

// G4.cpp:88-90 is an empty body; defined here as the G2 class defines it (G2.cpp:107-112)
void SteerableFiltersG4::computeMagnitudeAndPhase(const cv::Mat1f& g4, const cv::Mat1f& h4, cv::Mat1f& magnitude, cv::Mat1f& phase)
{
    if (g4.rows != h4.rows || g4.cols != h4.cols || g4.step != h4.step) detail::check(CVS_ERR_SIZE_MISMATCH, "computeMagnitudeAndPhase: g4/h4 differ");
    magnitude.create(g4.rows, g4.cols);
    phase.create(g4.rows, g4.cols);
    detail::check(cvs_magnitude_phase_host(kDevice, g4.ptr(0), h4.ptr(0), (size_t)g4.step, magnitude.ptr(0), phase.ptr(0), (size_t)magnitude.step, g4.rows,
                                   g4.cols),
          "SteerableFiltersG4::computeMagnitudeAndPhase");
}

_STEER_END
