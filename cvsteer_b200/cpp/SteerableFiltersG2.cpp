// fa::SteerableFiltersG2 over the C ABI of libcvsteer_b200 (drop-in for reference cvsteer/SteerableFiltersG2.cpp).
// Every method names the reference lines it replaces.  No pixel arithmetic happens here.
#include <cvsteer/SteerableFiltersG2.h>

#include <cstring>

#include "cvsteer_c.h"

_STEER_BEGIN

typedef cv::Mat1f Matf;

namespace {
const int kDevice = 0;  // one object = one image on one GPU; multi-GPU batches go through cvs_*_run_batch_dev

Matf taps(int which, int width, float spacing)
{
    Matf k(1, 2 * width + 1);
    cvs_g2_make_taps(which, width, spacing, k.ptr(0));
    return k;
}
inline float* outp(Matf* m, int rows, int cols)
{
    if (!m) return nullptr;
    m->create(rows, cols);  // the library never allocates caller memory
    return m->ptr(0);
}
}  // namespace

// G2.cpp:44-58 -- taps (host), then setup(image)
SteerableFiltersG2::SteerableFiltersG2(const Matf& image, int width, float spacing) : m_handle(nullptr), m_rows(0), m_cols(0), m_mirrorValid(0)
{
    detail::check(cvs_g2_create(&m_handle, kDevice, width, spacing), "SteerableFiltersG2::SteerableFiltersG2");
    m_g1 = taps(0, width, spacing), m_g2 = taps(1, width, spacing), m_g3 = taps(2, width, spacing);
    m_h1 = taps(3, width, spacing), m_h2 = taps(4, width, spacing), m_h3 = taps(5, width, spacing), m_h4 = taps(6, width, spacing);
    try {
        setup(image);
    } catch (...) {
        cvs_g2_destroy(m_handle);
        throw;
    }
}

SteerableFiltersG2::~SteerableFiltersG2() { cvs_g2_destroy(m_handle); }

// G2.cpp:60-100 -- 7x sepFilter2D + 16 products + C1..C3 + cartToPolar + wrap + *0.5, as ONE fused kernel
void SteerableFiltersG2::setup(const Matf& image)
{
    m_mirrorValid = 0;
    detail::check(cvs_g2_setup_host(m_handle, image.ptr(0), image.rows, image.cols, (size_t)image.step), "SteerableFiltersG2::setup");
    m_rows = image.rows, m_cols = image.cols;
#ifdef CVSTEER_EAGER_HOST_MIRRORS
    // strict drop-in for code that derives from this class and reads m_g2a ... m_theta without asking: fill every protected
    // member right away, as the reference's setup() does (12 plane downloads per image; the default is lazy)
    syncHostMirrors();
#endif
}

void SteerableFiltersG2::setup8u(const unsigned char* gray, int rows, int cols, size_t step)
{
    m_mirrorValid = 0;
    detail::check(cvs_g2_setup_host_u8(m_handle, gray, rows, cols, step), "SteerableFiltersG2::setup8u");
    m_rows = rows, m_cols = cols;
#ifdef CVSTEER_EAGER_HOST_MIRRORS
    syncHostMirrors();
#endif
}

const Matf& SteerableFiltersG2::mirror(int plane, Matf& m) const
{
    if (!(m_mirrorValid >> plane & 1u)) {
        m.create(m_rows, m_cols);
        detail::check(cvs_g2_get_plane_host(m_handle, plane, m.ptr(0), (size_t)m.step), "SteerableFiltersG2: download plane");
        m_mirrorValid |= 1u << plane;
    }
    return m;
}

void SteerableFiltersG2::syncHostMirrors() const
{
    Matf* planes[12] = {&m_g2a, &m_g2b, &m_g2c, &m_h2a, &m_h2b, &m_h2c, &m_h2d, &m_c1, &m_c2, &m_c3, &m_theta, &m_orientationStrength};
    for (int p = 0; p < 12; ++p) mirror(p, *planes[p]);
}

// G2.cpp:115-122
void SteerableFiltersG2::steer(const cv::Point& p, float theta, float& g2, float& h2)
{
    float out[5];
    detail::check(cvs_g2_steer_point(m_handle, p.x, p.y, theta, out), "SteerableFiltersG2::steer(point)");
    g2 = out[0], h2 = out[1];
}

// G2.cpp:124-134
void SteerableFiltersG2::steer(const cv::Point& p, float theta, float& g2, float& h2, float& e, float& magnitude, float& phase)
{
    float out[5];
    detail::check(cvs_g2_steer_point(m_handle, p.x, p.y, theta, out), "SteerableFiltersG2::steer(point)");
    g2 = out[0], h2 = out[1], e = out[2], magnitude = out[3], phase = out[4];
}

void SteerableFiltersG2::steerImpl(const Matf* theta, float thetaScalar, Matf* g2, Matf* h2, Matf* e, Matf* magnitude, Matf* phase)
{
    // The C ABI writes every requested output with ONE row step.  Mat::create() keeps an already allocated Mat of matching
    // size -- e.g. a ROI view whose step is wider than cols*4 -- so the outputs' own steps are honoured: all equal (the
    // normal case) -> written in place; otherwise the results land in dense temporaries and are copied row by row.
    Matf* dst[5] = {g2, h2, e, magnitude, phase};
    float* ptr[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t step = 0;
    bool uniform = true;
    for (int i = 0; i < 5; ++i) {
        ptr[i] = outp(dst[i], m_rows, m_cols);
        if (!dst[i]) continue;
        if (!step) step = (size_t)dst[i]->step;
        uniform = uniform && (size_t)dst[i]->step == step;
    }
    Matf tmp[5];
    if (!uniform) {
        step = (size_t)m_cols * sizeof(float);
        for (int i = 0; i < 5; ++i)
            if (dst[i]) tmp[i] = Matf(m_rows, m_cols), ptr[i] = tmp[i].ptr(0);
    }
    if (!theta) {
        detail::check(cvs_g2_steer_scalar_host(m_handle, thetaScalar, ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], step), "SteerableFiltersG2::steer(float)");
    } else {
        if (theta->rows != m_rows || theta->cols != m_cols) detail::check(CVS_ERR_SIZE_MISMATCH, "SteerableFiltersG2::steer: theta size differs from the image");
        // steer(getDominantOrientationAngle(), ...): the callers' idiom (example/steer.cpp:87, test/test.cpp:86).  When the
        // argument IS our own mirror of theta_d, steer from the device-resident map instead of uploading it again.
        const bool own = (m_mirrorValid >> 10 & 1u) && theta->ptr(0) == m_theta.ptr(0);
        detail::check(cvs_g2_steer_map_host(m_handle, own ? nullptr : theta->ptr(0), (size_t)theta->step, ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], step),
                      "SteerableFiltersG2::steer(Mat1f)");
    }
    if (!uniform)
        for (int i = 0; i < 5; ++i)
            if (dst[i])
                for (int r = 0; r < m_rows; ++r) std::memcpy(dst[i]->ptr(r), tmp[i].ptr(r), sizeof(float) * (size_t)m_cols);
}

// G2.cpp:137-145
void SteerableFiltersG2::steer(float theta, Matf& g2, Matf& h2) { steerImpl(nullptr, theta, &g2, &h2, nullptr, nullptr, nullptr); }
// G2.cpp:147-155
void SteerableFiltersG2::steer(const Matf& theta, Matf& g2, Matf& h2) { steerImpl(&theta, 0.f, &g2, &h2, nullptr, nullptr, nullptr); }
// G2.cpp:157-165
void SteerableFiltersG2::steer(float theta, Matf& g2, Matf& h2, Matf& e, Matf& magnitude, Matf& phase)
{
    steerImpl(nullptr, theta, &g2, &h2, &e, &magnitude, &phase);
}
// G2.cpp:167-177
void SteerableFiltersG2::steer(const Matf& theta, Matf& g2, Matf& h2, Matf& e, Matf& magnitude, Matf& phase)
{
    steerImpl(&theta, 0.f, &g2, &h2, &e, &magnitude, &phase);
}

// G2.cpp:107-112
void SteerableFiltersG2::computeMagnitudeAndPhase(const Matf& g2, const Matf& h2, Matf& magnitude, Matf& phase)
{
    if (g2.rows != h2.rows || g2.cols != h2.cols || g2.step != h2.step) detail::check(CVS_ERR_SIZE_MISMATCH, "computeMagnitudeAndPhase: g2/h2 differ");
    magnitude.create(g2.rows, g2.cols);
    phase.create(g2.rows, g2.cols);
    detail::check(cvs_magnitude_phase_host(kDevice, g2.ptr(0), h2.ptr(0), (size_t)g2.step, magnitude.ptr(0), phase.ptr(0), (size_t)magnitude.step, g2.rows,
                                   g2.cols),
          "SteerableFiltersG2::computeMagnitudeAndPhase");
}

// G2.cpp:179-186
void SteerableFiltersG2::phaseWeights(const Matf& phase, Matf& lambda, float phi, bool signum, float k)
{
    Matf out(phase.rows, phase.cols);
    detail::check(cvs_phase_weights_host(kDevice, phase.ptr(0), (size_t)phase.step, out.ptr(0), (size_t)out.step, phase.rows, phase.cols, phi, signum ? 1 : 0, k),
          "SteerableFiltersG2::phaseWeights");
    lambda = out;
}

namespace {
void find(int kind, const Matf& e, const Matf& phase, Matf& output, float k, const char* what)
{
    if (e.rows != phase.rows || e.cols != phase.cols || e.step != phase.step) detail::check(CVS_ERR_SIZE_MISMATCH, what);
    Matf out(e.rows, e.cols);  // separate buffer: the reference allows output to alias e
    detail::check(cvs_find_host(kDevice, kind, e.ptr(0), phase.ptr(0), (size_t)e.step, out.ptr(0), (size_t)out.step, e.rows, e.cols, k), what);
    output = out;
}
}  // namespace

// G2.cpp:201-212
void SteerableFiltersG2::findEdges(const Matf& e, const Matf& phase, Matf& output, float k) { find(0, e, phase, output, k, "SteerableFiltersG2::findEdges"); }
void SteerableFiltersG2::findDarkLines(const Matf& e, const Matf& phase, Matf& output, float k) { find(1, e, phase, output, k, "SteerableFiltersG2::findDarkLines"); }
void SteerableFiltersG2::findBrightLines(const Matf& e, const Matf& phase, Matf& output, float k) { find(2, e, phase, output, k, "SteerableFiltersG2::findBrightLines"); }

_STEER_END
