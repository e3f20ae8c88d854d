// fa::SteerableFiltersG2 -- drop-in for the reference class (cvsteer/SteerableFiltersG2.h:35-67): identical public
// signatures and protected member names; every body forwards to libcvsteer_b200 through the C ABI
// (include/cvsteer_c.h).  The set-up image's basis planes, C1..C3, dominant angle and strength stay RESIDENT ON THE
// GPU; the protected cv::Mat1f members are host mirrors that are downloaded lazily, the first time a getter (or
// syncHostMirrors()) asks for them, so code that only calls steer()/find*() never pays for 12 plane downloads.
#ifndef CVSTEER_B200_STEERABLEFILTERSG2_H_
#define CVSTEER_B200_STEERABLEFILTERSG2_H_

#include <cvsteer/SteerableFilters.h>

_STEER_BEGIN

class SteerableFiltersG2 : public SteerableFilters
{
public:
    SteerableFiltersG2(const cv::Mat1f& image, int width = 4, float spacing = 0.67);
    ~SteerableFiltersG2();
    SteerableFiltersG2(const SteerableFiltersG2&) = delete;  // owns a device handle
    SteerableFiltersG2& operator=(const SteerableFiltersG2&) = delete;

    const cv::Mat1f& getDominantOrientationAngle() const { return mirror(10, m_theta); }
    const cv::Mat1f& getDominantOrientationStrength() const { return mirror(11, m_orientationStrength); }

    void setup(const cv::Mat1f& image);

    // Steer filters at single point:
    void steer(const cv::Point& p, float theta, float& g2, float& h2);
    void steer(const cv::Point& p, float theta, float& g2, float& h2, float& e, float& magnitude, float& phase);
    void steer(const cv::Mat1f& theta, cv::Mat1f& g2, cv::Mat1f& h2);

    // Processing on entire images:
    void steer(float theta, cv::Mat1f& g2, cv::Mat1f& h2);
    void steer(float theta, cv::Mat1f& g2, cv::Mat1f& h2, cv::Mat1f& e, cv::Mat1f& magnitude, cv::Mat1f& phase);
    void steer(const cv::Mat1f& theta, cv::Mat1f& g2, cv::Mat1f& h2, cv::Mat1f& e, cv::Mat1f& magnitude, cv::Mat1f& phase);
    void computeMagnitudeAndPhase(const cv::Mat1f& g2, const cv::Mat1f& h2, cv::Mat1f& magnitude, cv::Mat1f& phase);

    void findEdges(const cv::Mat1f& e, const cv::Mat1f& phase, cv::Mat1f& output, float k = 2.0);
    void findDarkLines(const cv::Mat1f& e, const cv::Mat1f& phase, cv::Mat1f& output, float k = 2.0);
    void findBrightLines(const cv::Mat1f& e, const cv::Mat1f& phase, cv::Mat1f& output, float k = 2.0);

    static void phaseWeights(const cv::Mat1f& phase, cv::Mat1f& lambda, float phi, bool signum, float k);

    // 8-bit entry (not in the reference): uploads the gray bytes and converts on the device, instead of the implicit
    // Mat(8UC1) -> Mat1f host conversion both reference callers go through.
    void setup8u(const unsigned char* gray, int rows, int cols, size_t step);

protected:
    // download every protected plane member (for subclass code that reads m_g2a ... m_c3 directly)
    void syncHostMirrors() const;

    mutable cv::Mat1f m_dx, m_dy;
    cv::Mat1f m_g1, m_g2, m_g3, m_h1, m_h2, m_h3, m_h4;
    mutable cv::Mat1f m_g2a, m_g2b, m_g2c, m_h2a, m_h2b, m_h2c, m_h2d;
    mutable cv::Mat1f m_c1, m_c2, m_c3, m_theta, m_orientationStrength;

private:
    const cv::Mat1f& mirror(int plane, cv::Mat1f& m) const;
    void steerImpl(const cv::Mat1f* theta, float thetaScalar, cv::Mat1f* g2, cv::Mat1f* h2, cv::Mat1f* e, cv::Mat1f* magnitude,
                   cv::Mat1f* phase);
    cvs_g2* m_handle;
    int m_rows, m_cols;
    mutable unsigned m_mirrorValid;  // bit p set: host mirror of plane p is current
};

_STEER_END

#endif
