// fa::SteerableFiltersG4 -- drop-in for the reference class (cvsteer/SteerableFiltersG4.h:35-56).  As in the reference,
// there is no G4 orientation analysis: the getters return empty Mats (m_theta / m_orientationStrength are never
// assigned there either, SteerableFiltersG4.cpp:67-90).  computeMagnitudeAndPhase, an EMPTY body in the reference
// (G4.cpp:88-90), is implemented with the G2 class's definition (G2.cpp:107-112).
#ifndef CVSTEER_B200_STEERABLEFILTERSG4_H_
#define CVSTEER_B200_STEERABLEFILTERSG4_H_

#include <cvsteer/SteerableFilters.h>

_STEER_BEGIN

class SteerableFiltersG4 : public SteerableFilters
{
public:
    SteerableFiltersG4(const cv::Mat1f& image, int width = 6, float spacing = 0.5);
    ~SteerableFiltersG4();
    SteerableFiltersG4(const SteerableFiltersG4&) = delete;
    SteerableFiltersG4& operator=(const SteerableFiltersG4&) = delete;

    const cv::Mat1f& getDominantOrientationAngle() const { return m_theta; }
    const cv::Mat1f& getDominantOrientationStrength() const { return m_orientationStrength; }

    void setup(const cv::Mat1f& image);

    // Processing on entire images:
    void steer(const cv::Mat1f& theta, cv::Mat1f& g4, cv::Mat1f& h4);
    void steer(float theta, cv::Mat1f& g4, cv::Mat1f& h4);
    void computeMagnitudeAndPhase(const cv::Mat1f& g4, const cv::Mat1f& h4, cv::Mat1f& magnitude, cv::Mat1f& phase);

    // Extension (not in the reference, whose getters above stay empty because m_theta / m_orientationStrength are never
    // assigned): fills them with the dominant orientation of the G4/H4 oriented energy, defined like the G2 class's
    // (lowest-order Fourier terms of G4(theta)^2 + H4(theta)^2); see CVS_G4_THETA in cvsteer_c.h.
    void computeDominantOrientation();

protected:
    void syncHostMirrors() const;

    cv::Mat1f m_g1, m_g2, m_g3, m_g4, m_g5;
    cv::Mat1f m_h1, m_h2, m_h3, m_h4, m_h5, m_h6;
    mutable cv::Mat1f m_g4a, m_g4b, m_g4c, m_g4d, m_g4e;
    mutable cv::Mat1f m_h4a, m_h4b, m_h4c, m_h4d, m_h4e, m_h4f;
    cv::Mat1f m_c1, m_c2, m_c3, m_theta, m_orientationStrength;

private:
    cvs_g4* m_handle;
    int m_rows, m_cols;
};

_STEER_END

#endif
