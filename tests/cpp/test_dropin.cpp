// C++ side of tests/test_cpp_dropin_gpu.py: the reference's own test flow (test/test.cpp:80-90) written against the
// drop-in classes, exactly as a user of fa::SteerableFiltersG2 writes it.  Image I/O is raw binary (no codec in this
// image); the Python test generates the input, runs this program and compares every output with the oracle.
//   usage: test_dropin <in.u8> <rows> <cols> <outdir>
#include <cvsteer/SteerableFiltersG2.h>
#include <cvsteer/SteerableFiltersG4.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static void save(const std::string& dir, const char* name, const cv::Mat1f& m)
{
    FILE* f = fopen((dir + "/" + name + ".f32").c_str(), "wb");
    for (int r = 0; r < m.rows; ++r) fwrite(m.ptr(r), sizeof(float), m.cols, f);
    fclose(f);
}

// a subclass, to prove that protected members keep working for derived code (G2.h:62-66)
struct Probe : fa::SteerableFiltersG2 {
    using fa::SteerableFiltersG2::SteerableFiltersG2;
    const cv::Mat1f& g2a()
    {
        syncHostMirrors();
        return m_g2a;
    }
    const cv::Mat1f& c1() { return m_c1; }
    const cv::Mat1f& g2bRaw() { return m_g2b; }  // no sync: filled only by an eager-mirror build
    const cv::Mat1f& tapsG1() { return m_g1; }
};

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    const int rows = atoi(argv[2]), cols = atoi(argv[3]);
    const std::string dir = argv[4];
    std::vector<unsigned char> bytes((size_t)rows * cols);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(bytes.data(), 1, bytes.size(), f) != bytes.size()) return 3;
    fclose(f);
    cv::Mat1b fish(rows, cols, bytes.data());

    try {
        // --- test/test.cpp:84-90 verbatim (the 8-bit Mat converts to Mat1f implicitly, as in the reference)
        cv::Mat1f g2, h2, e, magnitude, phase, edges, linesDark, linesBright;
        Probe filters2(fish, 4, 0.67f);
#ifdef CVSTEER_EAGER_HOST_MIRRORS
        if (filters2.g2bRaw().empty() || filters2.g2bRaw().rows != rows) return 10;  // protected members filled by setup() itself
#else
        if (!filters2.g2bRaw().empty()) return 10;                                    // lazy by default
#endif
        filters2.steer(filters2.getDominantOrientationAngle(), g2, h2, e, magnitude, phase);
        filters2.findEdges(magnitude, phase, edges);
        filters2.findDarkLines(magnitude, phase, linesDark);
        filters2.findBrightLines(magnitude, phase, linesBright);
        save(dir, "theta", filters2.getDominantOrientationAngle());
        save(dir, "strength", filters2.getDominantOrientationStrength());
        save(dir, "g2", g2), save(dir, "h2", h2), save(dir, "e", e), save(dir, "magnitude", magnitude), save(dir, "phase", phase);
        save(dir, "edges", edges), save(dir, "linesDark", linesDark), save(dir, "linesBright", linesBright);
        save(dir, "g2a", filters2.g2a()), save(dir, "c1", filters2.c1()), save(dir, "taps_g1", filters2.tapsG1());

        // scalar + point overloads
        cv::Mat1f sg, sh;
        filters2.steer(0.3f, sg, sh);
        save(dir, "g2_s03", sg), save(dir, "h2_s03", sh);
        // outputs that are views with a wider row step (a ROI of a larger Mat): create() keeps them, every row must land at
        // the view's own step (reference: Mat expressions honour step); g2 and h2 views deliberately differ in step
        {
            std::vector<float> bufg((size_t)rows * (cols + 7), -1.f), bufh((size_t)rows * (cols + 3), -1.f);
            cv::Mat1f vg(rows, cols, bufg.data(), sizeof(float) * (cols + 7)), vh(rows, cols, bufh.data(), sizeof(float) * (cols + 3));
            filters2.steer(0.3f, vg, vh);
            if (vg.ptr(0) != bufg.data() || vh.ptr(0) != bufh.data()) return 7;  // must not have been reallocated
            for (int r = 0; r < rows; ++r) {
                if (memcmp(vg.ptr(r), sg.ptr(r), sizeof(float) * cols) || memcmp(vh.ptr(r), sh.ptr(r), sizeof(float) * cols)) return 8;
                for (int c = cols; c < cols + 3; ++c)
                    if (bufg[(size_t)r * (cols + 7) + c] != -1.f || bufh[(size_t)r * (cols + 3) + c] != -1.f) return 9;  // padding untouched
            }
        }
        float pg, ph, pe, pm, pp;
        filters2.steer(cv::Point(17, 5), 0.3f, pg, ph, pe, pm, pp);
        cv::Mat1f pt(1, 5);
        pt(0) = pg, pt(1) = ph, pt(2) = pe, pt(3) = pm, pt(4) = pp;
        save(dir, "point", pt);
        cv::Mat1f lam;
        fa::SteerableFiltersG2::phaseWeights(phase, lam, 1.0f, true, 2.0f);
        save(dir, "lambda", lam);

        // --- G4
        cv::Mat1f g4, h4, m4, p4;
        fa::SteerableFiltersG4 filters4(fish, 6, 0.5f);
        filters4.steer(filters2.getDominantOrientationAngle(), g4, h4);
        filters4.computeMagnitudeAndPhase(g4, h4, m4, p4);
        save(dir, "g4", g4), save(dir, "h4", h4), save(dir, "mag4", m4), save(dir, "phase4", p4);
        cv::Mat1f g4s, h4s;
        filters4.steer(0.3f, g4s, h4s);
        save(dir, "g4_s03", g4s);
        if (!filters4.getDominantOrientationAngle().empty()) return 5;  // never assigned in the reference either
        filters4.computeDominantOrientation();                          // extension
        save(dir, "theta4", filters4.getDominantOrientationAngle());
        save(dir, "strength4", filters4.getDominantOrientationStrength());

        // --- error behaviour: a failing C-ABI call surfaces as cv::Exception, like OpenCV's CV_Assert would
        bool threw = false;
        try {
            cv::Mat1f bad(3, 3), a, b;
            filters2.steer(bad, a, b);
        } catch (const cv::Exception&) {
            threw = true;
        }
        if (!threw) return 6;
    } catch (const std::exception& ex) {
        fprintf(stderr, "exception: %s\n", ex.what());
        return 4;
    }
    printf("dropin ok\n");
    return 0;
}
