// Internals shared by the translation units behind the C ABI (capi.cu, bands.cu).  Not installed, not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <utility>
#include <vector>

#include "../../include/cvsteer_c.h"
#include "launch.h"
#include "taps.h"

namespace cvsi {
using namespace cvs;

// thread-local error text of the calling thread (cvs_last_error); returns `code`
int fail(int code, const char* fmt, ...);

#define CU_TRY(expr)                                                                                         \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) return ::cvsi::fail(CVS_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

struct DevBuf {  // owning device allocation that only ever grows; freed on release() or destruction
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr, o.bytes = 0; }
    ~DevBuf() { release(); }
    cudaError_t reserve(size_t n)
    {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Filter {
    int family;  // 2 or 4
    int device;
    FamilyTaps taps;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};  // H2D / kernel / D2H pipeline of the host-batch call (lazy)
    // resident class state of the last setup()
    int rows = 0, cols = 0;
    size_t pitch = 0;         // bytes, multiple of 128 (TMA needs 16)
    DevBuf in, state, work, scratch;
    cudaStream_t scratch_stream[3] = {nullptr, nullptr, nullptr};  // last stream that used each slice of `scratch` (generic widths)
    bool scratch_used[3] = {false, false, false};
    cudaEvent_t scratch_ev = nullptr;
    int nstate = 0;           // planes in `state`: G2 12 (7 basis, c1..c3, theta, strength); G4 11
    bool ready = false;
    bool g4_orient_ready = false;  // planes 11, 12 of a G4 handle hold theta_d / strength of the current image
    LaunchInfo last{};

    float* state_plane(int i) const { return reinterpret_cast<float*>(static_cast<char*>(state.p) + (size_t)i * pitch * rows); }
    float* work_plane(int i) const { return reinterpret_cast<float*>(static_cast<char*>(work.p) + (size_t)i * pitch * rows); }
};

BatchGeom whole_frame_geom(const void* in, bool u8, int n, int rows, int cols, size_t in_pitch, size_t in_fs, size_t out_pitch,
                           size_t out_fs);
// `slot` of `nslots`: callers that keep several launches in flight on different streams give each stream its own slice of
// the generic-width scratch
int run_fused(Filter* f, const BatchGeom& g, unsigned mask, const SteerSpec& st, float* const* outs, cudaStream_t stream, int slot = 0,
              int nslots = 1);
int filter_create(Filter** out, int family, int device, int width, float spacing);
int filter_destroy(Filter* f);

struct BandPlanC {
    std::vector<int> rows;                       // image height per level
    std::vector<std::pair<int, int>> out, have;  // [lo, hi) per level: rows produced / rows that must be resident
    bool empty() const { return out[0].first >= out[0].second; }
};
// Same rule as cvsteer_b200/multi.py::plan_bands (band edges on multiples of 2^(levels-1) rows)
std::vector<BandPlanC> plan_bands_c(int rows, int world, int levels, int radius);

}  // namespace cvsi
