// Row-band mode behind the C ABI: ONE very large image split into row bands over several GPUs (SURVEY section 8e,
// BASELINE.json configs[4]).  A `cvs_bands` context is one rank (= one GPU) of such a run: it owns the rank's slice of every
// pyramid level (band + halo), launches the band-mode pyr_down / fused kernels level by level, and delivers its rows to the
// ROOT GPU's full-size planes in one of three ways:
//
//   CVS_GATHER_PEER_STORE  the fused kernel's output pointers ARE the root's planes (NVLink peer memory): compute and gather
//                          are one kernel, the transfer overlaps the arithmetic store by store, no local planes exist;
//   CVS_GATHER_PEER_COPY   the kernel writes local planes; a copy engine moves each level's rows to the root's planes on a
//                          second stream while the next level computes (SMs never wait on NVLink);
//   CVS_GATHER_NCCL        the same overlap with grouped ncclSend / ncclRecv per level (for ranks that cannot map the
//                          root's memory, e.g. across nodes).
//
// The reference has no counterpart (one image, one CPU thread: example/steer.cpp:69-124); parity is "bit-identical to the
// single-GPU whole-image pyramid" (tests/test_bands_gpu.py).  NCCL is bound at run time (dlopen: the copy already loaded
// into the process, e.g. by torch, else the system one), so the library has no link-time dependency on it.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "internal.h"

using namespace cvs;
using namespace cvsi;

namespace {

// ------------------------------------------------------------------------------------------------
// NCCL, bound lazily
// ------------------------------------------------------------------------------------------------
struct NcclUid {
    char internal[128];
};
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUid*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
enum { kNcclFloat = 7, kNcclSum = 0 };

NcclApi& nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
        if ((api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL))) break;  // the copy torch (or the caller) already loaded
    if (!api.lib)
        for (const char* n : names)
            if ((api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!api.lib) return api;
    auto sym = [&](const char* s) { return dlsym(api.lib, s); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv && api.AllReduce;
    return api;
}

#define NCCL_TRY(expr)                                                                                              \
    do {                                                                                                            \
        int r__ = (expr);                                                                                           \
        if (r__ != 0) return fail(CVS_ERR_CUDA, "%s: %s", #expr, nccl().GetErrorString ? nccl().GetErrorString(r__) : "nccl error"); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// the context
// ------------------------------------------------------------------------------------------------
// PEER_COPY: a level's output rows are produced in up to kMaxChunks launches so that the copy engine can move chunk k while
// chunk k+1 computes (matters when few GPUs share the image: at N = 2 a band's level-0 kernel takes 4 ms)
enum { kMaxChunks = 8, kMinChunkRows = 512 };

struct PlaneRef {
    float* p = nullptr;
    size_t pitch = 0;  // bytes
};

struct Bands {
    int device = 0, rank = 0, world = 1, root = 0;
    int rows = 0, cols = 0, levels = 1, width = 4;
    unsigned mask = 0;
    int nsel = 0;
    int sel[CVS_G2_NPLANES];          // selected plane ids, ascending
    std::vector<BandPlanC> plans;     // every rank's plan (the root needs them for the receives)
    std::vector<int> lc;              // columns per level
    std::vector<size_t> pitch;        // row pitch per level (bytes): level buffers, local planes and the root's own planes
    Filter* f = nullptr;
    std::vector<DevBuf> lv;           // this rank's slice of every level (rows plans[rank].have[l])
    DevBuf local;                     // this rank's output rows, all levels/planes back to back (not used by PEER_STORE)
    std::vector<size_t> local_off;    // [level * nsel + k] byte offset into `local`
    // the root's full-size planes as THIS rank can address them: own allocation (root), IPC mapping or same-process pointer
    std::vector<PlaneRef> rootp;      // [level * nsel + k]
    void* root_block = nullptr;       // allocation (root) or mapping (importer) behind rootp, when made by this library
    bool root_owned = false, root_mapped = false;
    void* comm = nullptr;             // ncclComm_t
    bool comm_owned = false;
    cudaStream_t xfer = nullptr;      // transfers run here, ordered against the compute stream by events
    std::vector<cudaEvent_t> ev;      // per (level, chunk): fused kernel of that chunk done
    cudaEvent_t ev_xfer = nullptr, ev_start = nullptr;
    float* token = nullptr;           // 1 float for the on-stream barrier

    const BandPlanC& plan() const { return plans[rank]; }
    int out_rows(int l, int r) const { return plans[r].out[l].second - plans[r].out[l].first; }
};

size_t root_layout(const Bands& b, std::vector<size_t>* offs)
{
    size_t off = 0;
    if (offs) offs->assign((size_t)b.levels * b.nsel, 0);
    for (int l = 0; l < b.levels; ++l)
        for (int k = 0; k < b.nsel; ++k) {
            if (offs) (*offs)[(size_t)l * b.nsel + k] = off;
            off += align_up(b.pitch[l] * (size_t)b.plans[0].rows[l], 256);
        }
    return off;
}

void set_root_block(Bands& b, void* base)
{
    std::vector<size_t> offs;
    root_layout(b, &offs);
    b.rootp.assign((size_t)b.levels * b.nsel, PlaneRef{});
    for (int l = 0; l < b.levels; ++l)
        for (int k = 0; k < b.nsel; ++k)
            b.rootp[(size_t)l * b.nsel + k] = PlaneRef{reinterpret_cast<float*>(static_cast<char*>(base) + offs[(size_t)l * b.nsel + k]), b.pitch[l]};
}

int ensure_local(Bands& b)
{
    if (b.local.p) return CVS_OK;
    size_t off = 0;
    b.local_off.assign((size_t)b.levels * b.nsel, 0);
    for (int l = 0; l < b.levels; ++l)
        for (int k = 0; k < b.nsel; ++k) {
            b.local_off[(size_t)l * b.nsel + k] = off;
            off += align_up(b.pitch[l] * (size_t)std::max(0, b.out_rows(l, b.rank)), 256);
        }
    CU_TRY(b.local.reserve(std::max<size_t>(off, 256)));
    return CVS_OK;
}

}  // namespace

struct cvs_bands : Bands {};

extern "C" int cvs_bands_create(cvs_bands** out, int device, int rank, int world, int root, int rows, int cols, int levels, unsigned mask,
                                int width, float spacing)
{
    if (!out) return fail(CVS_ERR_INVALID_ARG, "out is null");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world || root < 0 || root >= world) return fail(CVS_ERR_INVALID_ARG, "rank %d / root %d of %d", rank, root, world);
    if (rows <= 0 || cols <= 0 || levels < 1 || levels > 16) return fail(CVS_ERR_INVALID_ARG, "rows/cols/levels out of range");
    if (!mask || (mask >> CVS_G2_NPLANES)) return fail(CVS_ERR_INVALID_ARG, "mask 0x%x", mask);
    cvs_bands* b = new (std::nothrow) cvs_bands();
    if (!b) return fail(CVS_ERR_CUDA, "out of host memory");
    b->device = device, b->rank = rank, b->world = world, b->root = root;
    b->rows = rows, b->cols = cols, b->levels = levels, b->width = width, b->mask = mask;
    for (int p = 0; p < CVS_G2_NPLANES; ++p)
        if (mask >> p & 1u) b->sel[b->nsel++] = p;
    b->plans = plan_bands_c(rows, world, levels, width);
    b->lc.resize(levels);
    b->pitch.resize(levels);
    b->lc[0] = cols;
    for (int l = 1; l < levels; ++l) b->lc[l] = (b->lc[l - 1] + 1) / 2;
    for (int l = 0; l < levels; ++l) b->pitch[l] = align_up((size_t)b->lc[l] * 4, 128);
    int rc = filter_create(&b->f, 2, device, width, spacing);
    if (rc) {
        delete b;
        return rc;
    }
    auto bail = [&](int code) {
        cvs_bands_destroy(b);
        return code;
    };
    cudaError_t e = cudaSetDevice(device);
    b->lv.resize(levels);
    for (int l = 0; l < levels && e == cudaSuccess; ++l) {
        const int have = b->plan().have[l].second - b->plan().have[l].first;
        if (have > 0) e = b->lv[l].reserve(b->pitch[l] * (size_t)have);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->xfer, cudaStreamNonBlocking);
    b->ev.assign((size_t)levels * kMaxChunks, nullptr);
    for (size_t i = 0; i < b->ev.size() && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&b->ev[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_xfer, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_start, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->token), 256);
    if (e == cudaSuccess) e = cudaMemset(b->token, 0, 256);
    if (e != cudaSuccess) return bail(fail(CVS_ERR_CUDA, "cvs_bands_create: %s", cudaGetErrorString(e)));
    *out = b;
    return CVS_OK;
}

// First half of an orderly multi-process teardown, safe to call on every rank at the same time: waits for this rank's work,
// destroys the library-owned NCCL communicator (ncclCommDestroy may wait for the peers' calls, so every rank must get here
// without waiting on another rank first) and unmaps an imported root block.  The root's allocation stays valid until
// cvs_bands_destroy, which the root calls after its peers have detached.
extern "C" int cvs_bands_detach(cvs_bands* b)
{
    if (!b) return CVS_OK;
    CU_TRY(cudaSetDevice(b->device));
    CU_TRY(cudaDeviceSynchronize());
    if (b->comm && b->comm_owned && nccl().ok) nccl().CommDestroy(b->comm);
    b->comm = nullptr;
    b->comm_owned = false;
    if (b->root_block && b->root_mapped) {
        CU_TRY(cudaIpcCloseMemHandle(b->root_block));
        b->root_block = nullptr;
        b->root_mapped = false;
        b->rootp.clear();
    }
    return CVS_OK;
}

extern "C" int cvs_bands_destroy(cvs_bands* b)
{
    if (!b) return CVS_OK;
    cudaSetDevice(b->device);
    cudaDeviceSynchronize();
    if (b->comm && b->comm_owned && nccl().ok) nccl().CommDestroy(b->comm);
    if (b->root_block && b->root_mapped) cudaIpcCloseMemHandle(b->root_block);
    if (b->root_block && b->root_owned) cudaFree(b->root_block);
    for (cudaEvent_t e : b->ev)
        if (e) cudaEventDestroy(e);
    if (b->ev_xfer) cudaEventDestroy(b->ev_xfer);
    if (b->ev_start) cudaEventDestroy(b->ev_start);
    if (b->xfer) cudaStreamDestroy(b->xfer);
    if (b->token) cudaFree(b->token);
    b->lv.clear();
    b->local.release();
    filter_destroy(b->f);
    delete b;
    return CVS_OK;
}

extern "C" int cvs_bands_geometry(const cvs_bands* b, int rank, int level, int* level_rows, int* level_cols, int* out_lo, int* out_hi, int* have_lo,
                                  int* have_hi, size_t* pitch)
{
    if (!b || level < 0 || level >= b->levels || rank < -1 || rank >= b->world) return fail(CVS_ERR_INVALID_ARG, "cvs_bands_geometry: bad argument");
    const BandPlanC& p = b->plans[rank < 0 ? b->rank : rank];
    if (level_rows) *level_rows = p.rows[level];
    if (level_cols) *level_cols = b->lc[level];
    if (out_lo) *out_lo = p.out[level].first;
    if (out_hi) *out_hi = p.out[level].second;
    if (have_lo) *have_lo = p.have[level].first;
    if (have_hi) *have_hi = p.have[level].second;
    if (pitch) *pitch = b->pitch[level];
    return CVS_OK;
}

extern "C" int cvs_bands_input_dev(cvs_bands* b, float** ptr, size_t* pitch)
{
    if (!b || !ptr) return fail(CVS_ERR_INVALID_ARG, "null argument");
    *ptr = static_cast<float*>(b->lv[0].p);
    if (pitch) *pitch = b->pitch[0];
    return CVS_OK;
}

extern "C" int cvs_bands_upload_host(cvs_bands* b, const float* image, size_t step, void* stream)
{
    if (!b || !image) return fail(CVS_ERR_INVALID_ARG, "null argument");
    if (step < (size_t)b->cols * 4) return fail(CVS_ERR_INVALID_ARG, "step %zu < cols*4", step);
    if (b->plan().empty()) return CVS_OK;
    CU_TRY(cudaSetDevice(b->device));
    const int lo = b->plan().have[0].first, hi = b->plan().have[0].second;
    CU_TRY(cudaMemcpy2DAsync(b->lv[0].p, b->pitch[0], reinterpret_cast<const char*>(image) + (size_t)lo * step, step, (size_t)b->cols * 4, hi - lo,
                             cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return CVS_OK;
}

extern "C" int cvs_bands_root_bytes(const cvs_bands* b, size_t* bytes)
{
    if (!b || !bytes) return fail(CVS_ERR_INVALID_ARG, "null argument");
    *bytes = root_layout(*b, nullptr);
    return CVS_OK;
}

extern "C" int cvs_bands_root_export(cvs_bands* b, unsigned char handle[CVS_IPC_HANDLE_BYTES], void** base)
{
    if (!b) return fail(CVS_ERR_INVALID_ARG, "null argument");
    if (b->rank != b->root) return fail(CVS_ERR_INVALID_ARG, "cvs_bands_root_export is for the root rank (%d), this is rank %d", b->root, b->rank);
    CU_TRY(cudaSetDevice(b->device));
    if (!b->root_block) {
        CU_TRY(cudaMalloc(&b->root_block, root_layout(*b, nullptr)));
        b->root_owned = true;
        set_root_block(*b, b->root_block);
    }
    if (handle) {
        cudaIpcMemHandle_t h;
        CU_TRY(cudaIpcGetMemHandle(&h, b->root_block));
        memcpy(handle, &h, sizeof(h));
    }
    if (base) *base = b->root_block;
    return CVS_OK;
}

extern "C" int cvs_bands_root_import(cvs_bands* b, const unsigned char handle[CVS_IPC_HANDLE_BYTES])
{
    if (!b || !handle) return fail(CVS_ERR_INVALID_ARG, "null argument");
    if (b->root_block) return fail(CVS_ERR_INVALID_ARG, "root planes already attached");
    CU_TRY(cudaSetDevice(b->device));  // the IMPORTER's GPU: the mapping (and the NVLink peer path) is made for it
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    CU_TRY(cudaIpcOpenMemHandle(&b->root_block, h, cudaIpcMemLazyEnablePeerAccess));
    b->root_mapped = true;
    set_root_block(*b, b->root_block);
    return CVS_OK;
}

extern "C" int cvs_bands_root_attach(cvs_bands* b, void* base)
{
    if (!b || !base) return fail(CVS_ERR_INVALID_ARG, "null argument");
    if (b->root_block) return fail(CVS_ERR_INVALID_ARG, "root planes already attached");
    set_root_block(*b, base);  // same-process pointer (peer access enabled by the caller): not owned, not unmapped
    return CVS_OK;
}

extern "C" int cvs_bands_root_attach_planes(cvs_bands* b, float* const* const* planes, const size_t* pitches)
{
    if (!b || !planes || !pitches) return fail(CVS_ERR_INVALID_ARG, "null argument");
    b->rootp.assign((size_t)b->levels * b->nsel, PlaneRef{});
    for (int l = 0; l < b->levels; ++l) {
        if (!planes[l] || pitches[l] < (size_t)b->lc[l] * 4 || (pitches[l] & 3)) return fail(CVS_ERR_INVALID_ARG, "planes[%d] / pitches[%d]", l, l);
        for (int k = 0; k < b->nsel; ++k) {
            if (!planes[l][b->sel[k]]) return fail(CVS_ERR_INVALID_ARG, "planes[%d][%d] is null but selected by mask", l, b->sel[k]);
            b->rootp[(size_t)l * b->nsel + k] = PlaneRef{planes[l][b->sel[k]], pitches[l]};
        }
    }
    return CVS_OK;
}

extern "C" int cvs_bands_root_plane(cvs_bands* b, int level, int plane, float** ptr, size_t* pitch)
{
    if (!b || !ptr || level < 0 || level >= b->levels) return fail(CVS_ERR_INVALID_ARG, "bad argument");
    if (b->rootp.empty()) return fail(CVS_ERR_NOT_SETUP, "no root planes attached (cvs_bands_root_export / _import / _attach)");
    for (int k = 0; k < b->nsel; ++k)
        if (b->sel[k] == plane) {
            *ptr = b->rootp[(size_t)level * b->nsel + k].p;
            if (pitch) *pitch = b->rootp[(size_t)level * b->nsel + k].pitch;
            return CVS_OK;
        }
    return fail(CVS_ERR_INVALID_ARG, "plane %d is not selected by this context's mask", plane);
}

extern "C" int cvs_bands_local_plane(cvs_bands* b, int level, int plane, float** ptr, size_t* pitch)
{
    if (!b || !ptr || level < 0 || level >= b->levels) return fail(CVS_ERR_INVALID_ARG, "bad argument");
    CU_TRY(cudaSetDevice(b->device));
    int rc = ensure_local(*b);
    if (rc) return rc;
    for (int k = 0; k < b->nsel; ++k)
        if (b->sel[k] == plane) {
            *ptr = reinterpret_cast<float*>(static_cast<char*>(b->local.p) + b->local_off[(size_t)level * b->nsel + k]);
            if (pitch) *pitch = b->pitch[level];
            return CVS_OK;
        }
    return fail(CVS_ERR_INVALID_ARG, "plane %d is not selected by this context's mask", plane);
}

extern "C" int cvs_nccl_unique_id(unsigned char id[CVS_NCCL_ID_BYTES])
{
    if (!id) return fail(CVS_ERR_INVALID_ARG, "id is null");
    if (!nccl().ok) return fail(CVS_ERR_UNSUPPORTED, "NCCL not available: %s", dlerror() ? dlerror() : "libnccl.so.2 not found");
    NcclUid u;
    NCCL_TRY(nccl().GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return CVS_OK;
}

extern "C" int cvs_bands_nccl_init(cvs_bands* b, const unsigned char id[CVS_NCCL_ID_BYTES])
{
    if (!b || !id) return fail(CVS_ERR_INVALID_ARG, "null argument");
    if (!nccl().ok) return fail(CVS_ERR_UNSUPPORTED, "NCCL not available");
    if (b->comm) return fail(CVS_ERR_INVALID_ARG, "communicator already set");
    CU_TRY(cudaSetDevice(b->device));
    NcclUid u;
    memcpy(&u, id, sizeof(u));
    NCCL_TRY(nccl().CommInitRank(&b->comm, b->world, u, b->rank));
    b->comm_owned = true;
    return CVS_OK;
}

extern "C" int cvs_bands_nccl_attach(cvs_bands* b, void* comm)
{
    if (!b || !comm) return fail(CVS_ERR_INVALID_ARG, "null argument");
    if (!nccl().ok) return fail(CVS_ERR_UNSUPPORTED, "NCCL not available");
    if (b->comm) return fail(CVS_ERR_INVALID_ARG, "communicator already set");
    b->comm = comm;  // caller-owned ncclComm_t whose ranks are the band ranks
    return CVS_OK;
}

// One step: every level of this rank's band.  Asynchronous on `stream`; when `stream` reaches the end of this call's work,
// this rank's kernels, copies and sends (root: receives) have completed.
extern "C" int cvs_bands_run(cvs_bands* b, int gather, void* stream_)
{
    if (!b) return fail(CVS_ERR_INVALID_ARG, "context is null");
    if (gather < CVS_GATHER_NONE || gather > CVS_GATHER_PEER_COPY) return fail(CVS_ERR_INVALID_ARG, "gather mode %d", gather);
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    CU_TRY(cudaSetDevice(b->device));
    const bool is_root = b->rank == b->root;
    const bool to_root = gather != CVS_GATHER_NONE && b->world > 1;
    if (to_root && b->rootp.empty() && (is_root || gather != CVS_GATHER_NCCL))
        return fail(CVS_ERR_NOT_SETUP, "gather mode %d needs the root's planes (cvs_bands_root_export / _import / _attach)", gather);
    if (gather == CVS_GATHER_NCCL && b->world > 1 && !b->comm) return fail(CVS_ERR_NOT_SETUP, "CVS_GATHER_NCCL needs cvs_bands_nccl_init / _attach");
    // where this rank's kernels write: straight into the root's planes (root itself always when it has them; everyone in
    // PEER_STORE mode), else into local planes
    const bool direct = !b->rootp.empty() && (is_root || (to_root && gather == CVS_GATHER_PEER_STORE));
    if (!direct) {
        int rc = ensure_local(*b);
        if (rc) return rc;
    }
    const BandPlanC& plan = b->plan();
    const bool xfer = to_root && (gather == CVS_GATHER_NCCL || (gather == CVS_GATHER_PEER_COPY && !is_root));
    if (xfer) {  // transfers of this step must not start before earlier work on `s` (e.g. the input upload) is done
        CU_TRY(cudaEventRecord(b->ev_start, s));
        CU_TRY(cudaStreamWaitEvent(b->xfer, b->ev_start, 0));
    }
    SteerSpec st{};
    st.source = CVS_STEER_DOMINANT;
    for (int l = 0; l < b->levels; ++l) {
        const int lo = plan.out[l].first, hi = plan.out[l].second;
        const int have_lo = plan.have[l].first, have_hi = plan.have[l].second;
        // next level first: it only needs this level's input, and the (larger) fused kernel can then overlap nothing less
        if (l + 1 < b->levels && plan.have[l + 1].first < plan.have[l + 1].second && have_lo < have_hi) {
            BatchGeom g = whole_frame_geom(b->lv[l].p, false, 1, have_hi - have_lo, b->lc[l], b->pitch[l], 0, b->pitch[l + 1], 0);
            g.full_rows = plan.rows[l];
            g.y_origin = have_lo;
            g.out_row_begin = plan.have[l + 1].first, g.out_row_end = plan.have[l + 1].second, g.out_row_origin = plan.have[l + 1].first;
            CU_TRY(launch_pyr_down(g, static_cast<float*>(b->lv[l + 1].p), s));
        }
        // chunks of this level's output rows (one chunk unless the copy engine gathers)
        int nch = 1, chunk_rows = hi - lo;
        if (xfer && gather == CVS_GATHER_PEER_COPY && hi - lo >= 2 * kMinChunkRows) {
            chunk_rows = std::max<int>(kMinChunkRows, ((hi - lo + kMaxChunks - 1) / kMaxChunks + 63) / 64 * 64);
            nch = (hi - lo + chunk_rows - 1) / chunk_rows;
        }
        for (int c = 0; c < nch && lo < hi; ++c) {
            const int clo = lo + c * chunk_rows, chi = std::min(hi, clo + chunk_rows);
            float* outs[CVS_G2_NPLANES] = {nullptr};
            size_t opitch = b->pitch[l];
            for (int k = 0; k < b->nsel; ++k) {
                if (direct) {
                    const PlaneRef& r = b->rootp[(size_t)l * b->nsel + k];
                    outs[b->sel[k]] = reinterpret_cast<float*>(reinterpret_cast<char*>(r.p) + (size_t)clo * r.pitch);
                    opitch = r.pitch;
                } else {
                    outs[b->sel[k]] = reinterpret_cast<float*>(static_cast<char*>(b->local.p) + b->local_off[(size_t)l * b->nsel + k] +
                                                               (size_t)(clo - lo) * b->pitch[l]);
                }
            }
            BatchGeom g = whole_frame_geom(b->lv[l].p, false, 1, have_hi - have_lo, b->lc[l], b->pitch[l], 0, opitch, 0);
            g.full_rows = plan.rows[l];
            g.y_origin = have_lo;
            g.out_row_begin = clo, g.out_row_end = chi, g.out_row_origin = clo;
            int rc = run_fused(b->f, g, b->mask, st, outs, s);
            if (rc) return rc;
            if (!(xfer && gather == CVS_GATHER_PEER_COPY)) continue;
            // this chunk's rows travel on the transfer stream (copy engine) while the next chunk / level computes on `s`
            cudaEvent_t ev = b->ev[(size_t)l * kMaxChunks + c];
            CU_TRY(cudaEventRecord(ev, s));
            CU_TRY(cudaStreamWaitEvent(b->xfer, ev, 0));
            for (int k = 0; k < b->nsel; ++k) {
                const PlaneRef& r = b->rootp[(size_t)l * b->nsel + k];
                const char* src = static_cast<const char*>(b->local.p) + b->local_off[(size_t)l * b->nsel + k] + (size_t)(clo - lo) * b->pitch[l];
                char* dst = reinterpret_cast<char*>(r.p) + (size_t)clo * r.pitch;
                if (r.pitch == b->pitch[l])
                    CU_TRY(cudaMemcpyAsync(dst, src, b->pitch[l] * (size_t)(chi - clo), cudaMemcpyDefault, b->xfer));
                else
                    CU_TRY(cudaMemcpy2DAsync(dst, r.pitch, src, b->pitch[l], (size_t)b->lc[l] * 4, chi - clo, cudaMemcpyDefault, b->xfer));
            }
        }
        if (!xfer || gather == CVS_GATHER_PEER_COPY) continue;
        // NCCL: this level's rows travel on the transfer stream while the next level computes on `s`
        CU_TRY(cudaEventRecord(b->ev[(size_t)l * kMaxChunks], s));
        CU_TRY(cudaStreamWaitEvent(b->xfer, b->ev[(size_t)l * kMaxChunks], 0));
        {
            // one group per level; row blocks are contiguous (same pitch on both sides), so no staging copy
            NCCL_TRY(nccl().GroupStart());
            int nerr = 0;
            if (is_root) {
                for (int r = 0; r < b->world; ++r) {
                    if (r == b->root || b->out_rows(l, r) <= 0) continue;
                    for (int k = 0; k < b->nsel; ++k) {
                        const PlaneRef& pr = b->rootp[(size_t)l * b->nsel + k];
                        if (pr.pitch != b->pitch[l]) return fail(CVS_ERR_UNSUPPORTED, "CVS_GATHER_NCCL needs root planes with the context's own pitch");
                        nerr |= nccl().Recv(reinterpret_cast<char*>(pr.p) + (size_t)b->plans[r].out[l].first * pr.pitch,
                                            b->pitch[l] / 4 * (size_t)b->out_rows(l, r), kNcclFloat, r, b->comm, b->xfer);
                    }
                }
            } else if (lo < hi) {
                for (int k = 0; k < b->nsel; ++k)
                    nerr |= nccl().Send(static_cast<const char*>(b->local.p) + b->local_off[(size_t)l * b->nsel + k], b->pitch[l] / 4 * (size_t)(hi - lo),
                                        kNcclFloat, b->root, b->comm, b->xfer);
            }
            NCCL_TRY(nccl().GroupEnd());
            if (nerr) return fail(CVS_ERR_CUDA, "ncclSend/ncclRecv failed");
        }
    }
    if (xfer) {  // join: `s` continues only after this rank's transfers are done
        CU_TRY(cudaEventRecord(b->ev_xfer, b->xfer));
        CU_TRY(cudaStreamWaitEvent(s, b->ev_xfer, 0));
    }
    return CVS_OK;
}

// Cross-rank barrier ON the stream (a 1-element ncclAllReduce): after it, every rank's work enqueued before its own
// barrier call has completed -- what the root needs before reading planes filled by peer stores / peer copies.
extern "C" int cvs_bands_barrier(cvs_bands* b, void* stream)
{
    if (!b) return fail(CVS_ERR_INVALID_ARG, "context is null");
    if (b->world == 1) return CVS_OK;
    if (!b->comm) return fail(CVS_ERR_NOT_SETUP, "cvs_bands_barrier needs cvs_bands_nccl_init / _attach (or use the caller's own barrier)");
    CU_TRY(cudaSetDevice(b->device));
    NCCL_TRY(nccl().AllReduce(b->token, b->token, 1, kNcclFloat, kNcclSum, b->comm, static_cast<cudaStream_t>(stream)));
    return CVS_OK;
}

// ---- one process, several GPUs: device-resident outputs on devices[0] -------------------------------------------------
extern "C" int cvs_g2_run_bands_dev_multi(int n_devices, const int* devices, int width, float spacing, const float* image, int rows, int cols,
                                          size_t step, int levels, unsigned mask, int gather, float* const* const* root_planes,
                                          const size_t* root_pitches)
{
    if (n_devices <= 0 || n_devices > 64 || !image || !root_planes || !root_pitches) return fail(CVS_ERR_INVALID_ARG, "bad argument");
    if (gather != CVS_GATHER_PEER_STORE && gather != CVS_GATHER_PEER_COPY)
        return fail(CVS_ERR_INVALID_ARG, "one-process band runs gather by peer stores (2) or peer copies (3)");
    auto dev = [&](int i) { return devices ? devices[i] : i; };
    for (int i = 1; i < n_devices; ++i) {
        int rc = cvs_enable_peer_access(dev(i), dev(0));
        if (rc) return rc;
    }
    std::vector<cvs_bands*> ctx(n_devices, nullptr);
    std::vector<int> rc(n_devices, CVS_OK);
    std::vector<std::string> msg(n_devices);
    std::vector<std::thread> threads;
    for (int i = 0; i < n_devices; ++i)
        threads.emplace_back([&, i] {
            int r = cvs_bands_create(&ctx[i], dev(i), i, n_devices, 0, rows, cols, levels, mask, width, spacing);
            if (r == CVS_OK) r = cvs_bands_root_attach_planes(ctx[i], root_planes, root_pitches);
            Filter* f = r == CVS_OK ? ctx[i]->f : nullptr;
            if (r == CVS_OK) r = cvs_bands_upload_host(ctx[i], image, step, f->stream);
            if (r == CVS_OK) r = cvs_bands_run(ctx[i], gather, f->stream);
            if (r == CVS_OK && cudaStreamSynchronize(f->stream) != cudaSuccess) r = fail(CVS_ERR_CUDA, "band sync: %s", cudaGetErrorString(cudaGetLastError()));
            if (r != CVS_OK) msg[i] = cvs_last_error();
            rc[i] = r;
        });
    for (auto& t : threads) t.join();  // every GPU's stores / copies have landed in devices[0]'s planes
    for (auto* c : ctx) cvs_bands_destroy(c);
    for (int i = 0; i < n_devices; ++i)
        if (rc[i] != CVS_OK) return fail(rc[i], "device %d: %s", dev(i), msg[i].c_str());
    return CVS_OK;
}
