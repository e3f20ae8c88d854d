// Development probe: which k_march instantiations run?  (not part of the library)
#include <unistd.h>
#include <sys/wait.h>
#include "../cvsteer_b200/csrc/march_launch.cuh"
namespace cvs { std::atomic<unsigned long long> g_launches{0}; }
using namespace cvs;
struct G4FamBH32 : G4Fam { static constexpr int BH = 32; };
struct G4FamC1 : G4Fam { static constexpr int MIN_CTAS = 1; };
template <class Fam, unsigned MASK> int run(const char* nm, unsigned mask, int family)
{
    const int W = 256, H = 185, n = 1;
    float* d; cudaMalloc(&d, W * H * 4); cudaMemset(d, 0, W * H * 4);
    float* o; cudaMalloc(&o, (size_t)W * H * 4 * 16);
    BatchGeom g{}; g.in = d; g.n = n; g.cols = W; g.buf_rows = H; g.full_rows = H; g.out_row_end = H; g.in_pitch = W * 4; g.in_frame_stride = W * H * 4; g.out_pitch = W * 4; g.out_frame_stride = W * H * 4;
    float* outs[20]; for (int i = 0; i < 20; ++i) outs[i] = o + (size_t)(i % 16) * W * H;
    SteerSpec st{}; st.source = CVS_STEER_SCALAR; st.cos_t = 1; st.sin_t = 0;
    MarchArgs a = make_args(g, mask, st, outs, 15);
    FamilyTaps ft{}; ft.width = Fam::R; ft.nsets = family == 2 ? 7 : 11;
    for (int s = 0; s < ft.nsets; ++s) for (int i = 0; i < 2 * Fam::R + 1; ++i) ft.t[s][i] = 0.1f;
    TapTable<Fam::NSETS, Fam::R> tt; fill_tap_table<Fam>(ft, tt);
    dim3 grid((W + 127) / 128, (H + Fam::BH - 1) / Fam::BH, n);
    LaunchInfo li;
    cudaError_t e = launch_march_mask<Fam, MASK>(g, a, tt, grid, 0, &li, nm);
    cudaError_t e2 = cudaDeviceSynchronize();
    printf("%-28s %-24s launch=%s sync=%s\n", nm, li.name, cudaGetErrorString(e), cudaGetErrorString(e2));
    return 0;
}
template <class F> void child(F f) { fflush(stdout); pid_t c = fork(); if (!c) { f(); fflush(stdout); _exit(0); } int st; waitpid(c, &st, 0); }
int main()
{
    child([] { run<G2Fam, 0u>("g2 dyn", 0xFFFFF, 2); });
    child([] { run<G4Fam, 0u>("g4 dyn basis", CVS_G4_MASK_BASIS, 4); });
    child([] { run<G4Fam, 0u>("g4 dyn all", 0x7FFF, 4); });
    child([] { run<G4Fam, CVS_G4_MASK_BASIS>("g4 static basis", CVS_G4_MASK_BASIS, 4); });
    child([] { run<G4FamBH32, 0u>("g4 BH32 dyn", CVS_G4_MASK_BASIS, 4); });
    child([] { run<G4FamC1, 0u>("g4 minctas1 dyn", CVS_G4_MASK_BASIS, 4); });
    return 0;
}
