// pfbch2_fused.cu -- fused firpfbch2_crcf analysis kernel for M = 1024 (sm_100a).
//
// One persistent CTA per SM walks a contiguous range of frames.  Warp-specialised, producer /
// consumer through shared memory with mbarriers (no CTA-wide barrier in the steady state):
//
//   16 MAC warps (512 threads)   thread j owns the window pair (j, j+512): two 2m-sample complex
//                                histories and the two 2m-tap branches live in registers (6*2m
//                                registers).  Per frame: one new sample (its own, read once from
//                                HBM through a per-warp cp.async.bulk ring), two dot products
//                                (8m FFMA), two 8-byte stores into the frame's slot of the X ring.
//                                The 2m-slot circular histories are addressed statically by fully
//                                unrolling one period (4m frames).
//    4 FFT warps (128 threads)   warp f transforms frame 4g+f of group g: 1024 = 32 x 32, thread =
//                                32 points in registers, radix-2 DIT butterflies with constant
//                                twiddles, one in-place transposition through the slot (padded
//                                stride 33), inter-stage twiddles from a shared 32x32 table, then
//                                32 coalesced 256-byte row stores per frame straight to HBM.
//
// FP32 pipe and shared-memory pipe work of the two roles overlap by construction (the warp
// scheduler interleaves them); per input sample the kernel reads 8 B and writes 16 B of HBM.
// The 1/M output scale is folded into the taps (exact: M is a power of two).
//
// Reference semantics: src/multichannel/src/firpfbch2.c:244-282 (see firpfbch2.cu).
#include "pfbch2_fused.cuh"

#include <cmath>
#include <cstdint>

namespace {

constexpr int kM = 1024;
constexpr int kM2 = 512;
constexpr int kMacThreads = 512;
constexpr int kMacWarps = 16;
constexpr int kFftWarps = 4;
constexpr int kThreads = kMacThreads + 32 * kFftWarps;        // 640
constexpr int kGroup = 4;                                      // frames per group (one per FFT warp)
constexpr int kRingGroups = 4;                                 // X ring depth, in groups (power of two)
constexpr int kSlotStride = 32 * 33;                           // float2 per frame slot (padded transpose)
constexpr int kInStages = 4;                                   // per-warp input ring depth, in groups (= kRingGroups)
constexpr int kInStageBytes = kGroup * 32 * 8;                 // 4 frames x 32 samples x 8 B per warp

// shared memory carve-up (bytes)
constexpr int kOffTw = 0;                                      // float2 tw[32][32]
constexpr int kOffX = kOffTw + 32 * 32 * 8;                    // X ring
constexpr int kOffIn = kOffX + kRingGroups * kGroup * kSlotStride * 8;
constexpr int kOffBar = kOffIn + kMacWarps * kInStages * kInStageBytes;
constexpr int kNumBars = 2 * kRingGroups + kMacWarps * kInStages;
constexpr int kSmemBytes = kOffBar + kNumBars * 8;

// ---------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA engine)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, float2 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// ---------------------------------------------------------------------------------- register DFT
// cos / sin of 2 pi k / 32, k = 0..15
__device__ constexpr float kCos32[16] = {1.0f,
                                         0.98078528040323044913f,
                                         0.92387953251128675613f,
                                         0.83146961230254523708f,
                                         0.70710678118654752440f,
                                         0.55557023301960222474f,
                                         0.38268343236508977173f,
                                         0.19509032201612826785f,
                                         0.0f,
                                         -0.19509032201612826785f,
                                         -0.38268343236508977173f,
                                         -0.55557023301960222474f,
                                         -0.70710678118654752440f,
                                         -0.83146961230254523708f,
                                         -0.92387953251128675613f,
                                         -0.98078528040323044913f};
__device__ constexpr float kSin32[16] = {0.0f,
                                         0.19509032201612826785f,
                                         0.38268343236508977173f,
                                         0.55557023301960222474f,
                                         0.70710678118654752440f,
                                         0.83146961230254523708f,
                                         0.92387953251128675613f,
                                         0.98078528040323044913f,
                                         1.0f,
                                         0.98078528040323044913f,
                                         0.92387953251128675613f,
                                         0.83146961230254523708f,
                                         0.70710678118654752440f,
                                         0.55557023301960222474f,
                                         0.38268343236508977173f,
                                         0.19509032201612826785f};

// N-point DFT of v[] in registers, natural order in and out, e^{SIGN j 2 pi nk/N}.
// Decimation in time; every index is a compile-time constant, so the even/odd split is renaming.
template <int N, int SIGN>
struct RegDft {
    static __device__ __forceinline__ void run(float2 (&v)[N])
    {
        float2 e[N / 2], o[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; i++) {
            e[i] = v[2 * i];
            o[i] = v[2 * i + 1];
        }
        RegDft<N / 2, SIGN>::run(e);
        RegDft<N / 2, SIGN>::run(o);
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            constexpr float s = (float)SIGN;
            if (k == 0) {
                v[k] = make_float2(e[k].x + o[k].x, e[k].y + o[k].y);
                v[k + N / 2] = make_float2(e[k].x - o[k].x, e[k].y - o[k].y);
            } else if (4 * k == N) {
                // w = SIGN * j
                v[k] = make_float2(e[k].x - s * o[k].y, e[k].y + s * o[k].x);
                v[k + N / 2] = make_float2(e[k].x + s * o[k].y, e[k].y - s * o[k].x);
            } else {
                const float c = kCos32[k * (32 / N)];
                const float sn = s * kSin32[k * (32 / N)];
                // a + w b with two FMAs per component; a - w b = 2a - (a + w b)
                float re = fmaf(c, o[k].x, e[k].x);
                re = fmaf(-sn, o[k].y, re);
                float im = fmaf(c, o[k].y, e[k].y);
                im = fmaf(sn, o[k].x, im);
                v[k] = make_float2(re, im);
                v[k + N / 2] = make_float2(fmaf(2.0f, e[k].x, -re), fmaf(2.0f, e[k].y, -im));
            }
        }
    }
};
template <int SIGN>
struct RegDft<1, SIGN> {
    static __device__ __forceinline__ void run(float2 (&)[1]) {}
};

// ---------------------------------------------------------------------------------- the kernel
struct FusedArgs {
    const float2 *hist;        // H samples preceding x[0] (never null)
    const float2 *x;           // segment input
    float2 *y;                 // [n_frames][1024]
    const float *taps;         // [2][L][1024] as in firpfbch2.cu, pre-scaled by 1/M
    const float2 *tw;          // [32][32] inter-stage twiddles e^{+j 2 pi k1 n2 / 1024}
    int H;
    int par0;                  // parity of the segment's first frame
    int n_groups;              // groups of 4 frames in the segment
    int groups_per_cta;
};

template <int L>
__global__ void __launch_bounds__(kThreads, 1) pfbch2_fused_1024_kernel(const FusedArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_tw = s_base + kOffTw;
    const uint32_t s_x = s_base + kOffX;
    const uint32_t s_in = s_base + kOffIn;
    const uint32_t s_bar = s_base + kOffBar;
    // barriers: full[g] at s_bar + 8g, empty[g] at s_bar + 8(R+g), in[w][s] after those
    auto bar_full = [&](int g) { return s_bar + 8u * g; };
    auto bar_empty = [&](int g) { return s_bar + 8u * (kRingGroups + g); };
    auto bar_in = [&](int w, int s) { return s_bar + 8u * (2 * kRingGroups + w * kInStages + s); };

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const int g_begin = (int)blockIdx.x * a.groups_per_cta;
    int g_end = g_begin + a.groups_per_cta;
    if (g_end > a.n_groups) g_end = a.n_groups;
    const int n_groups = g_end - g_begin;                  // > 0 by construction of the grid
    const long long f_begin = (long long)g_begin * kGroup; // first frame of this CTA (even)

    // ---- one-time setup
    {
        float2 *tw_s = reinterpret_cast<float2 *>(smem + kOffTw);
        for (int i = tid; i < 32 * 32; i += kThreads) tw_s[i] = a.tw[i];
    }
    if (tid == 0) {
        for (int g = 0; g < kRingGroups; g++) {
            mbar_init(bar_full(g), kMacWarps);
            mbar_init(bar_empty(g), kFftWarps);
        }
        for (int w = 0; w < kMacWarps; w++)
            for (int s = 0; s < kInStages; s++) mbar_init(bar_in(w, s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < kMacWarps) {
        // =========================================================================== MAC warps
        // (no setmaxnreg: registers can only move between warpgroups of one CTA, and 640 x 96 is the launch allocation)
        const int j = tid;                                  // window pair (j, j + 512)
        // byte offset, inside a frame slot, of this thread's first dot product (window j or j+512
        // depending on the parity of the segment); the second one sits 512 windows away
        const uint32_t w1off = (uint32_t)(j + (a.par0 ? kM2 : 0)) * 8;
        const uint32_t w2delta = a.par0 ? (uint32_t)(-kM2 * 8) : (uint32_t)(kM2 * 8);   // warp-uniform

        float hA[L], hB[L];                                 // branch j, branch j + 512 (tap n = n-th newest)
        float2 dA[L], dB[L];                                // circular histories of the two windows
#pragma unroll
        for (int n = 0; n < L; n++) {
            hA[n] = __ldg(a.taps + (size_t)n * kM + j);                 // parity-0 table, window j -> branch j
            hB[n] = __ldg(a.taps + (size_t)n * kM + j + kM2);           // window j+512 -> branch j+512
        }
        // history: this thread's sample of frame F is stream index F*512 + 511 - j
        {
            const long long base = f_begin * kM2 + (kM2 - 1 - j);
#pragma unroll
            for (int n = 1; n < L; n++)                     // window 1: frames f_begin - 2n
                dA[L - n] = lcd::ld_stream(a.hist, a.x, base - 2LL * n * kM2, a.H);
            dA[0] = make_float2(0.f, 0.f);
#pragma unroll
            for (int n = 0; n < L; n++)                     // window 2: frames f_begin - 1 - 2n
                dB[L - 1 - n] = lcd::ld_stream(a.hist, a.x, base - (1LL + 2 * n) * kM2, a.H);
        }

        // per-warp input ring: stage s holds this warp's 32 samples of the 4 frames of one group
        const uint32_t in_warp = s_in + (uint32_t)warp * (kInStages * kInStageBytes);
        auto issue_group = [&](int g) {                     // lane 0 only
            const int s = g & (kInStages - 1);
            const uint32_t bar = bar_in(warp, s);
            // frame (f_begin + 4g + f), this warp: 32 samples starting at 512 - 32 (warp + 1)
            const float2 *src = a.x + (f_begin + (long long)g * kGroup) * kM2 + (kM2 - 32 * (warp + 1));
            mbar_arrive_expect_tx(bar, kInStageBytes);
#pragma unroll
            for (int f = 0; f < kGroup; f++)
                bulk_g2s(in_warp + s * kInStageBytes + f * 256, src + f * kM2, 256, bar);
        };
        if (lane == 0) {
            for (int g = 0; g < kInStages - 1 && g < n_groups; g++) issue_group(g);
        }
        const uint32_t in_lane = in_warp + (uint32_t)(31 - lane) * 8;
        const uint32_t bar_in_w = bar_in(warp, 0);

        int g = 0;
        while (g < n_groups) {
            // one period: L updates of each window = 2L frames = L/2 groups, fully unrolled
#pragma unroll
            for (int uu = 0; uu < L; uu++) {
                const int fg = (uu & 1) * 2;                // frame-in-group of this double frame's even frame
                const int gq = g & (kRingGroups - 1);       // stage of both rings (same depth)
                if ((uu & 1) == 0) {
                    if (g >= n_groups) break;
                    // keep the input ring full: fetch group g + kInStages - 1 into the stage read last group
                    __syncwarp();
                    if (lane == 0 && g + kInStages - 1 < n_groups) {
                        fence_proxy_async();
                        issue_group(g + kInStages - 1);
                    }
                    const uint32_t par = (uint32_t)(g >> 2) & 1u;
                    mbar_wait(bar_in_w + 8u * gq, par);
                    mbar_wait(bar_empty(gq), par ^ 1u);
                }
                const uint32_t in_g = in_lane + (uint32_t)gq * kInStageBytes;
                const uint32_t x_g = s_x + (uint32_t)gq * (kGroup * kSlotStride * 8) + w1off;

                // ---- even frame of the pair: window 1 gets the new sample
                dA[uu] = lds64(in_g + fg * 256);
                {
                    float2 o1 = make_float2(0.f, 0.f), o2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int n = L - 1; n >= 0; n--) {
                        const float2 s1 = dA[(uu - n + L) % L];
                        const float2 s2 = dB[(uu - 1 - n + 2 * L) % L];
                        o1.x = fmaf(hA[n], s1.x, o1.x);
                        o1.y = fmaf(hA[n], s1.y, o1.y);
                        o2.x = fmaf(hB[n], s2.x, o2.x);
                        o2.y = fmaf(hB[n], s2.y, o2.y);
                    }
                    const uint32_t slot = x_g + (uint32_t)fg * (kSlotStride * 8);
                    sts64(slot, o1);
                    sts64(slot + w2delta, o2);
                }
                // ---- odd frame of the pair: window 2 gets the new sample, branches swap
                dB[uu] = lds64(in_g + (fg + 1) * 256);
                {
                    float2 o1 = make_float2(0.f, 0.f), o2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int n = L - 1; n >= 0; n--) {
                        const float2 s1 = dA[(uu - n + L) % L];
                        const float2 s2 = dB[(uu - n + L) % L];
                        o1.x = fmaf(hB[n], s1.x, o1.x);
                        o1.y = fmaf(hB[n], s1.y, o1.y);
                        o2.x = fmaf(hA[n], s2.x, o2.x);
                        o2.y = fmaf(hA[n], s2.y, o2.y);
                    }
                    const uint32_t slot = x_g + (uint32_t)(fg + 1) * (kSlotStride * 8);
                    sts64(slot, o1);
                    sts64(slot + w2delta, o2);
                }
                if (uu & 1) {
                    // group complete: publish it to the FFT warps
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full(gq));
                    g++;
                }
            }
        }
    } else {
        // =========================================================================== FFT warps
        const int fw = warp - kMacWarps;                    // frame-in-group handled by this warp
        const uint32_t tw_lane = s_tw + (uint32_t)lane * 8; // tw[k1][lane] at + k1*256
        for (int g = 0; g < n_groups; g++) {
            const int gs = g & (kRingGroups - 1);
            mbar_wait(bar_full(gs), (uint32_t)(g >> 2) & 1u);
            const uint32_t slot = s_x + (uint32_t)(gs * kGroup + fw) * (kSlotStride * 8);

            float2 v[32];
            // pass 1: lane = n2, points X[32 n1 + n2]; 32-point DFT over n1; twiddle W^(n2 k1)
#pragma unroll
            for (int n1 = 0; n1 < 32; n1++) v[n1] = lds64(slot + (uint32_t)(32 * n1 + lane) * 8);
            RegDft<32, +1>::run(v);
#pragma unroll
            for (int k1 = 1; k1 < 32; k1++) v[k1] = lcd::cmul(v[k1], lds64(tw_lane + k1 * 256));
            __syncwarp();                                   // every lane has read its column
#pragma unroll
            for (int k1 = 0; k1 < 32; k1++) sts64(slot + (uint32_t)(k1 * 33 + lane) * 8, v[k1]);
            __syncwarp();
            // pass 2: lane = k1, points A[k1][n2]; 32-point DFT over n2 -> Y[k1 + 32 k2]
#pragma unroll
            for (int n2 = 0; n2 < 32; n2++) v[n2] = lds64(slot + (uint32_t)(lane * 33 + n2) * 8);
            RegDft<32, +1>::run(v);
            // the slot has been consumed (the DFT above needed every loaded value)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty(gs));
            float2 *yf = a.y + ((f_begin + (long long)g * kGroup + fw) * (long long)kM) + lane;
#pragma unroll
            for (int k2 = 0; k2 < 32; k2++) __stcs(yf + 32 * k2, v[k2]);
        }
    }
}

template <int L>
void launch_L(const FusedArgs &a, unsigned int grid, cudaStream_t st)
{
    static bool configured = false;
    if (!configured) {
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_fused_1024_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        configured = true;
    }
    pfbch2_fused_1024_kernel<L><<<grid, kThreads, kSmemBytes, st>>>(a);
    LC_LAUNCH_CHECK();
}

}  // namespace

namespace lc {
namespace pfb {

bool fused_pfbch2_supported(unsigned int M, unsigned int L)
{
    return M == 1024 && (L == 4 || L == 8 || L == 10 || L == 14);
}

void fused_pfbch2_make_twiddles(std::vector<float2> &tw)
{
    tw.resize(32 * 32);
    for (int k1 = 0; k1 < 32; k1++)
        for (int n2 = 0; n2 < 32; n2++) {
            double ang = 2.0 * M_PI * (double)(k1 * n2) / 1024.0;
            tw[k1 * 32 + n2] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

long long launch_pfbch2_fused(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps_scaled,
                              const float2 *d_tw, float2 *d_y, unsigned int M, unsigned int L,
                              unsigned long long n_frames, int parity0, cudaStream_t st)
{
    (void)M;
    long long n_groups = (long long)(n_frames / kGroup);
    if (n_groups == 0) return 0;
    if (n_groups > (1ll << 30)) n_groups = 1ll << 30;      // the caller loops for the rest
    int dev = 0, sms = 148;
    LC_CUDA_CHECK(cudaGetDevice(&dev));
    LC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // at least 8 groups per CTA so the history prologue and the pipeline fill amortise
    long long ctas = (n_groups + 7) / 8;
    if (ctas > sms) ctas = sms;
    long long per = (n_groups + ctas - 1) / ctas;
    ctas = (n_groups + per - 1) / per;
    FusedArgs a;
    a.hist = d_hist;
    a.x = d_x;
    a.y = d_y;
    a.taps = d_taps_scaled;
    a.tw = d_tw;
    a.H = H;
    a.par0 = parity0 & 1;
    a.n_groups = (int)n_groups;
    a.groups_per_cta = (int)per;
    switch (L) {
    case 4: launch_L<4>(a, (unsigned int)ctas, st); break;
    case 8: launch_L<8>(a, (unsigned int)ctas, st); break;
    case 10: launch_L<10>(a, (unsigned int)ctas, st); break;
    case 14: launch_L<14>(a, (unsigned int)ctas, st); break;
    default: fatal("launch_pfbch2_fused", "unsupported tap count");
    }
    return n_groups * kGroup;
}

}  // namespace pfb
}  // namespace lc
