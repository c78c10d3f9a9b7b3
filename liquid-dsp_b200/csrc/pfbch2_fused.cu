// pfbch2_fused.cu -- fused firpfbch2_crcf analysis kernel for M = 1024 (sm_100a).
//
// One persistent CTA per SM walks a contiguous range of frames.  Warp-specialised, producer /
// consumer through shared memory with mbarriers (no CTA-wide barrier in the steady state):
//
//   input                        one lane of the first FFT warp streams the input with cp.async.bulk
//                                (TMA engine): one 16 KB copy per group of 4 frames, four groups
//                                ahead, into a 4-stage ring; completion is counted in bytes on an
//                                mbarrier.  A stage is refilled once every MAC warp has published the
//                                group that read it, so no second barrier is needed.
//   16 MAC warps (512 threads)   thread j owns the window pair (j, j+512): two 2m-sample complex
//                                histories and the two 2m-tap branches live in registers (6*2m
//                                registers).  Per frame: one new sample (its own 8 bytes of the
//                                staged frame), two dot products (4m packed FFMA2: tap broadcast x
//                                {re, im} pair), two 8-byte stores into
//                                the frame's slot of the X ring.  The 2m-slot circular histories are
//                                addressed statically by fully unrolling one period (4m frames).
//    4 FFT warps (128 threads)   warp f transforms frame 4g+f of group g: 1024 = 32 x 32, thread =
//                                32 points in registers, radix-2 DIT butterflies with constant
//                                twiddles in packed FP32x2 arithmetic (3 instructions per general
//                                butterfly), one in-place transposition through the slot (padded
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
constexpr int kInStages = 4;                                   // input ring depth, in groups (= kRingGroups)
constexpr int kInStageBytes = kGroup * kM2 * 8;                // 4 frames x 512 samples x 8 B = 16 KB

// shared memory carve-up (bytes)
constexpr int kOffTw = 0;                                      // float2 tw[32][32]
constexpr int kOffX = kOffTw + 32 * 32 * 8;                    // X ring
constexpr int kOffIn = kOffX + kRingGroups * kGroup * kSlotStride * 8;
constexpr int kOffBar = kOffIn + kInStages * kInStageBytes;
constexpr int kNumBars = 2 * kRingGroups + kInStages;
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

// packed FP32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2): one instruction works on a {re, im} register
// pair; ptxas folds make_float2(t, t) into a scalar-broadcast operand and make_float2(-a.y, a.x) into a
// swapped, half-negated operand, so none of these cost extra moves.  The FMA pipe spends two cycles on
// a packed instruction (same flop rate as two scalar ones) but it takes a single issue slot.
__device__ __forceinline__ float2 bc(float t) { return make_float2(t, t); }
__device__ __forceinline__ float2 rot90(float2 a) { return make_float2(-a.y, a.x); }      //  j * a
__device__ __forceinline__ float2 rot270(float2 a) { return make_float2(a.y, -a.x); }     // -j * a
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
// complex product a * t
__device__ __forceinline__ float2 cmul2(float2 a, float2 t)
{
    return __ffma2_rn(bc(t.y), rot90(a), __fmul2_rn(bc(t.x), a));
}

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
            if (k == 0) {
                v[k] = __fadd2_rn(e[k], o[k]);
                v[k + N / 2] = __fadd2_rn(e[k], neg2(o[k]));
            } else if (4 * k == N) {
                // w = SIGN * j
                const float2 r = (SIGN > 0) ? rot90(o[k]) : rot270(o[k]);
                v[k] = __fadd2_rn(e[k], r);
                v[k + N / 2] = __fadd2_rn(e[k], neg2(r));
            } else {
                const float c = kCos32[k * (32 / N)];
                const float sn = (float)SIGN * kSin32[k * (32 / N)];
                // a + w b = a + c b + sn (j b) in two packed FMAs; a - w b = 2a - (a + w b) in one
                const float2 t = __ffma2_rn(bc(c), o[k], e[k]);
                v[k] = __ffma2_rn(bc(sn), rot90(o[k]), t);
                v[k + N / 2] = __ffma2_rn(bc(2.0f), e[k], neg2(v[k]));
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

// unit (float2) offset of transform input n inside a frame slot: rows of 32 padded to 33
__device__ __forceinline__ constexpr int slot_unit(int n) { return n + (n >> 5); }

template <int L>
__global__ void __launch_bounds__(kThreads, 1) pfbch2_fused_1024_kernel(const FusedArgs a)
{
    constexpr int P = L / 2;                               // groups per period
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_tw = s_base + kOffTw;
    const uint32_t s_x = s_base + kOffX;
    const uint32_t s_in = s_base + kOffIn;
    const uint32_t s_bar = s_base + kOffBar;
    // barriers: X ring full[g] / empty[g], input ring in_full[s]
    auto bar_full = [&](int g) { return s_bar + 8u * g; };
    auto bar_empty = [&](int g) { return s_bar + 8u * (kRingGroups + g); };
    auto bar_in_full = [&](int s) { return s_bar + 8u * (2 * kRingGroups + s); };

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
        for (int s = 0; s < kInStages; s++) mbar_init(bar_in_full(s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < kMacWarps) {
        // =========================================================================== MAC warps
        const int j = tid;                                  // window pair (j, j + 512)
        // this thread's first dot product belongs to window j (segment starts on an even frame) or
        // j + 512 (odd); the second one to the other.  Byte offsets inside a frame slot:
        const uint32_t x_thr = s_x + (uint32_t)(slot_unit(j) + (a.par0 ? slot_unit(kM2) : 0)) * 8;
        const uint32_t w2delta = a.par0 ? (uint32_t)(-slot_unit(kM2) * 8) : (uint32_t)(slot_unit(kM2) * 8);
        // this thread's sample inside a staged frame
        const uint32_t in_thr = s_in + (uint32_t)(kM2 - 1 - j) * 8;

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

        const int n_periods = (n_groups + P - 1) / P;      // the last one may be partial
        for (int p = 0; p < n_periods; p++) {
            // one period: L updates of each window = 2L frames = L/2 groups, fully unrolled so that
            // every history slot index below is a compile-time constant
#pragma unroll
            for (int uu = 0; uu < L; uu++) {
                const int g = p * P + (uu >> 1);            // group index inside this CTA
                if (g >= n_groups) continue;                // warp-uniform: tail of the last period
                const int gq = g & (kRingGroups - 1);       // stage of both rings (same depth)
                const int fg = (uu & 1) * 2;                // frame-in-group of this pair's even frame
                if ((uu & 1) == 0) {
                    const uint32_t par = (uint32_t)(g >> 2) & 1u;
                    mbar_wait(bar_in_full(gq), par);          // the group's 16 KB of input have landed
                    mbar_wait(bar_empty(gq), par ^ 1u);       // the FFT warps are done with this X slot
                }
                const uint32_t in_g = in_thr + (uint32_t)gq * kInStageBytes;
                const uint32_t x_g = x_thr + (uint32_t)gq * (kGroup * kSlotStride * 8);

                // ---- even frame of the pair: window 1 gets the new sample
                dA[uu] = lds64(in_g + fg * (kM2 * 8));
                {
                    float2 o1 = make_float2(0.f, 0.f), o2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int n = L - 1; n >= 0; n--) {
                        const float2 s1 = dA[(uu - n + L) % L];
                        const float2 s2 = dB[(uu - 1 - n + 2 * L) % L];
                        o1 = __ffma2_rn(bc(hA[n]), s1, o1);
                        o2 = __ffma2_rn(bc(hB[n]), s2, o2);
                    }
                    const uint32_t slot = x_g + (uint32_t)fg * (kSlotStride * 8);
                    sts64(slot, o1);
                    sts64(slot + w2delta, o2);
                }
                // ---- odd frame of the pair: window 2 gets the new sample, branches swap
                dB[uu] = lds64(in_g + (fg + 1) * (kM2 * 8));
                {
                    float2 o1 = make_float2(0.f, 0.f), o2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int n = L - 1; n >= 0; n--) {
                        const float2 s1 = dA[(uu - n + L) % L];
                        const float2 s2 = dB[(uu - n + L) % L];
                        o1 = __ffma2_rn(bc(hB[n]), s1, o1);
                        o2 = __ffma2_rn(bc(hA[n]), s2, o2);
                    }
                    const uint32_t slot = x_g + (uint32_t)(fg + 1) * (kSlotStride * 8);
                    sts64(slot, o1);
                    sts64(slot + w2delta, o2);
                }
                if (uu & 1) {
                    // group complete: publish it to the FFT warps (this also hands the input stage back:
                    // the producer refills stage gq only after full[gq] of this group has completed)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full(gq));
                }
            }
        }
    } else {
        // =========================================================================== FFT warps
        const int fw = warp - kMacWarps;                    // frame-in-group handled by this warp
        const uint32_t tw_lane = s_tw + (uint32_t)lane * 8; // tw[k1][lane] at + k1*256
        const bool producer = (fw == 0 && lane == 0);       // streams the input, 4 groups ahead
        auto fetch_group = [&](int gg) {
            // 4 consecutive frames are contiguous in the stream: one 16 KB bulk copy
            const int st = gg & (kInStages - 1);
            const float2 *src = a.x + (f_begin + (long long)gg * kGroup) * kM2;
            mbar_arrive_expect_tx(bar_in_full(st), kInStageBytes);
            bulk_g2s(s_in + (uint32_t)st * kInStageBytes, src, kInStageBytes, bar_in_full(st));
        };
        if (producer) {
            for (int gg = 0; gg < kInStages && gg < n_groups; gg++) fetch_group(gg);
        }
        for (int g = 0; g < n_groups; g++) {
            const int gs = g & (kRingGroups - 1);
            mbar_wait(bar_full(gs), (uint32_t)(g >> 2) & 1u);
            // every MAC warp has finished group g, so input stage gs is free again
            if (producer && g + kInStages < n_groups) fetch_group(g + kInStages);
            const uint32_t slot = s_x + (uint32_t)(gs * kGroup + fw) * (kSlotStride * 8) + (uint32_t)lane * 8;

            float2 v[32];
            // pass 1: lane = n2, points X[32 n1 + n2] (row n1 of the padded slot); DFT over n1
#pragma unroll
            for (int n1 = 0; n1 < 32; n1++) v[n1] = lds64(slot + (uint32_t)(n1 * 33) * 8);
            RegDft<32, +1>::run(v);
            __syncwarp();                                   // every lane has read its column
            // twiddle W^(n2 k1), then store transposed: row n2 = lane, column k1
            const uint32_t row = slot + (uint32_t)lane * (32 * 8);      // = slot base + lane*33*8
#pragma unroll
            for (int b = 0; b < 4; b++) {
                float2 t[8];
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (8 * b + i) t[i] = lds64(tw_lane + (8 * b + i) * 256);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int k1 = 8 * b + i;
                    if (k1) v[k1] = cmul2(v[k1], t[i]);
                    sts64(row + (uint32_t)k1 * 8, v[k1]);
                }
            }
            __syncwarp();
            // pass 2: lane = k1, points A[k1][n2] (row n2, column k1): same addresses as pass 1
#pragma unroll
            for (int n2 = 0; n2 < 32; n2++) v[n2] = lds64(slot + (uint32_t)(n2 * 33) * 8);
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
    if (n_groups > (1ll << 30)) n_groups = 1ll << 30;      // the caller finishes the rest
    int dev = 0, sms = 148;
    LC_CUDA_CHECK(cudaGetDevice(&dev));
    LC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // at least 8 groups (32 frames) per CTA so the history prologue and the pipeline fill amortise
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
