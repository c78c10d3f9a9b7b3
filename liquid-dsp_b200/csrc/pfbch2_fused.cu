// pfbch2_fused.cu -- fused firpfbch2_crcf analysis kernel for M = 1024 (sm_100a).
//
// One persistent CTA per SM (512 threads) walks a contiguous range of frames.  Warp-specialised,
// producer / consumer through shared memory with mbarriers (no CTA-wide barrier in steady state);
// registers are moved between the roles with setmaxnreg (launch 128/thread -> MAC 192, FFT 64):
//
//    8 MAC warps (256 threads)   thread j owns the four windows j, j+256, j+512, j+768, i.e. the window
//                                pairs (j, j+512) and (j+256, j+768): per pair two 2m-sample complex
//                                histories and the two 2m-tap branches live in registers (12*2m
//                                registers).  Per frame: two new samples (its own 16 bytes of the
//                                staged frame), four dot products (8m packed FFMA2: tap broadcast x
//                                {re, im} pair).  The four results are exactly the inputs of one
//                                radix-4 decimation-in-frequency butterfly of the 1024-point IDFT, so
//                                the thread also does that butterfly and its three twiddles and stores
//                                the frame as four independent 256-point problems.  The 2m-slot
//                                circular histories are addressed statically by fully unrolling one
//                                period (4m frames).
//    8 FFT warps (256 threads)   two warps per frame, four frames in flight: 256 = 16 x 16, thread =
//                                16 points in registers, radix-2 DIT butterflies with constant twiddles
//                                in packed FP32x2 arithmetic (3 instructions per general butterfly), one
//                                transposition through the frame slot (rows of 16 padded to 17), inter-
//                                pass twiddles from a shared 16x16 table of {c, c, -s, s} (a complex
//                                product is then two packed instructions on register pairs as they are
//                                loaded: no register moves), then 16 coalesced 256-byte row stores per
//                                thread straight to HBM.  Each FRAME slot of the X ring has its own
//                                "full" barrier, so the warp pairs of the four frames of a group start
//                                one frame time apart and the two FFT warps of a sub-partition run half a
//                                group period out of phase (one in its arithmetic while the other waits
//                                for shared memory).
//   input                        one lane of the first FFT warp streams the input with cp.async.bulk
//                                (TMA engine): one 16 KB copy per group of 4 frames, four groups
//                                ahead, into a 4-stage ring; completion is counted in bytes on an
//                                mbarrier.  A stage is refilled once every MAC warp has published the
//                                last frame of the group that read it, so no second barrier is needed.
//
// FP32 pipe and shared-memory pipe work of the two roles overlap by construction (the warp
// scheduler interleaves them); per input sample the kernel reads 8 B and writes 16 B of HBM.
// The 1/M output scale is folded into the taps (exact: M is a power of two).
//
// Reference semantics: src/multichannel/src/firpfbch2.c:244-282 (see firpfbch2.cu).
#include "pfbch2_fused.cuh"
#include "fused_common.cuh"

#include <cmath>
#include <cstdint>

namespace {

using namespace lcf;

constexpr int kM = 1024;
constexpr int kM2 = 512;
constexpr int kMacWarps = 8;
constexpr int kFftWarps = 8;
static_assert(32 * kMacWarps == 256, "thread j of the MAC role owns the windows j + 256 q");
constexpr int kThreads = 32 * (kMacWarps + kFftWarps);         // 512
constexpr int kGroup = 4;                                      // frames per group (two FFT warps each)
constexpr int kRingGroups = 4;                                 // X ring depth, in groups (power of two)
constexpr int kSubStride = 276;                                // float2 per 256-point sub-problem: 16 rows of 17 (+4)
constexpr int kSlotStride = 4 * kSubStride;                    // float2 per frame slot
constexpr int kInStages = 4;                                   // input ring depth, in groups (= kRingGroups)
constexpr int kInStageBytes = kGroup * kM2 * 8;                // 4 frames x 512 samples x 8 B = 16 KB

// shared memory carve-up (bytes)
constexpr int kOffTw4 = 0;                                     // float2 tw4[3][256]: W1024^(j p), p = 1..3
constexpr int kOffX = kOffTw4 + 3 * 256 * 8;                   // X ring
constexpr int kOffIn = kOffX + kRingGroups * kGroup * kSlotStride * 8;
// (the 16x16 table sits behind the rings: placed directly behind tw4 the kernel measured 3 % slower, the position of
// the rings themselves makes no difference -- profiles/r02_fused_smem_layout_search.log)
constexpr int kOffTw16 = kOffIn + kInStages * kInStageBytes;   // float4 tw16[16][16]: {c, c, -s, s} of W256^(k1 r2)
constexpr int kOffBar = kOffTw16 + 16 * 16 * 16;
constexpr int kSlots = kRingGroups * kGroup;                   // frame slots of the X ring
constexpr int kNumBars = kSlots + kRingGroups + kInStages;
constexpr int kSmemBytes = kOffBar + kNumBars * 8;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// ---------------------------------------------------------------------------------- the kernel
struct FusedArgs {
    const float2 *hist;        // H samples preceding x[0] (never null)
    const float2 *x;           // segment input
    float2 *y;                 // [n_frames][1024]
    const float *taps;         // [2][L][1024] as in firpfbch2.cu, pre-scaled by 1/M
    const float2 *tw;          // tw4[3][256] (float2) followed by tw16[16][16] as {c, c, -s, s} (float4)
    int H;
    int n_groups;              // groups of 4 frames in the segment
    int groups_per_cta;
};

// unit (float2) offset of element r of a 256-point sub-problem: rows of 16 padded to 17
__device__ __forceinline__ constexpr int sub_unit(int r) { return r + (r >> 4); }

// radix-4 DIF butterfly over the four windows of a thread + the three twiddles, stored as four
// 256-point sub-problems z_p[j], p = 0..3 (slot = this thread's address inside sub-problem 0)
__device__ __forceinline__ void radix4_store(uint32_t slot, float2 a0, float2 a1, float2 a2, float2 a3,
                                             float2 t1, float2 t2, float2 t3)
{
    const float2 s02 = __fadd2_rn(a0, a2), d02 = __fadd2_rn(a0, neg2(a2));
    const float2 s13 = __fadd2_rn(a1, a3), d13 = __fadd2_rn(a1, neg2(a3));
    const float2 u0 = __fadd2_rn(s02, s13);
    const float2 u2 = __fadd2_rn(s02, neg2(s13));
    const float2 u1 = __fadd2_rn(d02, rot90(d13));      // W4 = +j (backward transform)
    const float2 u3 = __fadd2_rn(d02, rot270(d13));
    sts64(slot, u0);
    sts64(slot + 1 * (kSubStride * 8), cmul2_rot(u1, t1));
    sts64(slot + 2 * (kSubStride * 8), cmul2_rot(u2, t2));
    sts64(slot + 3 * (kSubStride * 8), cmul2_rot(u3, t3));
}

// a * t with t given as {c, c, -s, s}: two packed instructions on the register pairs as loaded
__device__ __forceinline__ float2 cmulq(float2 a, float4 t)
{
    return __ffma2_rn(make_float2(t.z, t.w), make_float2(a.y, a.x), __fmul2_rn(make_float2(t.x, t.y), a));
}

template <int L, int PAR0, int X = 0>
__global__ void __launch_bounds__(kThreads, 1) pfbch2_fused_1024_kernel(const FusedArgs a)
{
    // X: profiling switches, compiled with -DLC_FUSED_EXPERIMENTS only (X = 0 is the product kernel):
    //   1 the MAC warps skip the dot products, 2 the FFT warps skip the DFT arithmetic -- what each role costs alone
    constexpr int P = L / 2;                               // groups per period
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_x = s_base + kOffX;
    const uint32_t s_in = s_base + kOffIn;
    const uint32_t s_bar = s_base + kOffBar;
    // barriers: X ring full[frame slot] / empty[group], input ring in_full[stage]
    auto bar_full = [&](int slot) { return s_bar + 8u * slot; };
    auto bar_empty = [&](int g) { return s_bar + 8u * (kSlots + g); };
    auto bar_in_full = [&](int s) { return s_bar + 8u * (kSlots + kRingGroups + s); };

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
        float2 *tw_s = reinterpret_cast<float2 *>(smem + kOffTw4), *twq_s = reinterpret_cast<float2 *>(smem + kOffTw16);
        for (int i = tid; i < 3 * 256; i += kThreads) tw_s[i] = a.tw[i];
        for (int i = tid; i < 2 * 16 * 16; i += kThreads) twq_s[i] = a.tw[3 * 256 + i];
    }
    if (tid == 0) {
        for (int sl = 0; sl < kSlots; sl++) mbar_init(bar_full(sl), kMacWarps);
        for (int g = 0; g < kRingGroups; g++) mbar_init(bar_empty(g), kFftWarps);
        for (int s = 0; s < kInStages; s++) mbar_init(bar_in_full(s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < kMacWarps) {
        // =========================================================================== MAC warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
        const int j = tid;                                  // windows j, j+256, j+512, j+768
        PairState<L> s0, s1;                                // pairs (j, j+512) and (j+256, j+768)
        pair_load<L, kM>(s0, a.taps, a.hist, a.x, a.H, j, f_begin);
        pair_load<L, kM>(s1, a.taps, a.hist, a.x, a.H, j + 256, f_begin);
        // this thread's two samples inside a staged frame, its z_0 element inside a frame slot, its twiddles
        const uint32_t in_thr = s_in + (uint32_t)(kM2 - 1 - j) * 8;       // pair 1 sits 256 samples lower
        const uint32_t x_thr = s_x + (uint32_t)sub_unit(j) * 8;
        const uint32_t tw_thr = s_base + kOffTw4 + (uint32_t)j * 8;

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
                    mbar_wait(bar_empty(gq), par ^ 1u);       // the FFT warps are done with this group of X slots
                }
                const uint32_t in_g = in_thr + (uint32_t)gq * kInStageBytes;
                const uint32_t x_g = x_thr + (uint32_t)gq * (kGroup * kSlotStride * 8);
                const float2 t1 = lds64(tw_thr), t2 = lds64(tw_thr + 256 * 8), t3 = lds64(tw_thr + 512 * 8);
                float2 o1, o2, o3, o4;

                // ---- even local frame: the window pushed on even frames gets the new sample
                s0.dA[uu] = lds64(in_g + fg * (kM2 * 8));
                s1.dA[uu] = lds64(in_g + fg * (kM2 * 8) - 256 * 8);
                if (X & 1) { o1 = s0.dA[uu]; o2 = s0.dB[(uu + 1) % L]; o3 = s1.dA[uu]; o4 = s1.dB[(uu + 1) % L]; }
                else {
                pair_mac<L, 0>(s0, uu, o1, o2);
                pair_mac<L, 0>(s1, uu, o3, o4);
                }
                // o1/o3 belong to windows j, j+256 when the segment starts on an even frame (PAR0 = 0),
                // to windows j+512, j+768 otherwise
                if (PAR0 == 0) radix4_store(x_g + (uint32_t)fg * (kSlotStride * 8), o1, o3, o2, o4, t1, t2, t3);
                else           radix4_store(x_g + (uint32_t)fg * (kSlotStride * 8), o2, o4, o1, o3, t1, t2, t3);
                // frame complete: publish it to its two FFT warps
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full(gq * kGroup + fg));

                // ---- odd local frame: the other window of each pair gets the new sample, branches swap
                s0.dB[uu] = lds64(in_g + (fg + 1) * (kM2 * 8));
                s1.dB[uu] = lds64(in_g + (fg + 1) * (kM2 * 8) - 256 * 8);
                if (X & 1) { o1 = s0.dA[uu]; o2 = s0.dB[uu]; o3 = s1.dA[uu]; o4 = s1.dB[uu]; }
                else {
                pair_mac<L, 1>(s0, uu, o1, o2);
                pair_mac<L, 1>(s1, uu, o3, o4);
                }
                if (PAR0 == 0) radix4_store(x_g + (uint32_t)(fg + 1) * (kSlotStride * 8), o1, o3, o2, o4, t1, t2, t3);
                else           radix4_store(x_g + (uint32_t)(fg + 1) * (kSlotStride * 8), o2, o4, o1, o3, t1, t2, t3);

                // (the last frame of a group also hands the input stage back: the producer refills stage gq
                // only after this barrier has completed)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full(gq * kGroup + fg + 1));
            }
        }
    } else {
        // =========================================================================== FFT warps
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        const int fwarp = warp - kMacWarps;                 // 0..7
        const int f = fwarp >> 1;                           // frame-in-group handled by this warp pair
        const int h = fwarp & 1;                            // which half of the work
        // pass A: lane = (sub, r2): sub-problem p = 2h + sub, column r2
        const int pA = 2 * h + (lane >> 4), r2 = lane & 15;
        const uint32_t offA = (uint32_t)(pA * kSubStride + r2) * 8;           // + r1 * 17 * 8 (loads)
        const uint32_t offAt = (uint32_t)(pA * kSubStride + r2 * 17) * 8;      // + k1 * 8 (transposed stores)
        const uint32_t tw_lane = s_base + kOffTw16 + (uint32_t)r2 * 16;        // tw16[k1][r2] at + k1 * 256
        // pass B: lane = (k1 - 8h, p): row k1 = 8h + lane/4 of sub-problem p = lane % 4
        const int pB = lane & 3, k1B = 8 * h + (lane >> 2);
        const uint32_t offB = (uint32_t)(pB * kSubStride + k1B) * 8;           // + r2 * 17 * 8
        // one lane of the LAST frame's warp pair streams the input, 4 groups ahead: that frame's "full" barrier
        // completes when every MAC warp has finished the group, i.e. has read the whole input stage
        const bool producer = (fwarp == 2 * (kGroup - 1) && lane == 0);
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
            mbar_wait(bar_full(gs * kGroup + f), (uint32_t)(g >> 2) & 1u);
            // (producer: every MAC warp has finished group g, so input stage gs is free again)
            if (producer && g + kInStages < n_groups) fetch_group(g + kInStages);
            const uint32_t slot = s_x + (uint32_t)(gs * kGroup + f) * (kSlotStride * 8);

            float2 v[16];
            // pass A: 16-point DFT down the column r2 of sub-problem pA, then twiddle W256^(r2 k1)
#pragma unroll
            for (int r1 = 0; r1 < 16; r1++) v[r1] = lds64(slot + offA + (uint32_t)(r1 * 17) * 8);
            if (!(X & 2)) {
            RegDft<16, +1>::run(v);
#pragma unroll
            for (int k1 = 1; k1 < 16; k1++) v[k1] = cmulq(v[k1], lds128(tw_lane + k1 * 256));
            }
            __syncwarp();                                   // both sub-problems of this warp fully read
            // transposed store: row r2, column k1 (conflict-free: lanes are 17 units apart)
#pragma unroll
            for (int k1 = 0; k1 < 16; k1++) sts64(slot + offAt + (uint32_t)k1 * 8, v[k1]);
            // the other warp of this frame holds the other two sub-problems
            asm volatile("bar.sync %0, 64;" ::"r"(1 + f) : "memory");
            // pass B: row k1 of sub-problem pB: 16-point DFT over r2 -> Y[4 (k1 + 16 k2) + p]
#pragma unroll
            for (int q = 0; q < 16; q++) v[q] = lds64(slot + offB + (uint32_t)(q * 17) * 8);
            if (!(X & 2)) RegDft<16, +1>::run(v);
            // the slot has been consumed (the DFT above needed every loaded value)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty(gs));
            float2 *yf = a.y + ((f_begin + (long long)g * kGroup + f) * (long long)kM) + 32 * h + lane;
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) __stcs(yf + 64 * k2, v[k2]);
        }
    }
}

template <int L>
void launch_L(const FusedArgs &a, int par0, unsigned int grid, cudaStream_t st)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured)) {
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_fused_1024_kernel<L, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_fused_1024_kernel<L, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    }
    if (par0) pfbch2_fused_1024_kernel<L, 1><<<grid, kThreads, kSmemBytes, st>>>(a);
    else      pfbch2_fused_1024_kernel<L, 0><<<grid, kThreads, kSmemBytes, st>>>(a);
    LC_LAUNCH_CHECK();
}

#ifdef LC_FUSED_EXPERIMENTS
template <int X>
void launch_X(const FusedArgs &a, unsigned int grid, cudaStream_t st)
{
    LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_fused_1024_kernel<14, 0, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    pfbch2_fused_1024_kernel<14, 0, X><<<grid, kThreads, kSmemBytes, st>>>(a);
    LC_LAUNCH_CHECK();
}
#endif

}  // namespace

namespace lc {
namespace pfb {

bool fused_pfbch2_supported(unsigned int M, unsigned int L)
{
    return M == 1024 && (L == 4 || L == 6 || L == 8 || L == 10 || L == 12 || L == 14);
}

void fused_pfbch2_make_twiddles(std::vector<float2> &tw)
{
    // tw4[p-1][j] = e^{+j 2 pi j p / 1024}, p = 1..3, j < 256; then tw16[k1][r2] = e^{+j 2 pi k1 r2 / 256} = c + j s
    // stored as the four floats {c, c, -s, s}: (c, c) * a + (-s, s) * swap(a) is the complex product
    tw.resize(3 * 256 + 2 * 16 * 16);
    for (int p = 1; p <= 3; p++)
        for (int j = 0; j < 256; j++) {
            double ang = 2.0 * M_PI * (double)(j * p) / 1024.0;
            tw[(p - 1) * 256 + j] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (int k1 = 0; k1 < 16; k1++)
        for (int r2 = 0; r2 < 16; r2++) {
            double ang = 2.0 * M_PI * (double)(k1 * r2) / 256.0;
            const float c = (float)cos(ang), sn = (float)sin(ang);
            tw[3 * 256 + 2 * (k1 * 16 + r2)] = make_float2(c, c);
            tw[3 * 256 + 2 * (k1 * 16 + r2) + 1] = make_float2(-sn, sn);
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
    sms = usable_sms(sms);
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
    a.n_groups = (int)n_groups;
    a.groups_per_cta = (int)per;
    g_last_chunk_frames.store((unsigned long long)per * kGroup);
#ifdef LC_FUSED_EXPERIMENTS
    if (const char *env = getenv("LIQUID_CUDA_FUSED_X")) {
        const int x = atoi(env);
        if (x && L == 14 && !(parity0 & 1)) {
            switch (x) {
            case 1: launch_X<1>(a, (unsigned int)ctas, st); break;
            case 2: launch_X<2>(a, (unsigned int)ctas, st); break;
            case 3: launch_X<3>(a, (unsigned int)ctas, st); break;
            default: fatal("launch_pfbch2_fused", "unknown experiment");
            }
            return n_groups * kGroup;
        }
    }
#endif
    switch (L) {
    case 4: launch_L<4>(a, parity0 & 1, (unsigned int)ctas, st); break;
    case 6: launch_L<6>(a, parity0 & 1, (unsigned int)ctas, st); break;
    case 8: launch_L<8>(a, parity0 & 1, (unsigned int)ctas, st); break;
    case 10: launch_L<10>(a, parity0 & 1, (unsigned int)ctas, st); break;
    case 12: launch_L<12>(a, parity0 & 1, (unsigned int)ctas, st); break;
    case 14: launch_L<14>(a, parity0 & 1, (unsigned int)ctas, st); break;
    default: fatal("launch_pfbch2_fused", "unsupported tap count");
    }
    return n_groups * kGroup;
}

}  // namespace pfb
}  // namespace lc
