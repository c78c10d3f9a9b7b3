// pfbch2_cluster.cu -- fused firpfbch2_crcf analysis kernel for M = 4096 (sm_100a), one frame spread
// over a thread-block cluster of 8 CTAs with the transform exchanged through distributed shared memory.
//
// Why a cluster: the per-window state of one frame (4096 windows x 2m samples + 2m taps: 864 KB at m = 9)
// does not fit the registers of one SM, and re-reading it from shared memory or L2 every frame would
// need several times the available bandwidth.  Eight SMs hold it in registers (108 KB each).
//
// The 4096-point backward DFT is split 64 x 64 (w = n0 + 64 n1, k = k1 + 64 k0):
//     Y[k1 + 64 k0] = sum_n0 W64^(n0 k0) { W4096^(n0 k1) sum_n1 W64^(n1 k1) X[n0 + 64 n1] }
//   * CTA c owns the windows with n0 in [8c, 8c+8): 512 windows = 256 pairs (w, w+2048), one pair per
//     MAC thread; its input samples are 64-byte runs of the stream, read once with cp.async.
//   * phase 1 (CTA c): eight 64-point DFTs over n1 (one per n0), times W4096^(n0 k1) -> Z[n0][k1]
//   * exchange: Z[n0][k1] goes to CTA k1/8 -- st.async straight from registers into the destination CTA's
//     shared memory, completion counted in bytes on the destination's mbarrier (256-byte runs per warp
//     store); receive slots are handed back with remote mbarrier arrives (credits)
//   * phase 2 (CTA c'): eight 64-point DFTs over n0 (one per k1 in [8c', 8c'+8)) -> Y[k1 + 64 k0], written
//     to HBM as 64-byte runs (the 8 CTAs of the cluster fill whole lines between them).
// Each 64-point DFT is 8 x 8: eight threads hold eight points each, two register radix-8 passes with one
// in-place transposition through shared memory (layouts padded so every access is bank-conflict-free).
//
// Warp roles per CTA (512 threads, registers moved with setmaxnreg 144 / 112):
//    8 MAC warps   thread = pair (n1, n0): 2 x 2m-sample histories + 2 x 2m taps in registers, per frame
//                  one 8-byte sample (own cp.async ring slot, no barrier needed), 4m packed FFMA2,
//                  two STS into the frame's X slot.
//    8 FFT warps   warp fw handles the frames = fw mod 8: phase 1 of frame f, then phase 2 of the frame one
//                  group (8 frames) earlier, so the exchange latency hides behind a group of work.  Each
//                  phase runs as two independent halves in one instruction stream (two dependency chains),
//                  and every 64-point DFT stays inside the warp: no block-level barrier anywhere.
//
// Reference semantics: src/multichannel/src/firpfbch2.c:244-282 (see firpfbch2.cu).  The 1/M output
// scale is folded into the taps (exact: M is a power of two).
#include "pfbch2_cluster.cuh"
#include "fused_common.cuh"

#include <cmath>
#include <cstdint>
#include <type_traits>

namespace {

using namespace lcf;

constexpr int kM = 4096;
constexpr int kM2 = 2048;
constexpr int kCluster = 8;
constexpr int kMacWarps = 8;
constexpr int kFftWarps = 8;
constexpr int kMacThreads = 32 * kMacWarps;                    // 256
constexpr int kThreads = 32 * (kMacWarps + kFftWarps);         // 512
constexpr int kGroup = 8;                                      // frames per group (one per FFT warp)
constexpr int kXStride = 74;                                   // units (float2) per n0 block of an X slot: 8 rows of 9 (+2)
constexpr int kXSlotUnits = 8 * kXStride;                      // 592
constexpr int kRecvGroups = 3;                                 // receive ring depth in groups
constexpr int kLag = 1;                                        // groups phase 2 trails phase 1 (< kRecvGroups)
constexpr int kRecvStride = 72;                                // units per source CTA of a receive slot: 64 (+8)
constexpr int kRecvSlotUnits = 8 * kRecvStride;                // 576
constexpr int kRecvSlots = kRecvGroups * kGroup;               // 24
constexpr int kFrameBytes = 64 * 64 * 8 / kCluster;            // bytes of Z one CTA receives per frame: 4096

// Ring depths that divide the MAC warps' unrolled period (2L frames = L double frames), so that every
// ring slot and most of every barrier parity is a compile-time constant inside the period:
//   XS  X ring (MAC -> phase 1) depth in frames, divides 2L;   D  input ring depth in double frames, divides L
template <int L> struct Rings;
template <> struct Rings<18> { static constexpr int XS = 12, D = 9; };
template <> struct Rings<14> { static constexpr int XS = 14, D = 7; };
template <> struct Rings<16> { static constexpr int XS = 16, D = 8; };
template <> struct Rings<12> { static constexpr int XS = 12, D = 6; };
template <> struct Rings<10> { static constexpr int XS = 10, D = 10; };
template <> struct Rings<8>  { static constexpr int XS = 16, D = 8; };
template <> struct Rings<6>  { static constexpr int XS = 12, D = 6; };
template <> struct Rings<4>  { static constexpr int XS = 8, D = 4; };
constexpr int kMaxXS = 16;

// shared memory carve-up (bytes); barriers first so that their offsets do not depend on L
constexpr int kOffBarRecvFull = 0;                             // [kRecvSlots]
constexpr int kOffBarRecvFree = kOffBarRecvFull + 8 * kRecvSlots;
constexpr int kOffBarXFull = kOffBarRecvFree + 8 * kRecvSlots; // [XS]
constexpr int kOffBarXEmpty = kOffBarXFull + 8 * kMaxXS;       // [XS]
constexpr int kOffTw8 = 1024;                                  // float2 tw8[8][8]: W64^(q r) at [r][q]
constexpr int kOffTz = kOffTw8 + 64 * 8;                       // float2 tz[2][8][32]: W4096^(n0 k1) of (half, s, lane)
constexpr int kOffRecv = kOffTz + 2 * 8 * 32 * 8;
constexpr int kOffX = kOffRecv + kRecvSlots * kRecvSlotUnits * 8;
template <int L> struct Smem {
    static constexpr int kOffIn = kOffX + Rings<L>::XS * kXSlotUnits * 8;
    static constexpr int kBytes = kOffIn + Rings<L>::D * 2 * kMacThreads * 8;
};
static_assert(kOffBarXEmpty + 8 * kMaxXS <= kOffTw8, "barrier block overflows");

struct ClusterArgs {
    const float2 *hist;        // H samples preceding x[0] (never null)
    const float2 *x;           // segment input
    float2 *y;                 // [n_frames][4096]
    const float *taps;         // [2][L][4096] as in firpfbch2.cu, pre-scaled by 1/M
    const float2 *tw;          // tw8[8][8] (r, q) followed by tw4096[64][64] (n0, k1)
    int H;
    int n_groups;              // groups of 8 frames in the segment
    int groups_per_cluster;
};

// ---------------------------------------------------------------------------------- cluster PTX helpers
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 8-byte store into another CTA's shared memory; the destination CTA's mbarrier counts the bytes
__device__ __forceinline__ void st_async64(uint32_t raddr, float2 v, uint32_t rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(raddr),
                 "f"(v.x), "f"(v.y), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t rbar)
{
    // relaxed: the credit only guards a write-after-read hazard and is issued after the values read from
    // the slot have been consumed (a release here costs MEMBAR.ALL.GPU + ERRBAR per arrive)
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// sender thread (warp half h, lane = 8 i + r) holds n0 = 8c + n0_of(h, i); two n0 of one half-warp are 4 apart
__device__ __forceinline__ constexpr int n0_of(int h, int i) { return (i & 1) * 4 + (i >> 1) + 2 * h; }
// position of Z[n0 = 8c + q][k1 = 8c' + r] inside source CTA c's 64-unit block of a receive slot:
// the sender's thread order (its warp writes one 256-byte run)
__device__ __forceinline__ constexpr int recv_pos(int q) { return ((q >> 1) & 1) * 32 + (((q >> 2) & 1) + 2 * (q & 1)) * 8; }

// complex product a * (c + j s) with the twiddle held as c and the pair {-s, s}: two packed instructions
// with no operand preparation (the swap of a is an operand modifier)
__device__ __forceinline__ float2 cmul3(float2 a, float c, float2 ms_s)
{
    return __ffma2_rn(ms_s, make_float2(a.y, a.x), __fmul2_rn(bc(c), a));
}
// same with the twiddle as loaded from a {c, s} table: the {-s, s} pair costs one integer sign flip (ALU pipe)
__device__ __forceinline__ float2 cmul_cs(float2 a, float2 t)
{
    return cmul3(a, t.x, make_float2(__int_as_float(__float_as_int(t.y) ^ (int)0x80000000), t.y));
}

template <int L, int PAR0>
__global__ void __launch_bounds__(kThreads, 1) pfbch2_cluster_4096_kernel(const ClusterArgs a)
{
    constexpr int XS = Rings<L>::XS, D = Rings<L>::D;
    static_assert((2 * L) % XS == 0 && L % D == 0 && XS >= kGroup && XS <= kMaxXS, "ring depths must divide the period");
    constexpr int K = 2 * L / XS;                          // X ring revolutions per period
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_tw8 = s_base + kOffTw8;
    const uint32_t s_x = s_base + kOffX;
    const uint32_t s_recv = s_base + kOffRecv;
    const uint32_t s_in = s_base + Smem<L>::kOffIn;
    // barriers: X ring full[slot] / empty[slot]; receive ring recv_full[slot] (bytes) / recv_free[slot] (credits)
    auto bar_x_full = [&](int s) { return s_base + kOffBarXFull + 8u * s; };
    auto bar_x_empty = [&](int s) { return s_base + kOffBarXEmpty + 8u * s; };
    auto bar_recv_full = [&](int s) { return s_base + kOffBarRecvFull + 8u * s; };
    auto bar_recv_free = [&](int s) { return s_base + kOffBarRecvFree + 8u * s; };

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int crank = (int)cluster_ctarank();
    const int cid = (int)blockIdx.x / kCluster;

    const int g_begin = cid * a.groups_per_cluster;
    int g_end = g_begin + a.groups_per_cluster;
    if (g_end > a.n_groups) g_end = a.n_groups;
    const int n_groups = g_end - g_begin;                  // > 0 by construction of the grid, same for the 8 CTAs
    const int n_frames = n_groups * kGroup;
    const long long f_begin = (long long)g_begin * kGroup; // first frame of this cluster (even)

    // ---- one-time setup
    if (tid < 64) reinterpret_cast<float2 *>(smem + kOffTw8)[tid] = a.tw[tid];
    {
        // tz[half][s][lane]: the phase-1 output twiddle W4096^(n0 k1) of FFT lane (i, q) in half `half`:
        // n0 = 8c + n0_of(half, i), k1 = q + 8 s
        const int half = tid >> 8, sd = (tid >> 5) & 7, ln = tid & 31;
        reinterpret_cast<float2 *>(smem + kOffTz)[tid] =
            __ldg(a.tw + 64 + (8 * crank + n0_of(half, ln >> 3)) * 64 + (ln & 7) + 8 * sd);
    }
    if (tid == 0) {
        for (int s = 0; s < XS; s++) {
            mbar_init(bar_x_full(s), kMacWarps);
            mbar_init(bar_x_empty(s), 1);                  // the FFT warp of the frame
        }
        for (int s = 0; s < kRecvSlots; s++) {
            mbar_init(bar_recv_full(s), 1);
            mbar_init(bar_recv_free(s), kCluster);         // one warp of each of the 8 receivers
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // arm the first use of every receive slot: one frame's worth of bytes from the 8 senders
        for (int s = 0; s < kRecvSlots; s++) mbar_arrive_expect_tx(bar_recv_full(s), kFrameBytes);
    }
    __syncthreads();
    cluster_sync_all();                                    // every CTA's barriers exist before any remote access

    if (warp < kMacWarps) {
        // =========================================================================== MAC warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        const int n1 = tid >> 3, n0l = tid & 7;             // a warp covers four 64-byte runs of the stream
        const int j = 8 * crank + n0l + 64 * n1;            // pair (j, j + 2048)
        PairState<L> s;
        pair_load<L, kM>(s, a.taps, a.hist, a.x, a.H, j, f_begin);
        const uint32_t in_thr = s_in + (uint32_t)(tid ^ 7) * 8;           // stream order inside a 64-byte run
        const uint32_t x_thr = s_x + (uint32_t)(n0l * kXStride + n1 + (n1 >> 3)) * 8;   // X[n0][n1]; X[n0][n1+32] is 36 units on
        const int n_dbl = n_frames / 2;
        // `nxt` walks the stream one double frame per step, D - 1 double frames ahead of the MACs
        const float2 *nxt = a.x + f_begin * kM2 + (kM2 - 1 - j);         // this pair's sample of local frame 0
#pragma unroll
        for (int u = 0; u < D - 1; u++) {
            if (u < n_dbl) {
                cp_async8(in_thr + (uint32_t)u * (2 * kMacThreads * 8), nxt);
                cp_async8(in_thr + (uint32_t)u * (2 * kMacThreads * 8) + kMacThreads * 8, nxt + kM2);
            }
            cp_async_commit();
            nxt += 2 * kM2;
        }

        const int n_periods = (n_dbl + L - 1) / L;          // the last one may be partial
        for (int p = 0; p < n_periods; p++) {
            const uint32_t pk = (uint32_t)(K * p) & 1u;     // parity of the X ring revolution this period starts in
            const int u0 = p * L;
            // one period: L updates of each window = 2L frames, fully unrolled so that every history slot,
            // ring slot and barrier index below is a compile-time constant
#pragma unroll
            for (int uu = 0; uu < L; uu++) {
                if (u0 + uu >= n_dbl) continue;             // warp-uniform: tail of the last period
                cp_async_wait<D - 2>();                     // double frame u has landed (own slots only)
                constexpr int kDbl = 2 * kMacThreads * 8;
                const uint32_t in_u = in_thr + (uint32_t)(uu % D) * kDbl;
                const float2 xe = lds64(in_u), xo = lds64(in_u + kMacThreads * 8);
                {
                    // refill the slot that was read one iteration ago with double frame u + D - 1
                    const uint32_t dst = in_thr + (uint32_t)((uu + D - 1) % D) * kDbl;
                    if (u0 + uu + D - 1 < n_dbl) {
                        cp_async8(dst, nxt);
                        cp_async8(dst + kMacThreads * 8, nxt + kM2);
                    }
                    cp_async_commit();                      // always: keeps the group count uniform
                    nxt += 2 * kM2;
                }
                float2 o1, o2;
                // ---- even local frame: the window pushed on even frames gets the new sample
                {
                    const int fl = 2 * uu, xs = fl % XS;              // constants after unrolling
                    s.dA[uu] = xe;
                    pair_mac_split<L, 0>(s, uu, o1, o2);
                    mbar_wait(bar_x_empty(xs), (pk ^ (uint32_t)((fl / XS) & 1)) ^ 1u);
                    const uint32_t xa = x_thr + (uint32_t)xs * (kXSlotUnits * 8);
                    // o1 is window j when the segment starts on an even frame (PAR0 = 0), window j + 2048 otherwise
                    sts64(xa, PAR0 == 0 ? o1 : o2);
                    sts64(xa + 36 * 8, PAR0 == 0 ? o2 : o1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_x_full(xs));
                }
                // ---- odd local frame: the other window of the pair gets the new sample, branches swap
                {
                    const int fl = 2 * uu + 1, xs = fl % XS;
                    s.dB[uu] = xo;
                    pair_mac_split<L, 1>(s, uu, o1, o2);
                    mbar_wait(bar_x_empty(xs), (pk ^ (uint32_t)((fl / XS) & 1)) ^ 1u);
                    const uint32_t xa = x_thr + (uint32_t)xs * (kXSlotUnits * 8);
                    sts64(xa, PAR0 == 0 ? o1 : o2);
                    sts64(xa + 36 * 8, PAR0 == 0 ? o2 : o1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_x_full(xs));
                }
            }
        }
        cp_async_wait<0>();
    } else {
        // =========================================================================== FFT warps
        asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
        const int fw = warp - kMacWarps;                    // 0..7: this warp handles the frames = fw mod 8
        const int i = lane >> 3, q = lane & 7;
        // Both phases run as two independent halves (A: va, B: vb) in one instruction stream, so each
        // hides the other's latencies, and every 64-point DFT stays inside the warp (only __syncwarp).
        // ---- phase 1: half hf holds n0 = 8c + n0_of(hf, i); column q in pass 1, row r = q in pass 2
        const uint32_t p1_colA = s_x + (uint32_t)(n0_of(0, i) * kXStride + q) * 8;       // cell (p, q): + 9 p units
        const uint32_t p1_colB = s_x + (uint32_t)(n0_of(1, i) * kXStride + q) * 8;
        const uint32_t p1_rowA = s_x + (uint32_t)(n0_of(0, i) * kXStride + 9 * q) * 8;   // cell (r = q, qq): + qq units
        const uint32_t p1_rowB = s_x + (uint32_t)(n0_of(1, i) * kXStride + 9 * q) * 8;
        const uint32_t tw_p1 = s_tw8 + (uint32_t)q * 8;                                   // tw8[r][q] at + 64 r
        const uint32_t tz_lane = s_base + kOffTz + (uint32_t)lane * 8;                    // tz[half][s][lane] at + 2048 half + 256 s
        // exchange: Z[n0][q + 8 s] goes to CTA s, position (half, lane) of source block `crank`
        const uint32_t send_off = (uint32_t)kOffRecv + (uint32_t)(crank * kRecvStride + lane) * 8;   // + 256 half
        const uint32_t rbase = map_to_cta(s_base, 0);                                     // CTA windows are equally spaced
        const uint32_t rstep = map_to_cta(s_base, 1) - rbase;
        // ---- phase 2: k1 = 8c + r with r = lane % 8; half hf holds column q' = Q(hf, lane / 8) in pass 1,
        //      row r2 = q' in pass 2 (the two q' of a half-warp differ in bits 0 and 2: conflict-free)
        const int qpA = (i == 0) ? 0 : (i == 1) ? 5 : (i == 2) ? 1 : 4;
        const int qpB = (i == 0) ? 2 : (i == 1) ? 7 : (i == 2) ? 3 : 6;
        const uint32_t p2_colA = s_recv + (uint32_t)(recv_pos(qpA) + q) * 8;              // cell (p', q'): + 72 p' units
        const uint32_t p2_colB = s_recv + (uint32_t)(recv_pos(qpB) + q) * 8;
        const uint32_t p2_rowA = s_recv + (uint32_t)(qpA * kRecvStride + q) * 8;          // cell (r2 = q', q''): + recv_pos(q'') units
        const uint32_t p2_rowB = s_recv + (uint32_t)(qpB * kRecvStride + q) * 8;
        const uint32_t tw_p2A = s_tw8 + (uint32_t)qpA * 8;
        const uint32_t tw_p2B = s_tw8 + (uint32_t)qpB * 8;
        float2 *y_thr = a.y + (f_begin + fw - (long long)kLag * kGroup) * (long long)kM + (8 * crank + q);   // + it * 8 frames + 64 q' + 512 s2

        auto twiddle8 = [&](float2 (&v)[8], uint32_t tw) {
#pragma unroll
            for (int r = 1; r < 8; r++) {
                v[r] = cmul_cs(v[r], lds64(tw + 64 * r));
            }
        };

        // ring state, advanced once per iteration
        int xs = fw;                                        // X slot of the phase-1 frame (advances by 8 mod XS)
        uint32_t xpar = 0;
        int g1 = 0;                                         // receive-ring group of the phase-1 frame, its use parity
        uint32_t use1 = 0, par1 = 0;
        int g2 = 0;                                         // same for the phase-2 frame
        uint32_t par2 = 0;
        auto body = [&](auto do1_c, auto do2_c, int it) {
            constexpr bool DO1 = decltype(do1_c)::value, DO2 = decltype(do2_c)::value;
            float2 va[8], vb[8];
            if (DO1) {
                // ------------------------------------------------ phase 1 of frame 8 it + fw
                const uint32_t xo = (uint32_t)xs * (kXSlotUnits * 8);
                mbar_wait(bar_x_full(xs), xpar);
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    va[p] = lds64(p1_colA + xo + (uint32_t)(9 * p) * 8);
                    vb[p] = lds64(p1_colB + xo + (uint32_t)(9 * p) * 8);
                }
                RegDft<8, +1>::run(va);
                RegDft<8, +1>::run(vb);
                twiddle8(va, tw_p1);
                twiddle8(vb, tw_p1);
                // in-place transposition: every thread rewrites exactly the cells it read
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    sts64(p1_colA + xo + (uint32_t)(9 * r) * 8, va[r]);
                    sts64(p1_colB + xo + (uint32_t)(9 * r) * 8, vb[r]);
                }
                __syncwarp();
#pragma unroll
                for (int qq = 0; qq < 8; qq++) {
                    va[qq] = lds64(p1_rowA + xo + (uint32_t)qq * 8);
                    vb[qq] = lds64(p1_rowB + xo + (uint32_t)qq * 8);
                }
                RegDft<8, +1>::run(va);
                RegDft<8, +1>::run(vb);
                __syncwarp();                               // the frame has left the X slot
                if (lane == 0) mbar_arrive(bar_x_empty(xs));
                // receive slot of this frame in every CTA; wait until all 8 receivers have handed it back
                const int rs1 = g1 * kGroup + fw;
                if (use1 > 0) mbar_wait(bar_recv_free(rs1), par1 ^ 1u);
                const uint32_t dst = send_off + (uint32_t)rs1 * (kRecvSlotUnits * 8);
                const uint32_t dbar = (uint32_t)kOffBarRecvFull + 8u * rs1;
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    const uint32_t rcta = rbase + (uint32_t)s * rstep;
                    st_async64(rcta + dst, cmul_cs(va[s], lds64(tz_lane + 256 * s)), rcta + dbar);
                    st_async64(rcta + dst + 256, cmul_cs(vb[s], lds64(tz_lane + 2048 + 256 * s)), rcta + dbar);
                }
                xs += kGroup;
                if (xs >= XS) {
                    xs -= XS;
                    xpar ^= 1u;
                }
                if (++g1 == kRecvGroups) {
                    g1 = 0;
                    use1++;
                    par1 ^= 1u;
                }
            }
            if (DO2) {
                // ------------------------------------------------ phase 2 of frame 8 (it - kLag) + fw
                const int rs2 = g2 * kGroup + fw;
                const uint32_t ro = (uint32_t)rs2 * (kRecvSlotUnits * 8);
                mbar_wait(bar_recv_full(rs2), par2);
                // re-arm the slot for its next use (bytes can only arrive after the credits below)
                if (lane == 0) mbar_arrive_expect_tx(bar_recv_full(rs2), kFrameBytes);
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    va[p] = lds64(p2_colA + ro + (uint32_t)(kRecvStride * p) * 8);
                    vb[p] = lds64(p2_colB + ro + (uint32_t)(kRecvStride * p) * 8);
                }
                RegDft<8, +1>::run(va);
                RegDft<8, +1>::run(vb);
                twiddle8(va, tw_p2A);
                twiddle8(vb, tw_p2B);
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    sts64(p2_colA + ro + (uint32_t)(kRecvStride * r) * 8, va[r]);
                    sts64(p2_colB + ro + (uint32_t)(kRecvStride * r) * 8, vb[r]);
                }
                __syncwarp();
#pragma unroll
                for (int qq = 0; qq < 8; qq++) {
                    va[qq] = lds64(p2_rowA + ro + (uint32_t)recv_pos(qq) * 8);
                    vb[qq] = lds64(p2_rowB + ro + (uint32_t)recv_pos(qq) * 8);
                }
                RegDft<8, +1>::run(va);
                RegDft<8, +1>::run(vb);
                __syncwarp();
                // hand the receive slot back to the 8 senders (every value read from it has been consumed by
                // the DFTs above): lane s credits CTA s
                if (lane < kCluster) mbar_arrive_remote(rbase + (uint32_t)lane * rstep + (uint32_t)kOffBarRecvFree + 8u * rs2);
                float2 *yf = y_thr + (long long)it * (kGroup * kM);
#pragma unroll
                for (int s2 = 0; s2 < 8; s2++) {
                    __stcs(yf + 64 * qpA + 512 * s2, va[s2]);
                    __stcs(yf + 64 * qpB + 512 * s2, vb[s2]);
                }
                if (++g2 == kRecvGroups) {
                    g2 = 0;
                    par2 ^= 1u;
                }
            }
        };
        using T = std::true_type;
        using F = std::false_type;
        int it = 0;
        for (; it < kLag && it < n_groups; it++) body(T{}, F{}, it);     // phase 1 runs for it < n_groups,
        for (; it < n_groups; it++) body(T{}, T{}, it);                  // phase 2 for it >= kLag
        for (; it < n_groups + kLag; it++)
            if (it >= kLag) body(F{}, T{}, it);
    }
    // no CTA may leave while its peers can still store into it or credit its barriers
    cluster_sync_all();
}

int g_max_clusters_dev[64];      // co-resident clusters of 8 per device, stored as n + 1 (0 = not measured yet)

template <int L>
long long run_L(ClusterArgs &a, long long n_groups, int par0, cudaStream_t st)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured)) {
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_cluster_4096_kernel<L, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<L>::kBytes));
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_cluster_4096_kernel<L, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<L>::kBytes));
    }
    int cur_dev = 0;
    LC_CUDA_CHECK(cudaGetDevice(&cur_dev));
    int &max_slot = g_max_clusters_dev[cur_dev & 63];
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = Smem<L>::kBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_slot == 0) {
        int sms = 148, n = 0;
        LC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev));
        cfg.gridDim = dim3((unsigned int)(sms / kCluster) * kCluster, 1, 1);
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, pfbch2_cluster_4096_kernel<L, 0>, &cfg);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            n = 0;
        }
        if (const char *env = getenv("LIQUID_CUDA_MAX_CLUSTERS")) {
            int lim = atoi(env);
            if (lim >= 0 && lim < n) n = lim;
        }
        max_slot = n + 1;
        if (getenv("LIQUID_CUDA_VERBOSE"))
            fprintf(stderr, "liquid_cuda: pfbch2 cluster kernel: %d co-resident clusters of %d CTAs on %d SMs\n", n, kCluster, sms);
    }
    const int g_max_clusters = max_slot - 1;
    if (g_max_clusters == 0) return 0;                     // clusters of 8 cannot be scheduled here
    // at least 4 groups (16 frames) per cluster so the history prologue and the pipeline fill amortise
    long long cl = (n_groups + 3) / 4;
    if (cl > g_max_clusters) cl = g_max_clusters;
    const long long per = (n_groups + cl - 1) / cl;
    cl = (n_groups + per - 1) / per;
    a.n_groups = (int)n_groups;
    a.groups_per_cluster = (int)per;
    lc::g_last_chunk_frames.store((unsigned long long)per * kGroup);
    cfg.gridDim = dim3((unsigned int)cl * kCluster, 1, 1);
    if (par0) LC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, pfbch2_cluster_4096_kernel<L, 1>, a));
    else      LC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, pfbch2_cluster_4096_kernel<L, 0>, a));
    LC_LAUNCH_CHECK();
    return n_groups * kGroup;
}

}  // namespace

namespace lc {
namespace pfb {

bool cluster_pfbch2_supported(unsigned int M, unsigned int L)
{
    return M == 4096 && (L == 4 || L == 6 || L == 8 || L == 10 || L == 12 || L == 14 || L == 16 || L == 18);
}

void cluster_pfbch2_make_twiddles(std::vector<float2> &tw)
{
    // tw8[r][q] = e^{+j 2 pi q r / 64}; then tw4096[n0][k1] = e^{+j 2 pi n0 k1 / 4096}
    tw.resize(64 + 64 * 64);
    for (int r = 0; r < 8; r++)
        for (int q = 0; q < 8; q++) {
            double ang = 2.0 * M_PI * (double)(q * r) / 64.0;
            tw[r * 8 + q] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (int n0 = 0; n0 < 64; n0++)
        for (int k1 = 0; k1 < 64; k1++) {
            double ang = 2.0 * M_PI * (double)(n0 * k1) / 4096.0;
            tw[64 + n0 * 64 + k1] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

int cluster_pfbch2_max_clusters()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    return g_max_clusters_dev[dev & 63] - 1;
}

long long launch_pfbch2_cluster(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps_scaled,
                                const float2 *d_tw, float2 *d_y, unsigned int M, unsigned int L,
                                unsigned long long n_frames, int parity0, cudaStream_t st)
{
    (void)M;
    long long n_groups = (long long)(n_frames / kGroup);
    if (n_groups == 0) return 0;
    if (n_groups > (1ll << 28)) n_groups = 1ll << 28;      // the caller finishes the rest
    ClusterArgs a;
    a.hist = d_hist;
    a.x = d_x;
    a.y = d_y;
    a.taps = d_taps_scaled;
    a.tw = d_tw;
    a.H = H;
    switch (L) {
    case 4: return run_L<4>(a, n_groups, parity0 & 1, st);
    case 6: return run_L<6>(a, n_groups, parity0 & 1, st);
    case 8: return run_L<8>(a, n_groups, parity0 & 1, st);
    case 10: return run_L<10>(a, n_groups, parity0 & 1, st);
    case 12: return run_L<12>(a, n_groups, parity0 & 1, st);
    case 14: return run_L<14>(a, n_groups, parity0 & 1, st);
    case 16: return run_L<16>(a, n_groups, parity0 & 1, st);
    case 18: return run_L<18>(a, n_groups, parity0 & 1, st);
    default: fatal("launch_pfbch2_cluster", "unsupported tap count");
    }
}

}  // namespace pfb
}  // namespace lc
