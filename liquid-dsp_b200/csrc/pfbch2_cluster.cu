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
// Warp roles per CTA (512 threads, registers moved with setmaxnreg 152 / 104):
//    8 MAC warps   thread = pair (n1, n0): 2 x 2m-sample histories + 2 x 2m taps in registers, per frame
//                  one 8-byte sample (own cp.async ring slot, no barrier needed), 4m packed FFMA2,
//                  two STS into the frame's X slot.
//    8 FFT warps   warp pair fp handles frames = fp mod 4: phase 1 of frame f, then phase 2 of the frame
//                  two groups earlier (the exchange latency hides behind a group of work).
//
// Reference semantics: src/multichannel/src/firpfbch2.c:244-282 (see firpfbch2.cu).  The 1/M output
// scale is folded into the taps (exact: M is a power of two).
#include "pfbch2_cluster.cuh"
#include "fused_common.cuh"

#include <cmath>
#include <cstdint>

namespace {

using namespace lcf;

constexpr int kM = 4096;
constexpr int kM2 = 2048;
constexpr int kCluster = 8;
constexpr int kMacWarps = 8;
constexpr int kFftWarps = 8;
constexpr int kMacThreads = 32 * kMacWarps;                    // 256
constexpr int kThreads = 32 * (kMacWarps + kFftWarps);         // 512
constexpr int kGroup = 4;                                      // frames per group (one per FFT warp pair)
constexpr int kXSlots = 8;                                     // X ring depth in frames (power of two)
constexpr int kXStride = 74;                                   // units (float2) per n0 block of an X slot: 8 rows of 9 (+2)
constexpr int kXSlotUnits = 8 * kXStride;                      // 592
constexpr int kRecvGroups = 4;                                 // receive ring depth in groups (power of two)
constexpr int kLag = 2;                                        // groups phase 2 trails phase 1 (< kRecvGroups)
constexpr int kRecvStride = 72;                                // units per source CTA of a receive slot: 64 (+8)
constexpr int kRecvSlotUnits = 8 * kRecvStride;                // 576
constexpr int kRecvSlots = kRecvGroups * kGroup;               // 16
constexpr int kInDepth = 8;                                    // input ring depth in double frames (power of two)
constexpr int kFrameBytes = 64 * 64 * 8 / kCluster;            // bytes of Z one CTA receives per frame: 4096

// shared memory carve-up (bytes)
constexpr int kOffTw64 = 0;                                    // float2 tw64[8][8]: W64^(q r)
constexpr int kOffX = kOffTw64 + 64 * 8;
constexpr int kOffRecv = kOffX + kXSlots * kXSlotUnits * 8;
constexpr int kOffIn = kOffRecv + kRecvSlots * kRecvSlotUnits * 8;
constexpr int kOffBar = kOffIn + kInDepth * 2 * kMacThreads * 8;
constexpr int kNumBars = 2 * kXSlots + 2 * kRecvSlots;
constexpr int kSmemBytes = kOffBar + kNumBars * 8;

struct ClusterArgs {
    const float2 *hist;        // H samples preceding x[0] (never null)
    const float2 *x;           // segment input
    float2 *y;                 // [n_frames][4096]
    const float *taps;         // [2][L][4096] as in firpfbch2.cu, pre-scaled by 1/M
    const float2 *tw;          // tw64[8][8] followed by tw4096[64][64] (n0, k1)
    int H;
    int n_groups;              // groups of 4 frames in the segment
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

template <int L, int PAR0>
__global__ void __launch_bounds__(kThreads, 1) pfbch2_cluster_4096_kernel(const ClusterArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_tw64 = s_base + kOffTw64;
    const uint32_t s_x = s_base + kOffX;
    const uint32_t s_recv = s_base + kOffRecv;
    const uint32_t s_in = s_base + kOffIn;
    const uint32_t s_bar = s_base + kOffBar;
    // barriers: X ring full[slot] / empty[slot]; receive ring recv_full[slot] (bytes) / recv_free[slot] (credits)
    auto bar_x_full = [&](int s) { return s_bar + 8u * s; };
    auto bar_x_empty = [&](int s) { return s_bar + 8u * (kXSlots + s); };
    auto bar_recv_full = [&](int s) { return s_bar + 8u * (2 * kXSlots + s); };
    auto bar_recv_free = [&](int s) { return s_bar + 8u * (2 * kXSlots + kRecvSlots + s); };

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
    if (tid < 64) reinterpret_cast<float2 *>(smem + kOffTw64)[tid] = a.tw[tid];
    if (tid == 0) {
        for (int s = 0; s < kXSlots; s++) {
            mbar_init(bar_x_full(s), kMacWarps);
            mbar_init(bar_x_empty(s), 2);                  // the two FFT warps of the frame
        }
        for (int s = 0; s < kRecvSlots; s++) {
            mbar_init(bar_recv_full(s), 1);
            mbar_init(bar_recv_free(s), 2 * kCluster);     // two warps of each of the 8 receivers
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // arm the first use of every receive slot: one frame's worth of bytes from the 8 senders
        for (int s = 0; s < kRecvSlots; s++) mbar_arrive_expect_tx(bar_recv_full(s), kFrameBytes);
    }
    __syncthreads();
    cluster_sync_all();                                    // every CTA's barriers exist before any remote access

    if (warp < kMacWarps) {
        // =========================================================================== MAC warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const int n1 = tid >> 3, n0l = tid & 7;             // a warp covers four 64-byte runs of the stream
        const int j = 8 * crank + n0l + 64 * n1;            // pair (j, j + 2048)
        PairState<L> s;
        pair_load<L, kM>(s, a.taps, a.hist, a.x, a.H, j, f_begin);
        const float2 *src = a.x + f_begin * kM2 + (kM2 - 1 - j);         // this pair's sample of local frame 0
        const uint32_t in_thr = s_in + (uint32_t)(tid ^ 7) * 8;           // stream order inside a 64-byte run; + (2 * dslot + odd) * 2048
        const uint32_t x_thr = s_x + (uint32_t)(n0l * kXStride + n1 + (n1 >> 3)) * 8;   // X[n0][n1]; X[n0][n1+32] is 36 units on
        const int n_dbl = n_frames / 2;

        auto fetch_dbl = [&](int u) {
            if (u < n_dbl) {
                const uint32_t dst = in_thr + (uint32_t)(u & (kInDepth - 1)) * (2 * kMacThreads * 8);
                cp_async8(dst, src + (long long)(2 * u) * kM2);
                cp_async8(dst + kMacThreads * 8, src + (long long)(2 * u + 1) * kM2);
            }
            cp_async_commit();                              // always: keeps the group count uniform
        };
        for (int u = 0; u < kInDepth - 1; u++) fetch_dbl(u);

        const int n_periods = (n_dbl + L - 1) / L;          // the last one may be partial
        for (int p = 0; p < n_periods; p++) {
            // one period: L updates of each window = 2L frames, fully unrolled so that every history slot
            // index below is a compile-time constant
#pragma unroll
            for (int uu = 0; uu < L; uu++) {
                const int u = p * L + uu;                   // double frame inside this cluster's range
                if (u >= n_dbl) continue;                   // warp-uniform: tail of the last period
                cp_async_wait<kInDepth - 2>();              // double frame u has landed (own slots only)
                const uint32_t in_u = in_thr + (uint32_t)(u & (kInDepth - 1)) * (2 * kMacThreads * 8);
                const float2 xe = lds64(in_u), xo = lds64(in_u + kMacThreads * 8);
                fetch_dbl(u + kInDepth - 1);                // refills the slot read one iteration ago
                float2 o1, o2;

                // ---- even local frame: the window pushed on even frames gets the new sample
                {
                    const int f = 2 * u, xs = f & (kXSlots - 1);
                    s.dA[uu] = xe;
                    pair_mac<L, 0>(s, uu, o1, o2);
                    mbar_wait(bar_x_empty(xs), (((uint32_t)f / kXSlots) & 1u) ^ 1u);
                    const uint32_t xa = x_thr + (uint32_t)xs * (kXSlotUnits * 8);
                    // o1 is window j when the segment starts on an even frame (PAR0 = 0), window j + 2048 otherwise
                    sts64(xa, PAR0 == 0 ? o1 : o2);
                    sts64(xa + 36 * 8, PAR0 == 0 ? o2 : o1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_x_full(xs));
                }
                // ---- odd local frame: the other window of the pair gets the new sample, branches swap
                {
                    const int f = 2 * u + 1, xs = f & (kXSlots - 1);
                    s.dB[uu] = xo;
                    pair_mac<L, 1>(s, uu, o1, o2);
                    mbar_wait(bar_x_empty(xs), (((uint32_t)f / kXSlots) & 1u) ^ 1u);
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
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
        const int fw = warp - kMacWarps;                    // 0..7
        const int fp = fw >> 1;                             // frame-in-group handled by this warp pair
        const int h = fw & 1;
        const int i = lane >> 3, q = lane & 7;              // phase 1: column q in pass 1, row r = q in pass 2
        // ---- phase 1 addressing: the 64-point DFT of n0 = 8c + n0l lives in 8 consecutive lanes
        const int n0l = n0_of(h, i);
        const uint32_t p1_col = (uint32_t)(n0l * kXStride + q) * 8;          // cell (p, q): + 9 p units
        const uint32_t p1_row = (uint32_t)(n0l * kXStride + 9 * q) * 8;      // cell (r = q, qq): + qq units
        float2 tw_p1[8];                                                     // W64^(q r), r = 1..7
#pragma unroll
        for (int r = 1; r < 8; r++) tw_p1[r] = lds64(s_tw64 + (uint32_t)(q * 8 + r) * 8);
        float2 tw_z[8];                                                      // W4096^(n0 (q + 8 s)), s = 0..7
#pragma unroll
        for (int s = 0; s < 8; s++) tw_z[s] = __ldg(a.tw + 64 + (8 * crank + n0l) * 64 + q + 8 * s);
        // exchange: this thread's Z[n0][q + 8 s] goes to CTA s, position (h, lane) of source block `crank`
        const uint32_t send_off = (uint32_t)kOffRecv + (uint32_t)(crank * kRecvStride + h * 32 + lane) * 8;
        const uint32_t rbase = map_to_cta(s_base, 0);                         // CTA windows are equally spaced
        const uint32_t rstep = map_to_cta(s_base, 1) - rbase;
        // ---- phase 2 addressing: k1 = 8c + r with r = lane % 8; column q' in pass 1, row r2 = q' in pass 2
        const int z = lane >> 3;
        const int qp = (h == 0) ? ((z == 0) ? 0 : (z == 1) ? 5 : (z == 2) ? 1 : 4)
                                : ((z == 0) ? 2 : (z == 1) ? 7 : (z == 2) ? 3 : 6);
        const uint32_t p2_col = (uint32_t)(recv_pos(qp) + q) * 8;            // cell (p', q'): + 72 p' units
        const uint32_t p2_row = (uint32_t)(qp * kRecvStride + q) * 8;        // cell (r2 = q', q''): + recv_pos(q'') units
        float2 tw_p2[8];                                                     // W64^(q' r2), r2 = 1..7
#pragma unroll
        for (int r = 1; r < 8; r++) tw_p2[r] = lds64(s_tw64 + (uint32_t)(qp * 8 + r) * 8);
        float2 *y_thr = a.y + f_begin * (long long)kM + (8 * crank + q) + 64 * qp;   // + frame * 4096 + 512 s2

        for (int it = 0; it < n_groups + kLag; it++) {
            if (it < n_groups) {
                // ------------------------------------------------ phase 1 of frame f
                const int f = it * kGroup + fp;
                const int xs = f & (kXSlots - 1);
                const uint32_t slot = s_x + (uint32_t)xs * (kXSlotUnits * 8);
                mbar_wait(bar_x_full(xs), ((uint32_t)f / kXSlots) & 1u);
                float2 v[8];
#pragma unroll
                for (int p = 0; p < 8; p++) v[p] = lds64(slot + p1_col + (uint32_t)(9 * p) * 8);
                RegDft<8, +1>::run(v);
#pragma unroll
                for (int r = 1; r < 8; r++) v[r] = cmul2(v[r], tw_p1[r]);
                // in-place transposition: this thread rewrites exactly the cells it read
#pragma unroll
                for (int r = 0; r < 8; r++) sts64(slot + p1_col + (uint32_t)(9 * r) * 8, v[r]);
                __syncwarp();
#pragma unroll
                for (int qq = 0; qq < 8; qq++) v[qq] = lds64(slot + p1_row + (uint32_t)qq * 8);
                __syncwarp();                               // the warp's four DFTs have left the slot
                if (lane == 0) mbar_arrive(bar_x_empty(xs));
                RegDft<8, +1>::run(v);
                // receive slot of frame f in every CTA; wait until all 8 receivers have handed it back
                const int rs = (it & (kRecvGroups - 1)) * kGroup + fp;
                const uint32_t use = (uint32_t)it / kRecvGroups;
                if (use > 0) mbar_wait(bar_recv_free(rs), (use - 1u) & 1u);
                const uint32_t dst = send_off + (uint32_t)rs * (kRecvSlotUnits * 8);
                const uint32_t dbar = (uint32_t)kOffBar + 8u * (2 * kXSlots + rs);
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    const uint32_t rcta = rbase + (uint32_t)s * rstep;
                    st_async64(rcta + dst, cmul2(v[s], tw_z[s]), rcta + dbar);
                }
            }
            if (it >= kLag) {
                // ------------------------------------------------ phase 2 of the frame two groups back
                const int g = it - kLag;
                const int rs = (g & (kRecvGroups - 1)) * kGroup + fp;
                const uint32_t use = (uint32_t)g / kRecvGroups;
                const uint32_t slot = s_recv + (uint32_t)rs * (kRecvSlotUnits * 8);
                mbar_wait(bar_recv_full(rs), use & 1u);
                // re-arm the slot for its next use (bytes can only arrive after the credits below)
                if (h == 0 && lane == 0) mbar_arrive_expect_tx(bar_recv_full(rs), kFrameBytes);
                float2 v[8];
#pragma unroll
                for (int p = 0; p < 8; p++) v[p] = lds64(slot + p2_col + (uint32_t)(kRecvStride * p) * 8);
                RegDft<8, +1>::run(v);
#pragma unroll
                for (int r = 1; r < 8; r++) v[r] = cmul2(v[r], tw_p2[r]);
#pragma unroll
                for (int r = 0; r < 8; r++) sts64(slot + p2_col + (uint32_t)(kRecvStride * r) * 8, v[r]);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + fp) : "memory");   // the DFT of one k1 spans both warps
#pragma unroll
                for (int qq = 0; qq < 8; qq++) v[qq] = lds64(slot + p2_row + (uint32_t)recv_pos(qq) * 8);
                RegDft<8, +1>::run(v);
                __syncwarp();
                // hand the slot back to the 8 senders (every value read from it has been consumed by the
                // DFT above): lane s credits CTA s
                if (lane < kCluster)
                    mbar_arrive_remote(rbase + (uint32_t)lane * rstep + (uint32_t)kOffBar + 8u * (2 * kXSlots + kRecvSlots + rs));
                float2 *yf = y_thr + (long long)(g * kGroup + fp) * kM;
#pragma unroll
                for (int s2 = 0; s2 < 8; s2++) __stcs(yf + 512 * s2, v[s2]);
            }
        }
    }
    // no CTA may leave while its peers can still store into it or credit its barriers
    cluster_sync_all();
}

int g_max_clusters = -1;         // co-resident clusters of 8 on this device (measured on first use)

template <int L>
long long run_L(ClusterArgs &a, long long n_groups, int par0, cudaStream_t st)
{
    static bool configured = false;
    if (!configured) {
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_cluster_4096_kernel<L, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_cluster_4096_kernel<L, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (g_max_clusters < 0) {
        int dev = 0, sms = 148, n = 0;
        LC_CUDA_CHECK(cudaGetDevice(&dev));
        LC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
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
        g_max_clusters = n;
        if (getenv("LIQUID_CUDA_VERBOSE"))
            fprintf(stderr, "liquid_cuda: pfbch2 cluster kernel: %d co-resident clusters of %d CTAs on %d SMs\n", n, kCluster, sms);
    }
    if (g_max_clusters == 0) return 0;                     // clusters of 8 cannot be scheduled here
    // at least 4 groups (16 frames) per cluster so the history prologue and the pipeline fill amortise
    long long cl = (n_groups + 3) / 4;
    if (cl > g_max_clusters) cl = g_max_clusters;
    const long long per = (n_groups + cl - 1) / cl;
    cl = (n_groups + per - 1) / per;
    a.n_groups = (int)n_groups;
    a.groups_per_cluster = (int)per;
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
    return M == 4096 && (L == 8 || L == 14 || L == 18);
}

void cluster_pfbch2_make_twiddles(std::vector<float2> &tw)
{
    // tw64[q][r] = e^{+j 2 pi q r / 64}; then tw4096[n0][k1] = e^{+j 2 pi n0 k1 / 4096}
    tw.resize(64 + 64 * 64);
    for (int q = 0; q < 8; q++)
        for (int r = 0; r < 8; r++) {
            double ang = 2.0 * M_PI * (double)(q * r) / 64.0;
            tw[q * 8 + r] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (int n0 = 0; n0 < 64; n0++)
        for (int k1 = 0; k1 < 64; k1++) {
            double ang = 2.0 * M_PI * (double)(n0 * k1) / 4096.0;
            tw[64 + n0 * 64 + k1] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

int cluster_pfbch2_max_clusters() { return g_max_clusters; }

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
    case 8: return run_L<8>(a, n_groups, parity0 & 1, st);
    case 14: return run_L<14>(a, n_groups, parity0 & 1, st);
    case 18: return run_L<18>(a, n_groups, parity0 & 1, st);
    default: fatal("launch_pfbch2_cluster", "unsupported tap count");
    }
}

}  // namespace pfb
}  // namespace lc
