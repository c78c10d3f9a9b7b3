// pfbch_fused.cu -- fused firpfbch_crcf analysis kernels for M = 16 and M = 256 (sm_100a).
//
// The critically sampled analyzer is HBM-bound (16 B per sample against ~50-80 flops), so the job of
// these kernels is to read every sample once, write every output once, and keep many bytes in flight:
//
//   phase 1  thread u owns window u (transform input u): the p taps of its branch and a register
//            window of p-1+16 samples; it loads its 16 samples of a 16-frame tile (one coalesced row
//            per frame), computes the 16 dot products with packed FFMA2 (tap broadcast x {re, im}) and
//            stores them into a shared tile; the last p-1 samples are carried to the next tile.
//   phase 2  M = 16 : each thread takes one whole frame of the tile (LDS.128 row), does a 16-point
//                     register DFT, writes the row back; the 16x16 tile then leaves as 256 contiguous
//                     complex outputs with 128-bit coalesced stores.  A 16-thread group (half a warp)
//                     is self-contained, so there is no CTA barrier at all; 16 groups per CTA walk 16
//                     consecutive sub-chunks of the stream.
//            M = 256: 16 threads per frame, 256 = 16 x 16, two passes of 16-point register DFTs with a
//                     transposition through the padded tile and inter-pass twiddles (as the 256-point
//                     sub-problems of pfbch2_fused.cu, forward sign); outputs leave as 128-byte rows.
//
// Reference semantics: src/multichannel/src/firpfbch.c:346-409 (push M samples, run with k = 0):
//   X_s[u] = sum_n h[(M-1-u) + nM] x[(s-n)M + u],   Y_s = DFT_forward_unscaled(X_s).
#include "pfbch_fused.cuh"
#include "pfbch2_tile.cuh"
#include "pfbch2_small.cuh"
#include "fused_common.cuh"

#include <cmath>

namespace {

using namespace lcf;

constexpr int kThreads = 256;
constexpr int kTile = 16;                      // frames per tile

struct PfbArgs {
    const float2 *hist;        // H samples preceding x[0] of the (single) stream, or null = zero state
    const float2 *x;           // [n_streams][stream_stride]
    float2 *y;                 // [n_streams][stream_stride]
    const float *taps;         // [P][M]: taps[n][i] = h[i + nM]
    const float2 *tw;          // M = 256 only: tw[k1][r2] = e^{-j 2 pi k1 r2 / 256}
    int H;
    int n_tiles;               // 16-frame tiles per stream
    int tiles_per_inst;        // consecutive tiles walked by one instance
    long long stream_stride;   // samples between streams
};

__device__ __forceinline__ float2 ld_hist0(const float2 *__restrict__ hist, const float2 *__restrict__ x,
                                           long long idx, int H)
{
    if (idx >= 0) return __ldg(x + idx);
    if (hist != nullptr && H + idx >= 0) return __ldg(hist + (H + idx));
    return make_float2(0.f, 0.f);
}

// phase 1 for one tile: win[0 .. P-2] carried history (oldest first), win[P-1+f] = sample of frame f
template <int M, int P>
__device__ __forceinline__ void mac_tile(const float (&h)[P], float2 (&win)[P - 1 + kTile], float2 (&out)[kTile])
{
#pragma unroll
    for (int f = 0; f < kTile; f++) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int n = P - 1; n >= 0; n--) acc = __ffma2_rn(bc(h[n]), win[P - 1 + f - n], acc);
        out[f] = acc;
    }
}

// ------------------------------------------------------------------------------------------ M = 16
template <int P>
__global__ void __launch_bounds__(kThreads) pfbch_fused_16_kernel(const PfbArgs a)
{
    constexpr int M = 16;
    constexpr int ROW = 18;                               // units per tile row: 144 B, keeps LDS.128 aligned
    __shared__ __align__(16) float2 tile[16][kTile * ROW]; // one 16x16 tile per 16-thread group

    const int q = threadIdx.x >> 4, u = threadIdx.x & 15;
    const long long inst = (long long)blockIdx.x * 16 + q;            // instance within the stream
    const int t0 = (int)(inst * a.tiles_per_inst);
    int nt = a.n_tiles - t0;
    if (nt > a.tiles_per_inst) nt = a.tiles_per_inst;
    const float2 *x = a.x + (long long)blockIdx.y * a.stream_stride;
    float2 *y = a.y + (long long)blockIdx.y * a.stream_stride;
    float2 *tl = tile[q];

    float h[P];
#pragma unroll
    for (int n = 0; n < P; n++) h[n] = __ldg(a.taps + n * M + (M - 1 - u));   // branch M-1-u feeds input u
    float2 win[P - 1 + kTile];
    if (nt > 0) {
        const long long s0 = (long long)t0 * kTile;
#pragma unroll
        for (int i = 0; i < P - 1; i++) win[i] = ld_hist0(a.hist, x, (s0 - (P - 1) + i) * M + u, a.H);
    }
    for (int it = 0; it < nt; it++) {
        const long long s0 = (long long)(t0 + it) * kTile;            // first frame of the tile
        // ---- phase 1
#pragma unroll
        for (int f = 0; f < kTile; f++) win[P - 1 + f] = __ldg(x + (s0 + f) * M + u);
        float2 out[kTile];
        mac_tile<M, P>(h, win, out);
#pragma unroll
        for (int f = 0; f < kTile; f++) tl[f * ROW + u] = out[f];
#pragma unroll
        for (int i = 0; i < P - 1; i++) win[i] = win[kTile + i];
        __syncwarp();
        // ---- phase 2: this thread transforms frame u of the tile
        float2 v[16];
        {
            const float4 *row = reinterpret_cast<const float4 *>(tl + u * ROW);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float4 t = row[i];
                v[2 * i] = make_float2(t.x, t.y);
                v[2 * i + 1] = make_float2(t.z, t.w);
            }
        }
        RegDft<16, -1>::run(v);
        {
            float4 *row = reinterpret_cast<float4 *>(tl + u * ROW);
#pragma unroll
            for (int i = 0; i < 8; i++) row[i] = make_float4(v[2 * i].x, v[2 * i].y, v[2 * i + 1].x, v[2 * i + 1].y);
        }
        __syncwarp();
        // ---- phase 3: the tile is 256 contiguous outputs; float4 index w = i*16 + u
        float4 *yt = reinterpret_cast<float4 *>(y + s0 * M);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = 2 * i + (u >> 3), c = (2 * u) & 15;
            __stcs(yt + i * 16 + u, *reinterpret_cast<const float4 *>(tl + r * ROW + c));
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------ M = 256
template <int P>
__global__ void __launch_bounds__(kThreads, P <= 4 ? 3 : 2) pfbch_fused_256_kernel(const PfbArgs a)
{
    constexpr int M = 256;
    constexpr int FR = 16 * 17;                           // units per frame: 16 rows of 16 padded to 17
    __shared__ float2 tile[kTile * FR];

    const int u = threadIdx.x;
    const int t0 = (int)((long long)blockIdx.x * a.tiles_per_inst);
    int nt = a.n_tiles - t0;
    if (nt > a.tiles_per_inst) nt = a.tiles_per_inst;
    const float2 *x = a.x + (long long)blockIdx.y * a.stream_stride;
    float2 *y = a.y + (long long)blockIdx.y * a.stream_stride;

    float h[P];
#pragma unroll
    for (int n = 0; n < P; n++) h[n] = __ldg(a.taps + n * M + (M - 1 - u));
    float2 win[P - 1 + kTile];
    {
        const long long s0 = (long long)t0 * kTile;
#pragma unroll
        for (int i = 0; i < P - 1; i++) win[i] = ld_hist0(a.hist, x, (s0 - (P - 1) + i) * M + u, a.H);
    }
    // phase 2 roles: 16 threads per frame
    const int f2 = u >> 4, r2 = u & 15;
    const float2 *twp = a.tw + r2;                         // tw[k1][r2] at + 16 k1
    // short filters: the next tile's 16 samples are loaded while phase 2 of the current tile runs (registers, one tile
    // ahead): 0.90 -> 0.97 of the roofline at p = 4; at p = 10 the same change measured 0.86 -> 0.80, so long filters load in place
    constexpr bool REGPF = true;
    float2 nx[kTile];
    auto fetch = [&](int it) {
        const long long s0 = (long long)(t0 + it) * kTile;
#pragma unroll
        for (int f = 0; f < kTile; f++) nx[f] = it < nt ? __ldg(x + (s0 + f) * M + u) : make_float2(0.f, 0.f);
    };
    if (REGPF) fetch(0);
    for (int it = 0; it < nt; it++) {                      // nt is CTA-uniform
        const long long s0 = (long long)(t0 + it) * kTile;
        // ---- phase 1
#pragma unroll
        for (int f = 0; f < kTile; f++) win[P - 1 + f] = REGPF ? nx[f] : __ldg(x + (s0 + f) * M + u);
        float2 out[kTile];
        mac_tile<M, P>(h, win, out);
#pragma unroll
        for (int f = 0; f < kTile; f++) tile[f * FR + u + (u >> 4)] = out[f];
#pragma unroll
        for (int i = 0; i < P - 1; i++) win[i] = win[kTile + i];
        if (REGPF) fetch(it + 1);
        __syncthreads();
        // ---- phase 2: frame f2, pass A down column r2, twiddle, transpose, pass B along row r2
        float2 *fr = tile + f2 * FR;
        float2 v[16];
#pragma unroll
        for (int r1 = 0; r1 < 16; r1++) v[r1] = fr[r1 * 17 + r2];
        RegDft<16, -1>::run(v);
#pragma unroll
        for (int k1 = 1; k1 < 16; k1++) v[k1] = cmul2_rot(v[k1], __ldg(twp + 16 * k1));
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) fr[r2 * 17 + k1] = v[k1];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 16; c++) v[c] = fr[c * 17 + r2];   // lane r2 now plays k1
        RegDft<16, -1>::run(v);
        float2 *yf = y + (s0 + f2) * M + r2;
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) __stcs(yf + 16 * k2, v[k2]);
        __syncthreads();                                   // tile free for the next phase 1
    }
}

// ------------------------------------------------------------------------------------------ M = 1024
// The reference's largest benchmark shape (firpfbch_crcf_benchmark.c:85).  One instance = one CTA of 256 threads.
//   phase 1  thread u owns the FOUR windows u + 256 q: four samples per frame, four P-tap dot products over sliding
//            register windows, then -- the four results are exactly the inputs of one radix-4 decimation-in-frequency
//            butterfly of the 1024-point forward DFT -- that butterfly and its three twiddles W1024^(u p): a frame
//            leaves phase 1 as four independent 256-point problems z_p[u] in the shared tile (4 frames per tile).
//   phase 2  64 threads per frame, lane = (p fastest, column): 256 = 16 x 16, a 16-point register DFT down the
//            column (in place), the inter-pass twiddle from a shared copy, a 64-thread named barrier, a 16-point
//            DFT along the row, and Y[4 (k1 + 16 k2) + p] leaves as 256-byte runs.  Sub-problems are 276 units
//            apart (16 rows of 17, + 4), so every access pattern is bank-conflict-free (as in pfbch2_fused.cu).
//   Two tile buffers leave one block barrier per tile; the next tile's 16 samples are prefetched into registers
//   (P = 4) or into the thread's own cp.async slot (longer filters).
constexpr int kQFr = 4;                                    // frames per tile
constexpr int kQSub = 276;                                 // units per 256-point sub-problem
constexpr int kQFrame = 4 * kQSub;                         // units per frame

struct QuadArgs {
    const float2 *hist;
    const float2 *x;
    float2 *y;
    const float *taps;         // [P][1024]
    const float2 *tw;          // tw4[p-1][u] = e^{-j 2 pi u p / 1024}, p = 1..3, u < 256; then tw256[k1][r2] = e^{-j 2 pi k1 r2 / 256}
    long long n_frames;
    long long stream_stride;
    int H;
    int n_tiles;
    int tiles_per_inst;
};

template <int P> constexpr bool kQuadRegPrefetch = (P <= 4);
template <int P>
constexpr size_t quad_smem_bytes()
{
    return sizeof(float2) * ((kQuadRegPrefetch<P> ? (size_t)0 : (size_t)4 * kQFr * kThreads) + (size_t)2 * kQFr * kQFrame + 16 * 18);
}

template <int P>
__global__ void __launch_bounds__(kThreads, P <= 4 ? 2 : 1) pfbch_quad_1024_kernel(const QuadArgs a)
{
    constexpr int M = 1024, HL = P - 1;
    constexpr bool REGPF = kQuadRegPrefetch<P>;
    constexpr int kSlotUnits = REGPF ? 0 : 4 * kQFr * kThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *slot = reinterpret_cast<float2 *>(smem_raw);              // [4 kQFr][kThreads]
    float2 *tile0 = slot + kSlotUnits;
    float2 *tws = tile0 + 2 * kQFr * kQFrame;                         // [r2][k1], pitch 18

    const int u = threadIdx.x;
    const long long t0 = (long long)blockIdx.x * a.tiles_per_inst;
    long long ntl = (long long)a.n_tiles - t0;
    if (ntl > a.tiles_per_inst) ntl = a.tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;
    const float2 *x = a.x + (long long)blockIdx.y * a.stream_stride;
    float2 *y = a.y + (long long)blockIdx.y * a.stream_stride;

    float2 nx[4 * kQFr];
    const uint32_t slot_u32 = smem_u32(slot + u);
    auto fetch = [&](int it) {
        const long long f0 = (t0 + it) * kQFr;
#pragma unroll
        for (int f = 0; f < kQFr; f++) {
            const bool ok = it < nt && f0 + f < a.n_frames;
            const float2 *src = x + (f0 + f) * M + u;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (REGPF) nx[4 * f + q] = ok ? __ldg(src + 256 * q) : make_float2(0.f, 0.f);
                else {
                    const uint32_t n = ok ? 8u : 0u;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(slot_u32 + (4 * f + q) * kThreads * 8),
                                 "l"(ok ? (const void *)(src + 256 * q) : (const void *)a.x), "r"(n) : "memory");
                }
            }
        }
        if (!REGPF) asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(0);
    tws[(u & 15) * 18 + (u >> 4)] = __ldg(a.tw + 768 + u);            // tw256[k1 = u >> 4][r2 = u & 15]
    const float2 t1 = __ldg(a.tw + u), t2 = __ldg(a.tw + 256 + u), t3 = __ldg(a.tw + 512 + u);
    float h[4][P];
    float2 w[4][HL + kQFr];
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
        for (int n = 0; n < P; n++) h[q][n] = __ldg(a.taps + n * M + (M - 1 - (u + 256 * q)));
        if (nt > 0) {
            const long long f0 = t0 * kQFr;
#pragma unroll
            for (int i = 0; i < HL; i++) w[q][i] = ld_hist0(a.hist, x, (f0 - HL + i) * M + u + 256 * q, a.H);
        }
    }
    // phase 2 role: frame f2, sub-problem pp (fastest), column / row c
    const int f2 = u >> 6, pp = u & 3, c = (u >> 2) & 15;
    const float4 *twc = reinterpret_cast<const float4 *>(tws + c * 18);
    const int z_off = (u >> 4) * 17 + (u & 15);                       // where z_p[u] lives inside its sub-problem
    __syncthreads();

    for (int it = 0; it < nt; it++) {
        const long long f0 = (t0 + it) * kQFr;
        float2 *tile = tile0 + (it & 1) * (kQFr * kQFrame);
        if (REGPF) {
#pragma unroll
            for (int f = 0; f < kQFr; f++)
#pragma unroll
                for (int q = 0; q < 4; q++) w[q][HL + f] = nx[4 * f + q];
            fetch(it + 1);
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int f = 0; f < kQFr; f++)
#pragma unroll
                for (int q = 0; q < 4; q++) w[q][HL + f] = slot[(4 * f + q) * kThreads + u];
        }
        // ---- phase 1
#pragma unroll
        for (int f = 0; f < kQFr; f++) {
            float2 d[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                d[q] = make_float2(0.f, 0.f);
#pragma unroll
                for (int n = P - 1; n >= 0; n--) d[q] = __ffma2_rn(bc(h[q][n]), w[q][HL + f - n], d[q]);
            }
            // radix-4 DIF butterfly of the forward transform (W4 = -j) and the three twiddles
            const float2 s02 = __fadd2_rn(d[0], d[2]), d02 = __fadd2_rn(d[0], neg2(d[2]));
            const float2 s13 = __fadd2_rn(d[1], d[3]), d13 = __fadd2_rn(d[1], neg2(d[3]));
            float2 *zf = tile + f * kQFrame + z_off;
            zf[0] = __fadd2_rn(s02, s13);
            zf[kQSub] = cmul2_rot(__fadd2_rn(d02, rot270(d13)), t1);
            zf[2 * kQSub] = cmul2_rot(__fadd2_rn(s02, neg2(s13)), t2);
            zf[3 * kQSub] = cmul2_rot(__fadd2_rn(d02, rot90(d13)), t3);
        }
        if (!REGPF) fetch(it + 1);
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int i = 0; i < HL; i++) w[q][i] = w[q][kQFr + i];
        __syncthreads();
        // ---- phase 2
        float2 *sp = tile + f2 * kQFrame + pp * kQSub;
        float2 v[16];
#pragma unroll
        for (int r1 = 0; r1 < 16; r1++) v[r1] = sp[r1 * 17 + c];
        RegDft<16, -1>::run(v);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 t = twc[k];
            if (k > 0) v[2 * k] = cmul2_rot(v[2 * k], make_float2(t.x, t.y));
            v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
        }
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) sp[k1 * 17 + c] = v[k1];
        asm volatile("bar.sync %0, 64;" ::"r"(1 + f2) : "memory");
#pragma unroll
        for (int cc = 0; cc < 16; cc++) v[cc] = sp[c * 17 + cc];      // this thread now plays row k1 = c
        RegDft<16, -1>::run(v);
        if (f0 + f2 < a.n_frames) {
            float2 *yf = y + (f0 + f2) * M + 4 * c + pp;
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) __stcs(yf + 64 * k2, v[k2]);
        }
    }
    if (!REGPF) asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int P>
void launch_quad_P(const QuadArgs &a, dim3 grid, cudaStream_t st)
{
    constexpr size_t smem = quad_smem_bytes<P>();
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch_quad_1024_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pfbch_quad_1024_kernel<P><<<grid, kThreads, smem, st>>>(a);
}

// ------------------------------------------------------------------------------------------ M = 256 synthesis
// firpfbch.c:314-336: v_s = unscaled backward DFT of frame s; y[sM + i] = sum_n h[i + nM] v_{s-n}[i].
//   phase A  16 threads per frame: the 256-point backward DFT straight from global memory (two passes of
//            16-point register DFTs, transposition through the padded tile), result rows left in the tile
//   phase B  thread i owns branch i: a register window of p-1+16 rows of v[.][i] and its p taps; 16 outputs
//            per tile, stored as coalesced rows.
// An instance that starts inside the stream first transforms the 16 frames before its range to rebuild the
// p-1 rows of history (6 % extra transforms at 256 frames per instance); the first instance takes them
// from the object's state (or zeros).
struct SynArgs {
    const float2 *hist_rows;   // hist_rows rows of v preceding frame 0 (single-stream stateful use), or null
    const float2 *x;           // [n_streams][stream_stride] channelised input frames
    float2 *y;
    const float *taps;         // [P][M]: taps[n][i] = h[i + nM]
    const float2 *tw;          // tw[k1][r2] = e^{+j 2 pi k1 r2 / 256}
    int n_hist_rows;
    int n_tiles;
    int tiles_per_inst;
    long long stream_stride;
};

template <int P>
__global__ void __launch_bounds__(kThreads) pfbch_synth_fused_256_kernel(const SynArgs a)
{
    constexpr int M = 256;
    constexpr int FR = 16 * 17;
    __shared__ float2 tile[kTile * FR];

    const int u = threadIdx.x;
    const int t0 = (int)((long long)blockIdx.x * a.tiles_per_inst);
    int nt = a.n_tiles - t0;
    if (nt > a.tiles_per_inst) nt = a.tiles_per_inst;
    const float2 *x = a.x + (long long)blockIdx.y * a.stream_stride;
    float2 *y = a.y + (long long)blockIdx.y * a.stream_stride;

    float h[P];
#pragma unroll
    for (int n = 0; n < P; n++) h[n] = __ldg(a.taps + n * M + u);
    float2 win[P - 1 + kTile];
    const int f2 = u >> 4, r2 = u & 15;
    const float2 *twp = a.tw + r2;

    // phase A for the 16 frames starting at frame s0: rows of v left in tile[f*FR + 0..255]
    auto transform_tile = [&](long long s0) {
        float2 *fr = tile + f2 * FR;
        const float2 *xf = x + (s0 + f2) * M + r2;
        float2 v[16];
#pragma unroll
        for (int r1 = 0; r1 < 16; r1++) v[r1] = __ldg(xf + 16 * r1);
        RegDft<16, +1>::run(v);
#pragma unroll
        for (int k1 = 1; k1 < 16; k1++) v[k1] = cmul2_rot(v[k1], __ldg(twp + 16 * k1));
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) fr[r2 * 17 + k1] = v[k1];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 16; c++) v[c] = fr[c * 17 + r2];
        RegDft<16, +1>::run(v);
        __syncwarp();                                      // the group has read everything it transposed
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) fr[r2 + 16 * k2] = v[k2];     // v[f][k1 + 16 k2], natural order
    };

    if (t0 > 0) {
        transform_tile((long long)t0 * kTile - kTile);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < P - 1; i++) win[i] = tile[(kTile - (P - 1) + i) * FR + u];
        __syncthreads();
    } else {
#pragma unroll
        for (int i = 0; i < P - 1; i++) {
            const int r = a.n_hist_rows - (P - 1) + i;
            win[i] = (a.hist_rows != nullptr && r >= 0) ? __ldg(a.hist_rows + (long long)r * M + u) : make_float2(0.f, 0.f);
        }
    }
    for (int it = 0; it < nt; it++) {
        const long long s0 = (long long)(t0 + it) * kTile;
        transform_tile(s0);
        __syncthreads();
#pragma unroll
        for (int f = 0; f < kTile; f++) win[P - 1 + f] = tile[f * FR + u];
        float2 out[kTile];
        mac_tile<M, P>(h, win, out);
#pragma unroll
        for (int f = 0; f < kTile; f++) __stcs(y + (s0 + f) * M + u, out[f]);
#pragma unroll
        for (int i = 0; i < P - 1; i++) win[i] = win[kTile + i];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------ M = 1024 synthesis
// firpfbch.c:314-336 at M = 1024: the transposed form of pfbch_quad_1024_kernel.  64 threads per frame, lane =
// (p fastest, row): the four 256-point backward problems X_p[k] = X[4k + p] -- 16-point DFTs along the rows out of
// register-prefetched inputs (256-byte runs of the frame), a 64-thread named barrier, twiddle and 16-point DFTs
// down the columns; then thread u does the radix-4 last stage v[u + 256 q] = sum_p (+j)^(q p) W1024^(u p) E_p[u],
// pushes the four values into register windows and forms its four outputs y[s][u + 256 q] = sum_n h[u + 256 q + nM]
// v_{s-n}[u + 256 q].  An instance that starts inside the stream first transforms the ceil((P-1)/4) tiles before
// its range; the object's state (the last P rows of v) is recomputed by the caller with the batched transform.
template <int P>
__global__ void __launch_bounds__(kThreads, P <= 4 ? 2 : 1) pfbch_quad_synth_1024_kernel(const SynArgs a, const long long n_frames)
{
    constexpr int M = 1024, HL = P - 1;
    constexpr int WT = (HL + kQFr - 1) / kQFr;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tile0 = reinterpret_cast<float2 *>(smem_raw);
    float2 *tws = tile0 + 2 * kQFr * kQFrame;                         // [c][k1], pitch 18
    const int u = threadIdx.x;
    const long long t0 = (long long)blockIdx.x * a.tiles_per_inst;
    long long ntl = (long long)a.n_tiles - t0;
    if (ntl > a.tiles_per_inst) ntl = a.tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;
    const float2 *x = a.x + (long long)blockIdx.y * a.stream_stride;
    float2 *y = a.y + (long long)blockIdx.y * a.stream_stride;

    const int f2 = u >> 6, pp = u & 3, idx = (u >> 2) & 15;
    float2 nx[16];
    auto fetch = [&](int it) {
        const long long f0 = (t0 + it) * kQFr;
        const bool ok = it < nt && f0 + f2 < n_frames;
        const float2 *src = x + (f0 + f2) * M + 4 * idx + pp;         // X_p[k1 = idx + 16 k2] at + 64 k2
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) nx[k2] = ok ? __ldg(src + 64 * k2) : make_float2(0.f, 0.f);
    };
    const int it0 = (t0 > 0 && nt > 0) ? -WT : 0;
    fetch(it0);
    tws[(u & 15) * 18 + (u >> 4)] = __ldg(a.tw + 768 + u);            // tw256[k1 = u >> 4][c = u & 15]
    const float2 t1 = __ldg(a.tw + u), t2 = __ldg(a.tw + 256 + u), t3 = __ldg(a.tw + 512 + u);
    float h[4][P];
    float2 w[4][HL + kQFr];
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
        for (int n = 0; n < P; n++) h[q][n] = __ldg(a.taps + n * M + u + 256 * q);
#pragma unroll
        for (int i = 0; i < HL; i++) {
            const int r = a.n_hist_rows - HL + i;
            w[q][i] = (t0 == 0 && a.hist_rows != nullptr && r >= 0) ? __ldg(a.hist_rows + (long long)r * M + u + 256 * q)
                                                                   : make_float2(0.f, 0.f);
        }
    }
    const float4 *twc = reinterpret_cast<const float4 *>(tws + idx * 18);
    const int z_off = (u >> 4) * 17 + (u & 15);
    __syncthreads();

    for (int it = it0; it < nt; it++) {
        const long long f0 = (t0 + it) * kQFr;
        float2 *tile = tile0 + ((it - it0) & 1) * (kQFr * kQFrame);
        float2 *sp = tile + f2 * kQFrame + pp * kQSub;
        float2 v[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) v[k2] = nx[k2];
        fetch(it + 1);
        RegDft<16, +1>::run(v);
#pragma unroll
        for (int c = 0; c < 16; c++) sp[idx * 17 + c] = v[c];
        asm volatile("bar.sync %0, 64;" ::"r"(1 + f2) : "memory");
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) v[k1] = sp[k1 * 17 + idx];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 t = twc[k];
            if (k > 0) v[2 * k] = cmul2_rot(v[2 * k], make_float2(t.x, t.y));
            v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
        }
        RegDft<16, +1>::run(v);
#pragma unroll
        for (int r1 = 0; r1 < 16; r1++) sp[r1 * 17 + idx] = v[r1];
        __syncthreads();
#pragma unroll
        for (int f = 0; f < kQFr; f++) {
            const float2 *zf = tile + f * kQFrame + z_off;
            const float2 b0 = zf[0], b1 = cmul2_rot(zf[kQSub], t1), b2 = cmul2_rot(zf[2 * kQSub], t2), b3 = cmul2_rot(zf[3 * kQSub], t3);
            const float2 s02 = __fadd2_rn(b0, b2), d02 = __fadd2_rn(b0, neg2(b2));
            const float2 s13 = __fadd2_rn(b1, b3), d13 = __fadd2_rn(b1, neg2(b3));
            w[0][HL + f] = __fadd2_rn(s02, s13);
            w[1][HL + f] = __fadd2_rn(d02, rot90(d13));               // W4 = +j (backward transform)
            w[2][HL + f] = __fadd2_rn(s02, neg2(s13));
            w[3][HL + f] = __fadd2_rn(d02, rot270(d13));
        }
        if (it >= 0) {
#pragma unroll
            for (int f = 0; f < kQFr; f++) {
                float2 acc[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    acc[q] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int n = P - 1; n >= 0; n--) acc[q] = __ffma2_rn(bc(h[q][n]), w[q][HL + f - n], acc[q]);
                }
                if (f0 + f < n_frames) {
#pragma unroll
                    for (int q = 0; q < 4; q++) __stcs(y + (f0 + f) * M + u + 256 * q, acc[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int i = 0; i < HL; i++) w[q][i] = w[q][kQFr + i];
    }
}

template <int P>
void launch_quad_synth_P(const SynArgs &a, long long n_frames, dim3 grid, cudaStream_t st)
{
    constexpr size_t smem = sizeof(float2) * ((size_t)2 * kQFr * kQFrame + 16 * 18);
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch_quad_synth_1024_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pfbch_quad_synth_1024_kernel<P><<<grid, kThreads, smem, st>>>(a, n_frames);
}

}  // namespace

namespace lc {
namespace pfb {

bool fused_pfbch_supported(unsigned int M, unsigned int p)
{
    return (M == 16 && (p == 4 || p == 6 || p == 8 || p == 10 || p == 12 || p == 14)) || (M == 256 && (p == 4 || p == 10)) || tile_pfbch_supported(M, p) ||
           small_pfbch_supported(M, p) || (M == 1024 && (p == 4 || p == 6 || p == 8 || p == 10 || p == 12 || p == 14));
}

bool fused_pfbch_synth_supported(unsigned int M, unsigned int p)
{
    return (M == 256 && (p == 4 || p == 6 || p == 8 || p == 10 || p == 12 || p == 14)) || (M == 1024 && (p == 4 || p == 6 || p == 8 || p == 10 || p == 12 || p == 14)) ||
           tile_pfbch_synth_supported(M, p);
}

void fused_pfbch_make_twiddles(std::vector<float2> &tw, int sign, unsigned int M, unsigned int p)
{
    if (sign > 0 && tile_pfbch_synth_supported(M, p)) {
        tile_pfbch2_make_twiddles(tw, M);                     // synthesizers for M = 64 / 128 / 512 (pfbch2_tile.cu): backward table
        return;
    }
    if (sign < 0 && small_pfbch_supported(M, p)) {
        tw.assign(1, make_float2(1.f, 0.f));                  // these kernels need no table
        return;
    }
    if (sign < 0 && tile_pfbch_supported(M, p)) {
        tile_pfbch_make_twiddles(tw, M);                      // analyzers for M = 64 / 128 / 512 (pfbch2_tile.cu)
        return;
    }
    if (M == 1024) {
        // pfbch_quad[_synth]_1024_kernel: tw4[p-1][u] = e^{sign j 2 pi u p / 1024}, then tw256[k1][r2] = e^{sign j 2 pi k1 r2 / 256}
        tw.resize(3 * 256 + 256);
        for (int pp = 1; pp <= 3; pp++)
            for (int u = 0; u < 256; u++) {
                double ang = (double)sign * 2.0 * M_PI * (double)(u * pp) / 1024.0;
                tw[(pp - 1) * 256 + u] = make_float2((float)cos(ang), (float)sin(ang));
            }
        for (int k1 = 0; k1 < 16; k1++)
            for (int r2 = 0; r2 < 16; r2++) {
                double ang = (double)sign * 2.0 * M_PI * (double)(k1 * r2) / 256.0;
                tw[768 + k1 * 16 + r2] = make_float2((float)cos(ang), (float)sin(ang));
            }
        return;
    }
    tw.resize(16 * 16);
    for (int k1 = 0; k1 < 16; k1++)
        for (int r2 = 0; r2 < 16; r2++) {
            double ang = (double)sign * 2.0 * M_PI * (double)(k1 * r2) / 256.0;
            tw[k1 * 16 + r2] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

long long launch_pfbch_fused(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, const float2 *d_tw,
                             float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                             unsigned int n_streams, long long stream_stride, cudaStream_t st)
{
    if (tile_pfbch_supported(M, p))
        return launch_pfbch_tile(d_hist, d_x, H, d_taps, d_tw, d_y, M, p, n_frames, n_streams, stream_stride, st);
    if (small_pfbch_supported(M, p))
        return launch_pfbch_small(d_hist, d_x, H, d_taps, d_y, M, p, n_frames, n_streams, stream_stride, st);
    if (M == 1024) {
        const long long nq = (long long)((n_frames + kQFr - 1) / kQFr);
        if (nq == 0 || nq > (1ll << 30) || n_streams == 0 || n_streams > 65535) return 0;
        QuadArgs q;
        q.hist = d_hist;
        q.x = d_x;
        q.y = d_y;
        q.taps = d_taps;
        q.tw = d_tw;
        q.H = H;
        q.n_frames = (long long)n_frames;
        q.stream_stride = stream_stride;
        q.n_tiles = (int)nq;
        long long tpi = 128;                                  // 512 frames per instance: < 2 % of history re-reads
        while (tpi > 1 && (nq / tpi) * n_streams < 148ll * 4) tpi >>= 1;
        q.tiles_per_inst = (int)tpi;
        dim3 grid((unsigned int)((nq + tpi - 1) / tpi), n_streams, 1);
        switch (p) {
        case 4: launch_quad_P<4>(q, grid, st); break;
        case 6: launch_quad_P<6>(q, grid, st); break;
        case 8: launch_quad_P<8>(q, grid, st); break;
        case 10: launch_quad_P<10>(q, grid, st); break;
        case 12: launch_quad_P<12>(q, grid, st); break;
        case 14: launch_quad_P<14>(q, grid, st); break;
        default: fatal("launch_pfbch_fused", "unsupported shape");
        }
        LC_LAUNCH_CHECK();
        return (long long)n_frames;
    }
    const long long n_tiles = (long long)(n_frames / kTile);
    if (n_tiles == 0 || n_tiles > (1ll << 30) || n_streams == 0 || n_streams > 65535) return 0;
    PfbArgs a;
    a.hist = d_hist;
    a.x = d_x;
    a.y = d_y;
    a.taps = d_taps;
    a.tw = d_tw;
    a.H = H;
    a.n_tiles = (int)n_tiles;
    a.stream_stride = stream_stride;
    // ~16 tiles (256 frames) per instance keeps the re-read of the p-1 history frames below 4 %
    // while leaving enough instances to fill the GPU; shrink for short streams
    const long long inst_per_cta = (M == 16) ? 16 : 1;
    long long tpi = 16;
    while (tpi > 1 && (n_tiles / tpi) * n_streams < 148ll * 8 * inst_per_cta) tpi >>= 1;
    a.tiles_per_inst = (int)tpi;
    const long long n_inst = (n_tiles + tpi - 1) / tpi;
    dim3 grid((unsigned int)((n_inst + inst_per_cta - 1) / inst_per_cta), n_streams, 1);
    if (M == 16 && p == 4) pfbch_fused_16_kernel<4><<<grid, kThreads, 0, st>>>(a);
    else if (M == 16 && p == 6) pfbch_fused_16_kernel<6><<<grid, kThreads, 0, st>>>(a);
    else if (M == 16 && p == 8) pfbch_fused_16_kernel<8><<<grid, kThreads, 0, st>>>(a);
    else if (M == 16 && p == 10) pfbch_fused_16_kernel<10><<<grid, kThreads, 0, st>>>(a);
    else if (M == 16 && p == 12) pfbch_fused_16_kernel<12><<<grid, kThreads, 0, st>>>(a);
    else if (M == 16 && p == 14) pfbch_fused_16_kernel<14><<<grid, kThreads, 0, st>>>(a);
    else if (M == 256 && p == 4) pfbch_fused_256_kernel<4><<<grid, kThreads, 0, st>>>(a);
    else if (M == 256 && p == 10) pfbch_fused_256_kernel<10><<<grid, kThreads, 0, st>>>(a);
    else fatal("launch_pfbch_fused", "unsupported shape");
    LC_LAUNCH_CHECK();
    return n_tiles * kTile;
}

long long launch_pfbch_synth_fused(const float2 *d_hist_rows, int n_hist_rows, const float2 *d_x, const float *d_taps,
                                   const float2 *d_tw, float2 *d_y, unsigned int M, unsigned int p,
                                   unsigned long long n_frames, unsigned int n_streams, long long stream_stride,
                                   cudaStream_t st)
{
    if (tile_pfbch_synth_supported(M, p))
        return launch_pfbch_tile_synth(d_hist_rows, n_hist_rows, d_x, d_taps, d_tw, d_y, M, p, n_frames, n_streams, stream_stride, st);
    if (M == 1024) {
        const long long nq = (long long)((n_frames + kQFr - 1) / kQFr);
        if (nq == 0 || nq > (1ll << 30) || n_streams == 0 || n_streams > 65535) return 0;
        SynArgs q;
        q.hist_rows = d_hist_rows;
        q.n_hist_rows = n_hist_rows;
        q.x = d_x;
        q.y = d_y;
        q.taps = d_taps;
        q.tw = d_tw;
        q.n_tiles = (int)nq;
        q.stream_stride = stream_stride;
        long long tpi = 128;                                  // never fewer than 8 tiles: the warm-up reaches back up to 3
        while (tpi > 8 && (nq / tpi) * n_streams < 148ll * 4) tpi >>= 1;
        q.tiles_per_inst = (int)tpi;
        dim3 grid((unsigned int)((nq + tpi - 1) / tpi), n_streams, 1);
        switch (p) {
        case 4: launch_quad_synth_P<4>(q, (long long)n_frames, grid, st); break;
        case 6: launch_quad_synth_P<6>(q, (long long)n_frames, grid, st); break;
        case 8: launch_quad_synth_P<8>(q, (long long)n_frames, grid, st); break;
        case 10: launch_quad_synth_P<10>(q, (long long)n_frames, grid, st); break;
        case 12: launch_quad_synth_P<12>(q, (long long)n_frames, grid, st); break;
        case 14: launch_quad_synth_P<14>(q, (long long)n_frames, grid, st); break;
        default: fatal("launch_pfbch_synth_fused", "unsupported shape");
        }
        LC_LAUNCH_CHECK();
        return (long long)n_frames;
    }
    const long long n_tiles = (long long)(n_frames / kTile);
    if (n_tiles == 0 || n_tiles > (1ll << 30) || n_streams == 0 || n_streams > 65535) return 0;
    SynArgs a;
    a.hist_rows = d_hist_rows;
    a.n_hist_rows = n_hist_rows;
    a.x = d_x;
    a.y = d_y;
    a.taps = d_taps;
    a.tw = d_tw;
    a.n_tiles = (int)n_tiles;
    a.stream_stride = stream_stride;
    long long tpi = 16;
    while (tpi > 1 && (n_tiles / tpi) * n_streams < 148ll * 8) tpi >>= 1;
    a.tiles_per_inst = (int)tpi;
    dim3 grid((unsigned int)((n_tiles + tpi - 1) / tpi), n_streams, 1);
    if (p == 4) pfbch_synth_fused_256_kernel<4><<<grid, kThreads, 0, st>>>(a);
    else if (p == 6) pfbch_synth_fused_256_kernel<6><<<grid, kThreads, 0, st>>>(a);
    else if (p == 8) pfbch_synth_fused_256_kernel<8><<<grid, kThreads, 0, st>>>(a);
    else if (p == 12) pfbch_synth_fused_256_kernel<12><<<grid, kThreads, 0, st>>>(a);
    else if (p == 14) pfbch_synth_fused_256_kernel<14><<<grid, kThreads, 0, st>>>(a);
    else if (p == 10) pfbch_synth_fused_256_kernel<10><<<grid, kThreads, 0, st>>>(a);
    else fatal("launch_pfbch_synth_fused", "unsupported shape");
    LC_LAUNCH_CHECK();
    return n_tiles * kTile;
}

}  // namespace pfb
}  // namespace lc
