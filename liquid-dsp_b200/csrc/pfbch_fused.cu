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
__global__ void __launch_bounds__(kThreads) pfbch_fused_256_kernel(const PfbArgs a)
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
    for (int it = 0; it < nt; it++) {                      // nt is CTA-uniform
        const long long s0 = (long long)(t0 + it) * kTile;
        // ---- phase 1
#pragma unroll
        for (int f = 0; f < kTile; f++) win[P - 1 + f] = __ldg(x + (s0 + f) * M + u);
        float2 out[kTile];
        mac_tile<M, P>(h, win, out);
#pragma unroll
        for (int f = 0; f < kTile; f++) tile[f * FR + u + (u >> 4)] = out[f];
#pragma unroll
        for (int i = 0; i < P - 1; i++) win[i] = win[kTile + i];
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

}  // namespace

namespace lc {
namespace pfb {

bool fused_pfbch_supported(unsigned int M, unsigned int p)
{
    return tile_pfbch_supported(M, p) || (M == 16 && (p == 4 || p == 8)) || (M == 256 && (p == 4 || p == 10));
}

bool fused_pfbch_synth_supported(unsigned int M, unsigned int p)
{
    return M == 256 && (p == 4 || p == 10);
}

void fused_pfbch_make_twiddles(std::vector<float2> &tw, int sign, unsigned int M, unsigned int p)
{
    if (sign < 0 && tile_pfbch_supported(M, p)) {
        tile_pfbch_make_twiddles(tw, M);                      // analyzers for M = 64 / 128 / 512 (pfbch2_tile.cu)
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
    else if (M == 16 && p == 8) pfbch_fused_16_kernel<8><<<grid, kThreads, 0, st>>>(a);
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
    (void)M;
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
    else if (p == 10) pfbch_synth_fused_256_kernel<10><<<grid, kThreads, 0, st>>>(a);
    else fatal("launch_pfbch_synth_fused", "unsupported shape");
    LC_LAUNCH_CHECK();
    return n_tiles * kTile;
}

}  // namespace pfb
}  // namespace lc
