// pfbch2_small.cu -- fused firpfbch2_crcf kernels for the smallest shapes, M = 4 / 8 / 16 / 32 (sm_100a): the low end of
// the reference's own benchmark (src/multichannel/bench/firpfbch2_crcf_benchmark.c:70-90, M = 4 and 16, m = 2).
//
// A frame is M/2 input samples and an M-point transform: far too little for a CTA, or even a warp.  An INSTANCE is
// M/2 threads (2 ... 16 lanes of a warp, 16 ... 2 instances per warp, 256 / (M/2) per CTA); it walks consecutive
// tiles of 16 frames of its own chunk of the stream and never synchronises with anything but itself (__syncwarp):
//   analyzer     phase 1  thread u owns the window pair (u, u + M/2): one new sample per frame, two L-tap dot products
//                         over a sliding register window (as pfbch2_tile.cu), stored as X[f][u], X[f][u + M/2];
//                phase 2  each thread takes whole frames (16 / (M/2) of them): one 16-byte-aligned row of M values,
//                         an M-point register DFT, the row written back;
//                phase 3  the tile is 16 M contiguous outputs: 128-bit coalesced stores by the instance.
//   synthesizer  the transposed order: the tile's 16 M channel samples arrive with 128-bit loads, each thread
//                transforms whole frames (times the reference's 1/M * M/2 = 1/2), then thread u pushes v[f][u] and
//                v[f][u + M/2] into register windows of the last 4m rows and forms y[t][u] with packed FFMA2.  The
//                kernel leaves the rows of v of the last 4m frames as the object's state; an instance that starts
//                inside the stream first transforms the tile before its range.
// The next tile's inputs are prefetched into registers while the current tile is computed.  The 1/M output scale of
// the analyzer is folded into the taps (exact: M is a power of two).
//
// Reference semantics: src/multichannel/src/firpfbch2.c:244-282 (analyzer), :287-335 (synthesizer).
#include "pfbch2_small.cuh"
#include "fused_common.cuh"

#include <cmath>

namespace {

using namespace lcf;

constexpr int kThreads = 256;
constexpr int kFr = 16;                         // frames per tile

struct SmallArgs {
    const float2 *hist;        // analyzer: H samples preceding x[0]; synthesizer: hist_rows rows of v; or null = zero state
    const float2 *x;
    float2 *y;
    float2 *hist_out;          // synthesizer: [NT][M], the rows of v of the last NT frames
    const float *taps;         // analyzer: [2][L][M] pre-scaled by 1/M (parity-0 half is used); synthesizer: [NT][M/2]
    long long n_frames;
    int H;                     // analyzer: samples of history; synthesizer: rows of history
    int par0;
    int n_tiles;
    int tiles_per_inst;
};

__device__ __forceinline__ float2 ld_hist0(const float2 *__restrict__ hist, const float2 *__restrict__ x, long long idx, int H)
{
    if (idx >= 0) return __ldg(x + idx);
    if (hist != nullptr && H + idx >= 0) return __ldg(hist + (H + idx));
    return make_float2(0.f, 0.f);
}

// row pitch of an instance's tile in float2: M + 2 keeps every row 16-byte aligned and spreads the rows that the
// lanes of a quarter-warp read with 128-bit loads over all bank groups (pitch / 2 is odd)
template <int M> constexpr int kPitch = M + 2;

// M-point backward DFT of row `row` of the tile, in place, times `scale` (1 = none)
template <int M>
__device__ __forceinline__ void transform_row(float2 *row, float scale)
{
    float2 v[M];
    const float4 *r4 = reinterpret_cast<const float4 *>(row);
#pragma unroll
    for (int i = 0; i < M / 2; i++) {
        const float4 t = r4[i];
        v[2 * i] = make_float2(t.x, t.y);
        v[2 * i + 1] = make_float2(t.z, t.w);
    }
    RegDft<M, +1>::run(v);
    float4 *w4 = reinterpret_cast<float4 *>(row);
#pragma unroll
    for (int i = 0; i < M / 2; i++)
        w4[i] = make_float4(scale * v[2 * i].x, scale * v[2 * i].y, scale * v[2 * i + 1].x, scale * v[2 * i + 1].y);
}

template <int M, int L>
__global__ void __launch_bounds__(kThreads) pfbch2_small_kernel(const SmallArgs a)
{
    constexpr int M2 = M / 2, PITCH = kPitch<M>, INST = kThreads / M2, HL = 2 * L - 1;
    extern __shared__ __align__(16) float2 tiles[];              // [INST][kFr][PITCH]
    const int inst_l = threadIdx.x / M2, u = threadIdx.x % M2;
    float2 *tile = tiles + inst_l * (kFr * PITCH);
    const long long inst = (long long)blockIdx.x * INST + inst_l;
    const long long t0 = inst * a.tiles_per_inst;
    long long ntl = (long long)a.n_tiles - t0;
    if (ntl > a.tiles_per_inst) ntl = a.tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;                        // uniform within an instance
    if (nt == 0) return;                                          // (whole instances leave: the others only __syncwarp)
    // the lanes of this instance inside its warp
    const unsigned int imask = (M2 == 32) ? 0xffffffffu : (((1u << M2) - 1u) << ((threadIdx.x & 31) / M2 * M2));

    float hA[L], hB[L];
#pragma unroll
    for (int n = 0; n < L; n++) {
        hA[n] = __ldg(a.taps + n * M + u);
        hB[n] = __ldg(a.taps + n * M + u + M2);
    }
    const float2 *xs = a.x + (M2 - 1 - u);
    float2 win[HL + kFr], nx[kFr];
    auto fetch = [&](int it) {
        const long long f0 = (t0 + it) * kFr;
#pragma unroll
        for (int f = 0; f < kFr; f++) nx[f] = (it < nt && f0 + f < a.n_frames) ? __ldg(xs + (f0 + f) * M2) : make_float2(0.f, 0.f);
    };
    fetch(0);
    {
        const long long f0 = t0 * kFr;
#pragma unroll
        for (int i = 0; i < HL; i++) win[i] = ld_hist0(a.hist, a.x, (f0 - HL + i) * M2 + (M2 - 1 - u), a.H);
    }
    for (int it = 0; it < nt; it++) {
        const long long f0 = (t0 + it) * kFr;                     // even
#pragma unroll
        for (int f = 0; f < kFr; f++) win[HL + f] = nx[f];
        fetch(it + 1);
        // ---- phase 1
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            float2 d1 = make_float2(0.f, 0.f), d2 = d1;
#pragma unroll
            for (int n = L - 1; n >= 0; n--) {
                d1 = __ffma2_rn(bc(hA[n]), win[HL + f - 2 * n], d1);
                d2 = __ffma2_rn(bc(hB[n]), win[HL + f - 1 - 2 * n], d2);
            }
            const bool odd = ((a.par0 + f) & 1) != 0;             // (X[u], X[u + M/2]) = (d1, d2) on even frames, (d2, d1) on odd ones
            tile[f * PITCH + u] = odd ? d2 : d1;
            tile[f * PITCH + u + M2] = odd ? d1 : d2;
        }
#pragma unroll
        for (int i = 0; i < HL; i++) win[i] = win[kFr + i];
        __syncwarp(imask);
        // ---- phase 2: whole frames
#pragma unroll
        for (int i = 0; i < kFr / M2; i++) transform_row<M>(tile + (u + M2 * i) * PITCH, 1.0f);
        __syncwarp(imask);
        // ---- phase 3: 16 M contiguous outputs, two per thread and store
        float4 *yt = reinterpret_cast<float4 *>(a.y + f0 * M);
#pragma unroll
        for (int i = 0; i < kFr; i++)                             // float4 i M/2 + u of the tile = row i, columns 2u, 2u + 1
            if (f0 + i < a.n_frames) __stcs(yt + i * M2 + u, *reinterpret_cast<const float4 *>(tile + i * PITCH + 2 * u));
        __syncwarp(imask);
    }
}

template <int M, int NT, int PAR0>
__global__ void __launch_bounds__(kThreads) pfbch2_small_synth_kernel(const SmallArgs a)
{
    constexpr int M2 = M / 2, PITCH = kPitch<M>, INST = kThreads / M2, HL = NT - 1;
    constexpr int WT = (HL + kFr - 1) / kFr;                      // warm-up tiles of an instance inside the stream
    extern __shared__ __align__(16) float2 tiles[];              // [INST][kFr][PITCH]
    const int inst_l = threadIdx.x / M2, u = threadIdx.x % M2;
    float2 *tile = tiles + inst_l * (kFr * PITCH);
    const long long inst = (long long)blockIdx.x * INST + inst_l;
    const long long t0 = inst * a.tiles_per_inst;
    long long ntl = (long long)a.n_tiles - t0;
    if (ntl > a.tiles_per_inst) ntl = a.tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;
    if (nt == 0) return;
    const unsigned int imask = (M2 == 32) ? 0xffffffffu : (((1u << M2) - 1u) << ((threadIdx.x & 31) / M2 * M2));

    float h[NT];
#pragma unroll
    for (int q = 0; q < NT; q++) h[q] = __ldg(a.taps + q * M2 + u);
    float2 wlo[HL + kFr], whi[HL + kFr];
#pragma unroll
    for (int i = 0; i < HL; i++) {
        wlo[i] = whi[i] = make_float2(0.f, 0.f);
        if (t0 == 0 && a.hist != nullptr) {
            wlo[i] = __ldg(a.hist + (size_t)(a.H - HL + i) * M + u);
            whi[i] = __ldg(a.hist + (size_t)(a.H - HL + i) * M + u + M2);
        }
    }
    // the tile's 16 M channel samples: float4 i M/2 + u = row i, columns 2u, 2u + 1
    float4 nx[kFr];
    auto fetch = [&](int it) {
        const long long f0 = (t0 + it) * kFr;
        const float4 *xt = reinterpret_cast<const float4 *>(a.x + f0 * M);
#pragma unroll
        for (int i = 0; i < kFr; i++)
            nx[i] = (it < nt && f0 + i < a.n_frames) ? __ldg(xt + i * M2 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    const int it0 = t0 > 0 ? -WT : 0;
    fetch(it0);
    for (int it = it0; it < nt; it++) {
        const long long f0 = (t0 + it) * kFr;
#pragma unroll
        for (int i = 0; i < kFr; i++) *reinterpret_cast<float4 *>(tile + i * PITCH + 2 * u) = nx[i];
        fetch(it + 1);
        __syncwarp(imask);
#pragma unroll
        for (int i = 0; i < kFr / M2; i++) transform_row<M>(tile + (u + M2 * i) * PITCH, 0.5f);
        __syncwarp(imask);
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            wlo[HL + f] = tile[f * PITCH + u];
            whi[HL + f] = tile[f * PITCH + u + M2];
        }
        __syncwarp(imask);
        if (it >= 0) {
#pragma unroll
            for (int f = 0; f < kFr; f++) {
                const bool odd = ((PAR0 + f) & 1) != 0;
                float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
                for (int q = 0; q < NT; q++) {
                    const float2 sv = odd ? whi[HL + f - q] : wlo[HL + f - q];
                    if (q & 1) a1 = __ffma2_rn(bc(h[q]), sv, a1);
                    else       a0 = __ffma2_rn(bc(h[q]), sv, a0);
                }
                if (f0 + f < a.n_frames) __stcs(a.y + (f0 + f) * M2 + u, __fadd2_rn(a0, a1));
            }
            if (f0 + kFr > a.n_frames - NT) {
#pragma unroll
                for (int f = 0; f < kFr; f++) {
                    const long long r = f0 + f - (a.n_frames - NT);
                    if (r >= 0 && f0 + f < a.n_frames) {
                        a.hist_out[r * M + u] = wlo[HL + f];
                        a.hist_out[r * M + u + M2] = whi[HL + f];
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < HL; i++) {
            wlo[i] = wlo[kFr + i];
            whi[i] = whi[kFr + i];
        }
    }
}

// ---------------------------------------------------------------------------------- M = 4: one thread per frame
// At M = 4 a frame is 16 bytes in and 32 bytes out: an instance of two lanes per chunk of the stream leaves every warp
// request scattered over 16 lines (measured 0.29 of the roofline whatever the tiling).  Here consecutive lanes take
// consecutive FRAMES: a lane re-reads the 2L frames its windows span (128-bit loads, coalesced across the warp; the 2L-fold
// overlap between neighbouring lanes is served by L1), does the M dot products and the M-point DFT in registers and
// writes its frame.  No shared memory, no synchronisation, no state carried between frames.
template <int M, int L>
__global__ void __launch_bounds__(kThreads) pfbch2_frame_kernel(const SmallArgs a)
{
    constexpr int M2 = M / 2, V = M2 / 2;                         // V: float4 per frame of input
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= a.n_frames) return;
    // samples of the frames t - (2L-1) ... t: fr[k] = frame t - k, stream order inside a frame
    float2 fr[2 * L][M2];
#pragma unroll
    for (int k = 0; k < 2 * L; k++) {
        const long long f = t - k;
        if (f >= 0) {
            const float4 *p = reinterpret_cast<const float4 *>(a.x + f * M2);
#pragma unroll
            for (int j = 0; j < V; j++) {
                const float4 q = __ldg(p + j);
                fr[k][2 * j] = make_float2(q.x, q.y);
                fr[k][2 * j + 1] = make_float2(q.z, q.w);
            }
        } else {
#pragma unroll
            for (int j = 0; j < M2; j++) fr[k][j] = ld_hist0(a.hist, a.x, f * M2 + j, a.H);
        }
    }
    const bool odd = ((a.par0 + (int)(t & 1)) & 1) != 0;
    float2 X[M];
#pragma unroll
    for (int u = 0; u < M2; u++) {
        // the pair's sample of frame T is stream index T M/2 + M/2-1-u:  d1 = sum hA[n] s_{T-2n}, d2 = sum hB[n] s_{T-1-2n}
        float2 d1 = make_float2(0.f, 0.f), d2 = d1;
#pragma unroll
        for (int n = L - 1; n >= 0; n--) {
            d1 = __ffma2_rn(bc(__ldg(a.taps + n * M + u)), fr[2 * n][M2 - 1 - u], d1);
            d2 = __ffma2_rn(bc(__ldg(a.taps + n * M + u + M2)), fr[2 * n + 1][M2 - 1 - u], d2);
        }
        X[u] = odd ? d2 : d1;
        X[u + M2] = odd ? d1 : d2;
    }
    RegDft<M, +1>::run(X);
    float4 *yo = reinterpret_cast<float4 *>(a.y + t * M);
#pragma unroll
    for (int j = 0; j < M2; j++) __stcs(yo + j, make_float4(X[2 * j].x, X[2 * j].y, X[2 * j + 1].x, X[2 * j + 1].y));
}

// synthesizer, one thread per output frame: the NT rows of v its taps span are recomputed from the channel frames
// (an M-point register DFT each), row t - q from the object's state when it lies before the block
template <int M, int NT>
__global__ void __launch_bounds__(kThreads) pfbch2_frame_synth_kernel(const SmallArgs a)
{
    constexpr int M2 = M / 2;
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= a.n_frames) return;
    const bool odd = ((a.par0 + (int)(t & 1)) & 1) != 0;
    float2 acc0[M2], acc1[M2];
#pragma unroll
    for (int i = 0; i < M2; i++) acc0[i] = acc1[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < NT; q++) {
        const long long f = t - q;
        float2 v[M];
        if (f >= 0) {
            const float4 *p = reinterpret_cast<const float4 *>(a.x + f * M);
#pragma unroll
            for (int j = 0; j < M2; j++) {
                const float4 w = __ldg(p + j);
                v[2 * j] = make_float2(w.x, w.y);
                v[2 * j + 1] = make_float2(w.z, w.w);
            }
            RegDft<M, +1>::run(v);
#pragma unroll
            for (int j = 0; j < M; j++) v[j] = __fmul2_rn(bc(0.5f), v[j]);        // the reference's (1/M) * (M/2)
        } else {
#pragma unroll
            for (int j = 0; j < M; j++)
                v[j] = (a.hist != nullptr && a.H + f >= 0) ? __ldg(a.hist + (a.H + f) * M + j) : make_float2(0.f, 0.f);
        }
        if (q == 0 && t >= a.n_frames - NT) {
            // the row of v of this frame is part of the object's state after the call
#pragma unroll
            for (int j = 0; j < M; j++) a.hist_out[(t - (a.n_frames - NT)) * M + j] = v[j];
        }
#pragma unroll
        for (int i = 0; i < M2; i++) {
            const float2 sv = odd ? v[i + M2] : v[i];
            const float c = __ldg(a.taps + q * M2 + i);
            if (q & 1) acc1[i] = __ffma2_rn(bc(c), sv, acc1[i]);
            else       acc0[i] = __ffma2_rn(bc(c), sv, acc0[i]);
        }
    }
    float2 *yo = a.y + t * M2;
#pragma unroll
    for (int i = 0; i < M2; i += 2) {
        const float2 r0 = __fadd2_rn(acc0[i], acc1[i]), r1 = __fadd2_rn(acc0[i + 1], acc1[i + 1]);
        __stcs(reinterpret_cast<float4 *>(yo + i), make_float4(r0.x, r0.y, r1.x, r1.y));
    }
}

// ---------------------------------------------------------------------------------- firpfbch (critically sampled), M = 4 / 8 / 32
// firpfbch.c:369-409: X_s[u] = sum_n h[(M-1-u) + nM] x[(s-n)M + u], Y_s = forward DFT.  (M = 16 has its own kernel in
// pfbch_fused.cu.)  M = 8 / 32: an instance of M lanes, thread u owns window u, whole-frame register DFTs, as above;
// M = 4: one thread per frame.  grid.y = stream.
struct PfbSmallArgs {
    const float2 *hist;
    const float2 *x;
    float2 *y;
    const float *taps;         // [P][M]: taps[n][i] = h[i + nM]
    long long n_frames;
    long long stream_stride;
    int H;
    int n_tiles;
    int tiles_per_inst;
};

template <int M>
__device__ __forceinline__ void transform_row_fwd(float2 *row)
{
    float2 v[M];
    const float4 *r4 = reinterpret_cast<const float4 *>(row);
#pragma unroll
    for (int i = 0; i < M / 2; i++) {
        const float4 t = r4[i];
        v[2 * i] = make_float2(t.x, t.y);
        v[2 * i + 1] = make_float2(t.z, t.w);
    }
    RegDft<M, -1>::run(v);
    float4 *w4 = reinterpret_cast<float4 *>(row);
#pragma unroll
    for (int i = 0; i < M / 2; i++) w4[i] = make_float4(v[2 * i].x, v[2 * i].y, v[2 * i + 1].x, v[2 * i + 1].y);
}

template <int M, int P>
__global__ void __launch_bounds__(kThreads) pfbch_small_kernel(const PfbSmallArgs a)
{
    constexpr int PITCH = kPitch<M>, INST = kThreads / M, HL = P - 1;
    extern __shared__ __align__(16) float2 tiles[];              // [INST][kFr][PITCH]
    const int inst_l = threadIdx.x / M, u = threadIdx.x % M;
    float2 *tile = tiles + inst_l * (kFr * PITCH);
    const long long inst = (long long)blockIdx.x * INST + inst_l;
    const long long t0 = inst * a.tiles_per_inst;
    long long ntl = (long long)a.n_tiles - t0;
    if (ntl > a.tiles_per_inst) ntl = a.tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;
    if (nt == 0) return;
    const unsigned int imask = (M == 32) ? 0xffffffffu : (((1u << M) - 1u) << ((threadIdx.x & 31) / M * M));
    const float2 *x = a.x + (long long)blockIdx.y * a.stream_stride;
    float2 *y = a.y + (long long)blockIdx.y * a.stream_stride;

    float h[P];
#pragma unroll
    for (int n = 0; n < P; n++) h[n] = __ldg(a.taps + n * M + (M - 1 - u));   // branch M-1-u feeds input u
    float2 win[HL + kFr], nx[kFr];
    auto fetch = [&](int it) {
        const long long f0 = (t0 + it) * kFr;
#pragma unroll
        for (int f = 0; f < kFr; f++) nx[f] = (it < nt && f0 + f < a.n_frames) ? __ldg(x + (f0 + f) * M + u) : make_float2(0.f, 0.f);
    };
    fetch(0);
    {
        const long long f0 = t0 * kFr;
#pragma unroll
        for (int i = 0; i < HL; i++) win[i] = ld_hist0(a.hist, x, (f0 - HL + i) * M + u, a.H);
    }
    for (int it = 0; it < nt; it++) {
        const long long f0 = (t0 + it) * kFr;
#pragma unroll
        for (int f = 0; f < kFr; f++) win[HL + f] = nx[f];
        fetch(it + 1);
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int n = P - 1; n >= 0; n--) acc = __ffma2_rn(bc(h[n]), win[HL + f - n], acc);
            tile[f * PITCH + u] = acc;
        }
#pragma unroll
        for (int i = 0; i < HL; i++) win[i] = win[kFr + i];
        __syncwarp(imask);
        if (M <= kFr) {
#pragma unroll
            for (int i = 0; i < (kFr / M > 0 ? kFr / M : 1); i++) transform_row_fwd<M>(tile + (u + M * i) * PITCH);
        } else if (u < kFr) {
            transform_row_fwd<M>(tile + u * PITCH);
        }
        __syncwarp(imask);
        // the tile is 16 M contiguous outputs = 8 M float4: float4 w = i M + u is row 2i + 2u / M, columns 2u % M, + 1
        float4 *yt = reinterpret_cast<float4 *>(y + f0 * M);
#pragma unroll
        for (int i = 0; i < kFr / 2; i++) {
            const int r = 2 * i + (2 * u) / M, c = (2 * u) % M;
            if (f0 + r < a.n_frames) __stcs(yt + i * M + u, *reinterpret_cast<const float4 *>(tile + r * PITCH + c));
        }
        __syncwarp(imask);
    }
}

// M = 4: one thread per frame (frame = 32 bytes in, 32 bytes out), the P frames its windows span re-read through L1
template <int P>
__global__ void __launch_bounds__(kThreads) pfbch_frame4_kernel(const PfbSmallArgs a)
{
    constexpr int M = 4;
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= a.n_frames) return;
    const float2 *x = a.x + (long long)blockIdx.y * a.stream_stride;
    float2 *y = a.y + (long long)blockIdx.y * a.stream_stride;
    float2 X[M];
#pragma unroll
    for (int u = 0; u < M; u++) X[u] = make_float2(0.f, 0.f);
#pragma unroll
    for (int n = P - 1; n >= 0; n--) {
        const long long f = t - n;
        float2 s[M];
        if (f >= 0) {
            const float4 *p = reinterpret_cast<const float4 *>(x + f * M);
            const float4 q0 = __ldg(p), q1 = __ldg(p + 1);
            s[0] = make_float2(q0.x, q0.y);
            s[1] = make_float2(q0.z, q0.w);
            s[2] = make_float2(q1.x, q1.y);
            s[3] = make_float2(q1.z, q1.w);
        } else {
#pragma unroll
            for (int u = 0; u < M; u++) s[u] = ld_hist0(a.hist, x, f * M + u, a.H);
        }
#pragma unroll
        for (int u = 0; u < M; u++) X[u] = __ffma2_rn(bc(__ldg(a.taps + n * M + (M - 1 - u))), s[u], X[u]);
    }
    RegDft<M, -1>::run(X);
    float4 *yo = reinterpret_cast<float4 *>(y + t * M);
    __stcs(yo, make_float4(X[0].x, X[0].y, X[1].x, X[1].y));
    __stcs(yo + 1, make_float4(X[2].x, X[2].y, X[3].x, X[3].y));
}

template <int M, int P>
void launch_pfb_small_MP(const PfbSmallArgs &a, dim3 grid, cudaStream_t st)
{
    constexpr size_t smem = sizeof(float2) * (size_t)(kThreads / M) * kFr * kPitch<M>;
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch_small_kernel<M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pfbch_small_kernel<M, P><<<grid, kThreads, smem, st>>>(a);
}

// grid: enough instances to fill the GPU, at most 64 tiles (1024 frames) each
void plan(SmallArgs &a, unsigned int M, long long n_frames, unsigned int &grid, long long min_tpi)
{
    const long long n_tiles = (n_frames + kFr - 1) / kFr;
    const long long inst_per_cta = kThreads / (M / 2);
    long long tpi = 64;
    while (tpi > min_tpi && (n_tiles / tpi) < 148ll * 4 * inst_per_cta) tpi >>= 1;
    a.n_frames = n_frames;
    a.n_tiles = (int)n_tiles;
    a.tiles_per_inst = (int)tpi;
    const long long n_inst = (n_tiles + tpi - 1) / tpi;
    grid = (unsigned int)((n_inst + inst_per_cta - 1) / inst_per_cta);
}

template <int M>
constexpr size_t small_smem_bytes() { return sizeof(float2) * (size_t)(kThreads / (M / 2)) * kFr * kPitch<M>; }

template <int M, int L>
void launch_analyzer_ML(const SmallArgs &a, unsigned int grid, cudaStream_t st)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_small_kernel<M, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes<M>()));
    pfbch2_small_kernel<M, L><<<grid, kThreads, small_smem_bytes<M>(), st>>>(a);
}

template <int M>
void launch_analyzer_M(const SmallArgs &a, unsigned int L, unsigned int grid, cudaStream_t st)
{
    switch (L) {
    case 4: launch_analyzer_ML<M, 4>(a, grid, st); break;
    case 6: launch_analyzer_ML<M, 6>(a, grid, st); break;
    case 8: launch_analyzer_ML<M, 8>(a, grid, st); break;
    case 10: launch_analyzer_ML<M, 10>(a, grid, st); break;
    case 12: launch_analyzer_ML<M, 12>(a, grid, st); break;
    case 14: launch_analyzer_ML<M, 14>(a, grid, st); break;
    default: lc::fatal("launch_pfbch2_small", "unsupported tap count");
    }
}

template <int M, int NT>
void launch_synth_MN(const SmallArgs &a, unsigned int grid, cudaStream_t st)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured)) {
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_small_synth_kernel<M, NT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes<M>()));
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_small_synth_kernel<M, NT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes<M>()));
    }
    if (a.par0) pfbch2_small_synth_kernel<M, NT, 1><<<grid, kThreads, small_smem_bytes<M>(), st>>>(a);
    else        pfbch2_small_synth_kernel<M, NT, 0><<<grid, kThreads, small_smem_bytes<M>(), st>>>(a);
}

template <int M>
void launch_synth_M(const SmallArgs &a, unsigned int NT, unsigned int grid, cudaStream_t st)
{
    switch (NT) {
    case 8: launch_synth_MN<M, 8>(a, grid, st); break;
    case 12: launch_synth_MN<M, 12>(a, grid, st); break;
    case 16: launch_synth_MN<M, 16>(a, grid, st); break;
    case 20: launch_synth_MN<M, 20>(a, grid, st); break;
    case 24: launch_synth_MN<M, 24>(a, grid, st); break;
    case 28: launch_synth_MN<M, 28>(a, grid, st); break;
    default: lc::fatal("launch_pfbch2_small_synth", "unsupported tap count");
    }
}

}  // namespace

namespace lc {
namespace pfb {

bool small_pfbch2_supported(unsigned int M, unsigned int L)
{
    return (M == 4 || M == 8 || M == 16 || M == 32) && (L == 4 || L == 6 || L == 8 || L == 10 || L == 12 || L == 14);
}

bool small_pfbch_supported(unsigned int M, unsigned int p)
{
    return ((M == 4 || M == 8 || M == 32) && (p == 4 || p == 6 || p == 8 || p == 10 || p == 12)) || ((M == 8 || M == 32) && p == 14);   // (M = 4, p = 14: the generic kernel measured faster)
}

long long launch_pfbch_small(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, float2 *d_y, unsigned int M,
                             unsigned int p, unsigned long long n_frames, unsigned int n_streams, long long stream_stride,
                             cudaStream_t st)
{
    if (n_frames == 0 || n_frames > (1ull << 34) || n_streams == 0 || n_streams > 65535) return 0;
    PfbSmallArgs a = {};
    a.hist = d_hist;
    a.x = d_x;
    a.y = d_y;
    a.taps = d_taps;
    a.H = H;
    a.n_frames = (long long)n_frames;
    a.stream_stride = stream_stride;
    if (M == 4) {
        dim3 g((unsigned int)((n_frames + kThreads - 1) / kThreads), n_streams, 1);
        if (p == 4) pfbch_frame4_kernel<4><<<g, kThreads, 0, st>>>(a);
        else if (p == 6) pfbch_frame4_kernel<6><<<g, kThreads, 0, st>>>(a);
        else if (p == 8) pfbch_frame4_kernel<8><<<g, kThreads, 0, st>>>(a);
        else if (p == 10) pfbch_frame4_kernel<10><<<g, kThreads, 0, st>>>(a);
        else if (p == 12) pfbch_frame4_kernel<12><<<g, kThreads, 0, st>>>(a);
        else pfbch_frame4_kernel<14><<<g, kThreads, 0, st>>>(a);
        LC_LAUNCH_CHECK();
        return (long long)n_frames;
    }
    const long long n_tiles = ((long long)n_frames + kFr - 1) / kFr;
    const long long inst_per_cta = kThreads / M;
    long long tpi = 64;
    while (tpi > 1 && (n_tiles / tpi) * n_streams < 148ll * 4 * inst_per_cta) tpi >>= 1;
    a.n_tiles = (int)n_tiles;
    a.tiles_per_inst = (int)tpi;
    const long long n_inst = (n_tiles + tpi - 1) / tpi;
    dim3 grid((unsigned int)((n_inst + inst_per_cta - 1) / inst_per_cta), n_streams, 1);
#define LC_PFB(MM, PP) launch_pfb_small_MP<MM, PP>(a, grid, st)
    if (M == 8 && p == 4) LC_PFB(8, 4);
    else if (M == 8 && p == 6) LC_PFB(8, 6);
    else if (M == 8 && p == 8) LC_PFB(8, 8);
    else if (M == 8 && p == 10) LC_PFB(8, 10);
    else if (M == 8 && p == 12) LC_PFB(8, 12);
    else if (M == 8 && p == 14) LC_PFB(8, 14);
    else if (M == 32 && p == 4) LC_PFB(32, 4);
    else if (M == 32 && p == 6) LC_PFB(32, 6);
    else if (M == 32 && p == 8) LC_PFB(32, 8);
    else if (M == 32 && p == 10) LC_PFB(32, 10);
    else if (M == 32 && p == 12) LC_PFB(32, 12);
    else if (M == 32 && p == 14) LC_PFB(32, 14);
    else fatal("launch_pfbch_small", "unsupported shape");
#undef LC_PFB
    LC_LAUNCH_CHECK();
    return (long long)n_frames;
}

bool small_pfbch2_synth_supported(unsigned int M, unsigned int n_taps)
{
    return ((M == 4 || M == 8 || M == 16 || M == 32) && (n_taps == 8 || n_taps == 12 || n_taps == 16)) ||
           ((M == 8 || M == 16 || M == 32) && (n_taps == 20 || n_taps == 24 || n_taps == 28));
}

long long launch_pfbch2_small(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps_scaled, float2 *d_y,
                              unsigned int M, unsigned int L, unsigned long long n_frames, int parity0, cudaStream_t st)
{
    if (n_frames == 0 || n_frames > (1ull << 34)) return 0;
    SmallArgs a = {};
    a.hist = d_hist;
    a.x = d_x;
    a.y = d_y;
    a.taps = d_taps_scaled;
    a.H = H;
    a.par0 = parity0 & 1;
    unsigned int grid = 0;
    plan(a, M, (long long)n_frames, grid, 1);
    if (M == 4) {                                                 // (M = 8 measured 0.45 here against 0.57 with the instance kernel)
        if (n_frames > (1ull << 31) * kThreads / 2) return 0;
        const unsigned int g = (unsigned int)((n_frames + kThreads - 1) / kThreads);
#define LC_FRAME(MM, LL) pfbch2_frame_kernel<MM, LL><<<g, kThreads, 0, st>>>(a)
        if (L == 4) LC_FRAME(4, 4);
        else if (L == 6) LC_FRAME(4, 6);
        else if (L == 8) LC_FRAME(4, 8);
        else if (L == 10) LC_FRAME(4, 10);
        else if (L == 12) LC_FRAME(4, 12);
        else if (L == 14) LC_FRAME(4, 14);
        else fatal("launch_pfbch2_small", "unsupported shape");
#undef LC_FRAME
        LC_LAUNCH_CHECK();
        return (long long)n_frames;
    }
    switch (M) {
    case 4: launch_analyzer_M<4>(a, L, grid, st); break;
    case 8: launch_analyzer_M<8>(a, L, grid, st); break;
    case 16: launch_analyzer_M<16>(a, L, grid, st); break;
    case 32: launch_analyzer_M<32>(a, L, grid, st); break;
    default: fatal("launch_pfbch2_small", "unsupported channel count");
    }
    LC_LAUNCH_CHECK();
    return (long long)n_frames;
}

long long launch_pfbch2_small_synth(const float2 *d_hist_rows, int hist_rows, const float2 *d_x, const float *d_taps,
                                    float2 *d_y, float2 *d_hist_out, unsigned int M, unsigned int n_taps,
                                    unsigned long long n_frames, int parity0, cudaStream_t st)
{
    if (n_frames < n_taps || n_frames > (1ull << 34)) return 0;   // the state rows all come from this call
    SmallArgs a = {};
    a.hist = d_hist_rows;
    a.H = hist_rows;
    a.x = d_x;
    a.y = d_y;
    a.hist_out = d_hist_out;
    a.taps = d_taps;
    a.par0 = parity0 & 1;
    unsigned int grid = 0;
    plan(a, M, (long long)n_frames, grid, 2);                     // the warm-up reaches one tile back
    if (M == 4) {
        const unsigned int g = (unsigned int)((n_frames + kThreads - 1) / kThreads);
#define LC_FRAME(MM, NN) pfbch2_frame_synth_kernel<MM, NN><<<g, kThreads, 0, st>>>(a)
        if (n_taps == 8) LC_FRAME(4, 8);
        else if (n_taps == 12) LC_FRAME(4, 12);
        else if (n_taps == 16) LC_FRAME(4, 16);
        else fatal("launch_pfbch2_small_synth", "unsupported shape");
#undef LC_FRAME
        LC_LAUNCH_CHECK();
        return (long long)n_frames;
    }
    switch (M) {
    case 4: launch_synth_M<4>(a, n_taps, grid, st); break;
    case 8: launch_synth_M<8>(a, n_taps, grid, st); break;
    case 16: launch_synth_M<16>(a, n_taps, grid, st); break;
    case 32: launch_synth_M<32>(a, n_taps, grid, st); break;
    default: fatal("launch_pfbch2_small_synth", "unsupported channel count");
    }
    LC_LAUNCH_CHECK();
    return (long long)n_frames;
}

}  // namespace pfb
}  // namespace lc
