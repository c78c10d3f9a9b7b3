// firfilt_kernels.cu -- FIR block filter, batched dot product and firpfb kernels.
//
// firfilt (src/filter/src/firfilt.c:346-358 = push + dotprod per sample) as a tiled convolution:
// a CTA stages TILE + n_taps - 1 input samples of one stream in shared memory with coalesced
// loads, the taps sit in shared memory (broadcast reads), each thread produces OPT consecutive
// outputs from a register window that slides one sample per tap.
#include "firfilt_kernels.cuh"

#include <cstdint>

namespace {

constexpr int kThreads = 256;
constexpr int kOpt = 4;                       // outputs per thread
constexpr int kTile = kThreads * kOpt;        // outputs per CTA tile

// smem: taps[n_taps] floats, then samples [kTile + n_taps - 1] float2 stored transposed so that
// thread t reading sample t*kOpt + j is conflict-free: slot(q) = (q % kOpt) * row + q / kOpt
__global__ void __launch_bounds__(kThreads)
firfilt_tiled_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                     const float *__restrict__ taps, unsigned int n_taps, float scale,
                     float2 *__restrict__ y, unsigned long long n, unsigned long long tiles_per_stream)
{
    extern __shared__ float smem_f[];
    float *h_s = smem_f;
    const unsigned int n_samp = kTile + n_taps - 1;
    const unsigned int row = (n_samp + kOpt - 1) / kOpt + 1;
    float2 *xs = reinterpret_cast<float2 *>(smem_f + ((n_taps + 1) & ~1u));

    const unsigned long long stream = blockIdx.x / tiles_per_stream;
    const unsigned long long tile = blockIdx.x % tiles_per_stream;
    const long long o0 = (long long)(tile * kTile);             // first output of the tile
    const float2 *xstr = x + stream * n;
    float2 *ystr = y + stream * n;

    for (unsigned int k = threadIdx.x; k < n_taps; k += kThreads) h_s[k] = taps[k];
    // sample q of the tile is stream sample o0 - (n_taps-1) + q
    for (unsigned int q = threadIdx.x; q < n_samp; q += kThreads) {
        long long idx = o0 - (long long)(n_taps - 1) + q;
        float2 v = make_float2(0.f, 0.f);
        if (idx >= (long long)n) v = make_float2(0.f, 0.f);
        else if (idx >= 0) v = __ldg(xstr + idx);
        else if (hist != nullptr && H + idx >= 0) v = __ldg(hist + (H + idx));
        xs[(q % kOpt) * row + q / kOpt] = v;
    }
    __syncthreads();

    // outputs o0 + t*kOpt + r, r < kOpt.  y[o] = sum_k h[k] xs_lin[o - o0 + (n_taps-1) - k]
    const unsigned int t = threadIdx.x;
    float2 acc[kOpt];
    float2 win[kOpt];                        // win[r] = xs_lin[t*kOpt + r + (n_taps-1) - k]
#pragma unroll
    for (int r = 0; r < kOpt; r++) {
        acc[r] = make_float2(0.f, 0.f);
        unsigned int q = t * kOpt + r + (n_taps - 1);
        win[r] = xs[(q % kOpt) * row + q / kOpt];
    }
    for (unsigned int k = 0; k < n_taps; k++) {
        const float c = h_s[k];
#pragma unroll
        for (int r = 0; r < kOpt; r++) {
            acc[r].x = fmaf(c, win[r].x, acc[r].x);
            acc[r].y = fmaf(c, win[r].y, acc[r].y);
        }
        // slide one sample towards the past
#pragma unroll
        for (int r = kOpt - 1; r > 0; r--) win[r] = win[r - 1];
        if (k + 1 < n_taps) {
            unsigned int q = t * kOpt + (n_taps - 1) - (k + 1);
            win[0] = xs[(q % kOpt) * row + q / kOpt];
        }
    }
#pragma unroll
    for (int r = 0; r < kOpt; r++) {
        long long o = o0 + (long long)t * kOpt + r;
        if (o < (long long)n) ystr[o] = make_float2(acc[r].x * scale, acc[r].y * scale);
    }
}

// ---------------------------------------------------------------------------------------------
// Register-blocked FIR for short filters (NT <= 64 taps, zero-padded up to NT): the FP32-pipe-bound
// shape of BASELINE config 2.  Each thread produces RO = 16 consecutive outputs; the NT taps sit in
// registers (the same in every thread), samples come from a staged tile: one LDS.64 per up to RO packed
// FFMA2 (tap broadcast x {re, im} pair), NT + RO - 1 loads for NT * RO multiply-adds.  Everything is
// unrolled, so tap and accumulator indices are compile-time constants:  y[o+r] += h[k] x[o+r-k]  is
// visited sample by sample (s = o+r-k).
//
// Every warp is its own pipeline: it owns a private double buffer, fetches its own 32*RO + NT-1 samples
// with cp.async (256-byte runs) and only ever reads what it fetched itself, so cp.async.wait_group +
// __syncwarp is all the synchronisation there is -- no block barrier, warps never wait for each other.
// The NT-1 samples of overlap between neighbouring warp tiles are re-read through L2, not from HBM.
constexpr int kRbThreads = 256;
constexpr int kRbWarps = kRbThreads / 32;

template <int NT, int RO>
struct RbCfg {
    static constexpr int kWarpTile = 32 * RO;               // outputs per warp tile
    static constexpr int NS = kWarpTile + NT - 1;           // staged samples per warp tile
    // staged transposed so that lane l reading sample l*RO + d is conflict-free:
    //   slot(q) = (q % RO) * ROW + q / RO
    // and ROW odd makes the 8-byte cp.async fill (a half-warp = 16 consecutive q = 16 rows) conflict-free too
    static constexpr int ROW = ((NS + RO - 1) / RO) | 1;
    static constexpr int NPASS = (NS + 31) / 32;
    static constexpr unsigned int kBufBytes = RO * ROW * 8;
    static constexpr size_t kSmem = (size_t)kRbWarps * 2 * kBufBytes;
    static_assert(RO == 16, "fill pattern is written for 16 outputs per thread");
};

__device__ __forceinline__ void cp_async8_zfill(unsigned int dst_smem, const float2 *src, bool valid)
{
    // 8-byte asynchronous global->shared copy; when !valid nothing is read and zeros are written
    const int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8_plain(unsigned int dst_smem, const float2 *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}

// persistent: every warp draws its next tile from a global counter (one atomicAdd per 512 outputs), so warps
// that the schedulers favour simply do more tiles and all of them finish together; with static striding
// (tile_counter == nullptr) a quarter of the warp slots sat empty on average while the slowest warps finished
template <int NT, int RO>
__global__ void __launch_bounds__(kRbThreads, 2)
firfilt_regblock_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                        const float *__restrict__ taps, unsigned int n_taps, float scale,
                        float2 *__restrict__ y, unsigned long long n_, unsigned int tiles_per_stream_,
                        unsigned int n_tiles_, unsigned int *__restrict__ tile_counter)
{
    using C = RbCfg<NT, RO>;
    constexpr int ROW = C::ROW, NS = C::NS, NPASS = C::NPASS, kWarpTile = C::kWarpTile;
    extern __shared__ float2 xs_dyn[];          // per warp: 2 buffers of RO * ROW samples
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int xs_warp = (unsigned int)__cvta_generic_to_shared(xs_dyn) + (unsigned int)warp * (2 * C::kBufBytes);
    // sample q = lane + 32 i of a tile goes to slot (lane % RO) * ROW + lane / RO + i * (32 / RO)
    const unsigned int fill_thr = xs_warp + (unsigned int)((lane % RO) * ROW + lane / RO) * 8;
    const unsigned int read_thr = xs_warp + (unsigned int)lane * 8;
    // keep the loop-invariant parameters in registers (constant-bank reloads inside the loop stall it)
    long long n = (long long)n_;
    unsigned int tiles_per_stream = tiles_per_stream_, n_tiles = n_tiles_;
    asm volatile("" : "+l"(n), "+r"(tiles_per_stream), "+r"(n_tiles));
    const unsigned int stride = gridDim.x * kRbWarps;

    float h[NT];
#pragma unroll
    for (int k = 0; k < NT; k++) h[k] = (k < (int)n_taps) ? __ldg(taps + k) * scale : 0.0f;

    auto prefetch = [&](unsigned int tile_id, unsigned int dst) {
        const unsigned int stream = tile_id / tiles_per_stream;
        const long long first = (long long)(tile_id - stream * tiles_per_stream) * kWarpTile - (long long)(NT - 1);
        const float2 *xstr = x + stream * (unsigned long long)n;
        if (first >= 0 && first + NS <= n) {
            // interior tile: every sample comes from the stream; addresses are base + immediates
            const float2 *src = xstr + first + lane;
#pragma unroll
            for (int i = 0; i < NPASS; i++)
                if (i < NPASS - 1 || lane + i * 32 < NS)
                    cp_async8_plain(dst + (unsigned int)(i * (32 / RO)) * 8, src + i * 32);
        } else {
            // first / last tile of a stream: history or zeros in front, zeros past the end
#pragma unroll 1
            for (int i = 0; i < NPASS; i++) {
                const int q = lane + i * 32;
                if (q < NS) {
                    const long long idx = first + q;
                    const float2 *src = xstr;
                    bool valid = false;
                    if (idx >= 0 && idx < n) { src = xstr + idx; valid = true; }
                    else if (idx < 0 && hist != nullptr && H + idx >= 0) { src = hist + (H + idx); valid = true; }
                    cp_async8_zfill(dst + (unsigned int)(i * (32 / RO)) * 8, src, valid);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // (tile ids fit 32 bits: the launcher checks; n_tiles + stride does not wrap because both are < 2^31)
    auto next_tile = [&](unsigned int current) -> unsigned int {
        if (tile_counter == nullptr) return current + stride;
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    unsigned int tile_id = (tile_counter == nullptr) ? blockIdx.x * kRbWarps + warp : next_tile(0);
    unsigned int buf = 0;                        // byte offset of the current buffer: 0 or kBufBytes
    if (tile_id < n_tiles) prefetch(tile_id, fill_thr);
    while (tile_id < n_tiles) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();                            // this tile is visible to the whole warp, and every lane has
                                                 // finished reading the other buffer (previous tile)
        const unsigned int upcoming = next_tile(tile_id);
        if (upcoming < n_tiles) prefetch(upcoming, fill_thr + (buf ^ C::kBufBytes));

        const unsigned int xb = read_thr + buf;
        float2 acc[RO];
#pragma unroll
        for (int r = 0; r < RO; r++) acc[r] = make_float2(0.f, 0.f);
        // tile-relative sample i = lane*RO + d, d = 0 .. NT+RO-2, feeds output r with tap k = (NT-1) + r - d
#pragma unroll
        for (int d = 0; d < NT - 1 + RO; d++) {
            float2 sv;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(sv.x), "=f"(sv.y) : "r"(xb + (unsigned int)((d % RO) * ROW + d / RO) * 8));
#pragma unroll
            for (int r = 0; r < RO; r++) {
                const int k = (NT - 1) + r - d;
                if (k >= 0 && k < NT) acc[r] = __ffma2_rn(make_float2(h[k], h[k]), sv, acc[r]);
            }
        }
        // the scale is folded into the taps: (h[k] * scale) * x differs from scale * (sum h x) only by
        // float rounding of the products (well inside the parity tolerance of the tests)
        const unsigned int stream = tile_id / tiles_per_stream;
        const long long o0 = (long long)(tile_id - stream * tiles_per_stream) * kWarpTile + (long long)lane * RO;
        float2 *yo = y + stream * (unsigned long long)n + o0;
        const long long t_first = o0 - (long long)lane * RO;     // first output of the warp tile
        if (t_first + kWarpTile <= n) {
            // whole tile inside the stream: the lanes' 128-byte runs go through the (consumed) sample buffer so that every
            // store instruction writes 512 contiguous bytes (see firfilt_ffa_kernel)
            __syncwarp();
#pragma unroll
            for (int r = 0; r < RO; r++)
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(xb + (unsigned int)(r * ROW) * 8), "f"(acc[r].x), "f"(acc[r].y) : "memory");
            __syncwarp();
            float4 *yt = reinterpret_cast<float4 *>(y + stream * (unsigned long long)n + t_first);
            const unsigned int xw = xs_warp + buf;
#pragma unroll
            for (int i = 0; i < kWarpTile / 64; i++) {
                const int q = 64 * i + 2 * lane;
                float2 a0, a1;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a0.x), "=f"(a0.y) : "r"(xw + (unsigned int)((q % RO) * ROW + q / RO) * 8));
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a1.x), "=f"(a1.y) : "r"(xw + (unsigned int)(((q + 1) % RO) * ROW + (q + 1) / RO) * 8));
                __stcs(yt + 32 * i + lane, make_float4(a0.x, a0.y, a1.x, a1.y));
            }
        } else if (o0 + RO <= n) {
#pragma unroll
            for (int r = 0; r < RO; r += 2)
                __stcs(reinterpret_cast<float4 *>(yo + r), make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y));
        } else {
#pragma unroll
            for (int r = 0; r < RO; r++)
                if (o0 + r < n) yo[r] = acc[r];
        }
        tile_id = upcoming;
        buf ^= C::kBufBytes;
    }
}

// ---------------------------------------------------------------------------------------------
// 2-parallel fast-FIR form of the kernel above for NT = 64 (BASELINE config 2), which is FP32-pipe bound: with
// x0[k] = x[2k], x1[k] = x[2k+1], H0[i] = h[2i], H1[i] = h[2i+1] (32 taps each),
//     A = H0 * x0,   B = H1 * x1,   C = (H0 + H1) * (x0 + x1)      (three 32-tap convolutions per output PAIR)
//     y[2k] = A[k] + B[k-1],        y[2k+1] = C[k] - A[k] - B[k]
// which is 800 packed FFMA2 + 55 packed adds per 16 outputs instead of 1024 FFMA2.  A lane still owns 16 consecutive
// outputs (8 pairs) of a warp tile and the same staged samples; the three tap sets sit in shared memory and pass
// through a short register window (a sample meets 8 - 9 consecutive taps, one new tap per sample: LDS.32 broadcast), so
// the taps need a dozen registers instead of 64.  Accumulator plan (17 pairs live at most):
// b[0..8] = B[k-1 .. k+7]; y0[p] starts as B[p-1] and takes A[p]; t[p] starts as B[p-1] - B[p] and takes C[p];
// y1[p] = t[p] - y0[p].  The sums H0 + H1 and x0 + x1 are rounded once more than the direct form (errors of a few 1e-7
// of the output scale: inside the parity tolerance, tested against the oracle and the direct kernel).
template <int NT>
__global__ void __launch_bounds__(kRbThreads, 2)
firfilt_ffa_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                   const float *__restrict__ taps, unsigned int n_taps, float scale,
                   float2 *__restrict__ y, unsigned long long n_, unsigned int tiles_per_stream_,
                   unsigned int n_tiles_, unsigned int *__restrict__ tile_counter)
{
    constexpr int RO = 16, HT = NT / 2, NP = RO / 2;            // HT taps per sub-filter, NP output pairs per lane
    using C = RbCfg<NT, RO>;
    constexpr int ROW = C::ROW, NS = C::NS, NPASS = C::NPASS, kWarpTile = C::kWarpTile;
    extern __shared__ float2 xs_dyn[];          // per warp: 2 buffers of RO * ROW samples; then the three tap sets
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int xs_warp = (unsigned int)__cvta_generic_to_shared(xs_dyn) + (unsigned int)warp * (2 * C::kBufBytes);
    float *tap_s = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(xs_dyn) + C::kSmem);   // [3][HT]: H0, H1, H0 + H1
    const unsigned int tap_u32 = (unsigned int)__cvta_generic_to_shared(tap_s);
    const unsigned int fill_thr = xs_warp + (unsigned int)((lane % RO) * ROW + lane / RO) * 8;
    const unsigned int read_thr = xs_warp + (unsigned int)lane * 8;
    long long n = (long long)n_;
    unsigned int tiles_per_stream = tiles_per_stream_, n_tiles = n_tiles_;
    asm volatile("" : "+l"(n), "+r"(tiles_per_stream), "+r"(n_tiles));
    const unsigned int stride = gridDim.x * kRbWarps;

    for (int i = threadIdx.x; i < HT; i += kRbThreads) {
        const float h0 = (2 * i < (int)n_taps) ? __ldg(taps + 2 * i) * scale : 0.0f;
        const float h1 = (2 * i + 1 < (int)n_taps) ? __ldg(taps + 2 * i + 1) * scale : 0.0f;
        tap_s[i] = h0;
        tap_s[HT + i] = h1;
        tap_s[2 * HT + i] = h0 + h1;
    }
    __syncthreads();

    auto prefetch = [&](unsigned int tile_id, unsigned int dst) {
        const unsigned int stream = tile_id / tiles_per_stream;
        const long long first = (long long)(tile_id - stream * tiles_per_stream) * kWarpTile - (long long)(NT - 1);
        const float2 *xstr = x + stream * (unsigned long long)n;
        if (first >= 0 && first + NS <= n) {
            const float2 *src = xstr + first + lane;
#pragma unroll
            for (int i = 0; i < NPASS; i++)
                if (i < NPASS - 1 || lane + i * 32 < NS)
                    cp_async8_plain(dst + (unsigned int)(i * (32 / RO)) * 8, src + i * 32);
        } else {
#pragma unroll 1
            for (int i = 0; i < NPASS; i++) {
                const int q = lane + i * 32;
                if (q < NS) {
                    const long long idx = first + q;
                    const float2 *src = xstr;
                    bool valid = false;
                    if (idx >= 0 && idx < n) { src = xstr + idx; valid = true; }
                    else if (idx < 0 && hist != nullptr && H + idx >= 0) { src = hist + (H + idx); valid = true; }
                    cp_async8_zfill(dst + (unsigned int)(i * (32 / RO)) * 8, src, valid);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto next_tile = [&](unsigned int current) -> unsigned int {
        if (tile_counter == nullptr) return current + stride;
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    // tile-relative sample d of this lane (d = 0 .. NT + RO - 2) is stream sample o0 - (NT-1) + d
    auto sample = [&](unsigned int xb, int d) -> float2 {
        float2 sv;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(sv.x), "=f"(sv.y) : "r"(xb + (unsigned int)((d % RO) * ROW + d / RO) * 8));
        return sv;
    };
    auto tap = [&](int set, int i) -> float {
        float t;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(tap_u32 + (unsigned int)(set * HT + i) * 4));
        return t;
    };

    unsigned int tile_id = (tile_counter == nullptr) ? blockIdx.x * kRbWarps + warp : next_tile(0);
    unsigned int buf = 0;
    if (tile_id < n_tiles) prefetch(tile_id, fill_thr);
    while (tile_id < n_tiles) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const unsigned int upcoming = next_tile(tile_id);
        if (upcoming < n_tiles) prefetch(upcoming, fill_thr + (buf ^ C::kBufBytes));
        const unsigned int xb = read_thr + buf;

        // With K0 the pair index of the lane's first output: x1 sample e (e = 0 .. HT+NP-1) = x1[K0 - HT + e] is tile sample
        // d = 2e, x0 sample e (e = 0 .. HT+NP-2) = x0[K0 - HT + 1 + e] is tile sample d = 2e + 1.
        // ---- B[j], j = -1 .. NP-1 (b[j + 1]):  B[K0 + j] = sum_i H1[i] x1[K0 + j - i], i.e. x1 sample e meets tap i = HT + j - e
        float2 b[NP + 1];
#pragma unroll
        for (int j = 0; j <= NP; j++) b[j] = make_float2(0.f, 0.f);
        {
            // samples and taps are loaded kAhead steps before they are used (everything is unrolled, so the small
            // rings below are plain registers): the FFMA2 bursts of a step never wait for their own loads
            constexpr int kAhead = 4, NE = HT + NP;
            float h1w[HT + NP + 1];                              // h1w[i + NP] = H1[i]: only a window of NP + 1 + kAhead is live
            float2 sv[NE];
#pragma unroll
            for (int e = 0; e < kAhead; e++) {
                sv[e] = sample(xb, 2 * e);
                if (HT - 1 - e >= 0) h1w[HT - 1 - e + NP] = tap(1, HT - 1 - e);
            }
#pragma unroll
            for (int e = 0; e < NE; e++) {
                if (e + kAhead < NE) {
                    sv[e + kAhead] = sample(xb, 2 * (e + kAhead));
                    // the tap that enters the window at sample e': i = HT - 1 - e' (meets j = -1)
                    if (HT - 1 - (e + kAhead) >= 0) h1w[HT - 1 - (e + kAhead) + NP] = tap(1, HT - 1 - (e + kAhead));
                }
#pragma unroll
                for (int j = -1; j < NP; j++) {
                    const int i = HT + j - e;
                    if (i >= 0 && i < HT) b[j + 1] = __ffma2_rn(make_float2(h1w[i + NP], h1w[i + NP]), sv[e], b[j + 1]);
                }
            }
        }
        float2 y0[NP], t[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            y0[p] = b[p];                                        // B[p - 1]
            t[p] = __fadd2_rn(b[p], make_float2(-b[p + 1].x, -b[p + 1].y));
        }
        // ---- A[p] into y0[p], C[p] into t[p]: x0 sample e meets tap i = HT - 1 + p - e of H0 (and of H0 + H1, with x0 + x1)
        {
            constexpr int kAhead = 3, NE = HT + NP - 1;
            float h0w[HT + NP], hsw[HT + NP];
            float2 s0[NE], ss[NE];
            auto load_step = [&](int e) {
                s0[e] = sample(xb, 2 * e + 1);
                ss[e] = sample(xb, 2 * e + 2);                   // x1[K0 - HT + 1 + e]: the partner of this x0 sample
                if (HT - 1 - e >= 0) {
                    h0w[HT - 1 - e + NP] = tap(0, HT - 1 - e);
                    hsw[HT - 1 - e + NP] = tap(2, HT - 1 - e);
                }
            };
#pragma unroll
            for (int e = 0; e < kAhead; e++) load_step(e);
#pragma unroll
            for (int e = 0; e < NE; e++) {
                if (e + kAhead < NE) load_step(e + kAhead);
                const float2 sum = __fadd2_rn(s0[e], ss[e]);
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    const int i = HT - 1 + p - e;
                    if (i >= 0 && i < HT) {
                        y0[p] = __ffma2_rn(make_float2(h0w[i + NP], h0w[i + NP]), s0[e], y0[p]);
                        t[p] = __ffma2_rn(make_float2(hsw[i + NP], hsw[i + NP]), sum, t[p]);
                    }
                }
            }
        }
        const unsigned int stream = tile_id / tiles_per_stream;
        const long long o0 = (long long)(tile_id - stream * tiles_per_stream) * kWarpTile + (long long)lane * RO;
        float2 *yo = y + stream * (unsigned long long)n + o0;
        const long long t_first = o0 - (long long)lane * RO;     // first output of the warp tile
        if (t_first + kWarpTile <= n) {
            // Whole tile inside the stream: the lanes' 128-byte runs go through the (consumed) sample buffer so that every
            // store instruction writes 512 contiguous bytes.  Stored straight from the registers, one instruction touched 32
            // different lines (16 bytes each, 128 bytes apart), and that -- not the FFMA2 count -- bounded the kernel:
            // the direct and the fast-FIR form both took 5.4 ms per 2^30 samples.
            __syncwarp();                                        // every lane has read its last sample
#pragma unroll
            for (int p = 0; p < NP; p++) {
                const float2 y1 = __fadd2_rn(t[p], make_float2(-y0[p].x, -y0[p].y));
                // output q = lane RO + r sits at slot (q % RO) ROW + q / RO, as the samples did
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(xb + (unsigned int)((2 * p) * ROW) * 8), "f"(y0[p].x), "f"(y0[p].y) : "memory");
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(xb + (unsigned int)((2 * p + 1) * ROW) * 8), "f"(y1.x), "f"(y1.y) : "memory");
            }
            __syncwarp();
            float4 *yt = reinterpret_cast<float4 *>(y + stream * (unsigned long long)n + t_first);
            const unsigned int xw = xs_warp + buf;
#pragma unroll
            for (int i = 0; i < kWarpTile / 64; i++) {
                // outputs q = 64 i + 2 lane, + 1
                const int q = 64 * i + 2 * lane;
                float2 a0, a1;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a0.x), "=f"(a0.y) : "r"(xw + (unsigned int)((q % RO) * ROW + q / RO) * 8));
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a1.x), "=f"(a1.y) : "r"(xw + (unsigned int)(((q + 1) % RO) * ROW + (q + 1) / RO) * 8));
                __stcs(yt + 32 * i + lane, make_float4(a0.x, a0.y, a1.x, a1.y));
            }
        } else {
#pragma unroll
            for (int p = 0; p < NP; p++) {
                const float2 y1 = __fadd2_rn(t[p], make_float2(-y0[p].x, -y0[p].y));
                if (o0 + 2 * p < n) yo[2 * p] = y0[p];
                if (o0 + 2 * p + 1 < n) yo[2 * p + 1] = y1;
            }
        }
        tile_id = upcoming;
        buf ^= C::kBufBytes;
    }
}

__global__ void firfilt_state_output_kernel(const float2 *__restrict__ hist, const float *__restrict__ taps,
                                            unsigned int n_taps, float scale, float2 *__restrict__ y)
{
    // one warp; strided partial sums, then a shuffle reduction
    float re = 0.f, im = 0.f;
    for (unsigned int k = threadIdx.x; k < n_taps; k += 32) {
        float2 s = hist[n_taps - 1 - k];
        re = fmaf(taps[k], s.x, re);
        im = fmaf(taps[k], s.y, im);
    }
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if (threadIdx.x == 0) *y = make_float2(re * scale, im * scale);
}

// one warp per dot product when n >= 32, one thread per dot product otherwise
__global__ void __launch_bounds__(kThreads)
dotprod_batch_kernel(const float *__restrict__ h, unsigned int n, const float2 *__restrict__ x,
                     unsigned long long n_vec, unsigned long long stride, float2 *__restrict__ y, int per_warp)
{
    if (per_warp) {
        const unsigned int lane = threadIdx.x & 31;
        unsigned long long v = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
        for (; v < n_vec; v += nw) {
            const float2 *xv = x + v * stride;
            float re = 0.f, im = 0.f;
            for (unsigned int i = lane; i < n; i += 32) {
                float2 s = __ldg(xv + i);
                float c = __ldg(h + i);
                re = fmaf(c, s.x, re);
                im = fmaf(c, s.y, im);
            }
            for (int o = 16; o > 0; o >>= 1) {
                re += __shfl_xor_sync(0xffffffffu, re, o);
                im += __shfl_xor_sync(0xffffffffu, im, o);
            }
            if (lane == 0) y[v] = make_float2(re, im);
        }
    } else {
        unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        const unsigned long long nt = (unsigned long long)gridDim.x * blockDim.x;
        for (; v < n_vec; v += nt) {
            const float2 *xv = x + v * stride;
            float re = 0.f, im = 0.f;
            for (unsigned int i = 0; i < n; i++) {
                float2 s = __ldg(xv + i);
                float c = __ldg(h + i);
                re = fmaf(c, s.x, re);
                im = fmaf(c, s.y, im);
            }
            y[v] = make_float2(re, im);
        }
    }
}

// n = 8 ... 255 on 16-byte aligned vectors (even n and stride): G = 2 / 4 / 8 / 16 lanes per dot product, each lane reading
// 16 bytes (two samples) at a time so that a group's load instruction is one contiguous run of 16 G bytes, a log2(G)-step
// shuffle reduction, and the warp's 32 / G results stored by its first lanes as one run.  (One warp per vector left every
// lane with two samples and ten shuffles at n = 64: 0.60 of the roofline.)
template <int G>
__global__ void __launch_bounds__(kThreads)
dotprod_group_kernel(const float *__restrict__ h, unsigned int n, const float2 *__restrict__ x, unsigned long long n_vec,
                     unsigned long long stride, float2 *__restrict__ y)
{
    constexpr int VPW = 32 / G;                                       // vectors per warp
    const unsigned int lane = threadIdx.x & 31, sub = lane % G, grp = lane / G;
    unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long n_wv = (n_vec + VPW - 1) / VPW;
    for (; w < n_wv; w += nw) {
        const unsigned long long v = w * VPW + grp;
        float2 acc = make_float2(0.f, 0.f), acc2 = acc;
        if (v < n_vec) {
            const float4 *xv = reinterpret_cast<const float4 *>(x + v * stride);
            const float2 *hv = reinterpret_cast<const float2 *>(h);
#pragma unroll 4
            for (unsigned int i = sub; i < n / 2; i += G) {
                const float4 s = __ldg(xv + i);
                const float2 c = __ldg(hv + i);
                acc = __ffma2_rn(make_float2(c.x, c.x), make_float2(s.x, s.y), acc);
                acc2 = __ffma2_rn(make_float2(c.y, c.y), make_float2(s.z, s.w), acc2);
            }
        }
        float re = acc.x + acc2.x, im = acc.y + acc2.y;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, o);
            im += __shfl_xor_sync(0xffffffffu, im, o);
        }
        // lane g of the warp fetches group g's result and stores it: VPW consecutive outputs in one run
        const float rre = __shfl_sync(0xffffffffu, re, (lane % VPW) * G), rim = __shfl_sync(0xffffffffu, im, (lane % VPW) * G);
        if (lane < VPW && w * VPW + lane < n_vec) y[w * VPW + lane] = make_float2(rre, rim);
    }
}

__global__ void __launch_bounds__(kThreads)
firpfb_block_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                    const float *__restrict__ sub, unsigned int sub_len, const unsigned int *__restrict__ idx,
                    float scale, float2 *__restrict__ y, unsigned long long n)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        // sub[f][sub_len-1-k] multiplies the k-th newest sample (firpfb.c:71-84 stores branches reversed)
        const float *hf = sub + (size_t)__ldg(idx + i) * sub_len;
        float re = 0.f, im = 0.f;
        for (unsigned int k = 0; k < sub_len; k++) {
            float2 s = lcd::ld_stream(hist, x, (long long)i - (long long)(sub_len - 1) + k, H);  // oldest first
            float c = __ldg(hf + k);
            re = fmaf(c, s.x, re);
            im = fmaf(c, s.y, im);
        }
        y[i] = make_float2(re * scale, im * scale);
    }
}

inline unsigned int grid_cap(unsigned long long blocks)
{
    const unsigned long long cap = 148ull * 8 * 16;
    return (unsigned int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace

namespace lc {
namespace firfilt {

void launch(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, unsigned int n_taps,
            float scale, float2 *d_y, unsigned long long n, unsigned int n_streams, cudaStream_t st,
            unsigned int *d_tile_counter)
{
    // short filters on 16-byte aligned, even-length streams: register-blocked packed kernel
    // (tile ids are 32-bit inside the kernel: streams * ceil(n / 512) must stay below 2^31)
    if (n_taps <= 256 && (n % 2) == 0 && ((uintptr_t)d_y % 16) == 0 && ((n + 511) / 512) * n_streams < (1ull << 31) &&
        !getenv("LIQUID_CUDA_NO_REGBLOCK")) {
        int dev = 0, sms = 148;
        LC_CUDA_CHECK(cudaGetDevice(&dev));
        LC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        auto run = [&](auto kern, auto cfg) {
            using C = decltype(cfg);
            const unsigned long long tiles_rb = (n + C::kWarpTile - 1) / C::kWarpTile;     // warp tiles per stream
            const unsigned long long total_rb = tiles_rb * n_streams;
            unsigned long long grid = (unsigned long long)sms * 2;      // two resident CTAs per SM
            const unsigned long long need = (total_rb + kRbWarps - 1) / kRbWarps;
            if (grid > need) grid = need;
            LC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
            if (d_tile_counter) LC_CUDA_CHECK(cudaMemsetAsync(d_tile_counter, 0, sizeof(unsigned int), st));
            kern<<<(unsigned int)grid, kRbThreads, C::kSmem, st>>>(d_hist, d_x, H, d_taps, n_taps, scale, d_y, n,
                                                                    (unsigned int)tiles_rb, (unsigned int)total_rb, d_tile_counter);
        };
        // 2-parallel fast-FIR form (the tap sets behind the sample buffers), instantiated for 16 / 32 / 64 / 96 / 128 / 192 / 256 taps; the direct
        // register-blocked form stays reachable for the parity tests (LIQUID_CUDA_FIRFILT_DIRECT, up to 64 taps)
        auto run_ffa = [&](auto nt_c) {
            constexpr int NT = decltype(nt_c)::value;
            using C = RbCfg<NT, 16>;
            constexpr size_t smem = C::kSmem + 3 * (NT / 2) * sizeof(float);
            const unsigned long long tiles_rb = (n + C::kWarpTile - 1) / C::kWarpTile;
            const unsigned long long total_rb = tiles_rb * n_streams;
            unsigned long long grid = (unsigned long long)sms * 2;
            const unsigned long long need = (total_rb + kRbWarps - 1) / kRbWarps;
            if (grid > need) grid = need;
            LC_CUDA_CHECK(cudaFuncSetAttribute(firfilt_ffa_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // tiles of up to 32 taps take a warp a fraction of a microsecond: drawn one by one from the global counter, all
            // 2368 warps queue on one address (measured at 16 taps: 183 GS/s with the counter, 388 GS/s with static striding;
            // ncu reports 0.33 ms for the launch that CUDA events time at 0.77 ms).  From 64 taps on the counter wins (254 vs 245)
            unsigned int *ctr = NT <= 32 ? nullptr : d_tile_counter;
            if (ctr) LC_CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), st));
            firfilt_ffa_kernel<NT><<<(unsigned int)grid, kRbThreads, smem, st>>>(d_hist, d_x, H, d_taps, n_taps, scale, d_y, n,
                                                                                (unsigned int)tiles_rb, (unsigned int)total_rb, ctr);
        };
        const bool direct = getenv("LIQUID_CUDA_FIRFILT_DIRECT") != nullptr && n_taps <= 64;
        if (direct && n_taps <= 16) run(firfilt_regblock_kernel<16, 16>, RbCfg<16, 16>{});
        else if (direct && n_taps <= 32) run(firfilt_regblock_kernel<32, 16>, RbCfg<32, 16>{});
        else if (direct) run(firfilt_regblock_kernel<64, 16>, RbCfg<64, 16>{});
        else if (n_taps <= 16) run_ffa(std::integral_constant<int, 16>{});
        else if (n_taps <= 32) run_ffa(std::integral_constant<int, 32>{});
        else if (n_taps <= 64) run_ffa(std::integral_constant<int, 64>{});
        else if (n_taps <= 96) run_ffa(std::integral_constant<int, 96>{});
        else if (n_taps <= 128) run_ffa(std::integral_constant<int, 128>{});
        else if (n_taps <= 192) run_ffa(std::integral_constant<int, 192>{});
        else run_ffa(std::integral_constant<int, 256>{});
        LC_LAUNCH_CHECK();
        return;
    }
    const unsigned long long tiles = (n + kTile - 1) / kTile;
    const unsigned int n_samp = kTile + n_taps - 1;
    const unsigned int row = (n_samp + kOpt - 1) / kOpt + 1;
    const size_t smem = sizeof(float) * ((n_taps + 1) & ~1u) + sizeof(float2) * (size_t)row * kOpt;
    if (smem > 200 * 1024) fatal("firfilt_crcf_execute_block", "filter too long for this build (more than ~20000 taps)");
    static std::atomic<unsigned long long> configured{0};
    if (smem > 48 * 1024 && first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(firfilt_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    unsigned long long total = tiles * n_streams;
    const unsigned long long max_grid = 1ull << 30;
    if (total > max_grid) fatal("firfilt_crcf_execute_streams_dev", "too many tiles for one launch");
    firfilt_tiled_kernel<<<(unsigned int)total, kThreads, smem, st>>>(d_hist, d_x, H, d_taps, n_taps, scale, d_y, n, tiles);
    LC_LAUNCH_CHECK();
}

void launch_state_output(const float2 *d_hist, const float *d_taps, unsigned int n_taps, float scale,
                         float2 *d_y, cudaStream_t st)
{
    firfilt_state_output_kernel<<<1, 32, 0, st>>>(d_hist, d_taps, n_taps, scale, d_y);
    LC_LAUNCH_CHECK();
}

void launch_dotprod_batch(const float *d_h, unsigned int n, const float2 *d_x, unsigned long long n_vec,
                          unsigned long long stride, float2 *d_y, cudaStream_t st)
{
    if (n >= 8 && n < 256 && (n % 2) == 0 && (stride % 2) == 0 && ((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_h % 8) == 0) {
        const unsigned int G = n < 32 ? 2 : n < 64 ? 4 : n < 128 ? 8 : 16;
        const unsigned long long n_wv = (n_vec + 32 / G - 1) / (32 / G);
        const unsigned int grid = grid_cap((n_wv * 32 + kThreads - 1) / kThreads);
        if (G == 2) dotprod_group_kernel<2><<<grid, kThreads, 0, st>>>(d_h, n, d_x, n_vec, stride, d_y);
        else if (G == 4) dotprod_group_kernel<4><<<grid, kThreads, 0, st>>>(d_h, n, d_x, n_vec, stride, d_y);
        else if (G == 8) dotprod_group_kernel<8><<<grid, kThreads, 0, st>>>(d_h, n, d_x, n_vec, stride, d_y);
        else dotprod_group_kernel<16><<<grid, kThreads, 0, st>>>(d_h, n, d_x, n_vec, stride, d_y);
        LC_LAUNCH_CHECK();
        return;
    }
    int per_warp = n >= 32;
    unsigned long long threads = per_warp ? n_vec * 32 : n_vec;
    dotprod_batch_kernel<<<grid_cap((threads + kThreads - 1) / kThreads), kThreads, 0, st>>>(d_h, n, d_x, n_vec, stride, d_y, per_warp);
    LC_LAUNCH_CHECK();
}

// The throughput form: a CTA stages the samples of 2048 outputs (+ sub_len - 1 of history; 8-byte cp.async) and the
// whole bank (rows padded to an odd length: the gather by filter index spreads over the banks) in shared memory; one
// thread per output, same fmaf chain as firpfb_block_kernel (oldest sample first): identical bits.
constexpr unsigned int kPfbTile = 2048;
__global__ void __launch_bounds__(kThreads)
firpfb_tile_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H, const float *__restrict__ sub,
                   unsigned int sub_len, unsigned int n_filters, const unsigned int *__restrict__ idx, float scale,
                   float2 *__restrict__ y, unsigned long long n)
{
    extern __shared__ float2 smem_p[];
    const unsigned int halo = sub_len - 1;
    const unsigned int row = sub_len | 1;                                // padded row length of the bank
    float2 *xs = smem_p;                                                 // staged sample q = sample i0 - halo + q
    float *bank = reinterpret_cast<float *>(xs + ((kPfbTile + halo + 1) & ~1u));
    const unsigned long long i0 = (unsigned long long)blockIdx.x * kPfbTile;
    const unsigned int nt = (n - i0 < kPfbTile) ? (unsigned int)(n - i0) : kPfbTile;
    for (unsigned int e = threadIdx.x; e < n_filters * sub_len; e += blockDim.x)
        bank[(e / sub_len) * row + e % sub_len] = __ldg(sub + e);
    for (unsigned int q = threadIdx.x; q < nt + halo; q += blockDim.x) {
        const long long s = (long long)i0 - (long long)halo + q;
        if (s >= 0) {
            const unsigned int dst = (unsigned int)__cvta_generic_to_shared(xs + q);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(x + s) : "memory");
        } else {
            xs[q] = __ldg(hist + (H + s));
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (unsigned int t = threadIdx.x; t < nt; t += blockDim.x) {
        const float *hf = bank + __ldg(idx + i0 + t) * row;
        float re = 0.f, im = 0.f;
        for (unsigned int k = 0; k < sub_len; k++) {
            const float2 s = xs[t + k];
            const float c = hf[k];
            re = fmaf(c, s.x, re);
            im = fmaf(c, s.y, im);
        }
        y[i0 + t] = make_float2(re * scale, im * scale);
    }
}

// Four consecutive outputs per thread: output 4 t + r meets sample 4 t + r + k at tap k, so a four-sample shift register
// takes ONE new sample per tap step for four outputs (the kernel above reads sub_len samples per output and is bound by
// shared-memory bandwidth: 12 LDS.64 + 12 LDS.32 per output at sub_len = 12).  Samples are staged transposed,
// slot(q) = (q % 4) ROW + q / 4 with ROW = 4 (mod 16): the 8-byte cp.async fill of 16 consecutive q and the threads'
// reads (consecutive t -> consecutive slots) are both conflict-free.  Same fmaf chain per output (oldest sample first):
// identical bits (tested against the kernels above).
__global__ void __launch_bounds__(kThreads)
firpfb_tile4_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H, const float *__restrict__ sub,
                    unsigned int sub_len, unsigned int n_filters, const unsigned int *__restrict__ idx, float scale,
                    float2 *__restrict__ y, unsigned long long n, unsigned int ROW)
{
    extern __shared__ float2 smem_p[];
    const unsigned int halo = sub_len - 1;
    const unsigned int row = sub_len | 1;                                // padded row length of the bank
    float2 *xs = smem_p;                                                 // staged sample q = sample i0 - halo + q, transposed
    float *bank = reinterpret_cast<float *>(xs + 4 * ROW);
    const unsigned long long i0 = (unsigned long long)blockIdx.x * kPfbTile;
    const unsigned int nt = (n - i0 < kPfbTile) ? (unsigned int)(n - i0) : kPfbTile;
    for (unsigned int e = threadIdx.x; e < n_filters * sub_len; e += blockDim.x)
        bank[(e / sub_len) * row + e % sub_len] = __ldg(sub + e);
    for (unsigned int q = threadIdx.x; q < kPfbTile + halo; q += blockDim.x) {
        const long long s = (long long)i0 - (long long)halo + q;
        float2 *dstp = xs + (q & 3) * ROW + (q >> 2);
        if (s >= 0 && s < (long long)n) {
            const unsigned int dst = (unsigned int)__cvta_generic_to_shared(dstp);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(x + s) : "memory");
        } else {
            *dstp = (s < 0) ? __ldg(hist + (H + s)) : make_float2(0.f, 0.f);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (unsigned int t = threadIdx.x; 4 * t < nt; t += blockDim.x) {
        // filter rows of the four outputs (outputs beyond the end of the block reuse row 0 and are not stored)
        const float *hf[4];
#pragma unroll
        for (int r = 0; r < 4; r++) hf[r] = bank + ((4 * t + r < nt) ? __ldg(idx + i0 + 4 * t + r) : 0u) * row;
        float2 w[4];
#pragma unroll
        for (int r = 0; r < 3; r++) w[r + 1] = xs[(r & 3) * ROW + t];    // samples 4 t + 0 .. 2 (slots of q = 4 t + r)
        float re[4] = {0.f, 0.f, 0.f, 0.f}, im[4] = {0.f, 0.f, 0.f, 0.f};
        for (unsigned int k = 0; k < sub_len; k++) {
            // window after the shift: w[r] = sample 4 t + r + k
            w[0] = w[1]; w[1] = w[2]; w[2] = w[3];
            const unsigned int q = 4 * t + 3 + k;
            w[3] = xs[(q & 3) * ROW + (q >> 2)];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const float c = hf[r][k];
                re[r] = fmaf(c, w[r].x, re[r]);
                im[r] = fmaf(c, w[r].y, im[r]);
            }
        }
        float2 *yo = y + i0 + 4 * t;
        if (4 * t + 3 < nt && ((uintptr_t)yo % 16) == 0) {
            __stcs(reinterpret_cast<float4 *>(yo), make_float4(re[0] * scale, im[0] * scale, re[1] * scale, im[1] * scale));
            __stcs(reinterpret_cast<float4 *>(yo) + 1, make_float4(re[2] * scale, im[2] * scale, re[3] * scale, im[3] * scale));
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (4 * t + r < nt) yo[r] = make_float2(re[r] * scale, im[r] * scale);
        }
    }
}

void launch_firpfb_block(const float2 *d_hist, const float2 *d_x, int H, const float *d_sub, unsigned int sub_len,
                         const unsigned int *d_idx, float scale, float2 *d_y, unsigned long long n, cudaStream_t st,
                         unsigned int n_filters)
{
    const size_t smem = sizeof(float2) * ((kPfbTile + sub_len) & ~1u) + sizeof(float) * (size_t)n_filters * (sub_len | 1);
    const unsigned long long tiles = (n + kPfbTile - 1) / kPfbTile;
    // ROW = 4 (mod 16) and >= ceil((tile + halo) / 4)
    const unsigned int need = (kPfbTile + sub_len - 1 + 3) / 4;
    const unsigned int ROW = ((need + 11) / 16) * 16 + 4;
    const size_t smem4 = sizeof(float2) * 4 * ROW + sizeof(float) * (size_t)n_filters * (sub_len | 1);
    if (n_filters && smem4 <= 48 * 1024 && n >= 4ull * kPfbTile && tiles < (1ull << 31) && !getenv("LIQUID_CUDA_FIRPFB_SIMPLE") &&
        !getenv("LIQUID_CUDA_FIRPFB_TILE1")) {
        firpfb_tile4_kernel<<<(unsigned int)tiles, kThreads, smem4, st>>>(d_hist, d_x, H, d_sub, sub_len, n_filters, d_idx, scale, d_y, n, ROW);
    } else if (n_filters && smem <= 48 * 1024 && n >= 4ull * kPfbTile && tiles < (1ull << 31) && !getenv("LIQUID_CUDA_FIRPFB_SIMPLE")) {
        firpfb_tile_kernel<<<(unsigned int)tiles, kThreads, smem, st>>>(d_hist, d_x, H, d_sub, sub_len, n_filters, d_idx, scale, d_y, n);
    } else {
        firpfb_block_kernel<<<grid_cap((n + kThreads - 1) / kThreads), kThreads, 0, st>>>(d_hist, d_x, H, d_sub, sub_len, d_idx, scale, d_y, n);
    }
    LC_LAUNCH_CHECK();
}

}  // namespace firfilt
}  // namespace lc
