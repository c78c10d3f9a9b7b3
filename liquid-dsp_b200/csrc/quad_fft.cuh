// quad_fft.cuh -- batched N-point transforms, N = 256 RA with RA = 4 / 8 / 16 (N = 1024 / 2048 / 4096), in the form of
// pfbch_quad_1024_kernel (pfbch_fused.cu): the same source / destination functors as tile_fft.cuh, for the callers that
// are a transform with something on either side (the fft_* batch API, ofdmframegen at long symbols).
//   one instance = one CTA of 256 threads walking tiles of 16 / RA transforms (a tile is always 16 sub-problems).
//   phase 1  thread u owns the RA inputs u + 256 q of a transform (each load instruction of a warp is one 256-byte run),
//            prefetched into registers a tile ahead: an RA-point register DFT over q (the first decimation-in-frequency
//            stage) and the RA - 1 twiddles W_N^(u p) leave the transform as RA independent 256-point problems z_p[u]
//            in the shared tile.
//   phase 2  16 RA threads per transform, lane = (p fastest, column): 256 = 16 x 16, a 16-point register DFT down the
//            column (in place), the inter-pass twiddle from a shared copy, a barrier over the transform's threads, a
//            16-point DFT along the row, and out[RA (k1 + 16 k2) + p] leaves as runs of 16 RA consecutive outputs.
//   Sub-problems are 272 + 16 / RA units apart (16 rows of 17, then a pad that spreads the RA sub-problems a half-warp
//   touches at once over the banks), so every shared-memory access pattern is conflict-free.  Two tile buffers leave one
//   block barrier per tile.
#pragma once
#include <cmath>
#include <vector>

#include "fused_common.cuh"

namespace lcquad {

using namespace lcf;

constexpr int kThreads = 256;
constexpr int kTwPitch = 18;

inline bool supported(unsigned int n) { return n == 1024 || n == 2048 || n == 4096; }

template <int RA> struct Pitch {
    static constexpr int SP = 272 + 16 / RA;
    static constexpr size_t kBytes = sizeof(float2) * ((size_t)2 * 16 * SP + 16 * 18);      // = smem_bytes<RA>(), usable in device code
};
template <int RA> constexpr size_t smem_bytes() { return sizeof(float2) * ((size_t)2 * 16 * Pitch<RA>::SP + 16 * kTwPitch); }

// host: tw1[p-1][u] = e^{sign j 2 pi u p / N}, p = 1 ... RA-1, u < 256; then tw256[k1][r2] = e^{sign j 2 pi k1 r2 / 256}
inline void make_twiddles(std::vector<float2> &tw, unsigned int n, int sign)
{
    const unsigned int RA = n / 256;
    tw.resize((size_t)(RA - 1) * 256 + 256);
    for (unsigned int p = 1; p < RA; p++)
        for (unsigned int u = 0; u < 256; u++) {
            const double ang = (double)sign * 2.0 * M_PI * (double)(u * p) / (double)n;
            tw[(size_t)(p - 1) * 256 + u] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (unsigned int k1 = 0; k1 < 16; k1++)
        for (unsigned int r2 = 0; r2 < 16; r2++) {
            const double ang = (double)sign * 2.0 * M_PI * (double)(k1 * r2) / 256.0;
            tw[(size_t)(RA - 1) * 256 + k1 * 16 + r2] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

// grid for n transforms of 256 RA points: instances of `tiles_per_inst` consecutive tiles
inline void plan(unsigned int RA, unsigned long long n, unsigned int &grid, int &tiles_per_inst)
{
    const long long fr = 16 / RA;
    const long long n_tiles = (long long)((n + fr - 1) / fr);
    long long tpi = 16;
    while (tpi > 1 && (n_tiles / tpi) < 148ll * 8) tpi >>= 1;
    tiles_per_inst = (int)tpi;
    grid = (unsigned int)((n_tiles + tpi - 1) / tpi);
}

#ifdef __CUDACC__
// the whole kernel body: call from a __global__ function launched with kThreads threads and smem_bytes<RA>() of dynamic
// shared memory (passed in as smem_raw)
// TWLD: the phase-1 twiddles are re-read (L1) in every tile instead of living in 2 (RA - 1) registers
// Post: post(tr0, n_valid) runs in every thread after the tile's outputs have gone to dst (which may be per-thread or
// shared-memory state of the caller's; a block barrier separates post of one tile from dst of the next)
struct NoPost {
    __device__ __forceinline__ void operator()(long long, int) const {}
};
template <int RA, int SIGN, bool TWLD, class Src, class Dst, class Post = NoPost>
__device__ __forceinline__ void run(unsigned char *smem_raw, const float2 *__restrict__ tw, long long n_tr, int tiles_per_inst,
                                    Src src, Dst &&dst, Post &&post = NoPost())
{
    constexpr int FR = 16 / RA, SP = Pitch<RA>::SP, GRP = 16 * RA;
    float2 *tile0 = reinterpret_cast<float2 *>(smem_raw);
    float2 *tws = tile0 + 2 * 16 * SP;                                // [r2][k1], pitch 18
    const int u = threadIdx.x;
    const long long n_tiles = (n_tr + FR - 1) / FR;
    const long long t0 = (long long)blockIdx.x * tiles_per_inst;
    long long ntl = n_tiles - t0;
    if (ntl > tiles_per_inst) ntl = tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;

    float2 nx[16];
    auto fetch = [&](int it) {
        const long long tr0 = (t0 + it) * FR;
#pragma unroll
        for (int f = 0; f < FR; f++) {
            const bool ok = it < nt && tr0 + f < n_tr;
#pragma unroll
            for (int q = 0; q < RA; q++) nx[f * RA + q] = ok ? src.load(tr0 + f, (unsigned int)(u + 256 * q)) : make_float2(0.f, 0.f);
        }
    };
    fetch(0);
    tws[(u & 15) * kTwPitch + (u >> 4)] = __ldg(tw + (RA - 1) * 256 + u);   // tw256[k1 = u >> 4][r2 = u & 15]
    float2 t1[RA - 1];
    if (!TWLD) {
#pragma unroll
        for (int p = 1; p < RA; p++) t1[p - 1] = __ldg(tw + (p - 1) * 256 + u);
    }
    // phase 2 role: transform f2 of the tile, sub-problem pp (fastest), column / row c
    const int f2 = u / GRP, pp = u % RA, c = (u % GRP) / RA;
    const float4 *twc = reinterpret_cast<const float4 *>(tws + c * kTwPitch);
    const int z_off = (u >> 4) * 17 + (u & 15);                       // where z_p[u] lives inside its sub-problem
    __syncthreads();

    for (int it = 0; it < nt; it++) {
        const long long tr0 = (t0 + it) * FR;
        float2 *tile = tile0 + (it & 1) * (16 * SP);
        // ---- phase 1
#pragma unroll
        for (int f = 0; f < FR; f++) {
            float2 d[RA];
#pragma unroll
            for (int q = 0; q < RA; q++) d[q] = src.finish(tr0 + f, (unsigned int)(u + 256 * q), nx[f * RA + q]);
            RegDft<RA, SIGN>::run(d);
            float2 *zf = tile + f * RA * SP + z_off;
            zf[0] = d[0];
#pragma unroll
            for (int p = 1; p < RA; p++) zf[p * SP] = cmul2_rot(d[p], TWLD ? __ldg(tw + (p - 1) * 256 + u) : t1[p - 1]);
        }
        fetch(it + 1);
        __syncthreads();
        // ---- phase 2
        float2 *sp = tile + (f2 * RA + pp) * SP;
        float2 v[16];
#pragma unroll
        for (int r1 = 0; r1 < 16; r1++) v[r1] = sp[r1 * 17 + c];
        RegDft<16, SIGN>::run(v);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 t = twc[k];
            if (k > 0) v[2 * k] = cmul2_rot(v[2 * k], make_float2(t.x, t.y));
            v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
        }
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) sp[k1 * 17 + c] = v[k1];
        if (GRP == kThreads) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + f2), "n"(GRP) : "memory");
#pragma unroll
        for (int cc = 0; cc < 16; cc++) v[cc] = sp[c * 17 + cc];      // this thread now plays row k1 = c
        RegDft<16, SIGN>::run(v);
        if (tr0 + f2 < n_tr) {
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) dst(tr0 + f2, (unsigned int)(RA * c + pp + GRP * k2), v[k2]);
        }
        post(tr0, (int)(n_tr - tr0 < FR ? n_tr - tr0 : FR));
    }
}

// Forward transform, product with a frequency response, backward transform, in one kernel and without an exchange in
// between: after phase 2 of the forward transform a thread holds X[RA c + pp + 16 RA k2], k2 < 16 -- with k2 = g + (16 / RA) q
// these are, for each g, exactly the RA inputs u' + 256 q (u' = RA c + pp + 16 RA g) that ONE thread of the backward
// transform's phase 1 needs.  So the thread multiplies its 16 values by H, does the 16 / RA first-stage DFTs of the backward
// transform in registers and writes the backward sub-problems into the second tile buffer.  Two block barriers per tile.
// tw_f / tw_b: the forward / backward tables of make_twiddles; Hf[N] in natural order.  UPPER: only outputs N/2 ... N-1
// are handed to dst (overlap-save keeps the second half).
template <int RA, bool UPPER, class Src, class Dst>
__device__ __forceinline__ void run_chain(unsigned char *smem_raw, const float2 *__restrict__ tw_f, const float2 *__restrict__ tw_b,
                                          const float2 *__restrict__ Hf, long long n_tr, int tiles_per_inst, Src src, Dst &&dst)
{
    constexpr int FR = 16 / RA, NG = 16 / RA, SP = Pitch<RA>::SP, GRP = 16 * RA;
    float2 *bufA = reinterpret_cast<float2 *>(smem_raw);
    float2 *bufB = bufA + 16 * SP;
    float2 *twsf = bufB + 16 * SP;                                    // [r2][k1], pitch 18: forward, then backward
    float2 *twsb = twsf + 16 * kTwPitch;
    const int u = threadIdx.x;
    const long long n_tiles = (n_tr + FR - 1) / FR;
    const long long t0 = (long long)blockIdx.x * tiles_per_inst;
    long long ntl = n_tiles - t0;
    if (ntl > tiles_per_inst) ntl = tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;

    float2 nx[16];
    auto fetch = [&](int it) {
        const long long tr0 = (t0 + it) * FR;
#pragma unroll
        for (int f = 0; f < FR; f++) {
            const bool ok = it < nt && tr0 + f < n_tr;
#pragma unroll
            for (int q = 0; q < RA; q++) nx[f * RA + q] = ok ? src.load(tr0 + f, (unsigned int)(u + 256 * q)) : make_float2(0.f, 0.f);
        }
    };
    fetch(0);
    twsf[(u & 15) * kTwPitch + (u >> 4)] = __ldg(tw_f + (RA - 1) * 256 + u);
    twsb[(u & 15) * kTwPitch + (u >> 4)] = __ldg(tw_b + (RA - 1) * 256 + u);
    const int f2 = u / GRP, pp = u % RA, c = (u % GRP) / RA;
    const float4 *twcf = reinterpret_cast<const float4 *>(twsf + c * kTwPitch);
    const float4 *twcb = reinterpret_cast<const float4 *>(twsb + c * kTwPitch);
    const int z_off = (u >> 4) * 17 + (u & 15);
    __syncthreads();

    for (int it = 0; it < nt; it++) {
        const long long tr0 = (t0 + it) * FR;
        // ---- forward phase 1 -> buffer A
#pragma unroll
        for (int f = 0; f < FR; f++) {
            float2 d[RA];
#pragma unroll
            for (int q = 0; q < RA; q++) d[q] = src.finish(tr0 + f, (unsigned int)(u + 256 * q), nx[f * RA + q]);
            RegDft<RA, -1>::run(d);
            float2 *zf = bufA + f * RA * SP + z_off;
            zf[0] = d[0];
#pragma unroll
            for (int p = 1; p < RA; p++) zf[p * SP] = cmul2_rot(d[p], __ldg(tw_f + (p - 1) * 256 + u));
        }
        fetch(it + 1);
        __syncthreads();
        float2 v[16];
        // ---- forward phase 2
        {
            float2 *sp = bufA + (f2 * RA + pp) * SP;
#pragma unroll
            for (int r1 = 0; r1 < 16; r1++) v[r1] = sp[r1 * 17 + c];
            RegDft<16, -1>::run(v);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 t = twcf[k];
                if (k > 0) v[2 * k] = cmul2_rot(v[2 * k], make_float2(t.x, t.y));
                v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
            }
#pragma unroll
            for (int k1 = 0; k1 < 16; k1++) sp[k1 * 17 + c] = v[k1];
            if (GRP == kThreads) __syncthreads();
            else asm volatile("bar.sync %0, %1;" ::"r"(1 + f2), "n"(GRP) : "memory");
#pragma unroll
            for (int cc = 0; cc < 16; cc++) v[cc] = sp[c * 17 + cc];
            RegDft<16, -1>::run(v);
        }
        // ---- product with H and backward phase 1 -> buffer B
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) v[k2] = cmul2_rot(v[k2], __ldg(Hf + RA * c + pp + GRP * k2));
#pragma unroll
        for (int g = 0; g < NG; g++) {
            float2 d[RA];
#pragma unroll
            for (int q = 0; q < RA; q++) d[q] = v[g + NG * q];
            RegDft<RA, +1>::run(d);
            const int up = RA * c + pp + GRP * g;
            float2 *zb = bufB + f2 * RA * SP + (up >> 4) * 17 + (up & 15);
            zb[0] = d[0];
#pragma unroll
            for (int p = 1; p < RA; p++) zb[p * SP] = cmul2_rot(d[p], __ldg(tw_b + (p - 1) * 256 + up));
        }
        __syncthreads();
        // ---- backward phase 2
        {
            float2 *sp = bufB + (f2 * RA + pp) * SP;
#pragma unroll
            for (int r1 = 0; r1 < 16; r1++) v[r1] = sp[r1 * 17 + c];
            RegDft<16, +1>::run(v);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 t = twcb[k];
                if (k > 0) v[2 * k] = cmul2_rot(v[2 * k], make_float2(t.x, t.y));
                v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
            }
#pragma unroll
            for (int k1 = 0; k1 < 16; k1++) sp[k1 * 17 + c] = v[k1];
            if (GRP == kThreads) __syncthreads();
            else asm volatile("bar.sync %0, %1;" ::"r"(1 + f2), "n"(GRP) : "memory");
#pragma unroll
            for (int cc = 0; cc < 16; cc++) v[cc] = sp[c * 17 + cc];
            RegDft<16, +1>::run(v);
        }
        if (tr0 + f2 < n_tr) {
#pragma unroll
            for (int k2 = UPPER ? 8 : 0; k2 < 16; k2++) dst(tr0 + f2, (unsigned int)(RA * c + pp + GRP * k2), v[k2]);
        }
    }
}
template <int RA> constexpr size_t chain_smem_bytes() { return sizeof(float2) * ((size_t)2 * 16 * Pitch<RA>::SP + 2 * 16 * kTwPitch); }
#endif

}  // namespace lcquad
