// fft.cu -- batched complex FFT on the device + the fft_* C API.
//
// Stands in for src/fft/src/fft_common.c:109-316 (plan / execute / run / shift) and the
// transforms behind it (fft_mixed_radix.c:128-191, fft_dft.c).  Conventions kept:
// FORWARD = e^{-j 2 pi nk/N}, BACKWARD = e^{+j 2 pi nk/N}, both unscaled (fft_radix2.c:68-70).
// The factorisation is NOT the reference's (any float32 FFT agrees with it to ~1e-7 relative):
//   power-of-two N : Stockham autosort, radix-4 stages (+ one radix-2 stage when log2 N is odd),
//                    whole transform in shared memory, several transforms per CTA when N is small
//   other composite N with prime factors <= 13 (N <= 12288): mixed-radix Stockham autosort in shared memory
//                    (stands in for fft_mixed_radix.c:128-191), factors 4, 2, 3, 5, 7, 11, 13, radix-R register DFTs
//   any other N    : table-driven O(N^2) DFT, one CTA per transform (slow, correct, on-device)
// Twiddles come from a per-plan table computed in double on the host.
#include <cmath>

#include "lc_common.cuh"
#include "fft_device.cuh"
#include "fused_common.cuh"
#include "tile_fft.cuh"
#include "quad_fft.cuh"
#include "../../include/liquid_cuda.h"

namespace {

using lcfft::fft_post;
using lcfft::kFftThreads;

// first-pass source / last-pass destination of the batched FFT: global memory, directly (consecutive threads
// read consecutive addresses; the last pass writes runs of Ns >= 16 consecutive outputs, post-scaled)
struct GlobalSrc {
    const float2 *x;
    unsigned int n;
    unsigned long long first, batch;
    __device__ __forceinline__ float2 operator()(unsigned int f, unsigned int idx) const
    {
        const unsigned long long b = first + f;
        return b < batch ? __ldg(x + b * n + idx) : make_float2(0.f, 0.f);
    }
};
struct GlobalDst {
    float2 *y;
    unsigned int n;
    unsigned long long first, batch;
    int post_mode;
    float pa, pb;
    __device__ __forceinline__ void operator()(unsigned int f, unsigned int idx, float2 v) const
    {
        const unsigned long long b = first + f;
        if (b < batch) y[b * n + idx] = fft_post(v, post_mode, pa, pb);
    }
};

// Stockham autosort FFT of n = 2^log2n points per transform, `per_cta` transforms per CTA (fft_device.cuh),
// first pass from global memory, last pass to global memory: a transform makes no extra trip through shared
// memory at either end, and transforms of up to 16 points never touch it.  In place (x == y) is fine: a CTA
// has read all the inputs of its transforms before it writes any of their outputs.
template <bool TWG>
__global__ void __launch_bounds__(kFftThreads)
fft_pow2_kernel(const float2 *__restrict__ x, float2 *__restrict__ y, const float2 *__restrict__ tw,
                unsigned int n, unsigned int log2n, unsigned int per_cta, unsigned long long batch,
                float sgn, int post_mode, float pa, float pb)
{
    extern __shared__ float2 smem[];
    lcfft::Smem sm = lcfft::carve(smem, n, per_cta);
    const unsigned long long first = (unsigned long long)blockIdx.x * per_cta;
    if (TWG) sm.tw = const_cast<float2 *>(tw);                // long transforms: the plan's table itself, through L1
    else if (log2n > 4) lcfft::stage_twiddles(sm, tw, n);     // published by the barrier after the first pass
    const GlobalSrc src{x, n, first, batch};
    const GlobalDst dst{y, n, first, batch, post_mode, pa, pb};
    if (sgn > 0.0f) lcfft::stockham_run<+1, TWG>(sm, n, log2n, per_cta, src, dst);
    else            lcfft::stockham_run<-1, TWG>(sm, n, log2n, per_cta, src, dst);
}

// n = 64 ... 512: the tile transform of tile_fft.cuh (register DFTs, one instance barrier per 8 transforms) between global
// memory and global memory; in place is fine (a tile's inputs are all read before any of its outputs is written)
struct TileSrc {
    const float2 *x;
    unsigned int n;
    __device__ __forceinline__ float2 load(long long t, unsigned int k) const { return __ldg(x + t * n + k); }
    __device__ __forceinline__ float2 finish(long long, unsigned int, float2 raw) const { return raw; }
};
struct TileDst {
    float2 *y;
    unsigned int n;
    int post_mode;
    float pa, pb;
    __device__ __forceinline__ void operator()(long long t, unsigned int k, float2 v) const
    {
        __stcs(y + t * n + k, fft_post(v, post_mode, pa, pb));
    }
};

template <int M, int SIGN>
__global__ void __launch_bounds__(lctile::kThreads, 2)
fft_tile_kernel(const float2 *__restrict__ x, float2 *__restrict__ y, const float2 *__restrict__ tw, long long batch,
                int tiles_per_inst, int post_mode, float pa, float pb)
{
    extern __shared__ __align__(16) unsigned char smem_tile[];
    lctile::run<M, SIGN>(smem_tile, tw, batch, tiles_per_inst, TileSrc{x, (unsigned int)M}, TileDst{y, (unsigned int)M, post_mode, pa, pb});
}

template <int M, int SIGN>
void launch_fft_tile(const float2 *d_x, float2 *d_y, const float2 *d_tw, unsigned long long batch, cudaStream_t st, int post_mode,
                     float a, float b)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(fft_tile_kernel<M, SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lctile::smem_bytes<M>()));
    unsigned int grid = 0;
    int tpi = 1;
    lctile::plan(M, batch, grid, tpi);
    fft_tile_kernel<M, SIGN><<<grid, lctile::kThreads, lctile::smem_bytes<M>(), st>>>(d_x, d_y, d_tw, (long long)batch, tpi, post_mode, a, b);
}

template <int M>
void launch_fft_tile_M(int sign, const float2 *d_x, float2 *d_y, const float2 *d_tw, unsigned long long batch, cudaStream_t st,
                       int post_mode, float a, float b)
{
    if (sign > 0) launch_fft_tile<M, +1>(d_x, d_y, d_tw, batch, st, post_mode, a, b);
    else          launch_fft_tile<M, -1>(d_x, d_y, d_tw, batch, st, post_mode, a, b);
}

// n = 4 / 8 / 16 / 32: one thread per transform, a warp per 32 consecutive transforms.  Read by the thread itself, a transform is
// 8 n bytes at a stride of 8 n: every load instruction of the warp touched 32 lines (n = 16: 0.31 of the roofline).  Here
// the warp moves its 32 n points with 128-bit accesses in stream order and the threads pick their rows out of a padded
// per-warp staging tile (row pitch n + 1 units: the row reads are conflict-free); same path back.  No block barrier.
template <int N> struct SmallCfg { static constexpr int WPC = N == 32 ? 4 : 8; };      // warps per CTA (static shared memory < 48 KB)
template <int N, int SIGN>
__global__ void __launch_bounds__(32 * SmallCfg<N>::WPC)
fft_small_kernel(const float2 *__restrict__ x, float2 *__restrict__ y, unsigned long long batch, int post_mode, float pa, float pb)
{
    constexpr int PITCH = N + 1, PER_LANE = N / 2;                    // float4 (two points) per lane per chunk
    constexpr int WPC = SmallCfg<N>::WPC;
    __shared__ float2 stage[WPC][32 * PITCH];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    float2 *st = stage[wl];
    const unsigned long long n_chunks = (batch + 31) / 32, total = batch * N;
    const unsigned long long w0 = (unsigned long long)blockIdx.x * WPC + wl, wstep = (unsigned long long)gridDim.x * WPC;
    for (unsigned long long ch = w0; ch < n_chunks; ch += wstep) {
        const unsigned long long base = ch * 32 * N;                   // first point of the chunk
        const float4 *x4 = reinterpret_cast<const float4 *>(x + base);
        float4 in[PER_LANE];
#pragma unroll
        for (int j = 0; j < PER_LANE; j++) {
            const int i = j * 32 + lane;                               // float4 index inside the chunk
            in[j] = base + 2ull * i < total ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < PER_LANE; j++) {
            const int pnt = 2 * (j * 32 + lane);
            float2 *d = st + (pnt / N) * PITCH + (pnt % N);
            d[0] = make_float2(in[j].x, in[j].y);
            d[1] = make_float2(in[j].z, in[j].w);
        }
        __syncwarp();
        float2 v[N];
#pragma unroll
        for (int k = 0; k < N; k++) v[k] = st[lane * PITCH + k];
        lcf::RegDft<N, SIGN>::run(v);
#pragma unroll
        for (int k = 0; k < N; k++) st[lane * PITCH + k] = fft_post(v[k], post_mode, pa, pb);
        __syncwarp();
        float4 *y4 = reinterpret_cast<float4 *>(y + base);
#pragma unroll
        for (int j = 0; j < PER_LANE; j++) {
            const int i = j * 32 + lane, pnt = 2 * i;
            const float2 *sp = st + (pnt / N) * PITCH + (pnt % N);
            const float2 a = sp[0], b = sp[1];
            if (base + 2ull * i < total) __stcs(y4 + i, make_float4(a.x, a.y, b.x, b.y));
        }
        __syncwarp();
    }
}

template <int N>
void launch_fft_small(int sign, const float2 *d_x, float2 *d_y, unsigned long long batch, cudaStream_t st, int post_mode, float a, float b)
{
    const unsigned long long n_chunks = (batch + 31) / 32;
    constexpr int WPC = SmallCfg<N>::WPC;
    unsigned long long ctas = (n_chunks + WPC - 1) / WPC;
    if (ctas > 148ull * 8 * 4) ctas = 148ull * 8 * 4;
    if (sign > 0) fft_small_kernel<N, +1><<<(unsigned int)ctas, 32 * WPC, 0, st>>>(d_x, d_y, batch, post_mode, a, b);
    else          fft_small_kernel<N, -1><<<(unsigned int)ctas, 32 * WPC, 0, st>>>(d_x, d_y, batch, post_mode, a, b);
}

// n = 1024 / 2048 / 4096: the quad transform of quad_fft.cuh (first radix-4 / 8 / 16 stage in the loading thread, then
// sixteen 256-point problems per tile); in place is fine for the same reason
template <int RA, int SIGN>
__global__ void __launch_bounds__(lcquad::kThreads, 2)
fft_quad_kernel(const float2 *__restrict__ x, float2 *__restrict__ y, const float2 *__restrict__ tw, long long batch,
                int tiles_per_inst, int post_mode, float pa, float pb)
{
    extern __shared__ __align__(16) unsigned char smem_quad[];
    lcquad::run<RA, SIGN, false>(smem_quad, tw, batch, tiles_per_inst, TileSrc{x, 256u * RA}, TileDst{y, 256u * RA, post_mode, pa, pb});
}

template <int RA, int SIGN>
void launch_fft_quad(const float2 *d_x, float2 *d_y, const float2 *d_tw, unsigned long long batch, cudaStream_t st, int post_mode,
                     float a, float b)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(fft_quad_kernel<RA, SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lcquad::smem_bytes<RA>()));
    unsigned int grid = 0;
    int tpi = 1;
    lcquad::plan(RA, batch, grid, tpi);
    fft_quad_kernel<RA, SIGN><<<grid, lcquad::kThreads, lcquad::smem_bytes<RA>(), st>>>(d_x, d_y, d_tw, (long long)batch, tpi, post_mode, a, b);
}

template <int RA>
void launch_fft_quad_RA(int sign, const float2 *d_x, float2 *d_y, const float2 *d_tw, unsigned long long batch, cudaStream_t st,
                        int post_mode, float a, float b)
{
    if (sign > 0) launch_fft_quad<RA, +1>(d_x, d_y, d_tw, batch, st, post_mode, a, b);
    else          launch_fft_quad<RA, -1>(d_x, d_y, d_tw, batch, st, post_mode, a, b);
}

// table-driven DFT, one CTA per transform; smem: n float2
__global__ void __launch_bounds__(kFftThreads)
dft_any_kernel(const float2 *__restrict__ x, float2 *__restrict__ y, const float2 *__restrict__ tw,
               unsigned int n, int post_mode, float pa, float pb)
{
    extern __shared__ float2 smem[];
    const unsigned long long b = blockIdx.x;
    for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) smem[i] = x[b * n + i];
    __syncthreads();
    for (unsigned int k = threadIdx.x; k < n; k += blockDim.x) {
        float2 acc = smem[0];
        unsigned int idx = 0;
        for (unsigned int i = 1; i < n; i++) {
            idx += k;
            if (idx >= n) idx -= n;
            acc = lcd::cadd(acc, lcd::cmul(smem[i], tw[idx]));
        }
        y[b * n + k] = fft_post(acc, post_mode, pa, pb);
    }
}

// ---- mixed-radix Stockham autosort for composite n
// One pass of radix R over a transform of n points: butterfly j (j < n / R) with k = j mod Ns (Ns = the product of the
// radices already done) takes in[j + r n/R], r < R, multiplies by e^{sgn j 2 pi r k / (Ns R)} (the plan's table, step
// n / (Ns R)), does the R-point DFT (roots from the same table, step n / R) and writes out[(j / Ns) Ns R + k + q Ns].
template <int R>
__device__ __forceinline__ void mixed_pass(const float2 *__restrict__ in, float2 *__restrict__ out, const float2 *__restrict__ tw,
                                           unsigned int n, unsigned int Ns, unsigned int per_cta, bool to_global,
                                           float2 *__restrict__ y, unsigned long long first, unsigned long long batch,
                                           int post_mode, float pa, float pb)
{
    const unsigned int nb = n / R, tstep = n / (Ns * R), rstep = n / R;
    for (unsigned int w = threadIdx.x; w < per_cta * nb; w += blockDim.x) {
        const unsigned int f = w / nb, j = w - f * nb;
        const unsigned int k = j % Ns;
        const float2 *src = in + f * n;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            v[r] = src[j + r * nb];
            if (r > 0) v[r] = lcd::cmul(v[r], __ldg(tw + (unsigned int)r * k * tstep));     // r k tstep < n
        }
        float2 o[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            float2 acc = v[0];
#pragma unroll
            for (int r = 1; r < R; r++) acc = lcd::cadd(acc, lcd::cmul(v[r], __ldg(tw + ((q * r) % R) * rstep)));
            o[q] = acc;
        }
        const unsigned int j0 = (j / Ns) * Ns * R + k;
        if (to_global) {
            const unsigned long long b = first + f;
            if (b < batch) {
#pragma unroll
                for (int q = 0; q < R; q++) y[b * n + j0 + q * Ns] = fft_post(o[q], post_mode, pa, pb);
            }
        } else {
            float2 *dst = out + f * n;
#pragma unroll
            for (int q = 0; q < R; q++) dst[j0 + q * Ns] = o[q];
        }
    }
}

struct MixedRadices {
    unsigned char r[16];
    unsigned int count;
};

// per_cta transforms per CTA, two ping-pong buffers of per_cta * n points; in place (x == y) is fine: a CTA has read all the
// inputs of its transforms before it writes any of their outputs
__global__ void __launch_bounds__(kFftThreads)
fft_mixed_kernel(const float2 *__restrict__ x, float2 *__restrict__ y, const float2 *__restrict__ tw, unsigned int n,
                 MixedRadices rad, unsigned int per_cta, unsigned long long batch, int post_mode, float pa, float pb)
{
    extern __shared__ float2 smem[];
    float2 *buf0 = smem, *buf1 = smem + (size_t)per_cta * n;
    const unsigned long long first = (unsigned long long)blockIdx.x * per_cta;
    for (unsigned int w = threadIdx.x; w < per_cta * n; w += blockDim.x) {
        const unsigned long long b = first + w / n;
        buf0[w] = b < batch ? __ldg(x + first * n + w) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    unsigned int Ns = 1;
    for (unsigned int s = 0; s < rad.count; s++) {
        const bool last = (s + 1 == rad.count);
        const unsigned int R = rad.r[s];
#define LC_PASS(RR) case RR: mixed_pass<RR>(buf0, buf1, tw, n, Ns, per_cta, last, y, first, batch, post_mode, pa, pb); break;
        switch (R) {
            LC_PASS(2) LC_PASS(3) LC_PASS(4) LC_PASS(5) LC_PASS(7) LC_PASS(11) LC_PASS(13)
        default: break;
        }
#undef LC_PASS
        Ns *= R;
        __syncthreads();
        float2 *t = buf0;
        buf0 = buf1;
        buf1 = t;
    }
}

}  // namespace

namespace lc {

void FftPlanDev::init(unsigned int n_, int sign_)
{
    ensure_device();
    n = n_;
    sign = sign_;
    std::vector<float2> tw(n);
    for (unsigned int i = 0; i < n; i++) {
        double a = (double)sign * 2.0 * M_PI * (double)i / (double)n;
        tw[i] = make_float2((float)cos(a), (float)sin(a));
    }
    LC_CUDA_CHECK(cudaMalloc(&d_tw, sizeof(float2) * n));
    LC_CUDA_CHECK(cudaMemcpy(d_tw, tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    // transforms above 3072 (power of two) / 6144 (other) points need more than the default 48 KB of dynamic
    // shared memory; every kernel a plan may launch is opted in, whatever n this particular plan has
    static std::atomic<unsigned long long> configured{0}, configured_mixed{0};
    if (first_time_on_device(configured)) {
        LC_CUDA_CHECK(cudaFuncSetAttribute(fft_pow2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        LC_CUDA_CHECK(cudaFuncSetAttribute(fft_pow2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        LC_CUDA_CHECK(cudaFuncSetAttribute(dft_any_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    // composite sizes: factors 4 first, then 2, 3, 5, 7, 11, 13; anything left over means no mixed-radix plan
    n_rad = 0;
    if (!is_pow2(n) && n >= 6 && (size_t)n * 2 * sizeof(float2) <= 192 * 1024) {
        unsigned int rest = n, cnt = 0;
        unsigned char fac[16];
        const unsigned int cand[7] = {4, 2, 3, 5, 7, 11, 13};
        for (unsigned int c : cand)
            while (rest % c == 0 && cnt < 16) {
                fac[cnt++] = (unsigned char)c;
                rest /= c;
            }
        if (rest == 1 && cnt >= 2) {
            n_rad = cnt;
            for (unsigned int i = 0; i < cnt; i++) radices[i] = fac[i];
        }
    }
    if (first_time_on_device(configured_mixed))
        LC_CUDA_CHECK(cudaFuncSetAttribute(fft_mixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (lctile::supported(n) || lcquad::supported(n)) {
        std::vector<float2> twt;
        if (lctile::supported(n)) lctile::make_twiddles(twt, n, sign);
        else lcquad::make_twiddles(twt, n, sign);
        LC_CUDA_CHECK(cudaMalloc(&d_tw_tile, sizeof(float2) * twt.size()));
        LC_CUDA_CHECK(cudaMemcpy(d_tw_tile, twt.data(), sizeof(float2) * twt.size(), cudaMemcpyHostToDevice));
    }
    if (is_pow2(n) && n >= 2) {
        const std::vector<float2> twp = lcfft::pass_ordered_twiddles(tw, n);
        LC_CUDA_CHECK(cudaMalloc(&d_twp, sizeof(float2) * n));
        LC_CUDA_CHECK(cudaMemcpy(d_twp, twp.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    }
}

void FftPlanDev::destroy()
{
    if (d_tw) cudaFree(d_tw);
    if (d_twp) cudaFree(d_twp);
    if (d_tw_tile) cudaFree(d_tw_tile);
    d_tw_tile = nullptr;
    d_tw = nullptr;
    d_twp = nullptr;
}

void FftPlanDev::run(const float2 *d_x, float2 *d_y, unsigned long long batch, cudaStream_t st,
                     int post_mode, float a, float b) const
{
    if (batch == 0) return;
    if (n == 1) {
        // degenerate transform: copy (+ post) through the generic kernel
        dft_any_kernel<<<(unsigned int)batch, 32, sizeof(float2), st>>>(d_x, d_y, d_tw, 1, post_mode, a, b);
        LC_LAUNCH_CHECK();
        return;
    }
    if ((n == 4 || n == 8 || n == 16 || n == 32) && (((uintptr_t)d_x | (uintptr_t)d_y) & 15) == 0 && !getenv("LIQUID_CUDA_FFT_NO_TILE") &&
        !getenv("LIQUID_CUDA_NO_FUSED")) {
        if (n == 4) launch_fft_small<4>(sign, d_x, d_y, batch, st, post_mode, a, b);
        else if (n == 8) launch_fft_small<8>(sign, d_x, d_y, batch, st, post_mode, a, b);
        else if (n == 16) launch_fft_small<16>(sign, d_x, d_y, batch, st, post_mode, a, b);
        else launch_fft_small<32>(sign, d_x, d_y, batch, st, post_mode, a, b);
        LC_LAUNCH_CHECK();
        return;
    }
    if (d_tw_tile && batch < (1ull << 40) && !getenv("LIQUID_CUDA_FFT_NO_TILE") && !getenv("LIQUID_CUDA_NO_FUSED")) {
        switch (n) {
        case 64: launch_fft_tile_M<64>(sign, d_x, d_y, d_tw_tile, batch, st, post_mode, a, b); break;
        case 128: launch_fft_tile_M<128>(sign, d_x, d_y, d_tw_tile, batch, st, post_mode, a, b); break;
        case 256: launch_fft_tile_M<256>(sign, d_x, d_y, d_tw_tile, batch, st, post_mode, a, b); break;
        case 512: launch_fft_tile_M<512>(sign, d_x, d_y, d_tw_tile, batch, st, post_mode, a, b); break;
        case 1024: launch_fft_quad_RA<4>(sign, d_x, d_y, d_tw_tile, batch, st, post_mode, a, b); break;
        case 2048: launch_fft_quad_RA<8>(sign, d_x, d_y, d_tw_tile, batch, st, post_mode, a, b); break;
        default: launch_fft_quad_RA<16>(sign, d_x, d_y, d_tw_tile, batch, st, post_mode, a, b); break;
        }
        LC_LAUNCH_CHECK();
        return;
    }
    if (lcfft::pow2_in_smem(n)) {
        // aim at 2048 points per CTA
        const unsigned int per_cta = lcfft::per_cta_for(n, batch);
        const bool twg = n >= lcfft::kTwGlobalFrom;
        const size_t smem = lcfft::smem_for(n, per_cta, !twg);
        const unsigned long long max_grid = 1ull << 30;
        unsigned long long done = 0;
        while (done < batch) {      // grid.x limit is 2^31-1; keep launches below 2^30 CTAs
            unsigned long long nb = batch - done;
            unsigned long long ctas = (nb + per_cta - 1) / per_cta;
            if (ctas > max_grid) { ctas = max_grid; nb = ctas * per_cta; }
            if (twg)
                fft_pow2_kernel<true><<<(unsigned int)ctas, kFftThreads, smem, st>>>(
                    d_x + done * n, d_y + done * n, d_twp, n, ilog2(n), per_cta, nb, (float)sign, post_mode, a, b);
            else
                fft_pow2_kernel<false><<<(unsigned int)ctas, kFftThreads, smem, st>>>(
                    d_x + done * n, d_y + done * n, d_twp, n, ilog2(n), per_cta, nb, (float)sign, post_mode, a, b);
            LC_LAUNCH_CHECK();
            done += nb;
        }
        return;
    }
    if (n_rad > 0 && !getenv("LIQUID_CUDA_FFT_NO_MIXED")) {
        MixedRadices rad;
        rad.count = n_rad;
        for (unsigned int i = 0; i < 16; i++) rad.r[i] = radices[i];
        unsigned int per_cta = n >= 1024 ? 1 : 2048 / n;               // ~2048 points per CTA
        if ((unsigned long long)per_cta > batch) per_cta = (unsigned int)batch;
        const size_t smem = (size_t)per_cta * n * 2 * sizeof(float2);
        const unsigned long long ctas = (batch + per_cta - 1) / per_cta;
        if (ctas < (1ull << 31)) {
            fft_mixed_kernel<<<(unsigned int)ctas, kFftThreads, smem, st>>>(d_x, d_y, d_tw, n, rad, per_cta, batch, post_mode, a, b);
            LC_LAUNCH_CHECK();
            return;
        }
    }
    if ((size_t)n * sizeof(float2) > 200 * 1024)
        fatal("fft_execute", "transform size too large for this build (non power-of-two sizes above 25600 are not supported)");
    dft_any_kernel<<<(unsigned int)batch, kFftThreads, n * sizeof(float2), st>>>(d_x, d_y, d_tw, n, post_mode, a, b);
    LC_LAUNCH_CHECK();
}

}  // namespace lc

// ------------------------------------------------------------------------------------------------
// C ABI: liquid.h:1148-1215
// ------------------------------------------------------------------------------------------------
struct fftplan_s {
    unsigned int n;
    int dir;
    lc_cf *x, *y;          // borrowed host arrays (fft_common.c:35-36: "not allocated")
    lc::FftPlanDev dev;
    lc::DevBuf din, dout;
    cudaStream_t stream;
};

extern "C" {

fftplan fft_create_plan(unsigned int _n, liquid_float_complex *_x, liquid_float_complex *_y, int _dir, int _flags)
{
    (void)_flags;
    if (_n == 0) lc::fatal("fft_create_plan", "fft size must be > 0");
    if (_dir != LIQUID_FFT_FORWARD && _dir != LIQUID_FFT_BACKWARD)
        lc::fatal("fft_create_plan", "unknown/invalid fft direction (only complex FORWARD/BACKWARD are provided)");
    lc::ensure_device();
    fftplan p = new fftplan_s();
    p->n = _n;
    p->dir = _dir;
    p->x = reinterpret_cast<lc_cf *>(_x);
    p->y = reinterpret_cast<lc_cf *>(_y);
    p->dev.init(_n, _dir == LIQUID_FFT_FORWARD ? -1 : +1);
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    return p;
}

void fft_destroy_plan(fftplan _p)
{
    if (!_p) return;
    _p->dev.destroy();
    _p->din.release();
    _p->dout.release();
    cudaStreamDestroy(_p->stream);
    delete _p;
}

void fft_print_plan(fftplan _p)
{
    printf("fft plan [%s], n=%u, %s\n", _p->dir == LIQUID_FFT_FORWARD ? "forward" : "reverse", _p->n,
           lc::is_pow2(_p->n) ? "CUDA Stockham radix-4" : (_p->dev.n_rad ? "CUDA mixed-radix Stockham" : "CUDA DFT"));
}

void fft_execute_batch_dev(fftplan _p, const void *_d_x, void *_d_y, unsigned long long _batch, void *_stream)
{
    _p->dev.run((const float2 *)_d_x, (float2 *)_d_y, _batch, lc::as_stream(_stream));
}

void fft_execute_batch(fftplan _p, liquid_float_complex *_x, liquid_float_complex *_y, unsigned int _batch)
{
    size_t bytes = (size_t)_batch * _p->n * sizeof(lc_cf);
    float2 *din = (float2 *)_p->din.get(bytes);
    float2 *dout = (float2 *)_p->dout.get(bytes);
    LC_CUDA_CHECK(cudaMemcpyAsync(din, _x, bytes, cudaMemcpyHostToDevice, _p->stream));
    _p->dev.run(din, dout, _batch, _p->stream);
    LC_CUDA_CHECK(cudaMemcpyAsync(_y, dout, bytes, cudaMemcpyDeviceToHost, _p->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_p->stream));
}

void fft_execute(fftplan _p)
{
    fft_execute_batch(_p, reinterpret_cast<liquid_float_complex *>(_p->x),
                      reinterpret_cast<liquid_float_complex *>(_p->y), 1);
}

void fft_run(unsigned int _n, liquid_float_complex *_x, liquid_float_complex *_y, int _dir, int _flags)
{
    fftplan p = fft_create_plan(_n, _x, _y, _dir, _flags);
    fft_execute(p);
    fft_destroy_plan(p);
}

// host-side index shuffle (fft_common.c:336-350)
void fft_shift(liquid_float_complex *_x, unsigned int _n)
{
    lc_cf *x = reinterpret_cast<lc_cf *>(_x);
    unsigned int n2 = (_n % 2) ? (_n - 1) / 2 : _n / 2;
    for (unsigned int i = 0; i < n2; i++) {
        lc_cf t = x[i];
        x[i] = x[i + n2];
        x[i + n2] = t;
    }
}

}  // extern "C"
