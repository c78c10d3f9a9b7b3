// spgram.cu -- spgramcf: windowed-FFT spectral periodogram, the usual consumer right after a channelizer.
//
// Replaces src/fft/src/spgramcf.c -> src/fft/src/spgram.c (create :57-107, create_kaiser :113-147,
// create_default :150-162, reset :180-199, push/write :202-219, execute :226-244, execute_psd :250-264,
// accumulate_psd :271-311, write_accumulation :318-334, estimate_psd :341-398).  SURVEY.md 8(f) item 2.
//
// The reference takes one transform at a time inside a per-sample loop; here a whole block of samples
// becomes a batch of T overlapping transforms:
//   spgram_accumulate_pow2_kernel  (power-of-two nfft that fits shared memory: the throughput path) CTA c walks a
//                                  run of consecutive transforms: taper straight into shared memory (zero-padded
//                                  to nfft), Stockham FFT of fft_device.cuh there, and |X|^2 folded into register
//                                  accumulators with the reference's recurrence
//                                      psd <- (1-alpha) psd + alpha |X|^2      (plain sum for estimate_psd)
//                                  in transform order; it leaves one partial spectrum.  Neither the tapered
//                                  segments nor the T spectra visit HBM.  Chunk 0 starts from the object's psd, so
//                                  a block of up to 64 transforms is folded in exactly the reference's order.
//   spgram_combine_kernel          psd <- decay_c psd + partial_c over the chunks, in order
//   spgram_transform_pow2_kernel   one transform's X (spgramcf_execute)
//   any other nfft                 gather kernel -> batched DFT of fft.cu -> |X|^2 -> spgram_fold_kernel -> combine
// Same state as the reference: the last window_len samples, psd[nfft] (1 after reset), sample_counter,
// num_transforms -- including the quirk that the first transform after a reset forces alpha = 1 for the
// rest of that accumulate_psd call (alpha is the function's own parameter, spgram.c:291-292).
#include <cmath>
#include <cstdint>

#include "lc_common.cuh"
#include "fft_device.cuh"
#include "tile_fft.cuh"
#include "quad_fft.cuh"
#include "../../include/liquid_cuda.h"
#include "host_pipeline.cuh"

struct spgramcf_s {
    unsigned int nfft, window_len;
    std::vector<float> w;           // scaled tapering window (host copy)
    float *d_w = nullptr;
    float2 *d_buf = nullptr;        // last window_len samples, oldest first
    float *d_psd = nullptr;         // [nfft] accumulation
    unsigned int sample_counter = 0, num_transforms = 0;
    lc::FftPlanDev fft;             // forward, nfft points
    lc::DevBuf spectra, partials, stage, hist_tmp, one, xout;
    cudaStream_t stream = nullptr;
};

namespace {

constexpr int kThreads = 256;
constexpr unsigned int kFoldChunk = 64;                    // transforms folded sequentially by one thread
constexpr size_t kScratchBytes = (size_t)256 << 20;        // spectra of one pass

// sample `idx` of the stream whose samples [-H, 0) live in hist and [0, n) in x (x may be null when n = 0)
__device__ __forceinline__ float2 ld_sample(const float2 *__restrict__ hist, const float2 *__restrict__ x,
                                            long long idx, int H)
{
    return idx >= 0 ? __ldg(x + idx) : __ldg(hist + (H + idx));
}

// transform t of the pass ends (newest sample) at stream index min(e0 + t*hop, e_max)
struct Segs {
    long long e0, hop, e_max;
    __device__ __forceinline__ long long end_of(unsigned long long t) const
    {
        const long long e = e0 + (long long)t * hop;
        return e < e_max ? e : e_max;
    }
};

// first-pass source of the periodogram transforms: element k of transform t is the tapered sample (zero beyond
// the window), read straight into the butterfly's registers                              spgram.c:233-237
struct TaperSrc {
    const float2 *hist, *x;
    const float *w;
    int H;
    unsigned int window_len;
    Segs segs;
    unsigned long long t0, t_end;          // transform t0 + f, valid below t_end
    __device__ __forceinline__ float2 operator()(unsigned int f, unsigned int k) const
    {
        const unsigned long long t = t0 + f;
        if (t >= t_end || k >= window_len) return make_float2(0.f, 0.f);
        const float2 s = ld_sample(hist, x, segs.end_of(t) - (long long)(window_len - 1) + k, H);
        const float c = __ldg(w + k);
        return make_float2(s.x * c, s.y * c);
    }
};

// last-pass destination of spgram_transform_pow2_kernel: |X|^2 (MODE 0) or X (MODE 1) to global memory
template <int MODE>
struct SpectrumDst {
    void *out;
    unsigned int nfft;
    unsigned long long t0, t_end;
    __device__ __forceinline__ void operator()(unsigned int f, unsigned int k, float2 X) const
    {
        const unsigned long long t = t0 + f;
        if (t >= t_end) return;
        if (MODE == 0) reinterpret_cast<float *>(out)[t * nfft + k] = X.x * X.x + X.y * X.y;
        else           reinterpret_cast<float2 *>(out)[t * nfft + k] = X;
    }
};

// power-of-two nfft: taper, transform and |X|^2 (MODE 0 -> float out) or X (MODE 1 -> float2 out) in one kernel
template <int MODE>
__global__ void __launch_bounds__(lcfft::kFftThreads)
spgram_transform_pow2_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                             const float *__restrict__ w, unsigned int window_len, const float2 *__restrict__ tw,
                             unsigned int nfft, unsigned int log2n, unsigned int per_cta, unsigned long long T,
                             Segs segs, void *__restrict__ out)
{
    extern __shared__ float2 smem_f2[];
    const lcfft::Smem sm = lcfft::carve(smem_f2, nfft, per_cta);
    if (log2n > 4) lcfft::stage_twiddles(sm, tw, nfft);
    const unsigned long long first = (unsigned long long)blockIdx.x * per_cta;
    const TaperSrc src{hist, x, w, H, window_len, segs, first, T};
    const SpectrumDst<MODE> dst{out, nfft, first, T};
    lcfft::stockham_run<-1>(sm, nfft, log2n, per_cta, src, dst);
}

// power-of-two nfft, the throughput path of accumulate_psd / estimate_psd: CTA c walks the `tpc` consecutive
// transforms [c tpc, (c+1) tpc) -- tapered samples straight into the first FFT pass's registers, Stockham passes
// in shared memory, and the fold of |X|^2 into per-thread register accumulators (bins tid, tid+256, ...) in
// transform order -- and leaves ONE partial spectrum; neither the tapered segments nor the T spectra ever visit
// HBM.  Consecutive transforms overlap by 50-75 % of their input, which L1/L2 serve.
__global__ void __launch_bounds__(lcfft::kFftThreads)
spgram_accumulate_pow2_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                              const float *__restrict__ w, unsigned int window_len, const float2 *__restrict__ tw,
                              unsigned int nfft, unsigned int log2n, unsigned int per_cta, unsigned long long T,
                              unsigned int tpc, Segs segs, const float *__restrict__ psd, float *__restrict__ partial,
                              float alpha, int seed_psd)
{
    extern __shared__ float2 smem_f2[];
    const lcfft::Smem sm = lcfft::carve(smem_f2, nfft, per_cta);
    lcfft::stage_twiddles(sm, tw, nfft);
    const unsigned long long t_begin = (unsigned long long)blockIdx.x * tpc;
    unsigned long long t_end = t_begin + tpc;
    if (t_end > T) t_end = T;
    constexpr int kMaxBins = 4096 / lcfft::kFftThreads;                // bins per thread at nfft = 4096
    float acc[kMaxBins];
#pragma unroll
    for (int b = 0; b < kMaxBins; b++) {
        const unsigned int k = threadIdx.x + b * lcfft::kFftThreads;
        acc[b] = (seed_psd && blockIdx.x == 0 && k < nfft) ? psd[k] : 0.0f;
    }
    const float keep = 1.0f - alpha;
    __syncthreads();                                                   // twiddle table staged
    for (unsigned long long g0 = t_begin; g0 < t_end; g0 += per_cta) {
        const TaperSrc src{hist, x, w, H, window_len, segs, g0, t_end};
        const float2 *res = lcfft::stockham_run_to_smem<-1>(sm, nfft, log2n, per_cta, src);
        for (unsigned int tt = 0; tt < per_cta && g0 + tt < t_end; tt++) {
#pragma unroll
            for (int b = 0; b < kMaxBins; b++) {
                const unsigned int k = threadIdx.x + b * lcfft::kFftThreads;
                if (k < nfft) {
                    const float2 X = res[lcfft::pidx(tt * nfft + k)];
                    const float P = X.x * X.x + X.y * X.y;
                    acc[b] = (alpha < 0.0f) ? acc[b] + P : keep * acc[b] + alpha * P;
                }
            }
        }
        __syncthreads();                                               // the buffers are rewritten next
    }
#pragma unroll
    for (int b = 0; b < kMaxBins; b++) {
        const unsigned int k = threadIdx.x + b * lcfft::kFftThreads;
        if (k < nfft) partial[(size_t)blockIdx.x * nfft + k] = acc[b];
    }
}

// ---- nfft = 64 ... 4096 on the tile / quad transforms (tile_fft.cuh, quad_fft.cuh): the same chunked fold as
// spgram_accumulate_pow2_kernel -- instance (tile) or CTA (quad) c folds `chunk` consecutive transforms in order and leaves
// partial spectrum c for spgram_combine_kernel -- without the block-wide shared-memory passes
struct TaperSrcT {
    const float2 *hist, *x;
    const float *w;
    int H;
    unsigned int window_len;
    Segs segs;
    // the raw sample (zero beyond the window), loaded a tile ahead; the taper is applied where the value is consumed
    __device__ __forceinline__ float2 load(long long t, unsigned int k) const
    {
        if (k >= window_len) return make_float2(0.f, 0.f);
        return ld_sample(hist, x, segs.end_of((unsigned long long)t) - (long long)(window_len - 1) + k, H);
    }
    __device__ __forceinline__ float2 finish(long long, unsigned int k, float2 raw) const
    {
        if (k >= window_len) return raw;
        const float c = __ldg(w + k);
        return make_float2(raw.x * c, raw.y * c);
    }
};

// tile transform: thread u is handed outputs u and u + M/2 of every transform of its instance, in order
struct FoldDstT {
    float lo, hi, alpha, keep;
    unsigned int M2;
    __device__ __forceinline__ void operator()(long long, unsigned int k, float2 X)
    {
        const float P = X.x * X.x + X.y * X.y;
        float &a = k < M2 ? lo : hi;
        a = (alpha < 0.0f) ? a + P : keep * a + alpha * P;
    }
};

template <int M>
__global__ void __launch_bounds__(lctile::kThreads, 2)
spgram_accumulate_tile_kernel(TaperSrcT src, const float2 *__restrict__ tw, long long T, int tiles_per_inst, unsigned int n_chunks,
                              const float *__restrict__ psd, float *__restrict__ partial, float alpha)
{
    extern __shared__ __align__(16) unsigned char smem_tile[];
    constexpr int M2 = M / 2;
    const unsigned int u = threadIdx.x % M2;
    const unsigned int inst = blockIdx.x * (lctile::kThreads / M2) + threadIdx.x / M2;
    FoldDstT dst{inst == 0 ? psd[u] : 0.0f, inst == 0 ? psd[u + M2] : 0.0f, alpha, 1.0f - alpha, (unsigned int)M2};
    lctile::run<M, -1>(smem_tile, tw, T, tiles_per_inst, src, dst);
    if (inst < n_chunks) {
        partial[(size_t)inst * M + u] = dst.lo;
        partial[(size_t)inst * M + u + M2] = dst.hi;
    }
}

// quad transform: the FR transforms of a tile leave through different thread groups, so |X|^2 goes through a
// shared [FR][N] buffer and thread tid folds bins tid + 256 b in transform order after a block barrier
template <int RA>
struct PowerDstQ {
    float *pbuf;
    __device__ __forceinline__ void operator()(long long t, unsigned int k, float2 X) const
    {
        pbuf[(unsigned int)(t % (16 / RA)) * (256 * RA) + k] = X.x * X.x + X.y * X.y;
    }
};
template <int RA>
struct FoldPostQ {
    const float *pbuf;
    float acc[RA];
    float alpha, keep;
    __device__ __forceinline__ void operator()(long long, int n_valid)
    {
        __syncthreads();
        for (int f = 0; f < n_valid; f++) {
#pragma unroll
            for (int b = 0; b < RA; b++) {
                const float P = pbuf[f * (256 * RA) + threadIdx.x + 256 * b];
                acc[b] = (alpha < 0.0f) ? acc[b] + P : keep * acc[b] + alpha * P;
            }
        }
    }
};

template <int RA>
__global__ void __launch_bounds__(lcquad::kThreads, 2)
spgram_accumulate_quad_kernel(TaperSrcT src, const float2 *__restrict__ tw, long long T, int tiles_per_inst,
                              const float *__restrict__ psd, float *__restrict__ partial, float alpha)
{
    extern __shared__ __align__(16) unsigned char smem_quad[];
    constexpr int N = 256 * RA;
    float *pbuf = reinterpret_cast<float *>(smem_quad + lcquad::Pitch<RA>::kBytes);
    FoldPostQ<RA> post;
    post.pbuf = pbuf;
    post.alpha = alpha;
    post.keep = 1.0f - alpha;
#pragma unroll
    for (int b = 0; b < RA; b++) post.acc[b] = blockIdx.x == 0 ? psd[threadIdx.x + 256 * b] : 0.0f;
    lcquad::run<RA, -1, (RA >= 8)>(smem_quad, tw, T, tiles_per_inst, src, PowerDstQ<RA>{pbuf}, post);
#pragma unroll
    for (int b = 0; b < RA; b++) partial[(size_t)blockIdx.x * N + threadIdx.x + 256 * b] = post.acc[b];
}

template <int M>
void launch_accumulate_tile(const TaperSrcT &src, const float2 *d_tw, unsigned long long T, int tpi, unsigned int n_chunks,
                            const float *d_psd, float *d_part, float alpha, cudaStream_t st)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(spgram_accumulate_tile_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)lctile::smem_bytes<M>()));
    const unsigned int ipc = lctile::kThreads / (M / 2);
    spgram_accumulate_tile_kernel<M><<<(n_chunks + ipc - 1) / ipc, lctile::kThreads, lctile::smem_bytes<M>(), st>>>(
        src, d_tw, (long long)T, tpi, n_chunks, d_psd, d_part, alpha);
}

template <int RA>
void launch_accumulate_quad(const TaperSrcT &src, const float2 *d_tw, unsigned long long T, int tpi, unsigned int n_chunks,
                            const float *d_psd, float *d_part, float alpha, cudaStream_t st)
{
    const size_t smem = lcquad::smem_bytes<RA>() + sizeof(float) * 4096;
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(spgram_accumulate_quad_kernel<RA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spgram_accumulate_quad_kernel<RA><<<n_chunks, lcquad::kThreads, smem, st>>>(src, d_tw, (long long)T, tpi, d_psd, d_part, alpha);
}

// any nfft: tapered, zero-padded segments [T][nfft] for the batched DFT
__global__ void __launch_bounds__(kThreads)
spgram_gather_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H, const float *__restrict__ w,
                     unsigned int window_len, unsigned int nfft, unsigned long long T, Segs segs, float2 *__restrict__ out)
{
    const unsigned long long total = T * nfft;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long t = i / nfft;
        const unsigned int k = (unsigned int)(i % nfft);
        float2 v = make_float2(0.f, 0.f);
        if (k < window_len) {
            const float2 s = ld_sample(hist, x, segs.end_of(t) - (long long)(window_len - 1) + k, H);
            const float c = __ldg(w + k);
            v = make_float2(s.x * c, s.y * c);
        }
        out[i] = v;
    }
}

__global__ void __launch_bounds__(kThreads)
spgram_magsq_kernel(const float2 *__restrict__ X, float *__restrict__ P, unsigned long long total)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float2 v = X[i];
        P[i] = v.x * v.x + v.y * v.y;
    }
}

// fold chunk c (transforms [c*chunk, min((c+1)*chunk, T))) of P[T][nfft] per bin k:
//   s = (seed_psd && c == 0) ? psd[k] : 0;   for t: s = (1-alpha) s + alpha P[t][k]   (alpha < 0: s += P[t][k])
__global__ void __launch_bounds__(kThreads)
spgram_fold_kernel(const float *__restrict__ P, const float *__restrict__ psd, float *__restrict__ partial,
                   unsigned int nfft, unsigned long long T, unsigned int chunk, float alpha, int seed_psd)
{
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int c = blockIdx.y;
    if (k >= nfft) return;
    const unsigned long long t0 = (unsigned long long)c * chunk;
    unsigned long long t1 = t0 + chunk;
    if (t1 > T) t1 = T;
    float s = (seed_psd && c == 0) ? psd[k] : 0.0f;
    if (alpha < 0.0f) {
        for (unsigned long long t = t0; t < t1; t++) s += __ldg(P + t * nfft + k);
    } else {
        const float keep = 1.0f - alpha;
        for (unsigned long long t = t0; t < t1; t++) s = keep * s + alpha * __ldg(P + t * nfft + k);
    }
    partial[(size_t)c * nfft + k] = s;
}

// psd[k] = partial[0][k]; for c >= 1: psd[k] = decay(c) psd[k] + partial[c][k], decay(c) = (1-alpha)^(transforms in
// chunk c) (all chunks but the last hold `chunk` transforms), or 1 for the plain sum
__global__ void __launch_bounds__(kThreads)
spgram_combine_kernel(const float *__restrict__ partial, float *__restrict__ psd, unsigned int nfft,
                      unsigned int n_chunks, unsigned int chunk, unsigned long long T, float alpha)
{
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nfft) return;
    float d_full = 1.0f, d_last = 1.0f;
    if (alpha >= 0.0f) {
        d_full = powf(1.0f - alpha, (float)chunk);
        d_last = powf(1.0f - alpha, (float)(T - (unsigned long long)(n_chunks - 1) * chunk));
    }
    float s = partial[k];
    for (unsigned int c = 1; c < n_chunks; c++)
        s = ((c + 1 == n_chunks) ? d_last : d_full) * s + __ldg(partial + (size_t)c * nfft + k);
    psd[k] = s;
}

inline unsigned int grid_for(unsigned long long total)
{
    unsigned long long b = (total + kThreads - 1) / kThreads;
    const unsigned long long cap = 148ull * 8 * 16;
    return (unsigned int)(b < cap ? (b ? b : 1) : cap);
}

// T transforms of the (hist | x) stream -> d_out: |X|^2 [T][nfft] floats (mode 0) or X [T][nfft] (mode 1)
void run_transforms(spgramcf q, const float2 *d_x, unsigned long long T, Segs segs, int mode, void *d_out, cudaStream_t st)
{
    if (T == 0) return;
    const unsigned int nfft = q->nfft;
    if (lcfft::pow2_in_smem(nfft)) {
        static std::atomic<unsigned long long> configured{0};
        if (lc::first_time_on_device(configured)) {
            LC_CUDA_CHECK(cudaFuncSetAttribute(spgram_transform_pow2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            LC_CUDA_CHECK(cudaFuncSetAttribute(spgram_transform_pow2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
        const unsigned int per_cta = lcfft::per_cta_for(nfft, T);
        const unsigned long long ctas = (T + per_cta - 1) / per_cta;
        const size_t smem = lcfft::smem_for(nfft, per_cta);
        if (mode == 0)
            spgram_transform_pow2_kernel<0><<<(unsigned int)ctas, lcfft::kFftThreads, smem, st>>>(
                q->d_buf, d_x, (int)q->window_len, q->d_w, q->window_len, q->fft.d_twp, nfft, lc::ilog2(nfft), per_cta, T, segs, d_out);
        else
            spgram_transform_pow2_kernel<1><<<(unsigned int)ctas, lcfft::kFftThreads, smem, st>>>(
                q->d_buf, d_x, (int)q->window_len, q->d_w, q->window_len, q->fft.d_twp, nfft, lc::ilog2(nfft), per_cta, T, segs, d_out);
        LC_LAUNCH_CHECK();
        return;
    }
    // any other size: tapered segments -> batched DFT (in place) -> |X|^2
    float2 *d_X = (mode == 1) ? (float2 *)d_out : (float2 *)q->xout.get(sizeof(float2) * T * nfft);
    spgram_gather_kernel<<<grid_for(T * nfft), kThreads, 0, st>>>(q->d_buf, d_x, (int)q->window_len, q->d_w, q->window_len,
                                                                 nfft, T, segs, d_X);
    LC_LAUNCH_CHECK();
    q->fft.run(d_X, d_X, T, st);
    if (mode == 0) {
        spgram_magsq_kernel<<<grid_for(T * nfft), kThreads, 0, st>>>(d_X, (float *)d_out, T * nfft);
        LC_LAUNCH_CHECK();
    }
}

// fold T spectra-in-time into q->d_psd: alpha in [0,1] -> the reference's recurrence seeded with psd;
// alpha < 0 -> plain sum on top of psd
void fold_into_psd(spgramcf q, const float *d_P, unsigned long long T, float alpha, int seed_psd, cudaStream_t st)
{
    const unsigned int nfft = q->nfft;
    unsigned int chunk = kFoldChunk;
    if ((T + chunk - 1) / chunk > 256) chunk = (unsigned int)((T + 255) / 256);
    const unsigned int n_chunks = (unsigned int)((T + chunk - 1) / chunk);
    float *d_part = (float *)q->partials.get(sizeof(float) * (size_t)n_chunks * nfft);
    dim3 grid((nfft + kThreads - 1) / kThreads, n_chunks);
    spgram_fold_kernel<<<grid, kThreads, 0, st>>>(d_P, q->d_psd, d_part, nfft, T, chunk, alpha, seed_psd);
    LC_LAUNCH_CHECK();
    spgram_combine_kernel<<<(nfft + kThreads - 1) / kThreads, kThreads, 0, st>>>(d_part, q->d_psd, nfft, n_chunks, chunk, T, alpha);
    LC_LAUNCH_CHECK();
}

// all T transforms of the (hist | x) stream folded into q->d_psd (alpha < 0: plain sum on top of it)
void accumulate_transforms(spgramcf q, const float2 *d_x, unsigned long long T, Segs segs, float alpha, cudaStream_t st)
{
    if (T == 0) return;
    const unsigned int nfft = q->nfft;
    if (q->fft.d_tw_tile && T < (1ull << 31) && !getenv("LIQUID_CUDA_SPGRAM_NO_TILE") && !getenv("LIQUID_CUDA_NO_FUSED")) {
        // transforms per tile, instances per CTA; chunks = about two waves of the CTAs that fit (2 per SM)
        const unsigned int per_tile = nfft <= 512 ? lctile::kFr : 4096 / nfft;
        const unsigned int ipc = nfft <= 512 ? lctile::kThreads / (nfft / 2) : 1;
        const unsigned long long n_tiles = (T + per_tile - 1) / per_tile;
        const unsigned long long target = 148ull * 2 * 2 * ipc;
        unsigned long long tpi = (n_tiles + target - 1) / target;
        if (tpi < 1) tpi = 1;
        const unsigned int n_chunks = (unsigned int)((n_tiles + tpi - 1) / tpi);
        float *d_part = (float *)q->partials.get(sizeof(float) * (size_t)n_chunks * nfft);
        const TaperSrcT src{q->d_buf, d_x, q->d_w, (int)q->window_len, q->window_len, segs};
        switch (nfft) {
        case 64: launch_accumulate_tile<64>(src, q->fft.d_tw_tile, T, (int)tpi, n_chunks, q->d_psd, d_part, alpha, st); break;
        case 128: launch_accumulate_tile<128>(src, q->fft.d_tw_tile, T, (int)tpi, n_chunks, q->d_psd, d_part, alpha, st); break;
        case 256: launch_accumulate_tile<256>(src, q->fft.d_tw_tile, T, (int)tpi, n_chunks, q->d_psd, d_part, alpha, st); break;
        case 512: launch_accumulate_tile<512>(src, q->fft.d_tw_tile, T, (int)tpi, n_chunks, q->d_psd, d_part, alpha, st); break;
        case 1024: launch_accumulate_quad<4>(src, q->fft.d_tw_tile, T, (int)tpi, n_chunks, q->d_psd, d_part, alpha, st); break;
        case 2048: launch_accumulate_quad<8>(src, q->fft.d_tw_tile, T, (int)tpi, n_chunks, q->d_psd, d_part, alpha, st); break;
        default: launch_accumulate_quad<16>(src, q->fft.d_tw_tile, T, (int)tpi, n_chunks, q->d_psd, d_part, alpha, st); break;
        }
        LC_LAUNCH_CHECK();
        spgram_combine_kernel<<<(nfft + kThreads - 1) / kThreads, kThreads, 0, st>>>(d_part, q->d_psd, nfft, n_chunks,
                                                                                  (unsigned int)(tpi * per_tile), T, alpha);
        LC_LAUNCH_CHECK();
        return;
    }
    if (lcfft::pow2_in_smem(nfft)) {
        static std::atomic<unsigned long long> configured{0};
        if (lc::first_time_on_device(configured))
            LC_CUDA_CHECK(cudaFuncSetAttribute(spgram_accumulate_pow2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        // transforms per CTA: a block of up to 64 transforms is one chunk (folded in exactly the reference's
        // order); beyond that aim at ~8 CTAs per SM, at most 2048 chunks for the combine pass
        unsigned long long tpc = kFoldChunk;
        if (T > tpc * 148 * 8) tpc = (T + 148 * 8 - 1) / (148 * 8);
        if ((T + tpc - 1) / tpc > 2048) tpc = (T + 2047) / 2048;
        const unsigned int per_cta = lcfft::per_cta_for(nfft, tpc, 4096);
        tpc = (tpc + per_cta - 1) / per_cta * per_cta;
        const unsigned int n_chunks = (unsigned int)((T + tpc - 1) / tpc);
        float *d_part = (float *)q->partials.get(sizeof(float) * (size_t)n_chunks * nfft);
        spgram_accumulate_pow2_kernel<<<n_chunks, lcfft::kFftThreads, lcfft::smem_for(nfft, per_cta), st>>>(
            q->d_buf, d_x, (int)q->window_len, q->d_w, q->window_len, q->fft.d_twp, nfft, lc::ilog2(nfft), per_cta, T,
            (unsigned int)tpc, segs, q->d_psd, d_part, alpha, 1);
        LC_LAUNCH_CHECK();
        spgram_combine_kernel<<<(nfft + kThreads - 1) / kThreads, kThreads, 0, st>>>(d_part, q->d_psd, nfft, n_chunks,
                                                                                  (unsigned int)tpc, T, alpha);
        LC_LAUNCH_CHECK();
        return;
    }
    // any other nfft: passes of gather -> batched DFT -> |X|^2 -> fold, cut so that the spectra fit the scratch budget
    unsigned long long lim = kScratchBytes / ((unsigned long long)nfft * sizeof(float2));
    if (lim == 0) lim = 1;
    for (unsigned long long done = 0; done < T;) {
        const unsigned long long nt = (T - done < lim) ? T - done : lim;
        float *d_P = (float *)q->spectra.get(sizeof(float) * nt * nfft);
        Segs sub{segs.e0 + (long long)done * segs.hop, segs.hop, segs.e_max};
        run_transforms(q, d_x, nt, sub, 0, d_P, st);
        fold_into_psd(q, d_P, nt, alpha, 1, st);
        done += nt;
    }
}

void set_psd_ones(spgramcf q)
{
    std::vector<float> ones(q->nfft, 1.0f);
    LC_CUDA_CHECK(cudaMemcpyAsync(q->d_psd, ones.data(), sizeof(float) * q->nfft, cudaMemcpyHostToDevice, q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(q->stream));
}

}  // namespace

extern "C" {

spgramcf spgramcf_create(unsigned int _nfft, float *_window, unsigned int _window_len)
{
    if (_nfft < 2) lc::fatal("spgramcf_create", "fft size must be at least 2");
    if (_window_len > _nfft) lc::fatal("spgramcf_create", "window size cannot exceed fft size");
    if (_window_len == 0) lc::fatal("spgramcf_create", "window size must be greater than zero");
    lc::ensure_device();
    spgramcf q = new spgramcf_s();
    q->nfft = _nfft;
    q->window_len = _window_len;
    // scale by window magnitude and FFT size: g = sqrt(2) / (rms(window) sqrt(nfft))      spgram.c:92-100
    float g = 0.0f;
    for (unsigned int i = 0; i < _window_len; i++) g += _window[i] * _window[i];
    g = (float)(1.41421356237309504880 / (double)(sqrtf(g / (float)_window_len) * sqrtf((float)_nfft)));
    q->w.resize(_window_len);
    for (unsigned int i = 0; i < _window_len; i++) q->w[i] = g * _window[i];
    LC_CUDA_CHECK(cudaMalloc(&q->d_w, sizeof(float) * _window_len));
    LC_CUDA_CHECK(cudaMemcpy(q->d_w, q->w.data(), sizeof(float) * _window_len, cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMalloc(&q->d_buf, sizeof(float2) * _window_len));
    LC_CUDA_CHECK(cudaMalloc(&q->d_psd, sizeof(float) * _nfft));
    q->fft.init(_nfft, -1);
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    spgramcf_reset(q);
    return q;
}

spgramcf spgramcf_create_kaiser(unsigned int _nfft, unsigned int _window_len, float _beta)
{
    if (_nfft < 2) lc::fatal("spgramcf_create_kaiser", "fft size must be at least 2");
    if (_window_len > _nfft) lc::fatal("spgramcf_create_kaiser", "window size cannot exceed fft size");
    if (_window_len == 0) lc::fatal("spgramcf_create_kaiser", "window size must be greater than zero");
    if (_beta <= 0.0f) lc::fatal("spgramcf_create_kaiser", "beta must be greater than zero");
    std::vector<float> w(_window_len);
    for (unsigned int i = 0; i < _window_len; i++) w[i] = lc::kaiser_window(i, _window_len, _beta, 0.0f);
    return spgramcf_create(_nfft, w.data(), _window_len);
}

spgramcf spgramcf_create_default(unsigned int _nfft)
{
    if (_nfft < 2) lc::fatal("spgramcf_create_default", "fft size must be at least 2");
    return spgramcf_create_kaiser(_nfft, _nfft / 2, 10.0f);       // spgram.c:158-161
}

void spgramcf_destroy(spgramcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    cudaFree(_q->d_w);
    cudaFree(_q->d_buf);
    cudaFree(_q->d_psd);
    _q->fft.destroy();
    _q->spectra.release();
    _q->partials.release();
    _q->stage.release();
    _q->hist_tmp.release();
    _q->one.release();
    _q->xout.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void spgramcf_reset(spgramcf _q)
{
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_buf, 0, sizeof(float2) * _q->window_len, _q->stream));
    _q->num_transforms = 0;
    _q->sample_counter = 0;
    set_psd_ones(_q);                    // unity = 0 dB (spgram.c:196-198)
}

void spgramcf_write(spgramcf _q, liquid_float_complex *_x, unsigned int _n)
{
    if (_n == 0) return;
    // only the last window_len samples can matter
    const unsigned int keep = _n < _q->window_len ? _n : _q->window_len;
    float2 *d = (float2 *)_q->stage.get(sizeof(float2) * keep);
    LC_CUDA_CHECK(cudaMemcpyAsync(d, reinterpret_cast<const lc_cf *>(_x) + (_n - keep), sizeof(float2) * keep,
                                  cudaMemcpyHostToDevice, _q->stream));
    lc::roll_history(_q->d_buf, _q->window_len, _q->hist_tmp, d, keep, _q->stream);
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

void spgramcf_push(spgramcf _q, liquid_float_complex _x) { spgramcf_write(_q, &_x, 1); }

void spgramcf_execute(spgramcf _q, liquid_float_complex *_X)
{
    // one transform of the buffer as it stands: the segment ends at stream index -1
    float2 *d_X = (float2 *)_q->spectra.get(sizeof(float2) * _q->nfft);
    Segs segs{-1, 0, -1};
    run_transforms(_q, nullptr, 1, segs, 1, d_X, _q->stream);
    if (_X) LC_CUDA_CHECK(cudaMemcpyAsync(_X, d_X, sizeof(float2) * _q->nfft, cudaMemcpyDeviceToHost, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

void spgramcf_execute_psd(spgramcf _q, float *_X)
{
    std::vector<lc_cf> X(_q->nfft);
    spgramcf_execute(_q, reinterpret_cast<liquid_float_complex *>(X.data()));
    const unsigned int h = _q->nfft / 2;
    for (unsigned int i = 0; i < _q->nfft; i++)
        _X[(i + h) % _q->nfft] = 10 * log10f(X[i].re * X[i].re + X[i].im * X[i].im + 1e-16f);     // spgram.c:259-263
}

// device-pointer form: _d_x holds _n new samples; state (buffer, psd, counters) advances as in the reference
void spgramcf_accumulate_psd_dev(spgramcf _q, const void *_d_x, float _alpha, unsigned long long _n, void *_stream)
{
    if (_alpha < 0.0f || _alpha > 1.0f) lc::fatal("spgramcf_accumulate_psd", "alpha must be in (0,1)");
    if (_n == 0) return;
    cudaStream_t st = lc::as_stream(_stream);
    const float2 *d_x = (const float2 *)_d_x;
    const unsigned int hop = _q->window_len / 2;                 // a transform every window_len/2 pushes
    unsigned long long T = 0;
    long long e0 = 0;
    if (hop > 0) {
        T = (_n + _q->sample_counter) / hop;
        e0 = (long long)hop - (long long)_q->sample_counter - 1;  // index in x of the push that triggers the first one
    }
    // the first transform after a reset forces alpha = 1 for the rest of this call (spgram.c:291-292)
    const float alpha = (_q->num_transforms == 0 && T > 0) ? 1.0f : _alpha;
    accumulate_transforms(_q, d_x, T, Segs{e0, (long long)hop, (long long)_n - 1}, alpha, st);
    lc::roll_history(_q->d_buf, _q->window_len, _q->hist_tmp, d_x, _n, st);
    if (hop > 0) {
        _q->sample_counter = (unsigned int)((_n + _q->sample_counter) % hop);
        _q->num_transforms += (unsigned int)T;
    } else {
        _q->sample_counter += (unsigned int)_n;                  // window_len = 1: the counter never meets 0
    }
}

void spgramcf_accumulate_psd(spgramcf _q, liquid_float_complex *_x, float _alpha, unsigned int _n)
{
    if (_alpha < 0.0f || _alpha > 1.0f) lc::fatal("spgramcf_accumulate_psd", "alpha must be in (0,1)");
    if (_n == 0) return;
    float2 *d = (float2 *)_q->stage.get(sizeof(float2) * (size_t)_n);
    LC_CUDA_CHECK(cudaMemcpyAsync(d, _x, sizeof(float2) * (size_t)_n, cudaMemcpyHostToDevice, _q->stream));
    spgramcf_accumulate_psd_dev(_q, d, _alpha, _n, _q->stream);
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

void spgramcf_write_accumulation(spgramcf _q, float *_x)
{
    std::vector<float> psd(_q->nfft);
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    LC_CUDA_CHECK(cudaMemcpy(psd.data(), _q->d_psd, sizeof(float) * _q->nfft, cudaMemcpyDeviceToHost));
    const unsigned int h = _q->nfft / 2;
    for (unsigned int i = 0; i < _q->nfft; i++) _x[(i + h) % _q->nfft] = 10 * log10f(psd[i]);      // spgram.c:332-333
}

// device-pointer form of estimate_psd: _psd (host, nfft floats) receives the fft-shifted average in dB
void spgramcf_estimate_psd_dev(spgramcf _q, const void *_d_x, unsigned long long _n, float *_psd, void *_stream)
{
    if (_n == 0) {
        fprintf(stderr, "warning: spgram%s_estimate_psd(), input size is zero\n", "cf");
        return;
    }
    cudaStream_t st = lc::as_stream(_stream);
    const float2 *d_x = (const float2 *)_d_x;
    spgramcf_reset(_q);
    LC_CUDA_CHECK(cudaStreamSynchronize(st));
    unsigned int delay = _q->nfft / 4;
    if (delay == 0) delay = 1;
    // a transform after every `delay` samples and after the last one            spgram.c:373-388
    const unsigned long long T = (_n + delay - 1) / delay;
    const unsigned int nfft = _q->nfft;
    // the sum is kept in d_psd while the passes run, then the reset state is restored
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_psd, 0, sizeof(float) * nfft, st));
    accumulate_transforms(_q, d_x, T, Segs{(long long)delay - 1, (long long)delay, (long long)_n - 1}, -1.0f, st);
    std::vector<float> sum(nfft);
    LC_CUDA_CHECK(cudaMemcpyAsync(sum.data(), _q->d_psd, sizeof(float) * nfft, cudaMemcpyDeviceToHost, st));
    lc::roll_history(_q->d_buf, _q->window_len, _q->hist_tmp, d_x, _n, st);
    LC_CUDA_CHECK(cudaStreamSynchronize(st));
    set_psd_ones(_q);                                           // estimate_psd leaves psd as reset left it
    for (unsigned int k = 0; k < nfft; k++)
        _psd[(k + nfft / 2) % nfft] = 10 * log10f(sum[k] / (float)T);
}

void spgramcf_estimate_psd(spgramcf _q, liquid_float_complex *_x, unsigned int _n, float *_psd)
{
    if (_n == 0) {
        fprintf(stderr, "warning: spgram%s_estimate_psd(), input size is zero\n", "cf");
        return;
    }
    float2 *d = (float2 *)_q->stage.get(sizeof(float2) * (size_t)_n);
    LC_CUDA_CHECK(cudaMemcpyAsync(d, _x, sizeof(float2) * (size_t)_n, cudaMemcpyHostToDevice, _q->stream));
    spgramcf_estimate_psd_dev(_q, d, _n, _psd, _q->stream);
}

unsigned int spgramcf_get_num_transforms(spgramcf _q) { return _q->num_transforms; }

}  // extern "C"

// ------------------------------------------------------------------------------------------ asgramcf
// ASCII spectrogram on top of spgramcf: liquid.h:1333-1395 -> src/fft/src/asgram.c (create :48-87, set_scale :113-128,
// execute :157-195, print :199-213).  The transform of the buffer is spgramcf's; one more tiny kernel shifts the
// spectrum, takes 10 log10 |X| (the reference's quantity: the magnitude, not its square), maps it to the ten level
// characters and finds the first maximum.
struct asgramcf_s {
    unsigned int nfft = 0;
    spgramcf periodogram = nullptr;
    float levels[10];
    char levelchar[10];
    float scale = 10.0f, offset = 0.0f;
    lc::DevBuf out;                 // [nfft] chars, then (peak value, peak frequency) floats
};

namespace {
struct AsgramLevels {
    float levels[10];
    char levelchar[10];
};

__global__ void asgram_levels_kernel(const float2 *__restrict__ X, unsigned int nfft, AsgramLevels lv, char *__restrict__ ascii,
                                     float *__restrict__ psd, float *__restrict__ peak)
{
    for (unsigned int i = threadIdx.x; i < nfft; i += blockDim.x) {
        const float2 v = X[(i + nfft / 2) % nfft];
        const float p = 10 * log10f(hypotf(v.x, v.y));
        psd[i] = p;
        char c = lv.levelchar[0];
        for (int j = 0; j < 10; j++)
            if (p > lv.levels[j]) c = lv.levelchar[j];
        ascii[i] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float best = psd[0];
        unsigned int arg = 0;
        for (unsigned int i = 1; i < nfft; i++)
            if (psd[i] > best) { best = psd[i]; arg = i; }
        peak[0] = best;
        peak[1] = (float)arg / (float)nfft - 0.5f;
    }
}
}  // namespace

extern "C" {

asgramcf asgramcf_create(unsigned int _nfft)
{
    if (_nfft < 2) lc::fatal("asgramcf_create", "fft size must be at least 2");
    asgramcf q = new asgramcf_s();
    q->nfft = _nfft;
    q->periodogram = spgramcf_create_kaiser(_nfft, _nfft, 10.0f);
    const char chars[10] = {' ', '.', ',', '-', '+', '*', '&', 'N', 'M', '#'};
    for (int i = 0; i < 10; i++) q->levelchar[i] = chars[i];
    asgramcf_set_scale(q, 0.0f, 10.0f);
    return q;
}

void asgramcf_destroy(asgramcf _q)
{
    if (!_q) return;
    spgramcf_destroy(_q->periodogram);
    _q->out.release();
    delete _q;
}

void asgramcf_reset(asgramcf _q) { spgramcf_reset(_q->periodogram); }

void asgramcf_set_scale(asgramcf _q, float _offset, float _scale)
{
    if (_scale <= 0.0f) lc::fatal("asgramcf_set_scale", "scale must be greater than zero");
    _q->offset = _offset;
    _q->scale = _scale;
    for (unsigned int i = 0; i < 10; i++) _q->levels[i] = _q->offset + i * _q->scale;
}

void asgramcf_push(asgramcf _q, liquid_float_complex _x) { spgramcf_push(_q->periodogram, _x); }
void asgramcf_write(asgramcf _q, liquid_float_complex *_x, unsigned int _n) { spgramcf_write(_q->periodogram, _x, _n); }

void asgramcf_execute(asgramcf _q, char *_ascii, float *_peakval, float *_peakfreq)
{
    spgramcf sp = _q->periodogram;
    spgramcf_execute(sp, nullptr);                                  // the spectrum stays on the device
    const float2 *d_X = (const float2 *)sp->spectra.get(sizeof(float2) * _q->nfft);
    const size_t chars = ((size_t)_q->nfft + 15) & ~(size_t)15;
    char *d_out = (char *)_q->out.get(chars + sizeof(float) * ((size_t)_q->nfft + 2));
    float *d_psd = (float *)(d_out + chars);
    float *d_peak = d_psd + _q->nfft;
    AsgramLevels lv;
    for (int i = 0; i < 10; i++) { lv.levels[i] = _q->levels[i]; lv.levelchar[i] = _q->levelchar[i]; }
    asgram_levels_kernel<<<1, 256, 0, sp->stream>>>(d_X, _q->nfft, lv, d_out, d_psd, d_peak);
    LC_LAUNCH_CHECK();
    float peak[2];
    LC_CUDA_CHECK(cudaMemcpyAsync(_ascii, d_out, _q->nfft, cudaMemcpyDeviceToHost, sp->stream));
    LC_CUDA_CHECK(cudaMemcpyAsync(peak, d_peak, sizeof(peak), cudaMemcpyDeviceToHost, sp->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(sp->stream));
    *_peakval = peak[0];
    *_peakfreq = peak[1];
}

void asgramcf_print(asgramcf _q)
{
    std::vector<char> ascii(_q->nfft + 1, '\0');
    float maxval, maxfreq;
    asgramcf_execute(_q, ascii.data(), &maxval, &maxfreq);
    printf(" > %s < pk%5.1f dB [%5.2f]\n", ascii.data(), maxval, maxfreq);
}

}  // extern "C"
