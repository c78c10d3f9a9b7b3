// fftfilt.cu -- fftfilt_crcf: FIR filtering of blocks of n samples through 2n-point transforms, the long-filter
// alternative to firfilt_crcf.  SURVEY.md 8(f) item 3.
//
// Replaces src/filter/src/fftfilt.c (create :68-129, destroy :131-151, reset :154-161, print :164-181,
// set_scale :184-189, execute :195-260, get_length :263-266), instantiated for crcf by src/filter/src/filter_crcf.c.
//
// The reference is overlap-add: pad the block with n zeros, transform, multiply by H = FFT(h | 0), transform back,
// add the tail saved from the previous block, keep the new tail.  The output is the causal linear convolution
// y[i] = scale sum_k h[k] x[i-k] (h_len - 1 <= n), which does not depend on how the blocks are cut.  Here each
// block is computed on its own by overlap-SAVE: transform the 2n samples [(b-1) n, (b+1) n) of the stream, multiply
// by H, transform back and keep the last n outputs (the circular wrap only reaches the first h_len - 1 <= n) -- the
// same one forward + one backward 2n-point transform per block, but no tail passed from block to block, so all
// blocks of a call run in parallel and nothing but the input (twice, the second time from L2) and the output
// touches HBM.  State: the last n input samples.  Results agree with the reference to float32 rounding.
//
// Power-of-two 2n that fits shared memory: one kernel, both transforms in shared memory (the Stockham passes of
// fft_device.cuh; the first pass reads the stream, the product with H is folded into the backward transform's
// first pass, the last pass writes only the kept half, scaled).  Any other n: gather, batched DFT, multiply,
// batched DFT, scatter through scratch buffers.
#include <cmath>
#include <cstdint>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "host_pipeline.cuh"
#include "fft_device.cuh"
#include "quad_fft.cuh"
#include "tile_fft.cuh"

struct fftfilt_crcf_s {
    unsigned int h_len = 0, n = 0;
    std::vector<float> h;
    float scale = 1.0f;              // user scale / (2n), as the reference stores it
    lc::FftPlanDev fwd, bwd;         // 2n points
    float2 *d_H = nullptr;           // [2n]
    float2 *d_hist = nullptr;        // last n input samples, oldest first
    lc::DevBuf stage_in, stage_out, hist_tmp, scratch0, scratch1;
    cudaStream_t stream = nullptr;
};

namespace {

constexpr int kThreads = 256;

// first pass of the forward transforms: element k of transform f of this CTA is stream sample (b0 + f - 1) n + k
struct SegmentSrc {
    const float2 *hist, *x;
    int H;                            // = n
    unsigned int n;
    unsigned long long b0, b_end;
    __device__ __forceinline__ float2 operator()(unsigned int f, unsigned int k) const
    {
        const unsigned long long b = b0 + f;
        if (b >= b_end) return make_float2(0.f, 0.f);
        return lcd::ld_stream(hist, x, (long long)b * n - (long long)n + k, H);
    }
};

// first pass of the backward transforms: the forward result (shared memory) times H
struct ProductSrc {
    const float2 *X;                  // padded shared-memory layout
    const float2 *Hf;                 // [2n], global
    unsigned int log2N;
    __device__ __forceinline__ float2 operator()(unsigned int f, unsigned int k) const
    {
        return lcd::cmul(X[lcfft::pidx((f << log2N) + k)], __ldg(Hf + k));
    }
};

// last pass of the backward transforms: keep elements n .. 2n-1, scaled
struct KeepDst {
    float2 *y;
    unsigned int n;
    float scale;
    unsigned long long b0, b_end;
    __device__ __forceinline__ void operator()(unsigned int f, unsigned int k, float2 v) const
    {
        const unsigned long long b = b0 + f;
        if (k < n || b >= b_end) return;
        y[b * n + (k - n)] = make_float2(v.x * scale, v.y * scale);
    }
};

template <bool TWG>
__global__ void __launch_bounds__(lcfft::kFftThreads)
fftfilt_pow2_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, const float2 *__restrict__ Hf,
                    const float2 *__restrict__ tw_f, const float2 *__restrict__ tw_b, float2 *__restrict__ y,
                    unsigned int n, unsigned int log2N, unsigned int per_cta, unsigned long long n_blocks, float scale)
{
    extern __shared__ float2 smem_f2[];
    const unsigned int N = 2 * n;
    lcfft::Smem sf = lcfft::carve(smem_f2, N, per_cta);                // buf0, buf1, forward table
    float2 *twb = sf.tw + N;                                           // backward table behind it
    if (TWG) {                                                         // long transforms: the plans' tables, through L1
        sf.tw = const_cast<float2 *>(tw_f);
        twb = const_cast<float2 *>(tw_b);
    }
    if (!TWG)
        for (unsigned int i = threadIdx.x; i < N; i += blockDim.x) {
            sf.tw[i] = __ldg(tw_f + i);
            twb[i] = __ldg(tw_b + i);
        }
    const unsigned long long b0 = (unsigned long long)blockIdx.x * per_cta;
    const SegmentSrc src{hist, x, (int)n, n, b0, n_blocks};
    const float2 *X = lcfft::stockham_run_to_smem<-1, TWG>(sf, N, log2N, per_cta, src);
    // the backward transform starts from the buffer that holds X
    float2 *other = (X == sf.buf0) ? sf.buf1 : sf.buf0;
    const lcfft::Smem sb{const_cast<float2 *>(X), other, twb};
    lcfft::stockham_run<+1, TWG>(sb, N, log2N, per_cta, ProductSrc{X, Hf, log2N}, KeepDst{y, n, scale, b0, n_blocks});
}

// ---- 2n = 1024 / 2048 / 4096: the chained quad transforms of quad_fft.cuh (forward, product with H and backward in
// registers and two shared tiles, no block-wide Stockham passes)
struct SegmentSrcT {
    const float2 *hist, *x;
    unsigned int n;
    // transform t covers stream samples (t - 1) n ... (t + 1) n - 1; the n samples before the call live in hist
    __device__ __forceinline__ float2 load(long long t, unsigned int k) const
    {
        const long long idx = (t - 1) * (long long)n + k;
        return idx >= 0 ? __ldg(x + idx) : __ldg(hist + (n + idx));
    }
    __device__ __forceinline__ float2 finish(long long, unsigned int, float2 raw) const { return raw; }
};
struct KeepDstT {
    float2 *y;
    unsigned int n;
    float scale;
    __device__ __forceinline__ void operator()(long long t, unsigned int k, float2 v) const
    {
        __stcs(y + t * n + (k - n), make_float2(v.x * scale, v.y * scale));
    }
};

template <int RA>
__global__ void __launch_bounds__(lcquad::kThreads, 2)
fftfilt_quad_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, const float2 *__restrict__ Hf,
                    const float2 *__restrict__ tw_f, const float2 *__restrict__ tw_b, float2 *__restrict__ y, long long n_blocks,
                    int tiles_per_inst, float scale)
{
    extern __shared__ __align__(16) unsigned char smem_quad[];
    lcquad::run_chain<RA, true>(smem_quad, tw_f, tw_b, Hf, n_blocks, tiles_per_inst, SegmentSrcT{hist, x, 128u * RA},
                                KeepDstT{y, 128u * RA, scale});
}

template <int RA>
void launch_fftfilt_quad(fftfilt_crcf q, const float2 *d_x, unsigned long long n_blocks, float2 *d_y, cudaStream_t st)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(fftfilt_quad_kernel<RA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)lcquad::chain_smem_bytes<RA>()));
    unsigned int grid = 0;
    int tpi = 1;
    lcquad::plan(RA, n_blocks, grid, tpi);
    fftfilt_quad_kernel<RA><<<grid, lcquad::kThreads, lcquad::chain_smem_bytes<RA>(), st>>>(
        q->d_hist, d_x, q->d_H, q->fwd.d_tw_tile, q->bwd.d_tw_tile, d_y, (long long)n_blocks, tpi, q->scale);
}

// ---- 2n = 256 / 512: the chained tile transforms of tile_fft.cuh (the backward transform is the transpose of the forward
// one and starts from the pair of outputs the forward transform leaves in each thread)
template <int M>
__global__ void __launch_bounds__(lctile::kThreads, 2)
fftfilt_tile_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, const float2 *__restrict__ Hf,
                    const float2 *__restrict__ tw_f, const float2 *__restrict__ tw_b, float2 *__restrict__ y, long long n_blocks,
                    int tiles_per_inst, float scale)
{
    extern __shared__ __align__(16) unsigned char smem_tile[];
    lctile::run_chain<M, true>(smem_tile, tw_f, tw_b, Hf, n_blocks, tiles_per_inst, SegmentSrcT{hist, x, (unsigned int)M / 2},
                               KeepDstT{y, (unsigned int)M / 2, scale});
}

template <int M>
void launch_fftfilt_tile(fftfilt_crcf q, const float2 *d_x, unsigned long long n_blocks, float2 *d_y, cudaStream_t st)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(fftfilt_tile_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)lctile::chain_smem_bytes<M>()));
    unsigned int grid = 0;
    int tpi = 1;
    lctile::plan(M, n_blocks, grid, tpi);
    fftfilt_tile_kernel<M><<<grid, lctile::kThreads, lctile::chain_smem_bytes<M>(), st>>>(
        q->d_hist, d_x, q->d_H, q->fwd.d_tw_tile, q->bwd.d_tw_tile, d_y, (long long)n_blocks, tpi, q->scale);
}

// ---- any other n: through scratch ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
fftfilt_gather_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, unsigned int n, float2 *__restrict__ seg,
                      unsigned long long total)
{
    const unsigned int N = 2 * n;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long b = i / N;
        const unsigned int k = (unsigned int)(i % N);
        seg[i] = lcd::ld_stream(hist, x, (long long)b * n - (long long)n + k, (int)n);
    }
}

__global__ void __launch_bounds__(kThreads)
fftfilt_multiply_kernel(float2 *__restrict__ X, const float2 *__restrict__ Hf, unsigned int N, unsigned long long total)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x)
        X[i] = lcd::cmul(X[i], __ldg(Hf + (unsigned int)(i % N)));
}

__global__ void __launch_bounds__(kThreads)
fftfilt_keep_kernel(const float2 *__restrict__ t, unsigned int n, float scale, float2 *__restrict__ y, unsigned long long total)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long b = i / n;
        const unsigned int k = (unsigned int)(i % n);
        const float2 v = t[b * 2 * n + n + k];
        y[i] = make_float2(v.x * scale, v.y * scale);
    }
}

inline unsigned int grid_for(unsigned long long total)
{
    unsigned long long b = (total + kThreads - 1) / kThreads;
    const unsigned long long cap = 148ull * 8 * 16;
    return (unsigned int)(b < cap ? (b ? b : 1) : cap);
}

size_t pow2_smem(unsigned int N, unsigned int per_cta)
{
    const bool staged = N < lcfft::kTwGlobalFrom;                      // forward and backward tables
    return lcfft::smem_for(N, per_cta, staged) + (staged ? sizeof(float2) * N : 0);
}

void blocks_dev(fftfilt_crcf q, const float2 *d_x, unsigned long long n_blocks, float2 *d_y, cudaStream_t st)
{
    if (n_blocks == 0) return;
    const unsigned int n = q->n, N = 2 * n;
    if (lcquad::supported(N) && q->fwd.d_tw_tile && q->bwd.d_tw_tile && n_blocks < (1ull << 31) && !getenv("LIQUID_CUDA_FFTFILT_GENERIC") &&
        !getenv("LIQUID_CUDA_FFTFILT_STOCKHAM") && !getenv("LIQUID_CUDA_NO_FUSED")) {
        if (N == 1024) launch_fftfilt_quad<4>(q, d_x, n_blocks, d_y, st);
        else if (N == 2048) launch_fftfilt_quad<8>(q, d_x, n_blocks, d_y, st);
        else launch_fftfilt_quad<16>(q, d_x, n_blocks, d_y, st);
        LC_LAUNCH_CHECK();
    } else if ((N == 256 || N == 512) && q->fwd.d_tw_tile && q->bwd.d_tw_tile && n_blocks < (1ull << 31) &&
               !getenv("LIQUID_CUDA_FFTFILT_GENERIC") && !getenv("LIQUID_CUDA_FFTFILT_STOCKHAM") && !getenv("LIQUID_CUDA_NO_FUSED")) {
        if (N == 256) launch_fftfilt_tile<256>(q, d_x, n_blocks, d_y, st);
        else launch_fftfilt_tile<512>(q, d_x, n_blocks, d_y, st);
        LC_LAUNCH_CHECK();
    } else if (lcfft::pow2_in_smem(N) && pow2_smem(N, 1) <= 200 * 1024 && !getenv("LIQUID_CUDA_FFTFILT_GENERIC")) {
        static std::atomic<unsigned long long> configured{0};
        if (lc::first_time_on_device(configured)) {
            LC_CUDA_CHECK(cudaFuncSetAttribute(fftfilt_pow2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            LC_CUDA_CHECK(cudaFuncSetAttribute(fftfilt_pow2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
        const unsigned int per_cta = lcfft::per_cta_for(N, n_blocks, 4096);
        const unsigned long long ctas = (n_blocks + per_cta - 1) / per_cta;
        if (ctas >= (1ull << 31)) lc::fatal("fftfilt_crcf_execute_block", "too many blocks for one launch");
        if (N >= lcfft::kTwGlobalFrom)
            fftfilt_pow2_kernel<true><<<(unsigned int)ctas, lcfft::kFftThreads, pow2_smem(N, per_cta), st>>>(
                q->d_hist, d_x, q->d_H, q->fwd.d_twp, q->bwd.d_twp, d_y, n, lc::ilog2(N), per_cta, n_blocks, q->scale);
        else
            fftfilt_pow2_kernel<false><<<(unsigned int)ctas, lcfft::kFftThreads, pow2_smem(N, per_cta), st>>>(
                q->d_hist, d_x, q->d_H, q->fwd.d_twp, q->bwd.d_twp, d_y, n, lc::ilog2(N), per_cta, n_blocks, q->scale);
        LC_LAUNCH_CHECK();
    } else {
        // pieces of at most ~256 MB of scratch per buffer
        unsigned long long lim = (256ull << 20) / (sizeof(float2) * N);
        if (lim == 0) lim = 1;
        const unsigned long long piece = n_blocks < lim ? n_blocks : lim;
        float2 *seg = (float2 *)q->scratch0.get(sizeof(float2) * piece * N);
        float2 *frq = (float2 *)q->scratch1.get(sizeof(float2) * piece * N);
        for (unsigned long long b = 0; b < n_blocks; b += piece) {
            const unsigned long long nb = (n_blocks - b < piece) ? n_blocks - b : piece;
            // blocks b .. b+nb-1 of the stream: the history of the first one is the previous n samples of x (or hist)
            const float2 *xs = d_x + b * n;
            const float2 *hs = b == 0 ? q->d_hist : d_x + (b - 1) * n;
            fftfilt_gather_kernel<<<grid_for(nb * N), kThreads, 0, st>>>(hs, xs, n, seg, nb * N);
            LC_LAUNCH_CHECK();
            q->fwd.run(seg, frq, nb, st);
            fftfilt_multiply_kernel<<<grid_for(nb * N), kThreads, 0, st>>>(frq, q->d_H, N, nb * N);
            LC_LAUNCH_CHECK();
            q->bwd.run(frq, seg, nb, st);
            fftfilt_keep_kernel<<<grid_for(nb * n), kThreads, 0, st>>>(seg, n, q->scale, d_y + b * n, nb * n);
            LC_LAUNCH_CHECK();
        }
    }
    lc::roll_history(q->d_hist, n, q->hist_tmp, d_x, n_blocks * n, st);
}

}  // namespace

extern "C" {

fftfilt_crcf fftfilt_crcf_create(float *_h, unsigned int _h_len, unsigned int _n)
{
    if (_h_len == 0) lc::fatal("fftfilt_crcf_create", "filter length must be greater than zero");
    if (_n < _h_len - 1) lc::fatal("fftfilt_crcf_create", "block length must be greater than _h_len-1");
    if (_n == 0) lc::fatal("fftfilt_crcf_create", "block length must be greater than zero");
    lc::ensure_device();
    fftfilt_crcf q = new fftfilt_crcf_s();
    q->h_len = _h_len;
    q->n = _n;
    q->h.assign(_h, _h + _h_len);
    const unsigned int N = 2 * _n;
    q->fwd.init(N, -1);
    q->bwd.init(N, +1);
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    LC_CUDA_CHECK(cudaMalloc(&q->d_H, sizeof(float2) * N));
    LC_CUDA_CHECK(cudaMalloc(&q->d_hist, sizeof(float2) * _n));
    // H = DFT of the zero-padded taps, computed by the same transform the blocks use (fftfilt.c:112-121)
    std::vector<float2> hp(N, make_float2(0.f, 0.f));
    for (unsigned int i = 0; i < _h_len; i++) hp[i].x = _h[i];
    float2 *d_tmp = (float2 *)q->scratch0.get(sizeof(float2) * N);
    LC_CUDA_CHECK(cudaMemcpyAsync(d_tmp, hp.data(), sizeof(float2) * N, cudaMemcpyHostToDevice, q->stream));
    q->fwd.run(d_tmp, q->d_H, 1, q->stream);
    LC_CUDA_CHECK(cudaStreamSynchronize(q->stream));
    fftfilt_crcf_set_scale(q, 1.0f);
    fftfilt_crcf_reset(q);
    return q;
}

void fftfilt_crcf_destroy(fftfilt_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    _q->fwd.destroy();
    _q->bwd.destroy();
    cudaFree(_q->d_H);
    cudaFree(_q->d_hist);
    _q->stage_in.release();
    _q->stage_out.release();
    _q->hist_tmp.release();
    _q->scratch0.release();
    _q->scratch1.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void fftfilt_crcf_reset(fftfilt_crcf _q)
{
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_hist, 0, sizeof(float2) * _q->n, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

void fftfilt_crcf_print(fftfilt_crcf _q)
{
    printf("fftfilt_crcf: [h_len=%u, n=%u] (cuda)\n", _q->h_len, _q->n);
    for (unsigned int i = 0; i < _q->h_len; i++) printf("  h(%3u) = %12.8f\n", i + 1, _q->h[_q->h_len - i - 1]);
    printf("  scale = %12.8f\n", _q->scale);
}

void fftfilt_crcf_set_scale(fftfilt_crcf _q, float _scale) { _q->scale = _scale / (float)(2 * _q->n); }

unsigned int fftfilt_crcf_get_length(fftfilt_crcf _q) { return _q->h_len; }

// _n_blocks blocks of n device-resident samples; the state is carried in stream order
void fftfilt_crcf_execute_block_dev(fftfilt_crcf _q, const void *_d_x, unsigned long long _n_blocks, void *_d_y, void *_stream)
{
    blocks_dev(_q, (const float2 *)_d_x, _n_blocks, (float2 *)_d_y, lc::as_stream(_stream));
}

void fftfilt_crcf_execute_block(fftfilt_crcf _q, liquid_float_complex *_x, unsigned int _n_blocks, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n_blocks, _q->n, _q->n,
                   [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { blocks_dev(_q, di, nu, dob, st); });
}

void fftfilt_crcf_execute(fftfilt_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    fftfilt_crcf_execute_block(_q, _x, 1, _y);
}

}  // extern "C"
