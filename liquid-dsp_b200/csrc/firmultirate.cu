// firmultirate.cu -- firdecim_crcf and firinterp_crcf: the polyphase FIR pattern without a transform
// (the natural pre-decimation / post-interpolation stages around a channelizer).  SURVEY.md 8(f) item 3.
//
// Replaces src/filter/src/firdecim.c (create :46-82, create_kaiser :88-118, create_prototype :126-163,
// clear :179-182, execute :189-206, execute_block :213-223) and src/filter/src/firinterp.c (create :43-80,
// create_kaiser :86-116, create_prototype :124-158, reset :178-181, execute :187-199, execute_block :205-217),
// both instantiated for crcf by src/filter/src/filter_crcf.c.
//
//   firdecim   y[i] = sum_k h[k] x[i M - k]: the reference pushes M samples per output and takes the dot
//              product right after the FIRST push.  Kernel: a CTA stages tile*M + h_len - 1 input samples in
//              shared memory with coalesced loads (every input sample is read from HBM once), the taps sit in
//              shared memory too (broadcast reads), each thread produces one output; the staged samples are
//              padded by one element per 16 so that the stride-M reads do not pile onto a few banks.
//   firinterp  y[i M + j] = sum_n h[j + n M] x[i - n] (a firpfb: one push, then the M branch outputs).  Kernel:
//              one thread per output; the M outputs of one input share their samples (broadcast within the
//              warp), taps are read coalesced.  The path is bound by its 8 M bytes of output per input sample.
// State, as in the reference: the last h_len (firdecim) / ceil(h_len / M) (firinterp) input samples.
#include <cmath>
#include <cstdint>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "host_pipeline.cuh"

struct firdecim_crcf_s {
    unsigned int M, h_len;
    std::vector<float> h;
    float *d_h = nullptr;
    float2 *d_hist = nullptr;        // last h_len samples, oldest first
    lc::DevBuf stage_in, stage_out, hist_tmp;
    cudaStream_t stream = nullptr;
};

struct firinterp_crcf_s {
    unsigned int M, h_len, sub_len;  // h_len = M * sub_len (zero-padded prototype)
    std::vector<float> h;
    float *d_h = nullptr;
    float2 *d_hist = nullptr;        // last sub_len samples, oldest first
    lc::DevBuf stage_in, stage_out, hist_tmp;
    cudaStream_t stream = nullptr;
};

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ unsigned int pad16(unsigned int i) { return i + (i >> 4); }

// outputs [blockIdx.x * tile, ...): smem = taps[h_len] floats, then the staged samples
__global__ void __launch_bounds__(kThreads)
firdecim_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H, const float *__restrict__ h,
                unsigned int h_len, unsigned int M, unsigned int tile, float2 *__restrict__ y, unsigned long long n_out)
{
    extern __shared__ float smem_f[];
    float *h_s = smem_f;
    float2 *xs = reinterpret_cast<float2 *>(smem_f + ((h_len + 1) & ~1u));
    const unsigned long long o0 = (unsigned long long)blockIdx.x * tile;
    unsigned int nt = tile;
    if (o0 + nt > n_out) nt = (unsigned int)(n_out - o0);
    for (unsigned int k = threadIdx.x; k < h_len; k += blockDim.x) h_s[k] = __ldg(h + k);
    // staged sample q is stream sample o0*M - (h_len-1) + q, q < (nt-1)*M + h_len
    const long long first = (long long)(o0 * M) - (long long)(h_len - 1);
    const unsigned int n_samp = (nt - 1) * M + h_len;
    for (unsigned int q = threadIdx.x; q < n_samp; q += blockDim.x) xs[pad16(q)] = lcd::ld_stream(hist, x, first + q, H);
    __syncthreads();
    for (unsigned int t = threadIdx.x; t < nt; t += blockDim.x) {
        // oldest sample first, as the reference's dot product over its window (firdecim.c:64-67, :200-203)
        float re = 0.f, im = 0.f;
        const unsigned int base = t * M;
        for (unsigned int i = 0; i < h_len; i++) {
            const float2 s = xs[pad16(base + i)];
            const float c = h_s[h_len - 1 - i];
            re = fmaf(c, s.x, re);
            im = fmaf(c, s.y, im);
        }
        y[o0 + t] = make_float2(re, im);
    }
}

// output o = i M + j:  sum_n h[j + n M] x[i - n], oldest sample first (firpfb.c:71-84 stores branches reversed)
__global__ void __launch_bounds__(kThreads)
firinterp_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H, const float *__restrict__ h,
                 unsigned int sub_len, unsigned int M, float2 *__restrict__ y, unsigned long long n_out)
{
    for (unsigned long long o = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; o < n_out;
         o += (unsigned long long)gridDim.x * blockDim.x) {
        const long long i = (long long)(o / M);
        const unsigned int j = (unsigned int)(o % M);
        float re = 0.f, im = 0.f;
        for (unsigned int k = 0; k < sub_len; k++) {
            const unsigned int n = sub_len - 1 - k;                   // tap index: n-th newest sample
            const float2 s = lcd::ld_stream(hist, x, i - (long long)n, H);
            const float c = __ldg(h + j + (size_t)n * M);
            re = fmaf(c, s.x, re);
            im = fmaf(c, s.y, im);
        }
        y[o] = make_float2(re, im);
    }
}

// The throughput form of firinterp: a CTA stages the inputs of its tile (+ sub_len - 1 of history) and the whole
// prototype in shared memory; a work item is (block of 8 consecutive inputs, branch j) and keeps a sliding window
// of 8 samples in registers, so each staged sample is read once per item and feeds 8 packed FFMA2 per tap.  Items
// are laid out branch-fastest: the M outputs of one input are written by adjacent lanes (full sectors).  Oldest
// sample first, as firinterp_kernel.
constexpr int kIR = 8;                                                // inputs per work item

__global__ void __launch_bounds__(kThreads)
firinterp_tile_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H, const float *__restrict__ h,
                      unsigned int sub_len, unsigned int M, unsigned int n_ib, float2 *__restrict__ y, unsigned long long n_in)
{
    extern __shared__ float2 smem_i[];
    const unsigned int T = n_ib * kIR;                                // inputs per CTA
    const unsigned int halo = sub_len - 1;
    float2 *xs = smem_i;                                              // staged sample q = input i0 - halo + q
    float *hs = reinterpret_cast<float *>(xs + ((T + halo + 1) & ~1u));
    const unsigned long long i0 = (unsigned long long)blockIdx.x * T;
    const unsigned int nt = (n_in - i0 < T) ? (unsigned int)(n_in - i0) : T;
    for (unsigned int k = threadIdx.x; k < M * sub_len; k += blockDim.x) hs[k] = __ldg(h + k);
    for (unsigned int q = threadIdx.x; q < nt + halo; q += blockDim.x) {
        const long long idx = (long long)i0 - (long long)halo + q;
        if (idx >= 0) {
            const unsigned int dst = (unsigned int)__cvta_generic_to_shared(xs + q);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(x + idx) : "memory");
        } else {
            xs[q] = __ldg(hist + (H + idx));
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (unsigned int w = threadIdx.x; w < n_ib * M; w += blockDim.x) {
        const unsigned int j = w % M, ib = w / M;
        const unsigned int t0 = ib * kIR;                             // first input of the item, relative to i0
        if (t0 >= nt) continue;
        // at step k (tap n = sub_len-1-k) input t0 + r needs sample t0 + r - n = staged element t0 + r + k
        float2 win[kIR];
#pragma unroll
        for (int r = 0; r < kIR - 1; r++) win[r] = xs[t0 + r];
        float2 acc[kIR];
#pragma unroll
        for (int r = 0; r < kIR; r++) acc[r] = make_float2(0.f, 0.f);
        for (unsigned int k0 = 0; k0 < sub_len; k0 += kIR) {
#pragma unroll
            for (int u = 0; u < kIR; u++) {
                const unsigned int k = k0 + u;
                if (k < sub_len) {
                    win[(u + kIR - 1) % kIR] = xs[t0 + kIR - 1 + k];   // the slot freed by the previous step
                    const float c = hs[j + (sub_len - 1 - k) * M];
                    const float2 cc = make_float2(c, c);
#pragma unroll
                    for (int r = 0; r < kIR; r++) acc[r] = __ffma2_rn(cc, win[(r + u) % kIR], acc[r]);
                }
            }
        }
        float2 *yo = y + (i0 + t0) * M + j;
#pragma unroll
        for (int r = 0; r < kIR; r++)
            if (t0 + r < nt) yo[(size_t)r * M] = acc[r];
    }
}

// The throughput form of firdecim: the filter is split into its M branches h_j[n] = h[n M + j] (zero-padded to
// sub_len M taps), y[o] = sum_j sum_n h_j[n] x[(o - n) M - j].  A CTA stages the inputs of kDR * n_ob outputs (+ the
// sub_len M - 1 samples of history) once; a work item is (block of 8 consecutive outputs, branch j) and slides an
// 8-sample register window over its branch's decimated sequence -- every staged sample is read once per item and
// feeds up to 8 packed FFMA2 per tap -- then the M partial sums of each output are added in branch order through
// shared memory.  The staged array is padded by M elements per 8 M so that the item groups of a warp spread over
// the banks.  (Summation order differs from the oldest-first dot product of firdecim_kernel: float32 rounding.)
constexpr int kDR = 8;                                                // outputs per work item

__global__ void __launch_bounds__(kThreads)
firdecim_tile_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H, const float *__restrict__ h,
                     unsigned int h_len, unsigned int sub_len, unsigned int M, unsigned int n_ob, float2 *__restrict__ y,
                     unsigned long long n_out)
{
    extern __shared__ float2 smem_d[];
    const unsigned int T = n_ob * kDR;                                // outputs per CTA
    const unsigned int rows = n_ob + (sub_len + kDR - 1) / kDR + 1;    // rows of 8 M staged samples (padded to 9 M)
    float2 *xs = smem_d;
    float2 *part = xs + (size_t)rows * (kDR + 1) * M;                  // [T][M + 1] partial sums
    float *hs = reinterpret_cast<float *>(part + (size_t)T * (M + 1)); // [sub_len M] taps, zero beyond h_len
    const unsigned long long o0 = (unsigned long long)blockIdx.x * T;
    const unsigned int nt = (n_out - o0 < T) ? (unsigned int)(n_out - o0) : T;
    for (unsigned int k = threadIdx.x; k < M * sub_len; k += blockDim.x) hs[k] = k < h_len ? __ldg(h + k) : 0.0f;
    // staged sample q is stream sample o0 M - sub_len M + 1 + q, q < (nt - 1 + sub_len) M
    const long long first = (long long)(o0 * M) - (long long)(sub_len * M) + 1;
    const unsigned int n_samp = (nt - 1 + sub_len) * M;
    const unsigned int row_rcp = 0xFFFFFFFFu / (kDR * M) + 1;          // q / (8 M) = umulhi(q, row_rcp) for the small q here
    for (unsigned int q = threadIdx.x; q < n_samp; q += blockDim.x) {
        const unsigned int pq = q + M * __umulhi(q, row_rcp);
        const long long idx = first + q;
        if (idx >= 0) {
            const unsigned int dst = (unsigned int)__cvta_generic_to_shared(xs + pq);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(x + idx) : "memory");
        } else {
            xs[pq] = (idx >= -(long long)H) ? __ldg(hist + (H + idx)) : make_float2(0.f, 0.f);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (unsigned int w = threadIdx.x; w < n_ob * M; w += blockDim.x) {
        const unsigned int j = w % M, ob = w / M;
        const unsigned int t0 = ob * kDR;
        if (t0 >= nt) continue;
        // element u of the branch's decimated sequence (u = 0 is output o0's oldest tap) sits at staged
        // u M + (M - 1 - j), padded row u / 8
        // (t0 is a multiple of 8: the first 7 elements share a row, every 8th step crosses into the next one)
        const float2 *px = xs + (size_t)t0 * M + (M - 1 - j) + (size_t)M * ob;      // element t0
        float2 win[kDR];
#pragma unroll
        for (int r = 0; r < kDR - 1; r++) win[r] = px[(size_t)r * M];
        px += (size_t)(kDR - 1) * M;                                                  // element t0 + 7 (same row)
        const float *ph = hs + (size_t)(sub_len - 1) * M + j;                         // tap sub_len - 1 of the branch
        float2 acc[kDR];
#pragma unroll
        for (int r = 0; r < kDR; r++) acc[r] = make_float2(0.f, 0.f);
        for (unsigned int k0 = 0; k0 < sub_len; k0 += kDR) {
#pragma unroll
            for (int u = 0; u < kDR; u++) {
                if (k0 + u < sub_len) {                               // step k = k0 + u: tap n = sub_len - 1 - k
                    win[(u + kDR - 1) % kDR] = *px;                   // element t0 + 7 + k
                    px += (u == 0) ? 2 * M : M;                       // element t0 + 8 + k0 starts the next row (one pad)
                    const float c = *ph;
                    ph -= M;
                    const float2 cc = make_float2(c, c);
#pragma unroll
                    for (int r = 0; r < kDR; r++) acc[r] = __ffma2_rn(cc, win[(r + u) % kDR], acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kDR; r++) part[(size_t)(t0 + r) * (M + 1) + j] = acc[r];
    }
    __syncthreads();
    for (unsigned int t = threadIdx.x; t < nt; t += blockDim.x) {
        const float2 *p = part + (size_t)t * (M + 1);
        float2 sum = p[0];
        for (unsigned int j = 1; j < M; j++) sum = __fadd2_rn(sum, p[j]);
        y[o0 + t] = sum;
    }
}

inline unsigned int grid_for(unsigned long long total)
{
    unsigned long long b = (total + kThreads - 1) / kThreads;
    const unsigned long long cap = 148ull * 8 * 16;
    return (unsigned int)(b < cap ? (b ? b : 1) : cap);
}

void decim_dev(firdecim_crcf q, const float2 *d_x, unsigned long long n_out, float2 *d_y, cudaStream_t st)
{
    if (n_out == 0) return;
    {
        // branch-decomposed tiled kernel when its staging fits ~96 KB
        const unsigned int sub_len = (q->h_len + q->M - 1) / q->M;
        unsigned int n_ob = (kThreads + q->M - 1) / q->M;
        if (n_ob < 4) n_ob = 4;
        const unsigned int T = n_ob * kDR;
        const size_t rows = n_ob + (sub_len + kDR - 1) / kDR + 1;
        const size_t smem_t = sizeof(float2) * (rows * (kDR + 1) * q->M + (size_t)T * (q->M + 1)) + sizeof(float) * (size_t)q->M * sub_len;
        const unsigned long long tiles = (n_out + T - 1) / T;
        if (smem_t <= 96 * 1024 && n_out >= 4ull * T && tiles < (1ull << 31) && !getenv("LIQUID_CUDA_FIRDECIM_SIMPLE")) {
            static std::atomic<unsigned long long> configured_t{0};
            if (lc::first_time_on_device(configured_t))
                LC_CUDA_CHECK(cudaFuncSetAttribute(firdecim_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            firdecim_tile_kernel<<<(unsigned int)tiles, kThreads, smem_t, st>>>(q->d_hist, d_x, (int)q->h_len, q->d_h, q->h_len,
                                                                             sub_len, q->M, n_ob, d_y, n_out);
            LC_LAUNCH_CHECK();
            lc::roll_history(q->d_hist, q->h_len, q->hist_tmp, d_x, n_out * q->M, st);
            return;
        }
    }
    // outputs per CTA: as many as fit ~96 KB of staged samples, at most 256
    const size_t budget = 96 * 1024;
    const size_t taps_bytes = sizeof(float) * ((q->h_len + 1) & ~1u);
    unsigned int tile = kThreads;
    auto bytes_for = [&](unsigned int t) {
        const size_t n_samp = (size_t)(t - 1) * q->M + q->h_len;
        return taps_bytes + sizeof(float2) * (n_samp + n_samp / 16 + 2);
    };
    while (tile > 1 && bytes_for(tile) > budget) tile >>= 1;
    const size_t smem = bytes_for(tile);
    if (smem > 200 * 1024) lc::fatal("firdecim_crcf_execute_block", "filter too long for this build");
    static std::atomic<unsigned long long> configured{0};
    if (smem > 48 * 1024 && lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(firdecim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const unsigned long long ctas = (n_out + tile - 1) / tile;
    if (ctas >= (1ull << 31)) lc::fatal("firdecim_crcf_execute_block", "too many outputs for one launch");
    firdecim_kernel<<<(unsigned int)ctas, kThreads, smem, st>>>(q->d_hist, d_x, (int)q->h_len, q->d_h, q->h_len, q->M, tile, d_y, n_out);
    LC_LAUNCH_CHECK();
    lc::roll_history(q->d_hist, q->h_len, q->hist_tmp, d_x, n_out * q->M, st);
}

void interp_dev(firinterp_crcf q, const float2 *d_x, unsigned long long n_in, float2 *d_y, cudaStream_t st)
{
    if (n_in == 0) return;
    // items per CTA: at least one per thread, at least 4 input blocks
    unsigned int n_ib = (kThreads + q->M - 1) / q->M;
    if (n_ib < 4) n_ib = 4;
    const unsigned int T = n_ib * kIR;
    const size_t smem = sizeof(float2) * ((T + q->sub_len) & ~1u) + sizeof(float) * (size_t)q->M * q->sub_len;
    const unsigned long long tiles = (n_in + T - 1) / T;
    if (smem <= 48 * 1024 && n_in >= 4ull * T && tiles < (1ull << 31) && !getenv("LIQUID_CUDA_FIRINTERP_SIMPLE")) {
        firinterp_tile_kernel<<<(unsigned int)tiles, kThreads, smem, st>>>(q->d_hist, d_x, (int)q->sub_len, q->d_h, q->sub_len,
                                                                          q->M, n_ib, d_y, n_in);
    } else {
        firinterp_kernel<<<grid_for(n_in * q->M), kThreads, 0, st>>>(q->d_hist, d_x, (int)q->sub_len, q->d_h, q->sub_len, q->M, d_y, n_in * q->M);
    }
    LC_LAUNCH_CHECK();
    lc::roll_history(q->d_hist, q->sub_len, q->hist_tmp, d_x, n_in, st);
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------ firdecim_crcf
firdecim_crcf firdecim_crcf_create(unsigned int _M, float *_h, unsigned int _h_len)
{
    if (_h_len == 0) lc::fatal("decim_crcf_create", "filter length must be greater than zero");
    if (_M == 0) lc::fatal("decim_crcf_create", "decimation factor must be greater than zero");
    lc::ensure_device();
    firdecim_crcf q = new firdecim_crcf_s();
    q->M = _M;
    q->h_len = _h_len;
    q->h.assign(_h, _h + _h_len);
    LC_CUDA_CHECK(cudaMalloc(&q->d_h, sizeof(float) * _h_len));
    LC_CUDA_CHECK(cudaMemcpy(q->d_h, q->h.data(), sizeof(float) * _h_len, cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMalloc(&q->d_hist, sizeof(float2) * _h_len));
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    firdecim_crcf_clear(q);
    return q;
}

firdecim_crcf firdecim_crcf_create_kaiser(unsigned int _M, unsigned int _m, float _As)
{
    if (_M < 2) lc::fatal("decim_crcf_create_kaiser", "decim factor must be greater than 1");
    if (_m == 0) lc::fatal("decim_crcf_create_kaiser", "filter delay must be greater than 0");
    if (_As < 0.0f) lc::fatal("decim_crcf_create_kaiser", "stop-band attenuation must be positive");
    // 2Mm+1 Kaiser taps at fc = 0.5/M; the first 2Mm are used                      firdecim.c:106-117
    const unsigned int h_len = 2 * _M * _m + 1;
    std::vector<float> hf(h_len);
    liquid_firdes_kaiser(h_len, 0.5f / (float)_M, _As, 0.0f, hf.data());
    return firdecim_crcf_create(_M, hf.data(), 2 * _M * _m);
}

firdecim_crcf firdecim_crcf_create_prototype(int _type, unsigned int _M, unsigned int _m, float _beta, float _dt)
{
    if (_M < 2) lc::fatal("decim_crcf_create_prototype", "decimation factor must be greater than 1");
    if (_m == 0) lc::fatal("decim_crcf_create_prototype", "filter delay must be greater than 0");
    if (_beta < 0.0f || _beta > 1.0f) lc::fatal("decim_crcf_create_prototype", "filter excess bandwidth factor must be in [0,1]");
    if (_dt < -1.0f || _dt > 1.0f) lc::fatal("decim_crcf_create_prototype", "filter fractional sample delay must be in [-1,1]");
    const unsigned int h_len = 2 * _M * _m + 1;                    // all 2Mm+1 taps here (firdecim.c:150-162)
    std::vector<float> h(h_len);
    liquid_firdes_prototype(_type, _M, _m, _beta, _dt, h.data());
    return firdecim_crcf_create(_M, h.data(), h_len);
}

void firdecim_crcf_destroy(firdecim_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    cudaFree(_q->d_h);
    cudaFree(_q->d_hist);
    _q->stage_in.release();
    _q->stage_out.release();
    _q->hist_tmp.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void firdecim_crcf_print(firdecim_crcf _q)
{
    printf("firdecim_crcf [%u] : %u taps (cuda)\n", _q->M, _q->h_len);
}

void firdecim_crcf_clear(firdecim_crcf _q)
{
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_hist, 0, sizeof(float2) * _q->h_len, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

// _n outputs from _n * M device-resident inputs; state carried in stream order
void firdecim_crcf_execute_block_dev(firdecim_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    decim_dev(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, lc::as_stream(_stream));
}

void firdecim_crcf_execute_block(firdecim_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x),
                   reinterpret_cast<lc_cf *>(_y), _n, _q->M, 1,
                   [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { decim_dev(_q, di, nu, dob, st); });
}

void firdecim_crcf_execute(firdecim_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    firdecim_crcf_execute_block(_q, _x, 1, _y);
}

// ------------------------------------------------------------------------------------------ firinterp_crcf
firinterp_crcf firinterp_crcf_create(unsigned int _M, float *_h, unsigned int _h_len)
{
    if (_M < 2) lc::fatal("firinterp_crcf_create", "interp factor must be greater than 1");
    if (_h_len < _M) lc::fatal("firinterp_crcf_create", "filter length cannot be less than interp factor");
    lc::ensure_device();
    firinterp_crcf q = new firinterp_crcf_s();
    q->M = _M;
    q->sub_len = 0;
    while (q->M * q->sub_len < _h_len) q->sub_len++;
    q->h_len = q->M * q->sub_len;                                  // prototype padded with zeros (firinterp.c:63-72)
    q->h.assign(q->h_len, 0.0f);
    for (unsigned int i = 0; i < _h_len; i++) q->h[i] = _h[i];
    LC_CUDA_CHECK(cudaMalloc(&q->d_h, sizeof(float) * q->h_len));
    LC_CUDA_CHECK(cudaMemcpy(q->d_h, q->h.data(), sizeof(float) * q->h_len, cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMalloc(&q->d_hist, sizeof(float2) * q->sub_len));
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    firinterp_crcf_reset(q);
    return q;
}

firinterp_crcf firinterp_crcf_create_kaiser(unsigned int _M, unsigned int _m, float _As)
{
    if (_M < 2) lc::fatal("firinterp_crcf_create_kaiser", "interp factor must be greater than 1");
    if (_m == 0) lc::fatal("firinterp_crcf_create_kaiser", "filter delay must be greater than 0");
    if (_As < 0.0f) lc::fatal("firinterp_crcf_create_kaiser", "stop-band attenuation must be positive");
    const unsigned int h_len = 2 * _M * _m + 1;
    std::vector<float> hf(h_len);
    liquid_firdes_kaiser(h_len, 0.5f / (float)_M, _As, 0.0f, hf.data());
    return firinterp_crcf_create(_M, hf.data(), 2 * _M * _m);       // firinterp.c:115
}

firinterp_crcf firinterp_crcf_create_prototype(int _type, unsigned int _k, unsigned int _m, float _beta, float _dt)
{
    if (_k < 2) lc::fatal("firinterp_crcf_create_prototype", "interp factor must be greater than 1");
    if (_m == 0) lc::fatal("firinterp_crcf_create_prototype", "filter delay must be greater than 0");
    if (_beta < 0.0f || _beta > 1.0f) lc::fatal("firinterp_crcf_create_prototype", "filter excess bandwidth factor must be in [0,1]");
    if (_dt < -1.0f || _dt > 1.0f) lc::fatal("firinterp_crcf_create_prototype", "filter fractional sample delay must be in [-1,1]");
    const unsigned int h_len = 2 * _k * _m + 1;
    std::vector<float> h(h_len);
    liquid_firdes_prototype(_type, _k, _m, _beta, _dt, h.data());
    return firinterp_crcf_create(_k, h.data(), h_len);
}

void firinterp_crcf_destroy(firinterp_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    cudaFree(_q->d_h);
    cudaFree(_q->d_hist);
    _q->stage_in.release();
    _q->stage_out.release();
    _q->hist_tmp.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void firinterp_crcf_print(firinterp_crcf _q)
{
    printf("interp():\n");
    printf("    interp  :   %u\n", _q->M);
    printf("    h_len   :   %u\n", _q->h_len);
}

void firinterp_crcf_reset(firinterp_crcf _q)
{
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_hist, 0, sizeof(float2) * _q->sub_len, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

// _n * M outputs from _n device-resident inputs; state carried in stream order
void firinterp_crcf_execute_block_dev(firinterp_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    interp_dev(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, lc::as_stream(_stream));
}

void firinterp_crcf_execute_block(firinterp_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x),
                   reinterpret_cast<lc_cf *>(_y), _n, 1, _q->M,
                   [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { interp_dev(_q, di, nu, dob, st); });
}

void firinterp_crcf_execute(firinterp_crcf _q, liquid_float_complex _x, liquid_float_complex *_y)
{
    firinterp_crcf_execute_block(_q, &_x, 1, _y);
}

}  // extern "C"
