// resamp2.cu -- resamp2_crcf (half-band filter, 2-channel analysis / synthesis bank, 2:1 decimator and 1:2
// interpolator) and msresamp2_crcf (their 2^S cascades).  SURVEY.md 8(f) item 3: the window + dot-product
// polyphase pattern with two branches and no transform, the usual rate changer in front of / behind a channelizer.
//
// Replaces src/filter/src/resamp2.c (create :65-123, recreate :130-168, clear :209-215, get_delay :218-221,
// filter_execute :228-264, analyzer_execute :271-290, synthesizer_execute :297-313, decim_execute :320-338,
// interp_execute :345-359) and src/filter/src/msresamp2.c (create :82-170, reset :213-225, get_delay :231-258,
// execute :264-283, interp :294-323, decim :329-354), both instantiated for crcf by src/filter/src/filter_crcf.c.
//
// The reference keeps two windows of 2m samples: w0 (delay branch, read at element m-1) and w1 (filter branch,
// dotted with the 2m odd taps of the 4m+1 tap prototype).  Here the object's state is exactly those two windows
// (device memory, oldest first) plus the toggle of the filter mode, so the five modes can be mixed on one object
// as in the reference.  A block of n units is one kernel: thread u derives the two branch sequences from the
// input on the fly (a[u] joins w0, b[u] joins w1):
//     decim        b[u] = x[2u]          a[u] = x[2u+1]            y[u]    = a[u-m] + F(u)
//     analyzer     b[u] = x[2u]/2        a[u] = x[2u+1]/2          y[2u]   = F(u) + a[u-m], y[2u+1] = F(u) - a[u-m]
//     synthesizer  a[u] = x[2u]+x[2u+1]  b[u] = x[2u]-x[2u+1]      y[2u]   = a[u-m],        y[2u+1] = F(u)
//     interp       a[u] = b[u] = x[u]                               y[2u]   = a[u-m],        y[2u+1] = F(u)
//     with F(u) = sum_{k<2m} hb[k] b[u-2m+1+k]; indices below zero read the windows.
// The filter mode pushes alternately into w0 and w1 (toggle) and filters the other window; in terms of one
// virtual sample stream: y0/y1[i] = (x[i-2m] +/- sum_k hb[k] x[i-1-2(2m-1-k)]) / 2.  A second tiny kernel
// writes the windows as they stand after the block into the other half of a ping-pong pair.
// msresamp2 runs its stages back to back on the device through two scratch buffers; the decimator's final
// 2^-S scaling rides on the last stage's store.
#include <cmath>
#include <cstdint>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "host_pipeline.cuh"

struct resamp2_crcf_s {
    unsigned int m = 0, len = 0;     // len = 2m taps on the filter branch
    float f0 = 0.f, As = 0.f;
    std::vector<float> h;            // 4m+1 tap prototype
    std::vector<float> hb;           // filter branch: odd taps, reversed
    float *d_hb = nullptr;
    float2 *d_win = nullptr;         // [2][2][len]: ping-pong pair of (w0 | w1), oldest first
    int cur = 0;                     // which half holds the state
    unsigned int toggle = 0;
    lc::DevBuf stage_in, stage_out;
    cudaStream_t stream = nullptr;
};

struct msresamp2_crcf_s {
    int interp = 0;
    unsigned int S = 0, M = 1;
    float fc = 0.f, f0 = 0.f, As = 0.f, zeta = 1.f;
    std::vector<unsigned int> m_stage;
    std::vector<float> fc_stage, f0_stage;
    std::vector<resamp2_crcf> hb;
    lc::DevBuf stage_in, stage_out, scratch0, scratch1;
    cudaStream_t stream = nullptr;
};

namespace {

constexpr int kThreads = 256;
enum Mode { kDecim = 0, kAnalyzer = 1, kSynth = 2, kInterp = 3 };

struct Windows {
    const float2 *w0, *w1;           // len samples each, oldest first
    unsigned int len;
};

__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 half2(float2 a) { return make_float2(0.5f * a.x, 0.5f * a.y); }

// element n of the delay-branch sequence (what the reference pushes into w0 at unit n)
template <int MODE>
__device__ __forceinline__ float2 seq_a(const Windows &w, const float2 *__restrict__ x, long long n)
{
    if (n < 0) return __ldg(w.w0 + ((long long)w.len + n));
    if (MODE == kDecim) return __ldg(x + 2 * n + 1);
    if (MODE == kAnalyzer) return half2(__ldg(x + 2 * n + 1));
    if (MODE == kSynth) return add2(__ldg(x + 2 * n), __ldg(x + 2 * n + 1));
    return __ldg(x + n);
}

// element n of the filter-branch sequence (pushed into w1)
template <int MODE>
__device__ __forceinline__ float2 seq_b(const Windows &w, const float2 *__restrict__ x, long long n)
{
    if (n < 0) return __ldg(w.w1 + ((long long)w.len + n));
    if (MODE == kDecim) return __ldg(x + 2 * n);
    if (MODE == kAnalyzer) return half2(__ldg(x + 2 * n));
    if (MODE == kSynth) return sub2(__ldg(x + 2 * n), __ldg(x + 2 * n + 1));
    return __ldg(x + n);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
resamp2_units_kernel(Windows w, const float *__restrict__ hb, unsigned int m, const float2 *__restrict__ x,
                     float2 *__restrict__ y, unsigned long long n, float scale, bool y_aligned16)
{
    extern __shared__ float hs[];
    for (unsigned int k = threadIdx.x; k < w.len; k += blockDim.x) hs[k] = __ldg(hb + k);
    __syncthreads();
    for (unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < n;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        const float2 d = seq_a<MODE>(w, x, (long long)u - (long long)m);
        float re = 0.f, im = 0.f;
        const long long b0 = (long long)u - (long long)w.len + 1;
        for (unsigned int k = 0; k < w.len; k++) {
            const float2 s = seq_b<MODE>(w, x, b0 + k);
            re = fmaf(hs[k], s.x, re);
            im = fmaf(hs[k], s.y, im);
        }
        const float2 f = make_float2(re, im);
        if (MODE == kDecim) {
            const float2 v = add2(d, f);
            y[u] = make_float2(v.x * scale, v.y * scale);
        } else {
            const float2 o0 = (MODE == kAnalyzer) ? add2(f, d) : d;
            const float2 o1 = (MODE == kAnalyzer) ? sub2(f, d) : f;
            if (y_aligned16) {
                reinterpret_cast<float4 *>(y)[u] = make_float4(o0.x, o0.y, o1.x, o1.y);
            } else {
                y[2 * u] = o0;
                y[2 * u + 1] = o1;
            }
        }
    }
}

// the windows after n units, written to the other half of the ping-pong pair
template <int MODE>
__global__ void resamp2_units_roll_kernel(Windows w, const float2 *__restrict__ x, unsigned long long n, float2 *__restrict__ next)
{
    const unsigned int j = threadIdx.x;
    if (j >= w.len) return;
    const long long idx = (long long)n - (long long)w.len + j;
    next[j] = seq_a<MODE>(w, x, idx);
    next[w.len + j] = seq_b<MODE>(w, x, idx);
}

// filter mode: sample idx of the virtual stream in which the samples alternate between the two windows; sample i
// joins window (g + i) & 1
__device__ __forceinline__ float2 filt_val(const Windows &w, const float2 *__restrict__ x, long long idx, unsigned int g)
{
    if (idx >= 0) return __ldg(x + idx);
    const unsigned int c = (unsigned int)(((long long)g + idx) & 1);
    const long long k = (-idx - 1) >> 1;
    return __ldg((c ? w.w1 : w.w0) + ((long long)w.len - 1 - k));
}

__global__ void __launch_bounds__(kThreads)
resamp2_filter_kernel(Windows w, const float *__restrict__ hb, unsigned int m, unsigned int g, const float2 *__restrict__ x,
                      float2 *__restrict__ y0, float2 *__restrict__ y1, unsigned long long n)
{
    extern __shared__ float hs[];
    for (unsigned int k = threadIdx.x; k < w.len; k += blockDim.x) hs[k] = __ldg(hb + k);
    __syncthreads();
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float2 d = filt_val(w, x, (long long)i - 2ll * m, g);
        float re = 0.f, im = 0.f;
        const long long b0 = (long long)i - 1 - 2ll * (w.len - 1);
        for (unsigned int k = 0; k < w.len; k++) {
            const float2 s = filt_val(w, x, b0 + 2ll * k, g);
            re = fmaf(hs[k], s.x, re);
            im = fmaf(hs[k], s.y, im);
        }
        y0[i] = make_float2(0.5f * (d.x + re), 0.5f * (d.y + im));
        y1[i] = make_float2(0.5f * (d.x - re), 0.5f * (d.y - im));
    }
}

__global__ void resamp2_filter_roll_kernel(Windows w, unsigned int g, const float2 *__restrict__ x, unsigned long long n,
                                           float2 *__restrict__ next)
{
    const unsigned int j = threadIdx.x;
    if (j >= 2 * w.len) return;
    const unsigned int c = j / w.len, pos = j % w.len;          // window, element (oldest first)
    const long long last = (long long)n - 1;
    const long long newest = ((((long long)g + last) & 1) == (long long)c) ? last : last - 1;
    next[j] = filt_val(w, x, newest - 2ll * (w.len - 1 - pos), g);
}

inline unsigned int grid_for(unsigned long long total)
{
    unsigned long long b = (total + kThreads - 1) / kThreads;
    const unsigned long long cap = 148ull * 8 * 16;
    return (unsigned int)(b < cap ? (b ? b : 1) : cap);
}

Windows windows_of(resamp2_crcf q)
{
    const float2 *base = q->d_win + (size_t)q->cur * 2 * q->len;
    return Windows{base, base + q->len, q->len};
}

float2 *next_windows(resamp2_crcf q) { return q->d_win + (size_t)(q->cur ^ 1) * 2 * q->len; }

template <int MODE>
void units_dev(resamp2_crcf q, const float2 *d_x, unsigned long long n, float2 *d_y, float scale, cudaStream_t st)
{
    if (n == 0) return;
    const Windows w = windows_of(q);
    resamp2_units_kernel<MODE><<<grid_for(n), kThreads, sizeof(float) * q->len, st>>>(
        w, q->d_hb, q->m, d_x, d_y, n, scale, (reinterpret_cast<uintptr_t>(d_y) & 15) == 0);
    LC_LAUNCH_CHECK();
    resamp2_units_roll_kernel<MODE><<<1, (q->len + 31) / 32 * 32, 0, st>>>(w, d_x, n, next_windows(q));
    LC_LAUNCH_CHECK();
    q->cur ^= 1;
}

void filter_dev(resamp2_crcf q, const float2 *d_x, unsigned long long n, float2 *d_y0, float2 *d_y1, cudaStream_t st)
{
    if (n == 0) return;
    const Windows w = windows_of(q);
    resamp2_filter_kernel<<<grid_for(n), kThreads, sizeof(float) * q->len, st>>>(w, q->d_hb, q->m, q->toggle, d_x, d_y0, d_y1, n);
    LC_LAUNCH_CHECK();
    resamp2_filter_roll_kernel<<<1, (2 * q->len + 31) / 32 * 32, 0, st>>>(w, q->toggle, d_x, n, next_windows(q));
    LC_LAUNCH_CHECK();
    q->cur ^= 1;
    q->toggle = (unsigned int)((q->toggle + n) & 1);
}

void design(resamp2_crcf q)
{
    // 4m+1 taps sinc(t/2) kaiser(t) cos(2 pi f0 t) in float, the phase in double as in resamp2.c:94-108;
    // the filter branch keeps the odd taps, reversed (:110-113)
    const unsigned int n = 4 * q->m + 1;
    q->h.resize(n);
    const float beta = lc::kaiser_beta_As(q->As);
    for (unsigned int i = 0; i < n; i++) {
        const float t = (float)i - (float)(n - 1) / 2.0f;
        const float a = lc::sinc_value(t / 2.0f);
        const float b = lc::kaiser_window(i, n, beta, 0);
        const float c = cosf((float)(2.0 * M_PI * (double)t * (double)q->f0));
        q->h[i] = a * b * c;
    }
    q->hb.resize(q->len);
    unsigned int j = 0;
    for (unsigned int i = 1; i < n; i += 2) q->hb[j++] = q->h[n - i - 1];
}

void check_args(const char *fn, unsigned int m, float f0)
{
    if (m < 2) lc::fatal(fn, "filter semi-length must be at least 2");
    if (f0 < -0.5f || f0 > 0.5f) lc::fatal(fn, "f0 must be in (-1,1)");
}

void allocate(resamp2_crcf q)
{
    LC_CUDA_CHECK(cudaMalloc(&q->d_hb, sizeof(float) * q->len));
    LC_CUDA_CHECK(cudaMemcpy(q->d_hb, q->hb.data(), sizeof(float) * q->len, cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMalloc(&q->d_win, sizeof(float2) * 4 * q->len));
}

// small host-pointer calls (the reference's per-sample entry points): stage, run, copy back
template <typename Run>
void host_small(resamp2_crcf q, const liquid_float_complex *x, size_t n_in, liquid_float_complex *y, size_t n_out, Run run)
{
    float2 *din = (float2 *)q->stage_in.get(sizeof(float2) * n_in);
    float2 *dout = (float2 *)q->stage_out.get(sizeof(float2) * n_out);
    LC_CUDA_CHECK(cudaMemcpyAsync(din, x, sizeof(float2) * n_in, cudaMemcpyHostToDevice, q->stream));
    run(din, dout, q->stream);
    LC_CUDA_CHECK(cudaMemcpyAsync(y, dout, sizeof(float2) * n_out, cudaMemcpyDeviceToHost, q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(q->stream));
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------ resamp2_crcf
resamp2_crcf resamp2_crcf_create(unsigned int _m, float _f0, float _As)
{
    check_args("resamp2_crcf_create", _m, _f0);
    lc::ensure_device();
    resamp2_crcf q = new resamp2_crcf_s();
    q->m = _m;
    q->len = 2 * _m;
    q->f0 = _f0;
    q->As = _As;
    design(q);
    allocate(q);
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    resamp2_crcf_clear(q);
    return q;
}

resamp2_crcf resamp2_crcf_recreate(resamp2_crcf _q, unsigned int _m, float _f0, float _As)
{
    // a new length rebuilds the object; otherwise only the taps change and the windows are kept (resamp2.c:130-168)
    if (_m != _q->m) {
        resamp2_crcf_destroy(_q);
        return resamp2_crcf_create(_m, _f0, _As);
    }
    check_args("resamp2_crcf_recreate", _m, _f0);
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    _q->f0 = _f0;
    _q->As = _As;
    design(_q);
    LC_CUDA_CHECK(cudaMemcpy(_q->d_hb, _q->hb.data(), sizeof(float) * _q->len, cudaMemcpyHostToDevice));
    return _q;
}

void resamp2_crcf_destroy(resamp2_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    cudaFree(_q->d_hb);
    cudaFree(_q->d_win);
    _q->stage_in.release();
    _q->stage_out.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void resamp2_crcf_print(resamp2_crcf _q)
{
    printf("fir half-band resampler: [%u taps, f0=%12.8f] (cuda)\n", 4 * _q->m + 1, _q->f0);
    for (unsigned int i = 0; i < 4 * _q->m + 1; i++) printf("  h(%4u) = %12.8f;\n", i + 1, _q->h[i]);
    printf("---\n");
    for (unsigned int i = 0; i < _q->len; i++) printf("  h1(%4u) = %12.8f;\n", i + 1, _q->hb[i]);
}

void resamp2_crcf_clear(resamp2_crcf _q)
{
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_win, 0, sizeof(float2) * 4 * _q->len, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    _q->cur = 0;
    _q->toggle = 0;
}

unsigned int resamp2_crcf_get_delay(resamp2_crcf _q) { return 2 * _q->m - 1; }

void resamp2_crcf_get_taps(resamp2_crcf _q, float *_h, float *_h1)
{
    if (_h) std::copy(_q->h.begin(), _q->h.end(), _h);
    if (_h1) std::copy(_q->hb.begin(), _q->hb.end(), _h1);
}

// device-pointer block forms; the state is carried in stream order
void resamp2_crcf_filter_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y0,
                                           void *_d_y1, void *_stream)
{
    filter_dev(_q, (const float2 *)_d_x, _n, (float2 *)_d_y0, (float2 *)_d_y1, lc::as_stream(_stream));
}

void resamp2_crcf_analyzer_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    units_dev<kAnalyzer>(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, 1.0f, lc::as_stream(_stream));
}

void resamp2_crcf_synthesizer_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    units_dev<kSynth>(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, 1.0f, lc::as_stream(_stream));
}

void resamp2_crcf_decim_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    units_dev<kDecim>(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, 1.0f, lc::as_stream(_stream));
}

void resamp2_crcf_interp_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    units_dev<kInterp>(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, 1.0f, lc::as_stream(_stream));
}

// host-pointer block forms: _n units (filter: samples; analyzer / synthesizer: pairs in and out; decim: outputs;
// interp: inputs)
void resamp2_crcf_filter_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n,
                                       liquid_float_complex *_y0, liquid_float_complex *_y1)
{
    const size_t chunk = 1u << 21;
    for (size_t i0 = 0; i0 < _n; i0 += chunk) {
        const size_t nc = std::min(chunk, (size_t)_n - i0);
        float2 *din = (float2 *)_q->stage_in.get(sizeof(float2) * nc);
        float2 *dout = (float2 *)_q->stage_out.get(sizeof(float2) * 2 * nc);
        LC_CUDA_CHECK(cudaMemcpyAsync(din, _x + i0, sizeof(float2) * nc, cudaMemcpyHostToDevice, _q->stream));
        filter_dev(_q, din, nc, dout, dout + nc, _q->stream);
        LC_CUDA_CHECK(cudaMemcpyAsync(_y0 + i0, dout, sizeof(float2) * nc, cudaMemcpyDeviceToHost, _q->stream));
        LC_CUDA_CHECK(cudaMemcpyAsync(_y1 + i0, dout + nc, sizeof(float2) * nc, cudaMemcpyDeviceToHost, _q->stream));
        LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    }
}

void resamp2_crcf_analyzer_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, 2, 2, [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { units_dev<kAnalyzer>(_q, di, nu, dob, 1.0f, st); });
}

void resamp2_crcf_synthesizer_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, 2, 2, [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { units_dev<kSynth>(_q, di, nu, dob, 1.0f, st); });
}

void resamp2_crcf_decim_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, 2, 1, [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { units_dev<kDecim>(_q, di, nu, dob, 1.0f, st); });
}

void resamp2_crcf_interp_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, 1, 2, [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { units_dev<kInterp>(_q, di, nu, dob, 1.0f, st); });
}

// the reference's per-call entry points
void resamp2_crcf_filter_execute(resamp2_crcf _q, liquid_float_complex _x, liquid_float_complex *_y0, liquid_float_complex *_y1)
{
    liquid_float_complex y[2];
    host_small(_q, &_x, 1, y, 2, [&](float2 *di, float2 *dob, cudaStream_t st) { filter_dev(_q, di, 1, dob, dob + 1, st); });
    *_y0 = y[0];
    *_y1 = y[1];
}

void resamp2_crcf_analyzer_execute(resamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    liquid_float_complex y[2];
    host_small(_q, _x, 2, y, 2, [&](float2 *di, float2 *dob, cudaStream_t st) { units_dev<kAnalyzer>(_q, di, 1, dob, 1.0f, st); });
    _y[0] = y[0];
    _y[1] = y[1];
}

void resamp2_crcf_synthesizer_execute(resamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    liquid_float_complex y[2];
    host_small(_q, _x, 2, y, 2, [&](float2 *di, float2 *dob, cudaStream_t st) { units_dev<kSynth>(_q, di, 1, dob, 1.0f, st); });
    _y[0] = y[0];
    _y[1] = y[1];
}

void resamp2_crcf_decim_execute(resamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    host_small(_q, _x, 2, _y, 1, [&](float2 *di, float2 *dob, cudaStream_t st) { units_dev<kDecim>(_q, di, 1, dob, 1.0f, st); });
}

void resamp2_crcf_interp_execute(resamp2_crcf _q, liquid_float_complex _x, liquid_float_complex *_y)
{
    liquid_float_complex y[2];
    host_small(_q, &_x, 1, y, 2, [&](float2 *di, float2 *dob, cudaStream_t st) { units_dev<kInterp>(_q, di, 1, dob, 1.0f, st); });
    _y[0] = y[0];
    _y[1] = y[1];
}

// ------------------------------------------------------------------------------------------ msresamp2_crcf
msresamp2_crcf msresamp2_crcf_create(int _type, unsigned int _num_stages, float _fc, float _f0, float _As)
{
    if (_num_stages > 16) lc::fatal("msresamp2_crcf_create", "number of stages should not exceed 16");
    if (_fc <= 0.0f || _fc >= 0.5f) lc::fatal("msresamp2_crcf_create", "cut-off frequency must be in (0,0.5)");
    if (_fc > 0.45f) {
        fprintf(stderr, "warning: msresamp2_crcf_create(), cut-off frequency greater than 0.45\n");
        fprintf(stderr, "    >> truncating to 0.45\n");
        _fc = 0.45f;
    }
    if (_f0 != 0.) {
        fprintf(stderr, "warning: msresamp2_crcf_create(), non-zero center frequency not yet supported\n");
        _f0 = 0.;
    }
    lc::ensure_device();
    msresamp2_crcf q = new msresamp2_crcf_s();
    q->interp = (_type == LIQUID_RESAMP_INTERP);
    q->S = _num_stages;
    q->M = 1u << _num_stages;
    q->fc = _fc;
    q->f0 = _f0;
    q->As = _As;
    q->zeta = 1.0f / (float)q->M;
    // stage i: cut-off fc / 2^(i+1), transition band (0.5 - fc_i) / 2, semi-length from Kaiser's estimate, at
    // least 3 (msresamp2.c:128-146)
    float fc = _fc, f0 = _f0;
    for (unsigned int i = 0; i < q->S; i++) {
        f0 = 0.5f * f0;
        fc = 0.5f * fc;
        const float ft = (0.5f - fc) / 2.0f;
        const unsigned int h_len = lc::estimate_req_filter_len(ft, _As);
        const unsigned int m = (unsigned int)ceilf((float)(h_len - 1) / 4.0f);
        q->fc_stage.push_back(fc);
        q->f0_stage.push_back(f0);
        q->m_stage.push_back(m < 3 ? 3 : m);
    }
    for (unsigned int i = 0; i < q->S; i++) q->hb.push_back(resamp2_crcf_create(q->m_stage[i], q->f0_stage[i], _As));
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    return q;
}

void msresamp2_crcf_destroy(msresamp2_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    for (resamp2_crcf s : _q->hb) resamp2_crcf_destroy(s);
    _q->stage_in.release();
    _q->stage_out.release();
    _q->scratch0.release();
    _q->scratch1.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void msresamp2_crcf_print(msresamp2_crcf _q)
{
    printf("multi-stage half-band resampler (cuda):\n");
    printf("    type                    : %s\n", _q->interp ? "interpolator" : "decimator");
    printf("    number of stages        : %u stage%s\n", _q->S, _q->S == 1 ? "" : "s");
    printf("    cut-off frequency, fc   : %12.8f Fs\n", _q->fc);
    printf("    center frequency, f0    : %12.8f Fs\n", _q->f0);
    printf("    stop-band attenuation   : %.2f dB\n", _q->As);
    for (unsigned int i = 0; i < _q->S; i++)
        printf("    stage[%2u]  {m=%3u, As=%6.2f dB, fc=%6.3f, f0=%6.3f}\n", i, _q->m_stage[i], _q->As, _q->fc_stage[i], _q->f0_stage[i]);
}

void msresamp2_crcf_reset(msresamp2_crcf _q)
{
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    for (resamp2_crcf s : _q->hb) resamp2_crcf_clear(s);
}

float msresamp2_crcf_get_delay(msresamp2_crcf _q)
{
    float delay = 0;
    for (unsigned int i = 0; i < _q->S; i++) {
        if (_q->interp) {
            delay *= 0.5f;
            delay += _q->m_stage[i];
        } else {
            delay *= 2;
            delay += 2 * _q->m_stage[_q->S - i - 1] - 1;
        }
    }
    return delay;
}

unsigned int msresamp2_crcf_get_stage_semilength(msresamp2_crcf _q, unsigned int _i) { return _q->m_stage[_i]; }

// _n units: interpolator _n inputs -> _n 2^S outputs (stages last to first, msresamp2.c:294-323); decimator
// _n 2^S inputs -> _n outputs, scaled by 2^-S (stages first to last, :329-354)
void msresamp2_crcf_execute_block_dev(msresamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    cudaStream_t st = lc::as_stream(_stream);
    if (_n == 0) return;
    if (_q->S == 0) {
        LC_CUDA_CHECK(cudaMemcpyAsync(_d_y, _d_x, sizeof(float2) * _n, cudaMemcpyDeviceToDevice, st));
        return;
    }
    const size_t half = (size_t)_n * (_q->M / 2);                  // the largest intermediate stream
    float2 *buf[2] = {nullptr, nullptr};
    if (_q->S > 1) {
        buf[0] = (float2 *)_q->scratch0.get(sizeof(float2) * half);
        buf[1] = (float2 *)_q->scratch1.get(sizeof(float2) * half);
    }
    const float2 *in = (const float2 *)_d_x;
    if (_q->interp) {
        unsigned long long cnt = _n;
        for (unsigned int s = 0; s < _q->S; s++) {
            float2 *out = (s == _q->S - 1) ? (float2 *)_d_y : buf[s & 1];
            units_dev<kInterp>(_q->hb[_q->S - s - 1], in, cnt, out, 1.0f, st);
            cnt *= 2;
            in = out;
        }
    } else {
        unsigned long long cnt = _n * _q->M;
        for (unsigned int s = 0; s < _q->S; s++) {
            cnt /= 2;
            const bool last = (s == _q->S - 1);
            float2 *out = last ? (float2 *)_d_y : buf[s & 1];
            units_dev<kDecim>(_q->hb[s], in, cnt, out, last ? _q->zeta : 1.0f, st);
            in = out;
        }
    }
}

void msresamp2_crcf_execute_block(msresamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    const size_t in_per = _q->interp ? 1 : _q->M, out_per = _q->interp ? _q->M : 1;
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, in_per, out_per,
                   [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { msresamp2_crcf_execute_block_dev(_q, di, nu, dob, st); });
}

void msresamp2_crcf_execute(msresamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    msresamp2_crcf_execute_block(_q, _x, 1, _y);
}

}  // extern "C"
