// firfilt.cu -- firfilt_crcf: direct-form FIR filter, complex samples, real taps.
//
// Replaces src/filter/src/firfilt.c (create :62-101, create_kaiser :109-131, create_rnyquist :140-172,
// create_rect :175-197, recreate :203-238, reset :253-263, set_scale :288, push :297-317,
// execute :322-338, execute_block :346-358, freqresponse :370-388, groupdelay :394-405).
//   y[i] = scale * sum_{k<n} h[k] x[i-k],  zero state after reset; x == y allowed in execute_block.
// State = the last n samples pushed.
#include <cmath>
#include <complex>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "host_pipeline.cuh"
#include "firfilt_kernels.cuh"

struct firfilt_crcf_s {
    unsigned int n;
    std::vector<float> h;          // taps in the caller's order h[0..n)
    float scale = 1.0f;
    float *d_taps = nullptr;       // h[0..n) on the device
    float2 *d_hist = nullptr;      // last n samples pushed, oldest first
    unsigned int *d_tile_counter = nullptr;   // work counter of the persistent short-filter kernel
    lc::DevBuf stage_in, stage_out, hist_tmp, one;
    cudaStream_t stream = nullptr;
};

namespace {

void upload_taps(firfilt_crcf q)
{
    if (q->d_taps) LC_CUDA_CHECK(cudaFree(q->d_taps));
    LC_CUDA_CHECK(cudaMalloc(&q->d_taps, sizeof(float) * q->n));
    LC_CUDA_CHECK(cudaMemcpy(q->d_taps, q->h.data(), sizeof(float) * q->n, cudaMemcpyHostToDevice));
}

}  // namespace

extern "C" {

firfilt_crcf firfilt_crcf_create(float *_h, unsigned int _n)
{
    if (_n == 0) lc::fatal("firfilt_crcf_create", "filter length must be greater than zero");
    lc::ensure_device();
    firfilt_crcf q = new firfilt_crcf_s();
    q->n = _n;
    q->h.assign(_h, _h + _n);
    upload_taps(q);
    LC_CUDA_CHECK(cudaMalloc(&q->d_hist, sizeof(float2) * _n));
    LC_CUDA_CHECK(cudaMalloc(&q->d_tile_counter, sizeof(unsigned int)));
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    firfilt_crcf_reset(q);
    return q;
}

firfilt_crcf firfilt_crcf_create_kaiser(unsigned int _n, float _fc, float _As, float _mu)
{
    if (_n == 0) lc::fatal("firfilt_crcf_create_kaiser", "filter length must be greater than zero");
    std::vector<float> h(_n);
    liquid_firdes_kaiser(_n, _fc, _As, _mu, h.data());
    return firfilt_crcf_create(h.data(), _n);
}

firfilt_crcf firfilt_crcf_create_rnyquist(int _type, unsigned int _k, unsigned int _m, float _beta, float _mu)
{
    if (_k < 2) lc::fatal("firfilt_crcf_create_rnyquist", "filter samples/symbol must be greater than 1");
    if (_m == 0) lc::fatal("firfilt_crcf_create_rnyquist", "filter delay must be greater than 0");
    if (_beta < 0.0f || _beta > 1.0f)
        lc::fatal("firfilt_crcf_create_rnyquist", "filter excess bandwidth factor must be in [0,1]");
    unsigned int h_len = 2 * _k * _m + 1;
    std::vector<float> h(h_len);
    liquid_firdes_prototype(_type, _k, _m, _beta, _mu, h.data());   // firfilt.c:161
    return firfilt_crcf_create(h.data(), h_len);
}

firfilt_crcf firfilt_crcf_create_rect(unsigned int _n)
{
    if (_n == 0 || _n > 1024) lc::fatal("firfilt_crcf_create_rect", "filter length must be in [1,1024]");
    std::vector<float> h(_n, 1.0f);
    return firfilt_crcf_create(h.data(), _n);
}

// new taps; the buffer contents survive only when the length is unchanged (firfilt.c:203-238)
firfilt_crcf firfilt_crcf_recreate(firfilt_crcf _q, float *_h, unsigned int _n)
{
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    if (_n != _q->n) {
        _q->n = _n;
        LC_CUDA_CHECK(cudaFree(_q->d_hist));
        LC_CUDA_CHECK(cudaMalloc(&_q->d_hist, sizeof(float2) * _n));
        LC_CUDA_CHECK(cudaMemset(_q->d_hist, 0, sizeof(float2) * _n));
    }
    _q->h.assign(_h, _h + _n);
    upload_taps(_q);
    return _q;
}

void firfilt_crcf_destroy(firfilt_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    cudaFree(_q->d_taps);
    cudaFree(_q->d_hist);
    cudaFree(_q->d_tile_counter);
    _q->stage_in.release();
    _q->stage_out.release();
    _q->hist_tmp.release();
    _q->one.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void firfilt_crcf_reset(firfilt_crcf _q)
{
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_hist, 0, sizeof(float2) * _q->n, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

void firfilt_crcf_print(firfilt_crcf _q)
{
    printf("firfilt_%s:\n", "crcf");
    for (unsigned int i = 0; i < _q->n; i++) printf("  h(%3u) = %12.8f\n", i + 1, _q->h[i]);
    printf("  scale = %12.8f\n", _q->scale);
}

void firfilt_crcf_set_scale(firfilt_crcf _q, float _scale) { _q->scale = _scale; }

unsigned int firfilt_crcf_get_length(firfilt_crcf _q) { return _q->n; }

void firfilt_crcf_execute_block_dev(firfilt_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    if (_n == 0) return;
    cudaStream_t st = lc::as_stream(_stream);
    const float2 *d_x = (const float2 *)_d_x;
    // the kernels are tiled: a tile reads the tail of its left neighbour's input, so the device form
    // cannot run in place (the host form firfilt_crcf_execute_block stages through separate device
    // buffers and keeps the reference's x == y allowance, firfilt.c:340-341)
    if (_d_x == _d_y) lc::fatal("firfilt_crcf_execute_block_dev", "in-place operation is not supported on device pointers");
    lc::firfilt::launch(_q->d_hist, d_x, (int)_q->n, _q->d_taps, _q->n, _q->scale, (float2 *)_d_y, _n, 1, st, _q->d_tile_counter);
    lc::roll_history(_q->d_hist, _q->n, _q->hist_tmp, d_x, _n, st);
}

void firfilt_crcf_execute_streams_dev(firfilt_crcf _q, const void *_d_x, unsigned int _n_streams,
                                      unsigned long long _n, void *_d_y, void *_stream)
{
    if (_n == 0 || _n_streams == 0) return;
    if (_d_x == _d_y) lc::fatal("firfilt_crcf_execute_streams_dev", "in-place operation is not supported for the multi-stream form");
    lc::firfilt::launch(nullptr, (const float2 *)_d_x, 0, _q->d_taps, _q->n, _q->scale, (float2 *)_d_y, _n,
                        _n_streams, lc::as_stream(_stream), _q->d_tile_counter);
}

void firfilt_crcf_execute_block(firfilt_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    // the staging buffers are distinct on the device, so host-side aliasing (x == y) is harmless
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x),
                   reinterpret_cast<lc_cf *>(_y), _n, 1, 1,
                   [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) {
                       firfilt_crcf_execute_block_dev(_q, di, nu, dob, st);
                   });
}

void firfilt_crcf_push(firfilt_crcf _q, liquid_float_complex _x)
{
    float2 *d_one = (float2 *)_q->one.get(sizeof(float2) * _q->n * 2);
    LC_CUDA_CHECK(cudaMemcpyAsync(d_one, &_x, sizeof(float2), cudaMemcpyHostToDevice, _q->stream));
    lc::roll_history(_q->d_hist, _q->n, _q->hist_tmp, d_one, 1, _q->stream);
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

// output for the buffer as it stands: y = scale * sum_k h[k] * (k-th newest sample)
void firfilt_crcf_execute(firfilt_crcf _q, liquid_float_complex *_y)
{
    float2 *d_y = (float2 *)_q->stage_out.get(sizeof(float2));
    lc::firfilt::launch_state_output(_q->d_hist, _q->d_taps, _q->n, _q->scale, d_y, _q->stream);
    LC_CUDA_CHECK(cudaMemcpyAsync(_y, d_y, sizeof(float2), cudaMemcpyDeviceToHost, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

// host-side analysis helpers (no sample data involved); firfilt.c:370-405
void firfilt_crcf_freqresponse(firfilt_crcf _q, float _fc, liquid_float_complex *_H)
{
    // the reference sums its reversed tap array against e^{j 2 pi fc i}: H = sum_i h[n-1-i] e^{j 2 pi fc i}
    std::complex<float> H = 0.0f;
    for (unsigned int i = 0; i < _q->n; i++) {
        std::complex<double> a(0.0, 2 * M_PI * (double)_fc * (double)i);
        std::complex<float> e = std::exp(std::complex<float>((float)a.real(), (float)a.imag()));
        H += _q->h[_q->n - 1 - i] * e;
    }
    H *= _q->scale;
    lc_cf out = {H.real(), H.imag()};
    memcpy(_H, &out, sizeof out);
}

float firfilt_crcf_groupdelay(firfilt_crcf _q, float _fc)
{
    // fir_group_delay (src/filter/src/group_delay.c): Re{ sum h[i] i e^{j2pi fc i} / sum h[i] e^{j2pi fc i} }
    std::complex<float> t0 = 0.0f, t1 = 0.0f;
    for (unsigned int i = 0; i < _q->n; i++) {
        std::complex<float> e = std::exp(std::complex<float>(0.0f, (float)(2 * M_PI * (double)_fc * (double)i)));
        t0 += _q->h[i] * e * (float)i;
        t1 += _q->h[i] * e;
    }
    return (t0 / t1).real();
}

}  // extern "C"
