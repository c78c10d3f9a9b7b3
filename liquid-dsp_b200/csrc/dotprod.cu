// dotprod.cu -- dotprod_crcf: y = sum_i h[i] x[i], real taps, complex samples.
//
// Replaces src/dotprod/src/dotprod.c:41-168 and the x86 variant dotprod_crcf.mmx.c:95-287.
// Inside the channelizer/filter kernels the dot product is fused (register FMAs); this object API
// exists for drop-in completeness and for the batched device entry point.
#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "firfilt_kernels.cuh"

struct dotprod_crcf_s {
    unsigned int n;
    std::vector<float> h;
    float *d_h = nullptr;
    lc::DevBuf dx, dy;
    cudaStream_t stream = nullptr;
};

namespace {
void dp_upload(dotprod_crcf q)
{
    if (q->d_h) LC_CUDA_CHECK(cudaFree(q->d_h));
    LC_CUDA_CHECK(cudaMalloc(&q->d_h, sizeof(float) * (q->n ? q->n : 1)));
    LC_CUDA_CHECK(cudaMemcpy(q->d_h, q->h.data(), sizeof(float) * q->n, cudaMemcpyHostToDevice));
}
}  // namespace

extern "C" {

dotprod_crcf dotprod_crcf_create(float *_h, unsigned int _n)
{
    lc::ensure_device();
    dotprod_crcf q = new dotprod_crcf_s();
    q->n = _n;
    q->h.assign(_h, _h + _n);
    dp_upload(q);
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    return q;
}

dotprod_crcf dotprod_crcf_recreate(dotprod_crcf _q, float *_h, unsigned int _n)
{
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    _q->n = _n;
    _q->h.assign(_h, _h + _n);
    dp_upload(_q);
    return _q;
}

void dotprod_crcf_destroy(dotprod_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    cudaFree(_q->d_h);
    _q->dx.release();
    _q->dy.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void dotprod_crcf_print(dotprod_crcf _q)
{
    printf("dotprod_crcf [cuda, %u coefficients]\n", _q->n);
    for (unsigned int i = 0; i < _q->n; i++) printf("  %3u : %12.9f\n", i, _q->h[i]);
}

void dotprod_crcf_execute_batch_dev(dotprod_crcf _q, const void *_d_x, unsigned long long _n_vec,
                                    unsigned long long _stride, void *_d_y, void *_stream)
{
    if (_n_vec == 0) return;
    lc::firfilt::launch_dotprod_batch(_q->d_h, _q->n, (const float2 *)_d_x, _n_vec, _stride, (float2 *)_d_y,
                                      lc::as_stream(_stream));
}

void dotprod_crcf_execute(dotprod_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    float2 *dx = (float2 *)_q->dx.get(sizeof(float2) * (_q->n ? _q->n : 1));
    float2 *dy = (float2 *)_q->dy.get(sizeof(float2));
    LC_CUDA_CHECK(cudaMemcpyAsync(dx, _x, sizeof(float2) * _q->n, cudaMemcpyHostToDevice, _q->stream));
    dotprod_crcf_execute_batch_dev(_q, dx, 1, _q->n, dy, _q->stream);
    LC_CUDA_CHECK(cudaMemcpyAsync(_y, dy, sizeof(float2), cudaMemcpyDeviceToHost, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

void dotprod_crcf_run(float *_h, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    dotprod_crcf q = dotprod_crcf_create(_h, _n);
    dotprod_crcf_execute(q, _x, _y);
    dotprod_crcf_destroy(q);
}

void dotprod_crcf_run4(float *_h, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    dotprod_crcf_run(_h, _x, _n, _y);
}

}  // extern "C"
