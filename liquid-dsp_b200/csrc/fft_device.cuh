// fft_device.cuh -- the shared-memory Stockham FFT used by the batched FFT kernel (fft.cu) and by the
// shape-generic fused analyzers (pfb_kernels.cu), and the post-scaling of a transform's output.
#pragma once
#include "lc_common.cuh"

namespace lcfft {

constexpr int kFftThreads = 256;

// post-processing of the last stage: 0 none, 1 divide by a, 2 multiply by a then by b
__device__ __forceinline__ float2 fft_post(float2 v, int mode, float a, float b)
{
    if (mode == 1) { v.x = v.x / a; v.y = v.y / a; }
    else if (mode == 2) { v.x *= a; v.y *= a; v.x *= b; v.y *= b; }
    return v;
}

// multiply by s*j  (s = +1 or -1)
__device__ __forceinline__ float2 mul_sj(float2 v, float s) { return make_float2(-s * v.y, s * v.x); }

// Stockham autosort FFT of `per_cta` transforms of n = 2^log2n points each, in shared memory: reads
// buf0, ping-pongs with buf1, returns the buffer that holds the result (natural order).  The caller
// has filled buf0 and synchronised; every thread of the CTA must call this.
__device__ __forceinline__ float2 *stockham_pow2(float2 *buf0, float2 *buf1, const float2 *__restrict__ tw,
                                                 unsigned int n, unsigned int log2n, unsigned int per_cta, float sgn)
{
    float2 *in = buf0, *out = buf1;
    unsigned int Ns = 1;
    unsigned int stages4 = log2n / 2;
    for (unsigned int s = 0; s < stages4; s++, Ns <<= 2) {
        const unsigned int q = n >> 2;                 // butterflies per transform
        for (unsigned int i = threadIdx.x; i < per_cta * q; i += blockDim.x) {
            unsigned int f = i / q, j = i % q;
            const float2 *src = in + (size_t)f * n;
            float2 *dst = out + (size_t)f * n;
            unsigned int k = j & (Ns - 1);
            unsigned int tstep = k * (n / (Ns << 2)); // index of W_n^(k n/(4Ns)) = W_{4Ns}^k
            float2 v0 = src[j];
            float2 v1 = lcd::cmul(src[j + q], tw[tstep]);
            float2 v2 = lcd::cmul(src[j + 2 * q], tw[2 * tstep]);
            float2 v3 = lcd::cmul(src[j + 3 * q], tw[3 * tstep]);
            float2 a02 = lcd::cadd(v0, v2), s02 = lcd::csub(v0, v2);
            float2 a13 = lcd::cadd(v1, v3), s13 = mul_sj(lcd::csub(v1, v3), sgn);
            unsigned int d = ((j - k) << 2) + k;
            dst[d]          = lcd::cadd(a02, a13);
            dst[d + Ns]     = lcd::cadd(s02, s13);
            dst[d + 2 * Ns] = lcd::csub(a02, a13);
            dst[d + 3 * Ns] = lcd::csub(s02, s13);
        }
        __syncthreads();
        float2 *t = in; in = out; out = t;
    }
    if (log2n & 1) {
        const unsigned int q = n >> 1;
        for (unsigned int i = threadIdx.x; i < per_cta * q; i += blockDim.x) {
            unsigned int f = i / q, j = i % q;
            const float2 *src = in + (size_t)f * n;
            float2 *dst = out + (size_t)f * n;
            unsigned int k = j & (Ns - 1);
            float2 v0 = src[j];
            float2 v1 = lcd::cmul(src[j + q], tw[k * (n / (Ns << 1))]);
            unsigned int d = ((j - k) << 1) + k;
            dst[d]      = lcd::cadd(v0, v1);
            dst[d + Ns] = lcd::csub(v0, v1);
        }
        __syncthreads();
        float2 *t = in; in = out; out = t;
    }
    return in;
}

// transforms per CTA for n-point transforms (aim at 2048 points per CTA) and the shared memory they need
inline unsigned int per_cta_for(unsigned int n, unsigned long long batch)
{
    unsigned int per_cta = 1;
    while (per_cta * n < 2048 && per_cta < batch) per_cta <<= 1;
    return per_cta;
}
inline size_t smem_for(unsigned int n, unsigned int per_cta) { return (size_t)2 * per_cta * n * sizeof(float2); }
inline bool pow2_in_smem(unsigned int n) { return lc::is_pow2(n) && n >= 2 && (size_t)n * 16 <= 200 * 1024; }

}  // namespace lcfft
