// microbenchmark: does the FFMA2 issue rate depend on where the operands come from?
//   mode 0  a[i] = a[i] * b + c            (two loop-invariant operand pairs: the operand reuse cache serves them)
//   mode 1  a[i] = bc(h[k]) * s + a[i]     (the FIR / polyphase MAC pattern: a different scalar tap register for
//                                           every instruction, one sample pair shared by 16 accumulators)
//   mode 2  as mode 1 with the sample pair changing every instruction too (rotating over 8 pairs)
// one CTA per SM, 128..1024 threads; prints cycles per FFMA2 per SM sub-partition (2.0 = nominal peak)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 bc(float t) { return make_float2(t, t); }

template <int MODE>
__global__ void k(float *out, const float *hin, int iters, long long *cyc)
{
    float2 a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float h[64];
#pragma unroll
    for (int j = 0; j < 64; j++) h[j] = hin[j] * 1e-3f;
    float2 s[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = make_float2(hin[64 + j], hin[72 + j]);
    const float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 64; kk++) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (MODE == 0) a[i] = __ffma2_rn(a[i], b, c);
                if (MODE == 1) a[i] = __ffma2_rn(bc(h[(kk + i) & 63]), s[kk & 7], a[i]);
                if (MODE == 2) a[i] = __ffma2_rn(bc(h[(kk + i) & 63]), s[(kk + i) & 7], a[i]);
            }
        }
    }
    const long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    float *out, *hin;
    long long *cyc, hc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&hin, 128 * 4);
    cudaMalloc(&cyc, 8);
    float hh[128];
    for (int i = 0; i < 128; i++) hh[i] = 0.001f * (i + 1);
    cudaMemcpy(hin, hh, sizeof hh, cudaMemcpyHostToDevice);
    const int iters = 200;
    for (int threads : {128, 256, 512}) {
        for (int mode = 0; mode < 3; mode++) {
            if (mode == 0) k<0><<<148, threads>>>(out, hin, iters, cyc);
            if (mode == 1) k<1><<<148, threads>>>(out, hin, iters, cyc);
            if (mode == 2) k<2><<<148, threads>>>(out, hin, iters, cyc);
            cudaDeviceSynchronize();
            cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
            const double ops_per_warp = (double)iters * 64 * 16, warps_per_smsp = threads / 128.0;
            printf("threads %4d mode %d: %.3f cycles per FFMA2 per SMSP\n", threads, mode, (double)hc / (ops_per_warp * warps_per_smsp));
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
