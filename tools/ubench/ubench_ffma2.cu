// microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) vs FADD2 on sm_100a, one CTA per SM
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, long long *cyc)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
                if (MODE == 1) { a[i] = __ffma2_rn(a[i], b, c); }
                if (MODE == 2) { a[i] = __fadd2_rn(a[i], c); }
                if (MODE == 3) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int mode = 0; mode < 4; mode++) {
            if (mode == 0) k<0><<<148, threads>>>(out, iters, cyc);
            if (mode == 1) k<1><<<148, threads>>>(out, iters, cyc);
            if (mode == 2) k<2><<<148, threads>>>(out, iters, cyc);
            if (mode == 3) k<3><<<148, threads>>>(out, iters, cyc);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            // complex-pair ops per thread = iters*32 ; warps per SMSP = threads/128
            double pair_ops_per_warp = (double)iters * 32;
            double warps_per_smsp = threads / 128.0;
            printf("threads %4d mode %d (%s): %lld cycles, %.3f cycles per pair-op per SMSP (=> %.2f scalar-flop-pairs/cycle/SMSP)\n",
                   threads, mode, mode == 0 ? "2xFFMA" : mode == 1 ? "FFMA2" : mode == 2 ? "FADD2" : "2xFADD", h,
                   (double)h / (pair_ops_per_warp * warps_per_smsp), (pair_ops_per_warp * warps_per_smsp) / (double)h);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
