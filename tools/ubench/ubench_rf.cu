// microbenchmark: FFMA2 (fma.rn.f32x2, scalar tap x sample pair + accumulator pair) issue rate against the
// number of distinct source registers an instruction reads and against operand reuse between neighbours.
//   mode 1  tap varies, sample fixed, 16 accumulators          mode 2  tap fixed, sample varies, 16 accumulators
//   mode 3  tap, sample and accumulator all change every instruction (16 each)
//   mode 4  as 3, neighbours (2k, 2k+1) share the sample       mode 5  as 3, neighbours share the tap
//   mode 6  as 3 with 4 accumulators (the MAC loop's four chains)   mode 7  as 4 with 4 accumulators
//   mode 8  as 3 with 8 accumulators                                mode 9  as 4 with 8 accumulators
// one CTA per SM, 1 or 2 warps per sub-partition; prints cycles per FFMA2 per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 bc(float t) { return make_float2(t, t); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float2 *out, const float *hin, int iters, long long *cyc)
{
    float2 a[16];
    float h[16];
    float2 s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
        h[i] = hin[i] * 1e-3f + threadIdx.x;
        s[i] = make_float2(hin[16 + i] + threadIdx.x, hin[32 + i] - threadIdx.x);
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int r = (i + kk) & 15;
                if (MODE == 1) a[i] = __ffma2_rn(bc(h[r]), s[0], a[i]);
                if (MODE == 2) a[i] = __ffma2_rn(bc(h[0]), s[r], a[i]);
                if (MODE == 3) a[i] = __ffma2_rn(bc(h[r]), s[(i + 5 * kk) & 15], a[i]);
                if (MODE == 4) a[i] = __ffma2_rn(bc(h[r]), s[((i >> 1) + 5 * kk) & 15], a[i]);
                if (MODE == 5) a[i] = __ffma2_rn(bc(h[((i >> 1) + kk) & 15]), s[(i + 5 * kk) & 15], a[i]);
                if (MODE == 6) a[i & 3] = __ffma2_rn(bc(h[r]), s[(i + 5 * kk) & 15], a[i & 3]);
                if (MODE == 7) a[i & 3] = __ffma2_rn(bc(h[r]), s[((i >> 1) + 5 * kk) & 15], a[i & 3]);
                if (MODE == 8) a[i & 7] = __ffma2_rn(bc(h[r]), s[(i + 5 * kk) & 15], a[i & 7]);
                if (MODE == 9) a[i & 7] = __ffma2_rn(bc(h[r]), s[((i >> 1) + 5 * kk) & 15], a[i & 7]);
            }
        }
    }
    const long long t1 = clock64();
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; i++) { r.x += a[i].x + h[i] + s[i].x; r.y += a[i].y + s[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(float2 *out, float *hin, long long *cyc)
{
    const int iters = 400;
    for (int threads : {128, 256}) {
        long long hc;
        k<MODE><<<148, threads>>>(out, hin, iters, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mode %d warps/SMSP %d: %.3f cycles per FFMA2 per SMSP\n", MODE, threads / 128,
               (double)hc / ((double)iters * 256 * (threads / 128.0)));
    }
}

int main()
{
    float2 *out;
    float *hin;
    long long *cyc;
    cudaMalloc(&out, 148 * 256 * 8);
    cudaMalloc(&hin, 128 * 4);
    cudaMalloc(&cyc, 8);
    float hh[128];
    for (int i = 0; i < 128; i++) hh[i] = 0.001f * (i + 1);
    cudaMemcpy(hin, hh, sizeof hh, cudaMemcpyHostToDevice);
    run<1>(out, hin, cyc);
    run<2>(out, hin, cyc);
    run<3>(out, hin, cyc);
    run<4>(out, hin, cyc);
    run<5>(out, hin, cyc);
    run<6>(out, hin, cyc);
    run<7>(out, hin, cyc);
    run<8>(out, hin, cyc);
    run<9>(out, hin, cyc);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
