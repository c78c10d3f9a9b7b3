// microbenchmark: does an FFMA2 (fma.rn.f32x2) leave the second cycle of its FMA-pipe occupancy free for
// another instruction of the same sub-partition (ALU op / shared-memory load), or does it hold the
// dispatch port for both cycles?  One CTA per SM, 1 / 2 / 4 warps per sub-partition.
//   K = independent ALU (LOP3/IADD3) or LDS instructions issued per FFMA2
// prints cycles per FFMA2 per sub-partition: 2.0 for every K <= 1 means the shadow slot is usable.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 bc(float t) { return make_float2(t, t); }

template <int KIND, int K2>   // KIND 0: ALU filler, 1: LDS.64 filler, 2: scalar FFMA pairs + ALU filler ; K2 = fillers per 2 FFMA2
__global__ void k(float *out, const float *hin, int iters, long long *cyc)
{
    __shared__ float2 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float2(hin[i & 63], hin[(i + 7) & 63]);
    float2 a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float h[16];
#pragma unroll
    for (int j = 0; j < 16; j++) h[j] = hin[j] * 1e-3f;
    float2 s[4];
#pragma unroll
    for (int j = 0; j < 4; j++) s[j] = make_float2(hin[64 + j], hin[72 + j]);
    unsigned int x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = threadIdx.x * 2654435761u + j;
    const unsigned int saddr = (unsigned int)__cvta_generic_to_shared(sm) + (threadIdx.x & 255) * 8;
    float2 ld[8];
#pragma unroll
    for (int j = 0; j < 8; j++) ld[j] = make_float2(0.f, 0.f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                if (KIND == 2) {
                    a[i].x = fmaf(h[(kk + i) & 15], s[kk & 3].x, a[i].x);
                    a[i].y = fmaf(h[(kk + i) & 15], s[kk & 3].y, a[i].y);
                    a[i + 1].x = fmaf(h[(kk + i + 1) & 15], s[kk & 3].x, a[i + 1].x);
                    a[i + 1].y = fmaf(h[(kk + i + 1) & 15], s[kk & 3].y, a[i + 1].y);
                } else {
                    a[i] = __ffma2_rn(bc(h[(kk + i) & 15]), s[kk & 3], a[i]);
                    a[i + 1] = __ffma2_rn(bc(h[(kk + i + 1) & 15]), s[kk & 3], a[i + 1]);
                }
#pragma unroll
                for (int f = 0; f < K2; f++) {
                    const int j = (i / 2 * K2 + f) & 7;
                    if (KIND == 1) {
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(ld[j].x), "=f"(ld[j].y) : "r"(saddr + 2048u * (unsigned)(j & 3)));
                    } else {
                        asm volatile("lop3.b32 %0, %0, %1, 0x9e3779b9, 0x96;" : "+r"(x[j]) : "r"(x[(j + 3) & 7]));
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) r += a[i].x + a[i].y;
#pragma unroll
    for (int j = 0; j < 8; j++) r += (float)x[j] + ld[j].x + ld[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND, int K2>
void run(const char *name, float *out, float *hin, long long *cyc)
{
    const int iters = 400;
    for (int threads : {128, 256, 512}) {
        long long hc;
        k<KIND, K2><<<148, threads>>>(out, hin, iters, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        const double ops_per_warp = (double)iters * 8 * 16, wps = threads / 128.0;
        printf("%-34s fillers per FFMA2 %.1f  warps/SMSP %d: %.3f cycles per FFMA2(-equivalent) per SMSP\n", name, K2 / 2.0,
               threads / 128, (double)hc / (ops_per_warp * wps));
    }
}

int main()
{
    float *out, *hin;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&hin, 128 * 4);
    cudaMalloc(&cyc, 8);
    float hh[128];
    for (int i = 0; i < 128; i++) hh[i] = 0.001f * (i + 1);
    cudaMemcpy(hin, hh, sizeof hh, cudaMemcpyHostToDevice);
    run<0, 0>("FFMA2 only", out, hin, cyc);
    run<0, 1>("FFMA2 + LOP3", out, hin, cyc);
    run<0, 2>("FFMA2 + LOP3", out, hin, cyc);
    run<0, 4>("FFMA2 + LOP3", out, hin, cyc);
    run<1, 1>("FFMA2 + LDS.64", out, hin, cyc);
    run<1, 2>("FFMA2 + LDS.64", out, hin, cyc);
    run<2, 0>("2xFFMA only", out, hin, cyc);
    run<2, 2>("2xFFMA + LOP3", out, hin, cyc);
    run<2, 4>("2xFFMA + LOP3", out, hin, cyc);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
