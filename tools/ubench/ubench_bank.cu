// microbenchmark: does the FFMA2 issue rate of the MAC pattern (scalar tap x sample pair + accumulator pair, all three
// changing from one instruction to the next) depend on WHERE the sample pair and the accumulator pair sit in the register
// file?  Sample pairs and accumulator pairs are created by 128-bit shared loads, so each float4 is an aligned register
// quad R4k .. R4k+3: .xy is the low half (registers = 0, 1 mod 4), .zw the high half (2, 3 mod 4).
//   mode 0  accumulator and sample always in OPPOSITE halves      mode 1  always in the SAME half
//   mode 2  mixed (half of the instructions same, half opposite)
// prints cycles per FFMA2 per sub-partition (2.0 = pipe peak); check the SASS for the register numbers actually used.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float t, float2 s, float2 a)
{
    unsigned long long A = reinterpret_cast<unsigned long long &>(a), S = reinterpret_cast<unsigned long long &>(s), D;
    asm("{\n\t.reg .b64 tt;\n\tmov.b64 tt, {%1, %1};\n\tfma.rn.f32x2 %0, tt, %2, %3;\n\t}" : "=l"(D) : "f"(t), "l"(S), "l"(A));
    return reinterpret_cast<float2 &>(D);
}
#define LO(q) make_float2((q).x, (q).y)
#define HI(q) make_float2((q).z, (q).w)
#define SETLO(q, v) { (q).x = (v).x; (q).y = (v).y; }
#define SETHI(q, v) { (q).z = (v).x; (q).w = (v).y; }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float4 *out, const float *hin, int iters, long long *cyc)
{
    __shared__ float4 sm[16];
    if (threadIdx.x < 16) sm[threadIdx.x] = make_float4(hin[threadIdx.x], hin[threadIdx.x + 16], hin[threadIdx.x + 32], hin[threadIdx.x + 48]);
    __syncthreads();
    float4 s[8], a[4];
    float h[16];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = sm[(i + threadIdx.x) & 15];      // LDS.128: aligned quads
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = sm[(i + 3 + threadIdx.x) & 15];
#pragma unroll
    for (int i = 0; i < 16; i++) h[i] = hin[64 + i] * 1e-3f + threadIdx.x * 1e-6f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                // accumulator pair i: quad i >> 1, half i & 1; sample quad j, half chosen by the mode
                const int j = (i + 3 * kk) & 7;
                const int ah = i & 1;
                const int sh = MODE == 0 ? 1 - ah : MODE == 1 ? ah : ((kk & 1) ? ah : 1 - ah);
                const float t = h[(i + kk) & 15];
                const float2 sv = sh ? HI(s[j]) : LO(s[j]);
                if (ah) { const float2 r = ffma2(t, sv, HI(a[i >> 1])); SETHI(a[i >> 1], r); }
                else    { const float2 r = ffma2(t, sv, LO(a[i >> 1])); SETLO(a[i >> 1], r); }
            }
        }
    }
    const long long t1 = clock64();
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; i++) { r.x += a[i].x; r.y += a[i].y; r.z += a[i].z; r.w += a[i].w; }
#pragma unroll
    for (int i = 0; i < 8; i++) { r.x += s[i].x; r.y += s[i].y; r.z += s[i].z; r.w += s[i].w; }
#pragma unroll
    for (int i = 0; i < 16; i++) r.x += h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(float4 *out, float *hin, long long *cyc)
{
    const int iters = 2000;
    for (int threads : {128, 256}) {
        long long hc;
        k<MODE><<<148, threads>>>(out, hin, iters, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mode %d warps/SMSP %d: %.3f cycles per FFMA2 per SMSP\n", MODE, threads / 128, (double)hc / ((double)iters * 128 * (threads / 128.0)));
    }
}

int main()
{
    float4 *out;
    float *hin;
    long long *cyc;
    cudaMalloc(&out, 148 * 256 * 16);
    cudaMalloc(&hin, 128 * 4);
    cudaMalloc(&cyc, 8);
    float hh[128];
    for (int i = 0; i < 128; i++) hh[i] = 0.001f * (i + 1);
    cudaMemcpy(hin, hh, sizeof hh, cudaMemcpyHostToDevice);
    run<0>(out, hin, cyc);
    run<1>(out, hin, cyc);
    run<2>(out, hin, cyc);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
