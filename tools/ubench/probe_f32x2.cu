#include <cuda_runtime.h>
__global__ void bc(float2 *p, const float *h)
{
    float2 a = p[threadIdx.x], c = p[threadIdx.x + 64];
    float t = h[threadIdx.x];
    // broadcast scalar tap
    float2 r = __ffma2_rn(make_float2(t, t), a, c);
    // swapped operand
    float2 r2 = __ffma2_rn(make_float2(a.y, a.x), c, r);
    // swapped + negated low
    float2 r3 = __ffma2_rn(make_float2(-a.y, a.x), c, r2);
    p[threadIdx.x + 128] = r3;
}
