// design.cu -- host-side prototype filter design in float32, create-time only.
//
// Follows the reference's formulas including their float/double promotion, so that the taps
// uploaded to the GPU equal the reference's to the last bit or two:
//   kaiser_beta_As          src/filter/src/firdes.c:224-236
//   liquid_firdes_kaiser    src/filter/src/firdes.c:244-280
//   kaiser()                src/math/src/math.c:289-310        (r = 2t/N, N not N-1)
//   sincf()                 src/math/src/math.c:128-139        (3-cosine product below 0.01)
//   liquid_besseli0f()      src/math/src/math.bessel.c:86-105  (32-term series through logs)
//   liquid_lngammaf()       src/math/src/math.gamma.c:43-71    (Stirling-type approximation)
//   liquid_firdes_rrcos     src/filter/src/rrcos.c:37-95
// Promotions are written out explicitly because C++ <cmath> overloads would otherwise keep
// expressions such as log(z) in float where the C reference evaluates them in double.
#include <cmath>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"

namespace {

const double kPi = 3.14159265358979323846;

float lngammaf_(float z)
{
    if (z < 10.0f) return lngammaf_(z + 1.0f) - logf(z);
    float g = (float)(0.5 * ((double)logf((float)(2 * kPi)) - log((double)z)));
    g += z * (logf(z + (1 / (12.0f * z - 0.1f / z))) - 1);
    return g;
}

float besseli0f_(float z)
{
    if (z == 0.0f) return 1.0f;
    float y = 0.0f;
    for (unsigned int k = 0; k < 32; k++) {
        float t = (float)k * logf(0.5f * z) - lngammaf_((float)k + 1.0f);
        y += expf(2 * t);
    }
    return y;
}

float sincf_(float x)
{
    if (fabsf(x) < 0.01f)
        return cosf((float)(kPi * (double)x / 2.0)) * cosf((float)(kPi * (double)x / 4.0)) *
               cosf((float)(kPi * (double)x / 8.0));
    return (float)((double)sinf((float)(kPi * (double)x)) / (kPi * (double)x));
}

float kaiser_(unsigned int n, unsigned int N, float beta, float mu)
{
    float t = (float)n - (float)(N - 1) / 2 + mu;
    float r = 2.0f * t / (float)N;
    float a = besseli0f_(beta * sqrtf(1 - r * r));
    float b = besseli0f_(beta);
    return a / b;
}

}  // namespace

namespace lc {

float kaiser_beta_As(float As)
{
    As = fabsf(As);
    float beta;
    if (As > 50.0f)      beta = 0.1102f * (As - 8.7f);
    else if (As > 21.0f) beta = (float)(0.5842 * (double)powf(As - 21, 0.4f) + (double)(0.07886f * (As - 21)));
    else                 beta = 0.0f;
    return beta;
}

int firdes_kaiser(unsigned int n, float fc, float As, float mu, float *h)
{
    if (mu < -0.5f || mu > 0.5f) return 1;
    if (fc < 0.0f || fc > 0.5f) return 2;
    if (n == 0) return 3;
    float beta = kaiser_beta_As(As);
    for (unsigned int i = 0; i < n; i++) {
        float t = (float)i - (float)(n - 1) / 2 + mu;
        float h1 = sincf_(2.0f * fc * t);
        float h2 = kaiser_(i, n, beta, mu);
        h[i] = h1 * h2;
    }
    return 0;
}

int firdes_rrcos(unsigned int k, unsigned int m, float beta, float dt, float *h)
{
    if (k < 1) return 1;
    if (m < 1) return 2;
    if (beta < 0.0f || beta > 1.0f) return 3;
    unsigned int h_len = 2 * k * m + 1;
    const float T = 1.0f;
    for (unsigned int n = 0; n < h_len; n++) {
        float nf = (float)n, kf = (float)k, mf = (float)m;
        float z = (nf + dt) / kf - mf;
        float t1 = cosf((float)((double)(1 + beta) * kPi * (double)z));
        float t2 = sinf((float)((double)(1 - beta) * kPi * (double)z));
        if ((double)fabsf(z) < 1e-5) {
            h[n] = (float)((double)(1 - beta) + (double)(4 * beta) / kPi);
        } else {
            float t3 = 1 / ((4 * beta * z));
            float g = 1 - 16 * beta * beta * z * z;
            g *= g;
            if ((double)g < 1e-5) {
                float g1 = (float)(1 + (double)2.0f / kPi);
                float g2 = sinf((float)((double)0.25f * kPi / (double)beta));
                float g3 = (float)(1 - (double)2.0f / kPi);
                float g4 = cosf((float)((double)0.25f * kPi / (double)beta));
                h[n] = beta / sqrtf(2.0f) * (g1 * g2 + g3 * g4);
            } else {
                float t4 = (float)((double)(4 * beta) /
                                   (kPi * (double)sqrtf(T) * (double)(1 - (16 * beta * beta * z * z))));
                h[n] = t4 * (t1 + (t2 * t3));
            }
        }
    }
    return 0;
}

}  // namespace lc

// ---- C ABI: liquid.h:1476, :1548, :1580 ----
extern "C" float kaiser_beta_As(float _As) { return lc::kaiser_beta_As(_As); }

extern "C" void liquid_firdes_kaiser(unsigned int _n, float _fc, float _As, float _mu, float *_h)
{
    switch (lc::firdes_kaiser(_n, _fc, _As, _mu, _h)) {
    case 1: fprintf(stderr, "error: liquid_firdes_kaiser(), _mu (%12.4e) out of range [-0.5,0.5]\n", _mu); exit(1);
    case 2: fprintf(stderr, "error: liquid_firdes_kaiser(), cutoff frequency (%12.4e) out of range (0, 0.5)\n", _fc); exit(1);
    case 3: fprintf(stderr, "error: liquid_firdes_kaiser(), filter length must be greater than zero\n"); exit(1);
    default: break;
    }
}

extern "C" void liquid_firdes_rrcos(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h)
{
    switch (lc::firdes_rrcos(_k, _m, _beta, _dt, _h)) {
    case 1: fprintf(stderr, "error: liquid_firdes_rrcos(): k must be greater than 0\n"); exit(1);
    case 2: fprintf(stderr, "error: liquid_firdes_rrcos(): m must be greater than 0\n"); exit(1);
    case 3: fprintf(stderr, "error: liquid_firdes_rrcos(): beta must be in [0,1]\n"); exit(1);
    default: break;
    }
}
