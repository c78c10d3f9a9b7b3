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
//   liquid_firdes_arkaiser / _rkaiser, rkaiser_approximate_rho   src/filter/src/rkaiser.c:47-148, :155-210, :339-516
//   estimate_req_filter_As, liquid_filter_isi                    src/filter/src/firdes.c:80-116, :503-531
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

float kaiser_window(unsigned int n, unsigned int N, float beta, float mu) { return kaiser_(n, N, beta, mu); }
float sinc_value(float x) { return sincf_(x); }

// Kaiser's length estimate (the reference build's method 0)                           firdes.c:52-75, :163-176
unsigned int estimate_req_filter_len(float df, float As) { return (unsigned int)((As - 7.95f) / (14.26f * df)); }

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

// ---- root-Nyquist Kaiser prototypes (rkaiser.c) -------------------------------------------------

float estimate_req_filter_As(float df, unsigned int N)
{
    // bisection on Kaiser's length estimate N(As) = (As - 7.95) / (14.26 df), 20 steps    firdes.c:80-116
    float lo = 0.01f, hi = 200.0f, As = 0.0f;
    for (int it = 0; it < 20; it++) {
        As = 0.5f * (hi + lo);
        const float n_hat = (As - 7.95f) / (14.26f * df);
        if (n_hat < (float)N) lo = As;
        else                  hi = As;
    }
    return As;
}

void filter_isi(const float *h, unsigned int k, unsigned int m, float *rms, float *peak)
{
    // normalised auto-correlation at the symbol lags k, 2k, ..., 2mk       firdes.c:420-438, :503-531
    const unsigned int n = 2 * k * m + 1;
    auto autocorr = [&](unsigned int lag) {
        float r = 0.0f;
        for (unsigned int i = lag; i < n; i++) r += h[i] * h[i - lag];
        return r;
    };
    const float r0 = autocorr(0);
    float acc = 0.0f, mx = 0.0f;
    for (unsigned int i = 1; i <= 2 * m; i++) {
        const float e = fabsf(autocorr(i * k) / r0);
        acc += e * e;
        if (i == 1 || e > mx) mx = e;
    }
    *rms = sqrtf(acc / (float)(2 * m));
    *peak = mx;
}

float rkaiser_approximate_rho(unsigned int m, float beta)
{
    // rho ~ c0 + c1 ln(beta) + c2 ln^2(beta), coefficients tabulated per m              rkaiser.c:155-210
    static const float tab[22][3] = {
        {0.75749731f, 0.06134303f, -0.08729663f}, {0.81151861f, 0.07437658f, -0.01427088f},
        {0.84249538f, 0.07684185f, -0.00536879f}, {0.86140782f, 0.07144126f, -0.00558652f},
        {0.87457740f, 0.06578694f, -0.00650447f}, {0.88438797f, 0.06074265f, -0.00736405f},
        {0.89216620f, 0.05669236f, -0.00791222f}, {0.89874983f, 0.05361696f, -0.00815301f},
        {0.90460032f, 0.05167952f, -0.00807893f}, {0.91034430f, 0.05130753f, -0.00746192f},
        {0.91587675f, 0.05180436f, -0.00670711f}, {0.92121875f, 0.05273801f, -0.00588351f},
        {0.92638195f, 0.05400764f, -0.00508452f}, {0.93123555f, 0.05516163f, -0.00437306f},
        {0.93564993f, 0.05596561f, -0.00388152f}, {0.93976742f, 0.05662274f, -0.00348280f},
        {0.94351703f, 0.05694120f, -0.00318821f}, {0.94557273f, 0.05227591f, -0.00400676f},
        {0.95001614f, 0.05681641f, -0.00300628f}, {0.95281708f, 0.05637607f, -0.00304790f},
        {0.95536256f, 0.05575880f, -0.00312988f}, {0.95754206f, 0.05426060f, -0.00385945f}};
    float c0, c1, c2;
    if (m >= 1 && m <= 22) {
        c0 = tab[m - 1][0];
        c1 = tab[m - 1][1];
        c2 = tab[m - 1][2];
    } else {
        c0 = (float)(0.056873 * (double)logf((float)m + 1e-3f) + 0.781388);
        c1 = 0.05426f;
        c2 = -0.00386f;
    }
    const float b = logf(beta);
    float rho = c0 + c1 * b + c2 * b * b;
    if (rho < 0.0f) rho = 0.0f;
    else if (rho > 1.0f) rho = 1.0f;
    return rho;
}

namespace {

// one Kaiser design for a given bandwidth adjustment rho; returns its RMS ISI          rkaiser.c:482-516
float rkaiser_design(unsigned int k, unsigned int m, float beta, float dt, float rho, float *h)
{
    const unsigned int n = 2 * k * m + 1;
    const float kf = (float)k;
    const float del = beta * rho / kf;                               // transition bandwidth
    const float As = estimate_req_filter_As(del, n);
    const float fc = 0.5f * (1 + beta * (1.0f - rho)) / kf;          // cut-off
    firdes_kaiser(n, fc, As, dt, h);
    float rms, peak;
    filter_isi(h, k, m, &rms, &peak);
    return rms;
}

void scale_to_energy(float *h, unsigned int n, unsigned int k)      // rkaiser.c:144-147
{
    float e2 = 0.0f;
    for (unsigned int i = 0; i < n; i++) e2 += h[i] * h[i];
    for (unsigned int i = 0; i < n; i++) h[i] *= sqrtf((float)k / e2);
}

int rnyquist_args(unsigned int k, unsigned int m, float beta, float dt)
{
    if (k < 2) return 1;
    if (m < 1) return 2;
    if (beta <= 0.0f || beta >= 1.0f) return 3;
    if (dt < -1.0f || dt > 1.0f) return 4;
    return 0;
}

}  // namespace

int firdes_arkaiser(unsigned int k, unsigned int m, float beta, float dt, float *h)
{
    if (int rc = rnyquist_args(k, m, beta, dt)) return rc;
    // closed-form fit of rho (double literals: evaluated in double as in the C reference) rkaiser.c:113-128
    const float c0 = (float)(0.762886 + 0.067663 * (double)logf((float)m));
    const float c1 = (float)0.065515;
    const float c2 = logf((float)(1 - 0.088 * (double)powf((float)m, (float)-1.6)));
    const float lb = logf(beta);
    float rho = c0 + c1 * lb + c2 * lb * lb;
    if (rho <= 0.0f || rho >= 1.0f) rho = rkaiser_approximate_rho(m, beta);
    const unsigned int n = 2 * k * m + 1;
    const float kf = (float)k;
    const float del = beta * rho / kf;
    const float As = estimate_req_filter_As(del, n);
    const float fc = 0.5f * (1 + beta * (1.0f - rho)) / kf;
    firdes_kaiser(n, fc, As, dt, h);
    scale_to_energy(h, n, k);
    return 0;
}

int firdes_rkaiser(unsigned int k, unsigned int m, float beta, float dt, float *h, float *rho_out)
{
    if (int rc = rnyquist_args(k, m, beta, dt)) return rc;
    // bracketed parabolic search for the rho with the least RMS ISI                     rkaiser.c:339-474
    const unsigned int n = 2 * k * m + 1;
    float best_rho = rkaiser_approximate_rho(m, beta), best_isi = 0.0f;
    float x1 = best_rho, dx = 0.2f;
    for (unsigned int round = 0; round < 14; round++) {
        float x0 = x1 - dx, x2 = x1 + dx;
        if (x0 <= 0.0f) x0 = 0.01f;
        if (x2 >= 1.0f) x2 = 0.99f;
        const float y0 = rkaiser_design(k, m, beta, dt, x0, h);
        const float y1 = rkaiser_design(k, m, beta, dt, x1, h);
        const float y2 = rkaiser_design(k, m, beta, dt, x2, h);
        if (round == 0 || y1 < best_isi) {
            best_rho = x1;
            best_isi = y1;
        }
        // vertex of the parabola through (x0,y0), (x1,y1), (x2,y2); the sums are float, the quotient double
        const double ta = y0 * (x1 * x1 - x2 * x2) + y1 * (x2 * x2 - x0 * x0) + y2 * (x0 * x0 - x1 * x1);
        const double tb = y0 * (x1 - x2) + y1 * (x2 - x0) + y2 * (x0 - x1);
        const float x_hat = (float)(0.5f * ta / tb);
        if (x_hat < x0 || x_hat > x2) break;                         // minimum left the bracket
        if (round > 3 && fabsf(x_hat - x1) < 1e-6f) break;           // converged
        x1 = x_hat;
        dx *= 0.5f;
    }
    rkaiser_design(k, m, beta, dt, best_rho, h);
    scale_to_energy(h, n, k);
    if (rho_out) *rho_out = best_rho;
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

namespace {
void rnyquist_fatal(const char *fn, int rc)
{
    switch (rc) {
    case 1: fprintf(stderr, "error: %s(), k must be at least 2\n", fn); exit(1);
    case 2: fprintf(stderr, "error: %s(), m must be at least 1\n", fn); exit(1);
    case 3: fprintf(stderr, "error: %s(), beta must be in (0,1)\n", fn); exit(1);
    case 4: fprintf(stderr, "error: %s(), dt must be in [-1,1]\n", fn); exit(1);
    default: break;
    }
}
}  // namespace

// ---- C ABI: liquid.h:1463, :1583, :1586, :1664 ----
extern "C" float estimate_req_filter_As(float _df, unsigned int _N) { return lc::estimate_req_filter_As(_df, _N); }

extern "C" unsigned int estimate_req_filter_len(float _df, float _As)
{
    if (_df > 0.5f || _df <= 0.0f) lc::fatal("estimate_req_filter_len", "invalid bandwidth");
    if (_As <= 0.0f) lc::fatal("estimate_req_filter_len", "invalid stopband level");
    return lc::estimate_req_filter_len(_df, _As);
}

extern "C" void liquid_filter_isi(float *_h, unsigned int _k, unsigned int _m, float *_rms, float *_max)
{
    lc::filter_isi(_h, _k, _m, _rms, _max);
}

extern "C" void liquid_firdes_arkaiser(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h)
{
    rnyquist_fatal("liquid_firdes_arkaiser", lc::firdes_arkaiser(_k, _m, _beta, _dt, _h));
}

extern "C" void liquid_firdes_rkaiser(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h)
{
    rnyquist_fatal("liquid_firdes_rkaiser", lc::firdes_rkaiser(_k, _m, _beta, _dt, _h, nullptr));
}

// (root-)Nyquist prototype of 2km+1 taps by liquid_firfilt_type           liquid.h:1444, firdes.c:290-370
// Provided: KAISER, ARKAISER, RKAISER, RRC.  The Parks-McClellan based ones (PM, hM3: firdespm.c) and the
// rcos / fexp / fsech / farcsech / GMSK families belong to filter-design subsystems outside this library.
extern "C" void liquid_firdes_prototype(int _type, unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h)
{
    const unsigned int h_len = 2 * _k * _m + 1;
    switch (_type) {
    case LIQUID_FIRFILT_KAISER: {
        const float fc = 0.5f / (float)_k;                          // cut-off
        const float df = _beta / (float)_k;                         // transition bandwidth
        liquid_firdes_kaiser(h_len, fc, lc::estimate_req_filter_As(df, h_len), _dt, _h);
        return;
    }
    case LIQUID_FIRFILT_ARKAISER: liquid_firdes_arkaiser(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_RKAISER: liquid_firdes_rkaiser(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_RRC: liquid_firdes_rrcos(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_PM:
    case LIQUID_FIRFILT_RCOS:
    case LIQUID_FIRFILT_FEXP:
    case LIQUID_FIRFILT_FSECH:
    case LIQUID_FIRFILT_FARCSECH:
    case LIQUID_FIRFILT_hM3:
    case LIQUID_FIRFILT_GMSKTX:
    case LIQUID_FIRFILT_GMSKRX:
    case LIQUID_FIRFILT_RFEXP:
    case LIQUID_FIRFILT_RFSECH:
    case LIQUID_FIRFILT_RFARCSECH:
        fprintf(stderr, "error: liquid_firdes_prototype(), filter type '%d' is not provided by liquid_cuda "
                        "(KAISER, ARKAISER, RKAISER and RRC are; pass external taps to the create() form)\n", _type);
        exit(1);
    default:
        fprintf(stderr, "error: liquid_firdes_prototype(), invalid root-Nyquist filter type '%d'\n", _type);
        exit(1);
    }
}
