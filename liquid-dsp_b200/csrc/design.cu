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

namespace lc {

// ---- Parks-McClellan equiripple design (firdespm.c) and the harris-Moerder prototype built on it (hM3.c) ---------
// Band-pass (even-symmetric) filters, the only kind firdespm_compute_taps() of the reference produces.  All of it in
// double precision with the reference's operation order: frequency grid of 20 points per approximating function on
// the bands (firdespm.c:214-223, :346-455), Remez exchange with a barycentric Lagrange interpolant on the r + 1
// extremal points (:456-497, poly.lagrange.c:104-151), extremal search keeping alternating signs and dropping the
// smallest surplus extrema (:518-659), at most 40 rounds or until the extremal errors agree to 1e-3 (:663-686), then
// the taps from the interpolant sampled on h_len points (:689-748).
namespace {
struct PmDesign {
    unsigned int h_len, s, n, r, grid = 0;
    std::vector<double> F, D, W, E, x, alpha, c;
    std::vector<unsigned int> iext;
    double rho = 0.0;
    unsigned int exchanges = 0;

    static double lagrange_val(const std::vector<double> &x, const std::vector<double> &y, const std::vector<double> &w, double x0)
    {
        double t0 = 0.0, t1 = 0.0;
        const float tol = 1e-6f;
        for (size_t j = 0; j < x.size(); j++) {
            const double g = x0 - x[j];
            if (fabs(g) < tol) return y[j];
            t0 += w[j] * y[j] / g;
            t1 += w[j] / g;
        }
        return t0 / t1;
    }

    void interpolant()
    {
        for (unsigned int i = 0; i <= r; i++) x[i] = cos(2 * M_PI * F[iext[i]]);
        for (unsigned int j = 0; j <= r; j++) {               // barycentric weights, normalised by the first
            double w = 1.;
            for (unsigned int k = 0; k <= r; k++)
                if (j != k) w *= (x[j] - x[k]);
            alpha[j] = 1. / w;
        }
        const double w0 = alpha[0];
        for (unsigned int j = 0; j <= r; j++) alpha[j] /= w0;
        double t0 = 0.0, t1 = 0.0;
        for (unsigned int i = 0; i <= r; i++) {
            t0 += alpha[i] * D[iext[i]];
            t1 += alpha[i] / W[iext[i]] * (i % 2 ? -1.0 : 1.0);
        }
        rho = t0 / t1;
        for (unsigned int i = 0; i <= r; i++) c[i] = D[iext[i]] - (i % 2 ? -1 : 1) * rho / W[iext[i]];
    }

    void errors()
    {
        for (unsigned int i = 0; i < grid; i++) {
            const double xf = cos(2 * M_PI * F[i]);
            E[i] = W[i] * (D[i] - lagrange_val(x, c, alpha, xf));
        }
    }

    void exchange(unsigned int num_bands)
    {
        std::vector<unsigned int> found;
        found.reserve(2 * r + 2 * num_bands);
        found.push_back(0);                                    // both ends are always candidates
        for (unsigned int i = 1; i + 1 < grid; i++)
            if ((E[i] >= 0.0 && E[i - 1] <= E[i] && E[i + 1] <= E[i]) || (E[i] < 0.0 && E[i - 1] >= E[i] && E[i + 1] >= E[i]))
                found.push_back(i);
        found.push_back(grid - 1);
        if (found.size() < r + 1) {
            fprintf(stderr, "warning: firdespm_iext_search(), too few extrema found (expected %u, found %u); exiting prematurely\n",
                    r + 1, (unsigned int)found.size());
            exchanges = 0;
            return;
        }
        size_t extra = found.size() - r - 1;
        while (extra) {
            unsigned int sign = E[found[0]] > 0.0;
            size_t imin = 0;
            bool alternating = true;
            for (size_t i = 1; i < found.size(); i++) {
                if (fabs(E[found[i]]) < fabs(E[found[imin]])) imin = i;
                if (sign && E[found[i]] < 0.0) sign = 0;
                else if (!sign && E[found[i]] >= 0.0) sign = 1;
                else {                                         // two neighbours of one sign: drop the smaller
                    imin = fabs(E[found[i]]) < fabs(E[found[i - 1]]) ? i : i - 1;
                    alternating = false;
                    break;
                }
            }
            if (alternating && extra == 1) imin = fabs(E[found.front()]) < fabs(E[found.back()]) ? 0 : found.size() - 1;
            found.erase(found.begin() + (long)imin);
            extra--;
        }
        exchanges = 0;
        for (unsigned int i = 0; i <= r; i++) exchanges += iext[i] == found[i] ? 0 : 1;
        for (unsigned int i = 0; i <= r; i++) iext[i] = found[i];
    }

    bool converged() const
    {
        if (exchanges == 0) return true;
        const double tol = 1e-3f;
        double emin = 0.0, emax = 0.0;
        for (unsigned int i = 0; i <= r; i++) {
            const double e = fabs(E[iext[i]]);
            if (i == 0 || e < emin) emin = e;
            if (i == 0 || e > emax) emax = e;
        }
        return (emax - emin) / emax < tol;
    }
};
}  // namespace

// 0 ok; 1 invalid bands; 2 invalid weights; 3 no bands; 4 invalid weight type; 5 band type without a tap formula
int firdespm_bandpass(unsigned int h_len, unsigned int num_bands, const float *bands, const float *des, const float *weights,
                      const int *wtype, float *h)
{
    bool bands_ok = true, weights_ok = true;
    for (unsigned int i = 0; i < 2 * num_bands; i++) bands_ok &= bands[i] >= 0.0 && bands[i] <= 0.5;
    for (unsigned int i = 1; i < 2 * num_bands; i++) bands_ok &= bands[i] >= bands[i - 1];
    for (unsigned int i = 0; i < num_bands; i++) weights_ok &= (weights == nullptr || weights[i] > 0);
    if (!bands_ok) return 1;
    if (!weights_ok) return 2;
    if (num_bands == 0) return 3;
    PmDesign q;
    q.h_len = h_len;
    q.s = h_len % 2;
    q.n = (h_len - q.s) / 2;
    q.r = q.n + q.s;
    const unsigned int density = 20;
    const double df = 0.5 / (density * q.r);
    // the grid: every band sampled at df, its last point forced onto the upper edge
    for (unsigned int b = 0; b < num_bands; b++) {
        const double f0 = bands[2 * b], f1 = bands[2 * b + 1];
        unsigned int pts = (unsigned int)((f1 - f0) / df + 0.5);
        if (pts < 1) pts = 1;
        for (unsigned int j = 0; j < pts; j++) {
            double fw = 1.0f;                                  // (a double in the reference too: firdespm.c:381)
            const int wt = wtype ? wtype[b] : LIQUID_FIRDESPM_FLATWEIGHT;
            if (wt == LIQUID_FIRDESPM_FLATWEIGHT) fw = 1.0f;
            else if (wt == LIQUID_FIRDESPM_EXPWEIGHT) fw = expf((float)(2.0f * j * df));
            else if (wt == LIQUID_FIRDESPM_LINWEIGHT) fw = 1.0f + 2.7f * j * df;
            else return 4;
            q.F.push_back(f0 + j * df);
            q.D.push_back((double)des[b]);
            q.W.push_back((double)(weights ? weights[b] : 1.0f) * fw);
        }
        q.F.back() = f1;
    }
    q.grid = (unsigned int)q.F.size();
    if (q.s == 0)                                               // even length: the cos(pi f) factor moves into D and W
        for (unsigned int i = 0; i < q.grid; i++) {
            q.D[i] /= cos(M_PI * q.F[i]);
            q.W[i] *= cos(M_PI * q.F[i]);
        }
    q.E.assign(q.grid, 0.0);
    q.x.assign(q.r + 1, 0.0);
    q.alpha.assign(q.r + 1, 0.0);
    q.c.assign(q.r + 1, 0.0);
    q.iext.resize(q.r + 1);
    for (unsigned int i = 0; i <= q.r; i++) q.iext[i] = (i * (q.grid - 1)) / q.r;
    for (unsigned int it = 0; it < 40; it++) {
        q.interpolant();
        q.errors();
        q.exchange(num_bands);
        if (q.converged()) break;
    }
    q.interpolant();
    const unsigned int p = q.r - q.s + 1;
    std::vector<double> G(p);
    for (unsigned int i = 0; i < p; i++) {
        const double f = (double)i / (double)h_len;
        const double cf = PmDesign::lagrange_val(q.x, q.c, q.alpha, cos(2 * M_PI * f));
        const double g = q.s == 1 ? 1.0 : cos(M_PI * i / h_len);
        G[i] = cf * g;
    }
    for (unsigned int i = 0; i < h_len; i++) {
        double v = G[0];
        const double f = ((double)i - (double)(p - 1) + 0.5 * (1 - q.s)) / (double)h_len;
        for (unsigned int j = 1; j < q.r; j++) v += 2.0 * G[j] * cos(2 * M_PI * f * j);
        h[i] = (float)(v / (double)h_len);
    }
    return 0;
}

// hM3.c:49-126: three-band equiripple design (pass band to fc(1-beta), 1/sqrt(2) at fc, stop band from fc(1+beta)
// with exponential weight), the pass-band edge walked towards fc in 100 steps while the inter-symbol interference
// keeps falling; unit energy per symbol.  0 ok, else message id
int firdes_hM3(unsigned int k, unsigned int m, float beta, float dt, float *h_out)
{
    (void)dt;
    if (k < 2) return 1;
    if (m < 1) return 2;
    if (beta < 0.0f || beta > 1.0f) return 3;
    const unsigned int n = 2 * k * m + 1;
    const float fc = (float)(1.0 / (float)(2 * k));
    float fp = (float)(fc * (1.0 - beta));
    const float fs = (float)(fc * (1.0 + beta));
    float bands[6] = {0.0f, fp, fc, fc, fs, 0.5f};
    const float des[3] = {1.0f, 1.0f / sqrtf(2.0f), 0.0f};
    const float weights[3] = {1.0f, 1.0f, 1.0f};
    const int wtype[3] = {LIQUID_FIRDESPM_FLATWEIGHT, LIQUID_FIRDESPM_FLATWEIGHT, LIQUID_FIRDESPM_EXPWEIGHT};
    std::vector<float> h(n);
    firdespm_bandpass(n, 3, bands, des, weights, wtype, h.data());
    std::copy(h.begin(), h.end(), h_out);
    float isi_rms, isi_max;
    filter_isi(h.data(), k, m, &isi_rms, &isi_max);
    float isi_rms_min = isi_rms;
    const unsigned int pmax = 100;
    for (unsigned int p = 0; p < pmax; p++) {
        fp = (float)(fc * (1.0 - beta * p / (float)pmax));
        bands[1] = fp;
        firdespm_bandpass(n, 3, bands, des, weights, wtype, h.data());
        filter_isi(h.data(), k, m, &isi_rms, &isi_max);
        if (isi_rms > isi_rms_min) break;
        isi_rms_min = isi_rms;
        std::copy(h.begin(), h.end(), h_out);
    }
    float e2 = 0.0f;
    for (unsigned int i = 0; i < n; i++) e2 += h_out[i] * h_out[i];
    for (unsigned int i = 0; i < n; i++) h_out[i] *= sqrtf(k / e2);
    return 0;
}

}  // namespace lc

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
    if (_df > 0.5f || _df <= 0.0f) {
        fprintf(stderr, "error: estimate_req_filter_len(), invalid bandwidth : %f\n", _df);
        exit(1);
    }
    if (_As <= 0.0f) {
        fprintf(stderr, "error: estimate_req_filter_len(), invalid stopband level : %f\n", _As);
        exit(1);
    }
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
namespace lc {
namespace {
// unscaled DFT of a short odd-length design array (the reference uses its float FFT here, fnyquist.c:86, gmsk.c:152-179;
// a direct sum in double agrees with it to float rounding): y[k] = sum_i x[i] e^{sign j 2 pi i k / n}
void design_dft(const std::vector<lc_cf> &x, std::vector<lc_cf> &y, int sign)
{
    const unsigned int n = (unsigned int)x.size();
    y.resize(n);
    for (unsigned int k = 0; k < n; k++) {
        double re = 0.0, im = 0.0;
        for (unsigned int i = 0; i < n; i++) {
            const double a = (double)sign * 2.0 * M_PI * (double)(((unsigned long long)i * k) % n) / (double)n;
            const double c = cos(a), sn = sin(a);
            re += x[i].re * c - x[i].im * sn;
            im += x[i].re * sn + x[i].im * c;
        }
        y[k] = lc_cf{(float)re, (float)im};
    }
}

int nyquist_args(unsigned int k, unsigned int m, float beta)       // 0 ok, else message id
{
    if (k < 1) return 1;
    if (m < 1) return 2;
    if (beta < 0.0f || beta > 1.0f) return 3;
    return 0;
}

float asechf_(float z)                                              // fnyquist.c:253-262
{
    if (z <= 0.0f || z > 1.0f) {
        fprintf(stderr, "warning: liquid_asechf(), input out of range\n");
        return 0.0f;
    }
    const float zi = 1.0f / z;
    return logf(sqrtf(zi - 1.0f) * sqrtf(zi + 1.0f) + zi);
}
}  // namespace

// rcos.c:39-80: raised cosine, cos(pi beta z) sinc(z) / (1 - 4 beta^2 z^2) with the removable singularity patched
int firdes_rcos(unsigned int k, unsigned int m, float beta, float dt, float *h)
{
    if (int rc = nyquist_args(k, m, beta)) return rc;
    const unsigned int h_len = 2 * k * m + 1;
    for (unsigned int n = 0; n < h_len; n++) {
        const float z = ((float)n + dt) / (float)k - (float)m;
        const float t1 = cosf((float)(beta * M_PI * z));
        const float t2 = sincf_(z);
        const float t3 = 1 - 4.0f * beta * beta * z * z;
        if (fabsf(t3) < 1e-3f) h[n] = (float)(sinf((float)(M_PI / (2.0 * beta))) * beta * 0.5f);
        else h[n] = t1 * t2 / t3;
    }
    return 0;
}

// fnyquist.c:47-93, :121-318: flipped-exponential / -sech / -arcsech Nyquist filters designed in the frequency domain
// (kind 0 / 1 / 2), optionally the square root of the response, inverse transform, shift and scale
int firdes_fnyquist(int kind, int root, unsigned int k, unsigned int m, float beta, float dt, float *h)
{
    (void)dt;
    if (int rc = nyquist_args(k, m, beta)) return rc;
    const unsigned int h_len = 2 * k * m + 1;
    const float f0 = 0.5f * (1.0f - beta) / (float)k, f1 = 0.5f * (1.0f) / (float)k, f2 = 0.5f * (1.0f + beta) / (float)k;
    const float B = 0.5f / (float)k;
    const float gamma = (kind == 0 ? logf(2.0f) : logf(sqrtf(3.0f) + 2.0f)) / (beta * B);
    const float zeta = 1.0f / (2.0f * beta * B);
    std::vector<lc_cf> H(h_len), t;
    for (unsigned int i = 0; i < h_len; i++) {
        float f = (float)i / (float)h_len;
        if (f > 0.5f) f = f - 1.0f;
        f = fabsf(f);
        float v;
        if (f < f0) v = 1.0f;
        else if (f > f0 && f < f2) {
            if (kind == 0) v = f < f1 ? expf(gamma * (B * (1 - beta) - f)) : 1.0f - expf(gamma * (f - (1 + beta) * B));
            else if (kind == 1) v = f < f1 ? 1.0f / coshf(gamma * (f - B * (1 - beta))) : 1.0f - 1.0f / coshf(gamma * (B * (1 + beta) - f));
            else v = f < f1 ? 1.0f - (zeta / gamma) * asechf_(zeta * (B * (1 + beta) - f)) : (zeta / gamma) * asechf_(zeta * (f - B * (1 - beta)));
        } else v = 0.0f;
        H[i] = lc_cf{root ? sqrtf(v) : v, 0.0f};
    }
    design_dft(H, t, +1);
    for (unsigned int i = 0; i < h_len; i++) h[i] = t[(i + k * m + 1) % h_len].re * (float)k / (float)h_len;
    return 0;
}

// gmsk.c:39-78: Gaussian pulse-shaping filter (difference of two Q functions), normalised to an integral of pi/2, times k
int firdes_gmsktx(unsigned int k, unsigned int m, float beta, float dt, float *h)
{
    if (int rc = nyquist_args(k, m, beta)) return rc;
    const unsigned int h_len = 2 * k * m + 1;
    const float c0 = 1.0f / sqrtf(logf(2.0f));
    auto Q = [](float z) { return 0.5f * (1.0f - erff((float)(z * M_SQRT1_2))); };
    for (unsigned int i = 0; i < h_len; i++) {
        const float t = (float)i / (float)k - (float)m + dt;
        h[i] = Q((float)(2 * M_PI * beta * (t - 0.5f) * c0)) - Q((float)(2 * M_PI * beta * (t + 0.5f) * c0));
    }
    float e = 0.0f;
    for (unsigned int i = 0; i < h_len; i++) e += h[i];
    for (unsigned int i = 0; i < h_len; i++) h[i] = (float)(h[i] * (M_PI / (2.0f * e)));
    for (unsigned int i = 0; i < h_len; i++) h[i] *= (float)k;
    return 0;
}

// gmsk.c:91-193: receive filter that, cascaded with the transmit filter, approximates a Kaiser Nyquist prototype:
// ratio of the two frequency responses (minima removed, small correction added), times a 60 dB Kaiser "gain" filter
int firdes_gmskrx(unsigned int k, unsigned int m, float beta, float dt, float *h)
{
    (void)dt;
    if (int rc = nyquist_args(k, m, beta)) return rc;
    const unsigned int h_len = 2 * k * m + 1;
    const float delta = 1e-3f;
    std::vector<float> ht(h_len), hp(h_len), gp(h_len);
    firdes_gmsktx(k, m, beta, 0.0f, ht.data());
    {
        const float fc = 0.5f / (float)k, df = beta / (float)k;       // the KAISER case of liquid_firdes_prototype
        firdes_kaiser(h_len, fc, estimate_req_filter_As(df, h_len), 0.0f, hp.data());
    }
    firdes_kaiser(h_len, (float)((0.7f + 0.1 * beta) / (float)k), 60.0f, 0.0f, gp.data());
    std::vector<lc_cf> a(h_len), b(h_len), c(h_len), Hp, Gp, Ht;
    for (unsigned int i = 0; i < h_len; i++) {
        a[i] = lc_cf{hp[(i + k * m) % h_len], 0.0f};
        b[i] = lc_cf{gp[(i + k * m) % h_len], 0.0f};
        c[i] = lc_cf{ht[(i + k * m) % h_len], 0.0f};
    }
    design_dft(a, Hp, -1);
    design_dft(b, Gp, -1);
    design_dft(c, Ht, -1);
    float Ht_min = 0.0f, Hp_min = 0.0f, Gp_min = 0.0f;
    for (unsigned int i = 0; i < h_len; i++) {
        if (i == 0 || Ht[i].re < Ht_min) Ht_min = Ht[i].re;
        if (i == 0 || Hp[i].re < Hp_min) Hp_min = Hp[i].re;
        if (i == 0 || Gp[i].re < Gp_min) Gp_min = Gp[i].re;
    }
    std::vector<lc_cf> Hh(h_len), hh;
    for (unsigned int i = 0; i < h_len; i++) {
        float v = (Hp[i].re - Hp_min + delta) / (Ht[i].re - Ht_min + delta);
        v *= (Gp[i].re - Gp_min) / Gp[0].re;
        Hh[i] = lc_cf{v, 0.0f};
    }
    design_dft(Hh, hh, +1);
    for (unsigned int i = 0; i < h_len; i++) {
        const float hr = hh[(i + k * m + 1) % h_len].re / (float)(k * h_len);
        h[i] = hr * k * k;
    }
    return 0;
}

}  // namespace lc

namespace {
// (these design functions of the reference punctuate "error: <fn>(): <msg>")
[[noreturn]] void fatal_colon(const char *fn, const char *msg)
{
    fprintf(stderr, "error: %s(): %s\n", fn, msg);
    exit(1);
}
void nyquist_fatal(const char *fn, int rc)
{
    if (rc == 1) fatal_colon(fn, "k must be greater than 0");
    if (rc == 2) fatal_colon(fn, "m must be greater than 0");
    if (rc) fatal_colon(fn, "beta must be in [0,1]");
}
}  // namespace

extern "C" void liquid_firdes_rcos(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h)
{
    nyquist_fatal("liquid_firdes_rcos", lc::firdes_rcos(_k, _m, _beta, _dt, _h));
}
#define LC_FNYQUIST(NAME, KIND, ROOT)                                                                          \
    extern "C" void liquid_firdes_##NAME(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h) \
    {                                                                                                          \
        nyquist_fatal("liquid_firdes_fnyquist", lc::firdes_fnyquist(KIND, ROOT, _k, _m, _beta, _dt, _h));      \
    }
LC_FNYQUIST(fexp, 0, 0)
LC_FNYQUIST(rfexp, 0, 1)
LC_FNYQUIST(fsech, 1, 0)
LC_FNYQUIST(rfsech, 1, 1)
LC_FNYQUIST(farcsech, 2, 0)
LC_FNYQUIST(rfarcsech, 2, 1)
#undef LC_FNYQUIST
extern "C" void liquid_firdes_gmsktx(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h)
{
    nyquist_fatal("liquid_firdes_gmsktx", lc::firdes_gmsktx(_k, _m, _beta, _dt, _h));
}
extern "C" void liquid_firdes_gmskrx(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h)
{
    nyquist_fatal("liquid_firdes_gmskrx", lc::firdes_gmskrx(_k, _m, _beta, _dt, _h));
}

extern "C" void firdespm_run(unsigned int _h_len, unsigned int _num_bands, float *_bands, float *_des, float *_weights,
                             int *_wtype, int _btype, float *_h)
{
    if (_btype != LIQUID_FIRDESPM_BANDPASS) {
        // (the reference only warns and leaves _h untouched for these: firdespm.c:737-743)
        fprintf(stderr, "warning: firdespm_compute_taps(), filter configuration not yet supported\n");
        return;
    }
    switch (lc::firdespm_bandpass(_h_len, _num_bands, _bands, _des, _weights, _wtype, _h)) {
    case 0: return;
    case 1: lc::fatal("firdespm_create", "invalid bands");
    case 2: lc::fatal("firdespm_create", "invalid weights (must be positive)");
    case 3: lc::fatal("firdespm_create", "number of bands must be > 0");
    default: lc::fatal("firdespm_init_grid", "invalid weighting specifyer");
    }
}

// the designer as an object (liquid.h:1514-1540, firdespm.c:142-343): the specification is kept, execute() designs
struct firdespm_s {
    unsigned int h_len, num_bands;
    std::vector<float> bands, des, weights;
    std::vector<int> wtype;
    int btype;
};

extern "C" firdespm firdespm_create(unsigned int _h_len, unsigned int _num_bands, float *_bands, float *_des, float *_weights,
                                    int *_wtype, int _btype)
{
    bool bands_ok = true, weights_ok = true;
    for (unsigned int i = 0; i < 2 * _num_bands; i++) bands_ok &= _bands[i] >= 0.0 && _bands[i] <= 0.5;
    for (unsigned int i = 1; i < 2 * _num_bands; i++) bands_ok &= _bands[i] >= _bands[i - 1];
    for (unsigned int i = 0; i < _num_bands; i++) weights_ok &= (_weights == nullptr || _weights[i] > 0);
    if (!bands_ok) lc::fatal("firdespm_create", "invalid bands");
    if (!weights_ok) lc::fatal("firdespm_create", "invalid weights (must be positive)");
    if (_num_bands == 0) lc::fatal("firdespm_create", "number of bands must be > 0");
    firdespm q = new firdespm_s();
    q->h_len = _h_len;
    q->num_bands = _num_bands;
    q->bands.assign(_bands, _bands + 2 * _num_bands);
    q->des.assign(_des, _des + _num_bands);
    q->weights.assign(_num_bands, 1.0f);
    if (_weights) q->weights.assign(_weights, _weights + _num_bands);
    q->wtype.assign(_num_bands, LIQUID_FIRDESPM_FLATWEIGHT);
    if (_wtype) q->wtype.assign(_wtype, _wtype + _num_bands);
    q->btype = _btype;
    return q;
}

extern "C" void firdespm_destroy(firdespm _q) { delete _q; }

extern "C" void firdespm_print(firdespm _q)
{
    printf("firdespm:               ");
    for (unsigned int i = 0; i < _q->num_bands; i++) printf("      band %-5u", i);
    printf("\n  lower band edge       ");
    for (unsigned int i = 0; i < _q->num_bands; i++) printf("%16.8f", _q->bands[2 * i]);
    printf("\n  upper band edge       ");
    for (unsigned int i = 0; i < _q->num_bands; i++) printf("%16.8f", _q->bands[2 * i + 1]);
    printf("\n  desired value         ");
    for (unsigned int i = 0; i < _q->num_bands; i++) printf("%16.8f", _q->des[i]);
    printf("\n  weighting             ");
    for (unsigned int i = 0; i < _q->num_bands; i++) printf("%16.8f", _q->weights[i]);
    printf("\n");
}

extern "C" void firdespm_execute(firdespm _q, float *_h)
{
    firdespm_run(_q->h_len, _q->num_bands, _q->bands.data(), _q->des.data(), _q->weights.data(), _q->wtype.data(), _q->btype, _h);
}

extern "C" void liquid_firdes_hM3(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h)
{
    switch (lc::firdes_hM3(_k, _m, _beta, _dt, _h)) {
    case 0: return;
    case 1: fatal_colon("liquid_firdes_hM3", "k must be greater than 1");
    case 2: fatal_colon("liquid_firdes_hM3", "m must be greater than 0");
    default: fatal_colon("liquid_firdes_hM3", "beta must be in [0,1]");
    }
}

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
    case LIQUID_FIRFILT_hM3: liquid_firdes_hM3(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_PM: {
        // firdes.c:303-323: three bands around the cut-off, k / k/2 / 0, flat weights (the timing offset is ignored)
        const float fc = 0.5f / (float)_k, df = _beta / (float)_k;
        float bands[6] = {0.0f, fc - 0.5f * df, fc, fc, fc + 0.5f * df, 0.5f};
        float des[3] = {(float)_k, 0.5f * _k, 0.0f};
        float weights[3] = {1.0f, 1.0f, 1.0f};
        int wtype[3] = {LIQUID_FIRDESPM_FLATWEIGHT, LIQUID_FIRDESPM_FLATWEIGHT, LIQUID_FIRDESPM_FLATWEIGHT};
        firdespm_run(h_len, 3, bands, des, weights, wtype, LIQUID_FIRDESPM_BANDPASS, _h);
        return;
    }
    case LIQUID_FIRFILT_RCOS: liquid_firdes_rcos(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_FEXP: liquid_firdes_fexp(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_FSECH: liquid_firdes_fsech(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_FARCSECH: liquid_firdes_farcsech(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_GMSKTX: liquid_firdes_gmsktx(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_GMSKRX: liquid_firdes_gmskrx(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_RFEXP: liquid_firdes_rfexp(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_RFSECH: liquid_firdes_rfsech(_k, _m, _beta, _dt, _h); return;
    case LIQUID_FIRFILT_RFARCSECH: liquid_firdes_rfarcsech(_k, _m, _beta, _dt, _h); return;
    default:
        fprintf(stderr, "error: liquid_firdes_prototype(), invalid root-Nyquist filter type '%d'\n", _type);
        exit(1);
    }
}
