/*
 * liquid_oracle.c -- CPU restatement of liquid-dsp's polyphase-channelizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see liquid_oracle.h).  Parity status: PINNED against the
 * reference's golden vectors and against the compiled reference (oracle/_ref).
 *
 * Written from the behaviour of the reference, not from its text: state is held as plain
 * "last L samples, oldest first" arrays instead of the reference's power-of-two ring with
 * memmove, FFT plans are replaced by a recursive function, objects are flat structs.
 * Paths cited are relative to /root/reference.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -pthread liquid_oracle.c -lm
 *   (-ffp-contract=off keeps a*b+c as two roundings, as the SSE mul/add intrinsics of
 *    src/dotprod/src/dotprod_crcf.mmx.c do.)
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "liquid_oracle.h"

/* ------------------------------------------------------------------------------------ */
/* small helpers                                                                         */
/* ------------------------------------------------------------------------------------ */

static orc_cf cf(float re, float im) { orc_cf v; v.re = re; v.im = im; return v; }

static orc_cf cf_mul(orc_cf a, orc_cf b)
{
    /* C99 complex multiply as the reference's `float complex` products evaluate */
    return cf(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}

static void *xmalloc(size_t n)
{
    void *p = malloc(n ? n : 1);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}

/* ------------------------------------------------------------------------------------ */
/* scalar math                                                                           */
/* ------------------------------------------------------------------------------------ */

/* floor(log2 x) + 1, 0 for x == 0.        src/utility/src/msb_index.c:110-135 */
unsigned int orc_msb_index(unsigned int x)
{
    unsigned int bits = 0;
    while (x) { bits++; x >>= 1; }
    return bits;
}

/* ln Gamma(z): upward recursion to z >= 10, then a Stirling-type closed form.
 * src/math/src/math.gamma.c:43-71 (the double-typed 0.5 and log() are kept) */
float orc_lngammaf(float z)
{
    if (z < 0) return NAN;                      /* reference exits; the path never gets here */
    if (z < 10.0f)
        return orc_lngammaf(z + 1.0f) - logf(z);
    float g = 0.5 * (logf(2 * M_PI) - log(z));
    g += z * (logf(z + (1 / (12.0f * z - 0.1f / z))) - 1);
    return g;
}

/* I0(z) by 32 terms of sum_k ((z/2)^k / k!)^2 evaluated through logs.
 * src/math/src/math.bessel.c:86-105 */
float orc_besseli0f(float z)
{
    if (z == 0.0f) return 1.0f;
    float y = 0.0f;
    for (unsigned int k = 0; k < 32; k++) {
        float t = k * logf(0.5f * z) - orc_lngammaf((float)k + 1.0f);
        y += expf(2 * t);
    }
    return y;
}

/* sin(pi x)/(pi x); three-cosine product below |x| < 0.01.   src/math/src/math.c:128-139 */
float orc_sincf(float x)
{
    if (fabsf(x) < 0.01f)
        return cosf(M_PI * x / 2.0f) * cosf(M_PI * x / 4.0f) * cosf(M_PI * x / 8.0f);
    return sinf(M_PI * x) / (M_PI * x);
}

/* Kaiser window sample; note r = 2t/N (not N-1).    src/math/src/math.c:289-310 */
float orc_kaiser(unsigned int n, unsigned int N, float beta, float mu)
{
    float t = (float)n - (float)(N - 1) / 2 + mu;
    float r = 2.0f * t / (float)(N);
    float a = orc_besseli0f(beta * sqrtf(1 - r * r));
    float b = orc_besseli0f(beta);
    return a / b;
}

/* src/filter/src/firdes.c:224-236 */
float orc_kaiser_beta_As(float As)
{
    As = fabsf(As);
    float beta;
    if (As > 50.0f)      beta = 0.1102f * (As - 8.7f);
    else if (As > 21.0f) beta = 0.5842 * powf(As - 21, 0.4f) + 0.07886f * (As - 21);
    else                 beta = 0.0f;
    return beta;
}

/* Kaiser-windowed sinc low-pass.   src/filter/src/firdes.c:244-280
 * returns 0 on success, -1 on the argument errors the reference exits on */
int orc_firdes_kaiser(unsigned int n, float fc, float As, float mu, float *h)
{
    if (mu < -0.5f || mu > 0.5f) return -1;
    if (fc < 0.0f || fc > 0.5f)  return -1;
    if (n == 0)                  return -1;
    float beta = orc_kaiser_beta_As(As);
    for (unsigned int i = 0; i < n; i++) {
        float t  = (float)i - (float)(n - 1) / 2 + mu;
        float h1 = orc_sincf(2.0f * fc * t);
        float h2 = orc_kaiser(i, n, beta, mu);
        h[i] = h1 * h2;
    }
    return 0;
}

/* square-root raised cosine, 2km+1 taps.   src/filter/src/rrcos.c:37-95 */
int orc_firdes_rrcos(unsigned int k, unsigned int m, float beta, float dt, float *h)
{
    if (k < 1 || m < 1 || beta < 0.0f || beta > 1.0f) return -1;
    unsigned int h_len = 2 * k * m + 1;
    float T = 1.0f;
    for (unsigned int n = 0; n < h_len; n++) {
        float nf = (float)n, kf = (float)k, mf = (float)m;
        float z  = (nf + dt) / kf - mf;
        float t1 = cosf((1 + beta) * M_PI * z);
        float t2 = sinf((1 - beta) * M_PI * z);
        if (fabsf(z) < 1e-5) {
            h[n] = 1 - beta + 4 * beta / M_PI;
        } else {
            float t3 = 1 / ((4 * beta * z));
            float g  = 1 - 16 * beta * beta * z * z;
            g *= g;
            if (g < 1e-5) {
                float g1 = 1 + 2.0f / M_PI;
                float g2 = sinf(0.25f * M_PI / beta);
                float g3 = 1 - 2.0f / M_PI;
                float g4 = cosf(0.25f * M_PI / beta);
                h[n] = beta / sqrtf(2.0f) * (g1 * g2 + g3 * g4);
            } else {
                float t4 = 4 * beta / (M_PI * sqrtf(T) * (1 - (16 * beta * beta * z * z)));
                h[n] = t4 * (t1 + (t2 * t3));
            }
        }
    }
    return 0;
}

/* stop-band attenuation a Kaiser design of N taps reaches over a transition band df: 20 bisection
 * steps on Kaiser's length estimate N = (As - 7.95) / (14.26 df).   src/filter/src/firdes.c:80-116, :163-177 */
float orc_estimate_req_filter_As(float df, unsigned int N)
{
    float lo = 0.01f, hi = 200.0f, As = 0.0f;
    for (unsigned int it = 0; it < 20; it++) {
        As = 0.5f * (hi + lo);
        float n_hat = (As - 7.95f) / (14.26f * df);
        if (n_hat < (float)N) lo = As;
        else                  hi = As;
    }
    return As;
}

/* auto-correlation of h at a lag (float32 running sum).   firdes.c:420-438 */
static float orc_filter_autocorr(const float *h, unsigned int n, unsigned int lag)
{
    if (lag >= n) return 0.0f;
    float r = 0.0f;
    for (unsigned int i = lag; i < n; i++) r += h[i] * h[i - lag];
    return r;
}

/* RMS and peak inter-symbol interference of a 2km+1-tap filter: normalised auto-correlation at the
 * symbol lags k, 2k, ..., 2mk.   firdes.c:503-531 */
void orc_filter_isi(const float *h, unsigned int k, unsigned int m, float *rms, float *peak)
{
    unsigned int n = 2 * k * m + 1;
    float r0 = orc_filter_autocorr(h, n, 0);
    float acc = 0.0f, mx = 0.0f;
    for (unsigned int i = 1; i <= 2 * m; i++) {
        float e = fabsf(orc_filter_autocorr(h, n, i * k) / r0);
        acc += e * e;
        if (i == 1 || e > mx) mx = e;
    }
    *rms = sqrtf(acc / (float)(2 * m));
    *peak = mx;
}

/* tabulated bandwidth-adjustment estimate rho(m, beta) = c0 + c1 ln(beta) + c2 ln^2(beta), clamped to
 * [0,1].   src/filter/src/rkaiser.c:155-210 */
float orc_rkaiser_approximate_rho(unsigned int m, float beta)
{
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
        c0 = tab[m - 1][0]; c1 = tab[m - 1][1]; c2 = tab[m - 1][2];
    } else {
        c0 = 0.056873 * logf(m + 1e-3f) + 0.781388;
        c1 = 0.05426f;
        c2 = -0.00386f;
    }
    float b = logf(beta);
    float rho = c0 + c1 * b + c2 * b * b;
    if (rho < 0.0f) rho = 0.0f;
    else if (rho > 1.0f) rho = 1.0f;
    return rho;
}

/* Kaiser-windowed sinc whose transition band is rho * beta / k wide, cut-off moved so that the filter is
 * (approximately) root-Nyquist; returns its RMS ISI.   rkaiser.c:482-516 */
static float orc_rkaiser_design_isi(unsigned int k, unsigned int m, float beta, float dt, float rho, float *h)
{
    unsigned int n = 2 * k * m + 1;
    float kf = (float)k;
    float del = beta * rho / kf;
    float As = orc_estimate_req_filter_As(del, n);
    float fc = 0.5f * (1 + beta * (1.0f - rho)) / kf;
    orc_firdes_kaiser(n, fc, As, dt, h);
    float rms, peak;
    orc_filter_isi(h, k, m, &rms, &peak);
    return rms;
}

/* scale to energy k.   rkaiser.c:144-147, :468-470 */
static void orc_normalize_energy(float *h, unsigned int n, unsigned int k)
{
    float e2 = 0.0f;
    for (unsigned int i = 0; i < n; i++) e2 += h[i] * h[i];
    for (unsigned int i = 0; i < n; i++) h[i] *= sqrtf(k / e2);
}

/* approximate root-Nyquist Kaiser: rho from a closed-form fit, one design.   rkaiser.c:82-148 */
int orc_firdes_arkaiser(unsigned int k, unsigned int m, float beta, float dt, float *h)
{
    if (k < 2 || m < 1 || beta <= 0.0f || beta >= 1.0f || dt < -1.0f || dt > 1.0f) return -1;
    float c0 = 0.762886 + 0.067663 * logf(m);
    float c1 = 0.065515;
    float c2 = logf(1 - 0.088 * powf(m, -1.6));
    float lb = logf(beta);
    float rho = c0 + c1 * lb + c2 * lb * lb;
    if (rho <= 0.0f || rho >= 1.0f) rho = orc_rkaiser_approximate_rho(m, beta);
    unsigned int n = 2 * k * m + 1;
    float kf = (float)k;
    float del = beta * rho / kf;
    float As = orc_estimate_req_filter_As(del, n);
    float fc = 0.5f * (1 + beta * (1.0f - rho)) / kf;
    orc_firdes_kaiser(n, fc, As, dt, h);
    orc_normalize_energy(h, n, k);
    return 0;
}

/* root-Nyquist Kaiser: rho found by a bracketed parabolic search (at most 14 rounds of three designs)
 * for the least RMS ISI, then one more design at the best rho seen.   rkaiser.c:47-76, :339-474 */
int orc_firdes_rkaiser(unsigned int k, unsigned int m, float beta, float dt, float *h, float *rho_out)
{
    if (k < 2 || m < 1 || beta <= 0.0f || beta >= 1.0f || dt < -1.0f || dt > 1.0f) return -1;
    unsigned int n = 2 * k * m + 1;
    float rho_hat = orc_rkaiser_approximate_rho(m, beta);
    float rho_opt = rho_hat, y_opt = 0.0f;
    float x1 = rho_hat, dx = 0.2f;
    const float tol = 1e-6f;
    for (unsigned int p = 0; p < 14; p++) {
        float x0 = x1 - dx, x2 = x1 + dx;
        if (x0 <= 0.0f) x0 = 0.01f;
        if (x2 >= 1.0f) x2 = 0.99f;
        float y0 = orc_rkaiser_design_isi(k, m, beta, dt, x0, h);
        float y1 = orc_rkaiser_design_isi(k, m, beta, dt, x1, h);
        float y2 = orc_rkaiser_design_isi(k, m, beta, dt, x2, h);
        if (p == 0 || y1 < y_opt) { rho_opt = x1; y_opt = y1; }
        /* vertex of the parabola through the three points (float arithmetic, widened to double only for the division) */
        double ta = y0 * (x1 * x1 - x2 * x2) + y1 * (x2 * x2 - x0 * x0) + y2 * (x0 * x0 - x1 * x1);
        double tb = y0 * (x1 - x2) + y1 * (x2 - x0) + y2 * (x0 - x1);
        float x_hat = 0.5f * ta / tb;
        if (x_hat < x0 || x_hat > x2) break;
        float del = x_hat - x1;
        if (p > 3 && fabsf(del) < tol) break;
        x1 = x_hat;
        dx *= 0.5f;
    }
    orc_rkaiser_design_isi(k, m, beta, dt, rho_opt, h);
    orc_normalize_energy(h, n, k);
    if (rho_out) *rho_out = rho_opt;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* dot products                                                                          */
/* ------------------------------------------------------------------------------------ */

/* sequential sum.   src/dotprod/src/dotprod.c:41-56 */
void orc_dotprod_crcf_run(const float *h, const orc_cf *x, unsigned int n, orc_cf *y)
{
    float re = 0.0f, im = 0.0f;
    for (unsigned int i = 0; i < n; i++) {
        re += h[i] * x[i].re;
        im += h[i] * x[i].im;
    }
    *y = cf(re, im);
}

/* same order as _run (four per trip, then the tail).   src/dotprod/src/dotprod.c:63-89 */
void orc_dotprod_crcf_run4(const float *h, const orc_cf *x, unsigned int n, orc_cf *y)
{
    orc_dotprod_crcf_run(h, x, n, y);
}

/* The x86 build keeps four float lanes {re,im,re,im} (sixteen when n >= 32), folds them
 * at the end and finishes the tail sequentially.   src/dotprod/src/dotprod_crcf.mmx.c:144-287
 * Here the 2n-float view of x is walked with explicit lane accumulators. */
void orc_dotprod_crcf_run_sse_order(const float *h, const orc_cf *x, unsigned int n_taps, orc_cf *y)
{
    const float *xf = (const float *)x;
    unsigned int n = 2 * n_taps;                 /* floats */
    float w0, w1;
    unsigned int i;
    if (n_taps < 32) {
        float s[4] = {0, 0, 0, 0};
        unsigned int t = (n >> 2) << 2;
        for (i = 0; i < t; i += 4)
            for (unsigned int l = 0; l < 4; l++)
                s[l] += xf[i + l] * h[(i + l) >> 1];
        w0 = s[0] + s[2];
        w1 = s[1] + s[3];
    } else {
        float s[4][4];
        memset(s, 0, sizeof s);
        unsigned int r = (n >> 4) << 2;
        for (i = 0; i < r; i += 4)
            for (unsigned int a = 0; a < 4; a++)
                for (unsigned int l = 0; l < 4; l++) {
                    unsigned int j = 4 * i + 4 * a + l;
                    s[a][l] += xf[j] * h[j >> 1];
                }
        for (unsigned int l = 0; l < 4; l++) {
            s[0][l] = s[0][l] + s[1][l];
            s[2][l] = s[2][l] + s[3][l];
            s[0][l] = s[0][l] + s[2][l];
        }
        w0 = s[0][0] + s[0][2];
        w1 = s[0][1] + s[0][3];
        i  = 4 * r;
    }
    for (; i < n; i += 2) {
        w0 += xf[i] * h[i >> 1];
        w1 += xf[i + 1] * h[(i + 1) >> 1];
    }
    *y = cf(w0, w1);
}

/* the dot product the filter objects use: x86 lane order, as configure selects on this
 * platform (configure.ac:156-167) */
static orc_cf dot_crcf(const float *h, const orc_cf *x, unsigned int n)
{
    orc_cf y;
    orc_dotprod_crcf_run_sse_order(h, x, n, &y);
    return y;
}

/* ------------------------------------------------------------------------------------ */
/* FFT                                                                                   */
/* ------------------------------------------------------------------------------------ */


static unsigned int smallest_factor(unsigned int n)
{
    for (unsigned int f = 2; f * f <= n; f++)
        if (n % f == 0) return f;
    return n;
}

/* e^{j d 2 pi i / n} with the reference's float/double mix.   fft_mixed_radix.c:103-105 */
static orc_cf twiddle_calc(int dir, unsigned int i, unsigned int n)
{
    float d   = (dir == ORC_FFT_FORWARD) ? -1.0f : 1.0f;
    float arg = (float)((double)(d * 2) * M_PI * (double)(float)i / (double)(float)n);
    return cf(cosf(arg), sinf(arg));
}

/* per-(n, dir) twiddle tables, built once (the reference builds them per plan) */
struct tw_entry { unsigned int n; int dir; orc_cf *w; struct tw_entry *next; };
static struct tw_entry *tw_head = NULL;
static pthread_mutex_t tw_lock = PTHREAD_MUTEX_INITIALIZER;

static const orc_cf *twiddles(int dir, unsigned int n)
{
    pthread_mutex_lock(&tw_lock);
    struct tw_entry *e;
    for (e = tw_head; e; e = e->next)
        if (e->n == n && e->dir == dir) break;
    if (!e) {
        e = xmalloc(sizeof *e);
        e->n = n; e->dir = dir;
        e->w = xmalloc(n * sizeof *e->w);
        for (unsigned int i = 0; i < n; i++) e->w[i] = twiddle_calc(dir, i, n);
        e->next = tw_head;
        tw_head = e;
    }
    pthread_mutex_unlock(&tw_lock);
    return e->w;
}

/* O(n^2) DFT in float.  Stands in for the codelets (fft_dft.c:158-520), the generic
 * dotprod DFT (fft_dft.c:124-153) and Rader (fft_rader.c, fft_rader2.c): same transform,
 * different float rounding (~1e-7 relative). */
static void dft_direct(unsigned int n, const orc_cf *x, orc_cf *y, int dir)
{
    const orc_cf *w = twiddles(dir, n);
    for (unsigned int k = 0; k < n; k++) {
        float re = x[0].re, im = x[0].im;
        for (unsigned int i = 1; i < n; i++) {
            orc_cf p = cf_mul(x[i], w[(unsigned int)(((unsigned long long)k * i) % n)]);
            re += p.re;
            im += p.im;
        }
        y[k] = cf(re, im);
    }
}

/* Planner + mixed-radix Cooley-Tukey.
 *   method choice   src/fft/src/fft_utilities.c:34-68   (n<=8, 11, 13, 16, 17 -> DFT)
 *   radix choice    src/fft/src/fft_mixed_radix.c:194-242 (16, 8, 4, 2, else least factor)
 *   execute         src/fft/src/fft_mixed_radix.c:128-191 */
static void fft_rec(unsigned int n, const orc_cf *x, orc_cf *y, int dir)
{
    if (n == 1) { y[0] = x[0]; return; }
    unsigned int f = smallest_factor(n);
    if (n <= 8 || n == 11 || n == 13 || n == 16 || n == 17 || f == n) {
        dft_direct(n, x, y, dir);
        return;
    }
    unsigned int Q;
    if (n % 2 == 0) Q = (n % 16 == 0) ? 16 : (n % 8 == 0) ? 8 : (n % 4 == 0) ? 4 : 2;
    else            Q = f;
    unsigned int P = n / Q;
    unsigned int tl = Q > P ? Q : P;
    orc_cf *buf = xmalloc(n * sizeof *buf);
    orc_cf *t0  = xmalloc(tl * sizeof *t0);
    orc_cf *t1  = xmalloc(tl * sizeof *t1);
    const orc_cf *w = twiddles(dir, n);
    memcpy(buf, x, n * sizeof *buf);
    /* Q strided P-point transforms, each followed by the twiddle W^(i k) */
    for (unsigned int i = 0; i < Q; i++) {
        for (unsigned int k = 0; k < P; k++) t0[k] = buf[Q * k + i];
        fft_rec(P, t0, t1, dir);
        for (unsigned int k = 0; k < P; k++) buf[Q * k + i] = cf_mul(t1[k], w[i * k]);
    }
    /* P contiguous Q-point transforms, written transposed */
    for (unsigned int i = 0; i < P; i++) {
        for (unsigned int k = 0; k < Q; k++) t0[k] = buf[Q * i + k];
        fft_rec(Q, t0, t1, dir);
        for (unsigned int k = 0; k < Q; k++) y[k * P + i] = t1[k];
    }
    free(buf); free(t0); free(t1);
}

/* unscaled DFT; FORWARD = e^{-j..}, BACKWARD = e^{+j..}.   src/fft/src/fft_common.c:285-316 */
void orc_fft_run(unsigned int n, const orc_cf *x, orc_cf *y, int dir)
{
    if (n == 0) return;
    if (x == y) {                                 /* allow in-place use */
        orc_cf *tmp = xmalloc(n * sizeof *tmp);
        memcpy(tmp, x, n * sizeof *tmp);
        fft_rec(n, tmp, y, dir);
        free(tmp);
    } else {
        fft_rec(n, x, y, dir);
    }
}

/* ------------------------------------------------------------------------------------ */
/* sliding window: "the last len samples, oldest first".  src/buffer/src/window.c:200-214 */
/* ------------------------------------------------------------------------------------ */

static void win_push(orc_cf *w, unsigned int len, orc_cf v)
{
    memmove(w, w + 1, (len - 1) * sizeof *w);
    w[len - 1] = v;
}

/* branch i of a prototype, stored newest-tap-last so it lines up with an oldest-first
 * window:  sub[len-1-n] = h[i + n M].   firpfbch.c:115-124, firpfbch2.c:101-109, firpfb.c:71-84 */
static float *polyphase_split(const float *h, unsigned int M, unsigned int len)
{
    float *sub = xmalloc((size_t)M * len * sizeof *sub);
    for (unsigned int i = 0; i < M; i++)
        for (unsigned int n = 0; n < len; n++)
            sub[(size_t)i * len + (len - 1 - n)] = h[i + n * M];
    return sub;
}

/* ------------------------------------------------------------------------------------ */
/* firfilt_crcf                                                                          */
/* ------------------------------------------------------------------------------------ */

struct orc_firfilt_s {
    unsigned int n;
    float *hr;          /* taps reversed: hr[n-1-k] = h[k]      firfilt.c:87-90 */
    orc_cf *w;          /* last n samples, oldest first */
    float scale;        /* firfilt.c:96 */
};

orc_firfilt orc_firfilt_create(const float *h, unsigned int n)
{
    if (n == 0) return NULL;                      /* firfilt.c:66-69 (reference exits) */
    orc_firfilt q = xmalloc(sizeof *q);
    q->n  = n;
    q->hr = xmalloc(n * sizeof *q->hr);
    q->w  = xmalloc(n * sizeof *q->w);
    for (unsigned int k = 0; k < n; k++) q->hr[n - 1 - k] = h[k];
    q->scale = 1;
    orc_firfilt_reset(q);
    return q;
}

/* firfilt.c:109-131 */
orc_firfilt orc_firfilt_create_kaiser(unsigned int n, float fc, float As, float mu)
{
    if (n == 0) return NULL;
    float *h = xmalloc(n * sizeof *h);
    if (orc_firdes_kaiser(n, fc, As, mu, h)) { free(h); return NULL; }
    orc_firfilt q = orc_firfilt_create(h, n);
    free(h);
    return q;
}

void orc_firfilt_destroy(orc_firfilt q) { if (q) { free(q->hr); free(q->w); free(q); } }
void orc_firfilt_reset(orc_firfilt q) { memset(q->w, 0, q->n * sizeof *q->w); }   /* firfilt.c:253-263 */
void orc_firfilt_set_scale(orc_firfilt q, float s) { q->scale = s; }               /* firfilt.c:288 */
void orc_firfilt_push(orc_firfilt q, orc_cf x) { win_push(q->w, q->n, x); }        /* firfilt.c:297-317 */
unsigned int orc_firfilt_get_length(orc_firfilt q) { return q->n; }

/* firfilt.c:322-338: dot product, then the scale */
void orc_firfilt_execute(orc_firfilt q, orc_cf *y)
{
    orc_cf r = dot_crcf(q->hr, q->w, q->n);
    *y = cf(r.re * q->scale, r.im * q->scale);
}

/* firfilt.c:346-358; x == y allowed */
void orc_firfilt_execute_block(orc_firfilt q, const orc_cf *x, unsigned int n, orc_cf *y)
{
    for (unsigned int i = 0; i < n; i++) {
        orc_firfilt_push(q, x[i]);
        orc_firfilt_execute(q, &y[i]);
    }
}

/* ------------------------------------------------------------------------------------ */
/* firpfb_crcf: M sub-filters sharing one window                                         */
/* ------------------------------------------------------------------------------------ */

struct orc_firpfb_s {
    unsigned int M, sub_len;
    float *sub;         /* [M][sub_len] */
    orc_cf *w;
    float scale;
};

/* firpfb.c:46-95; sub_len = h_len / M (integer division, trailing taps dropped) */
orc_firpfb orc_firpfb_create(unsigned int M, const float *h, unsigned int h_len)
{
    if (M == 0 || h_len == 0) return NULL;
    orc_firpfb q = xmalloc(sizeof *q);
    q->M = M;
    q->sub_len = h_len / M;
    q->sub = polyphase_split(h, M, q->sub_len);
    q->w = xmalloc((q->sub_len ? q->sub_len : 1) * sizeof *q->w);
    q->scale = 1;
    orc_firpfb_reset(q);
    return q;
}

/* firpfb.c:105-137: 2Mm+1 taps, cut-off fc/M */
orc_firpfb orc_firpfb_create_kaiser(unsigned int M, unsigned int m, float fc, float As)
{
    if (M == 0 || m == 0 || fc < 0.0f || fc > 0.5f || As < 0.0f) return NULL;
    unsigned int H = 2 * M * m + 1;
    float *h = xmalloc(H * sizeof *h);
    orc_firdes_kaiser(H, fc / (float)M, As, 0.0f, h);
    orc_firpfb q = orc_firpfb_create(M, h, H);
    free(h);
    return q;
}

void orc_firpfb_destroy(orc_firpfb q) { if (q) { free(q->sub); free(q->w); free(q); } }
void orc_firpfb_reset(orc_firpfb q) { memset(q->w, 0, q->sub_len * sizeof *q->w); }
void orc_firpfb_set_scale(orc_firpfb q, float s) { q->scale = s; }
void orc_firpfb_push(orc_firpfb q, orc_cf x) { win_push(q->w, q->sub_len, x); }   /* firpfb.c:315-319 */

/* firpfb.c:325-346 */
void orc_firpfb_execute(orc_firpfb q, unsigned int i, orc_cf *y)
{
    orc_cf r = dot_crcf(q->sub + (size_t)i * q->sub_len, q->w, q->sub_len);
    *y = cf(r.re * q->scale, r.im * q->scale);
}

/* ------------------------------------------------------------------------------------ */
/* firpfbch_crcf: critically sampled channelizer                                         */
/* ------------------------------------------------------------------------------------ */

struct orc_firpfbch_s {
    int type;
    unsigned int M, p;
    float *sub;                 /* [M][p] */
    orc_cf *w;                  /* [M][p] windows */
    unsigned int filter_index;  /* analyzer: window the next sample goes to */
    orc_cf *X, *x;              /* transform in / out */
};

/* firpfbch.c:73-142 */
orc_firpfbch orc_firpfbch_create(int type, unsigned int M, unsigned int p, const float *h)
{
    if ((type != ORC_ANALYZER && type != ORC_SYNTHESIZER) || M == 0 || p == 0) return NULL;
    orc_firpfbch q = xmalloc(sizeof *q);
    q->type = type; q->M = M; q->p = p;
    q->sub = polyphase_split(h, M, p);
    q->w = xmalloc((size_t)M * p * sizeof *q->w);
    q->X = xmalloc(M * sizeof *q->X);
    q->x = xmalloc(M * sizeof *q->x);
    orc_firpfbch_reset(q);
    return q;
}

/* firpfbch.c:150-184: 2Mm+1 Kaiser taps at fc = 0.5/M, not normalised, first 2Mm used */
orc_firpfbch orc_firpfbch_create_kaiser(int type, unsigned int M, unsigned int m, float As)
{
    if (M == 0 || m == 0) return NULL;
    As = fabsf(As);
    unsigned int h_len = 2 * M * m + 1;
    float *h = xmalloc(h_len * sizeof *h);
    float fc = 0.5f / (float)M;
    orc_firdes_kaiser(h_len, fc, As, 0.0f, h);
    orc_firpfbch q = orc_firpfbch_create(type, M, 2 * m, h);
    free(h);
    return q;
}

/* firpfbch.c:193-256: root-Nyquist prototype at k = M (ftype 7 = arkaiser, 8 = rkaiser, 9 = RRC; hM3 is
 * not restated), first 2Mm taps, time-reversed for the analyzer */
orc_firpfbch orc_firpfbch_create_rnyquist(int type, unsigned int M, unsigned int m, float beta, int ftype)
{
    if ((type != ORC_ANALYZER && type != ORC_SYNTHESIZER) || M == 0 || m == 0) return NULL;
    unsigned int h_len = 2 * M * m + 1, g_len = 2 * M * m;
    float *h = xmalloc(h_len * sizeof *h);
    float *g = xmalloc(g_len * sizeof *g);
    int rc;
    if (ftype == 7)      rc = orc_firdes_arkaiser(M, m, beta, 0.0f, h);
    else if (ftype == 8) rc = orc_firdes_rkaiser(M, m, beta, 0.0f, h, NULL);
    else if (ftype == 9) rc = orc_firdes_rrcos(M, m, beta, 0.0f, h);
    else                 rc = -1;
    if (rc) { free(h); free(g); return NULL; }
    for (unsigned int i = 0; i < g_len; i++)
        g[i] = (type == ORC_SYNTHESIZER) ? h[i] : h[g_len - i - 1];
    orc_firpfbch q = orc_firpfbch_create(type, M, 2 * m, g);
    free(h); free(g);
    return q;
}

orc_firpfbch orc_firpfbch_create_rrc(int type, unsigned int M, unsigned int m, float beta)
{
    return orc_firpfbch_create_rnyquist(type, M, m, beta, 9);
}

void orc_firpfbch_destroy(orc_firpfbch q)
{
    if (q) { free(q->sub); free(q->w); free(q->X); free(q->x); free(q); }
}

/* firpfbch.c:284-293 */
void orc_firpfbch_reset(orc_firpfbch q)
{
    memset(q->w, 0, (size_t)q->M * q->p * sizeof *q->w);
    q->filter_index = q->M - 1;
}

/* firpfbch.c:369-377 */
void orc_firpfbch_analyzer_push(orc_firpfbch q, orc_cf x)
{
    win_push(q->w + (size_t)q->filter_index * q->p, q->p, x);
    q->filter_index = (q->filter_index + q->M - 1) % q->M;
}

/* firpfbch.c:383-409: branch i reads window (i+k)%M, lands in transform input M-1-i;
 * forward transform, no scaling */
void orc_firpfbch_analyzer_run(orc_firpfbch q, unsigned int k, orc_cf *y)
{
    for (unsigned int i = 0; i < q->M; i++) {
        unsigned int idx = (i + k) % q->M;
        q->X[q->M - i - 1] = dot_crcf(q->sub + (size_t)i * q->p, q->w + (size_t)idx * q->p, q->p);
    }
    orc_fft_run(q->M, q->X, q->x, ORC_FFT_FORWARD);
    memcpy(y, q->x, q->M * sizeof *y);
}

/* firpfbch.c:346-359 */
void orc_firpfbch_analyzer_execute(orc_firpfbch q, const orc_cf *x, orc_cf *y)
{
    for (unsigned int i = 0; i < q->M; i++) orc_firpfbch_analyzer_push(q, x[i]);
    orc_firpfbch_analyzer_run(q, 0, y);
}

/* firpfbch.c:314-336: backward transform (unscaled), then one FIR step per branch */
void orc_firpfbch_synthesizer_execute(orc_firpfbch q, const orc_cf *x, orc_cf *y)
{
    memcpy(q->X, x, q->M * sizeof *q->X);
    orc_fft_run(q->M, q->X, q->x, ORC_FFT_BACKWARD);
    for (unsigned int i = 0; i < q->M; i++) {
        orc_cf *w = q->w + (size_t)i * q->p;
        win_push(w, q->p, q->x[i]);
        y[i] = dot_crcf(q->sub + (size_t)i * q->p, w, q->p);
    }
}

void orc_firpfbch_execute_frames(orc_firpfbch q, const orc_cf *x, unsigned int n_frames, orc_cf *y)
{
    for (unsigned int t = 0; t < n_frames; t++) {
        if (q->type == ORC_ANALYZER) orc_firpfbch_analyzer_execute(q, x + (size_t)t * q->M, y + (size_t)t * q->M);
        else                         orc_firpfbch_synthesizer_execute(q, x + (size_t)t * q->M, y + (size_t)t * q->M);
    }
}

/* ------------------------------------------------------------------------------------ */
/* firpfbch_cccf: the same channelizer with complex taps                                 */
/* src/multichannel/src/firpfbch_cccf.c:35-52 instantiates firpfbch.c with TC = float     */
/* complex and dotprod_cccf (src/dotprod/src/dotprod_cccf.c -> dotprod.c:162-168:         */
/* r += h[i] * x[i], plain complex product, no conjugate)                                 */
/* ------------------------------------------------------------------------------------ */

struct orc_firpfbch_cccf_s {
    int type;
    unsigned int M, p;
    orc_cf *sub;                /* [M][p], newest tap last */
    orc_cf *w;                  /* [M][p] windows */
    unsigned int filter_index;
    orc_cf *X, *x;
};

static orc_cf dot_cccf(const orc_cf *h, const orc_cf *x, unsigned int n)
{
    orc_cf r = cf(0.0f, 0.0f);
    for (unsigned int i = 0; i < n; i++) {
        r.re += h[i].re * x[i].re - h[i].im * x[i].im;
        r.im += h[i].re * x[i].im + h[i].im * x[i].re;
    }
    return r;
}

/* firpfbch.c:73-142 with complex h */
orc_firpfbch_cccf orc_firpfbch_cccf_create(int type, unsigned int M, unsigned int p, const orc_cf *h)
{
    if ((type != ORC_ANALYZER && type != ORC_SYNTHESIZER) || M == 0 || p == 0) return NULL;
    orc_firpfbch_cccf q = xmalloc(sizeof *q);
    q->type = type; q->M = M; q->p = p;
    q->sub = xmalloc((size_t)M * p * sizeof *q->sub);
    for (unsigned int i = 0; i < M; i++)
        for (unsigned int n = 0; n < p; n++)
            q->sub[(size_t)i * p + (p - 1 - n)] = h[i + n * M];      /* firpfbch.c:115-124 */
    q->w = xmalloc((size_t)M * p * sizeof *q->w);
    q->X = xmalloc(M * sizeof *q->X);
    q->x = xmalloc(M * sizeof *q->x);
    memset(q->w, 0, (size_t)M * p * sizeof *q->w);
    q->filter_index = M - 1;
    return q;
}

void orc_firpfbch_cccf_destroy(orc_firpfbch_cccf q)
{
    if (q) { free(q->sub); free(q->w); free(q->X); free(q->x); free(q); }
}

/* firpfbch.c:346-409 (analyzer: M pushes, run with k = 0) and :314-336 (synthesizer) */
void orc_firpfbch_cccf_execute_frames(orc_firpfbch_cccf q, const orc_cf *x, unsigned int n_frames, orc_cf *y)
{
    const unsigned int M = q->M, p = q->p;
    for (unsigned int t = 0; t < n_frames; t++) {
        const orc_cf *xt = x + (size_t)t * M;
        orc_cf *yt = y + (size_t)t * M;
        if (q->type == ORC_ANALYZER) {
            for (unsigned int i = 0; i < M; i++) {
                win_push(q->w + (size_t)q->filter_index * p, p, xt[i]);
                q->filter_index = (q->filter_index + M - 1) % M;
            }
            for (unsigned int i = 0; i < M; i++)
                q->X[M - i - 1] = dot_cccf(q->sub + (size_t)i * p, q->w + (size_t)i * p, p);
            orc_fft_run(M, q->X, q->x, ORC_FFT_FORWARD);
            memcpy(yt, q->x, M * sizeof *yt);
        } else {
            memcpy(q->X, xt, M * sizeof *q->X);
            orc_fft_run(M, q->X, q->x, ORC_FFT_BACKWARD);
            for (unsigned int i = 0; i < M; i++) {
                orc_cf *w = q->w + (size_t)i * p;
                win_push(w, p, q->x[i]);
                yt[i] = dot_cccf(q->sub + (size_t)i * p, w, p);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* firdecim_crcf / firinterp_crcf: the polyphase pattern without a   */
/* transform.  src/filter/src/firdecim.c, firinterp.c                                      */
/* ------------------------------------------------------------------------------------ */

struct orc_firdecim_s {
    unsigned int M, h_len;
    float *hr;                  /* taps reversed: lines up with an oldest-first window (firdecim.c:64-67) */
    orc_cf *w;                  /* window of h_len samples */
};

/* firdecim.c:46-82 */
orc_firdecim orc_firdecim_create(unsigned int M, const float *h, unsigned int h_len)
{
    if (h_len == 0 || M == 0) return NULL;
    orc_firdecim q = xmalloc(sizeof *q);
    q->M = M; q->h_len = h_len;
    q->hr = xmalloc(h_len * sizeof *q->hr);
    for (unsigned int i = 0; i < h_len; i++) q->hr[i] = h[h_len - i - 1];
    q->w = xmalloc(h_len * sizeof *q->w);
    memset(q->w, 0, h_len * sizeof *q->w);
    return q;
}

/* firdecim.c:88-118: 2Mm+1 Kaiser taps at fc = 0.5/M, the first 2Mm used */
orc_firdecim orc_firdecim_create_kaiser(unsigned int M, unsigned int m, float As)
{
    if (M < 2 || m == 0 || As < 0.0f) return NULL;
    unsigned int h_len = 2 * M * m + 1;
    float *h = xmalloc(h_len * sizeof *h);
    orc_firdes_kaiser(h_len, 0.5f / (float)M, As, 0.0f, h);
    orc_firdecim q = orc_firdecim_create(M, h, 2 * M * m);
    free(h);
    return q;
}

void orc_firdecim_destroy(orc_firdecim q) { if (q) { free(q->hr); free(q->w); free(q); } }
void orc_firdecim_clear(orc_firdecim q) { memset(q->w, 0, q->h_len * sizeof *q->w); }   /* firdecim.c:179-182 */

/* firdecim.c:189-223: M pushes per output, the dot product is taken right after the FIRST push */
void orc_firdecim_execute_block(orc_firdecim q, const orc_cf *x, unsigned int n, orc_cf *y)
{
    for (unsigned int i = 0; i < n; i++)
        for (unsigned int j = 0; j < q->M; j++) {
            win_push(q->w, q->h_len, x[(size_t)i * q->M + j]);
            if (j == 0) y[i] = dot_crcf(q->hr, q->w, q->h_len);
        }
}

struct orc_firinterp_s {
    unsigned int M, sub_len;
    float *sub;                 /* [M][sub_len], newest tap last (the firpfb layout, firpfb.c:71-84) */
    orc_cf *w;                  /* window of sub_len samples */
};

/* firinterp.c:43-80: prototype zero-padded to M * ceil(h_len / M), split into the M branches of a firpfb */
orc_firinterp orc_firinterp_create(unsigned int M, const float *h, unsigned int h_len)
{
    if (M < 2 || h_len < M) return NULL;
    orc_firinterp q = xmalloc(sizeof *q);
    q->M = M;
    q->sub_len = 0;
    while (M * q->sub_len < h_len) q->sub_len++;
    unsigned int H = M * q->sub_len;
    float *hp = xmalloc(H * sizeof *hp);
    for (unsigned int i = 0; i < H; i++) hp[i] = i < h_len ? h[i] : 0.0f;
    q->sub = polyphase_split(hp, M, q->sub_len);
    free(hp);
    q->w = xmalloc(q->sub_len * sizeof *q->w);
    memset(q->w, 0, q->sub_len * sizeof *q->w);
    return q;
}

/* firinterp.c:86-116 */
orc_firinterp orc_firinterp_create_kaiser(unsigned int M, unsigned int m, float As)
{
    if (M < 2 || m == 0 || As < 0.0f) return NULL;
    unsigned int h_len = 2 * M * m + 1;
    float *h = xmalloc(h_len * sizeof *h);
    orc_firdes_kaiser(h_len, 0.5f / (float)M, As, 0.0f, h);
    orc_firinterp q = orc_firinterp_create(M, h, 2 * M * m);
    free(h);
    return q;
}

void orc_firinterp_destroy(orc_firinterp q) { if (q) { free(q->sub); free(q->w); free(q); } }
void orc_firinterp_reset(orc_firinterp q) { memset(q->w, 0, q->sub_len * sizeof *q->w); }   /* firinterp.c:178-181 */

/* firinterp.c:187-217: one push, then the M branch outputs */
void orc_firinterp_execute_block(orc_firinterp q, const orc_cf *x, unsigned int n, orc_cf *y)
{
    for (unsigned int i = 0; i < n; i++) {
        win_push(q->w, q->sub_len, x[i]);
        for (unsigned int j = 0; j < q->M; j++)
            y[(size_t)i * q->M + j] = dot_crcf(q->sub + (size_t)j * q->sub_len, q->w, q->sub_len);
    }
}

/* ------------------------------------------------------------------------------------ */
/* resamp2_crcf / msresamp2_crcf: half-band filter, 2:1 resamplers    */
/* and their 2^S cascades.  src/filter/src/resamp2.c, msresamp2.c                           */
/* ------------------------------------------------------------------------------------ */

struct orc_resamp2_s {
    unsigned int m, len;        /* len = 2m taps on the filter branch */
    float *hb;                  /* filter-branch taps, lined up with an oldest-first window */
    orc_cf *w0, *w1;            /* delay-branch and filter-branch windows, len samples each */
    unsigned int toggle;
};

/* firdes.c:52-75 (Kaiser's estimate, the build's method 0) */
unsigned int orc_estimate_req_filter_len(float df, float As)
{
    return (unsigned int)((As - 7.95f) / (14.26f * df));
}

/* resamp2.c:65-123: 4m+1 tap prototype sinc(t/2) * kaiser * cos(2 pi f0 t); the filter branch keeps the
 * odd-indexed taps, reversed */
orc_resamp2 orc_resamp2_create(unsigned int m, float f0, float As)
{
    if (m < 2 || f0 < -0.5f || f0 > 0.5f) return NULL;
    orc_resamp2 q = xmalloc(sizeof *q);
    q->m = m; q->len = 2 * m;
    unsigned int n = 4 * m + 1;
    float *h = xmalloc(n * sizeof *h);
    float beta = orc_kaiser_beta_As(As);
    for (unsigned int i = 0; i < n; i++) {
        float t = (float)i - (float)(n - 1) / 2.0f;
        float a = orc_sincf(t / 2.0f);
        float b = orc_kaiser(i, n, beta, 0);
        float c = cosf(2.0f * M_PI * t * f0);
        h[i] = a * b * c;
    }
    q->hb = xmalloc(q->len * sizeof *q->hb);
    for (unsigned int j = 0, i = 1; i < n; i += 2) q->hb[j++] = h[n - i - 1];
    free(h);
    q->w0 = xmalloc(q->len * sizeof *q->w0);
    q->w1 = xmalloc(q->len * sizeof *q->w1);
    orc_resamp2_clear(q);
    return q;
}

void orc_resamp2_destroy(orc_resamp2 q) { if (q) { free(q->hb); free(q->w0); free(q->w1); free(q); } }

/* resamp2.c:209-215 */
void orc_resamp2_clear(orc_resamp2 q)
{
    memset(q->w0, 0, q->len * sizeof *q->w0);
    memset(q->w1, 0, q->len * sizeof *q->w1);
    q->toggle = 0;
}

void orc_resamp2_get_taps(orc_resamp2 q, float *hb) { memcpy(hb, q->hb, q->len * sizeof *hb); }

static orc_cf cf_add(orc_cf a, orc_cf b) { return cf(a.re + b.re, a.im + b.im); }
static orc_cf cf_sub(orc_cf a, orc_cf b) { return cf(a.re - b.re, a.im - b.im); }
static orc_cf cf_half(orc_cf a) { return cf(0.5f * a.re, 0.5f * a.im); }

/* resamp2.c:228-264: n samples through the half-band filter -> low band y0[n], high band y1[n] */
void orc_resamp2_filter_execute(orc_resamp2 q, const orc_cf *x, unsigned int n, orc_cf *y0, orc_cf *y1)
{
    for (unsigned int i = 0; i < n; i++) {
        orc_cf *wp = q->toggle ? q->w1 : q->w0;     /* the window this sample joins; the other one is filtered */
        orc_cf *wf = q->toggle ? q->w0 : q->w1;
        win_push(wp, q->len, x[i]);
        orc_cf d = wp[q->m - 1];
        orc_cf f = dot_crcf(q->hb, wf, q->len);
        q->toggle = 1 - q->toggle;
        y0[i] = cf_half(cf_add(d, f));
        y1[i] = cf_half(cf_sub(d, f));
    }
}

/* resamp2.c:271-290: n input pairs -> n (low, high) pairs at half rate */
void orc_resamp2_analyzer_execute(orc_resamp2 q, const orc_cf *x, unsigned int n, orc_cf *y)
{
    for (unsigned int i = 0; i < n; i++) {
        win_push(q->w1, q->len, cf_half(x[2 * i]));
        orc_cf f = dot_crcf(q->hb, q->w1, q->len);
        win_push(q->w0, q->len, cf_half(x[2 * i + 1]));
        orc_cf d = q->w0[q->m - 1];
        y[2 * i] = cf_add(f, d);
        y[2 * i + 1] = cf_sub(f, d);
    }
}

/* resamp2.c:297-313 */
void orc_resamp2_synthesizer_execute(orc_resamp2 q, const orc_cf *x, unsigned int n, orc_cf *y)
{
    for (unsigned int i = 0; i < n; i++) {
        win_push(q->w0, q->len, cf_add(x[2 * i], x[2 * i + 1]));
        y[2 * i] = q->w0[q->m - 1];
        win_push(q->w1, q->len, cf_sub(x[2 * i], x[2 * i + 1]));
        y[2 * i + 1] = dot_crcf(q->hb, q->w1, q->len);
    }
}

/* resamp2.c:320-338: n outputs from 2n inputs */
void orc_resamp2_decim_execute(orc_resamp2 q, const orc_cf *x, unsigned int n, orc_cf *y)
{
    for (unsigned int i = 0; i < n; i++) {
        win_push(q->w1, q->len, x[2 * i]);
        orc_cf f = dot_crcf(q->hb, q->w1, q->len);
        win_push(q->w0, q->len, x[2 * i + 1]);
        y[i] = cf_add(q->w0[q->m - 1], f);
    }
}

/* resamp2.c:345-359: 2n outputs from n inputs */
void orc_resamp2_interp_execute(orc_resamp2 q, const orc_cf *x, unsigned int n, orc_cf *y)
{
    for (unsigned int i = 0; i < n; i++) {
        win_push(q->w0, q->len, x[i]);
        y[2 * i] = q->w0[q->m - 1];
        win_push(q->w1, q->len, x[i]);
        y[2 * i + 1] = dot_crcf(q->hb, q->w1, q->len);
    }
}

struct orc_msresamp2_s {
    int interp;
    unsigned int S;             /* stages; rate 2^S */
    unsigned int *m;            /* semi-length of each stage */
    orc_resamp2 *hb;
};

/* msresamp2.c:82-170: stage i has cut-off fc/2^(i+1), transition (0.5 - fc_i)/2, semi-length from
 * Kaiser's length estimate (at least 3) */
orc_msresamp2 orc_msresamp2_create(int type, unsigned int S, float fc, float f0, float As)
{
    if (S > 16 || fc <= 0.0f || fc >= 0.5f) return NULL;
    if (fc > 0.45f) fc = 0.45f;
    f0 = 0.0f;                                   /* msresamp2.c:108-111: centre frequency unsupported */
    orc_msresamp2 q = xmalloc(sizeof *q);
    q->interp = (type == 0);                     /* LIQUID_RESAMP_INTERP = 0 */
    q->S = S;
    q->m = xmalloc((S ? S : 1) * sizeof *q->m);
    q->hb = xmalloc((S ? S : 1) * sizeof *q->hb);
    for (unsigned int i = 0; i < S; i++) {
        f0 = 0.5f * f0;
        fc = 0.5f * fc;
        float ft = (0.5f - fc) / 2.0f;
        unsigned int h_len = orc_estimate_req_filter_len(ft, As);
        unsigned int m = (unsigned int)ceilf((float)(h_len - 1) / 4.0f);
        q->m[i] = m < 3 ? 3 : m;
        q->hb[i] = orc_resamp2_create(q->m[i], f0, As);
    }
    return q;
}

void orc_msresamp2_destroy(orc_msresamp2 q)
{
    if (!q) return;
    for (unsigned int i = 0; i < q->S; i++) orc_resamp2_destroy(q->hb[i]);
    free(q->hb); free(q->m); free(q);
}

void orc_msresamp2_reset(orc_msresamp2 q) { for (unsigned int i = 0; i < q->S; i++) orc_resamp2_clear(q->hb[i]); }
unsigned int orc_msresamp2_stage_m(orc_msresamp2 q, unsigned int i) { return q->m[i]; }

/* msresamp2.c:231-258 */
float orc_msresamp2_get_delay(orc_msresamp2 q)
{
    float delay = 0;
    for (unsigned int i = 0; i < q->S; i++) {
        if (q->interp) { delay *= 0.5f; delay += q->m[i]; }
        else           { delay *= 2; delay += 2 * q->m[q->S - i - 1] - 1; }
    }
    return delay;
}

/* msresamp2.c:264-354 over a block of n units (interp: n inputs -> n 2^S outputs, stages last to first;
 * decim: n 2^S inputs -> n outputs scaled by 2^-S, stages first to last).  The stages only see their own
 * sample streams, so running each stage over the whole block equals the reference's unit-by-unit order. */
void orc_msresamp2_execute_block(orc_msresamp2 q, const orc_cf *x, unsigned int n, orc_cf *y)
{
    unsigned int M = 1u << q->S;
    if (q->S == 0) { memcpy(y, x, n * sizeof *y); return; }
    size_t cap = (size_t)n * M;
    orc_cf *a = xmalloc(cap * sizeof *a), *b = xmalloc(cap * sizeof *b);
    if (q->interp) {
        memcpy(a, x, n * sizeof *a);
        size_t cnt = n;
        for (unsigned int s = 0; s < q->S; s++) {
            orc_resamp2_interp_execute(q->hb[q->S - s - 1], a, (unsigned int)cnt, b);
            cnt *= 2;
            orc_cf *t = a; a = b; b = t;
        }
        memcpy(y, a, cnt * sizeof *y);
    } else {
        memcpy(a, x, cap * sizeof *a);
        size_t cnt = cap;
        for (unsigned int s = 0; s < q->S; s++) {
            cnt /= 2;
            orc_resamp2_decim_execute(q->hb[s], a, (unsigned int)cnt, b);
            orc_cf *t = a; a = b; b = t;
        }
        float zeta = 1.0f / (float)M;
        for (unsigned int i = 0; i < n; i++) y[i] = cf(a[i].re * zeta, a[i].im * zeta);
    }
    free(a); free(b);
}

/* ------------------------------------------------------------------------------------ */
/* fftfilt_crcf: FIR filtering by overlap-add with 2n-point transforms  */
/* src/filter/src/fftfilt.c                                                                  */
/* ------------------------------------------------------------------------------------ */

struct orc_fftfilt_s {
    unsigned int h_len, n;
    orc_cf *H;                  /* transform of the zero-padded taps, 2n bins */
    orc_cf *tail;               /* second half of the previous block's result */
    orc_cf *tb, *fb;            /* time / frequency work arrays, 2n each */
    float scale;
};

/* fftfilt.c:68-129 */
orc_fftfilt orc_fftfilt_create(const float *h, unsigned int h_len, unsigned int n)
{
    if (h_len == 0 || n < h_len - 1) return NULL;
    orc_fftfilt q = xmalloc(sizeof *q);
    q->h_len = h_len; q->n = n;
    unsigned int N = 2 * n;
    q->H = xmalloc(N * sizeof *q->H);
    q->tb = xmalloc(N * sizeof *q->tb);
    q->fb = xmalloc(N * sizeof *q->fb);
    q->tail = xmalloc(n * sizeof *q->tail);
    for (unsigned int i = 0; i < N; i++) q->tb[i] = cf(i < h_len ? h[i] : 0.0f, 0.0f);
    orc_fft_run(N, q->tb, q->H, ORC_FFT_FORWARD);
    orc_fftfilt_set_scale(q, 1.0f);
    orc_fftfilt_reset(q);
    return q;
}

void orc_fftfilt_destroy(orc_fftfilt q) { if (q) { free(q->H); free(q->tb); free(q->fb); free(q->tail); free(q); } }
void orc_fftfilt_reset(orc_fftfilt q) { memset(q->tail, 0, q->n * sizeof *q->tail); }            /* fftfilt.c:155-161 */
void orc_fftfilt_set_scale(orc_fftfilt q, float scale) { q->scale = scale / (float)(2 * q->n); }  /* fftfilt.c:184-189 */

/* fftfilt.c:195-260 over n_blocks blocks of n samples */
void orc_fftfilt_execute_blocks(orc_fftfilt q, const orc_cf *x, unsigned int n_blocks, orc_cf *y)
{
    unsigned int n = q->n, N = 2 * q->n;
    for (unsigned int b = 0; b < n_blocks; b++) {
        for (unsigned int i = 0; i < n; i++) { q->tb[i] = x[(size_t)b * n + i]; q->tb[n + i] = cf(0.0f, 0.0f); }
        orc_fft_run(N, q->tb, q->fb, ORC_FFT_FORWARD);
        for (unsigned int i = 0; i < N; i++) q->fb[i] = cf_mul(q->fb[i], q->H[i]);
        orc_fft_run(N, q->fb, q->tb, ORC_FFT_BACKWARD);
        for (unsigned int i = 0; i < n; i++) {
            orc_cf v = cf(q->tb[i].re + q->tail[i].re, q->tb[i].im + q->tail[i].im);
            y[(size_t)b * n + i] = cf(v.re * q->scale, v.im * q->scale);
        }
        memcpy(q->tail, q->tb + n, n * sizeof *q->tail);
    }
}

/* ------------------------------------------------------------------------------------ */
/* spgramcf: windowed-FFT spectral periodogram                      */
/* src/fft/src/spgramcf.c:27-51 instantiating src/fft/src/spgram.c                        */
/* ------------------------------------------------------------------------------------ */

struct orc_spgramcf_s {
    unsigned int nfft, window_len;
    orc_cf *buf;                /* last window_len samples, oldest first */
    float *w;                   /* scaled tapering window */
    orc_cf *x, *X;              /* transform in / out */
    float *psd;                 /* accumulation */
    unsigned int sample_counter, num_transforms;
};

/* spgram.c:180-199 */
void orc_spgramcf_reset(orc_spgramcf q)
{
    memset(q->buf, 0, q->window_len * sizeof *q->buf);
    memset(q->x, 0, q->nfft * sizeof *q->x);
    q->num_transforms = 0;
    q->sample_counter = 0;
    for (unsigned int i = 0; i < q->nfft; i++) q->psd[i] = 1;
}

/* spgram.c:57-107: the window is scaled by sqrt(2) / (rms(window) sqrt(nfft)) */
orc_spgramcf orc_spgramcf_create(unsigned int nfft, const float *window, unsigned int window_len)
{
    if (nfft < 2 || window_len > nfft || window_len == 0) return NULL;
    orc_spgramcf q = xmalloc(sizeof *q);
    q->nfft = nfft;
    q->window_len = window_len;
    q->x = xmalloc(nfft * sizeof *q->x);
    q->X = xmalloc(nfft * sizeof *q->X);
    q->psd = xmalloc(nfft * sizeof *q->psd);
    q->buf = xmalloc(window_len * sizeof *q->buf);
    q->w = xmalloc(window_len * sizeof *q->w);
    float g = 0.0f;
    for (unsigned int i = 0; i < window_len; i++) g += window[i] * window[i];
    g = M_SQRT2 / (sqrtf(g / window_len) * sqrtf((float)nfft));
    for (unsigned int i = 0; i < window_len; i++) q->w[i] = g * window[i];
    orc_spgramcf_reset(q);
    return q;
}

/* spgram.c:113-147: Kaiser-Bessel taper kaiser(i, window_len, beta, 0) */
orc_spgramcf orc_spgramcf_create_kaiser(unsigned int nfft, unsigned int window_len, float beta)
{
    if (nfft < 2 || window_len > nfft || window_len == 0 || beta <= 0.0f) return NULL;
    float *w = xmalloc(window_len * sizeof *w);
    for (unsigned int i = 0; i < window_len; i++) w[i] = orc_kaiser(i, window_len, beta, 0.0f);
    orc_spgramcf q = orc_spgramcf_create(nfft, w, window_len);
    free(w);
    return q;
}

/* spgram.c:150-162: window nfft/2, beta 10 */
orc_spgramcf orc_spgramcf_create_default(unsigned int nfft)
{
    if (nfft < 2) return NULL;
    return orc_spgramcf_create_kaiser(nfft, nfft / 2, 10.0f);
}

void orc_spgramcf_destroy(orc_spgramcf q)
{
    if (q) { free(q->x); free(q->X); free(q->psd); free(q->buf); free(q->w); free(q); }
}

void orc_spgramcf_push(orc_spgramcf q, orc_cf v) { win_push(q->buf, q->window_len, v); }   /* spgram.c:202-207 */

void orc_spgramcf_write(orc_spgramcf q, const orc_cf *x, unsigned int n)                    /* spgram.c:213-219 */
{
    for (unsigned int i = 0; i < n; i++) win_push(q->buf, q->window_len, x[i]);
}

/* spgram.c:226-244: taper the buffer into the first window_len transform inputs (the rest stay
 * zero), forward transform */
void orc_spgramcf_execute(orc_spgramcf q, orc_cf *X)
{
    for (unsigned int i = 0; i < q->window_len; i++)
        q->x[i] = cf(q->buf[i].re * q->w[i], q->buf[i].im * q->w[i]);
    orc_fft_run(q->nfft, q->x, q->X, ORC_FFT_FORWARD);
    if (X) memmove(X, q->X, q->nfft * sizeof *X);
}

static float magsq(orc_cf v) { return v.re * v.re + v.im * v.im; }   /* crealf(X * conjf(X)) */

/* spgram.c:250-264: fft-shifted dB */
void orc_spgramcf_execute_psd(orc_spgramcf q, float *out)
{
    orc_spgramcf_execute(q, NULL);
    unsigned int h = q->nfft / 2;
    for (unsigned int i = 0; i < q->nfft; i++) out[(i + h) % q->nfft] = 10 * log10f(magsq(q->X[i]) + 1e-16f);
}

/* spgram.c:271-311: a transform every window_len/2 pushes, psd <- (1-alpha) psd + alpha |X|^2; the
 * first transform after a reset overrides alpha with 1 -- and, alpha being the function's own
 * parameter, it stays 1 for the rest of THAT call */
int orc_spgramcf_accumulate_psd(orc_spgramcf q, const orc_cf *x, float alpha, unsigned int n)
{
    if (alpha < 0.0f || alpha > 1.0f) return -1;
    for (unsigned int i = 0; i < n; i++) {
        win_push(q->buf, q->window_len, x[i]);
        q->sample_counter++;
        if (q->sample_counter == q->window_len / 2) {
            if (q->num_transforms == 0) alpha = 1.0f;
            orc_spgramcf_execute(q, NULL);
            for (unsigned int k = 0; k < q->nfft; k++)
                q->psd[k] = (1.0f - alpha) * q->psd[k] + alpha * magsq(q->X[k]);
            q->sample_counter = 0;
            q->num_transforms++;
        }
    }
    return 0;
}

/* spgram.c:318-334 */
void orc_spgramcf_write_accumulation(orc_spgramcf q, float *out)
{
    unsigned int h = q->nfft / 2;
    for (unsigned int i = 0; i < q->nfft; i++) out[(i + h) % q->nfft] = 10 * log10f(q->psd[i]);
}

/* spgram.c:341-398: reset, a transform every nfft/4 samples (and after the last sample), plain
 * average of |X|^2, fft-shifted dB */
void orc_spgramcf_estimate_psd(orc_spgramcf q, const orc_cf *x, unsigned int n, float *psd)
{
    if (n == 0) return;
    orc_spgramcf_reset(q);
    unsigned int delay = q->nfft / 4;
    if (delay == 0) delay = 1;
    for (unsigned int i = 0; i < q->nfft; i++) psd[i] = 0.0f;
    unsigned int num = 0;
    for (unsigned int i = 0; i < n; i++) {
        win_push(q->buf, q->window_len, x[i]);
        if (((i + 1) % delay) == 0 || i == n - 1) {
            orc_spgramcf_execute(q, NULL);
            for (unsigned int k = 0; k < q->nfft; k++) psd[(k + q->nfft / 2) % q->nfft] += magsq(q->X[k]);
            num++;
        }
    }
    for (unsigned int i = 0; i < q->nfft; i++) psd[i] = 10 * log10f(psd[i] / (float)num);
}

unsigned int orc_spgramcf_num_transforms(orc_spgramcf q) { return q->num_transforms; }

/* asgramcf: ASCII spectrogram over spgramcf.  src/fft/src/asgram.c:48-195 */
struct orc_asgramcf_s {
    unsigned int nfft;
    orc_spgramcf sp;
    float levels[10];
};

orc_asgramcf orc_asgramcf_create(unsigned int nfft)
{
    if (nfft < 2) return NULL;
    orc_asgramcf q = xmalloc(sizeof *q);
    q->nfft = nfft;
    q->sp = orc_spgramcf_create_kaiser(nfft, nfft, 10.0f);          /* asgram.c:66-69 */
    orc_asgramcf_set_scale(q, 0.0f, 10.0f);
    return q;
}

void orc_asgramcf_destroy(orc_asgramcf q) { if (q) { orc_spgramcf_destroy(q->sp); free(q); } }
void orc_asgramcf_reset(orc_asgramcf q) { orc_spgramcf_reset(q->sp); }
void orc_asgramcf_set_scale(orc_asgramcf q, float offset, float scale)       /* asgram.c:113-128 */
{
    for (unsigned int i = 0; i < 10; i++) q->levels[i] = offset + i * scale;
}
void orc_asgramcf_write(orc_asgramcf q, const orc_cf *x, unsigned int n) { orc_spgramcf_write(q->sp, x, n); }

/* asgram.c:157-195: shifted 10 log10 |X| (the magnitude), level characters, first maximum */
void orc_asgramcf_execute(orc_asgramcf q, char *ascii, float *peakval, float *peakfreq)
{
    static const char lc[10] = {' ', '.', ',', '-', '+', '*', '&', 'N', 'M', '#'};
    orc_cf *X = xmalloc(q->nfft * sizeof *X);
    orc_spgramcf_execute(q->sp, X);
    for (unsigned int i = 0; i < q->nfft; i++) {
        orc_cf v = X[(i + q->nfft / 2) % q->nfft];
        float p = 10 * log10f(hypotf(v.re, v.im));
        if (i == 0 || p > *peakval) { *peakval = p; *peakfreq = (float)i / (float)q->nfft - 0.5f; }
        ascii[i] = lc[0];
        for (unsigned int j = 0; j < 10; j++)
            if (p > q->levels[j]) ascii[i] = lc[j];
    }
    free(X);
}

/* ------------------------------------------------------------------------------------ */
/* ofdmframegen: OFDM symbol generator -- subcarrier allocation,    */
/* PLCP sequences, IFFT, cyclic prefix and tapered overlap.                               */
/* src/multichannel/src/ofdmframe.common.c, ofdmframegen.c, src/sequence/src/msequence.c  */
/* ------------------------------------------------------------------------------------ */

/* default m-sequence of register length m (msequence.c:41-58: generator polynomial shifted right by one, initial
 * state 1 << (m-1)); advance (:166-177): output bit = parity(v & g), pushed into the register */
struct mseq { unsigned int m, g, n, v; };
static const unsigned int mseq_poly[16] = {0, 0, 0x0003, 0x0005, 0x0009, 0x0012, 0x0021, 0x0044, 0x008E,
                                           0x0108, 0x0204, 0x0402, 0x0829, 0x100d, 0x2015, 0x4001};
static struct mseq mseq_default(unsigned int m)
{
    struct mseq s = {m, mseq_poly[m], (1u << m) - 1, 1u << (m - 1)};
    return s;
}
static unsigned int mseq_advance(struct mseq *s)
{
    unsigned int b = (unsigned int)__builtin_parity(s->v & s->g);
    s->v = ((s->v << 1) | b) & s->n;
    return b;
}
static unsigned int mseq_symbol(struct mseq *s, unsigned int bps)
{
    unsigned int v = 0;
    for (unsigned int i = 0; i < bps; i++) v = (v << 1) | mseq_advance(s);
    return v;
}

static unsigned int nextpow2(unsigned int x)          /* math.c:143-157 */
{
    unsigned int n = 0;
    x--;
    while (x > 0) { x >>= 1; n++; }
    return n;
}

/* ofdmframe.common.c:176-214 */
void orc_ofdmframe_init_default_sctype(unsigned int M, unsigned char *p)
{
    unsigned int M2 = M / 2, G = M / 10;
    if (G < 2) G = 2;
    unsigned int P = (M > 34) ? 8 : 4, P2 = P / 2;
    for (unsigned int i = 0; i < M; i++) p[i] = 0;
    for (unsigned int i = 1; i + G < M2; i++) {
        unsigned char t = (((i + P2) % P) == 0) ? 1 : 2;
        p[i] = t;
        p[M - i] = t;
    }
}

/* ofdmframe.common.c:41-101 (which = 0: short sequence, even enabled subcarriers only) and :109-167 (which = 1: long
 * sequence, register one longer): +-1 from bit 0 of 3-bit symbols, backward transform, scaled by 1/sqrt(enabled) */
static unsigned int ofdm_plcp(const unsigned char *p, unsigned int M, int which, orc_cf *S, orc_cf *s)
{
    unsigned int m = nextpow2(M);
    if (m < 4) m = 4; else if (m > 8) m = 8;
    if (which) m++;
    struct mseq ms = mseq_default(m);
    unsigned int enabled = 0;
    for (unsigned int i = 0; i < M; i++) {
        unsigned int b = mseq_symbol(&ms, 3) & 1;
        if (p[i] == 0 || (!which && (i % 2))) S[i] = cf(0.0f, 0.0f);
        else { S[i] = cf(b ? 1.0f : -1.0f, 0.0f); enabled++; }
    }
    orc_fft_run(M, S, s, ORC_FFT_BACKWARD);
    float g = 1.0f / sqrtf(enabled);
    for (unsigned int i = 0; i < M; i++) s[i] = cf(s[i].re * g, s[i].im * g);
    return enabled;
}

struct orc_ofdmframegen_s {
    unsigned int M, cp, tl, n_pilot, n_data;
    unsigned char *p;
    float *taper, g_data;
    orc_cf *postfix, *X, *x, *S0, *s0, *S1, *s1;
    struct mseq pilot;
};

/* ofdmframegen.c:85-171 */
orc_ofdmframegen orc_ofdmframegen_create(unsigned int M, unsigned int cp, unsigned int tl, const unsigned char *p)
{
    if (M < 2 || (M % 2) || cp > M || tl > cp) return NULL;
    orc_ofdmframegen q = xmalloc(sizeof *q);
    q->M = M; q->cp = cp; q->tl = tl;
    q->p = xmalloc(M);
    if (p) memcpy(q->p, p, M); else orc_ofdmframe_init_default_sctype(M, q->p);
    q->n_pilot = q->n_data = 0;
    for (unsigned int i = 0; i < M; i++) { q->n_pilot += q->p[i] == 1; q->n_data += q->p[i] == 2; }
    if (q->n_data == 0 || q->n_pilot < 2) { free(q->p); free(q); return NULL; }
    q->X = xmalloc(M * sizeof(orc_cf)); q->x = xmalloc(M * sizeof(orc_cf));
    q->S0 = xmalloc(M * sizeof(orc_cf)); q->s0 = xmalloc(M * sizeof(orc_cf));
    q->S1 = xmalloc(M * sizeof(orc_cf)); q->s1 = xmalloc(M * sizeof(orc_cf));
    ofdm_plcp(q->p, M, 0, q->S0, q->s0);
    ofdm_plcp(q->p, M, 1, q->S1, q->s1);
    q->taper = xmalloc((tl ? tl : 1) * sizeof(float));
    q->postfix = xmalloc((tl ? tl : 1) * sizeof(orc_cf));
    for (unsigned int i = 0; i < tl; i++) {
        float t = ((float)i + 0.5f) / (float)tl;
        float g = sinf(M_PI_2 * t);
        q->taper[i] = g * g;
    }
    q->g_data = 1.0f / sqrtf(q->n_pilot + q->n_data);
    orc_ofdmframegen_reset(q);
    return q;
}

void orc_ofdmframegen_destroy(orc_ofdmframegen q)
{
    if (!q) return;
    free(q->p); free(q->taper); free(q->postfix); free(q->X); free(q->x);
    free(q->S0); free(q->s0); free(q->S1); free(q->s1); free(q);
}

/* ofdmframegen.c:214-222 */
void orc_ofdmframegen_reset(orc_ofdmframegen q)
{
    q->pilot = mseq_default(8);
    memset(q->postfix, 0, q->tl * sizeof(orc_cf));
}

void orc_ofdmframegen_get_sctype(orc_ofdmframegen q, unsigned char *p) { memcpy(p, q->p, q->M); }

/* ofdmframegen.c:345-361: cyclic prefix, taper against the previous symbol's postfix, keep the new postfix */
static void ofdm_gensymbol(orc_ofdmframegen q, orc_cf *y)
{
    memcpy(y, q->x + (q->M - q->cp), q->cp * sizeof(orc_cf));
    memcpy(y + q->cp, q->x, q->M * sizeof(orc_cf));
    for (unsigned int i = 0; i < q->tl; i++) {
        float a = q->taper[i], b = q->taper[q->tl - i - 1];
        y[i] = cf(y[i].re * a, y[i].im * a);
        y[i] = cf(y[i].re + q->postfix[i].re * b, y[i].im + q->postfix[i].im * b);
    }
    memcpy(q->postfix, q->x, q->tl * sizeof(orc_cf));
}

/* ofdmframegen.c:237-250 */
void orc_ofdmframegen_write_S0a(orc_ofdmframegen q, orc_cf *y)
{
    for (unsigned int i = 0; i < q->M + q->cp; i++) y[i] = q->s0[(i + q->M - 2 * q->cp) % q->M];
    for (unsigned int i = 0; i < q->tl; i++) y[i] = cf(y[i].re * q->taper[i], y[i].im * q->taper[i]);
}

/* ofdmframegen.c:252-264 */
void orc_ofdmframegen_write_S0b(orc_ofdmframegen q, orc_cf *y)
{
    for (unsigned int i = 0; i < q->M + q->cp; i++) y[i] = q->s0[(i + q->M - q->cp) % q->M];
    memcpy(q->postfix, q->s0, q->tl * sizeof(orc_cf));
}

/* ofdmframegen.c:266-272 */
void orc_ofdmframegen_write_S1(orc_ofdmframegen q, orc_cf *y)
{
    memcpy(q->x, q->s1, q->M * sizeof(orc_cf));
    ofdm_gensymbol(q, y);
}

/* ofdmframegen.c:279-311 over n symbols: x [n][M] -> y [n][M + cp] */
void orc_ofdmframegen_writesymbols(orc_ofdmframegen q, const orc_cf *x, unsigned int n, orc_cf *y)
{
    for (unsigned int s = 0; s < n; s++) {
        const orc_cf *xs = x + (size_t)s * q->M;
        for (unsigned int i = 0; i < q->M; i++) {
            unsigned int k = (i + q->M / 2) % q->M;
            if (q->p[k] == 0) q->X[k] = cf(0.0f, 0.0f);
            else if (q->p[k] == 1) q->X[k] = cf((mseq_advance(&q->pilot) ? 1.0f : -1.0f) * q->g_data, 0.0f);
            else q->X[k] = cf(xs[k].re * q->g_data, xs[k].im * q->g_data);
        }
        orc_fft_run(q->M, q->X, q->x, ORC_FFT_BACKWARD);
        ofdm_gensymbol(q, y + (size_t)s * (q->M + q->cp));
    }
}

/* ofdmframegen.c:314-322 */
void orc_ofdmframegen_writetail(orc_ofdmframegen q, orc_cf *y)
{
    for (unsigned int i = 0; i < q->tl; i++) {
        float b = q->taper[q->tl - i - 1];
        y[i] = cf(q->postfix[i].re * b, q->postfix[i].im * b);
    }
}

/* ------------------------------------------------------------------------------------ */
/* firpfbch2_crcf: channelizer with output rate 2 Fs / M                                 */
/* ------------------------------------------------------------------------------------ */

struct orc_firpfbch2_s {
    int type;
    unsigned int M, M2, m, len;     /* len = 2m taps per branch */
    float *sub;                     /* [M][len] */
    orc_cf *w0, *w1;                /* [M][len] windows (w1: synthesizer only) */
    orc_cf *X, *x;
    int flag;                       /* frame parity */
};

/* firpfbch2.c:66-128 */
orc_firpfbch2 orc_firpfbch2_create(int type, unsigned int M, unsigned int m, const float *h)
{
    if ((type != ORC_ANALYZER && type != ORC_SYNTHESIZER) || M < 2 || (M % 2) || m < 1) return NULL;
    orc_firpfbch2 q = xmalloc(sizeof *q);
    q->type = type; q->M = M; q->M2 = M / 2; q->m = m; q->len = 2 * m;
    q->sub = polyphase_split(h, M, q->len);
    q->w0 = xmalloc((size_t)M * q->len * sizeof *q->w0);
    q->w1 = xmalloc((size_t)M * q->len * sizeof *q->w1);
    q->X = xmalloc(M * sizeof *q->X);
    q->x = xmalloc(M * sizeof *q->x);
    orc_firpfbch2_reset(q);
    return q;
}

/* firpfbch2.c:135-183: 2Mm+1 Kaiser taps, fc = 1/M (analyzer) or 0.5/M (synthesizer),
 * scaled so that the float32 running sum equals M, last tap unused */
orc_firpfbch2 orc_firpfbch2_create_kaiser(int type, unsigned int M, unsigned int m, float As)
{
    if ((type != ORC_ANALYZER && type != ORC_SYNTHESIZER) || M < 2 || (M % 2) || m < 1) return NULL;
    unsigned int h_len = 2 * M * m + 1;
    float *hf = xmalloc(h_len * sizeof *hf);
    float fc = (type == ORC_ANALYZER) ? 1.0f / (float)M : 0.5f / (float)M;
    orc_firdes_kaiser(h_len, fc, As, 0.0f, hf);
    float hf_sum = 0.0f;
    for (unsigned int i = 0; i < h_len; i++) hf_sum += hf[i];
    for (unsigned int i = 0; i < h_len; i++) hf[i] = hf[i] * (float)M / hf_sum;
    orc_firpfbch2 q = orc_firpfbch2_create(type, M, m, hf);
    free(hf);
    return q;
}

void orc_firpfbch2_destroy(orc_firpfbch2 q)
{
    if (q) { free(q->sub); free(q->w0); free(q->w1); free(q->X); free(q->x); free(q); }
}

/* firpfbch2.c:213-225 */
void orc_firpfbch2_reset(orc_firpfbch2 q)
{
    memset(q->w0, 0, (size_t)q->M * q->len * sizeof *q->w0);
    memset(q->w1, 0, (size_t)q->M * q->len * sizeof *q->w1);
    q->flag = 0;
}

/* firpfbch2.c:244-282 */
static void pfbch2_analyze(orc_firpfbch2 q, const orc_cf *x, orc_cf *y)
{
    unsigned int M = q->M, M2 = q->M2, len = q->len;
    /* M/2 new samples go into the lower (even frame) or upper (odd frame) half of the
     * windows, walking downwards */
    unsigned int base = q->flag ? M : M2;
    for (unsigned int i = 0; i < M2; i++)
        win_push(q->w0 + (size_t)(base - i - 1) * len, len, x[i]);
    /* branch i meets window (offset+i)%M; the result sits at the window's index */
    unsigned int offset = q->flag ? M2 : 0;
    for (unsigned int i = 0; i < M; i++) {
        unsigned int b = (offset + i) % M;
        q->X[b] = dot_crcf(q->sub + (size_t)i * len, q->w0 + (size_t)b * len, len);
    }
    orc_fft_run(M, q->X, q->x, ORC_FFT_BACKWARD);
    for (unsigned int i = 0; i < M; i++)
        y[i] = cf(q->x[i].re / (float)M, q->x[i].im / (float)M);
    q->flag = 1 - q->flag;
}

/* firpfbch2.c:287-335 */
static void pfbch2_synthesize(orc_firpfbch2 q, const orc_cf *x, orc_cf *y)
{
    unsigned int M = q->M, M2 = q->M2, len = q->len;
    memcpy(q->X, x, M * sizeof *q->X);
    orc_fft_run(M, q->X, q->x, ORC_FFT_BACKWARD);
    /* two separate float scalings, as the reference does them */
    float g0 = 1.0f / (float)M, g1 = (float)M2;
    for (unsigned int i = 0; i < M; i++) { q->x[i].re *= g0; q->x[i].im *= g0; }
    for (unsigned int i = 0; i < M; i++) { q->x[i].re *= g1; q->x[i].im *= g1; }
    orc_cf *buf = (q->flag == 0) ? q->w1 : q->w0;
    for (unsigned int i = 0; i < M; i++)
        win_push(buf + (size_t)i * len, len, q->x[i]);
    for (unsigned int i = 0; i < M2; i++) {
        unsigned int b = (q->flag == 0) ? i : i + M2;
        const orc_cf *r0 = q->w0 + (size_t)b * len;
        const orc_cf *r1 = q->w1 + (size_t)b * len;
        const orc_cf *p0 = q->flag ? r0 : r1;
        const orc_cf *p1 = q->flag ? r1 : r0;
        orc_cf y0 = dot_crcf(q->sub + (size_t)i * len, p0, len);
        orc_cf y1 = dot_crcf(q->sub + (size_t)(i + M2) * len, p1, len);
        y[i] = cf(y0.re + y1.re, y0.im + y1.im);
    }
    q->flag = 1 - q->flag;
}

/* firpfbch2.c:342-358 */
void orc_firpfbch2_execute(orc_firpfbch2 q, const orc_cf *x, orc_cf *y)
{
    if (q->type == ORC_ANALYZER) pfbch2_analyze(q, x, y);
    else                         pfbch2_synthesize(q, x, y);
}

void orc_firpfbch2_execute_frames(orc_firpfbch2 q, const orc_cf *x, unsigned int n_frames, orc_cf *y)
{
    size_t in = (q->type == ORC_ANALYZER) ? q->M2 : q->M;
    size_t out = (q->type == ORC_ANALYZER) ? q->M : q->M2;
    for (unsigned int t = 0; t < n_frames; t++)
        orc_firpfbch2_execute(q, x + t * in, y + t * out);
}

/* ------------------------------------------------------------------------------------ */
/* synthetic input + CPU baseline driver (not a restatement: bench plumbing)             */
/* ------------------------------------------------------------------------------------ */

static unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* u(n,c) = (mix64(seed + 2n + c) >> 40) * 2^-24 - 0.5       */
void orc_synth_input(unsigned long long seed, unsigned long long first, unsigned long long n, orc_cf *x)
{
    for (unsigned long long i = 0; i < n; i++) {
        unsigned long long k = seed + 2 * (first + i);
        x[i].re = (float)(mix64(k) >> 40) * (1.0f / 16777216.0f) - 0.5f;
        x[i].im = (float)(mix64(k + 1) >> 40) * (1.0f / 16777216.0f) - 0.5f;
    }
}

struct bench_arg { unsigned int M, m; float As; unsigned long long n; unsigned int id; double secs; };

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void *bench_thread(void *p)
{
    struct bench_arg *a = p;
    orc_firpfbch2 q = orc_firpfbch2_create_kaiser(ORC_ANALYZER, a->M, a->m, a->As);
    unsigned int M2 = a->M / 2;
    unsigned long long frames = a->n / M2;
    unsigned int blk = 256;                                    /* frames per generated block */
    orc_cf *x = xmalloc((size_t)blk * M2 * sizeof *x);
    orc_cf *y = xmalloc((size_t)blk * a->M * sizeof *y);
    orc_synth_input(0x5EED0000ULL + a->id, 0, (unsigned long long)blk * M2, x);
    double t0 = now_s();
    for (unsigned long long f = 0; f < frames; f += blk) {
        unsigned int nb = (frames - f) < blk ? (unsigned int)(frames - f) : blk;
        orc_firpfbch2_execute_frames(q, x, nb, y);
    }
    a->secs = now_s() - t0;
    free(x); free(y);
    orc_firpfbch2_destroy(q);
    return NULL;
}

double orc_bench_firpfbch2_analysis(unsigned int M, unsigned int m, float As,
                                    unsigned long long n_samples, unsigned int n_threads)
{
    pthread_t *th = xmalloc(n_threads * sizeof *th);
    struct bench_arg *args = xmalloc(n_threads * sizeof *args);
    for (unsigned int i = 0; i < n_threads; i++) {
        args[i].M = M; args[i].m = m; args[i].As = As; args[i].n = n_samples; args[i].id = i; args[i].secs = 0;
        pthread_create(&th[i], NULL, bench_thread, &args[i]);
    }
    double worst = 0;
    for (unsigned int i = 0; i < n_threads; i++) {
        pthread_join(th[i], NULL);
        if (args[i].secs > worst) worst = args[i].secs;
    }
    free(th); free(args);
    return worst;
}
