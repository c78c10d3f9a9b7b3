// ofdmframesync.cu -- OFDM frame synchronizer (receive side).
//
// Replaces src/multichannel/src/ofdmframesync.c: create :140-270, reset :327-352, execute :354-400, detection
// states seekplcp / S0a / S0b / S1 :415-700, rxsymbols :714-750, S0 metrics :754-772, gain estimates :779-840,
// polynomial equalizer fit :905-985, rxsymbol :989-1100.  Same API (liquid.h:5878-5911), same callback contract.
//
// Split of the work:
//   * ACQUISITION (seek PLCP, the two short symbols, the long symbol, the equalizer fit) is a sample-by-sample
//     state machine with one small transform every M/2 .. M samples: it runs on the host, as the reference does.
//   * Once the long symbol is found the object is in the PAYLOAD state, where the reference spends its time: per
//     sample an NCO mix (256-entry sine table indexed from a float32 phase that is accumulated sample by sample),
//     per symbol an M-point transform, the per-subcarrier equalizer product, the pilot phases, a first-order fit,
//     the de-rotation of every subcarrier, and a feedback of the common phase change into the NCO frequency.  All
//     of that runs on the device: ofdm_rxsymbols_kernel takes every sample that is left in an execute() call,
//     walks the symbols IN ORDER (the NCO feedback makes symbol s + 1's mixing depend on symbol s's pilots) and
//     works in parallel inside a symbol; it leaves one row of M equalised subcarriers per symbol, and the host
//     then makes the callbacks.  A callback that returns non-zero (or resets the object) ends the frame: the rows
//     after it are dropped, the window of mixed samples is rewound from what the kernel left behind, and the
//     remaining samples go back through acquisition -- exactly the stream of callbacks the reference produces.
//     The phase accumulation keeps the reference's float32 order (one lane, sequentially: its table index must
//     not flip), everything else is per-sample / per-subcarrier parallel.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "ofdm_common.cuh"

namespace {

using namespace lcofdm;

constexpr int kRxThreads = 256;
constexpr unsigned int kMaxBatchSymbols = 512;

struct RxState {                 // what the payload state carries from symbol to symbol (device resident)
    float theta, d_theta;        // NCO phase and frequency (nco.c:40-41)
    float phi_prime, p1_prime;   // last common pilot phase, filtered pilot phase slope
    unsigned int num_symbols;
    unsigned int pilot_pos;      // position in the 255-periodic pilot m-sequence
};

struct RxArgs {
    const float2 *x;             // raw samples of this call's payload stretch
    float2 *mixed;               // the same samples after the NCO (kept for a rewind)
    float2 *win;                 // [M + cp] the mixed-sample window, oldest first (in: before x[0]; out: after the last sample)
    float2 *X;                   // [n_sym][M] equalised symbols
    RxState *st;
    const float2 *R;             // [M] composite equalizer gain B / G
    const float2 *twd;           // [M] e^{-j 2 pi i / M}
    const unsigned char *p;      // [M] subcarrier types
    const unsigned int *pilot_k; // [M_pilot] pilot subcarriers in the order (i + M/2) % M
    const float *pilot_x;        // [M_pilot] their signed frequency index
    const float *pilot_v;        // [255] pilot values +-1
    const float *sintab;         // [256] nco.c:58-59
    unsigned int M, cp, backoff, M_pilot, log2M;   // log2M = 0: M is not a power of two (direct transform)
    unsigned int n_in, t0, n_sym;                   // samples, samples until the first symbol completes, symbols
};

// float32 product of two complex numbers as C99 computes it without contraction: (ac - bd) + j (ad + bc)
__device__ __forceinline__ float2 cmul_c99(float2 a, float2 b)
{
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

// one CTA, sequential over symbols, parallel inside a symbol
__global__ void __launch_bounds__(kRxThreads) ofdm_rxsymbols_kernel(const RxArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned int M = a.M, W = a.M + a.cp;
    float2 *win = reinterpret_cast<float2 *>(smem_raw);                  // [W]
    float2 *fa = win + W;                                                 // [M] transform ping
    float2 *fb = fa + M;                                                  // [M] transform pong
    float *yph = reinterpret_cast<float *>(fb + M);                       // [M_pilot]
    unsigned char *tidx = reinterpret_cast<unsigned char *>(yph + a.M_pilot);   // [W + 4] table indices of a stretch
    __shared__ RxState st;
    __shared__ float pp[2];
    const unsigned int tid = threadIdx.x;

    for (unsigned int i = tid; i < W; i += kRxThreads) win[i] = a.win[i];
    if (tid == 0) st = *a.st;
    __syncthreads();

    unsigned int done = 0;                                               // samples consumed
    for (unsigned int s = 0; s <= a.n_sym; s++) {
        // the stretch of new samples that completes symbol s (the last stretch is the tail after the last symbol)
        unsigned int n_new = (s < a.n_sym) ? (s == 0 ? a.t0 : W) : a.n_in - done;
        // ---- NCO: phase indices in the reference's float32 order (nco.c:134-138, 364-376), one lane
        for (unsigned int c0 = 0; c0 < n_new; c0 += W) {                 // stretches longer than the window: in pieces
            const unsigned int nc = min(W, n_new - c0);
            if (tid == 0) {
                float th = st.theta;
                const float dth = st.d_theta;
                for (unsigned int k = 0; k < nc; k++) {
                    tidx[k] = (unsigned char)(((unsigned int)(__fadd_rn(__fadd_rn(__fmul_rn(th, 40.743665f), 512.0f), 0.5f))) & 0xffu);
                    th = __fadd_rn(th, dth);
                    if ((double)th > M_PI) th = (float)((double)th - 2 * M_PI);
                    else if ((double)th < -M_PI) th = (float)((double)th + 2 * M_PI);
                }
                st.theta = th;
            }
            __syncthreads();
            // ---- mix down, append to the window (shift left by nc)
            float2 keep[8];                                               // W <= 8 * 256 handled by the launcher
            unsigned int nk = 0;
            if (nc < W)
                for (unsigned int i = tid; i < W - nc; i += kRxThreads) keep[nk++] = win[i + nc];
            __syncthreads();
            nk = 0;
            if (nc < W)
                for (unsigned int i = tid; i < W - nc; i += kRxThreads) win[i] = keep[nk++];
            for (unsigned int k = tid; k < nc; k += kRxThreads) {
                const unsigned int ix = tidx[k];
                const float sn = a.sintab[ix], cs = a.sintab[(ix + 64) & 0xff];
                const float2 y = cmul_c99(a.x[done + c0 + k], make_float2(cs, -sn));
                a.mixed[done + c0 + k] = y;
                win[W - nc + k] = y;
            }
            __syncthreads();
        }
        done += n_new;
        if (s == a.n_sym) break;

        // ---- M-point forward transform of win[cp - backoff .. + M)
        const float2 *src = win + (a.cp - a.backoff);
        float2 *X = fa;
        if (a.log2M) {
            // radix-2 Stockham autosort, ping-pong between fa and fb
            const float2 *in = src;
            float2 *out = fa;
            for (unsigned int Ns = 1; Ns < M; Ns <<= 1) {
                for (unsigned int j = tid; j < M / 2; j += kRxThreads) {
                    const unsigned int k = j & (Ns - 1);
                    const float2 w = a.twd[k * (M / (2 * Ns))];
                    const float2 u = in[j], v = lcd::cmul(in[j + M / 2], w);
                    const unsigned int o = ((j - k) << 1) + k;
                    out[o] = lcd::cadd(u, v);
                    out[o + Ns] = lcd::csub(u, v);
                }
                __syncthreads();
                X = out;
                in = out;
                out = (out == fa) ? fb : fa;
            }
        } else {
            for (unsigned int k = tid; k < M; k += kRxThreads) {
                float2 acc = make_float2(0.f, 0.f);
                unsigned int idx = 0;
                for (unsigned int i = 0; i < M; i++) {
                    acc = lcd::cadd(acc, lcd::cmul(src[i], a.twd[idx]));
                    idx += k;
                    if (idx >= M) idx -= M;
                }
                fa[k] = acc;
            }
            __syncthreads();
        }
        // ---- equalizer gain (ofdmframesync.c:993-994), pilot phases (:1004-1027)
        for (unsigned int i = tid; i < M; i += kRxThreads) X[i] = cmul_c99(X[i], a.R[i]);
        __syncthreads();
        for (unsigned int n = tid; n < a.M_pilot; n += kRxThreads) {
            const float2 v = X[a.pilot_k[n]];
            const float pv = a.pilot_v[(st.pilot_pos + n) % kPilotPeriod];
            yph[n] = atan2f(v.y * pv, v.x * pv);
        }
        __syncthreads();
        if (tid == 0) {
            // unwrap (:1035-1041), first-order least squares through the reference's normal equations (poly.common.c:46-90,
            // Gauss-Jordan with row pivoting on [X'X | I]), slope filter (:1046-1049)
            for (unsigned int i = 1; i < a.M_pilot; i++) {
                while ((double)(yph[i] - yph[i - 1]) > M_PI) yph[i] = (float)((double)yph[i] - 2 * M_PI);
                while ((double)(yph[i] - yph[i - 1]) < -M_PI) yph[i] = (float)((double)yph[i] + 2 * M_PI);
            }
            float sx = 0.f, sxx = 0.f, sy = 0.f, sxy = 0.f, n1 = 0.f;
            for (unsigned int i = 0; i < a.M_pilot; i++) {
                const float xv = a.pilot_x[i];
                n1 = __fadd_rn(n1, 1.0f);
                sx = __fadd_rn(sx, xv);
                sxx = __fadd_rn(sxx, __fmul_rn(xv, xv));
                sy = __fadd_rn(sy, yph[i]);
                sxy = __fadd_rn(sxy, __fmul_rn(xv, yph[i]));
            }
            float m[2][4] = {{n1, sx, 1.f, 0.f}, {sx, sxx, 0.f, 1.f}};
            for (int r = 0; r < 2; r++) {
                int ro = r;
                if (r == 0 && fabsf(m[1][0]) > fabsf(m[0][0])) ro = 1;
                if (ro != r)
                    for (int c = 0; c < 4; c++) { const float t = m[r][c]; m[r][c] = m[ro][c]; m[ro][c] = t; }
                const float v = m[r][r];
                const int o = 1 - r;
                const float g = m[o][r] / v;
                for (int c = 0; c < 4; c++) m[o][c] = __fsub_rn(__fmul_rn(g, m[r][c]), m[o][c]);
            }
            for (int r = 0; r < 2; r++) {
                const float g = 1.0f / m[r][r];
                for (int c = 0; c < 4; c++) m[r][c] = __fmul_rn(m[r][c], g);
            }
            const float p0 = __fadd_rn(__fmul_rn(m[0][2], sy), __fmul_rn(m[0][3], sxy));
            float p1 = __fadd_rn(__fmul_rn(m[1][2], sy), __fmul_rn(m[1][3], sxy));
            const float alpha = 0.3f;
            p1 = __fadd_rn(__fmul_rn(alpha, p1), __fmul_rn(1 - alpha, st.p1_prime));
            st.p1_prime = p1;
            pp[0] = p0;
            pp[1] = p1;
            // NCO frequency from the change of the common phase (:1079-1089; the reference unwraps with M_2_PI = 2/pi)
            if (st.num_symbols > 0) {
                float d = __fsub_rn(p0, st.phi_prime);
                while ((double)d > M_PI) d = (float)((double)d - M_2_PI);
                while ((double)d < -M_PI) d = (float)((double)d + M_2_PI);
                st.d_theta = __fadd_rn(st.d_theta, __fmul_rn(1e-3f, d));
            }
            st.phi_prime = p0;
            st.num_symbols++;
            st.pilot_pos = (st.pilot_pos + a.M_pilot) % kPilotPeriod;
        }
        __syncthreads();
        // ---- de-rotate (:1066-1076) and leave the row
        for (unsigned int i = tid; i < M; i += kRxThreads) {
            float2 v = make_float2(0.f, 0.f);
            if (a.p[i] != OFDMFRAME_SCTYPE_NULL) {
                const float fx = (i > M / 2) ? (float)i - (float)M : (float)i;
                const float th = __fadd_rn(__fmul_rn(pp[0], 1.0f), __fmul_rn(pp[1], fx));      // polyf_val: p0 * 1 + p1 * x
                v = cmul_c99(X[i], make_float2(cosf(-th), sinf(-th)));
            }
            a.X[(size_t)s * M + i] = v;
        }
        __syncthreads();
    }
    for (unsigned int i = tid; i < W; i += kRxThreads) a.win[i] = win[i];
    if (tid == 0) *a.st = st;
}

// ---------------------------------------------------------------------------------- host helpers
typedef std::vector<lc_cf> cvec;

inline lc_cf cmulh(lc_cf a, lc_cf b) { return lc_cf{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline lc_cf conjh(lc_cf a) { return lc_cf{a.re, -a.im}; }
inline float cabsh(lc_cf a) { return hypotf(a.re, a.im); }
inline float cargh(lc_cf a) { return atan2f(a.im, a.re); }

// y = sum x[i] e^{sign j 2 pi i k / M}: the acquisition's few transforms (one per M/2 .. M samples), on the host
void host_dft(const lc_cf *x, lc_cf *y, unsigned int M, int sign, const cvec &tw)
{
    if ((M & (M - 1)) == 0) {
        cvec a(x, x + M), b(M);
        lc_cf *in = a.data(), *out = b.data();
        for (unsigned int Ns = 1; Ns < M; Ns <<= 1) {
            for (unsigned int j = 0; j < M / 2; j++) {
                const unsigned int k = j & (Ns - 1);
                lc_cf w = tw[k * (M / (2 * Ns))];
                if (sign > 0) w.im = -w.im;
                const lc_cf u = in[j], v = cmulh(in[j + M / 2], w);
                const unsigned int o = ((j - k) << 1) + k;
                out[o] = lc_cf{u.re + v.re, u.im + v.im};
                out[o + Ns] = lc_cf{u.re - v.re, u.im - v.im};
            }
            std::swap(in, out);
        }
        std::copy(in, in + M, y);
        return;
    }
    for (unsigned int k = 0; k < M; k++) {
        double ar = 0.0, ai = 0.0;
        unsigned int idx = 0;
        for (unsigned int i = 0; i < M; i++) {
            lc_cf w = tw[idx];
            if (sign > 0) w.im = -w.im;
            ar += (double)x[i].re * w.re - (double)x[i].im * w.im;
            ai += (double)x[i].re * w.im + (double)x[i].im * w.re;
            idx += k;
            if (idx >= M) idx -= M;
        }
        y[k] = lc_cf{(float)ar, (float)ai};
    }
}

// least squares polynomial fit of k coefficients (poly.common.c:46-90: normal equations, Gauss-Jordan inverse)
void host_polyfit(const std::vector<float> &x, const std::vector<float> &y, unsigned int k, float *p)
{
    const unsigned int n = (unsigned int)x.size();
    std::vector<float> X((size_t)n * k);
    for (unsigned int r = 0; r < n; r++) {
        float v = 1.f;
        for (unsigned int c = 0; c < k; c++) {
            X[(size_t)r * k + c] = v;
            v *= x[r];
        }
    }
    std::vector<float> Xty(k, 0.f), X2((size_t)k * k, 0.f);
    for (unsigned int c = 0; c < k; c++) {
        float s = 0.f;
        for (unsigned int r = 0; r < n; r++) s += X[(size_t)r * k + c] * y[r];
        Xty[c] = s;
        for (unsigned int c2 = 0; c2 < k; c2++) {
            float t = 0.f;
            for (unsigned int r = 0; r < n; r++) t += X[(size_t)r * k + c] * X[(size_t)r * k + c2];
            X2[(size_t)c * k + c2] = t;
        }
    }
    // [X2 | I] -> [I | inv]   (matrix.inv.c: row pivot on the column maximum, pivot g * row_r - row, scale by the diagonal)
    const unsigned int C = 2 * k;
    std::vector<float> m((size_t)k * C, 0.f);
    for (unsigned int r = 0; r < k; r++) {
        for (unsigned int c = 0; c < k; c++) m[(size_t)r * C + c] = X2[(size_t)r * k + c];
        m[(size_t)r * C + k + r] = 1.f;
    }
    for (unsigned int r = 0; r < k; r++) {
        unsigned int ro = r;
        float vmax = 0.f;
        for (unsigned int rh = r; rh < k; rh++) {
            const float v = fabsf(m[(size_t)rh * C + r]);
            if (v > vmax || rh == r) { ro = rh; vmax = v; }
        }
        if (ro != r)
            for (unsigned int c = 0; c < C; c++) std::swap(m[(size_t)r * C + c], m[(size_t)ro * C + c]);
        const float v = m[(size_t)r * C + r];
        if (v == 0.f) continue;
        for (unsigned int r2 = 0; r2 < k; r2++) {
            if (r2 == r) continue;
            const float g = m[(size_t)r2 * C + r] / v;
            for (unsigned int c = 0; c < C; c++) m[(size_t)r2 * C + c] = g * m[(size_t)r * C + c] - m[(size_t)r2 * C + c];
        }
    }
    for (unsigned int r = 0; r < k; r++) {
        const float g = 1.f / m[(size_t)r * C + r];
        for (unsigned int c = 0; c < C; c++) m[(size_t)r * C + c] *= g;
    }
    for (unsigned int r = 0; r < k; r++) {
        float s = 0.f;
        for (unsigned int c = 0; c < k; c++) s += m[(size_t)r * C + k + c] * Xty[c];
        p[r] = s;
    }
}

inline float host_polyval(const float *p, unsigned int k, float x)
{
    float xk = 1.f, y = 0.f;
    for (unsigned int i = 0; i < k; i++) {
        y += p[i] * xk;
        xk *= x;
    }
    return y;
}

}  // namespace

enum { ST_SEEKPLCP = 0, ST_PLCPSHORT0, ST_PLCPSHORT1, ST_PLCPLONG, ST_RXSYMBOLS };

struct ofdmframesync_s {
    unsigned int M, M2, cp_len, backoff;
    std::vector<unsigned char> p;
    unsigned int M_null, M_pilot, M_data, M_S0, M_S1;
    cvec S0, s0, S1, s1, G0, G1, G, B, R, tw;
    std::vector<float> sintab;
    float g0 = 1.0f;
    int state = ST_SEEKPLCP;
    float theta = 0.f, d_theta = 0.f;          // NCO
    float phi_prime = 0.f, p1_prime = 0.f;
    unsigned int pilot_pos = 0;
    unsigned int timer = 0, num_symbols = 0;
    lc_cf s_hat_0{0.f, 0.f}, s_hat_1{0.f, 0.f};
    float plcp_detect_thresh = 0.35f, plcp_sync_thresh = 0.30f;
    ofdmframesync_callback callback = nullptr;
    void *userdata = nullptr;
    cvec win, win_saved;                        // the last M + cp (mixed) samples, oldest first; a copy for the rewind
    bool reset_seen = false;                    // set by reset(): a callback may call it on its own object
    // device side
    cudaStream_t stream = nullptr;
    float2 *d_R = nullptr, *d_tw = nullptr, *d_win = nullptr;
    unsigned char *d_p = nullptr;
    unsigned int *d_pilot_k = nullptr;
    float *d_pilot_x = nullptr, *d_pilot_v = nullptr, *d_sintab = nullptr;
    RxState *d_st = nullptr;
    lc::DevBuf d_x, d_mixed, d_X;
    std::vector<lc_cf> hX;
    size_t smem_bytes = 0;
    unsigned int log2M = 0;
};

extern "C" void ofdmframesync_reset(ofdmframesync _q);

namespace {

void nco_mix_push(ofdmframesync q, lc_cf x)
{
    if (q->state != ST_SEEKPLCP) {
        const unsigned int ix = ((unsigned int)(q->theta * 40.743665f + 512.0f + 0.5f)) & 0xff;      // nco.c:364-376
        const float sn = q->sintab[ix], cs = q->sintab[(ix + 64) & 0xff];
        x = cmulh(x, lc_cf{cs, -sn});                                                                // nco.c:255-264
        q->theta += q->d_theta;                                                                        // nco.c:134-138
        if (q->theta > M_PI) q->theta -= 2 * M_PI;
        else if (q->theta < -M_PI) q->theta += 2 * M_PI;
    }
    const size_t W = q->win.size();
    memmove(q->win.data(), q->win.data() + 1, (W - 1) * sizeof(lc_cf));
    q->win[W - 1] = x;
}

// gain estimate from a short (which = 0) or long (1) training symbol starting at x (ofdmframesync.c:779-840)
void estimate_gain(ofdmframesync q, const lc_cf *x, int which, cvec &G)
{
    cvec X(q->M);
    host_dft(x, X.data(), q->M, -1, q->tw);
    const float gain = sqrtf((float)(which ? q->M_S1 : q->M_S0)) / (float)q->M;
    const cvec &S = which ? q->S1 : q->S0;
    for (unsigned int i = 0; i < q->M; i++) {
        if (q->p[i] != OFDMFRAME_SCTYPE_NULL && (which || (i % 2) == 0)) G[i] = cmulh(X[i], conjh(S[i]));
        else G[i] = lc_cf{0.f, 0.f};
        G[i].re *= gain;
        G[i].im *= gain;
    }
}

lc_cf s0_metrics(ofdmframesync q, const cvec &G)
{
    lc_cf s{0.f, 0.f};
    for (unsigned int i = 0; i < q->M; i += 2) {
        const lc_cf t = cmulh(G[(i + 2) % q->M], conjh(G[i]));
        s.re += t.re;
        s.im += t.im;
    }
    s.re /= (float)q->M_S0;
    s.im /= (float)q->M_S0;
    return s;
}

void seekplcp(ofdmframesync q)
{
    q->timer++;
    if (q->timer < q->M) return;
    q->timer = 0;
    const lc_cf *rc = q->win.data();
    float g = 0.0f;
    for (unsigned int i = q->cp_len; i < q->M + q->cp_len; i++) g += rc[i].re * rc[i].re + rc[i].im * rc[i].im;
    g = (float)q->M / g;
    estimate_gain(q, rc + q->cp_len, 0, q->G0);
    lc_cf s_hat = s0_metrics(q, q->G0);
    s_hat.re *= g;
    s_hat.im *= g;
    const float tau_hat = (float)((double)(cargh(s_hat) * (float)q->M2) / (2 * M_PI));
    q->g0 = g;
    if (cabsh(s_hat) > q->plcp_detect_thresh) {
        const int dt = (int)roundf(tau_hat);
        q->timer = (q->M + dt) % q->M2;
        q->timer += q->M;
        q->state = ST_PLCPSHORT0;
    }
}

void plcp_s0a(ofdmframesync q)
{
    q->timer++;
    if (q->timer < q->M2) return;
    q->timer = 0;
    estimate_gain(q, q->win.data() + q->cp_len, 0, q->G0);
    lc_cf s_hat = s0_metrics(q, q->G0);
    s_hat.re *= q->g0;
    s_hat.im *= q->g0;
    q->s_hat_0 = s_hat;
    q->state = ST_PLCPSHORT1;
}

void plcp_s0b(ofdmframesync q)
{
    q->timer++;
    if (q->timer < q->M2) return;
    q->timer = q->M + q->cp_len - q->backoff;
    const lc_cf *rc = q->win.data();
    estimate_gain(q, rc + q->cp_len, 0, q->G1);
    lc_cf s_hat = s0_metrics(q, q->G1);
    s_hat.re *= q->g0;
    s_hat.im *= q->g0;
    q->s_hat_1 = s_hat;
    const float tau_prime =
        (float)((double)(cargh(lc_cf{q->s_hat_0.re + q->s_hat_1.re, q->s_hat_0.im + q->s_hat_1.im}) * (float)q->M2) / (2 * M_PI));
    q->timer -= (int)roundf(tau_prime);
    // carrier offset, ML estimate over the two halves of the short symbol (:597-603)
    lc_cf t0{0.f, 0.f};
    for (unsigned int i = 0; i < q->M2; i++) {
        const lc_cf a = cmulh(conjh(rc[i]), q->s0[i]);
        const lc_cf b = cmulh(rc[i + q->M2], conjh(q->s0[i + q->M2]));
        const lc_cf t = cmulh(a, b);
        t0.re += t.re;
        t0.im += t.im;
    }
    q->d_theta = cargh(t0) / (float)q->M2;
    q->state = ST_PLCPLONG;
}

void eqgain_poly(ofdmframesync q, unsigned int order)
{
    const unsigned int N = q->M_pilot + q->M_data;
    if (order > N - 1) order = N - 1;
    if (order > 10) order = 10;
    std::vector<float> xf(N), ya(N), yg(N);
    unsigned int n = 0;
    for (unsigned int i = 0; i < q->M; i++) {
        const unsigned int k = (i + q->M2) % q->M;
        if (q->p[k] == OFDMFRAME_SCTYPE_NULL) continue;
        xf[n] = ((k > q->M2) ? (float)k - (float)q->M : (float)k) / (float)q->M;
        ya[n] = cabsh(q->G[k]);
        yg[n] = cargh(q->G[k]);
        n++;
    }
    for (unsigned int i = 1; i < N; i++) {
        while ((yg[i] - yg[i - 1]) > M_PI) yg[i] -= 2 * M_PI;
        while ((yg[i] - yg[i - 1]) < -M_PI) yg[i] += 2 * M_PI;
    }
    float pa[11], pg[11];
    host_polyfit(xf, ya, order + 1, pa);
    host_polyfit(xf, yg, order + 1, pg);
    for (unsigned int i = 0; i < q->M; i++) {
        const float f = ((i > q->M2) ? (float)i - (float)q->M : (float)i) / (float)q->M;
        const float A = host_polyval(pa, order + 1, f), th = host_polyval(pg, order + 1, f);
        q->G[i] = (q->p[i] == OFDMFRAME_SCTYPE_NULL) ? lc_cf{0.f, 0.f} : lc_cf{A * cosf(th), A * sinf(th)};
    }
}

void plcp_s1(ofdmframesync q)
{
    q->timer--;
    if (q->timer > 0) return;
    q->num_symbols++;
    estimate_gain(q, q->win.data() + q->cp_len, 1, q->G);
    lc_cf g_hat{0.f, 0.f};
    for (unsigned int i = 0; i < q->M; i++) {
        const lc_cf t = cmulh(q->G[(i + 1) % q->M], conjh(q->G[i]));
        g_hat.re += t.re;
        g_hat.im += t.im;
    }
    g_hat.re = g_hat.re / (float)q->M_S1 * q->g0;
    g_hat.im = g_hat.im / (float)q->M_S1 * q->g0;
    const float ph = (float)((double)((float)q->backoff * 2.0f) * M_PI / (double)(float)q->M);
    g_hat = cmulh(g_hat, lc_cf{cosf(ph), sinf(ph)});
    if (cabsh(g_hat) > q->plcp_sync_thresh && fabsf(cargh(g_hat)) < 0.1f * M_PI) {
        q->state = ST_RXSYMBOLS;
        q->timer = q->M + q->cp_len + q->backoff;
        q->num_symbols = 0;
        const float g = (float)q->M / sqrtf((float)(q->M_pilot + q->M_data));
        for (unsigned int i = 0; i < q->M; i++) {
            q->G[i].re *= g;
            q->G[i].im *= g;
            q->G[i] = cmulh(q->G[i], q->B[i]);
        }
        unsigned int order = 4;
        if (order >= q->M_pilot + q->M_data) order = q->M_pilot + q->M_data - 1;
        eqgain_poly(q, order);
        for (unsigned int i = 0; i < q->M; i++) {
            // R = B / G (complex division as C99 does it for finite operands)
            const float d = q->G[i].re * q->G[i].re + q->G[i].im * q->G[i].im;
            const lc_cf n = cmulh(q->B[i], conjh(q->G[i]));
            q->R[i] = lc_cf{n.re / d, n.im / d};
        }
        LC_CUDA_CHECK(cudaMemcpyAsync(q->d_R, q->R.data(), sizeof(float2) * q->M, cudaMemcpyHostToDevice, q->stream));
        return;
    }
    // not found yet: look again half a symbol later, give up after 16 tries (:693-712)
    if (q->num_symbols == 16) ofdmframesync_reset(q);
    q->timer = q->M2;
}

// the payload state over x[0 .. n): returns the number of samples consumed (all of them unless a callback ended the frame)
unsigned int rx_payload(ofdmframesync q, const lc_cf *x, unsigned int n)
{
    const unsigned int W = q->M + q->cp_len;
    // symbols that complete inside this stretch
    unsigned int n_sym = (n >= q->timer) ? 1 + (n - q->timer) / W : 0;
    if (n_sym > kMaxBatchSymbols) n_sym = kMaxBatchSymbols;
    const unsigned int n_use = (n_sym == kMaxBatchSymbols) ? q->timer + (n_sym - 1) * W : n;     // a capped batch ends on a symbol
    float2 *d_x = (float2 *)q->d_x.get(sizeof(float2) * n_use);
    float2 *d_mixed = (float2 *)q->d_mixed.get(sizeof(float2) * n_use);
    float2 *d_X = (float2 *)q->d_X.get(sizeof(float2) * (size_t)(n_sym ? n_sym : 1) * q->M);
    RxState st{q->theta, q->d_theta, q->phi_prime, q->p1_prime, q->num_symbols, q->pilot_pos};
    LC_CUDA_CHECK(cudaMemcpyAsync(d_x, x, sizeof(float2) * n_use, cudaMemcpyHostToDevice, q->stream));
    LC_CUDA_CHECK(cudaMemcpyAsync(q->d_win, q->win.data(), sizeof(float2) * W, cudaMemcpyHostToDevice, q->stream));
    LC_CUDA_CHECK(cudaMemcpyAsync(q->d_st, &st, sizeof st, cudaMemcpyHostToDevice, q->stream));
    RxArgs a;
    a.x = d_x;
    a.mixed = d_mixed;
    a.win = q->d_win;
    a.X = d_X;
    a.st = q->d_st;
    a.R = q->d_R;
    a.twd = q->d_tw;
    a.p = q->d_p;
    a.pilot_k = q->d_pilot_k;
    a.pilot_x = q->d_pilot_x;
    a.pilot_v = q->d_pilot_v;
    a.sintab = q->d_sintab;
    a.M = q->M;
    a.cp = q->cp_len;
    a.backoff = q->backoff;
    a.M_pilot = q->M_pilot;
    a.log2M = q->log2M;
    a.n_in = n_use;
    a.t0 = q->timer;
    a.n_sym = n_sym;
    ofdm_rxsymbols_kernel<<<1, kRxThreads, q->smem_bytes, q->stream>>>(a);
    LC_LAUNCH_CHECK();
    q->hX.resize((size_t)(n_sym ? n_sym : 1) * q->M);
    if (n_sym) LC_CUDA_CHECK(cudaMemcpyAsync(q->hX.data(), d_X, sizeof(float2) * (size_t)n_sym * q->M, cudaMemcpyDeviceToHost, q->stream));
    LC_CUDA_CHECK(cudaMemcpyAsync(q->win.data(), q->d_win, sizeof(float2) * W, cudaMemcpyDeviceToHost, q->stream));
    LC_CUDA_CHECK(cudaMemcpyAsync(&st, q->d_st, sizeof st, cudaMemcpyDeviceToHost, q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(q->stream));
    // the callbacks, in order; the first one that ends the frame drops everything after it
    const unsigned int t0 = q->timer;
    for (unsigned int s = 0; s < n_sym; s++) {
        q->reset_seen = false;
        int rv = 0;
        if (q->callback)
            rv = q->callback(reinterpret_cast<liquid_float_complex *>(q->hX.data() + (size_t)s * q->M), q->p.data(), q->M, q->userdata);
        if (rv != 0 || q->reset_seen) {
            if (rv != 0) ofdmframesync_reset(q);                      // (:741-742)
            // rewind: the window as it stood when symbol s completed = the last W of (window before | mixed[0 .. end));
            // the rest of the stretch is not payload any more
            const unsigned int end = t0 + s * W;                      // samples consumed up to and including symbol s
            const unsigned int take = end < W ? end : W;
            cvec mixed(take);
            LC_CUDA_CHECK(cudaMemcpy(mixed.data(), d_mixed + (end - take), sizeof(float2) * take, cudaMemcpyDeviceToHost));
            std::copy(q->win_saved.begin() + take, q->win_saved.end(), q->win.begin());
            std::copy(mixed.begin(), mixed.end(), q->win.begin() + (W - take));
            q->timer = W;                                              // (:745-746: set after the callback, reset or not)
            return end;
        }
    }
    // no callback ended the frame: adopt the kernel's state
    q->theta = st.theta;
    q->d_theta = st.d_theta;
    q->phi_prime = st.phi_prime;
    q->p1_prime = st.p1_prime;
    q->num_symbols = st.num_symbols;
    q->pilot_pos = st.pilot_pos;
    if (n_sym == 0) q->timer -= n_use;
    else q->timer = W - (n_use - (t0 + (n_sym - 1) * W));
    return n_use;
}

}  // namespace

extern "C" {

ofdmframesync ofdmframesync_create(unsigned int _M, unsigned int _cp_len, unsigned int _taper_len, unsigned char *_p,
                                   ofdmframesync_callback _callback, void *_userdata)
{
    (void)_taper_len;
    if (_M < 8) fprintf(stderr, "warning: ofdmframesync_create(), less than 8 subcarriers\n");
    else if (_M % 2) lc::fatal("ofdmframesync_create", "number of subcarriers must be even");
    else if (_cp_len > _M) lc::fatal("ofdmframesync_create", "cyclic prefix length cannot exceed number of subcarriers");
    std::vector<unsigned char> p(_M);
    if (_p == nullptr) ofdmframe_init_default_sctype(_M, p.data());
    else std::copy(_p, _p + _M, p.begin());
    unsigned int n0, n1, n2;
    ofdmframe_validate_sctype(p.data(), _M, &n0, &n1, &n2);
    if (n1 + n2 == 0) lc::fatal("ofdmframesync_create", "must have at least one enabled subcarrier");
    if (n2 == 0) lc::fatal("ofdmframesync_create", "must have at least one data subcarriers");
    if (n1 < 2) lc::fatal("ofdmframesync_create", "must have at least two pilot subcarriers");
    lc::ensure_device();
    ofdmframesync q = new ofdmframesync_s();
    q->M = _M;
    q->M2 = _M / 2;
    q->cp_len = _cp_len;
    q->p = p;
    q->M_null = n0;
    q->M_pilot = n1;
    q->M_data = n2;
    q->callback = _callback;
    q->userdata = _userdata;
    q->backoff = _cp_len < 2 ? _cp_len : 2;
    const unsigned int M = _M, W = _M + _cp_len;
    if (W > 8 * kRxThreads) lc::fatal("ofdmframesync_create", "symbols longer than 2048 samples are not supported by this build");
    q->win.assign(W, lc_cf{0.f, 0.f});
    q->win_saved.assign(W, lc_cf{0.f, 0.f});
    q->tw.resize(M);
    for (unsigned int i = 0; i < M; i++) {
        const double ang = -2.0 * M_PI * (double)i / (double)M;
        q->tw[i] = lc_cf{(float)cos(ang), (float)sin(ang)};
    }
    // training symbols: frequency domain, and the short one in the time domain (ofdmframe.common.c:41-100)
    q->M_S0 = plcp_freq(p, M, 0, q->S0);
    q->M_S1 = plcp_freq(p, M, 1, q->S1);
    if (q->M_S0 == 0) lc::fatal("ofdmframe_init_S0", "no subcarriers enabled; check allocation");
    if (q->M_S1 == 0) lc::fatal("ofdmframe_init_S1", "no subcarriers enabled; check allocation");
    q->s0.resize(M);
    q->s1.resize(M);
    host_dft(q->S0.data(), q->s0.data(), M, +1, q->tw);
    host_dft(q->S1.data(), q->s1.data(), M, +1, q->tw);
    for (unsigned int i = 0; i < M; i++) {
        const float g0 = 1.0f / sqrtf((float)q->M_S0), g1 = 1.0f / sqrtf((float)q->M_S1);
        q->s0[i].re *= g0; q->s0[i].im *= g0;
        q->s1[i].re *= g1; q->s1[i].im *= g1;
    }
    q->G0.assign(M, lc_cf{0.f, 0.f});
    q->G1 = q->G0;
    q->G = q->G0;
    q->R = q->G0;
    q->B.resize(M);
    const float phi = (float)((double)((float)q->backoff * 2.0f) * M_PI / (double)(float)M);
    for (unsigned int i = 0; i < M; i++) q->B[i] = lc_cf{cosf((float)i * phi), sinf((float)i * phi)};
    q->sintab.resize(256);
    for (unsigned int i = 0; i < 256; i++) q->sintab[i] = sinf((float)(2.0f * M_PI * (float)i / 256.0f));      // nco.c:58-59
    // pilots in the order the reference visits them, their signed frequency index, the 255 pilot values (:1000-1020)
    std::vector<unsigned int> pk;
    std::vector<float> px;
    for (unsigned int i = 0; i < M; i++) {
        const unsigned int k = (i + q->M2) % M;
        if (p[k] != OFDMFRAME_SCTYPE_PILOT) continue;
        pk.push_back(k);
        px.push_back((k > q->M2) ? (float)k - (float)M : (float)k);
    }
    std::vector<float> pv(kPilotPeriod);
    Mseq ms(8);
    for (unsigned int i = 0; i < kPilotPeriod; i++) pv[i] = ms.advance() ? 1.0f : -1.0f;
    q->log2M = 0;
    if ((M & (M - 1)) == 0)
        while ((1u << q->log2M) < M) q->log2M++;
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    LC_CUDA_CHECK(cudaMalloc(&q->d_R, sizeof(float2) * M));
    LC_CUDA_CHECK(cudaMalloc(&q->d_tw, sizeof(float2) * M));
    LC_CUDA_CHECK(cudaMalloc(&q->d_win, sizeof(float2) * W));
    LC_CUDA_CHECK(cudaMalloc(&q->d_p, M));
    LC_CUDA_CHECK(cudaMalloc(&q->d_pilot_k, sizeof(unsigned int) * pk.size()));
    LC_CUDA_CHECK(cudaMalloc(&q->d_pilot_x, sizeof(float) * px.size()));
    LC_CUDA_CHECK(cudaMalloc(&q->d_pilot_v, sizeof(float) * kPilotPeriod));
    LC_CUDA_CHECK(cudaMalloc(&q->d_sintab, sizeof(float) * 256));
    LC_CUDA_CHECK(cudaMalloc(&q->d_st, sizeof(RxState)));
    LC_CUDA_CHECK(cudaMemcpy(q->d_tw, q->tw.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMemcpy(q->d_p, p.data(), M, cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMemcpy(q->d_pilot_k, pk.data(), sizeof(unsigned int) * pk.size(), cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMemcpy(q->d_pilot_x, px.data(), sizeof(float) * px.size(), cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMemcpy(q->d_pilot_v, pv.data(), sizeof(float) * kPilotPeriod, cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMemcpy(q->d_sintab, q->sintab.data(), sizeof(float) * 256, cudaMemcpyHostToDevice));
    q->smem_bytes = sizeof(float2) * (W + 2 * (size_t)M) + sizeof(float) * q->M_pilot + W + 16;
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(ofdm_rxsymbols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ofdmframesync_reset(q);
    return q;
}

void ofdmframesync_destroy(ofdmframesync _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    cudaFree(_q->d_R);
    cudaFree(_q->d_tw);
    cudaFree(_q->d_win);
    cudaFree(_q->d_p);
    cudaFree(_q->d_pilot_k);
    cudaFree(_q->d_pilot_x);
    cudaFree(_q->d_pilot_v);
    cudaFree(_q->d_sintab);
    cudaFree(_q->d_st);
    _q->d_x.release();
    _q->d_mixed.release();
    _q->d_X.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void ofdmframesync_print(ofdmframesync _q)
{
    printf("ofdmframesync:\n");
    printf("    num subcarriers     :   %-u\n", _q->M);
    printf("    cyclic prefix len   :   %-u\n", _q->cp_len);
}

void ofdmframesync_reset(ofdmframesync _q)
{
    _q->theta = 0.f;
    _q->d_theta = 0.f;
    _q->pilot_pos = 0;
    _q->timer = 0;
    _q->num_symbols = 0;
    _q->s_hat_0 = lc_cf{0.f, 0.f};
    _q->s_hat_1 = lc_cf{0.f, 0.f};
    _q->phi_prime = 0.0f;
    _q->p1_prime = 0.0f;
    _q->plcp_detect_thresh = (_q->M > 44) ? 0.35f : 0.35f + 0.01f * (44 - _q->M);
    _q->plcp_sync_thresh = (_q->M > 44) ? 0.30f : 0.30f + 0.01f * (44 - _q->M);
    _q->state = ST_SEEKPLCP;
    _q->reset_seen = true;
}

void ofdmframesync_execute(ofdmframesync _q, liquid_float_complex *_x, unsigned int _n)
{
    const lc_cf *x = reinterpret_cast<const lc_cf *>(_x);
    unsigned int i = 0;
    while (i < _n) {
        if (_q->state == ST_RXSYMBOLS) {
            // every sample left in this call goes to the device in one launch
            _q->win_saved = _q->win;
            i += rx_payload(_q, x + i, _n - i);
            continue;
        }
        nco_mix_push(_q, x[i++]);
        switch (_q->state) {
        case ST_SEEKPLCP: seekplcp(_q); break;
        case ST_PLCPSHORT0: plcp_s0a(_q); break;
        case ST_PLCPSHORT1: plcp_s0b(_q); break;
        case ST_PLCPLONG: plcp_s1(_q); break;
        default: break;
        }
    }
}

float ofdmframesync_get_rssi(ofdmframesync _q) { return -10.0f * log10f(_q->g0); }
float ofdmframesync_get_cfo(ofdmframesync _q) { return _q->d_theta; }
int ofdmframesync_is_frame_open(ofdmframesync _q) { return _q->state == ST_RXSYMBOLS; }

/* the reference's debug trace (a GNU Octave script of its internal buffers) is not reproduced */
void ofdmframesync_debug_enable(ofdmframesync) {}
void ofdmframesync_debug_disable(ofdmframesync) {}
void ofdmframesync_debug_print(ofdmframesync, const char *_filename)
{
    fprintf(stderr, "ofdmframesync_debug_print(): debugging trace is not available in libliquid_cuda (%s not written)\n", _filename);
}

}  // extern "C"
