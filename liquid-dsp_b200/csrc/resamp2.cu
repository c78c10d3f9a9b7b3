// resamp2.cu -- resamp2_crcf (half-band filter, 2-channel analysis / synthesis bank, 2:1 decimator and 1:2
// interpolator) and msresamp2_crcf (their 2^S cascades).  SURVEY.md 8(f) item 3: the window + dot-product
// polyphase pattern with two branches and no transform, the usual rate changer in front of / behind a channelizer.
//
// Replaces src/filter/src/resamp2.c (create :65-123, recreate :130-168, clear :209-215, get_delay :218-221,
// filter_execute :228-264, analyzer_execute :271-290, synthesizer_execute :297-313, decim_execute :320-338,
// interp_execute :345-359) and src/filter/src/msresamp2.c (create :82-170, reset :213-225, get_delay :231-258,
// execute :264-283, interp :294-323, decim :329-354), both instantiated for crcf by src/filter/src/filter_crcf.c.
//
// The reference keeps two windows of 2m samples: w0 (delay branch, read at element m-1) and w1 (filter branch,
// dotted with the 2m odd taps of the 4m+1 tap prototype).  Here the object's state is exactly those two windows
// (device memory, oldest first) plus the toggle of the filter mode, so the five modes can be mixed on one object
// as in the reference.  A block of n units is one kernel: thread u derives the two branch sequences from the
// input on the fly (a[u] joins w0, b[u] joins w1):
//     decim        b[u] = x[2u]          a[u] = x[2u+1]            y[u]    = a[u-m] + F(u)
//     analyzer     b[u] = x[2u]/2        a[u] = x[2u+1]/2          y[2u]   = F(u) + a[u-m], y[2u+1] = F(u) - a[u-m]
//     synthesizer  a[u] = x[2u]+x[2u+1]  b[u] = x[2u]-x[2u+1]      y[2u]   = a[u-m],        y[2u+1] = F(u)
//     interp       a[u] = b[u] = x[u]                               y[2u]   = a[u-m],        y[2u+1] = F(u)
//     with F(u) = sum_{k<2m} hb[k] b[u-2m+1+k]; indices below zero read the windows.
// The filter mode pushes alternately into w0 and w1 (toggle) and filters the other window; in terms of one
// virtual sample stream: y0/y1[i] = (x[i-2m] +/- sum_k hb[k] x[i-1-2(2m-1-k)]) / 2.  A second tiny kernel
// writes the windows as they stand after the block into the other half of a ping-pong pair.
// msresamp2 runs its stages back to back on the device through two scratch buffers; the decimator's final
// 2^-S scaling rides on the last stage's store.
#include <cmath>
#include <cstdint>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "host_pipeline.cuh"

struct resamp2_crcf_s {
    unsigned int m = 0, len = 0;     // len = 2m taps on the filter branch
    float f0 = 0.f, As = 0.f;
    std::vector<float> h;            // 4m+1 tap prototype
    std::vector<float> hb;           // filter branch: odd taps, reversed
    float *d_hb = nullptr;
    float2 *d_win = nullptr;         // [2][2][len]: ping-pong pair of (w0 | w1), oldest first
    int cur = 0;                     // which half holds the state
    unsigned int toggle = 0;
    lc::DevBuf stage_in, stage_out;
    cudaStream_t stream = nullptr;
};

struct msresamp2_crcf_s {
    int interp = 0;
    unsigned int S = 0, M = 1;
    float fc = 0.f, f0 = 0.f, As = 0.f, zeta = 1.f;
    std::vector<unsigned int> m_stage;
    std::vector<float> fc_stage, f0_stage;
    std::vector<resamp2_crcf> hb;
    lc::DevBuf stage_in, stage_out, scratch0, scratch1;
    cudaStream_t stream = nullptr;
};

namespace {

constexpr int kThreads = 256;
enum Mode { kDecim = 0, kAnalyzer = 1, kSynth = 2, kInterp = 3 };

struct Windows {
    const float2 *w0, *w1;           // len samples each, oldest first
    unsigned int len;
};

__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 half2(float2 a) { return make_float2(0.5f * a.x, 0.5f * a.y); }

// element n of the delay-branch sequence (what the reference pushes into w0 at unit n)
template <int MODE>
__device__ __forceinline__ float2 seq_a(const Windows &w, const float2 *__restrict__ x, long long n)
{
    if (n < 0) return __ldg(w.w0 + ((long long)w.len + n));
    if (MODE == kDecim) return __ldg(x + 2 * n + 1);
    if (MODE == kAnalyzer) return half2(__ldg(x + 2 * n + 1));
    if (MODE == kSynth) return add2(__ldg(x + 2 * n), __ldg(x + 2 * n + 1));
    return __ldg(x + n);
}

// element n of the filter-branch sequence (pushed into w1)
template <int MODE>
__device__ __forceinline__ float2 seq_b(const Windows &w, const float2 *__restrict__ x, long long n)
{
    if (n < 0) return __ldg(w.w1 + ((long long)w.len + n));
    if (MODE == kDecim) return __ldg(x + 2 * n);
    if (MODE == kAnalyzer) return half2(__ldg(x + 2 * n));
    if (MODE == kSynth) return sub2(__ldg(x + 2 * n), __ldg(x + 2 * n + 1));
    return __ldg(x + n);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
resamp2_units_kernel(Windows w, const float *__restrict__ hb, unsigned int m, const float2 *__restrict__ x,
                     float2 *__restrict__ y, unsigned long long n, float scale, bool y_aligned16)
{
    extern __shared__ float hs[];
    for (unsigned int k = threadIdx.x; k < w.len; k += blockDim.x) hs[k] = __ldg(hb + k);
    __syncthreads();
    for (unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < n;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        const float2 d = seq_a<MODE>(w, x, (long long)u - (long long)m);
        float re = 0.f, im = 0.f;
        const long long b0 = (long long)u - (long long)w.len + 1;
        for (unsigned int k = 0; k < w.len; k++) {
            const float2 s = seq_b<MODE>(w, x, b0 + k);
            re = fmaf(hs[k], s.x, re);
            im = fmaf(hs[k], s.y, im);
        }
        const float2 f = make_float2(re, im);
        if (MODE == kDecim) {
            const float2 v = add2(d, f);
            y[u] = make_float2(v.x * scale, v.y * scale);
        } else {
            const float2 o0 = (MODE == kAnalyzer) ? add2(f, d) : d;
            const float2 o1 = (MODE == kAnalyzer) ? sub2(f, d) : f;
            if (y_aligned16) {
                reinterpret_cast<float4 *>(y)[u] = make_float4(o0.x, o0.y, o1.x, o1.y);
            } else {
                y[2 * u] = o0;
                y[2 * u + 1] = o1;
            }
        }
    }
}

// two adjacent units (t, t + 1), t even, from a staged filter-branch sequence: unit t's taps start at element t.
// Packed FFMA2 on (re, im): the same two IEEE fused multiply-adds per tap as the scalar chains, in the same order.
__device__ __forceinline__ void pair_fir(const float4 *__restrict__ sb4, unsigned int t2, const float *__restrict__ hs,
                                         unsigned int len, float2 &acc0, float2 &acc1)
{
    const float2 *hs2 = reinterpret_cast<const float2 *>(hs);
    acc0 = make_float2(0.f, 0.f);
    acc1 = make_float2(0.f, 0.f);
    float4 cur = sb4[t2];                                                // elements t, t + 1
    for (unsigned int k2 = 0; k2 < len / 2; k2++) {
        const float4 nxt = sb4[t2 + k2 + 1];
        const float2 h = hs2[k2];                                        // taps 2 k2, 2 k2 + 1
        const float2 h0 = make_float2(h.x, h.x), h1 = make_float2(h.y, h.y);
        acc0 = __ffma2_rn(h0, make_float2(cur.x, cur.y), acc0);
        acc0 = __ffma2_rn(h1, make_float2(cur.z, cur.w), acc0);
        acc1 = __ffma2_rn(h0, make_float2(cur.z, cur.w), acc1);
        acc1 = __ffma2_rn(h1, make_float2(nxt.x, nxt.y), acc1);
        cur = nxt;
    }
}

__device__ __forceinline__ void cp_async8(float2 *dst_smem, const float2 *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned int)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// The throughput form of the above (16-byte aligned x and y): a CTA stages the two branch sequences of kTile units
// plus 2m - 1 units of history in shared memory -- one coalesced 16-byte load per input pair, every input byte
// read from HBM once -- and each thread then produces two adjacent units from 16-byte shared-memory reads
// (consecutive lanes, consecutive addresses).  Same fmaf chain per output as resamp2_units_kernel: identical bits.
constexpr unsigned int kTile = 2048;

__host__ __device__ inline unsigned int tile_hs_f2(unsigned int len) { return (len / 2 + 1) & ~1u; }   // taps, in float2, even
__host__ __device__ inline unsigned int tile_sb_f2(unsigned int len) { return (kTile + len + 2 + 1) & ~1u; }
inline size_t tile_smem(unsigned int len) { return sizeof(float2) * (tile_hs_f2(len) + 2 * tile_sb_f2(len)); }

template <int MODE>
__global__ void __launch_bounds__(kThreads)
resamp2_tile_kernel(Windows w, const float *__restrict__ hb, unsigned int m, const float2 *__restrict__ x,
                    float2 *__restrict__ y, unsigned long long n, float scale)
{
    extern __shared__ float4 smem_f4[];
    float2 *sm = reinterpret_cast<float2 *>(smem_f4);
    float *hs = reinterpret_cast<float *>(sm);
    float2 *sb = sm + tile_hs_f2(w.len);
    float2 *sa = sb + tile_sb_f2(w.len);
    const unsigned int halo = w.len - 1;
    const unsigned long long u0 = (unsigned long long)blockIdx.x * kTile;
    const unsigned int nt = (n - u0 < kTile) ? (unsigned int)(n - u0) : kTile;
    for (unsigned int k = threadIdx.x; k < w.len; k += blockDim.x) hs[k] = __ldg(hb + k);
    // staged element q holds unit u0 - halo + q
    for (unsigned int q = threadIdx.x; q < halo + nt; q += blockDim.x) {
        const long long idx = (long long)u0 - (long long)halo + q;
        if (idx < 0) {
            sb[q] = __ldg(w.w1 + ((long long)w.len + idx));
            sa[q] = __ldg(w.w0 + ((long long)w.len + idx));
        } else if (MODE == kInterp) {
            cp_async8(sb + q, x + idx);
            cp_async8(sa + q, x + idx);
        } else if (MODE == kDecim) {
            cp_async8(sb + q, x + 2 * idx);
            cp_async8(sa + q, x + 2 * idx + 1);
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + idx);      // (x[2 idx], x[2 idx + 1])
            const float2 e = make_float2(v.x, v.y), o = make_float2(v.z, v.w);
            if (MODE == kAnalyzer) { sb[q] = half2(e); sa[q] = half2(o); }
            if (MODE == kSynth) { sa[q] = add2(e, o); sb[q] = sub2(e, o); }
        }
    }
    cp_async_wait_all();
    __syncthreads();
    const float4 *sb4 = reinterpret_cast<const float4 *>(sb);
    for (unsigned int t2 = threadIdx.x; 2 * t2 < nt; t2 += blockDim.x) {
        const unsigned int t = 2 * t2;                                   // units u0 + t and u0 + t + 1
        float2 f0, f1;
        pair_fir(sb4, t2, hs, w.len, f0, f1);
        const float re0 = f0.x, im0 = f0.y, re1 = f1.x, im1 = f1.y;
        const float2 d0 = sa[halo + t - m], d1 = sa[halo + t + 1 - m];
        const bool second = (t + 1 < nt);
        const unsigned long long u = u0 + t;
        if (MODE == kDecim) {
            const float2 v0 = make_float2((d0.x + re0) * scale, (d0.y + im0) * scale);
            const float2 v1 = make_float2((d1.x + re1) * scale, (d1.y + im1) * scale);
            if (second) *reinterpret_cast<float4 *>(y + u) = make_float4(v0.x, v0.y, v1.x, v1.y);
            else        y[u] = v0;
        } else if (MODE == kAnalyzer) {
            reinterpret_cast<float4 *>(y)[u] = make_float4(re0 + d0.x, im0 + d0.y, re0 - d0.x, im0 - d0.y);
            if (second) reinterpret_cast<float4 *>(y)[u + 1] = make_float4(re1 + d1.x, im1 + d1.y, re1 - d1.x, im1 - d1.y);
        } else {
            reinterpret_cast<float4 *>(y)[u] = make_float4(d0.x, d0.y, re0, im0);
            if (second) reinterpret_cast<float4 *>(y)[u + 1] = make_float4(d1.x, d1.y, re1, im1);
        }
    }
}

// the windows after n units, written to the other half of the ping-pong pair
template <int MODE>
__global__ void resamp2_units_roll_kernel(Windows w, const float2 *__restrict__ x, unsigned long long n, float2 *__restrict__ next)
{
    const unsigned int j = threadIdx.x;
    if (j >= w.len) return;
    const long long idx = (long long)n - (long long)w.len + j;
    next[j] = seq_a<MODE>(w, x, idx);
    next[w.len + j] = seq_b<MODE>(w, x, idx);
}

// filter mode: sample idx of the virtual stream in which the samples alternate between the two windows; sample i
// joins window (g + i) & 1
__device__ __forceinline__ float2 filt_val(const Windows &w, const float2 *__restrict__ x, long long idx, unsigned int g)
{
    if (idx >= 0) return __ldg(x + idx);
    const unsigned int c = (unsigned int)(((long long)g + idx) & 1);
    const long long k = (-idx - 1) >> 1;
    return __ldg((c ? w.w1 : w.w0) + ((long long)w.len - 1 - k));
}

__global__ void __launch_bounds__(kThreads)
resamp2_filter_kernel(Windows w, const float *__restrict__ hb, unsigned int m, unsigned int g, const float2 *__restrict__ x,
                      float2 *__restrict__ y0, float2 *__restrict__ y1, unsigned long long n)
{
    extern __shared__ float hs[];
    for (unsigned int k = threadIdx.x; k < w.len; k += blockDim.x) hs[k] = __ldg(hb + k);
    __syncthreads();
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float2 d = filt_val(w, x, (long long)i - 2ll * m, g);
        float re = 0.f, im = 0.f;
        const long long b0 = (long long)i - 1 - 2ll * (w.len - 1);
        for (unsigned int k = 0; k < w.len; k++) {
            const float2 s = filt_val(w, x, b0 + 2ll * k, g);
            re = fmaf(hs[k], s.x, re);
            im = fmaf(hs[k], s.y, im);
        }
        y0[i] = make_float2(0.5f * (d.x + re), 0.5f * (d.y + im));
        y1[i] = make_float2(0.5f * (d.x - re), 0.5f * (d.y - im));
    }
}

__global__ void resamp2_filter_roll_kernel(Windows w, unsigned int g, const float2 *__restrict__ x, unsigned long long n,
                                           float2 *__restrict__ next)
{
    const unsigned int j = threadIdx.x;
    if (j >= 2 * w.len) return;
    const unsigned int c = j / w.len, pos = j % w.len;          // window, element (oldest first)
    const long long last = (long long)n - 1;
    const long long newest = ((((long long)g + last) & 1) == (long long)c) ? last : last - 1;
    next[j] = filt_val(w, x, newest - 2ll * (w.len - 1 - pos), g);
}

// ---- msresamp2 decimator, all stages in ONE kernel -------------------------------------------------------------
// A CTA produces kT final outputs.  Working backwards, stage s must compute the units [lo_s, hi_s) with
// lo_{s-1} = 2 (lo_s - (len_s - 1)), hi_{s-1} = 2 hi_s: each stage recomputes the 2m - 1 units of history the next one
// needs (a few percent of extra arithmetic) instead of reading them back from HBM, so the input is read once and only
// the final stream is written.  A stage's two branch sequences live de-interleaved in shared memory (the pair of
// outputs (u, u+1), u even, is the next stage's (b, a) pair at unit u/2), units below zero come from that stage's
// windows, and the last CTA leaves every stage's windows as they stand after the block.
// The inner loop is register-blocked: a thread owns 8 adjacent units, the filter length is a template parameter and
// the tap loop is fully unrolled, so each staged element is read from shared memory once per thread (16-byte reads;
// the arrays are padded by 16 bytes per 64 so that the 64-byte lane stride is conflict-free) and feeds up to 8 packed
// FFMA2.  Same fused multiply-add chains and the same d + f, * 2^-S roundings as the stage-by-stage form: identical bits.
constexpr int kMaxFusedStages = 8;
constexpr unsigned int kMaxFusedLen = 32;
constexpr unsigned int kBlockUnits = 8;

struct CascadeStage {
    const float2 *w0, *w1;        // current windows (oldest first)
    float2 *next;                 // the other half of the ping-pong pair: (w0 | w1) after the block
    const float *hb;
    unsigned int len, m;
};

struct CascadeArgs {
    CascadeStage st[kMaxFusedStages];
    unsigned int S, T;            // stages, final outputs per CTA (a multiple of 8)
    unsigned int cap[2];          // float2 capacity (padded layout) of one staged array in region 0 / 1
    unsigned int back[kMaxFusedStages];   // stage s computes the units [(o_lo << (S-1-s)) - back[s], o_hi << (S-1-s))
    float zeta;
};

// padded position of staged element q: 2 elements (16 bytes) of padding after every 8
__host__ __device__ inline unsigned int phys(unsigned int q) { return q + 2 * (q >> 3); }

// units of stage s the tile whose final outputs are [o_lo, o_hi) must compute
__host__ __device__ inline void cascade_range(const CascadeArgs &a, unsigned int s, long long o_lo, long long o_hi,
                                              long long &lo, long long &hi)
{
    lo = o_lo;
    hi = o_hi;
    for (unsigned int j = a.S - 1; j > s; j--) {
        lo = 2 * (lo - (long long)(a.st[j].len - 1));
        hi = 2 * hi;
    }
}

// One stage over the units [lo, hi) of a tile.  Staged element q of sb / sa is unit lo - (LEN - 1) + q of the
// filter / delay branch.  Units below zero (start of the stream) are skipped: the next stage reads its windows there.
template <int LEN, bool FINAL>
__device__ __forceinline__ void cascade_stage(const float2 *__restrict__ sb, const float2 *__restrict__ sa,
                                              const float *__restrict__ hs, unsigned int m, long long lo, unsigned int cnt,
                                              float2 *__restrict__ nb, float2 *__restrict__ na, float2 *__restrict__ y,
                                              float zeta)
{
    constexpr int R = kBlockUnits;
    constexpr int NE = R + LEN;                                       // elements a thread reads (one spare: LEN + R - 1 used)
    const unsigned int first = lo < 0 ? (unsigned int)(-lo) : 0;      // first unit (relative to lo) that exists
    for (unsigned int t = (first & ~(R - 1u)) + R * threadIdx.x; t < cnt; t += R * blockDim.x) {
        float2 e[NE];
        const float4 *p = reinterpret_cast<const float4 *>(sb + phys(t));   // t is a multiple of 8: 16-byte aligned
#pragma unroll
        for (int j = 0; j < NE / 2; j++) {
            // elements t + 2j, t + 2j + 1: pairs never straddle the padding
            const float4 v = p[(2 * j + 2 * ((2 * j) >> 3)) / 2];
            e[2 * j] = make_float2(v.x, v.y);
            e[2 * j + 1] = make_float2(v.z, v.w);
        }
        float2 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < LEN; k++) {
            const float2 h = make_float2(hs[k], hs[k]);
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = __ffma2_rn(h, e[k + r], acc[r]);
        }
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = add2(sa[phys(LEN - 1 + t + r - m)], acc[r]);
        const bool whole = (t >= first) && (t + R <= cnt);
        if (!FINAL) {
            // outputs (t + 2j, t + 2j + 1) are the next stage's (b, a) pair at element t/2 + j
            const unsigned int q = phys(t / 2);                       // t/2 is a multiple of 4: 16-byte aligned, no padding inside
            if (whole) {
                *reinterpret_cast<float4 *>(nb + q) = make_float4(v[0].x, v[0].y, v[2].x, v[2].y);
                *reinterpret_cast<float4 *>(nb + q + 2) = make_float4(v[4].x, v[4].y, v[6].x, v[6].y);
                *reinterpret_cast<float4 *>(na + q) = make_float4(v[1].x, v[1].y, v[3].x, v[3].y);
                *reinterpret_cast<float4 *>(na + q + 2) = make_float4(v[5].x, v[5].y, v[7].x, v[7].y);
            } else {
#pragma unroll
                for (int j = 0; j < R / 2; j++)
                    if (t + 2 * j >= first && t + 2 * j < cnt) {      // first and cnt are even here
                        nb[q + j] = v[2 * j];
                        na[q + j] = v[2 * j + 1];
                    }
            }
        } else {
            float2 *yo = y + lo + t;                                  // lo >= 0 on the final stage, a multiple of 8
            if (whole) {
#pragma unroll
                for (int j = 0; j < R / 2; j++)
                    reinterpret_cast<float4 *>(yo)[j] = make_float4(v[2 * j].x * zeta, v[2 * j].y * zeta,
                                                                    v[2 * j + 1].x * zeta, v[2 * j + 1].y * zeta);
            } else {
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (t + r < cnt) yo[r] = make_float2(v[r].x * zeta, v[r].y * zeta);
            }
        }
    }
}

template <bool FINAL>
__device__ __forceinline__ void cascade_stage_any(unsigned int len, const float2 *sb, const float2 *sa, const float *hs,
                                                  unsigned int m, long long lo, unsigned int cnt, float2 *nb, float2 *na,
                                                  float2 *y, float zeta)
{
    switch (len) {
#define LC_STAGE(L) case L: cascade_stage<L, FINAL>(sb, sa, hs, m, lo, cnt, nb, na, y, zeta); break;
        LC_STAGE(4) LC_STAGE(6) LC_STAGE(8) LC_STAGE(10) LC_STAGE(12) LC_STAGE(14) LC_STAGE(16) LC_STAGE(18)
        LC_STAGE(20) LC_STAGE(22) LC_STAGE(24) LC_STAGE(26) LC_STAGE(28) LC_STAGE(30) LC_STAGE(32)
#undef LC_STAGE
    default: break;                                                   // the host only launches supported lengths
    }
}

__global__ void __launch_bounds__(kThreads)
msresamp2_decim_fused_kernel(CascadeArgs a, const float2 *__restrict__ x, float2 *__restrict__ y, unsigned long long n_out)
{
    extern __shared__ float4 smem_f4[];
    float *hs = reinterpret_cast<float *>(smem_f4);                        // kMaxFusedLen taps of the running stage
    float2 *sm = reinterpret_cast<float2 *>(smem_f4) + kMaxFusedLen / 2;   // region 0 = [0, 2 cap0), region 1 after it
    const long long o_lo = (long long)blockIdx.x * a.T;
    const long long o_hi = (o_lo + a.T < (long long)n_out) ? o_lo + a.T : (long long)n_out;
    const bool last_cta = (blockIdx.x == gridDim.x - 1);

    // stage 0's sequences from the input: pair idx = (x[2 idx], x[2 idx + 1]) = (b, a); every load of a thread is in
    // flight at once (cp.async), nothing passes through registers
    {
        const long long lo = (o_lo << (a.S - 1)) - (long long)a.back[0], hi = o_hi << (a.S - 1);
        const CascadeStage &g = a.st[0];
        const long long org = lo - (long long)(g.len - 1);
        float2 *sb = sm, *sa = sm + a.cap[0];
        const unsigned int cnt = (unsigned int)(hi - org);
        if (org >= 0) {
            // 32-bit shared addresses, one 64-bit pointer bump per pair
            const float2 *src = x + 2 * org + 2 * threadIdx.x;
            const unsigned int sb_addr = (unsigned int)__cvta_generic_to_shared(sb);
            const unsigned int a_off = a.cap[0] * (unsigned int)sizeof(float2);
#pragma unroll 4
            for (unsigned int q = threadIdx.x; q < cnt; q += kThreads, src += 2 * kThreads) {
                const unsigned int dst = sb_addr + phys(q) * (unsigned int)sizeof(float2);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + a_off), "l"(src + 1) : "memory");
            }
        } else {
            for (unsigned int q = threadIdx.x; q < cnt; q += blockDim.x) {
                const long long idx = org + q;
                if (idx >= 0) {
                    cp_async8(sb + phys(q), x + 2 * idx);
                    cp_async8(sa + phys(q), x + 2 * idx + 1);
                } else if (idx >= -(long long)g.len) {
                    // (the staged range reaches below the window near the start of the stream; those elements
                    // are never used: the first unit computed is unit 0)
                    sb[phys(q)] = __ldg(g.w1 + ((long long)g.len + idx));
                    sa[phys(q)] = __ldg(g.w0 + ((long long)g.len + idx));
                }
            }
        }
        cp_async_wait_all();
    }
    for (unsigned int s = 0; s < a.S; s++) {
        const CascadeStage &g = a.st[s];
        const long long lo = (o_lo << (a.S - 1 - s)) - (long long)a.back[s], hi = o_hi << (a.S - 1 - s);
        const unsigned int r = s & 1;
        const unsigned int cur_off = r ? 2 * a.cap[0] : 0, nxt_off = r ? 0 : 2 * a.cap[0];
        const float2 *sb = sm + cur_off, *sa = sm + cur_off + a.cap[r];
        float2 *nb = sm + nxt_off, *na = sm + nxt_off + a.cap[r ^ 1];
        const bool final_stage = (s == a.S - 1);
        for (unsigned int k = threadIdx.x; k < g.len; k += blockDim.x) hs[k] = __ldg(g.hb + k);
        if (!final_stage && lo < 0) {
            // the next stage's units below zero: its windows (unit org' + q with org' = lo / 2)
            const CascadeStage &gn = a.st[s + 1];
            const long long orgn = lo / 2;
            for (long long q = threadIdx.x; orgn + q < 0; q += blockDim.x) {
                if (orgn + q < -(long long)gn.len) continue;
                nb[phys((unsigned int)q)] = __ldg(gn.w1 + ((long long)gn.len + orgn + q));
                na[phys((unsigned int)q)] = __ldg(gn.w0 + ((long long)gn.len + orgn + q));
            }
        }
        __syncthreads();                                                   // this stage's sequences and taps are in place
        if (last_cta) {
            // the windows after the block: the last len units of this stage's sequences
            const long long org = lo - (long long)(g.len - 1);
            for (unsigned int j = threadIdx.x; j < g.len; j += blockDim.x) {
                const unsigned int q = phys((unsigned int)(hi - (long long)g.len + j - org));
                g.next[j] = sa[q];
                g.next[g.len + j] = sb[q];
            }
        }
        const unsigned int cnt = (unsigned int)(hi - lo);
        if (final_stage) cascade_stage_any<true>(g.len, sb, sa, hs, g.m, lo, cnt, nb, na, y, a.zeta);
        else             cascade_stage_any<false>(g.len, sb, sa, hs, g.m, lo, cnt, nb, na, y, a.zeta);
        __syncthreads();                                                   // hs and this stage's region are free again
    }
}

// padded float2 capacity one array of stage s needs for a full tile (with the slack of the blocked reads)
inline unsigned int cascade_capacity(const CascadeArgs &a, unsigned int s)
{
    long long lo, hi;
    cascade_range(a, s, 0, a.T, lo, hi);
    const unsigned int elems = (unsigned int)(hi - lo) + 2 * kBlockUnits + a.st[s].len + 2;
    return (phys(elems) + 3) & ~3u;
}

inline unsigned int grid_for(unsigned long long total)
{
    unsigned long long b = (total + kThreads - 1) / kThreads;
    const unsigned long long cap = 148ull * 8 * 16;
    return (unsigned int)(b < cap ? (b ? b : 1) : cap);
}

Windows windows_of(resamp2_crcf q)
{
    const float2 *base = q->d_win + (size_t)q->cur * 2 * q->len;
    return Windows{base, base + q->len, q->len};
}

float2 *next_windows(resamp2_crcf q) { return q->d_win + (size_t)(q->cur ^ 1) * 2 * q->len; }

template <int MODE>
void units_dev(resamp2_crcf q, const float2 *d_x, unsigned long long n, float2 *d_y, float scale, cudaStream_t st)
{
    if (n == 0) return;
    const Windows w = windows_of(q);
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_y)) & 15) == 0;
    const size_t smem = tile_smem(q->len);
    const unsigned long long tiles = (n + kTile - 1) / kTile;
    if (aligned && n >= 4 * kTile && smem <= 48 * 1024 && tiles < (1ull << 31) && !getenv("LIQUID_CUDA_RESAMP2_SIMPLE")) {
        resamp2_tile_kernel<MODE><<<(unsigned int)tiles, kThreads, smem, st>>>(w, q->d_hb, q->m, d_x, d_y, n, scale);
    } else {
        resamp2_units_kernel<MODE><<<grid_for(n), kThreads, sizeof(float) * q->len, st>>>(
            w, q->d_hb, q->m, d_x, d_y, n, scale, (reinterpret_cast<uintptr_t>(d_y) & 15) == 0);
    }
    LC_LAUNCH_CHECK();
    resamp2_units_roll_kernel<MODE><<<1, (q->len + 31) / 32 * 32, 0, st>>>(w, d_x, n, next_windows(q));
    LC_LAUNCH_CHECK();
    q->cur ^= 1;
}

void filter_dev(resamp2_crcf q, const float2 *d_x, unsigned long long n, float2 *d_y0, float2 *d_y1, cudaStream_t st)
{
    if (n == 0) return;
    const Windows w = windows_of(q);
    resamp2_filter_kernel<<<grid_for(n), kThreads, sizeof(float) * q->len, st>>>(w, q->d_hb, q->m, q->toggle, d_x, d_y0, d_y1, n);
    LC_LAUNCH_CHECK();
    resamp2_filter_roll_kernel<<<1, (2 * q->len + 31) / 32 * 32, 0, st>>>(w, q->toggle, d_x, n, next_windows(q));
    LC_LAUNCH_CHECK();
    q->cur ^= 1;
    q->toggle = (unsigned int)((q->toggle + n) & 1);
}

void design(resamp2_crcf q)
{
    // 4m+1 taps sinc(t/2) kaiser(t) cos(2 pi f0 t) in float, the phase in double as in resamp2.c:94-108;
    // the filter branch keeps the odd taps, reversed (:110-113)
    const unsigned int n = 4 * q->m + 1;
    q->h.resize(n);
    const float beta = lc::kaiser_beta_As(q->As);
    for (unsigned int i = 0; i < n; i++) {
        const float t = (float)i - (float)(n - 1) / 2.0f;
        const float a = lc::sinc_value(t / 2.0f);
        const float b = lc::kaiser_window(i, n, beta, 0);
        const float c = cosf((float)(2.0 * M_PI * (double)t * (double)q->f0));
        q->h[i] = a * b * c;
    }
    q->hb.resize(q->len);
    unsigned int j = 0;
    for (unsigned int i = 1; i < n; i += 2) q->hb[j++] = q->h[n - i - 1];
}

void check_args(const char *fn, unsigned int m, float f0)
{
    if (m < 2) lc::fatal(fn, "filter semi-length must be at least 2");
    if (f0 < -0.5f || f0 > 0.5f) {
        fprintf(stderr, "error: %s(), f0 (%12.4e) must be in (-1,1)\n", fn, f0);
        exit(1);
    }
}

void allocate(resamp2_crcf q)
{
    LC_CUDA_CHECK(cudaMalloc(&q->d_hb, sizeof(float) * q->len));
    LC_CUDA_CHECK(cudaMemcpy(q->d_hb, q->hb.data(), sizeof(float) * q->len, cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMalloc(&q->d_win, sizeof(float2) * 4 * q->len));
}

// small host-pointer calls (the reference's per-sample entry points): stage, run, copy back
template <typename Run>
void host_small(resamp2_crcf q, const liquid_float_complex *x, size_t n_in, liquid_float_complex *y, size_t n_out, Run run)
{
    float2 *din = (float2 *)q->stage_in.get(sizeof(float2) * n_in);
    float2 *dout = (float2 *)q->stage_out.get(sizeof(float2) * n_out);
    LC_CUDA_CHECK(cudaMemcpyAsync(din, x, sizeof(float2) * n_in, cudaMemcpyHostToDevice, q->stream));
    run(din, dout, q->stream);
    LC_CUDA_CHECK(cudaMemcpyAsync(y, dout, sizeof(float2) * n_out, cudaMemcpyDeviceToHost, q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(q->stream));
}

// the whole decimating cascade in one launch; false when the shape is left to the stage-by-stage form
bool decim_fused_dev(msresamp2_crcf q, const float2 *d_x, unsigned long long n_out, float2 *d_y, cudaStream_t st)
{
    if (q->S < 2 || q->S > (unsigned int)kMaxFusedStages || getenv("LIQUID_CUDA_MSRESAMP2_STAGED")) return false;
    if (((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_y)) & 15) != 0) return false;
    CascadeArgs a;
    a.S = q->S;
    a.zeta = q->zeta;
    for (unsigned int s = 0; s < q->S; s++) {
        resamp2_crcf h = q->hb[s];
        if (h->len > kMaxFusedLen || h->len < 4) return false;
        const Windows w = windows_of(h);
        a.st[s] = CascadeStage{w.w0, w.w1, next_windows(h), h->d_hb, h->len, h->m};
    }
    for (unsigned int s = 0; s < q->S; s++) {
        long long lo, hi;
        cascade_range(a, s, 0, 0, lo, hi);
        a.back[s] = (unsigned int)(-lo);
    }
    // final outputs per CTA: a multiple of 8, the largest for which stage 0 (the widest) is at most one 8-unit task
    // per thread and the staged sequences fit ~72 KB (three CTAs per SM)
    size_t smem = 0;
    unsigned int best = 0;
    for (unsigned int T = 8; T <= 2048; T += 8) {
        a.T = T;
        long long lo, hi;
        cascade_range(a, 0, 0, T, lo, hi);
        if ((unsigned long long)(hi - lo) > (unsigned long long)kBlockUnits * kThreads) break;
        a.cap[0] = a.cap[1] = 0;
        for (unsigned int s = 0; s < q->S; s++) a.cap[s & 1] = std::max(a.cap[s & 1], cascade_capacity(a, s));
        if (sizeof(float) * kMaxFusedLen + sizeof(float2) * 2 * ((size_t)a.cap[0] + a.cap[1]) > 72 * 1024) break;
        best = T;
    }
    if (best == 0) return false;
    a.T = best;
    a.cap[0] = a.cap[1] = 0;
    for (unsigned int s = 0; s < q->S; s++) a.cap[s & 1] = std::max(a.cap[s & 1], cascade_capacity(a, s));
    smem = sizeof(float) * kMaxFusedLen + sizeof(float2) * 2 * ((size_t)a.cap[0] + a.cap[1]);
    if (n_out < 8ull * a.T) return false;
    const unsigned long long tiles = (n_out + a.T - 1) / a.T;
    if (tiles >= (1ull << 31)) return false;
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
    {
        LC_CUDA_CHECK(cudaFuncSetAttribute(msresamp2_decim_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        // several CTAs per SM hide each other's load and barrier phases: ask for the full shared-memory carve-out
        LC_CUDA_CHECK(cudaFuncSetAttribute(msresamp2_decim_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
    }
    msresamp2_decim_fused_kernel<<<(unsigned int)tiles, kThreads, smem, st>>>(a, d_x, d_y, n_out);
    LC_LAUNCH_CHECK();
    for (unsigned int s = 0; s < q->S; s++) q->hb[s]->cur ^= 1;
    return true;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------ resamp2_crcf
resamp2_crcf resamp2_crcf_create(unsigned int _m, float _f0, float _As)
{
    check_args("resamp2_crcf_create", _m, _f0);
    lc::ensure_device();
    resamp2_crcf q = new resamp2_crcf_s();
    q->m = _m;
    q->len = 2 * _m;
    q->f0 = _f0;
    q->As = _As;
    design(q);
    allocate(q);
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    resamp2_crcf_clear(q);
    return q;
}

resamp2_crcf resamp2_crcf_recreate(resamp2_crcf _q, unsigned int _m, float _f0, float _As)
{
    // a new length rebuilds the object; otherwise only the taps change and the windows are kept (resamp2.c:130-168)
    if (_m != _q->m) {
        resamp2_crcf_destroy(_q);
        return resamp2_crcf_create(_m, _f0, _As);
    }
    check_args("resamp2_crcf_recreate", _m, _f0);
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    _q->f0 = _f0;
    _q->As = _As;
    design(_q);
    LC_CUDA_CHECK(cudaMemcpy(_q->d_hb, _q->hb.data(), sizeof(float) * _q->len, cudaMemcpyHostToDevice));
    return _q;
}

void resamp2_crcf_destroy(resamp2_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    cudaFree(_q->d_hb);
    cudaFree(_q->d_win);
    _q->stage_in.release();
    _q->stage_out.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void resamp2_crcf_print(resamp2_crcf _q)
{
    printf("fir half-band resampler: [%u taps, f0=%12.8f] (cuda)\n", 4 * _q->m + 1, _q->f0);
    for (unsigned int i = 0; i < 4 * _q->m + 1; i++) printf("  h(%4u) = %12.8f;\n", i + 1, _q->h[i]);
    printf("---\n");
    for (unsigned int i = 0; i < _q->len; i++) printf("  h1(%4u) = %12.8f;\n", i + 1, _q->hb[i]);
}

void resamp2_crcf_clear(resamp2_crcf _q)
{
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_win, 0, sizeof(float2) * 4 * _q->len, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    _q->cur = 0;
    _q->toggle = 0;
}

unsigned int resamp2_crcf_get_delay(resamp2_crcf _q) { return 2 * _q->m - 1; }

void resamp2_crcf_get_taps(resamp2_crcf _q, float *_h, float *_h1)
{
    if (_h) std::copy(_q->h.begin(), _q->h.end(), _h);
    if (_h1) std::copy(_q->hb.begin(), _q->hb.end(), _h1);
}

// device-pointer block forms; the state is carried in stream order
void resamp2_crcf_filter_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y0,
                                           void *_d_y1, void *_stream)
{
    filter_dev(_q, (const float2 *)_d_x, _n, (float2 *)_d_y0, (float2 *)_d_y1, lc::as_stream(_stream));
}

void resamp2_crcf_analyzer_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    units_dev<kAnalyzer>(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, 1.0f, lc::as_stream(_stream));
}

void resamp2_crcf_synthesizer_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    units_dev<kSynth>(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, 1.0f, lc::as_stream(_stream));
}

void resamp2_crcf_decim_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    units_dev<kDecim>(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, 1.0f, lc::as_stream(_stream));
}

void resamp2_crcf_interp_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    units_dev<kInterp>(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, 1.0f, lc::as_stream(_stream));
}

// host-pointer block forms: _n units (filter: samples; analyzer / synthesizer: pairs in and out; decim: outputs;
// interp: inputs)
void resamp2_crcf_filter_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n,
                                       liquid_float_complex *_y0, liquid_float_complex *_y1)
{
    const size_t chunk = 1u << 21;
    for (size_t i0 = 0; i0 < _n; i0 += chunk) {
        const size_t nc = std::min(chunk, (size_t)_n - i0);
        float2 *din = (float2 *)_q->stage_in.get(sizeof(float2) * nc);
        float2 *dout = (float2 *)_q->stage_out.get(sizeof(float2) * 2 * nc);
        LC_CUDA_CHECK(cudaMemcpyAsync(din, _x + i0, sizeof(float2) * nc, cudaMemcpyHostToDevice, _q->stream));
        filter_dev(_q, din, nc, dout, dout + nc, _q->stream);
        LC_CUDA_CHECK(cudaMemcpyAsync(_y0 + i0, dout, sizeof(float2) * nc, cudaMemcpyDeviceToHost, _q->stream));
        LC_CUDA_CHECK(cudaMemcpyAsync(_y1 + i0, dout + nc, sizeof(float2) * nc, cudaMemcpyDeviceToHost, _q->stream));
        LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    }
}

void resamp2_crcf_analyzer_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, 2, 2, [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { units_dev<kAnalyzer>(_q, di, nu, dob, 1.0f, st); });
}

void resamp2_crcf_synthesizer_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, 2, 2, [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { units_dev<kSynth>(_q, di, nu, dob, 1.0f, st); });
}

void resamp2_crcf_decim_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, 2, 1, [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { units_dev<kDecim>(_q, di, nu, dob, 1.0f, st); });
}

void resamp2_crcf_interp_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, 1, 2, [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { units_dev<kInterp>(_q, di, nu, dob, 1.0f, st); });
}

// the reference's per-call entry points
void resamp2_crcf_filter_execute(resamp2_crcf _q, liquid_float_complex _x, liquid_float_complex *_y0, liquid_float_complex *_y1)
{
    liquid_float_complex y[2];
    host_small(_q, &_x, 1, y, 2, [&](float2 *di, float2 *dob, cudaStream_t st) { filter_dev(_q, di, 1, dob, dob + 1, st); });
    *_y0 = y[0];
    *_y1 = y[1];
}

void resamp2_crcf_analyzer_execute(resamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    liquid_float_complex y[2];
    host_small(_q, _x, 2, y, 2, [&](float2 *di, float2 *dob, cudaStream_t st) { units_dev<kAnalyzer>(_q, di, 1, dob, 1.0f, st); });
    _y[0] = y[0];
    _y[1] = y[1];
}

void resamp2_crcf_synthesizer_execute(resamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    liquid_float_complex y[2];
    host_small(_q, _x, 2, y, 2, [&](float2 *di, float2 *dob, cudaStream_t st) { units_dev<kSynth>(_q, di, 1, dob, 1.0f, st); });
    _y[0] = y[0];
    _y[1] = y[1];
}

void resamp2_crcf_decim_execute(resamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    host_small(_q, _x, 2, _y, 1, [&](float2 *di, float2 *dob, cudaStream_t st) { units_dev<kDecim>(_q, di, 1, dob, 1.0f, st); });
}

void resamp2_crcf_interp_execute(resamp2_crcf _q, liquid_float_complex _x, liquid_float_complex *_y)
{
    liquid_float_complex y[2];
    host_small(_q, &_x, 1, y, 2, [&](float2 *di, float2 *dob, cudaStream_t st) { units_dev<kInterp>(_q, di, 1, dob, 1.0f, st); });
    _y[0] = y[0];
    _y[1] = y[1];
}

// ------------------------------------------------------------------------------------------ msresamp2_crcf
msresamp2_crcf msresamp2_crcf_create(int _type, unsigned int _num_stages, float _fc, float _f0, float _As)
{
    if (_num_stages > 16) lc::fatal("msresamp2_crcf_create", "number of stages should not exceed 16");
    if (_fc <= 0.0f || _fc >= 0.5f) lc::fatal("msresamp2_crcf_create", "cut-off frequency must be in (0,0.5)");
    if (_fc > 0.45f) {
        fprintf(stderr, "warning: msresamp2_crcf_create(), cut-off frequency greater than 0.45\n");
        fprintf(stderr, "    >> truncating to 0.45\n");
        _fc = 0.45f;
    }
    if (_f0 != 0.) {
        fprintf(stderr, "warning: msresamp2_crcf_create(), non-zero center frequency not yet supported\n");
        _f0 = 0.;
    }
    lc::ensure_device();
    msresamp2_crcf q = new msresamp2_crcf_s();
    q->interp = (_type == LIQUID_RESAMP_INTERP);
    q->S = _num_stages;
    q->M = 1u << _num_stages;
    q->fc = _fc;
    q->f0 = _f0;
    q->As = _As;
    q->zeta = 1.0f / (float)q->M;
    // stage i: cut-off fc / 2^(i+1), transition band (0.5 - fc_i) / 2, semi-length from Kaiser's estimate, at
    // least 3 (msresamp2.c:128-146)
    float fc = _fc, f0 = _f0;
    for (unsigned int i = 0; i < q->S; i++) {
        f0 = 0.5f * f0;
        fc = 0.5f * fc;
        const float ft = (0.5f - fc) / 2.0f;
        const unsigned int h_len = lc::estimate_req_filter_len(ft, _As);
        const unsigned int m = (unsigned int)ceilf((float)(h_len - 1) / 4.0f);
        q->fc_stage.push_back(fc);
        q->f0_stage.push_back(f0);
        q->m_stage.push_back(m < 3 ? 3 : m);
    }
    for (unsigned int i = 0; i < q->S; i++) q->hb.push_back(resamp2_crcf_create(q->m_stage[i], q->f0_stage[i], _As));
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    return q;
}

void msresamp2_crcf_destroy(msresamp2_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    for (resamp2_crcf s : _q->hb) resamp2_crcf_destroy(s);
    _q->stage_in.release();
    _q->stage_out.release();
    _q->scratch0.release();
    _q->scratch1.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void msresamp2_crcf_print(msresamp2_crcf _q)
{
    printf("multi-stage half-band resampler (cuda):\n");
    printf("    type                    : %s\n", _q->interp ? "interpolator" : "decimator");
    printf("    number of stages        : %u stage%s\n", _q->S, _q->S == 1 ? "" : "s");
    printf("    cut-off frequency, fc   : %12.8f Fs\n", _q->fc);
    printf("    center frequency, f0    : %12.8f Fs\n", _q->f0);
    printf("    stop-band attenuation   : %.2f dB\n", _q->As);
    for (unsigned int i = 0; i < _q->S; i++)
        printf("    stage[%2u]  {m=%3u, As=%6.2f dB, fc=%6.3f, f0=%6.3f}\n", i, _q->m_stage[i], _q->As, _q->fc_stage[i], _q->f0_stage[i]);
}

void msresamp2_crcf_reset(msresamp2_crcf _q)
{
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    for (resamp2_crcf s : _q->hb) resamp2_crcf_clear(s);
}

float msresamp2_crcf_get_delay(msresamp2_crcf _q)
{
    float delay = 0;
    for (unsigned int i = 0; i < _q->S; i++) {
        if (_q->interp) {
            delay *= 0.5f;
            delay += _q->m_stage[i];
        } else {
            delay *= 2;
            delay += 2 * _q->m_stage[_q->S - i - 1] - 1;
        }
    }
    return delay;
}

unsigned int msresamp2_crcf_get_stage_semilength(msresamp2_crcf _q, unsigned int _i) { return _q->m_stage[_i]; }

// _n units: interpolator _n inputs -> _n 2^S outputs (stages last to first, msresamp2.c:294-323); decimator
// _n 2^S inputs -> _n outputs, scaled by 2^-S (stages first to last, :329-354)
void msresamp2_crcf_execute_block_dev(msresamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    cudaStream_t st = lc::as_stream(_stream);
    if (_n == 0) return;
    if (_q->S == 0) {
        LC_CUDA_CHECK(cudaMemcpyAsync(_d_y, _d_x, sizeof(float2) * _n, cudaMemcpyDeviceToDevice, st));
        return;
    }
    if (!_q->interp && decim_fused_dev(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, st)) return;
    const size_t half = (size_t)_n * (_q->M / 2);                  // the largest intermediate stream
    float2 *buf[2] = {nullptr, nullptr};
    if (_q->S > 1) {
        buf[0] = (float2 *)_q->scratch0.get(sizeof(float2) * half);
        buf[1] = (float2 *)_q->scratch1.get(sizeof(float2) * half);
    }
    const float2 *in = (const float2 *)_d_x;
    if (_q->interp) {
        unsigned long long cnt = _n;
        for (unsigned int s = 0; s < _q->S; s++) {
            float2 *out = (s == _q->S - 1) ? (float2 *)_d_y : buf[s & 1];
            units_dev<kInterp>(_q->hb[_q->S - s - 1], in, cnt, out, 1.0f, st);
            cnt *= 2;
            in = out;
        }
    } else {
        unsigned long long cnt = _n * _q->M;
        for (unsigned int s = 0; s < _q->S; s++) {
            cnt /= 2;
            const bool last = (s == _q->S - 1);
            float2 *out = last ? (float2 *)_d_y : buf[s & 1];
            units_dev<kDecim>(_q->hb[s], in, cnt, out, last ? _q->zeta : 1.0f, st);
            in = out;
        }
    }
}

void msresamp2_crcf_execute_block(msresamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    const size_t in_per = _q->interp ? 1 : _q->M, out_per = _q->interp ? _q->M : 1;
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, in_per, out_per,
                   [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { msresamp2_crcf_execute_block_dev(_q, di, nu, dob, st); });
}

void msresamp2_crcf_execute(msresamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    msresamp2_crcf_execute_block(_q, _x, 1, _y);
}

}  // extern "C"
