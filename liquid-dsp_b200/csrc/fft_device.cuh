// fft_device.cuh -- the shared-memory FFT used by the batched FFT kernel (fft.cu), by the shape-generic
// one-kernel analyzers (pfb_kernels.cu) and by the periodogram (spgram.cu), and the post-scaling of a
// transform's output.
//
// Power-of-two n, `per_cta` transforms side by side in shared memory.  Stockham autosort passes of radix 16
// (then one pass of radix 8 / 4 / 2 for what is left: 1024 = 16 * 16 * 4, three passes instead of five
// radix-4 ones): a thread takes R points at stride n/R (consecutive threads read consecutive addresses),
// multiplies by W_{R Ns}^(k m) from a table staged in shared memory, does the R-point DFT in registers
// (RegDft, packed FP32x2) and writes them to their autosort positions.  Both buffers are stored with one
// pad element per 16 (pidx), which makes the stride-R writes of the first pass bank-conflict-free as well.
#pragma once
#include <vector>
#include "lc_common.cuh"
#include "fused_common.cuh"

namespace lcfft {

constexpr int kFftThreads = 256;

// post-processing of the last stage: 0 none, 1 divide by a, 2 multiply by a then by b
__device__ __forceinline__ float2 fft_post(float2 v, int mode, float a, float b)
{
    if (mode == 1) { v.x = v.x / a; v.y = v.y / a; }
    else if (mode == 2) { v.x *= a; v.y *= a; v.x *= b; v.y *= b; }
    return v;
}

// physical position of logical element i of a buffer (i counts across the per_cta transforms)
__host__ __device__ __forceinline__ unsigned int pidx(unsigned int i) { return i + (i >> 4); }

// the two ping-pong buffers and the twiddle table, carved out of the kernel's dynamic shared memory
struct Smem {
    float2 *buf0, *buf1, *tw;
};
__host__ __device__ __forceinline__ unsigned int padded_len(unsigned int n, unsigned int per_cta)
{
    return (pidx(per_cta * n - 1) + 2) & ~1u;               // even: keeps the next region 16-byte aligned
}
__device__ __forceinline__ Smem carve(float2 *base, unsigned int n, unsigned int per_cta)
{
    const unsigned int p = padded_len(n, per_cta);
    return Smem{base, base + p, base + 2 * p};
}
// cooperative copy of the plan's PASS-ORDERED table (FftPlanDev::d_twp, n entries: for every pass after the first,
// in order, the R-1 rows W_{R Ns}^(k m), k < Ns, of that pass; zero padded); a barrier must follow before the FFT
// (the callers' own barrier after filling buf0 serves)
// From kTwGlobalFrom points on the FFT-only kernels (fft.cu, fftfilt.cu, ofdmframe.cu; kernel template parameter TWG)
// do NOT stage the table: a CTA holds one or two such transforms, so its table would be as large as its data; the
// passes then read the plan's table in global memory through L1 (stockham_run<SIGN, true>: coalesced, because of
// the pass order) and the shared memory saved goes into occupancy.  The analyzers and the periodogram, whose fill
// phases live on L1 themselves, keep the staged copy (measured: slower without it).
constexpr unsigned int kTwGlobalFrom = 1024;
__device__ __forceinline__ void stage_twiddles(const Smem &s, const float2 *__restrict__ tw, unsigned int n)
{
    for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) s.tw[i] = __ldg(tw + i);
}

// element accessors of a pass: Src()(f, idx) returns logical element idx of transform f of this CTA,
// Dst()(f, idx, v) stores it.  The shared-memory ones address the padded buffers; callers supply their own for
// the first pass (read global memory, taper a segment, compute a polyphase dot product ...) and the last pass
// (write global memory with a scale ...), so a transform makes no extra trip through shared memory at its ends.
struct SmemSrc {
    const float2 *p;
    unsigned int log2n;
    __device__ __forceinline__ float2 operator()(unsigned int f, unsigned int idx) const { return p[pidx((f << log2n) + idx)]; }
};
struct SmemDst {
    float2 *p;
    unsigned int log2n;
    __device__ __forceinline__ void operator()(unsigned int f, unsigned int idx, float2 v) const { p[pidx((f << log2n) + idx)] = v; }
};

// one Stockham autosort pass of radix R over the per_cta transforms: butterfly j of transform f reads elements
// j + m q (q = n / R), multiplies by W_{R Ns}^(k m) (k = j mod Ns), does the R-point DFT in registers and writes
// elements (j - k) R + k + m Ns
template <int R, int SIGN, bool TWG, class Src, class Dst>
__device__ __forceinline__ void stockham_pass(const Src &src, const Dst &dst, const float2 *__restrict__ twp,
                                              unsigned int n, unsigned int log2n, unsigned int Ns, unsigned int per_cta)
{
    constexpr unsigned int kLog2R = (R == 16) ? 4 : (R == 8) ? 3 : (R == 4) ? 2 : 1;
    const unsigned int log2q = log2n - kLog2R;
    const unsigned int q = 1u << log2q;                        // butterflies per transform
    for (unsigned int i = threadIdx.x; i < (per_cta << log2q); i += blockDim.x) {
        const unsigned int f = i >> log2q, j = i & (q - 1);
        const unsigned int k = j & (Ns - 1);
        float2 v[R];
#pragma unroll
        for (int m = 0; m < R; m++) v[m] = src(f, j + m * q);
        if (Ns > 1) {
            // this pass's table holds W_{R Ns}^(k m) at [(m-1) Ns + k]: the lanes of a warp (consecutive k) read
            // consecutive entries
#pragma unroll
            for (int m = 1; m < R; m++) v[m] = lcf::cmul2(v[m], TWG ? __ldg(twp + (m - 1) * Ns + k) : twp[(m - 1) * Ns + k]);
        }
        lcf::RegDft<R, SIGN>::run(v);
        const unsigned int d = (j - k) * R + k;
#pragma unroll
        for (int m = 0; m < R; m++) dst(f, d + m * Ns, v[m]);
    }
}

// all passes of the per_cta transforms: radix 16 while at least 16 points are left, then one pass of radix
// 8 / 4 / 2.  The first pass reads through `first`, the last pass writes through `last`; passes in between
// ping-pong in shared memory with a block barrier after each.  When the transform has more than one pass the
// twiddle table must have been staged (stage_twiddles) -- the barrier after the first pass publishes it.
// TWG: the table pointer s.tw is the plan's table in GLOBAL memory (read through L1, coalesced thanks to the pass
// order) instead of a staged copy -- for long transforms, where a CTA's table would be as large as its data
template <int SIGN, bool TWG = false, class Src, class Dst>
__device__ __forceinline__ void stockham_run(const Smem &s, unsigned int n, unsigned int log2n, unsigned int per_cta,
                                             const Src &first, const Dst &last)
{
    float2 *in = s.buf0, *out = s.buf1;
    unsigned int Ns = 1, left = log2n;
    bool is_first = true;
    const float2 *twp = s.tw;                                  // table of the running pass (pass-ordered layout)
#define LC_PASS(R)                                                                                               \
    do {                                                                                                         \
        const bool is_last = (left == ((R == 16) ? 4u : (R == 8) ? 3u : (R == 4) ? 2u : 1u));                    \
        const SmemSrc ssrc{in, log2n};                                                                           \
        const SmemDst sdst{out, log2n};                                                                          \
        if (is_first && is_last)  stockham_pass<R, SIGN, TWG>(first, last, twp, n, log2n, Ns, per_cta);               \
        else if (is_first)        stockham_pass<R, SIGN, TWG>(first, sdst, twp, n, log2n, Ns, per_cta);               \
        else if (is_last)         stockham_pass<R, SIGN, TWG>(ssrc, last, twp, n, log2n, Ns, per_cta);                \
        else                      stockham_pass<R, SIGN, TWG>(ssrc, sdst, twp, n, log2n, Ns, per_cta);                \
        if (!is_last) __syncthreads();                                                                           \
    } while (0)
    while (left >= 4) {
        LC_PASS(16);
        if (Ns > 1) twp += 15 * Ns;
        Ns <<= 4;
        left -= 4;
        is_first = false;
        float2 *t = in; in = out; out = t;
    }
    if (left == 3)      LC_PASS(8);
    else if (left == 2) LC_PASS(4);
    else if (left == 1) LC_PASS(2);
#undef LC_PASS
}

// the same with the result left in shared memory (padded layout, natural order): returns the buffer that
// holds it; a block barrier has been passed when it returns
template <int SIGN, bool TWG = false, class Src>
__device__ __forceinline__ const float2 *stockham_run_to_smem(const Smem &s, unsigned int n, unsigned int log2n,
                                                              unsigned int per_cta, const Src &first)
{
    // the buffer the last pass writes: passes alternate buf1, buf0, buf1, ...
    const unsigned int n_passes = (log2n + 3) / 4;
    float2 *res = (n_passes & 1) ? s.buf1 : s.buf0;
    stockham_run<SIGN, TWG>(s, n, log2n, per_cta, first, SmemDst{res, log2n});
    __syncthreads();
    return res;
}

// host: the pass-ordered table from the flat one tw[i] = e^{sign j 2 pi i / n} (same values, so results do not change;
// the flat table read at k m tstride made the lanes of a warp hit one bank in the middle passes of long transforms)
inline std::vector<float2> pass_ordered_twiddles(const std::vector<float2> &flat, unsigned int n)
{
    std::vector<float2> t;
    unsigned int left = 0;
    while ((1u << left) < n) left++;
    unsigned int Ns = 1;
    auto emit = [&](unsigned int R) {
        if (Ns > 1) {
            const unsigned int tstride = (n / R) / Ns;
            for (unsigned int m = 1; m < R; m++)
                for (unsigned int k = 0; k < Ns; k++) t.push_back(flat[(size_t)k * m * tstride]);
        }
    };
    while (left >= 4) { emit(16); Ns <<= 4; left -= 4; }
    if (left == 3) emit(8);
    else if (left == 2) emit(4);
    else if (left == 1) emit(2);
    t.resize(n, make_float2(0.f, 0.f));
    return t;
}

// transforms per CTA for n-point transforms and the shared memory they need.  2048 points per CTA keeps four
// CTAs per SM (what the analyzers, whose fill phase is load-latency bound, want); 4096 gives every one of the 256
// threads a butterfly in a radix-16 pass (what the periodogram's long FFT-only loop wants)
inline unsigned int per_cta_for(unsigned int n, unsigned long long batch, unsigned int target_points = 2048)
{
    unsigned int per_cta = 1;
    while (per_cta * n < target_points && per_cta < batch) per_cta <<= 1;
    return per_cta;
}
inline size_t smem_for(unsigned int n, unsigned int per_cta, bool staged_table = true)
{
    return ((size_t)2 * padded_len(n, per_cta) + (staged_table ? n : 0)) * sizeof(float2);
}
inline bool pow2_in_smem(unsigned int n) { return lc::is_pow2(n) && n >= 2 && (size_t)n * 16 <= 200 * 1024; }

}  // namespace lcfft
