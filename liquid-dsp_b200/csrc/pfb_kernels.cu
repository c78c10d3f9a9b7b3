// pfb_kernels.cu -- generic polyphase multiply-accumulate kernels (any M, any tap count).
//
// One thread per output element, channel index fastest so that every load and store of a warp is
// a contiguous run; samples are re-read through L1/L2 (each is used by 2m..4m outputs).  These are
// the shape-agnostic kernels; the shapes the benchmarks run use the fused kernels in
// pfbch2_fused.cu, which keep the sliding windows in registers.
#include "pfb_kernels.cuh"
#include "fft_device.cuh"
#include "fused_common.cuh"

namespace {

constexpr int kThreads = 256;

// acc += tap * sample for real (crcf) and complex (cccf: plain product, no conjugate) taps
__device__ __forceinline__ void tap_mac(float c, float2 s, float &re, float &im)
{
    re = fmaf(c, s.x, re);
    im = fmaf(c, s.y, im);
}
__device__ __forceinline__ void tap_mac(float2 c, float2 s, float &re, float &im)
{
    re = fmaf(c.x, s.x, re);
    re = fmaf(-c.y, s.y, re);
    im = fmaf(c.x, s.y, im);
    im = fmaf(c.y, s.x, im);
}

// transform input w of frame t (relative to x[0]) of the firpfbch2 analyzer       firpfbch2.c:253-271
__device__ __forceinline__ float2 pfbch2_mac_elem(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                                                  const float *__restrict__ taps, unsigned int M, unsigned int L,
                                                  int parity0, long long t, unsigned int w)
{
    const unsigned int M2 = M >> 1;
    const int par = (parity0 + (int)(t & 1)) & 1;
    const bool lower = w < M2;                      // lower half is pushed on even frames
    const long long tu = (par == (lower ? 0 : 1)) ? t : t - 1;     // frame of the newest sample
    const long long off = lower ? (long long)(M2 - 1 - w) : (long long)(M - 1 - w);
    const float *tp = taps + (size_t)par * L * M + w;
    float re = 0.f, im = 0.f;
    for (unsigned int n = 0; n < L; n++) {
        long long idx = (tu - 2 * (long long)n) * (long long)M2 + off;
        float2 s = lcd::ld_stream(hist, x, idx, H);
        float c = __ldg(tp + (size_t)n * M);
        re = fmaf(c, s.x, re);
        im = fmaf(c, s.y, im);
    }
    return make_float2(re, im);
}

__global__ void __launch_bounds__(kThreads)
pfbch2_mac_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                  const float *__restrict__ taps, float2 *__restrict__ X, unsigned int M, unsigned int L,
                  unsigned long long n_frames, int parity0, long long t_first)
{
    const unsigned long long total = n_frames * M;
    for (unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
         gid += (unsigned long long)gridDim.x * blockDim.x)
        X[gid] = pfbch2_mac_elem(hist, x, H, taps, M, L, parity0, t_first + (long long)(gid / M), (unsigned int)(gid % M));
}

// last-pass destination of the one-kernel analyzers: global memory, scaled
struct FrameDst {
    float2 *y;
    unsigned int M;
    unsigned long long first, n_frames;
    int post_mode;
    float pa;
    __device__ __forceinline__ void operator()(unsigned int f, unsigned int k, float2 v) const
    {
        if (first + f < n_frames) y[(first + f) * M + k] = lcfft::fft_post(v, post_mode, pa, 1.0f);
    }
};

// firpfbch2 analyzer for power-of-two M in ONE kernel: the polyphase MACs of `per_cta` frames are written
// straight into shared memory (all threads busy with independent dot products: computing them inside the
// first FFT pass instead serialises 16 of them per thread and halved the rate), transformed there (Stockham,
// backward), and the last pass writes global memory divided by M -- the transform input never visits HBM.
// Samples are re-read through L1/L2 (2m uses each).
__global__ void __launch_bounds__(lcfft::kFftThreads)
pfbch2_analyze_pow2_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                           const float *__restrict__ taps, const float2 *__restrict__ tw, float2 *__restrict__ y,
                           unsigned int M, unsigned int log2M, unsigned int L, unsigned int per_cta,
                           unsigned long long n_frames, int parity0, long long t_first)
{
    extern __shared__ float2 smem_f2[];
    const lcfft::Smem sm = lcfft::carve(smem_f2, M, per_cta);
    lcfft::stage_twiddles(sm, tw, M);
    const unsigned long long first = (unsigned long long)blockIdx.x * per_cta;
    {
        // fill: thread = (window, frame lane); everything that depends on the window alone is hoisted out of
        // the frame loop and the taps walk by pointer -- the same fmaf sequence as pfbch2_mac_elem, bit for bit
        const unsigned int M2 = M >> 1;
        const unsigned int nw = M < blockDim.x ? M : blockDim.x;       // windows covered per pass (power of two)
        const unsigned int fl = blockDim.x / nw;                        // frame lanes
        const unsigned int w0 = threadIdx.x % nw, f0 = threadIdx.x / nw;
        for (unsigned int w = w0; w < M; w += nw) {
            const bool lower = w < M2;                                  // lower half is pushed on even frames
            const long long off = lower ? (long long)(M2 - 1 - w) : (long long)(M - 1 - w);
            for (unsigned int fi = f0; fi < per_cta; fi += fl) {
                const unsigned long long f = first + fi;
                float re = 0.f, im = 0.f;
                if (f < n_frames) {
                    const long long t = t_first + (long long)f;
                    const int par = (parity0 + (int)(t & 1)) & 1;
                    const long long tu = (par == (lower ? 0 : 1)) ? t : t - 1;      // frame of the newest sample
                    const long long idx0 = tu * (long long)M2 + off;              // tap n reads idx0 - n M
                    const float *tp = taps + (size_t)par * L * M + w;
                    if (idx0 - (long long)(L - 1) * M >= 0) {
                        const float2 *sp = x + idx0;
                        for (unsigned int n = 0; n < L; n++, sp -= M, tp += M) {
                            const float2 sv = __ldg(sp);
                            const float c = __ldg(tp);
                            re = fmaf(c, sv.x, re);
                            im = fmaf(c, sv.y, im);
                        }
                    } else {
                        for (unsigned int n = 0; n < L; n++, tp += M) {
                            const float2 sv = lcd::ld_stream(hist, x, idx0 - (long long)n * M, H);
                            const float c = __ldg(tp);
                            re = fmaf(c, sv.x, re);
                            im = fmaf(c, sv.y, im);
                        }
                    }
                }
                sm.buf0[lcfft::pidx(fi * M + w)] = make_float2(re, im);
            }
        }
    }
    __syncthreads();
    const FrameDst dst{y, M, first, n_frames, 1, (float)M};
    lcfft::stockham_run<+1>(sm, M, log2M, per_cta, lcfft::SmemSrc{sm.buf0, log2M}, dst);
}

__global__ void __launch_bounds__(kThreads)
pfbch2_synth_fir_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ v, int hist_rows,
                        const float *__restrict__ taps, float2 *__restrict__ y, unsigned int M,
                        unsigned int n_taps, unsigned long long n_frames, int parity0)
{
    const unsigned int M2 = M >> 1;
    const unsigned long long total = n_frames * M2;
    for (unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
         gid += (unsigned long long)gridDim.x * blockDim.x) {
        const long long t = (long long)(gid / M2);
        const unsigned int i = (unsigned int)(gid % M2);
        const int par = (parity0 + (int)(t & 1)) & 1;
        const unsigned int j = i + (par ? M2 : 0);
        // even and odd taps are the reference's two dot products (dp[i], dp[i+M/2]); summed apart
        float re0 = 0.f, im0 = 0.f, re1 = 0.f, im1 = 0.f;
        for (unsigned int q = 0; q < n_taps; q++) {
            long long r = t - (long long)q;
            float2 s = (r >= 0) ? __ldg(v + (size_t)r * M + j) : __ldg(hist + (size_t)(hist_rows + r) * M + j);
            float c = __ldg(taps + (size_t)q * M2 + i);
            if (q & 1) { re1 = fmaf(c, s.x, re1); im1 = fmaf(c, s.y, im1); }
            else       { re0 = fmaf(c, s.x, re0); im0 = fmaf(c, s.y, im0); }
        }
        y[gid] = make_float2(re0 + re1, im0 + im1);
    }
}

// firpfbch2 synthesis FIR with the rows of v kept in registers (power-of-two-ish shapes: M/2 a multiple of 128,
// NT = 4m taps per output in {8, 16, 20, 28}: 4 NT registers of state + NT taps).  Thread i owns output column i: the last NT values of
// v[.][i] and of v[.][i + M/2] live in two circular register buffers and the NT taps h[i + q M/2] beside them;
// per frame it loads two new values (coalesced rows), does NT packed FFMA2 and stores one output.  A CTA walks
// a chunk of frames that starts on a multiple of NT, so that -- one period of NT frames fully unrolled -- every
// buffer slot index is a compile-time constant; the chunk's first NT-1 rows of history come from v itself (or
// from the object's state in front of frame 0).  Each row of v is read from HBM once instead of 4m times
// through L2.  Same summation order as pfbch2_synth_fir_kernel (even and odd taps apart), so the results are
// bit-identical to it.
template <int NT, int PAR0>
__global__ void __launch_bounds__(128, NT > 16 ? 2 : 3)
pfbch2_synth_fir_reg_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ v, int hist_rows,
                            const float *__restrict__ taps, float2 *__restrict__ y, unsigned int M,
                            unsigned long long n_frames, unsigned int frames_per_chunk)
{
    using namespace lcf;
    const unsigned int M2 = M >> 1;
    const unsigned int i = blockIdx.x * 128 + threadIdx.x;            // output column, < M2 by construction
    const long long t0 = (long long)blockIdx.y * frames_per_chunk;    // multiple of NT (hence even)
    long long t_end = t0 + frames_per_chunk;
    if (t_end > (long long)n_frames) t_end = (long long)n_frames;
    float h[NT];
#pragma unroll
    for (int q = 0; q < NT; q++) h[q] = __ldg(taps + (size_t)q * M2 + i);
    float2 lo[NT], hi[NT];                                             // row r sits in slot r mod NT
    auto row = [&](long long r) -> const float2 * {
        return r >= 0 ? v + (size_t)r * M : hist + (size_t)(hist_rows + r) * M;
    };
    // thread-private prefetch ring in shared memory: slot tt holds the (lo, hi) pair of frame tp + tt for this
    // thread's column, fetched kAhead frames early by cp.async; only the owning thread touches its slots, so
    // cp.async.wait_group is the only synchronisation
    extern __shared__ float2 s_ring[];                                  // [NT][2][128]
    constexpr int kAhead = NT - 2;
    const unsigned int ring = (unsigned int)__cvta_generic_to_shared(s_ring) + threadIdx.x * 8u;
    auto fetch = [&](long long t, int slot) {
        if (t < t_end) {
            const float2 *p = v + (size_t)t * M + i;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ring + (unsigned int)slot * 2048u), "l"(p) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ring + (unsigned int)slot * 2048u + 1024u), "l"(p + M2) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < kAhead; s++) fetch(t0 + s, s);
#pragma unroll
    for (int s = 1; s < NT; s++) {                                     // rows t0 - NT + s, s = 1 .. NT-1
        const float2 *p = row(t0 - NT + s);
        lo[s] = __ldg(p + i);
        hi[s] = __ldg(p + i + M2);
    }
    lo[0] = hi[0] = make_float2(0.f, 0.f);
    for (long long tp = t0; tp < t_end; tp += NT) {
#pragma unroll
        for (int tt = 0; tt < NT; tt++) {
            const long long t = tp + tt;
            if (t >= t_end) break;                                     // uniform: tail of the last period
            fetch(t + kAhead, (tt + kAhead) % NT);
            asm volatile("cp.async.wait_group %0;" ::"n"(kAhead) : "memory");
            lo[tt] = s_ring[tt * 256 + threadIdx.x];
            hi[tt] = s_ring[tt * 256 + 128 + threadIdx.x];
            const bool odd = ((PAR0 + tt) & 1) != 0;                   // t0 and NT are even
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < NT; q++) {
                const float2 sv = odd ? hi[(tt - q + NT) % NT] : lo[(tt - q + NT) % NT];
                if (q & 1) a1 = __ffma2_rn(bc(h[q]), sv, a1);
                else       a0 = __ffma2_rn(bc(h[q]), sv, a0);
            }
            y[(size_t)t * M2 + i] = __fadd2_rn(a0, a1);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int NT>
bool launch_synth_fir_reg(const float2 *d_hist_rows, const float2 *d_v, int hist_rows, const float *d_taps, float2 *d_y,
                          unsigned int M, unsigned long long n_frames, int parity0, cudaStream_t st)
{
    const unsigned int fpc = NT * ((256 + NT - 1) / NT);               // frames per chunk: a multiple of NT near 256
    const unsigned long long chunks = (n_frames + fpc - 1) / fpc;
    if (chunks > 65535) return false;                                  // grid.y limit: the caller falls back
    dim3 grid((M / 2) / 128, (unsigned int)chunks);
    constexpr size_t smem = (size_t)NT * 2048;                         // the prefetch ring
    static std::atomic<unsigned long long> configured{0};
    if (smem > 48 * 1024 && lc::first_time_on_device(configured)) {
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_synth_fir_reg_kernel<NT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_synth_fir_reg_kernel<NT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (parity0 & 1)
        pfbch2_synth_fir_reg_kernel<NT, 1><<<grid, 128, smem, st>>>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames, fpc);
    else
        pfbch2_synth_fir_reg_kernel<NT, 0><<<grid, 128, smem, st>>>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames, fpc);
    LC_LAUNCH_CHECK();
    return true;
}

// transform input u of run s of the firpfbch analyzer (M pushes per run, alignment index k)  firpfbch.c:369-409
template <typename TapT>
__device__ __forceinline__ float2 pfbch_mac_elem(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                                                 const TapT *__restrict__ taps, unsigned int M, unsigned int p,
                                                 unsigned int phase, unsigned int first, unsigned int k,
                                                 long long s, unsigned int u)
{
    const unsigned int i = M - 1 - u;                       // branch
    const unsigned int win = (i + k) % M;                   // window it reads (firpfbch.c:395)
    // pushes are numbered from `phase` (push c goes to window M-1-(c mod M), firpfbch.c:369-377);
    // x[0] is push number `phase`; run s happens after `first` + s*M samples of x, i.e. after
    // pushes up to C-1
    const long long C = (long long)phase + (long long)first + s * (long long)M;
    const long long rr = (long long)(M - 1 - win);
    long long d = (C - 1 - rr) % (long long)M;
    if (d < 0) d += M;
    const long long last = C - 1 - d - (long long)phase;    // index into x of the newest sample
    const TapT *tp = taps + i;
    float re = 0.f, im = 0.f;
    for (unsigned int n = 0; n < p; n++) {
        float2 smp = lcd::ld_stream(hist, x, last - (long long)n * M, H);
        tap_mac(__ldg(tp + (size_t)n * M), smp, re, im);
    }
    return make_float2(re, im);
}

template <typename TapT>
__global__ void __launch_bounds__(kThreads)
pfbch_mac_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                 const TapT *__restrict__ taps, float2 *__restrict__ X, unsigned int M, unsigned int p,
                 unsigned long long n_frames, unsigned int phase, unsigned int first, unsigned int k)
{
    const unsigned long long total = n_frames * M;
    for (unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
         gid += (unsigned long long)gridDim.x * blockDim.x)
        X[gid] = pfbch_mac_elem<TapT>(hist, x, H, taps, M, p, phase, first, k, (long long)(gid / M), (unsigned int)(gid % M));
}

// The same kernel with the window of a thread kept in registers while it walks its frames (LW = 2m taps, a
// compile-time constant): a window takes a new sample every other frame, so a lane that owns a contiguous run of
// frames loads LW + (run / 2) samples instead of LW per frame, and its two tap sets (by parity) once.  Same fmaf
// sequence per output as above: identical bits.
template <int LW>
__global__ void __launch_bounds__(lcfft::kFftThreads)
pfbch2_analyze_pow2_win_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                               const float *__restrict__ taps, const float2 *__restrict__ tw, float2 *__restrict__ y,
                               unsigned int M, unsigned int log2M, unsigned int per_cta,
                               unsigned long long n_frames, int parity0, long long t_first)
{
    extern __shared__ float2 smem_f2[];
    const lcfft::Smem sm = lcfft::carve(smem_f2, M, per_cta);
    lcfft::stage_twiddles(sm, tw, M);
    const unsigned long long first = (unsigned long long)blockIdx.x * per_cta;
    {
        const unsigned int M2 = M >> 1;
        const unsigned int nw = M < blockDim.x ? M : blockDim.x;       // windows covered per pass (power of two)
        const unsigned int fl = blockDim.x / nw;                        // frame lanes
        const unsigned int w0 = threadIdx.x % nw, f0 = threadIdx.x / nw;
        const unsigned int run = (per_cta + fl - 1) / fl;               // contiguous frames per lane
        const unsigned int fa = f0 * run;
        const unsigned int fb = (fa + run < per_cta) ? fa + run : per_cta;
        for (unsigned int w = w0; w < M; w += nw) {
            const bool lower = w < M2;
            const long long off = lower ? (long long)(M2 - 1 - w) : (long long)(M - 1 - w);
            float2 win[LW];                                             // win[n] = sample (tu - 2n) M2 + off
            float c0[LW], c1[LW];                                       // both tap sets of the window (by parity)
#pragma unroll
            for (int n = 0; n < LW; n++) {
                c0[n] = __ldg(taps + w + (size_t)n * M);
                c1[n] = __ldg(taps + (size_t)LW * M + w + (size_t)n * M);
            }
            long long cur_tu = 0;
            bool have = false;
            for (unsigned int fi = fa; fi < fb; fi++) {
                const unsigned long long f = first + fi;
                float re = 0.f, im = 0.f;
                if (f < n_frames) {
                    const long long t = t_first + (long long)f;
                    const int par = (parity0 + (int)(t & 1)) & 1;
                    const long long tu = (par == (lower ? 0 : 1)) ? t : t - 1;
                    if (!have || tu != cur_tu) {
                        if (have && tu == cur_tu + 2) {
#pragma unroll
                            for (int n = LW - 1; n > 0; n--) win[n] = win[n - 1];
                            win[0] = lcd::ld_stream(hist, x, tu * (long long)M2 + off, H);
                        } else {
#pragma unroll
                            for (int n = 0; n < LW; n++)
                                win[n] = lcd::ld_stream(hist, x, (tu - 2 * (long long)n) * (long long)M2 + off, H);
                        }
                        cur_tu = tu;
                        have = true;
                    }
                    if (par) {
#pragma unroll
                        for (int n = 0; n < LW; n++) {
                            re = fmaf(c1[n], win[n].x, re);
                            im = fmaf(c1[n], win[n].y, im);
                        }
                    } else {
#pragma unroll
                        for (int n = 0; n < LW; n++) {
                            re = fmaf(c0[n], win[n].x, re);
                            im = fmaf(c0[n], win[n].y, im);
                        }
                    }
                }
                sm.buf0[lcfft::pidx(fi * M + w)] = make_float2(re, im);
            }
        }
    }
    __syncthreads();
    const FrameDst dst{y, M, first, n_frames, 1, (float)M};
    lcfft::stockham_run<+1>(sm, M, log2M, per_cta, lcfft::SmemSrc{sm.buf0, log2M}, dst);
}

// firpfbch analyzer for power-of-two M in ONE kernel (see pfbch2_analyze_pow2_kernel): MAC into shared memory,
// forward Stockham FFT there, unscaled store by the last pass
template <typename TapT>
__global__ void __launch_bounds__(lcfft::kFftThreads)
pfbch_analyze_pow2_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                          const TapT *__restrict__ taps, const float2 *__restrict__ tw, float2 *__restrict__ y,
                          unsigned int M, unsigned int log2M, unsigned int p, unsigned int per_cta,
                          unsigned long long n_frames, unsigned int phase, unsigned int first_in, unsigned int k)
{
    extern __shared__ float2 smem_f2[];
    const lcfft::Smem sm = lcfft::carve(smem_f2, M, per_cta);
    lcfft::stage_twiddles(sm, tw, M);
    const unsigned long long first = (unsigned long long)blockIdx.x * per_cta;
    {
        // fill: thread = (transform input u, frame lane).  The window branch M-1-u reads and the offset d of its
        // newest sample do not depend on the run index (C = phase + first + s M is congruent mod M for all s),
        // so they are hoisted; same tap_mac sequence as pfbch_mac_elem, bit for bit
        const unsigned int nw = M < blockDim.x ? M : blockDim.x;
        const unsigned int fl = blockDim.x / nw;
        const unsigned int u0 = threadIdx.x % nw, f0 = threadIdx.x / nw;
        for (unsigned int u = u0; u < M; u += nw) {
            const unsigned int i = M - 1 - u;                           // branch
            const unsigned int win = (i + k) % M;                       // window it reads (firpfbch.c:395)
            const long long rr = (long long)(M - 1 - win);
            long long d = ((long long)phase + (long long)first_in - 1 - rr) % (long long)M;
            if (d < 0) d += M;
            const long long last0 = (long long)first_in - 1 - d;        // newest sample of run 0 (index into x)
            const TapT *tp0 = taps + i;
            for (unsigned int fi = f0; fi < per_cta; fi += fl) {
                const unsigned long long f = first + fi;
                float re = 0.f, im = 0.f;
                if (f < n_frames) {
                    const long long last = last0 + (long long)f * M;
                    const TapT *tp = tp0;
                    if (last - (long long)(p - 1) * M >= 0) {
                        const float2 *sp = x + last;
                        for (unsigned int n = 0; n < p; n++, sp -= M, tp += M) tap_mac(__ldg(tp), __ldg(sp), re, im);
                    } else {
                        for (unsigned int n = 0; n < p; n++, tp += M)
                            tap_mac(__ldg(tp), lcd::ld_stream(hist, x, last - (long long)n * M, H), re, im);
                    }
                }
                sm.buf0[lcfft::pidx(fi * M + u)] = make_float2(re, im);
            }
        }
    }
    __syncthreads();
    const FrameDst dst{y, M, first, n_frames, 0, 1.0f};
    lcfft::stockham_run<-1>(sm, M, log2M, per_cta, lcfft::SmemSrc{sm.buf0, log2M}, dst);
}

template <typename TapT>
__global__ void __launch_bounds__(kThreads)
pfbch_synth_fir_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ v, int hist_rows,
                       const TapT *__restrict__ taps, float2 *__restrict__ y, unsigned int M, unsigned int p,
                       unsigned long long n_frames)
{
    const unsigned long long total = n_frames * M;
    for (unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
         gid += (unsigned long long)gridDim.x * blockDim.x) {
        const long long s = (long long)(gid / M);
        const unsigned int i = (unsigned int)(gid % M);
        float re = 0.f, im = 0.f;
        for (unsigned int n = 0; n < p; n++) {
            long long r = s - (long long)n;
            float2 smp = (r >= 0) ? __ldg(v + (size_t)r * M + i) : __ldg(hist + (size_t)(hist_rows + r) * M + i);
            tap_mac(__ldg(taps + (size_t)n * M + i), smp, re, im);
        }
        y[gid] = make_float2(re, im);
    }
}

// ---- column-walking MAC kernels of the two-kernel fallback (shapes without a dedicated kernel: M = 2048, the critically
// sampled M = 4096, filter lengths above the instantiated ones).  The element-per-thread MAC kernels above re-read every
// sample 2m times with 64-bit index arithmetic per tap; here a thread owns ONE window (transform input), keeps its 2m
// samples and taps in registers and walks a chunk of frames: one new sample per push, coalesced across the windows of a
// warp, eight pushes loaded ahead of their use.  Same fmaf sequence per output (newest sample first): identical bits.
constexpr int kColThreads = 128;
constexpr int kColGroup = 8;                               // pushes per group

// firpfbch2 analyzer: window w takes a sample every other frame and is read every frame with the tap set of the
// frame's parity.  Frames are counted from b = the chunk's first frame rounded down to the window's push parity.
template <int L>
__global__ void __launch_bounds__(kColThreads)
pfbch2_mac_cols_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H, const float *__restrict__ taps,
                       float2 *__restrict__ X, unsigned int M, long long n_frames, int parity0, long long t_first, int fpc)
{
    const unsigned int M2 = M >> 1;
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;      // blockDim.x divides M/2: a CTA's windows share a half
    const bool lower = w < M2;
    const long long off = lower ? (long long)(M2 - 1 - w) : (long long)(M - 1 - w);
    const long long c0 = (long long)blockIdx.y * fpc;                   // first frame of the chunk (relative to the block)
    long long c1 = c0 + fpc;
    if (c1 > n_frames) c1 = n_frames;
    // absolute frame tt = t_first + t; the window is pushed on frames whose parity (parity0 + tt) & 1 equals (lower ? 0 : 1)
    const long long tt0 = t_first + c0;
    const int par0 = (parity0 + (int)(tt0 & 1)) & 1;
    const long long b = (par0 == (lower ? 0 : 1)) ? tt0 : tt0 - 1;      // a push frame at or just before the chunk
    const int pb = (parity0 + (int)(b & 1)) & 1;                         // tap set of frame b (= lower ? 0 : 1)
    float h0[L], h1[L];                                                  // tap sets of frames b + 2r and b + 2r + 1
#pragma unroll
    for (int n = 0; n < L; n++) {
        h0[n] = __ldg(taps + (size_t)pb * L * M + (size_t)n * M + w);
        h1[n] = __ldg(taps + (size_t)(1 - pb) * L * M + (size_t)n * M + w);
    }
    float2 A[L - 1 + kColGroup];                                         // A[L - 1 + r] = sample of push frame b + 2 (j0 + r)
#pragma unroll
    for (int n = 1; n < L; n++) A[L - 1 - n] = lcd::ld_stream(hist, x, (b - 2LL * n) * (long long)M2 + off, H);
    for (long long i0 = 0; b + i0 < t_first + c1; i0 += 2 * kColGroup) {
#pragma unroll
        for (int r = 0; r < kColGroup; r++) {
            const long long tu = b + i0 + 2 * r;
            A[L - 1 + r] = tu < t_first + c1 ? lcd::ld_stream(hist, x, tu * (long long)M2 + off, H) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < kColGroup; r++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const long long tt = b + i0 + 2 * r + e;
                if (tt >= tt0 && tt < t_first + c1) {
                    float re = 0.f, im = 0.f;
#pragma unroll
                    for (int n = 0; n < L; n++) {
                        const float c = e ? h1[n] : h0[n];
                        re = fmaf(c, A[L - 1 + r - n].x, re);
                        im = fmaf(c, A[L - 1 + r - n].y, im);
                    }
                    X[(tt - t_first) * (long long)M + w] = make_float2(re, im);
                }
            }
        }
#pragma unroll
        for (int n = 0; n < L - 1; n++) A[n] = A[n + kColGroup];
    }
}

// firpfbch analyzer, block-aligned (phase 0, k = 0): transform input u of frame t reads x[(t - n) M + u] with tap
// taps[n M + (M - 1 - u)]
template <int P>
__global__ void __launch_bounds__(kColThreads)
pfbch_mac_cols_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H, const float *__restrict__ taps,
                      float2 *__restrict__ X, unsigned int M, long long n_frames, long long f_first, int fpc)
{
    const unsigned int u = blockIdx.x * blockDim.x + threadIdx.x;
    const long long c0 = (long long)blockIdx.y * fpc;
    long long c1 = c0 + fpc;
    if (c1 > n_frames) c1 = n_frames;
    float h[P];
#pragma unroll
    for (int n = 0; n < P; n++) h[n] = __ldg(taps + (size_t)n * M + (M - 1 - u));
    float2 A[P - 1 + kColGroup];
#pragma unroll
    for (int n = 1; n < P; n++) A[P - 1 - n] = lcd::ld_stream(hist, x, (f_first + c0 - n) * (long long)M + u, H);
    for (long long t0 = c0; t0 < c1; t0 += kColGroup) {
#pragma unroll
        for (int r = 0; r < kColGroup; r++)
            A[P - 1 + r] = t0 + r < c1 ? lcd::ld_stream(hist, x, (f_first + t0 + r) * (long long)M + u, H) : make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kColGroup; r++) {
            if (t0 + r < c1) {
                float re = 0.f, im = 0.f;
#pragma unroll
                for (int n = 0; n < P; n++) {
                    re = fmaf(h[n], A[P - 1 + r - n].x, re);
                    im = fmaf(h[n], A[P - 1 + r - n].y, im);
                }
                X[(t0 + r) * (long long)M + u] = make_float2(re, im);
            }
        }
#pragma unroll
        for (int n = 0; n < P - 1; n++) A[n] = A[n + kColGroup];
    }
}

inline unsigned int grid_for(unsigned long long total)
{
    unsigned long long b = (total + kThreads - 1) / kThreads;
    const unsigned long long cap = 148ull * 8 * 16;     // grid-stride beyond this
    return (unsigned int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace

namespace lc {
namespace pfb {

void launch_pfbch2_mac(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, float2 *d_X,
                       unsigned int M, unsigned int L, unsigned long long n_frames, int parity0, cudaStream_t st,
                       long long t_first)
{
    const unsigned int col_threads = (M / 2) % kColThreads == 0 ? kColThreads : (M / 2) % 64 == 0 ? 64 : 32;
    if ((M / 2) % 32 == 0 && n_frames >= 64 && !getenv("LIQUID_CUDA_MAC_PLAIN")) {
        const int fpc = 256;
        const unsigned long long chunks = (n_frames + fpc - 1) / fpc;
        if (chunks <= 65535) {
            dim3 grid(M / col_threads, (unsigned int)chunks);
            bool done = true;
            switch (L) {
#define LC_COLS(LL)                                                                                                        \
    case LL:                                                                                                               \
        pfbch2_mac_cols_kernel<LL><<<grid, col_threads, 0, st>>>(d_hist, d_x, H, d_taps, d_X, M, (long long)n_frames, parity0, \
                                                                 t_first, fpc);                                            \
        break;
                LC_COLS(2) LC_COLS(4) LC_COLS(6) LC_COLS(8) LC_COLS(10) LC_COLS(12) LC_COLS(14) LC_COLS(16) LC_COLS(18) LC_COLS(20)
#undef LC_COLS
            default: done = false;
            }
            if (done) {
                LC_LAUNCH_CHECK();
                return;
            }
        }
    }
    pfbch2_mac_kernel<<<grid_for(n_frames * M), kThreads, 0, st>>>(d_hist, d_x, H, d_taps, d_X, M, L, n_frames, parity0, t_first);
    LC_LAUNCH_CHECK();
}

void launch_pfbch2_synth_fir(const float2 *d_hist_rows, const float2 *d_v, int hist_rows, const float *d_taps,
                             float2 *d_y, unsigned int M, unsigned int n_taps, unsigned long long n_frames,
                             int parity0, cudaStream_t st)
{
    // register-window kernel for the shapes it is instantiated for (needs >= n_taps - 1 rows of history)
    if ((M / 2) % 128 == 0 && hist_rows >= (int)n_taps - 1 && n_frames >= 64 && !getenv("LIQUID_CUDA_NO_FUSED")) {
        bool done = false;
        switch (n_taps) {
        case 8: done = launch_synth_fir_reg<8>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames, parity0, st); break;
        case 12: done = launch_synth_fir_reg<12>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames, parity0, st); break;
        case 16: done = launch_synth_fir_reg<16>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames, parity0, st); break;
        case 20: done = launch_synth_fir_reg<20>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames, parity0, st); break;
        case 24: done = launch_synth_fir_reg<24>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames, parity0, st); break;
        case 28: done = launch_synth_fir_reg<28>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames, parity0, st); break;
        case 32: done = launch_synth_fir_reg<32>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames, parity0, st); break;
        case 36: done = launch_synth_fir_reg<36>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames, parity0, st); break;
        default: break;
        }
        if (done) return;
    }
    pfbch2_synth_fir_kernel<<<grid_for(n_frames * (M / 2)), kThreads, 0, st>>>(d_hist_rows, d_v, hist_rows, d_taps, d_y,
                                                                              M, n_taps, n_frames, parity0);
    LC_LAUNCH_CHECK();
}

void launch_pfbch_mac(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, float2 *d_X,
                      unsigned int M, unsigned int p, unsigned long long n_frames, unsigned int phase,
                      unsigned int first, unsigned int k, cudaStream_t st)
{
    const unsigned int col_threads = M % kColThreads == 0 ? kColThreads : 64;
    if (phase == 0 && k == 0 && first >= M && first % M == 0 && M % 64 == 0 && n_frames >= 64 && !getenv("LIQUID_CUDA_MAC_PLAIN")) {
        const int fpc = 256;
        const unsigned long long chunks = (n_frames + fpc - 1) / fpc;
        if (chunks <= 65535) {
            dim3 grid(M / col_threads, (unsigned int)chunks);
            const long long f_first = (long long)(first / M) - 1;         // frame of x the first run reads last
            bool done = true;
            switch (p) {
#define LC_COLS(PP)                                                                                                       \
    case PP:                                                                                                              \
        pfbch_mac_cols_kernel<PP><<<grid, col_threads, 0, st>>>(d_hist, d_x, H, d_taps, d_X, M, (long long)n_frames, f_first, fpc); \
        break;
                LC_COLS(2) LC_COLS(4) LC_COLS(6) LC_COLS(8) LC_COLS(10) LC_COLS(12) LC_COLS(14) LC_COLS(16) LC_COLS(18) LC_COLS(20)
#undef LC_COLS
            default: done = false;
            }
            if (done) {
                LC_LAUNCH_CHECK();
                return;
            }
        }
    }
    pfbch_mac_kernel<float><<<grid_for(n_frames * M), kThreads, 0, st>>>(d_hist, d_x, H, d_taps, d_X, M, p, n_frames, phase, first, k);
    LC_LAUNCH_CHECK();
}

void launch_pfbch_mac_cc(const float2 *d_hist, const float2 *d_x, int H, const float2 *d_taps, float2 *d_X,
                         unsigned int M, unsigned int p, unsigned long long n_frames, unsigned int phase,
                         unsigned int first, unsigned int k, cudaStream_t st)
{
    pfbch_mac_kernel<float2><<<grid_for(n_frames * M), kThreads, 0, st>>>(d_hist, d_x, H, d_taps, d_X, M, p, n_frames, phase, first, k);
    LC_LAUNCH_CHECK();
}

namespace {
// firpfbch synthesis FIR with a column's window and taps in registers (P taps, compile time): thread i owns output
// column i of a chunk of frames, loads each row of v once (8 rows in flight) and slides its window by one per frame,
// instead of P loads per output through L2.  Same tap_mac sequence as pfbch_synth_fir_kernel: identical bits.
constexpr unsigned int kSynthChunk = 256;                  // frames per CTA
template <typename TapT, int P>
__global__ void __launch_bounds__(128)
pfbch_synth_fir_win_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ v, int hist_rows,
                           const TapT *__restrict__ taps, float2 *__restrict__ y, unsigned int M, unsigned long long n_frames)
{
    const unsigned int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= M) return;
    const long long s0 = (long long)blockIdx.y * kSynthChunk;
    long long s_end = s0 + kSynthChunk;
    if (s_end > (long long)n_frames) s_end = (long long)n_frames;
    TapT c[P];
#pragma unroll
    for (int n = 0; n < P; n++) c[n] = __ldg(taps + (size_t)n * M + i);
    float2 win[P];                                         // win[n] = row s - n
#pragma unroll
    for (int n = 1; n < P; n++) {                          // rows s0 - 1 ... s0 - (P-1), shifted in by the first frame
        const long long r = s0 - n;
        win[n - 1] = (r >= 0) ? __ldg(v + (size_t)r * M + i) : __ldg(hist + (size_t)(hist_rows + r) * M + i);
    }
    constexpr int U = 8;
    for (long long s = s0; s < s_end; s += U) {
        float2 rows[U];
#pragma unroll
        for (int u = 0; u < U; u++) rows[u] = (s + u < s_end) ? __ldg(v + (size_t)(s + u) * M + i) : make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (s + u < s_end) {
#pragma unroll
                for (int n = P - 1; n > 0; n--) win[n] = win[n - 1];
                win[0] = rows[u];
                float re = 0.f, im = 0.f;
#pragma unroll
                for (int n = 0; n < P; n++) tap_mac(c[n], win[n], re, im);
                y[(size_t)(s + u) * M + i] = make_float2(re, im);
            }
        }
    }
}

template <typename TapT>
bool launch_synth_fir_win(const float2 *d_hist_rows, const float2 *d_v, int hist_rows, const TapT *d_taps, float2 *d_y,
                          unsigned int M, unsigned int p, unsigned long long n_frames, cudaStream_t st)
{
    const unsigned long long chunks = (n_frames + kSynthChunk - 1) / kSynthChunk;
    if (n_frames < 2 * kSynthChunk || chunks > 65535 || hist_rows + 1 < (int)p || getenv("LIQUID_CUDA_SYNTH_NO_WINDOW")) return false;
    const dim3 grid((M + 127) / 128, (unsigned int)chunks);
    switch (p) {
#define LC_WIN(PP)                                                                                                   \
    case PP:                                                                                                         \
        pfbch_synth_fir_win_kernel<TapT, PP><<<grid, 128, 0, st>>>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, n_frames); \
        break;
        LC_WIN(2) LC_WIN(3) LC_WIN(4) LC_WIN(5) LC_WIN(6) LC_WIN(7) LC_WIN(8) LC_WIN(10) LC_WIN(12) LC_WIN(14) LC_WIN(16) LC_WIN(18) LC_WIN(20)
#undef LC_WIN
    default: return false;
    }
    LC_LAUNCH_CHECK();
    return true;
}
}  // namespace

void launch_pfbch_synth_fir(const float2 *d_hist_rows, const float2 *d_v, int hist_rows, const float *d_taps,
                            float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames, cudaStream_t st)
{
    if (launch_synth_fir_win<float>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, p, n_frames, st)) return;
    pfbch_synth_fir_kernel<float><<<grid_for(n_frames * M), kThreads, 0, st>>>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, p, n_frames);
    LC_LAUNCH_CHECK();
}

void launch_pfbch_synth_fir_cc(const float2 *d_hist_rows, const float2 *d_v, int hist_rows, const float2 *d_taps,
                               float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames, cudaStream_t st)
{
    if (launch_synth_fir_win<float2>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, p, n_frames, st)) return;
    pfbch_synth_fir_kernel<float2><<<grid_for(n_frames * M), kThreads, 0, st>>>(d_hist_rows, d_v, hist_rows, d_taps, d_y, M, p, n_frames);
    LC_LAUNCH_CHECK();
}

namespace {
// firpfbch analyzer with the window and the taps of a thread in registers (P taps, compile time): the window of a
// branch takes one new sample per frame, so a lane that owns a contiguous run of frames loads P + run samples and
// its P taps once, instead of 2 P loads per output.  Same tap_mac sequence as pfbch_analyze_pow2_kernel: identical bits.
template <typename TapT, int P>
__global__ void __launch_bounds__(lcfft::kFftThreads)
pfbch_analyze_pow2_win_kernel(const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                              const TapT *__restrict__ taps, const float2 *__restrict__ tw, float2 *__restrict__ y,
                              unsigned int M, unsigned int log2M, unsigned int per_cta,
                              unsigned long long n_frames, unsigned int phase, unsigned int first_in, unsigned int k)
{
    extern __shared__ float2 smem_f2[];
    const lcfft::Smem sm = lcfft::carve(smem_f2, M, per_cta);
    lcfft::stage_twiddles(sm, tw, M);
    const unsigned long long first = (unsigned long long)blockIdx.x * per_cta;
    {
        const unsigned int nw = M < blockDim.x ? M : blockDim.x;
        const unsigned int fl = blockDim.x / nw;
        const unsigned int u0 = threadIdx.x % nw, f0 = threadIdx.x / nw;
        const unsigned int run = (per_cta + fl - 1) / fl;               // contiguous frames per lane
        const unsigned int fa = f0 * run;
        const unsigned int fb = (fa + run < per_cta) ? fa + run : per_cta;
        for (unsigned int u = u0; u < M; u += nw) {
            const unsigned int i = M - 1 - u;                           // branch
            const unsigned int win_i = (i + k) % M;                     // window it reads (firpfbch.c:395)
            const long long rr = (long long)(M - 1 - win_i);
            long long d = ((long long)phase + (long long)first_in - 1 - rr) % (long long)M;
            if (d < 0) d += M;
            const long long last0 = (long long)first_in - 1 - d;        // newest sample of run 0 (index into x)
            TapT c[P];
#pragma unroll
            for (int n = 0; n < P; n++) c[n] = __ldg(taps + i + (size_t)n * M);
            float2 win[P];                                              // win[n] = sample last - n M
            bool have = false;
            for (unsigned int fi = fa; fi < fb; fi++) {
                const unsigned long long f = first + fi;
                float re = 0.f, im = 0.f;
                if (f < n_frames) {
                    const long long last = last0 + (long long)f * M;
                    if (have) {
#pragma unroll
                        for (int n = P - 1; n > 0; n--) win[n] = win[n - 1];
                        win[0] = lcd::ld_stream(hist, x, last, H);
                    } else {
#pragma unroll
                        for (int n = 0; n < P; n++) win[n] = lcd::ld_stream(hist, x, last - (long long)n * M, H);
                        have = true;
                    }
#pragma unroll
                    for (int n = 0; n < P; n++) tap_mac(c[n], win[n], re, im);
                }
                sm.buf0[lcfft::pidx(fi * M + u)] = make_float2(re, im);
            }
        }
    }
    __syncthreads();
    const FrameDst dst{y, M, first, n_frames, 0, 1.0f};
    lcfft::stockham_run<-1>(sm, M, log2M, per_cta, lcfft::SmemSrc{sm.buf0, log2M}, dst);
}

template <typename Kern>
void ensure_big_smem(Kern kern)
{
    // transforms above 3072 points need more than the default 48 KB of dynamic shared memory
    LC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}
}  // namespace

bool analyze_pow2_supported(unsigned int M) { return lcfft::pow2_in_smem(M); }

void launch_pfbch2_analyze_pow2(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, const float2 *d_tw,
                                float2 *d_y, unsigned int M, unsigned int L, unsigned long long n_frames, int parity0,
                                cudaStream_t st, long long t_first)
{
    static std::atomic<unsigned long long> configured{0};
    if (first_time_on_device(configured)) {
        ensure_big_smem(pfbch2_analyze_pow2_kernel);
#define LC_WIN(LW) ensure_big_smem(pfbch2_analyze_pow2_win_kernel<LW>);
        LC_WIN(4) LC_WIN(6) LC_WIN(8) LC_WIN(10) LC_WIN(12) LC_WIN(14) LC_WIN(16) LC_WIN(18) LC_WIN(20)
#undef LC_WIN
    }
    const unsigned int per_cta = lcfft::per_cta_for(M, n_frames);
    const unsigned long long ctas = (n_frames + per_cta - 1) / per_cta;
    if (ctas >= (1ull << 31)) fatal("firpfbch2_crcf_execute_block", "too many frames for one launch");
    const size_t smem = lcfft::smem_for(M, per_cta);
    // a lane that owns at least 4 consecutive frames keeps its window in registers
    const unsigned int lanes = lcfft::kFftThreads / (M < (unsigned int)lcfft::kFftThreads ? M : (unsigned int)lcfft::kFftThreads);
    const bool sliding = per_cta / lanes >= 4 && !getenv("LIQUID_CUDA_ANALYZE_NO_WINDOW");
    switch (sliding ? L : 0u) {
#define LC_WIN(LW)                                                                                             \
    case LW:                                                                                                   \
        pfbch2_analyze_pow2_win_kernel<LW><<<(unsigned int)ctas, lcfft::kFftThreads, smem, st>>>(              \
            d_hist, d_x, H, d_taps, d_tw, d_y, M, ilog2(M), per_cta, n_frames, parity0, t_first);              \
        break;
        LC_WIN(4) LC_WIN(6) LC_WIN(8) LC_WIN(10) LC_WIN(12) LC_WIN(14) LC_WIN(16) LC_WIN(18) LC_WIN(20)
#undef LC_WIN
    default:
        pfbch2_analyze_pow2_kernel<<<(unsigned int)ctas, lcfft::kFftThreads, smem, st>>>(
            d_hist, d_x, H, d_taps, d_tw, d_y, M, ilog2(M), L, per_cta, n_frames, parity0, t_first);
    }
    LC_LAUNCH_CHECK();
}

template <typename TapT>
static void launch_pfbch_analyze_pow2_t(const float2 *d_hist, const float2 *d_x, int H, const TapT *d_taps,
                                        const float2 *d_tw, float2 *d_y, unsigned int M, unsigned int p,
                                        unsigned long long n_frames, unsigned int phase, unsigned int first,
                                        unsigned int k, cudaStream_t st)
{
    static std::atomic<unsigned long long> configured{0};
#define LC_WIN_ALL LC_WIN(2) LC_WIN(3) LC_WIN(4) LC_WIN(5) LC_WIN(6) LC_WIN(7) LC_WIN(8) LC_WIN(10) LC_WIN(12) LC_WIN(14) LC_WIN(16) LC_WIN(18) LC_WIN(20)
    if (first_time_on_device(configured)) {
        ensure_big_smem(pfbch_analyze_pow2_kernel<TapT>);
#define LC_WIN(PP) ensure_big_smem(pfbch_analyze_pow2_win_kernel<TapT, PP>);
        LC_WIN_ALL
#undef LC_WIN
    }
    const unsigned int per_cta = lcfft::per_cta_for(M, n_frames);
    const unsigned long long ctas = (n_frames + per_cta - 1) / per_cta;
    if (ctas >= (1ull << 31)) fatal("firpfbch_analyzer_execute_block", "too many frames for one launch");
    const size_t smem = lcfft::smem_for(M, per_cta);
    // a lane that owns at least 4 consecutive frames keeps its window and taps in registers
    const unsigned int lanes = lcfft::kFftThreads / (M < (unsigned int)lcfft::kFftThreads ? M : (unsigned int)lcfft::kFftThreads);
    const bool sliding = per_cta / lanes >= 4 && !getenv("LIQUID_CUDA_ANALYZE_NO_WINDOW");
    switch (sliding ? p : 0u) {
#define LC_WIN(PP)                                                                                              \
    case PP:                                                                                                    \
        pfbch_analyze_pow2_win_kernel<TapT, PP><<<(unsigned int)ctas, lcfft::kFftThreads, smem, st>>>(          \
            d_hist, d_x, H, d_taps, d_tw, d_y, M, ilog2(M), per_cta, n_frames, phase, first, k);                \
        break;
        LC_WIN_ALL
#undef LC_WIN
    default:
        pfbch_analyze_pow2_kernel<TapT><<<(unsigned int)ctas, lcfft::kFftThreads, smem, st>>>(
            d_hist, d_x, H, d_taps, d_tw, d_y, M, ilog2(M), p, per_cta, n_frames, phase, first, k);
    }
#undef LC_WIN_ALL
    LC_LAUNCH_CHECK();
}

void launch_pfbch_analyze_pow2(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, const float2 *d_tw,
                               float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                               unsigned int phase, unsigned int first, unsigned int k, cudaStream_t st)
{
    launch_pfbch_analyze_pow2_t<float>(d_hist, d_x, H, d_taps, d_tw, d_y, M, p, n_frames, phase, first, k, st);
}

void launch_pfbch_analyze_pow2_cc(const float2 *d_hist, const float2 *d_x, int H, const float2 *d_taps, const float2 *d_tw,
                                  float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                                  unsigned int phase, unsigned int first, unsigned int k, cudaStream_t st)
{
    launch_pfbch_analyze_pow2_t<float2>(d_hist, d_x, H, d_taps, d_tw, d_y, M, p, n_frames, phase, first, k, st);
}

}  // namespace pfb
}  // namespace lc
