// lc_common.cuh -- shared host/device plumbing of libliquid_cuda (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define LC_VERSION "liquid_cuda 0.1 (sm_100a)"

// complex float as the C ABI sees it: {re, im}; same layout as float2 / C99 float complex
struct lc_cf { float re, im; };
static_assert(sizeof(lc_cf) == 8, "complex float must be 8 bytes");

#define LC_CUDA_CHECK(expr)                                                                     \
    do {                                                                                        \
        cudaError_t lc_err__ = (expr);                                                          \
        if (lc_err__ != cudaSuccess) {                                                          \
            fprintf(stderr, "liquid_cuda: CUDA error %s (%s) at %s:%d: %s\n",                   \
                    cudaGetErrorName(lc_err__), cudaGetErrorString(lc_err__), __FILE__,         \
                    __LINE__, #expr);                                                           \
            abort();                                                                            \
        }                                                                                       \
    } while (0)

namespace lc {

// reference error convention: print "error: <fn>(), <msg>" and exit(1)  (e.g. firpfbch2.c:72-81)
[[noreturn]] inline void fatal(const char *fn, const char *msg)
{
    fprintf(stderr, "error: %s(), %s\n", fn, msg);
    exit(1);
}

// selects the device once (env LIQUID_CUDA_DEVICE or the current one) and aborts loudly when
// there is no usable sm_100 GPU: this library has no CPU path.
void ensure_device();
void choose_device(int dev);       // liquid_cuda_set_device: takes precedence over LIQUID_CUDA_DEVICE, applies to every thread

extern std::atomic<unsigned long long> g_launches;
// frames each CTA (fused M=1024 kernel) or cluster (M=4096 kernel) walked in the most recent fused
// channelizer launch, 0 if none yet: lets a test aim its spot checks at the chunk boundaries
extern std::atomic<unsigned long long> g_last_chunk_frames;
// SMs the persistent one-CTA-per-SM kernels leave free (liquid_cuda_set_reserved_sms): room for a
// collective's kernels or a small side-stream launch to run beside them instead of after them
extern std::atomic<int> g_reserved_sms;
inline int usable_sms(int sms)
{
    const int r = g_reserved_sms.load(std::memory_order_relaxed);
    return (r > 0 && r < sms) ? sms - r : sms;
}
inline void count_launch(unsigned long long n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// after every kernel launch
#define LC_LAUNCH_CHECK()                          \
    do {                                           \
        LC_CUDA_CHECK(cudaGetLastError());         \
        lc::count_launch();                        \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void *get(size_t bytes)
    {
        if (bytes > cap) {
            if (p) LC_CUDA_CHECK(cudaFree(p));
            LC_CUDA_CHECK(cudaMalloc(&p, bytes));
            cap = bytes;
        }
        return p;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Function attributes (opt-in dynamic shared memory) belong to a device: a process that drives several GPUs
// through liquid_cuda_set_device must configure each kernel once per device.  `mask` is a function-local
// static of the caller; returns true the first time it is asked on the current device.
inline bool first_time_on_device(std::atomic<unsigned long long> &mask)
{
    int dev = 0;
    LC_CUDA_CHECK(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    return (mask.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool is_pow2(unsigned int n) { return n && !(n & (n - 1)); }
inline unsigned int ilog2(unsigned int n)
{
    unsigned int l = 0;
    while ((1u << l) < n) l++;
    return l;
}

// ---- host-side prototype design (float32, reference formulas) : design.cu ----
float kaiser_beta_As(float As);
float kaiser_window(unsigned int n, unsigned int N, float beta, float mu);   // kaiser(), src/math/src/math.c:289-310
int firdes_kaiser(unsigned int n, float fc, float As, float mu, float *h);   // 0 ok, else message id
int firdes_rrcos(unsigned int k, unsigned int m, float beta, float dt, float *h);
float estimate_req_filter_As(float df, unsigned int N);
unsigned int estimate_req_filter_len(float df, float As);                   // firdes.c:52-75
float sinc_value(float x);                                                   // sincf(), src/math/src/math.c:128-139
void filter_isi(const float *h, unsigned int k, unsigned int m, float *rms, float *peak);
float rkaiser_approximate_rho(unsigned int m, float beta);
int firdes_arkaiser(unsigned int k, unsigned int m, float beta, float dt, float *h);          // 0 ok, else message id
int firdes_rkaiser(unsigned int k, unsigned int m, float beta, float dt, float *h, float *rho_out);
int firdespm_bandpass(unsigned int h_len, unsigned int num_bands, const float *bands, const float *des, const float *weights,
                      const int *wtype, float *h);                           // firdespm.c (band-pass type); 0 ok
int firdes_hM3(unsigned int k, unsigned int m, float beta, float dt, float *h);  // hM3.c:49-126; 0 ok, else message id

// ---- batched FFT building block shared by the channelizers and the fft_* API : fft.cu ----
// d_tw: twiddle table of n entries e^{sign j 2 pi i/n} (only used for non power-of-two n)
struct FftPlanDev {
    unsigned int n = 0;
    int sign = 0;              // -1 forward (e^{-j}), +1 backward (e^{+j})
    float2 *d_tw = nullptr;    // [n], device
    float2 *d_twp = nullptr;   // [n], device: the same values in the pass order of the shared-memory FFT (power-of-two n)
    float2 *d_tw_tile = nullptr;   // n = 64 ... 512: the table of the tile transform (tile_fft.cuh); n = 1024 ... 4096: of the quad transform (quad_fft.cuh)
    unsigned char radices[16] = {0};   // mixed-radix factorisation of a composite n (factors 2 ... 13), n_rad = 0: none
    unsigned int n_rad = 0;
    void init(unsigned int n_, int sign_);
    void destroy();
    // y[b][k] = post( sum_i x[b][i] e^{sign j 2 pi i k / n} ); post = identity (mode 0),
    // divide by div (mode 1) or multiply by mul0 then mul1 (mode 2).  in-place allowed.
    void run(const float2 *d_x, float2 *d_y, unsigned long long batch, cudaStream_t st,
             int post_mode = 0, float a = 1.0f, float b = 1.0f) const;
};

}  // namespace lc

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace lcd {

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// sample idx of a stream whose samples [-H, 0) live in hist[0..H) and [0, ..) in x
__device__ __forceinline__ float2 ld_stream(const float2 *__restrict__ hist, const float2 *__restrict__ x,
                                            long long idx, int H)
{
    return idx >= 0 ? __ldg(x + idx) : __ldg(hist + (H + idx));
}

}  // namespace lcd
#endif
