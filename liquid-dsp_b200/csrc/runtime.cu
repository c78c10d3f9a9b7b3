// runtime.cu -- device selection, memory helpers, launch counter, synthetic input generator.
#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"

namespace lc {

std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_reserved_sms{0};
std::atomic<unsigned long long> g_last_chunk_frames{0};
// The device this library works on: LIQUID_CUDA_DEVICE unless liquid_cuda_set_device was called (an explicit call
// wins), else whatever the calling thread's current device is.  The CUDA current device is per-thread state, so a
// thread that has not been here yet is switched to the chosen device on its first call; the sm_100 check runs once
// per device.
static std::atomic<int> g_chosen_device{-1};          // -1: none chosen (follow the thread's current device)
static std::atomic<unsigned long long> g_device_checked{0};

static void check_device(int dev)
{
    const unsigned long long bit = 1ull << (dev & 63);
    if (g_device_checked.load(std::memory_order_acquire) & bit) return;
    cudaDeviceProp prop;
    LC_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        fprintf(stderr, "liquid_cuda: device %d (%s) is sm_%d%d; this build carries sm_100a code only\n",
                dev, prop.name, prop.major, prop.minor);
        abort();
    }
    if (getenv("LIQUID_CUDA_VERBOSE"))      // one line when the library takes over a device (e.g. under LD_PRELOAD)
        fprintf(stderr, "liquid_cuda: active on device %d (%s, sm_%d%d, %d SMs)\n", dev, prop.name, prop.major, prop.minor,
                prop.multiProcessorCount);
    g_device_checked.fetch_or(bit, std::memory_order_acq_rel);
}

void ensure_device()
{
    static thread_local int t_ready_for = -2;          // the chosen device this thread was last aligned with
    const int chosen = g_chosen_device.load(std::memory_order_acquire);
    if (t_ready_for == chosen && chosen != -1) return;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        fprintf(stderr, "liquid_cuda: no CUDA device available (%s); this library has no CPU fallback\n",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        abort();
    }
    int dev = chosen;
    if (dev == -1) {
        if (const char *env = getenv("LIQUID_CUDA_DEVICE")) {
            dev = atoi(env);
            if (dev < 0 || dev >= count) {
                fprintf(stderr, "liquid_cuda: LIQUID_CUDA_DEVICE=%s out of range (have %d devices)\n", env, count);
                abort();
            }
            int expected = -1;
            g_chosen_device.compare_exchange_strong(expected, dev);
            dev = g_chosen_device.load();
        }
    }
    if (dev >= 0) {
        LC_CUDA_CHECK(cudaSetDevice(dev));
        t_ready_for = dev;
    } else {
        LC_CUDA_CHECK(cudaGetDevice(&dev));            // no choice made: the thread's own current device
    }
    check_device(dev);
}

void choose_device(int dev)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) {
        fprintf(stderr, "liquid_cuda: liquid_cuda_set_device(%d): no such device (have %d)\n", dev, count);
        abort();
    }
    g_chosen_device.store(dev, std::memory_order_release);
    ensure_device();
}

}  // namespace lc

// u(n,c) = (mix64(seed + 2n + c) >> 40) * 2^-24 - 0.5     
__device__ __forceinline__ unsigned long long lc_mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void lc_synth_input_kernel(float2 *__restrict__ x, unsigned long long seed,
                                      unsigned long long first, unsigned long long n)
{
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long k = seed + 2 * (first + i);
        float re = (float)(lc_mix64(k) >> 40) * (1.0f / 16777216.0f) - 0.5f;
        float im = (float)(lc_mix64(k + 1) >> 40) * (1.0f / 16777216.0f) - 0.5f;
        x[i] = make_float2(re, im);
    }
}

extern "C" {

const char *liquid_cuda_version(void) { return LC_VERSION; }

int liquid_cuda_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}

void liquid_cuda_set_device(int _device) { lc::choose_device(_device); }

void *liquid_cuda_malloc(unsigned long long _bytes)
{
    lc::ensure_device();
    void *p = nullptr;
    LC_CUDA_CHECK(cudaMalloc(&p, _bytes ? _bytes : 1));
    return p;
}

void liquid_cuda_free(void *_dptr) { if (_dptr) LC_CUDA_CHECK(cudaFree(_dptr)); }

void *liquid_cuda_host_alloc(unsigned long long _bytes)
{
    lc::ensure_device();
    void *p = nullptr;
    LC_CUDA_CHECK(cudaHostAlloc(&p, _bytes ? _bytes : 1, cudaHostAllocDefault));
    return p;
}

void liquid_cuda_host_free(void *_hptr) { if (_hptr) LC_CUDA_CHECK(cudaFreeHost(_hptr)); }

int liquid_cuda_host_register(void *_hptr, unsigned long long _bytes)
{
    lc::ensure_device();
    if (cudaHostRegister(_hptr, _bytes, cudaHostRegisterDefault) == cudaSuccess) return 0;
    (void)cudaGetLastError();
    return -1;
}

void liquid_cuda_host_unregister(void *_hptr)
{
    if (_hptr && cudaHostUnregister(_hptr) != cudaSuccess) (void)cudaGetLastError();
}

void liquid_cuda_memcpy_h2d(void *_dst, const void *_src, unsigned long long _bytes)
{
    LC_CUDA_CHECK(cudaMemcpy(_dst, _src, _bytes, cudaMemcpyHostToDevice));
}

void liquid_cuda_memcpy_d2h(void *_dst, const void *_src, unsigned long long _bytes)
{
    LC_CUDA_CHECK(cudaMemcpy(_dst, _src, _bytes, cudaMemcpyDeviceToHost));
}

void liquid_cuda_sync(void) { LC_CUDA_CHECK(cudaDeviceSynchronize()); }

unsigned long long liquid_cuda_launch_count(void) { return lc::g_launches.load(); }

unsigned long long liquid_cuda_last_chunk_frames(void) { return lc::g_last_chunk_frames.load(); }

void liquid_cuda_set_reserved_sms(int _n) { lc::g_reserved_sms.store(_n < 0 ? 0 : _n); }
int liquid_cuda_get_reserved_sms(void) { return lc::g_reserved_sms.load(); }

void liquid_cuda_synth_input_dev(void *_d_x, unsigned long long _seed, unsigned long long _first,
                                 unsigned long long _n, void *_stream)
{
    lc::ensure_device();
    if (_n == 0) return;
    unsigned long long blocks = (_n + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    lc_synth_input_kernel<<<(unsigned int)blocks, 256, 0, lc::as_stream(_stream)>>>((float2 *)_d_x, _seed, _first, _n);
    LC_LAUNCH_CHECK();
}

}  // extern "C"
