// host_pipeline.cuh -- host-pointer block execution shared by the *_execute_block entry points.
//
// The caller's host arrays are cut into chunks; chunk c+1's H2D copy, chunk c's kernels and chunk
// c-1's D2H copy run on three streams, double-buffered, so that with pinned host memory both copy
// engines and the SMs overlap.  The per-chunk device work (`run`) carries the object's state in
// stream order on `compute`.  Returns when all output has landed in `y`.
#pragma once
#include <algorithm>

#include "lc_common.cuh"

namespace lc {

constexpr int kPipeEvents = 4;                 // ring of per-chunk events (two chunks are in flight at most)
struct CopyStreams {
    cudaStream_t in = nullptr, out = nullptr;
    // created once per thread and device, reused by every call: a chunk takes the three events of slot c % 4
    // (a stream wait captures the record that precedes it, so re-recording a slot later is harmless)
    cudaEvent_t ev_in[kPipeEvents] = {}, ev_k[kPipeEvents] = {}, ev_out[kPipeEvents] = {};
    void ensure()
    {
        if (!in) {
            LC_CUDA_CHECK(cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking));
            LC_CUDA_CHECK(cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking));
            for (int i = 0; i < kPipeEvents; i++) {
                LC_CUDA_CHECK(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
                LC_CUDA_CHECK(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
                LC_CUDA_CHECK(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
            }
        }
    }
};

// A caller's malloc'd (pageable) array makes cudaMemcpyAsync synchronous and staged through the driver's bounce
// buffer: measured 0.64 GS/s end to end against 3.2 GS/s from pinned arrays (firpfbch2 M=1024, bench.py e2e).
// Page-locking the arrays for the duration of each call (cudaHostRegister, LIQUID_CUDA_PAGEABLE=register) is
// worse still -- 0.12 GS/s, registering 6 GB costs seconds -- so the default leaves pageable arrays to the driver;
// a caller that reuses its arrays pins them once with liquid_cuda_host_register and gets the pinned rate.
struct ScopedPin {
    void *p = nullptr;
    ScopedPin(const void *ptr, size_t bytes)
    {
        const char *e = getenv("LIQUID_CUDA_PAGEABLE");
        if (!e || strcmp(e, "register") || bytes < (8u << 20)) return;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
            (void)cudaGetLastError();
            return;
        }
        if (at.type != cudaMemoryTypeUnregistered) return;
        if (cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterDefault) == cudaSuccess) p = const_cast<void *>(ptr);
        else (void)cudaGetLastError();              // e.g. read-only mapping: fall back to the pageable path
    }
    ~ScopedPin()
    {
        if (p) cudaHostUnregister(p);
    }
    ScopedPin(const ScopedPin &) = delete;
    ScopedPin &operator=(const ScopedPin &) = delete;
};

inline size_t pipeline_chunk_bytes()
{
    static const size_t v = [] {
        const char *e = getenv("LIQUID_CUDA_CHUNK_MB");
        long mb = e ? atol(e) : 0;
        return (size_t)(mb >= 1 && mb <= 4096 ? mb : 32) << 20;
    }();
    return v;
}

// the calling thread's copy streams on the CURRENT device (streams belong to a device; a process may drive
// several through liquid_cuda_set_device)
inline CopyStreams &copy_streams()
{
    static thread_local CopyStreams cs[64];
    int dev = 0;
    LC_CUDA_CHECK(cudaGetDevice(&dev));
    CopyStreams &c = cs[dev & 63];
    c.ensure();
    return c;
}

// n_units units of work; unit u reads in_per complex samples at x + u*in_per and writes out_per at
// y + u*out_per.  run(d_in, n_units_in_chunk, d_out, stream) enqueues the device work of one chunk.
template <typename Run>
void host_block(cudaStream_t compute, DevBuf &stage_in, DevBuf &stage_out, const lc_cf *x, lc_cf *y,
                size_t n_units, size_t in_per, size_t out_per, Run run, size_t chunk_bytes = 0)
{
    if (n_units == 0) return;
    if (chunk_bytes == 0) chunk_bytes = pipeline_chunk_bytes();
    const bool alias = (const void *)x == (const void *)y;         // firfilt allows in-place blocks
    ScopedPin pin_x(x, n_units * in_per * sizeof(float2));
    ScopedPin pin_y(alias ? nullptr : y, alias ? 0 : n_units * out_per * sizeof(float2));
    const size_t per = std::max(in_per, out_per) * sizeof(float2);
    const size_t chunk_units = std::max<size_t>(1, chunk_bytes / per);
    const size_t n_chunks = (n_units + chunk_units - 1) / chunk_units;
    const size_t bufs = n_chunks > 1 ? 2 : 1;
    const size_t cu = std::min(chunk_units, n_units);
    float2 *din = (float2 *)stage_in.get(bufs * cu * in_per * sizeof(float2));
    float2 *dout = (float2 *)stage_out.get(bufs * cu * out_per * sizeof(float2));
    CopyStreams &cs = copy_streams();
    for (size_t c = 0; c < n_chunks; c++) {
        const int e = (int)(c % kPipeEvents), e_old = (int)((c + kPipeEvents - bufs) % kPipeEvents);
        const size_t u0 = c * chunk_units, nu = std::min(chunk_units, n_units - u0);
        float2 *di = din + (c % bufs) * cu * in_per;
        float2 *dob = dout + (c % bufs) * cu * out_per;
        if (c >= bufs) {                               // buffer reuse: wait for its last consumers
            LC_CUDA_CHECK(cudaStreamWaitEvent(cs.in, cs.ev_k[e_old], 0));
            LC_CUDA_CHECK(cudaStreamWaitEvent(compute, cs.ev_out[e_old], 0));
        }
        LC_CUDA_CHECK(cudaMemcpyAsync(di, x + u0 * in_per, nu * in_per * sizeof(float2), cudaMemcpyHostToDevice, cs.in));
        LC_CUDA_CHECK(cudaEventRecord(cs.ev_in[e], cs.in));
        LC_CUDA_CHECK(cudaStreamWaitEvent(compute, cs.ev_in[e], 0));
        run(di, nu, dob, compute);
        LC_CUDA_CHECK(cudaEventRecord(cs.ev_k[e], compute));
        LC_CUDA_CHECK(cudaStreamWaitEvent(cs.out, cs.ev_k[e], 0));
        LC_CUDA_CHECK(cudaMemcpyAsync(y + u0 * out_per, dob, nu * out_per * sizeof(float2), cudaMemcpyDeviceToHost, cs.out));
        LC_CUDA_CHECK(cudaEventRecord(cs.ev_out[e], cs.out));
    }
    LC_CUDA_CHECK(cudaStreamSynchronize(cs.out));
    LC_CUDA_CHECK(cudaStreamSynchronize(compute));
}

// new state = last H samples of [hist | x(n samples)], all on the device, in stream order
inline void roll_history(float2 *d_hist, size_t H, DevBuf &tmpbuf, const float2 *d_x, unsigned long long n, cudaStream_t st)
{
    if (n == 0 || H == 0) return;
    if (n >= H) {
        LC_CUDA_CHECK(cudaMemcpyAsync(d_hist, d_x + (n - H), sizeof(float2) * H, cudaMemcpyDeviceToDevice, st));
    } else {
        float2 *tmp = (float2 *)tmpbuf.get(sizeof(float2) * H);
        LC_CUDA_CHECK(cudaMemcpyAsync(tmp, d_hist + n, sizeof(float2) * (H - n), cudaMemcpyDeviceToDevice, st));
        LC_CUDA_CHECK(cudaMemcpyAsync(tmp + (H - n), d_x, sizeof(float2) * n, cudaMemcpyDeviceToDevice, st));
        LC_CUDA_CHECK(cudaMemcpyAsync(d_hist, tmp, sizeof(float2) * H, cudaMemcpyDeviceToDevice, st));
    }
}

}  // namespace lc
