// firfilt_kernels.cuh -- launchers of the FIR block-filter kernels (firfilt_kernels.cu).
#pragma once
#include "lc_common.cuh"

namespace lc {
namespace firfilt {

// y[s][i] = scale * sum_{k<n_taps} h[k] x[s][i-k] for n_streams streams of n samples (stream s at x + s*n).
// Samples before the start of a stream come from hist (last H samples, single-stream stateful use) or
// are zero when hist == nullptr.
// d_tile_counter: one device word owned by the calling object (zeroed here, in stream order) that the
// persistent short-filter kernel uses to hand out tiles dynamically; nullptr = static striding.
void launch(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, unsigned int n_taps,
            float scale, float2 *d_y, unsigned long long n, unsigned int n_streams, cudaStream_t st,
            unsigned int *d_tile_counter = nullptr);

// y = scale * sum_k h[k] * hist[n_taps-1-k]   (output for the buffer as it stands)
void launch_state_output(const float2 *d_hist, const float *d_taps, unsigned int n_taps, float scale,
                         float2 *d_y, cudaStream_t st);

// n_vec dot products y[v] = sum_i h[i] x[v*stride + i]
void launch_dotprod_batch(const float *d_h, unsigned int n, const float2 *d_x, unsigned long long n_vec,
                          unsigned long long stride, float2 *d_y, cudaStream_t st);

// firpfb: for i < n: window <- x[i]; y[i] = scale * sum_k sub[idx[i]][k] * window[k]
// n_filters: rows of d_sub (0 = unknown: the bank is not staged in shared memory)
void launch_firpfb_block(const float2 *d_hist, const float2 *d_x, int H, const float *d_sub, unsigned int sub_len,
                         const unsigned int *d_idx, float scale, float2 *d_y, unsigned long long n, cudaStream_t st,
                         unsigned int n_filters = 0);

}  // namespace firfilt
}  // namespace lc
