// pfbch2_small.cuh -- launchers of the fused firpfbch2 kernels for M = 4 / 8 / 16 / 32 (pfbch2_small.cu).
#pragma once
#include "lc_common.cuh"

namespace lc {
namespace pfb {

// analyzer: L = 2m in {4, 8, 10}; covers every frame.  d_taps_scaled: analyzer tap table [2][L][M] times 1/M.
// d_y must be 16-byte aligned.
bool small_pfbch2_supported(unsigned int M, unsigned int L);
long long launch_pfbch2_small(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps_scaled, float2 *d_y,
                              unsigned int M, unsigned int L, unsigned long long n_frames, int parity0, cudaStream_t st);

// synthesizer: n_taps = 4m in {8, 16}; covers all n_frames (>= n_taps) or returns 0.  d_x must be 16-byte aligned;
// d_hist_out (may not alias d_hist_rows) receives the rows of v of the last n_taps frames.
bool small_pfbch2_synth_supported(unsigned int M, unsigned int n_taps);
long long launch_pfbch2_small_synth(const float2 *d_hist_rows, int hist_rows, const float2 *d_x, const float *d_taps,
                                    float2 *d_y, float2 *d_hist_out, unsigned int M, unsigned int n_taps,
                                    unsigned long long n_frames, int parity0, cudaStream_t st);

// firpfbch (critically sampled) analyzer, M = 4 / 8 / 32, p = 2m in {4, 8, 10}: same contract as launch_pfbch_fused
// (pfbch_fused.cuh), every frame covered.  d_x and d_y must be 16-byte aligned.
bool small_pfbch_supported(unsigned int M, unsigned int p);
long long launch_pfbch_small(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, float2 *d_y, unsigned int M,
                             unsigned int p, unsigned long long n_frames, unsigned int n_streams, long long stream_stride,
                             cudaStream_t st);

}  // namespace pfb
}  // namespace lc
