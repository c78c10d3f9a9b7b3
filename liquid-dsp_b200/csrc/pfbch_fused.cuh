// pfbch_fused.cuh -- launcher of the fused firpfbch analysis kernels (pfbch_fused.cu).
#pragma once
#include "lc_common.cuh"

namespace lc {
namespace pfb {

bool fused_pfbch_supported(unsigned int M, unsigned int p);

bool fused_pfbch_synth_supported(unsigned int M, unsigned int p);

// inter-pass twiddles of the M = 256 kernels: e^{sign j 2 pi k1 r2 / 256} (-1 analyzer, +1 synthesizer)
void fused_pfbch_make_twiddles(std::vector<float2> &tw, int sign, unsigned int M = 256, unsigned int p = 4);

// Runs the fused analyzer over the largest multiple of 16 frames <= n_frames of each of n_streams
// streams (stream s at d_x + s*stream_stride, output at d_y + s*stream_stride) and returns how many
// frames per stream it covered.  d_hist: the H samples preceding d_x[0] (single-stream stateful use) or
// nullptr for the zero state.  d_x and d_y must be 16-byte aligned, pushes aligned (phase 0).
long long launch_pfbch_fused(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, const float2 *d_tw,
                             float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                             unsigned int n_streams, long long stream_stride, cudaStream_t st);

// Fused synthesizer (M = 256): covers the largest multiple of 16 frames <= n_frames of each stream and
// returns the frames covered.  d_hist_rows: n_hist_rows rows of IDFT output preceding frame 0 (stateful
// single-stream use) or nullptr for the zero state.
long long launch_pfbch_synth_fused(const float2 *d_hist_rows, int n_hist_rows, const float2 *d_x, const float *d_taps,
                                   const float2 *d_tw, float2 *d_y, unsigned int M, unsigned int p,
                                   unsigned long long n_frames, unsigned int n_streams, long long stream_stride,
                                   cudaStream_t st);

}  // namespace pfb
}  // namespace lc
