// pfbch2_fused.cuh -- launcher of the fused firpfbch2 analysis kernel (pfbch2_fused.cu).
#pragma once
#include "lc_common.cuh"

namespace lc {
namespace pfb {

// shapes the fused kernel is instantiated for
bool fused_pfbch2_supported(unsigned int M, unsigned int L);

// inter-stage twiddle table the kernel expects (uploaded once per object)
void fused_pfbch2_make_twiddles(std::vector<float2> &tw);

// Runs the fused kernel over the largest multiple of 4 frames (one group) <= n_frames and returns how
// many frames it covered (the caller finishes the remaining 0..3 frames with the generic kernels).
// d_taps_scaled: analyzer tap table [2][L][M] already multiplied by 1/M.  d_x must be 16-byte aligned.
long long launch_pfbch2_fused(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps_scaled,
                              const float2 *d_tw, float2 *d_y, unsigned int M, unsigned int L,
                              unsigned long long n_frames, int parity0, cudaStream_t st);

}  // namespace pfb
}  // namespace lc
