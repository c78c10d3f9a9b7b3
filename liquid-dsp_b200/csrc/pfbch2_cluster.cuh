// pfbch2_cluster.cuh -- launcher of the cluster-of-8 fused firpfbch2 analysis kernel for M = 4096
// (pfbch2_cluster.cu).
#pragma once
#include "lc_common.cuh"

namespace lc {
namespace pfb {

// shapes the cluster kernel is instantiated for
bool cluster_pfbch2_supported(unsigned int M, unsigned int L);

// twiddle tables the kernel expects (uploaded once per object)
void cluster_pfbch2_make_twiddles(std::vector<float2> &tw);

// clusters of 8 CTAs the device keeps resident (-1 before the first launch, 0 = cannot be scheduled)
int cluster_pfbch2_max_clusters();

// Runs the kernel over the largest multiple of 8 frames <= n_frames and returns how many frames it
// covered (0 when clusters of 8 cannot be scheduled; the caller finishes the rest with the generic
// kernels).  d_taps_scaled: analyzer tap table [2][L][M] already multiplied by 1/M.
long long launch_pfbch2_cluster(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps_scaled,
                                const float2 *d_tw, float2 *d_y, unsigned int M, unsigned int L,
                                unsigned long long n_frames, int parity0, cudaStream_t st);

}  // namespace pfb
}  // namespace lc
