// pfbch2_tile.cuh -- launcher of the fused firpfbch2 analysis kernels for M = 64 ... 512 (pfbch2_tile.cu).
#pragma once
#include "lc_common.cuh"

namespace lc {
namespace pfb {

// shapes the tile kernel is instantiated for (M = 64 / 128 / 256 / 512, L = 2m = 4 / 8 / 10 / 14)
bool tile_pfbch2_supported(unsigned int M, unsigned int L);

// twiddle table the kernel expects (uploaded once per object)
void tile_pfbch2_make_twiddles(std::vector<float2> &tw, unsigned int M);

// Runs the kernel over the largest multiple of 8 frames <= n_frames and returns how many frames it covered
// (the caller finishes the remaining 0..7 frames with the generic kernels).
// d_taps_scaled: analyzer tap table [2][L][M] already multiplied by 1/M.
long long launch_pfbch2_tile(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps_scaled,
                             const float2 *d_tw, float2 *d_y, unsigned int M, unsigned int L,
                             unsigned long long n_frames, int parity0, cudaStream_t st);

// Synthesizer: M = 64 ... 512 with n_taps = 4m in {8, 16, 20}; M = 1024 with n_taps in {8, 16}.  Covers all n_frames (>= n_taps) or returns 0 (caller takes the two-kernel
// path).  d_hist_rows: hist_rows rows of v (scaled IDFT output) preceding frame 0; d_hist_out (may not alias it)
// receives the rows of v of the last n_taps frames.  d_taps: [n_taps][M/2]; d_tw: the analyzer's table.
bool tile_pfbch2_synth_supported(unsigned int M, unsigned int n_taps);
void tile_pfbch2_synth_make_twiddles(std::vector<float2> &tw, unsigned int M);   // M = 1024 has its own table
long long launch_pfbch2_tile_synth(const float2 *d_hist_rows, int hist_rows, const float2 *d_x, const float *d_taps,
                                   const float2 *d_tw, float2 *d_y, float2 *d_hist_out, unsigned int M, unsigned int n_taps,
                                   unsigned long long n_frames, int parity0, cudaStream_t st);

// firpfbch (critically sampled) analyzer, M = 64 / 128 / 512, p = 2m in {4, 8, 10}: same contract as
// launch_pfbch_fused (pfbch_fused.cuh) except that every frame is covered (ragged last tile in the kernel).
bool tile_pfbch_supported(unsigned int M, unsigned int p);
void tile_pfbch_make_twiddles(std::vector<float2> &tw, unsigned int M);
long long launch_pfbch_tile(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, const float2 *d_tw,
                            float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                            unsigned int n_streams, long long stream_stride, cudaStream_t st);

// firpfbch synthesizer, M = 64 / 128 / 512, p in {4, 8, 10}: same contract as launch_pfbch_synth_fused (pfbch_fused.cuh),
// every frame covered; d_tw: tile_pfbch2_make_twiddles (backward).
bool tile_pfbch_synth_supported(unsigned int M, unsigned int p);
long long launch_pfbch_tile_synth(const float2 *d_hist_rows, int hist_rows, const float2 *d_x, const float *d_taps,
                                  const float2 *d_tw, float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                                  unsigned int n_streams, long long stream_stride, cudaStream_t st);

// the same analyzer with complex taps (firpfbch_cccf), M = 64 / 128 / 256 / 512; d_tw: tile_pfbch_make_twiddles (forward)
bool tile_pfbch_cc_supported(unsigned int M, unsigned int p);
long long launch_pfbch_tile_cc(const float2 *d_hist, const float2 *d_x, int H, const float2 *d_taps, const float2 *d_tw,
                               float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames, cudaStream_t st);

}  // namespace pfb
}  // namespace lc
