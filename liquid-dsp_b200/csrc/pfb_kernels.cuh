// pfb_kernels.cuh -- launchers of the polyphase multiply-accumulate kernels (pfb_kernels.cu).
#pragma once
#include "lc_common.cuh"

namespace lc {
namespace pfb {

// firpfbch2 analysis MAC: X[t][w] = sum_n taps[par(t)][n][w] * (n-th newest sample of window w at frame
// t_first + t); parity0 is the parity of frame 0 (the frame that starts at x[0]); d_X points at row t_first
void launch_pfbch2_mac(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, float2 *d_X,
                       unsigned int M, unsigned int L, unsigned long long n_frames, int parity0, cudaStream_t st,
                       long long t_first = 0);

// firpfbch2 synthesis FIR over scaled-IDFT rows v: y[t][i] = sum_q taps[q][i] * v[t-q][i + (par(t) ? M/2 : 0)]
void launch_pfbch2_synth_fir(const float2 *d_hist_rows, const float2 *d_v, int hist_rows, const float *d_taps,
                             float2 *d_y, unsigned int M, unsigned int n_taps, unsigned long long n_frames,
                             int parity0, cudaStream_t st);

// firpfbch analysis MAC with arbitrary push phase and alignment index k:
// X[s][M-1-i] = sum_n h[i + nM] * (n-th newest sample of window (i+k)%M after first + s*M samples of x);
// `phase` = pushes since reset (mod M) before x[0]
void launch_pfbch_mac(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, float2 *d_X,
                      unsigned int M, unsigned int p, unsigned long long n_frames, unsigned int phase,
                      unsigned int first, unsigned int k, cudaStream_t st);

// firpfbch synthesis FIR over IDFT rows v: y[s][i] = sum_n taps[n][i] * v[s-n][i]
void launch_pfbch_synth_fir(const float2 *d_hist_rows, const float2 *d_v, int hist_rows, const float *d_taps,
                            float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames, cudaStream_t st);

// One-kernel analyzers for power-of-two M whose transform fits shared memory (M <= 4096 at 2 buffers):
// polyphase MAC straight into shared memory, Stockham FFT there, scaled store.  d_tw: the FftPlanDev
// table of the object (backward for firpfbch2, forward for firpfbch).
bool analyze_pow2_supported(unsigned int M);
void launch_pfbch2_analyze_pow2(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, const float2 *d_tw,
                                float2 *d_y, unsigned int M, unsigned int L, unsigned long long n_frames, int parity0,
                                cudaStream_t st, long long t_first = 0);
void launch_pfbch_analyze_pow2(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, const float2 *d_tw,
                               float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                               unsigned int phase, unsigned int first, unsigned int k, cudaStream_t st);
void launch_pfbch_analyze_pow2_cc(const float2 *d_hist, const float2 *d_x, int H, const float2 *d_taps, const float2 *d_tw,
                                  float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                                  unsigned int phase, unsigned int first, unsigned int k, cudaStream_t st);

// the same two with complex taps (firpfbch_cccf): plain complex product, no conjugate
void launch_pfbch_mac_cc(const float2 *d_hist, const float2 *d_x, int H, const float2 *d_taps, float2 *d_X,
                         unsigned int M, unsigned int p, unsigned long long n_frames, unsigned int phase,
                         unsigned int first, unsigned int k, cudaStream_t st);
void launch_pfbch_synth_fir_cc(const float2 *d_hist_rows, const float2 *d_v, int hist_rows, const float2 *d_taps,
                               float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames, cudaStream_t st);

}  // namespace pfb
}  // namespace lc
