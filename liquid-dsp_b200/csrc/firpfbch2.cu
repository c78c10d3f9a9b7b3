// firpfbch2.cu -- firpfbch2_crcf: polyphase filterbank channelizer with output rate 2 Fs / M.
//
// Replaces src/multichannel/src/firpfbch2.c (create :66-128, create_kaiser :135-183, reset :213-225,
// analyzer :244-282, synthesizer :287-335).  Same API, state and alignment semantics; the per-frame
// window objects become a flat history of the last 2mM samples.
//
// Data flow of the analyzer (derived in DESIGN.md):
//   frame t consumes x[t*M/2 .. (t+1)*M/2);   sample x[t*M/2 + i] belongs to window
//   w = M/2-1-i (t even) or M-1-i (t odd);    window w is filtered with branch w on even frames
//   and branch (w+M/2)%M on odd frames;       the result is input w of an unscaled backward DFT;
//   the output is divided by M.
#include <cmath>
#include <cstdint>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "pfb_kernels.cuh"
#include "host_pipeline.cuh"
#include "pfbch2_fused.cuh"
#include "pfbch2_cluster.cuh"
#include "pfbch2_tile.cuh"
#include "pfbch2_small.cuh"

struct firpfbch2_crcf_s {
    int type;
    unsigned int M, M2, m, L;      // L = 2m taps per branch
    unsigned int H;                // history length kept as state (samples)
    std::vector<float> h;          // prototype, first 2mM taps (host copy for print)
    float *d_taps = nullptr;       // analyzer: [2][L][M] (parity, tap, window); synthesizer: [4m][M/2]
    float *d_taps_scaled = nullptr;// analyzer, fused kernel only: d_taps * (1/M)
    float2 *d_tw_fused = nullptr;  // fused kernel's inter-stage twiddles
    bool fused = false;            // M = 1024: one CTA per frame (pfbch2_fused.cu)
    bool clustered = false;        // M = 4096: one frame per cluster of 8 CTAs (pfbch2_cluster.cu)
    bool small = false;            // M = 4 ... 32: several instances per warp (pfbch2_small.cu)
    bool tiled = false;            // M = 64 ... 512: register windows + shared-memory tile of 8 frames (pfbch2_tile.cu)
    float2 *d_hist = nullptr;      // last H samples (analyzer: inputs; synthesizer: scaled IDFT rows)
    int flag = 0;                  // frame parity (firpfbch2.c:57)
    lc::FftPlanDev fft;            // M-point backward transform
    lc::DevBuf scratch, stage_in, stage_out, hist_tmp;
    cudaStream_t stream = nullptr;
};

namespace {

// the stateful core on device pointers; hist/x/y may not alias
void analyze_dev(firpfbch2_crcf q, const float2 *d_hist, const float2 *d_x, int parity0,
                 unsigned long long n_frames, float2 *d_y, cudaStream_t st)
{
    if (n_frames == 0) return;
    unsigned long long done = 0;
    if (q->fused && n_frames >= 32 && ((uintptr_t)d_x % 16) == 0 && !getenv("LIQUID_CUDA_NO_FUSED"))
        done = (unsigned long long)lc::pfb::launch_pfbch2_fused(d_hist, d_x, (int)q->H, q->d_taps_scaled, q->d_tw_fused,
                                                                d_y, q->M, q->L, n_frames, parity0, st);
    else if (q->clustered && n_frames >= 32 && !getenv("LIQUID_CUDA_NO_FUSED"))
        done = (unsigned long long)lc::pfb::launch_pfbch2_cluster(d_hist, d_x, (int)q->H, q->d_taps_scaled, q->d_tw_fused,
                                                                  d_y, q->M, q->L, n_frames, parity0, st);
    else if (q->small && ((uintptr_t)d_y % 16) == 0 && ((uintptr_t)d_x % 16) == 0 && !getenv("LIQUID_CUDA_NO_FUSED"))
        done = (unsigned long long)lc::pfb::launch_pfbch2_small(d_hist, d_x, (int)q->H, q->d_taps_scaled, d_y, q->M, q->L,
                                                                n_frames, parity0, st);
    else if (q->tiled && !getenv("LIQUID_CUDA_NO_FUSED"))
        done = (unsigned long long)lc::pfb::launch_pfbch2_tile(d_hist, d_x, (int)q->H, q->d_taps_scaled, q->d_tw_fused,
                                                               d_y, q->M, q->L, n_frames, parity0, st);
    if (done == n_frames) return;
    // shape-generic path (and the few frames the fused kernels leave)
    float2 *d_yr = d_y + done * q->M;
    // large blocks of a shape whose transform runs on the tile / quad kernels: column-walking MAC kernel + batched transform
    // (two kernels at the HBM rate) beat the one-kernel shared-memory analyzer, whose fill re-reads every sample 2m times
    const bool two_fast = q->M % 64 == 0 && n_frames - done >= 64 && q->fft.d_tw_tile != nullptr && !getenv("LIQUID_CUDA_ONE_KERNEL") &&
                          !getenv("LIQUID_CUDA_NO_FUSED") && !getenv("LIQUID_CUDA_MAC_PLAIN");
    if (lc::pfb::analyze_pow2_supported(q->M) && !getenv("LIQUID_CUDA_TWO_KERNEL") && !two_fast) {
        // power-of-two M: polyphase MAC, backward FFT and /M in one kernel (shared memory)
        lc::pfb::launch_pfbch2_analyze_pow2(d_hist, d_x, (int)q->H, q->d_taps, q->fft.d_twp, d_yr, q->M, q->L,
                                            n_frames - done, parity0, st, (long long)done);
        return;
    }
    // any other M: polyphase MAC into the output buffer, then an in-place batched backward DFT with /M
    lc::pfb::launch_pfbch2_mac(d_hist, d_x, (int)q->H, q->d_taps, d_yr, q->M, q->L, n_frames - done, parity0, st,
                               (long long)done);
    q->fft.run(d_yr, d_yr, n_frames - done, st, /*post: divide*/ 1, (float)q->M, 1.0f);
}

void synthesize_dev(firpfbch2_crcf q, const float2 *d_hist, const float2 *d_x, int parity0,
                    unsigned long long n_frames, float2 *d_y, float2 *d_v, cudaStream_t st)
{
    if (n_frames == 0) return;
    // v = IDFT(X) * (1/M) * (M/2), two float multiplies as firpfbch2.c:301-305
    q->fft.run(d_x, d_v, n_frames, st, /*post: two multiplies*/ 2, 1.0f / (float)q->M, (float)q->M2);
    lc::pfb::launch_pfbch2_synth_fir(d_hist, d_v, (int)(q->H / q->M), q->d_taps, d_y, q->M, 2 * q->L, n_frames, parity0, st);
}

}  // namespace

extern "C" {

firpfbch2_crcf firpfbch2_crcf_create(int _type, unsigned int _M, unsigned int _m, float *_h)
{
    if (_type != LIQUID_ANALYZER && _type != LIQUID_SYNTHESIZER) {
        fprintf(stderr, "error: firpfbch2_%s_create(), invalid type %d\n", "crcf", _type);
        exit(1);
    } else if (_M < 2 || _M % 2) {
        lc::fatal("firpfbch2_crcf_create", "number of channels must be greater than 2 and even");
    } else if (_m < 1) {
        lc::fatal("firpfbch2_crcf_create", "filter semi-length must be at least 1");
    }
    lc::ensure_device();

    firpfbch2_crcf q = new firpfbch2_crcf_s();
    q->type = _type;
    q->M = _M;
    q->M2 = _M / 2;
    q->m = _m;
    q->L = 2 * _m;
    const unsigned int M = _M, M2 = _M / 2, L = q->L;
    q->h.assign(_h, _h + (size_t)L * M);          // only the first 2mM taps are used (firpfbch2.c:104-105)

    std::vector<float> taps;
    if (_type == LIQUID_ANALYZER) {
        // taps[par][n][w]: tap n (n-th newest sample) of the branch that meets window w on a frame of
        // parity par: branch w (even), branch (w + M/2) % M (odd)          firpfbch2.c:260-271
        taps.resize((size_t)2 * L * M);
        for (unsigned int par = 0; par < 2; par++)
            for (unsigned int n = 0; n < L; n++)
                for (unsigned int w = 0; w < M; w++) {
                    unsigned int br = par ? (w + M2) % M : w;
                    taps[((size_t)par * L + n) * M + w] = _h[br + n * M];
                }
        q->H = L * M;                              // >= 2mM - M/2 samples ever looked back at
    } else {
        // taps[qq][i] = h[i + qq*M/2], qq < 4m: output i of a frame sums taps over the last 4m
        // scaled-IDFT rows                                              firpfbch2.c:313-333
        taps.resize((size_t)2 * L * M2);
        for (unsigned int qq = 0; qq < 2 * L; qq++)
            for (unsigned int i = 0; i < M2; i++)
                taps[(size_t)qq * M2 + i] = _h[i + qq * M2];
        q->H = (2 * L) * M;                        // 4m rows of M (4m-1 needed)
    }
    LC_CUDA_CHECK(cudaMalloc(&q->d_taps, taps.size() * sizeof(float)));
    LC_CUDA_CHECK(cudaMemcpy(q->d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    const bool want_fused = _type == LIQUID_ANALYZER && lc::pfb::fused_pfbch2_supported(M, L);
    const bool want_cluster = _type == LIQUID_ANALYZER && lc::pfb::cluster_pfbch2_supported(M, L);
    const bool want_tile = _type == LIQUID_ANALYZER && lc::pfb::tile_pfbch2_supported(M, L);
    if (_type == LIQUID_SYNTHESIZER && lc::pfb::small_pfbch2_synth_supported(M, 2 * L)) q->small = true;
    const bool want_small = _type == LIQUID_ANALYZER && lc::pfb::small_pfbch2_supported(M, L);
    if (want_small) {
        std::vector<float> ts(taps);
        for (float &v : ts) v *= 1.0f / (float)M;
        LC_CUDA_CHECK(cudaMalloc(&q->d_taps_scaled, ts.size() * sizeof(float)));
        LC_CUDA_CHECK(cudaMemcpy(q->d_taps_scaled, ts.data(), ts.size() * sizeof(float), cudaMemcpyHostToDevice));
        q->small = true;
    }
    if (_type == LIQUID_SYNTHESIZER && lc::pfb::tile_pfbch2_synth_supported(M, 2 * L)) {
        std::vector<float2> tw;
        lc::pfb::tile_pfbch2_synth_make_twiddles(tw, M);
        LC_CUDA_CHECK(cudaMalloc(&q->d_tw_fused, tw.size() * sizeof(float2)));
        LC_CUDA_CHECK(cudaMemcpy(q->d_tw_fused, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
        q->tiled = true;
    }
    if (want_fused || want_cluster || want_tile) {
        // the fused kernels fold the 1/M output scale into the taps (exact: M is a power of two)
        std::vector<float> ts(taps);
        for (float &v : ts) v *= 1.0f / (float)M;
        LC_CUDA_CHECK(cudaMalloc(&q->d_taps_scaled, ts.size() * sizeof(float)));
        LC_CUDA_CHECK(cudaMemcpy(q->d_taps_scaled, ts.data(), ts.size() * sizeof(float), cudaMemcpyHostToDevice));
        std::vector<float2> tw;
        if (want_fused)        lc::pfb::fused_pfbch2_make_twiddles(tw);
        else if (want_cluster) lc::pfb::cluster_pfbch2_make_twiddles(tw);
        else                   lc::pfb::tile_pfbch2_make_twiddles(tw, M);
        LC_CUDA_CHECK(cudaMalloc(&q->d_tw_fused, tw.size() * sizeof(float2)));
        LC_CUDA_CHECK(cudaMemcpy(q->d_tw_fused, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
        q->fused = want_fused;
        q->clustered = want_cluster;
        q->tiled = want_tile;
    }
    LC_CUDA_CHECK(cudaMalloc(&q->d_hist, sizeof(float2) * q->H));
    q->fft.init(M, +1);                            // backward, unscaled (firpfbch2.c:115)
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    firpfbch2_crcf_reset(q);
    return q;
}

firpfbch2_crcf firpfbch2_crcf_create_kaiser(int _type, unsigned int _M, unsigned int _m, float _As)
{
    if (_type != LIQUID_ANALYZER && _type != LIQUID_SYNTHESIZER) {
        fprintf(stderr, "error: firpfbch2_%s_create_kaiser(), invalid type %d\n", "crcf", _type);
        exit(1);
    } else if (_M < 2 || _M % 2) {
        lc::fatal("firpfbch2_crcf_create_kaiser", "number of channels must be greater than 2 and even");
    } else if (_m < 1) {
        lc::fatal("firpfbch2_crcf_create_kaiser", "filter semi-length must be at least 1");
    }
    // 2Mm+1 Kaiser taps, cut-off 1/M (analyzer) or 0.5/M (synthesizer), scaled so that the float32
    // running sum equals M; create() then drops the last tap            firpfbch2.c:153-175
    unsigned int h_len = 2 * _M * _m + 1;
    std::vector<float> hf(h_len);
    float fc = (_type == LIQUID_ANALYZER) ? 1.0f / (float)_M : 0.5f / (float)_M;
    liquid_firdes_kaiser(h_len, fc, _As, 0.0f, hf.data());
    float hf_sum = 0.0f;
    for (unsigned int i = 0; i < h_len; i++) hf_sum += hf[i];
    for (unsigned int i = 0; i < h_len; i++) hf[i] = hf[i] * (float)_M / hf_sum;
    return firpfbch2_crcf_create(_type, _M, _m, hf.data());
}

void firpfbch2_crcf_destroy(firpfbch2_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    cudaFree(_q->d_taps);
    cudaFree(_q->d_taps_scaled);
    cudaFree(_q->d_tw_fused);
    cudaFree(_q->d_hist);
    _q->fft.destroy();
    _q->scratch.release();
    _q->stage_in.release();
    _q->stage_out.release();
    _q->hist_tmp.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void firpfbch2_crcf_reset(firpfbch2_crcf _q)
{
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_hist, 0, sizeof(float2) * _q->H, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    _q->flag = 0;
}

void firpfbch2_crcf_print(firpfbch2_crcf _q)
{
    printf("firpfbch2_%s:\n", "crcf");
    printf("    channels    :   %u\n", _q->M);
    printf("    h_len       :   %u\n", 2 * _q->M * _q->m);
    printf("    semi-length :   %u\n", _q->m);
    // per-branch taps in the order the reference's dotprod objects hold them (reversed sub-filters)
    for (unsigned int i = 0; i < _q->M; i++) {
        printf("dotprod_crcf [cuda, %u coefficients]\n", _q->L);
        for (unsigned int n = 0; n < _q->L; n++)
            printf("  %3u : %12.9f\n", n, _q->h[i + (_q->L - 1 - n) * _q->M]);
    }
}

unsigned int firpfbch2_crcf_get_state_len(firpfbch2_crcf _q) { return _q->H; }
unsigned int firpfbch2_crcf_get_num_channels(firpfbch2_crcf _q) { return _q->M; }

void firpfbch2_crcf_get_state(firpfbch2_crcf _q, liquid_float_complex *_hist, int *_flag)
{
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    LC_CUDA_CHECK(cudaMemcpy(_hist, _q->d_hist, sizeof(float2) * _q->H, cudaMemcpyDeviceToHost));
    if (_flag) *_flag = _q->flag;
}

void firpfbch2_crcf_set_state(firpfbch2_crcf _q, const liquid_float_complex *_hist, int _flag)
{
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    LC_CUDA_CHECK(cudaMemcpy(_q->d_hist, _hist, sizeof(float2) * _q->H, cudaMemcpyHostToDevice));
    _q->flag = _flag & 1;
}

void firpfbch2_crcf_execute_block_dev(firpfbch2_crcf _q, const void *_d_x, unsigned long long _n_frames,
                                      void *_d_y, void *_stream)
{
    cudaStream_t st = lc::as_stream(_stream);
    if (_n_frames == 0) return;
    if (_q->type == LIQUID_ANALYZER) {
        analyze_dev(_q, _q->d_hist, (const float2 *)_d_x, _q->flag, _n_frames, (float2 *)_d_y, st);
        lc::roll_history(_q->d_hist, _q->H, _q->hist_tmp, (const float2 *)_d_x, _n_frames * _q->M2, st);
    } else if (_q->small && _n_frames >= 2 * _q->L && ((uintptr_t)_d_x % 16) == 0 && !getenv("LIQUID_CUDA_NO_FUSED")) {
        float2 *d_new = (float2 *)_q->hist_tmp.get(sizeof(float2) * _q->H);
        lc::pfb::launch_pfbch2_small_synth(_q->d_hist, (int)(_q->H / _q->M), (const float2 *)_d_x, _q->d_taps, (float2 *)_d_y,
                                           d_new, _q->M, 2 * _q->L, _n_frames, _q->flag, st);
        LC_CUDA_CHECK(cudaMemcpyAsync(_q->d_hist, d_new, sizeof(float2) * _q->H, cudaMemcpyDeviceToDevice, st));
    } else if (_q->tiled && _n_frames >= 2 * _q->L && !getenv("LIQUID_CUDA_NO_FUSED")) {
        // fused synthesizer: the rows of v stay on the chip, the kernel leaves the last 4m of them as the new state
        float2 *d_new = (float2 *)_q->hist_tmp.get(sizeof(float2) * _q->H);
        lc::pfb::launch_pfbch2_tile_synth(_q->d_hist, (int)(_q->H / _q->M), (const float2 *)_d_x, _q->d_taps, _q->d_tw_fused,
                                          (float2 *)_d_y, d_new, _q->M, 2 * _q->L, _n_frames, _q->flag, st);
        LC_CUDA_CHECK(cudaMemcpyAsync(_q->d_hist, d_new, sizeof(float2) * _q->H, cudaMemcpyDeviceToDevice, st));
    } else {
        float2 *d_v = (float2 *)_q->scratch.get(sizeof(float2) * _n_frames * _q->M);
        synthesize_dev(_q, _q->d_hist, (const float2 *)_d_x, _q->flag, _n_frames, (float2 *)_d_y, d_v, st);
        lc::roll_history(_q->d_hist, _q->H, _q->hist_tmp, d_v, _n_frames * _q->M, st);
    }
    _q->flag = (int)((_q->flag + _n_frames) & 1);
}

void firpfbch2_crcf_execute_segment_dev(firpfbch2_crcf _q, const void *_d_halo, const void *_d_x,
                                        unsigned long long _first_frame, unsigned long long _n_frames,
                                        void *_d_y, void *_stream)
{
    if (_q->type != LIQUID_ANALYZER)
        lc::fatal("firpfbch2_crcf_execute_segment_dev", "segment form is provided for the analyzer only");
    analyze_dev(_q, (const float2 *)_d_halo, (const float2 *)_d_x, (int)(_first_frame & 1), _n_frames,
                (float2 *)_d_y, lc::as_stream(_stream));
}

void firpfbch2_crcf_execute_block(firpfbch2_crcf _q, liquid_float_complex *_x, unsigned int _n_frames,
                                  liquid_float_complex *_y)
{
    const size_t in_per = (_q->type == LIQUID_ANALYZER) ? _q->M2 : _q->M;
    const size_t out_per = (_q->type == LIQUID_ANALYZER) ? _q->M : _q->M2;
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x),
                   reinterpret_cast<lc_cf *>(_y), _n_frames, in_per, out_per,
                   [&](float2 *di, size_t nf, float2 *dob, cudaStream_t st) {
                       firpfbch2_crcf_execute_block_dev(_q, di, nf, dob, st);
                   });
}

void firpfbch2_crcf_execute_sharded(firpfbch2_crcf _q, liquid_cuda_comm _c, liquid_float_complex *_x,
                                    unsigned long long _first_frame, unsigned int _n_frames, liquid_float_complex *_y)
{
    if (_q->type != LIQUID_ANALYZER)
        lc::fatal("firpfbch2_crcf_execute_sharded", "the sharded form is provided for the analyzer only");
    if (_first_frame & 1) lc::fatal("firpfbch2_crcf_execute_sharded", "a segment must start on an even frame");
    const size_t H = _q->H, n = (size_t)_n_frames * _q->M2;
    if (liquid_cuda_comm_world(_c) > 1 && n < H)
        lc::fatal("firpfbch2_crcf_execute_sharded", "a segment must hold at least the 2mM samples its right neighbour needs");
    // the object's history becomes the left neighbour's tail (rank 0: the reset state), then the segment runs
    // through the ordinary chunked H2D / compute / D2H pipeline
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_hist, 0, sizeof(float2) * H, _q->stream));
    if (liquid_cuda_comm_world(_c) > 1) {
        float2 *d_tail = (float2 *)_q->hist_tmp.get(sizeof(float2) * H);
        LC_CUDA_CHECK(cudaMemcpyAsync(d_tail, reinterpret_cast<const float2 *>(_x) + (n - H), sizeof(float2) * H,
                                      cudaMemcpyHostToDevice, _q->stream));
        liquid_cuda_comm_exchange_halo_dev(_c, _q->d_hist, d_tail, H, _q->stream);
    }
    _q->flag = 0;
    firpfbch2_crcf_execute_block(_q, _x, _n_frames, _y);
}

void firpfbch2_crcf_execute(firpfbch2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    firpfbch2_crcf_execute_block(_q, _x, 1, _y);
}

}  // extern "C"
