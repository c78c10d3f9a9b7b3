// firpfbch_cccf.cu -- firpfbch_cccf: the critically sampled polyphase channelizer with complex taps.
//
// Replaces src/multichannel/src/firpfbch_cccf.c, which instantiates firpfbch.c (create :73-142,
// create_kaiser :150-184, create_rnyquist :193-256, reset :284-293, synthesizer :314-336, analyzer
// :346-409) with TC = float complex and dotprod_cccf (plain complex product, no conjugate).
// SURVEY.md section 8(f) item 1.
//
// Two execution paths behind one handle:
//   * taps whose imaginary parts are all zero (everything create_kaiser / create_rnyquist produce, and
//     any real prototype handed to create): an inner firpfbch_crcf object does the work, so these run on
//     the fused kernels of pfbch_fused.cu -- a complex multiply by c + 0j equals the real multiply
//     except in the sign of a zero;
//   * genuinely complex taps: the shape-generic kernels of pfb_kernels.cu with complex taps
//     (polyphase MAC -> batched FFT for the analyzer, batched IFFT -> per-branch FIR for the
//     synthesizer), same state handling as firpfbch.cu.
#include <cmath>
#include <cstdint>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "pfb_kernels.cuh"
#include "host_pipeline.cuh"
#include "pfbch2_tile.cuh"

struct firpfbch_cccf_s {
    int type;
    unsigned int M, p;
    unsigned int H;                 // state length: p*M samples (analyzer: inputs; synthesizer: IDFT rows)
    std::vector<lc_cf> h;           // M*p prototype taps (host copy, for print)
    firpfbch_crcf real = nullptr;   // set when every tap is real: the crcf object carries state and kernels
    float2 *d_taps = nullptr;       // [p][M]: taps[n][i] = h[i + nM]
    float2 *d_hist = nullptr;
    float2 *d_tw_tile = nullptr;    // forward twiddles of the complex-tap tile analyzer (M = 64 ... 512), or null
    unsigned int phase = 0;         // analyzer: pushes since reset, mod M
    lc::FftPlanDev fft;
    lc::DevBuf scratch, stage_in, stage_out, hist_tmp, one;
    cudaStream_t stream = nullptr;
};

namespace {

// complex-tap analyzer step: one kernel when M is a power of two that fits shared memory, else polyphase
// MAC into y followed by an in-place forward DFT; unscaled
void analyze_generic_cc(firpfbch_cccf q, const float2 *d_x, float2 *d_y, unsigned long long n, unsigned int first,
                        unsigned int k, cudaStream_t st)
{
    if (q->d_tw_tile && q->phase == 0 && first == q->M && k == 0 && !getenv("LIQUID_CUDA_NO_FUSED")) {
        // whole frames from an aligned state: the tile kernel of pfbch2_tile.cu with complex taps
        lc::pfb::launch_pfbch_tile_cc(q->d_hist, d_x, (int)q->H, q->d_taps, q->d_tw_tile, d_y, q->M, q->p, n, st);
        return;
    }
    if (lc::pfb::analyze_pow2_supported(q->M) && !getenv("LIQUID_CUDA_TWO_KERNEL")) {
        lc::pfb::launch_pfbch_analyze_pow2_cc(q->d_hist, d_x, (int)q->H, q->d_taps, q->fft.d_twp, d_y, q->M, q->p, n,
                                              q->phase, first, k, st);
        return;
    }
    lc::pfb::launch_pfbch_mac_cc(q->d_hist, d_x, (int)q->H, q->d_taps, d_y, q->M, q->p, n, q->phase, first, k, st);
    q->fft.run(d_y, d_y, n, st);
}

}  // namespace

extern "C" {

firpfbch_cccf firpfbch_cccf_create(int _type, unsigned int _M, unsigned int _p, liquid_float_complex *_h)
{
    if (_type != LIQUID_ANALYZER && _type != LIQUID_SYNTHESIZER) {
        fprintf(stderr, "error: firpfbch_%s_create(), invalid type %d\n", "cccf", _type);
        exit(1);
    } else if (_M == 0) {
        lc::fatal("firpfbch_cccf_create", "number of channels must be greater than 0");
    } else if (_p == 0) {
        lc::fatal("firpfbch_cccf_create", "invalid filter size (must be greater than 0)");
    }
    lc::ensure_device();
    firpfbch_cccf q = new firpfbch_cccf_s();
    q->type = _type;
    q->M = _M;
    q->p = _p;
    q->H = _p * _M;
    const lc_cf *hc = reinterpret_cast<const lc_cf *>(_h);
    q->h.assign(hc, hc + (size_t)_M * _p);
    bool all_real = true;
    for (const lc_cf &t : q->h) all_real = all_real && (t.im == 0.0f);
    if (all_real && !getenv("LIQUID_CUDA_CCCF_GENERIC")) {
        std::vector<float> hr(q->h.size());
        for (size_t i = 0; i < hr.size(); i++) hr[i] = q->h[i].re;
        q->real = firpfbch_crcf_create(_type, _M, _p, hr.data());
        return q;
    }
    // h is already laid out as taps[n][i] = h[i + nM]  (firpfbch.c:115-124 sub-samples it the same way)
    LC_CUDA_CHECK(cudaMalloc(&q->d_taps, sizeof(float2) * q->h.size()));
    LC_CUDA_CHECK(cudaMemcpy(q->d_taps, q->h.data(), sizeof(float2) * q->h.size(), cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMalloc(&q->d_hist, sizeof(float2) * q->H));
    q->fft.init(_M, _type == LIQUID_ANALYZER ? -1 : +1);     // firpfbch.c:131-135
    if (_type == LIQUID_ANALYZER && lc::pfb::tile_pfbch_cc_supported(_M, _p)) {
        std::vector<float2> tw;
        lc::pfb::tile_pfbch_make_twiddles(tw, _M);
        LC_CUDA_CHECK(cudaMalloc(&q->d_tw_tile, tw.size() * sizeof(float2)));
        LC_CUDA_CHECK(cudaMemcpy(q->d_tw_tile, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    firpfbch_cccf_reset(q);
    return q;
}

firpfbch_cccf firpfbch_cccf_create_kaiser(int _type, unsigned int _M, unsigned int _m, float _As)
{
    if (_M == 0) lc::fatal("firpfbch_cccf_create_kaiser", "number of channels must be greater than 0");
    if (_m == 0) lc::fatal("firpfbch_cccf_create_kaiser", "invalid filter size (must be greater than 0)");
    _As = fabsf(_As);
    // 2Mm+1 Kaiser taps at fc = 0.5/M, no normalisation; the first 2Mm are used  (firpfbch.c:164-180)
    const unsigned int h_len = 2 * _M * _m + 1;
    std::vector<float> h(h_len);
    liquid_firdes_kaiser(h_len, 0.5f / (float)_M, _As, 0.0f, h.data());
    std::vector<lc_cf> hc(2 * _M * _m);
    for (size_t i = 0; i < hc.size(); i++) hc[i] = lc_cf{h[i], 0.0f};
    return firpfbch_cccf_create(_type, _M, 2 * _m, reinterpret_cast<liquid_float_complex *>(hc.data()));
}

firpfbch_cccf firpfbch_cccf_create_rnyquist(int _type, unsigned int _M, unsigned int _m, float _beta, int _ftype)
{
    if (_type != LIQUID_ANALYZER && _type != LIQUID_SYNTHESIZER) {
        fprintf(stderr, "error: firpfbch_%s_create_rnyquist(), invalid type %d\n", "cccf", _type);
        exit(1);
    } else if (_M == 0) {
        lc::fatal("firpfbch_cccf_create_rnyquist", "number of channels must be greater than 0");
    } else if (_m == 0) {
        lc::fatal("firpfbch_cccf_create_rnyquist", "invalid filter size (must be greater than 0)");
    }
    if (_ftype != LIQUID_FIRFILT_ARKAISER && _ftype != LIQUID_FIRFILT_RKAISER && _ftype != LIQUID_FIRFILT_RRC &&
        _ftype != LIQUID_FIRFILT_hM3) {
        fprintf(stderr, "error: firpfbch_%s_create_rnyquist(), unknown/invalid prototype (%d)\n", "cccf", _ftype);
        exit(1);
    }
    const unsigned int g_len = 2 * _M * _m;
    std::vector<float> h(g_len + 1);
    liquid_firdes_prototype(_ftype, _M, _m, _beta, 0.0f, h.data());
    // first 2Mm taps; matched filter g(-t) for the analyzer  (firpfbch.c:238-248)
    std::vector<lc_cf> g(g_len);
    for (unsigned int i = 0; i < g_len; i++) g[i] = lc_cf{(_type == LIQUID_SYNTHESIZER) ? h[i] : h[g_len - i - 1], 0.0f};
    return firpfbch_cccf_create(_type, _M, 2 * _m, reinterpret_cast<liquid_float_complex *>(g.data()));
}

void firpfbch_cccf_destroy(firpfbch_cccf _q)
{
    if (!_q) return;
    if (_q->real) {
        firpfbch_crcf_destroy(_q->real);
    } else {
        cudaStreamSynchronize(_q->stream);
        cudaFree(_q->d_taps);
        cudaFree(_q->d_hist);
        cudaFree(_q->d_tw_tile);
        _q->fft.destroy();
        _q->scratch.release();
        _q->stage_in.release();
        _q->stage_out.release();
        _q->hist_tmp.release();
        _q->one.release();
        cudaStreamDestroy(_q->stream);
    }
    delete _q;
}

void firpfbch_cccf_reset(firpfbch_cccf _q)
{
    if (_q->real) {
        firpfbch_crcf_reset(_q->real);
        return;
    }
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_hist, 0, sizeof(float2) * _q->H, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    _q->phase = 0;          // filter_index = M-1 (firpfbch.c:292)
}

void firpfbch_cccf_print(firpfbch_cccf _q)
{
    printf("firpfbch (%s) [%u channels]:\n", _q->type == LIQUID_ANALYZER ? "analyzer" : "synthesizer", _q->M);
    for (unsigned int i = 0; i < _q->M * _q->p; i++)
        printf("  h[%3u] = %12.8f + %12.8f*j\n", i, _q->h[i].re, _q->h[i].im);
}

// device pointers, whole frames (M samples in, M samples out per frame); state carried in stream order
void firpfbch_cccf_execute_block_dev(firpfbch_cccf _q, const void *_d_x, unsigned long long _n_frames,
                                     void *_d_y, void *_stream)
{
    if (_n_frames == 0) return;
    if (_q->real) {
        firpfbch_crcf_execute_block_dev(_q->real, _d_x, _n_frames, _d_y, _stream);
        return;
    }
    cudaStream_t st = lc::as_stream(_stream);
    const float2 *d_x = (const float2 *)_d_x;
    float2 *d_y = (float2 *)_d_y;
    const unsigned int M = _q->M, p = _q->p;
    if (_q->type == LIQUID_ANALYZER) {
        // per frame: M pushes, then run(k = 0): polyphase MAC, forward FFT, unscaled
        analyze_generic_cc(_q, d_x, d_y, _n_frames, M, 0, st);
        lc::roll_history(_q->d_hist, _q->H, _q->hist_tmp, d_x, _n_frames * M, st);
    } else {
        // v = IDFT(X) per frame, then y[s][i] = sum_n h[i + nM] v[s-n][i]; new state = last p rows of v
        float2 *d_v = (float2 *)_q->scratch.get(sizeof(float2) * _n_frames * M);
        _q->fft.run(d_x, d_v, _n_frames, st);
        lc::pfb::launch_pfbch_synth_fir_cc(_q->d_hist, d_v, (int)p, _q->d_taps, d_y, M, p, _n_frames, st);
        lc::roll_history(_q->d_hist, _q->H, _q->hist_tmp, d_v, _n_frames * M, st);
    }
}

static void pfbch_cccf_host_block(firpfbch_cccf _q, liquid_float_complex *_x, unsigned int _n_frames,
                                  liquid_float_complex *_y)
{
    if (_q->real) {
        if (_q->type == LIQUID_ANALYZER) firpfbch_crcf_analyzer_execute_block(_q->real, _x, _n_frames, _y);
        else                             firpfbch_crcf_synthesizer_execute_block(_q->real, _x, _n_frames, _y);
        return;
    }
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x),
                   reinterpret_cast<lc_cf *>(_y), _n_frames, _q->M, _q->M,
                   [&](float2 *di, size_t nf, float2 *dob, cudaStream_t st) {
                       firpfbch_cccf_execute_block_dev(_q, di, nf, dob, st);
                   });
}

void firpfbch_cccf_analyzer_execute_block(firpfbch_cccf _q, liquid_float_complex *_x, unsigned int _n_frames,
                                          liquid_float_complex *_y)
{
    if (_q->type != LIQUID_ANALYZER) lc::fatal("firpfbch_cccf_analyzer_execute_block", "object is not an analyzer");
    pfbch_cccf_host_block(_q, _x, _n_frames, _y);
}

void firpfbch_cccf_synthesizer_execute_block(firpfbch_cccf _q, liquid_float_complex *_x, unsigned int _n_frames,
                                             liquid_float_complex *_y)
{
    if (_q->type != LIQUID_SYNTHESIZER) lc::fatal("firpfbch_cccf_synthesizer_execute_block", "object is not a synthesizer");
    pfbch_cccf_host_block(_q, _x, _n_frames, _y);
}

void firpfbch_cccf_synthesizer_execute(firpfbch_cccf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    firpfbch_cccf_synthesizer_execute_block(_q, _x, 1, _y);
}

void firpfbch_cccf_analyzer_execute(firpfbch_cccf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    firpfbch_cccf_analyzer_execute_block(_q, _x, 1, _y);
}

void firpfbch_cccf_analyzer_push(firpfbch_cccf _q, liquid_float_complex _x)
{
    if (_q->type != LIQUID_ANALYZER) lc::fatal("firpfbch_cccf_analyzer_push", "object is not an analyzer");
    if (_q->real) {
        firpfbch_crcf_analyzer_push(_q->real, _x);
        return;
    }
    float2 *d_one = (float2 *)_q->one.get(sizeof(float2));
    LC_CUDA_CHECK(cudaMemcpyAsync(d_one, &_x, sizeof(float2), cudaMemcpyHostToDevice, _q->stream));
    lc::roll_history(_q->d_hist, _q->H, _q->hist_tmp, d_one, 1, _q->stream);
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    _q->phase = (_q->phase + 1) % _q->M;
}

void firpfbch_cccf_analyzer_run(firpfbch_cccf _q, unsigned int _k, liquid_float_complex *_X)
{
    if (_q->type != LIQUID_ANALYZER) lc::fatal("firpfbch_cccf_analyzer_run", "object is not an analyzer");
    if (_q->real) {
        firpfbch_crcf_analyzer_run(_q->real, _k, _X);
        return;
    }
    float2 *d_y = (float2 *)_q->stage_out.get(sizeof(float2) * _q->M);
    // no new samples: one run on the state as it stands (first = 0), alignment index _k
    analyze_generic_cc(_q, _q->d_hist /*unused*/, d_y, 1, 0, _k % _q->M, _q->stream);
    LC_CUDA_CHECK(cudaMemcpyAsync(_X, d_y, sizeof(float2) * _q->M, cudaMemcpyDeviceToHost, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

}  // extern "C"
