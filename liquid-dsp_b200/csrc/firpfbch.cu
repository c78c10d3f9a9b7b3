// firpfbch.cu -- firpfbch_crcf: critically sampled M-channel polyphase analysis / synthesis channelizer.
//
// Replaces src/multichannel/src/firpfbch.c (create :73-142, create_kaiser :150-184, create_rnyquist
// :193-256, reset :284-293, synthesizer_execute :314-336, analyzer_execute :346-359, analyzer_push
// :369-377, analyzer_run :383-409).  The per-branch window objects become a flat history of the last
// p*M pushed samples plus a push counter (the reference's filter_index is M-1-(count mod M)).
//
// Analyzer:  push c goes to window M-1-(c mod M); a run reads, for branch i, window (i+k)%M and puts
//            sum_n h[i+nM] * (n-th newest of that window) into input M-1-i of an unscaled forward DFT.
// Synthesizer: v_s = unscaled backward DFT of frame s; y[sM+i] = sum_n h[i+nM] v_{s-n}[i].
#include <cmath>
#include <cstdint>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "pfb_kernels.cuh"
#include "host_pipeline.cuh"
#include "pfbch_fused.cuh"

struct firpfbch_crcf_s {
    int type;
    unsigned int M, p;
    unsigned int H;                 // state length: p*M samples (analyzer: inputs; synthesizer: IDFT rows)
    std::vector<float> h;           // M*p prototype taps (host copy, for print)
    float *d_taps = nullptr;        // [p][M]: taps[n][i] = h[i + nM]
    float2 *d_hist = nullptr;
    float2 *d_zero = nullptr;       // H zeros: the reset state, for the stateless multi-stream form
    float2 *d_tw_fused = nullptr;   // fused M = 256 kernel's inter-pass twiddles
    bool fused = false;
    unsigned int phase = 0;         // analyzer: pushes since reset, mod M
    lc::FftPlanDev fft;
    lc::DevBuf scratch, stage_in, stage_out, hist_tmp, one;
    cudaStream_t stream = nullptr;
};

namespace {

// shape-generic analyzer step: n runs of (M pushes, run(k)) -- one kernel when M is a power of two that
// fits shared memory, else polyphase MAC into y followed by an in-place forward DFT; unscaled
void analyze_generic(firpfbch_crcf q, const float2 *d_hist, const float2 *d_x, float2 *d_y, unsigned long long n,
                     unsigned int phase, unsigned int first, unsigned int k, cudaStream_t st)
{
    // (see firpfbch2.cu: column-walking MAC + tile / quad transform for large aligned blocks)
    const bool two_fast = phase == 0 && k == 0 && first >= q->M && first % q->M == 0 && q->M % 64 == 0 && n >= 64 &&
                          q->fft.d_tw_tile != nullptr && !getenv("LIQUID_CUDA_ONE_KERNEL") && !getenv("LIQUID_CUDA_NO_FUSED") &&
                          !getenv("LIQUID_CUDA_MAC_PLAIN");
    if (lc::pfb::analyze_pow2_supported(q->M) && !getenv("LIQUID_CUDA_TWO_KERNEL") && !two_fast) {
        lc::pfb::launch_pfbch_analyze_pow2(d_hist, d_x, (int)q->H, q->d_taps, q->fft.d_twp, d_y, q->M, q->p, n, phase, first, k, st);
        return;
    }
    lc::pfb::launch_pfbch_mac(d_hist, d_x, (int)q->H, q->d_taps, d_y, q->M, q->p, n, phase, first, k, st);
    q->fft.run(d_y, d_y, n, st);
}

}  // namespace

extern "C" {

firpfbch_crcf firpfbch_crcf_create(int _type, unsigned int _M, unsigned int _p, float *_h)
{
    if (_type != LIQUID_ANALYZER && _type != LIQUID_SYNTHESIZER) {
        fprintf(stderr, "error: firpfbch_%s_create(), invalid type %d\n", "crcf", _type);
        exit(1);
    } else if (_M == 0) {
        lc::fatal("firpfbch_crcf_create", "number of channels must be greater than 0");
    } else if (_p == 0) {
        lc::fatal("firpfbch_crcf_create", "invalid filter size (must be greater than 0)");
    }
    lc::ensure_device();
    firpfbch_crcf q = new firpfbch_crcf_s();
    q->type = _type;
    q->M = _M;
    q->p = _p;
    q->H = _p * _M;
    q->h.assign(_h, _h + (size_t)_M * _p);
    // h is already laid out as taps[n][i] = h[i + nM]  (firpfbch.c:115-124 sub-samples it the same way)
    LC_CUDA_CHECK(cudaMalloc(&q->d_taps, sizeof(float) * q->h.size()));
    LC_CUDA_CHECK(cudaMemcpy(q->d_taps, q->h.data(), sizeof(float) * q->h.size(), cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMalloc(&q->d_hist, sizeof(float2) * q->H));
    LC_CUDA_CHECK(cudaMalloc(&q->d_zero, sizeof(float2) * q->H));
    LC_CUDA_CHECK(cudaMemset(q->d_zero, 0, sizeof(float2) * q->H));
    if ((_type == LIQUID_ANALYZER && lc::pfb::fused_pfbch_supported(_M, _p)) ||
        (_type == LIQUID_SYNTHESIZER && lc::pfb::fused_pfbch_synth_supported(_M, _p))) {
        std::vector<float2> tw;
        lc::pfb::fused_pfbch_make_twiddles(tw, _type == LIQUID_ANALYZER ? -1 : +1, _M, _p);
        LC_CUDA_CHECK(cudaMalloc(&q->d_tw_fused, tw.size() * sizeof(float2)));
        LC_CUDA_CHECK(cudaMemcpy(q->d_tw_fused, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
        q->fused = true;
    }
    q->fft.init(_M, _type == LIQUID_ANALYZER ? -1 : +1);     // firpfbch.c:131-135
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    firpfbch_crcf_reset(q);
    return q;
}

firpfbch_crcf firpfbch_crcf_create_kaiser(int _type, unsigned int _M, unsigned int _m, float _As)
{
    if (_M == 0) lc::fatal("firpfbch_crcf_create_kaiser", "number of channels must be greater than 0");
    if (_m == 0) lc::fatal("firpfbch_crcf_create_kaiser", "invalid filter size (must be greater than 0)");
    _As = fabsf(_As);
    // 2Mm+1 Kaiser taps at fc = 0.5/M, no normalisation; the first 2Mm are used  (firpfbch.c:164-180)
    unsigned int h_len = 2 * _M * _m + 1;
    std::vector<float> h(h_len);
    float fc = 0.5f / (float)_M;
    liquid_firdes_kaiser(h_len, fc, _As, 0.0f, h.data());
    return firpfbch_crcf_create(_type, _M, 2 * _m, h.data());
}

firpfbch_crcf firpfbch_crcf_create_rnyquist(int _type, unsigned int _M, unsigned int _m, float _beta, int _ftype)
{
    if (_type != LIQUID_ANALYZER && _type != LIQUID_SYNTHESIZER) {
        fprintf(stderr, "error: firpfbch_%s_create_rnyquist(), invalid type %d\n", "crcf", _type);
        exit(1);
    } else if (_M == 0) {
        lc::fatal("firpfbch_crcf_create_rnyquist", "number of channels must be greater than 0");
    } else if (_m == 0) {
        lc::fatal("firpfbch_crcf_create_rnyquist", "invalid filter size (must be greater than 0)");
    }
    unsigned int h_len = 2 * _M * _m + 1;
    std::vector<float> h(h_len);
    switch (_ftype) {                                   // firpfbch.c:216-235
    case LIQUID_FIRFILT_ARKAISER:
    case LIQUID_FIRFILT_RKAISER:
    case LIQUID_FIRFILT_RRC:
    case LIQUID_FIRFILT_hM3:
        liquid_firdes_prototype(_ftype, _M, _m, _beta, 0.0f, h.data());
        break;
    default:
        fprintf(stderr, "error: firpfbch_%s_create_rnyquist(), unknown/invalid prototype (%d)\n", "crcf", _ftype);
        exit(1);
    }
    // first 2Mm taps; matched filter g(-t) for the analyzer  (firpfbch.c:238-248)
    unsigned int g_len = 2 * _M * _m;
    std::vector<float> g(g_len);
    for (unsigned int i = 0; i < g_len; i++) g[i] = (_type == LIQUID_SYNTHESIZER) ? h[i] : h[g_len - i - 1];
    return firpfbch_crcf_create(_type, _M, 2 * _m, g.data());
}

void firpfbch_crcf_destroy(firpfbch_crcf _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    cudaFree(_q->d_taps);
    cudaFree(_q->d_hist);
    cudaFree(_q->d_zero);
    cudaFree(_q->d_tw_fused);
    _q->fft.destroy();
    _q->scratch.release();
    _q->stage_in.release();
    _q->stage_out.release();
    _q->hist_tmp.release();
    _q->one.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void firpfbch_crcf_reset(firpfbch_crcf _q)
{
    LC_CUDA_CHECK(cudaMemsetAsync(_q->d_hist, 0, sizeof(float2) * _q->H, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    _q->phase = 0;          // filter_index = M-1 (firpfbch.c:292)
}

void firpfbch_crcf_print(firpfbch_crcf _q)
{
    printf("firpfbch (%s) [%u channels]:\n", _q->type == LIQUID_ANALYZER ? "analyzer" : "synthesizer", _q->M);
    for (unsigned int i = 0; i < _q->M * _q->p; i++)
        printf("  h[%3u] = %12.8f + %12.8f*j\n", i, _q->h[i], 0.0f);
}

unsigned int firpfbch_crcf_get_state_len(firpfbch_crcf _q) { return _q->H; }

void firpfbch_crcf_get_state(firpfbch_crcf _q, liquid_float_complex *_hist)
{
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    LC_CUDA_CHECK(cudaMemcpy(_hist, _q->d_hist, sizeof(float2) * _q->H, cudaMemcpyDeviceToHost));
}

void firpfbch_crcf_set_state(firpfbch_crcf _q, const liquid_float_complex *_hist)
{
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    LC_CUDA_CHECK(cudaMemcpy(_q->d_hist, _hist, sizeof(float2) * _q->H, cudaMemcpyHostToDevice));
}

void firpfbch_crcf_execute_block_dev(firpfbch_crcf _q, const void *_d_x, unsigned long long _n_frames,
                                     void *_d_y, void *_stream)
{
    if (_n_frames == 0) return;
    cudaStream_t st = lc::as_stream(_stream);
    const float2 *d_x = (const float2 *)_d_x;
    float2 *d_y = (float2 *)_d_y;
    if (_q->type == LIQUID_ANALYZER) {
        // M pushes then run(k=0) per frame
        unsigned long long done = 0;
        if (_q->fused && _q->phase == 0 && _n_frames >= 16 && ((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_y % 16) == 0 &&
            !getenv("LIQUID_CUDA_NO_FUSED"))
            done = (unsigned long long)lc::pfb::launch_pfbch_fused(_q->d_hist, d_x, (int)_q->H, _q->d_taps, _q->d_tw_fused,
                                                                   d_y, _q->M, _q->p, _n_frames, 1, 0, st);
        if (done < _n_frames) {
            // shape-generic path (and the < 16 frames the fused kernel leaves)
            analyze_generic(_q, _q->d_hist, d_x, d_y + done * _q->M, _n_frames - done, _q->phase,
                            (unsigned int)((done + 1) * _q->M), 0, st);
        }
        lc::roll_history(_q->d_hist, _q->H, _q->hist_tmp, d_x, _n_frames * _q->M, st);
    } else {
        const unsigned int M = _q->M, p = _q->p;
        unsigned long long done = 0;
        if (_q->fused && _n_frames >= 16 && ((uintptr_t)d_x % 16) == 0 && !getenv("LIQUID_CUDA_NO_FUSED"))
            done = (unsigned long long)lc::pfb::launch_pfbch_synth_fused(_q->d_hist, (int)p, d_x, _q->d_taps, _q->d_tw_fused,
                                                                         d_y, M, p, _n_frames, 1, 0, st);
        // rows of v the generic tail and the new state need: frames [v0, n)
        const unsigned long long v0 = (done > p) ? done - p : 0;
        float2 *d_v = (float2 *)_q->scratch.get(sizeof(float2) * (_n_frames - v0 + p) * M);
        if (done < _n_frames || v0 < _n_frames) _q->fft.run(d_x + v0 * M, d_v, _n_frames - v0, st);
        if (done < _n_frames) {
            // shape-generic path (and the < 16 frames the fused kernel leaves)
            const float2 *hist_rows = (done == 0) ? _q->d_hist : d_v + (done - v0 - p) * M;   // p rows before frame `done`
            lc::pfb::launch_pfbch_synth_fir(hist_rows, d_v + (done - v0) * M, (int)p, _q->d_taps, d_y + done * M, M, p,
                                            _n_frames - done, st);
        }
        // new state: the last p rows of v (d_v holds frames [v0, n))
        lc::roll_history(_q->d_hist, _q->H, _q->hist_tmp, d_v, (_n_frames - v0) * M, st);
    }
}

// n_streams independent streams of n_frames frames each, every stream from the reset state; the
// object's own state is not touched
void firpfbch_crcf_execute_streams_dev(firpfbch_crcf _q, const void *_d_x, unsigned int _n_streams,
                                       unsigned long long _n_frames, void *_d_y, void *_stream)
{
    if (_n_frames == 0 || _n_streams == 0) return;
    cudaStream_t st = lc::as_stream(_stream);
    const float2 *d_x = (const float2 *)_d_x;
    float2 *d_y = (float2 *)_d_y;
    const unsigned int M = _q->M, p = _q->p;
    const long long stride = (long long)_n_frames * M;
    const bool ok = _q->fused && _n_frames >= 16 && ((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_y % 16) == 0 &&
                    (stride % 2) == 0 && !getenv("LIQUID_CUDA_NO_FUSED");
    unsigned long long done = 0;
    if (_q->type == LIQUID_ANALYZER) {
        if (ok)
            done = (unsigned long long)lc::pfb::launch_pfbch_fused(nullptr, d_x, 0, _q->d_taps, _q->d_tw_fused, d_y, M, p,
                                                                   _n_frames, _n_streams, stride, st);
        if (done < _n_frames) {
            for (unsigned int s = 0; s < _n_streams; s++) {
                float2 *ys = d_y + s * stride + done * M;
                analyze_generic(_q, _q->d_zero, d_x + s * stride, ys, _n_frames - done, 0, (unsigned int)((done + 1) * M), 0, st);
            }
        }
    } else {
        if (ok)
            done = (unsigned long long)lc::pfb::launch_pfbch_synth_fused(nullptr, 0, d_x, _q->d_taps, _q->d_tw_fused, d_y, M, p,
                                                                         _n_frames, _n_streams, stride, st);
        if (done < _n_frames) {
            const unsigned long long v0 = (done > p) ? done - p : 0;
            float2 *d_v = (float2 *)_q->scratch.get(sizeof(float2) * (_n_frames - v0 + p) * M);
            for (unsigned int s = 0; s < _n_streams; s++) {
                _q->fft.run(d_x + s * stride + v0 * M, d_v, _n_frames - v0, st);
                const float2 *hist_rows = (done == 0) ? _q->d_zero : d_v + (done - v0 - p) * M;
                lc::pfb::launch_pfbch_synth_fir(hist_rows, d_v + (done - v0) * M, (int)p, _q->d_taps,
                                                d_y + s * stride + done * M, M, p, _n_frames - done, st);
            }
        }
    }
}

static void pfbch_host_block(firpfbch_crcf _q, liquid_float_complex *_x, unsigned int _n_frames, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x),
                   reinterpret_cast<lc_cf *>(_y), _n_frames, _q->M, _q->M,
                   [&](float2 *di, size_t nf, float2 *dob, cudaStream_t st) {
                       firpfbch_crcf_execute_block_dev(_q, di, nf, dob, st);
                   });
}

void firpfbch_crcf_analyzer_execute_block(firpfbch_crcf _q, liquid_float_complex *_x, unsigned int _n_frames,
                                          liquid_float_complex *_y)
{
    if (_q->type != LIQUID_ANALYZER) lc::fatal("firpfbch_crcf_analyzer_execute_block", "object is not an analyzer");
    pfbch_host_block(_q, _x, _n_frames, _y);
}

void firpfbch_crcf_synthesizer_execute_block(firpfbch_crcf _q, liquid_float_complex *_x, unsigned int _n_frames,
                                             liquid_float_complex *_y)
{
    if (_q->type != LIQUID_SYNTHESIZER) lc::fatal("firpfbch_crcf_synthesizer_execute_block", "object is not a synthesizer");
    pfbch_host_block(_q, _x, _n_frames, _y);
}

// The reference does not check the object's type in these two (firpfbch.c:314, :346); calling the
// wrong one there runs the other algorithm on a plan of the wrong direction.  Here it is an error.
void firpfbch_crcf_synthesizer_execute(firpfbch_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    firpfbch_crcf_synthesizer_execute_block(_q, _x, 1, _y);
}

void firpfbch_crcf_analyzer_execute(firpfbch_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    firpfbch_crcf_analyzer_execute_block(_q, _x, 1, _y);
}

void firpfbch_crcf_analyzer_push(firpfbch_crcf _q, liquid_float_complex _x)
{
    if (_q->type != LIQUID_ANALYZER) lc::fatal("firpfbch_crcf_analyzer_push", "object is not an analyzer");
    float2 *d_one = (float2 *)_q->one.get(sizeof(float2));
    LC_CUDA_CHECK(cudaMemcpyAsync(d_one, &_x, sizeof(float2), cudaMemcpyHostToDevice, _q->stream));
    lc::roll_history(_q->d_hist, _q->H, _q->hist_tmp, d_one, 1, _q->stream);
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    _q->phase = (_q->phase + 1) % _q->M;
}

void firpfbch_crcf_analyzer_run(firpfbch_crcf _q, unsigned int _k, liquid_float_complex *_X)
{
    if (_q->type != LIQUID_ANALYZER) lc::fatal("firpfbch_crcf_analyzer_run", "object is not an analyzer");
    float2 *d_y = (float2 *)_q->stage_out.get(sizeof(float2) * _q->M);
    // no new samples: one run on the state as it stands (first = 0), alignment index _k
    analyze_generic(_q, _q->d_hist, _q->d_hist /*unused*/, d_y, 1, _q->phase, 0, _k % _q->M, _q->stream);
    LC_CUDA_CHECK(cudaMemcpyAsync(_X, d_y, sizeof(float2) * _q->M, cudaMemcpyDeviceToHost, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

}  // extern "C"
