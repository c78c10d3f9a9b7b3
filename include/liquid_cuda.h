/*
 * liquid_cuda.h -- C ABI of libliquid_cuda.so: the B200 (sm_100a) implementation of liquid-dsp's
 * polyphase filterbank channelizer hot path.
 *
 * Part 1 re-declares, with the reference's exact names and signatures, the slice of
 * include/liquid.h (reference, /root/reference) that this library replaces; each block cites the
 * reference declaration (liquid.h:line) and the reference definition it stands in for.
 * A program that already includes <liquid.h> keeps doing so and simply links -lliquid_cuda ahead of
 * -lliquid: Part 1 is skipped (guarded by __LIQUID_H__) and only Part 2 is added.
 *
 * Part 2 are the batched / device-pointer / state entry points a GPU needs (the reference's
 * one-frame-per-call execute is kept for drop-in use but moves 4 KB per call).
 *
 * Conventions kept from the reference: create* allocates and returns an opaque handle, copies the
 * taps (caller keeps its array); destroy frees everything; invalid arguments print
 * "error: <fn>(), <msg>" to stderr and exit(1); execute functions return void; objects are not
 * thread-safe, distinct objects are independent.  CUDA failures print the CUDA error and abort().
 * There is no CPU fallback anywhere in this library: without a CUDA device every create aborts.
 *
 * All pointers are plain host pointers unless the name ends in _dev (device pointers, asynchronous
 * on the given cudaStream_t passed as void*; NULL = the legacy default stream).
 *
 * Stream contract: a *_dev call updates the object's state (history, frame parity) in stream order on the
 * CALLER's stream.  Every other entry point of the same object (execute, execute_block, reset, get_state,
 * set_state, push ...) works on the object's own private stream and synchronises only that one.  Between a
 * *_dev call on stream S and any non-_dev call on the same object the caller must synchronise S
 * (cudaStreamSynchronize / liquid_cuda_sync); consecutive *_dev calls need nothing as long as they are issued
 * on the same stream (or on streams the caller has ordered with events).  Distinct objects never interact.
 *
 * Device selection: liquid_cuda_set_device(d) (or env LIQUID_CUDA_DEVICE when it is never called) selects the
 * device for every thread of the process; each thread is switched to it on its first call into the library.
 */
#ifndef LIQUID_CUDA_H
#define LIQUID_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

/* ============================================================================================
 * Part 1 -- the reference API surface (skipped when <liquid.h> is already included)
 * ============================================================================================ */
#ifndef __LIQUID_H__

/* complex float: interleaved {re, im}, layout- and ABI-compatible with C99 `float complex`
 * (liquid.h:77-88) */
#if defined(__STDC_VERSION__) && !defined(__cplusplus) && !defined(LIQUID_CUDA_NO_COMPLEX_H)
#  include <complex.h>
typedef float _Complex liquid_float_complex;
#else
typedef struct { float real; float imag; } liquid_float_complex;
#endif

#define LIQUID_ANALYZER     0       /* liquid.h:5651 */
#define LIQUID_SYNTHESIZER  1       /* liquid.h:5652 */
#define LIQUID_FFT_FORWARD  (+1)    /* liquid.h:1122 */
#define LIQUID_FFT_BACKWARD (-1)    /* liquid.h:1123 */
/* enum liquid_firfilt_type, liquid.h:1414-1435 (values only; prototypes marked * are designed here) */
#define LIQUID_FIRFILT_UNKNOWN   0
#define LIQUID_FIRFILT_KAISER    1  /* * */
#define LIQUID_FIRFILT_PM        2  /* * */
#define LIQUID_FIRFILT_RCOS      3  /* * */
#define LIQUID_FIRFILT_FEXP      4  /* * */
#define LIQUID_FIRFILT_FSECH     5  /* * */
#define LIQUID_FIRFILT_FARCSECH  6  /* * */
#define LIQUID_FIRFILT_ARKAISER  7  /* * */
#define LIQUID_FIRFILT_RKAISER   8  /* * */
#define LIQUID_FIRFILT_RRC       9  /* * */
#define LIQUID_FIRFILT_hM3       10 /* * */
#define LIQUID_FIRFILT_GMSKTX    11 /* * */
#define LIQUID_FIRFILT_GMSKRX    12 /* * */
#define LIQUID_FIRFILT_RFEXP     13 /* * */
#define LIQUID_FIRFILT_RFSECH    14 /* * */
#define LIQUID_FIRFILT_RFARCSECH 15 /* * */

/* ---- prototype design: liquid.h:1476, :1548, :1580 -> src/filter/src/firdes.c:224-280, rrcos.c:37-95 */
float kaiser_beta_As(float _As);
void  liquid_firdes_kaiser(unsigned int _n, float _fc, float _As, float _mu, float *_h);
void  liquid_firdes_rrcos(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
/* root-Nyquist Kaiser prototypes and their helpers: liquid.h:1463, :1583, :1586, :1664, :1444
 * -> src/filter/src/rkaiser.c:47-148, :339-516, firdes.c:80-116, :290-370, :503-531 */
float estimate_req_filter_As(float _df, unsigned int _N);
void  liquid_filter_isi(float *_h, unsigned int _k, unsigned int _m, float *_rms, float *_max);
void  liquid_firdes_arkaiser(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
void  liquid_firdes_rkaiser(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
void  liquid_firdes_prototype(int _type, unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
/* Parks-McClellan equiripple design (band-pass / even-symmetric filters) and the harris-Moerder root-Nyquist
 * prototype built on it: liquid.h:1482-1512, :1592 -> src/filter/src/firdespm.c:115-748, hM3.c:49-126 (host code) */
#define LIQUID_FIRDESPM_BANDPASS        0   /* liquid_firdespm_btype */
#define LIQUID_FIRDESPM_DIFFERENTIATOR  1
#define LIQUID_FIRDESPM_HILBERT         2
#define LIQUID_FIRDESPM_FLATWEIGHT      0   /* liquid_firdespm_wtype */
#define LIQUID_FIRDESPM_EXPWEIGHT       1
#define LIQUID_FIRDESPM_LINWEIGHT       2
void  firdespm_run(unsigned int _h_len, unsigned int _num_bands, float *_bands, float *_des, float *_weights,
                   int *_wtype, int _btype, float *_h);
typedef struct firdespm_s * firdespm;            /* liquid.h:1514-1540 -> firdespm.c:142-343 */
firdespm firdespm_create(unsigned int _h_len, unsigned int _num_bands, float *_bands, float *_des, float *_weights,
                         int *_wtype, int _btype);
void  firdespm_destroy(firdespm _q);
void  firdespm_print(firdespm _q);
void  firdespm_execute(firdespm _q, float *_h);
void  liquid_firdes_hM3(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
/* the remaining families of liquid_firdes_prototype (host code): raised cosine, flipped exponential / sech / arcsech
 * Nyquist and root-Nyquist filters, GMSK transmit / receive: liquid.h:1573-1605 -> src/filter/src/rcos.c:39-80,
 * fnyquist.c:47-318, gmsk.c:39-193 */
void  liquid_firdes_rcos(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
void  liquid_firdes_fexp(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
void  liquid_firdes_rfexp(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
void  liquid_firdes_fsech(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
void  liquid_firdes_rfsech(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
void  liquid_firdes_farcsech(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
void  liquid_firdes_rfarcsech(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
void  liquid_firdes_gmsktx(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);
void  liquid_firdes_gmskrx(unsigned int _k, unsigned int _m, float _beta, float _dt, float *_h);

/* ---- firdecim_crcf / firinterp_crcf: liquid.h:2665-2722, :2498-2556 -> src/filter/src/firdecim.c:46-223,
 *      firinterp.c:43-217 (instantiated by filter_crcf.c).  SURVEY.md 8(f) item 3. */
typedef struct firdecim_crcf_s * firdecim_crcf;
firdecim_crcf firdecim_crcf_create(unsigned int _M, float *_h, unsigned int _h_len);
firdecim_crcf firdecim_crcf_create_kaiser(unsigned int _M, unsigned int _m, float _As);
firdecim_crcf firdecim_crcf_create_prototype(int _type, unsigned int _M, unsigned int _m, float _beta, float _dt);
void firdecim_crcf_destroy(firdecim_crcf _q);
void firdecim_crcf_print(firdecim_crcf _q);
void firdecim_crcf_clear(firdecim_crcf _q);
void firdecim_crcf_execute(firdecim_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y);
void firdecim_crcf_execute_block(firdecim_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);
typedef struct firinterp_crcf_s * firinterp_crcf;
firinterp_crcf firinterp_crcf_create(unsigned int _M, float *_h, unsigned int _h_len);
firinterp_crcf firinterp_crcf_create_kaiser(unsigned int _M, unsigned int _m, float _As);
firinterp_crcf firinterp_crcf_create_prototype(int _type, unsigned int _k, unsigned int _m, float _beta, float _dt);
void firinterp_crcf_destroy(firinterp_crcf _q);
void firinterp_crcf_print(firinterp_crcf _q);
void firinterp_crcf_reset(firinterp_crcf _q);
void firinterp_crcf_execute(firinterp_crcf _q, liquid_float_complex _x, liquid_float_complex *_y);
void firinterp_crcf_execute_block(firinterp_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);

/* ---- resamp2_crcf (half-band filter / 2:1 resamplers): liquid.h:2840-2925 -> src/filter/src/resamp2.c:65-359;
 *      msresamp2_crcf (2^S cascades): liquid.h:3022-3080 -> src/filter/src/msresamp2.c:82-354;
 *      estimate_req_filter_len: liquid.h:1500 -> src/filter/src/firdes.c:52-75.  SURVEY.md 8(f) item 3. */
#define LIQUID_RESAMP_INTERP 0      /* enum liquid_resamp_type, liquid.h:3022-3025 */
#define LIQUID_RESAMP_DECIM  1
unsigned int estimate_req_filter_len(float _df, float _As);
typedef struct resamp2_crcf_s * resamp2_crcf;
resamp2_crcf resamp2_crcf_create(unsigned int _m, float _f0, float _As);
resamp2_crcf resamp2_crcf_recreate(resamp2_crcf _q, unsigned int _m, float _f0, float _As);
void resamp2_crcf_destroy(resamp2_crcf _q);
void resamp2_crcf_print(resamp2_crcf _q);
void resamp2_crcf_clear(resamp2_crcf _q);
unsigned int resamp2_crcf_get_delay(resamp2_crcf _q);
void resamp2_crcf_filter_execute(resamp2_crcf _q, liquid_float_complex _x, liquid_float_complex *_y0, liquid_float_complex *_y1);
void resamp2_crcf_analyzer_execute(resamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y);
void resamp2_crcf_synthesizer_execute(resamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y);
void resamp2_crcf_decim_execute(resamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y);
void resamp2_crcf_interp_execute(resamp2_crcf _q, liquid_float_complex _x, liquid_float_complex *_y);
typedef struct msresamp2_crcf_s * msresamp2_crcf;
msresamp2_crcf msresamp2_crcf_create(int _type, unsigned int _num_stages, float _fc, float _f0, float _As);
void msresamp2_crcf_destroy(msresamp2_crcf _q);
void msresamp2_crcf_print(msresamp2_crcf _q);
void msresamp2_crcf_reset(msresamp2_crcf _q);
float msresamp2_crcf_get_delay(msresamp2_crcf _q);
void msresamp2_crcf_execute(msresamp2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y);

/* ---- ofdmframegen (OFDM symbol generator) and the subcarrier-allocation helpers: liquid.h:5803-5876 ->
 *      src/multichannel/src/ofdmframegen.c:85-361, ofdmframe.common.c:176-282.  SURVEY.md 8(f) item 4. */
#define OFDMFRAME_SCTYPE_NULL   0
#define OFDMFRAME_SCTYPE_PILOT  1
#define OFDMFRAME_SCTYPE_DATA   2
void ofdmframe_init_default_sctype(unsigned int _M, unsigned char *_p);
void ofdmframe_validate_sctype(unsigned char *_p, unsigned int _M, unsigned int *_M_null, unsigned int *_M_pilot,
                               unsigned int *_M_data);
void ofdmframe_print_sctype(unsigned char *_p, unsigned int _M);
typedef struct ofdmframegen_s * ofdmframegen;
ofdmframegen ofdmframegen_create(unsigned int _M, unsigned int _cp_len, unsigned int _taper_len, unsigned char *_p);
void ofdmframegen_destroy(ofdmframegen _q);
void ofdmframegen_print(ofdmframegen _q);
void ofdmframegen_reset(ofdmframegen _q);
void ofdmframegen_write_S0a(ofdmframegen _q, liquid_float_complex *_y);
void ofdmframegen_write_S0b(ofdmframegen _q, liquid_float_complex *_y);
void ofdmframegen_write_S1(ofdmframegen _q, liquid_float_complex *_y);
void ofdmframegen_writesymbol(ofdmframegen _q, liquid_float_complex *_x, liquid_float_complex *_y);
void ofdmframegen_writetail(ofdmframegen _q, liquid_float_complex *_x);

/* ---- ofdmframesync: liquid.h:5878-5911 -> src/multichannel/src/ofdmframesync.c.  Acquisition (the PLCP search and the
 *      equalizer fit) runs on the host as in the reference; the payload state -- NCO mix, M-point transform, equalizer
 *      product, pilot phase fit, de-rotation, NCO feedback, for every symbol left in an execute() call -- is one
 *      device launch, and the callbacks are made in order afterwards (csrc/ofdmframesync.cu). */
typedef int (*ofdmframesync_callback)(liquid_float_complex *_y, unsigned char *_p, unsigned int _M, void *_userdata);
typedef struct ofdmframesync_s * ofdmframesync;
ofdmframesync ofdmframesync_create(unsigned int _M, unsigned int _cp_len, unsigned int _taper_len, unsigned char *_p,
                                   ofdmframesync_callback _callback, void *_userdata);
void ofdmframesync_destroy(ofdmframesync _q);
void ofdmframesync_print(ofdmframesync _q);
void ofdmframesync_reset(ofdmframesync _q);
void ofdmframesync_execute(ofdmframesync _q, liquid_float_complex *_x, unsigned int _n);
float ofdmframesync_get_rssi(ofdmframesync _q);
float ofdmframesync_get_cfo(ofdmframesync _q);
void ofdmframesync_debug_enable(ofdmframesync _q);
void ofdmframesync_debug_disable(ofdmframesync _q);
void ofdmframesync_debug_print(ofdmframesync _q, const char *_filename);

/* ---- fftfilt_crcf (FIR filtering through 2n-point transforms): liquid.h:2192-2232 -> src/filter/src/fftfilt.c:68-266.
 *      SURVEY.md 8(f) item 3.  x and y must not overlap in the device-pointer form. */
typedef struct fftfilt_crcf_s * fftfilt_crcf;
fftfilt_crcf fftfilt_crcf_create(float *_h, unsigned int _h_len, unsigned int _n);
void fftfilt_crcf_destroy(fftfilt_crcf _q);
void fftfilt_crcf_reset(fftfilt_crcf _q);
void fftfilt_crcf_print(fftfilt_crcf _q);
void fftfilt_crcf_set_scale(fftfilt_crcf _q, float _scale);
void fftfilt_crcf_execute(fftfilt_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y);
unsigned int fftfilt_crcf_get_length(fftfilt_crcf _q);

/* ---- spgramcf (windowed-FFT spectral periodogram): liquid.h:1230-1316 -> src/fft/src/spgramcf.c -> spgram.c:57-398.
 *      SURVEY.md 8(f) item 2. */
typedef struct spgramcf_s * spgramcf;
spgramcf spgramcf_create(unsigned int _nfft, float *_window, unsigned int _window_len);
spgramcf spgramcf_create_kaiser(unsigned int _nfft, unsigned int _window_len, float _beta);
spgramcf spgramcf_create_default(unsigned int _nfft);
void spgramcf_destroy(spgramcf _q);
void spgramcf_reset(spgramcf _q);
void spgramcf_push(spgramcf _q, liquid_float_complex _x);
void spgramcf_write(spgramcf _q, liquid_float_complex *_x, unsigned int _n);
void spgramcf_execute(spgramcf _q, liquid_float_complex *_X);
void spgramcf_execute_psd(spgramcf _q, float *_X);
void spgramcf_accumulate_psd(spgramcf _q, liquid_float_complex *_x, float _alpha, unsigned int _n);
void spgramcf_write_accumulation(spgramcf _q, float *_x);
void spgramcf_estimate_psd(spgramcf _q, liquid_float_complex *_x, unsigned int _n, float *_psd);

/* ---- asgramcf (ASCII spectrogram over spgramcf): liquid.h:1333-1395 -> src/fft/src/asgram.c:48-213.
 *      _ascii holds nfft characters (no terminator is written, as in the reference). */
typedef struct asgramcf_s * asgramcf;
asgramcf asgramcf_create(unsigned int _nfft);
void asgramcf_destroy(asgramcf _q);
void asgramcf_reset(asgramcf _q);
void asgramcf_set_scale(asgramcf _q, float _offset, float _scale);
void asgramcf_push(asgramcf _q, liquid_float_complex _x);
void asgramcf_write(asgramcf _q, liquid_float_complex *_x, unsigned int _n);
void asgramcf_execute(asgramcf _q, char *_ascii, float *_peakval, float *_peakfreq);
void asgramcf_print(asgramcf _q);

/* ---- firpfbch2_crcf: liquid.h:5754-5793 -> src/multichannel/src/firpfbch2.c:66-358 */
typedef struct firpfbch2_crcf_s * firpfbch2_crcf;
firpfbch2_crcf firpfbch2_crcf_create(int _type, unsigned int _M, unsigned int _m, float *_h);
firpfbch2_crcf firpfbch2_crcf_create_kaiser(int _type, unsigned int _M, unsigned int _m, float _As);
void firpfbch2_crcf_destroy(firpfbch2_crcf _q);
void firpfbch2_crcf_reset(firpfbch2_crcf _q);
void firpfbch2_crcf_print(firpfbch2_crcf _q);
/* analyzer: _x[M/2] -> _y[M]; synthesizer: _x[M] -> _y[M/2] */
void firpfbch2_crcf_execute(firpfbch2_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y);

/* ---- firpfbch_crcf: liquid.h:5667-5728 -> src/multichannel/src/firpfbch.c:73-409 */
typedef struct firpfbch_crcf_s * firpfbch_crcf;
firpfbch_crcf firpfbch_crcf_create(int _type, unsigned int _M, unsigned int _p, float *_h);
firpfbch_crcf firpfbch_crcf_create_kaiser(int _type, unsigned int _M, unsigned int _m, float _As);
firpfbch_crcf firpfbch_crcf_create_rnyquist(int _type, unsigned int _M, unsigned int _m, float _beta, int _ftype);
void firpfbch_crcf_destroy(firpfbch_crcf _q);
void firpfbch_crcf_reset(firpfbch_crcf _q);
void firpfbch_crcf_print(firpfbch_crcf _q);
void firpfbch_crcf_synthesizer_execute(firpfbch_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y);
void firpfbch_crcf_analyzer_execute(firpfbch_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y);

/* ---- firpfbch_cccf (complex taps): liquid.h:5660-5664, :5682-5728 -> src/multichannel/src/firpfbch_cccf.c:35-52
 *      instantiating firpfbch.c:73-409 with TC = float complex and dotprod_cccf.  SURVEY.md 8(f) item 1. */
typedef struct firpfbch_cccf_s * firpfbch_cccf;
firpfbch_cccf firpfbch_cccf_create(int _type, unsigned int _M, unsigned int _p, liquid_float_complex *_h);
firpfbch_cccf firpfbch_cccf_create_kaiser(int _type, unsigned int _M, unsigned int _m, float _As);
firpfbch_cccf firpfbch_cccf_create_rnyquist(int _type, unsigned int _M, unsigned int _m, float _beta, int _ftype);
void firpfbch_cccf_destroy(firpfbch_cccf _q);
void firpfbch_cccf_reset(firpfbch_cccf _q);
void firpfbch_cccf_print(firpfbch_cccf _q);
void firpfbch_cccf_synthesizer_execute(firpfbch_cccf _q, liquid_float_complex *_x, liquid_float_complex *_y);
void firpfbch_cccf_analyzer_execute(firpfbch_cccf _q, liquid_float_complex *_x, liquid_float_complex *_y);
/* exported by the reference but absent from liquid.h, as for crcf (firpfbch.c:369, :383) */
void firpfbch_cccf_analyzer_push(firpfbch_cccf _q, liquid_float_complex _x);
void firpfbch_cccf_analyzer_run(firpfbch_cccf _q, unsigned int _k, liquid_float_complex *_X);

/* ---- firfilt_crcf: liquid.h:1985-2074 -> src/filter/src/firfilt.c:62-405 */
typedef struct firfilt_crcf_s * firfilt_crcf;
firfilt_crcf firfilt_crcf_create(float *_h, unsigned int _n);
firfilt_crcf firfilt_crcf_create_kaiser(unsigned int _n, float _fc, float _As, float _mu);
firfilt_crcf firfilt_crcf_create_rnyquist(int _type, unsigned int _k, unsigned int _m, float _beta, float _mu);
firfilt_crcf firfilt_crcf_create_rect(unsigned int _n);
firfilt_crcf firfilt_crcf_recreate(firfilt_crcf _q, float *_h, unsigned int _n);
void firfilt_crcf_destroy(firfilt_crcf _q);
void firfilt_crcf_reset(firfilt_crcf _q);
void firfilt_crcf_print(firfilt_crcf _q);
void firfilt_crcf_set_scale(firfilt_crcf _q, float _scale);
void firfilt_crcf_push(firfilt_crcf _q, liquid_float_complex _x);
void firfilt_crcf_execute(firfilt_crcf _q, liquid_float_complex *_y);
void firfilt_crcf_execute_block(firfilt_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);
unsigned int firfilt_crcf_get_length(firfilt_crcf _q);
void  firfilt_crcf_freqresponse(firfilt_crcf _q, float _fc, liquid_float_complex *_H);
float firfilt_crcf_groupdelay(firfilt_crcf _q, float _fc);

/* ---- firpfb_crcf: liquid.h:2392-2470 -> src/filter/src/firpfb.c:46-346 */
typedef struct firpfb_crcf_s * firpfb_crcf;
firpfb_crcf firpfb_crcf_create(unsigned int _M, float *_h, unsigned int _h_len);
firpfb_crcf firpfb_crcf_create_kaiser(unsigned int _M, unsigned int _m, float _fc, float _As);
firpfb_crcf firpfb_crcf_create_rnyquist(int _type, unsigned int _M, unsigned int _k, unsigned int _m, float _beta);
firpfb_crcf firpfb_crcf_create_drnyquist(int _type, unsigned int _M, unsigned int _k, unsigned int _m, float _beta);   /* liquid.h:2432 */
firpfb_crcf firpfb_crcf_recreate(firpfb_crcf _q, unsigned int _M, float *_h, unsigned int _h_len);
void firpfb_crcf_destroy(firpfb_crcf _q);
void firpfb_crcf_print(firpfb_crcf _q);
void firpfb_crcf_set_scale(firpfb_crcf _q, float _g);
void firpfb_crcf_reset(firpfb_crcf _q);
void firpfb_crcf_push(firpfb_crcf _q, liquid_float_complex _x);
void firpfb_crcf_execute(firpfb_crcf _q, unsigned int _i, liquid_float_complex *_y);

/* ---- dotprod_crcf: liquid.h:503-541 -> src/dotprod/src/dotprod.c:41-168, dotprod_crcf.mmx.c:95-287 */
typedef struct dotprod_crcf_s * dotprod_crcf;
void dotprod_crcf_run(float *_h, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);
void dotprod_crcf_run4(float *_h, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);
dotprod_crcf dotprod_crcf_create(float *_h, unsigned int _n);
dotprod_crcf dotprod_crcf_recreate(dotprod_crcf _q, float *_h, unsigned int _n);
void dotprod_crcf_destroy(dotprod_crcf _q);
void dotprod_crcf_print(dotprod_crcf _q);
void dotprod_crcf_execute(dotprod_crcf _q, liquid_float_complex *_x, liquid_float_complex *_y);

/* ---- complex FFT: liquid.h:1148-1215 -> src/fft/src/fft_common.c:109-316
 * unscaled both ways; FORWARD = e^{-j}, BACKWARD = e^{+j}; the plan borrows _x/_y (not copied) */
typedef struct fftplan_s * fftplan;
fftplan fft_create_plan(unsigned int _n, liquid_float_complex *_x, liquid_float_complex *_y, int _dir, int _flags);
void fft_destroy_plan(fftplan _p);
void fft_print_plan(fftplan _p);
void fft_execute(fftplan _p);
void fft_run(unsigned int _n, liquid_float_complex *_x, liquid_float_complex *_y, int _dir, int _flags);
void fft_shift(liquid_float_complex *_x, unsigned int _n);

#endif /* __LIQUID_H__ */

/* exported by the reference but absent from liquid.h (src/multichannel/src/firpfbch.c:369, :383;
 * used by sandbox/firpfbch_analysis_alignment_test.c) */
void firpfbch_crcf_analyzer_push(firpfbch_crcf _q, liquid_float_complex _x);
void firpfbch_crcf_analyzer_run(firpfbch_crcf _q, unsigned int _k, liquid_float_complex *_X);

/* ============================================================================================
 * Part 2 -- batched / device / state extensions (not in the reference)
 * ============================================================================================ */

/* library / device plumbing */
const char *liquid_cuda_version(void);
int   liquid_cuda_device_count(void);
void  liquid_cuda_set_device(int _device);            /* also honoured: env LIQUID_CUDA_DEVICE */
void *liquid_cuda_malloc(unsigned long long _bytes);  /* device memory */
void  liquid_cuda_free(void *_dptr);
void *liquid_cuda_host_alloc(unsigned long long _bytes);   /* pinned host memory */
void  liquid_cuda_host_free(void *_hptr);
/* page-lock an array the caller allocated itself (malloc, a file mapping) for as long as it is handed to the
 * *_execute_block entry points; returns 0 on success.  Pageable arrays work without it, at the rate of the driver's
 * staged copies (about a fifth of the pinned rate, bench.py e2e.pageable_host_arrays); env LIQUID_CUDA_PAGEABLE=register
 * page-locks arrays of 8 MB or more for the duration of each call instead (slower unless calls are very long). */
int   liquid_cuda_host_register(void *_hptr, unsigned long long _bytes);
void  liquid_cuda_host_unregister(void *_hptr);
void  liquid_cuda_memcpy_h2d(void *_dst, const void *_src, unsigned long long _bytes);
void  liquid_cuda_memcpy_d2h(void *_dst, const void *_src, unsigned long long _bytes);
void  liquid_cuda_sync(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
unsigned long long liquid_cuda_launch_count(void);
/* frames each CTA / cluster walked in the most recent fused channelizer launch (0 = none yet): test aid */
unsigned long long liquid_cuda_last_chunk_frames(void);
/* SMs the persistent one-CTA-per-SM kernels leave free (default 0): set to 1-2 when a collective (NCCL
 * send/recv of the halo) or a side-stream launch should run beside the channelizer kernel, not after it */
void  liquid_cuda_set_reserved_sms(int _n);
int   liquid_cuda_get_reserved_sms(void);
/* fill _x[_n] on the device with the bench's synthetic stream
 * u(n,c) = (splitmix64(seed + 2n + c) >> 40) * 2^-24 - 0.5 ; _first = index of _x[0] in the stream */
void  liquid_cuda_synth_input_dev(void *_d_x, unsigned long long _seed, unsigned long long _first,
                                  unsigned long long _n, void *_stream);

/* ---- firpfbch2_crcf ---- */
/* n_frames consecutive execute() calls in one go; identical results and end state.
 * analyzer: _x[n_frames*M/2] -> _y[n_frames*M]; synthesizer: _x[n_frames*M] -> _y[n_frames*M/2] */
void firpfbch2_crcf_execute_block(firpfbch2_crcf _q, liquid_float_complex *_x,
                                  unsigned int _n_frames, liquid_float_complex *_y);
/* same on device pointers, asynchronous on _stream; state is carried in the object */
void firpfbch2_crcf_execute_block_dev(firpfbch2_crcf _q, const void *_d_x,
                                      unsigned long long _n_frames, void *_d_y, void *_stream);
/* stateless segment form used for time-sharded multi-GPU runs (analyzer only): _d_halo holds the
 * firpfbch2_crcf_get_state_len() input samples that precede _d_x[0] in the stream (zeros at the
 * stream start); _first_frame gives the parity of the first frame.  Object state is not touched. */
void firpfbch2_crcf_execute_segment_dev(firpfbch2_crcf _q, const void *_d_halo, const void *_d_x,
                                        unsigned long long _first_frame, unsigned long long _n_frames,
                                        void *_d_y, void *_stream);
/* state = the last get_state_len() samples pushed (analyzer: input samples; synthesizer: scaled IDFT
 * outputs) plus the frame parity flag -- the checkpoint, and the multi-GPU halo */
unsigned int firpfbch2_crcf_get_state_len(firpfbch2_crcf _q);
void firpfbch2_crcf_get_state(firpfbch2_crcf _q, liquid_float_complex *_hist, int *_flag);
void firpfbch2_crcf_set_state(firpfbch2_crcf _q, const liquid_float_complex *_hist, int _flag);
unsigned int firpfbch2_crcf_get_num_channels(firpfbch2_crcf _q);

/* ---- multi-GPU driver: one long stream time-sharded over `world` processes, one GPU each ----
 * Rank r owns the frames [first_frame, first_frame + n_frames) of the stream, first_frame even (every segment
 * starts with the reference's flag = 0, firpfbch2.c:57).  The communicator wraps an NCCL communicator (NCCL is
 * loaded at run time, so single-GPU programs never need it): rank 0 calls liquid_cuda_comm_get_unique_id and
 * hands the liquid_cuda_comm_unique_id_bytes() bytes to the other ranks by any means (MPI_Bcast, a file, a
 * socket), then every rank calls liquid_cuda_comm_create.  Creating a communicator with world > 1 sets
 * NCCL_MAX/MIN_P2P_NCHANNELS=1 unless already set and reserves two SMs (liquid_cuda_set_reserved_sms) so that
 * the exchange runs beside the persistent channelizer kernel; destroy restores the reservation. */
typedef struct liquid_cuda_comm_s *liquid_cuda_comm;
int  liquid_cuda_comm_unique_id_bytes(void);
void liquid_cuda_comm_get_unique_id(void *_id);
liquid_cuda_comm liquid_cuda_comm_create(int _rank, int _world, const void *_id);    /* _id may be NULL when world == 1 */
void liquid_cuda_comm_destroy(liquid_cuda_comm _c);
int  liquid_cuda_comm_rank(liquid_cuda_comm _c);
int  liquid_cuda_comm_world(liquid_cuda_comm _c);
/* the exchange step on its own: _n_samples complex samples from _d_tail go to rank + 1, the left neighbour's
 * arrive in _d_head (rank 0 receives nothing), one ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd on _stream */
void liquid_cuda_comm_exchange_halo_dev(liquid_cuda_comm _c, void *_d_head, const void *_d_tail,
                                        unsigned long long _n_samples, void *_stream);
/* one pass of the sharded analyzer.  _d_x_with_headroom: get_state_len() samples of head-room followed by this
 * rank's n_frames * M/2 input samples; the left neighbour's last get_state_len() samples land in the head-room
 * (rank 0's head-room is read as it is: zeros = the reference's reset state).  The interior frames run on
 * _stream while the halo is in flight, the 4m halo-dependent frames follow on the communicator's side stream,
 * and _stream waits for them: on return everything is ordered on _stream.  Object state is not touched. */
void firpfbch2_crcf_execute_sharded_dev(firpfbch2_crcf _q, liquid_cuda_comm _c, void *_d_x_with_headroom,
                                        unsigned long long _first_frame, unsigned long long _n_frames,
                                        void *_d_y, void *_stream);
/* the same pass on host pointers: _x[_n_frames*M/2] is this rank's segment, _y[_n_frames*M] its output.  The last
 * get_state_len() samples go to the right neighbour, the left neighbour's become the object's history (rank 0: the
 * reset state), then the segment runs through the chunked H2D / compute / D2H pipeline of execute_block.  On return
 * the object holds the state after the segment, as after execute_block. */
void firpfbch2_crcf_execute_sharded(firpfbch2_crcf _q, liquid_cuda_comm _c, liquid_float_complex *_x,
                                    unsigned long long _first_frame, unsigned int _n_frames, liquid_float_complex *_y);

/* ---- firpfbch_crcf ---- */
void firpfbch_crcf_analyzer_execute_block(firpfbch_crcf _q, liquid_float_complex *_x,
                                          unsigned int _n_frames, liquid_float_complex *_y);
void firpfbch_crcf_synthesizer_execute_block(firpfbch_crcf _q, liquid_float_complex *_x,
                                             unsigned int _n_frames, liquid_float_complex *_y);
/* device pointers; dispatches on the object's type */
void firpfbch_crcf_execute_block_dev(firpfbch_crcf _q, const void *_d_x,
                                     unsigned long long _n_frames, void *_d_y, void *_stream);
/* _n_streams independent streams of _n_frames frames each (stream s at _d_x + s*_n_frames*M, same for
 * _d_y), every stream from the reset state; dispatches on the object's type; object state untouched */
void firpfbch_crcf_execute_streams_dev(firpfbch_crcf _q, const void *_d_x, unsigned int _n_streams,
                                       unsigned long long _n_frames, void *_d_y, void *_stream);
unsigned int firpfbch_crcf_get_state_len(firpfbch_crcf _q);
void firpfbch_crcf_get_state(firpfbch_crcf _q, liquid_float_complex *_hist);
void firpfbch_crcf_set_state(firpfbch_crcf _q, const liquid_float_complex *_hist);

/* ---- ofdmframesync: 1 while the object is in the payload state (between the long training symbol and the callback
 *      that ends the frame) ---- */
int ofdmframesync_is_frame_open(ofdmframesync _q);

/* ---- firdecim_crcf / firinterp_crcf: device-pointer forms (_n = outputs for firdecim, inputs for firinterp) ---- */
void firdecim_crcf_execute_block_dev(firdecim_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream);
void firinterp_crcf_execute_block_dev(firinterp_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream);

/* ---- resamp2_crcf / msresamp2_crcf: block forms over _n units (filter: samples; analyzer / synthesizer: pairs in,
 *      pairs out; decim: outputs from 2 _n inputs; interp: inputs giving 2 _n outputs; msresamp2: _n inputs giving
 *      _n 2^S outputs, or _n outputs from _n 2^S inputs), host and device pointers; get_taps copies the 4m+1 tap
 *      prototype and / or the 2m filter-branch taps (either pointer may be NULL) ---- */
void resamp2_crcf_get_taps(resamp2_crcf _q, float *_h, float *_h1);
void resamp2_crcf_filter_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n,
                                       liquid_float_complex *_y0, liquid_float_complex *_y1);
void resamp2_crcf_analyzer_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);
void resamp2_crcf_synthesizer_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);
void resamp2_crcf_decim_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);
void resamp2_crcf_interp_execute_block(resamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);
void resamp2_crcf_filter_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y0,
                                           void *_d_y1, void *_stream);
void resamp2_crcf_analyzer_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream);
void resamp2_crcf_synthesizer_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream);
void resamp2_crcf_decim_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream);
void resamp2_crcf_interp_execute_block_dev(resamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream);
unsigned int msresamp2_crcf_get_stage_semilength(msresamp2_crcf _q, unsigned int _i);
void msresamp2_crcf_execute_block(msresamp2_crcf _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);
void msresamp2_crcf_execute_block_dev(msresamp2_crcf _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream);

/* ---- ofdmframegen: _n data symbols in one call (_x [n][M] -> _y [n][M + cp_len]), host / device pointers ---- */
void ofdmframegen_get_sctype(ofdmframegen _q, unsigned char *_p);
void ofdmframegen_writesymbols(ofdmframegen _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y);
void ofdmframegen_writesymbols_dev(ofdmframegen _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream);

/* ---- fftfilt_crcf: _n_blocks blocks of n samples in one call (all blocks run in parallel), host / device pointers ---- */
void fftfilt_crcf_execute_block(fftfilt_crcf _q, liquid_float_complex *_x, unsigned int _n_blocks, liquid_float_complex *_y);
void fftfilt_crcf_execute_block_dev(fftfilt_crcf _q, const void *_d_x, unsigned long long _n_blocks, void *_d_y, void *_stream);

/* ---- spgramcf: device-pointer forms (the whole block becomes one batch of overlapping transforms) ---- */
void spgramcf_accumulate_psd_dev(spgramcf _q, const void *_d_x, float _alpha, unsigned long long _n, void *_stream);
void spgramcf_estimate_psd_dev(spgramcf _q, const void *_d_x, unsigned long long _n, float *_psd, void *_stream);
unsigned int spgramcf_get_num_transforms(spgramcf _q);

/* ---- firpfbch_cccf ---- */
void firpfbch_cccf_analyzer_execute_block(firpfbch_cccf _q, liquid_float_complex *_x,
                                          unsigned int _n_frames, liquid_float_complex *_y);
void firpfbch_cccf_synthesizer_execute_block(firpfbch_cccf _q, liquid_float_complex *_x,
                                             unsigned int _n_frames, liquid_float_complex *_y);
void firpfbch_cccf_execute_block_dev(firpfbch_cccf _q, const void *_d_x,
                                     unsigned long long _n_frames, void *_d_y, void *_stream);

/* ---- firfilt_crcf ---- */
/* _n_streams independent streams of _n samples each (stream s at _d_x + s*_n), every stream filtered
 * from the zero state with the object's taps and scale; the object's own state is not touched */
void firfilt_crcf_execute_streams_dev(firfilt_crcf _q, const void *_d_x, unsigned int _n_streams,
                                      unsigned long long _n, void *_d_y, void *_stream);
/* stateful single-stream block on device pointers (what execute_block does after its H2D copy) */
void firfilt_crcf_execute_block_dev(firfilt_crcf _q, const void *_d_x, unsigned long long _n,
                                    void *_d_y, void *_stream);

/* ---- firpfb_crcf ---- */
/* for i < _n: push _x[i], then _y[i] = execute(_idx[i]) */
void firpfb_crcf_execute_block(firpfb_crcf _q, liquid_float_complex *_x, const unsigned int *_idx,
                               unsigned int _n, liquid_float_complex *_y);
void firpfb_crcf_execute_block_dev(firpfb_crcf _q, const void *_d_x, const unsigned int *_d_idx, unsigned long long _n,
                                   void *_d_y, void *_stream);

/* ---- dotprod_crcf ---- */
/* _n_vec dot products: vector v starts at _d_x + v*_stride (complex samples) */
void dotprod_crcf_execute_batch_dev(dotprod_crcf _q, const void *_d_x, unsigned long long _n_vec,
                                    unsigned long long _stride, void *_d_y, void *_stream);

/* ---- FFT ---- */
/* _batch transforms of the plan's size on device pointers (row-major [batch][n]); in-place allowed */
void fft_execute_batch_dev(fftplan _p, const void *_d_x, void *_d_y, unsigned long long _batch, void *_stream);
void fft_execute_batch(fftplan _p, liquid_float_complex *_x, liquid_float_complex *_y, unsigned int _batch);

#ifdef __cplusplus
}
#endif
#endif /* LIQUID_CUDA_H */
