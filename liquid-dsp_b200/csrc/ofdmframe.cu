// ofdmframe.cu -- ofdmframegen: the OFDM symbol generator, the other tenant of src/multichannel.  SURVEY.md 8(f) item 4
// (the per-symbol path: subcarrier mapping, IFFT, cyclic prefix, tapered overlap; frame synchronisation is serial and
// stays out of scope).
//
// Replaces src/multichannel/src/ofdmframegen.c (create :85-171, destroy :174-198, print :201-211, reset :214-222,
// write_S0a :237-250, write_S0b :252-264, write_S1 :266-272, writesymbol :279-311, writetail :314-322, gensymbol
// :345-361) and src/multichannel/src/ofdmframe.common.c (init_S0 :41-101, init_S1 :109-167, init_default_sctype
// :176-214, validate_sctype :224-257, print_sctype :261-282); the pilot / PLCP bits come from the default m-sequences
// of src/sequence/src/msequence.c (:41-58, advance :166-177).
//
// A block of n data symbols is data-parallel except for two thin threads of state, both made index-computable: the
// pilot bits (register length 8, period 255: symbol s of the block takes bits pos + s n_pilot + rank(k) of a 255-entry
// table, rank = position of pilot k in the reference's fftshifted traversal) and the tapered overlap (a symbol's head
// needs the first taper_len samples of the previous symbol's IFFT, which are sitting unmodified in that symbol's body
// in the output).  Power-of-two M that fits shared memory: ONE kernel maps subcarriers straight into the first pass of
// the shared-memory IFFT (lcfft::stockham_run) and its last pass writes body + tapered cyclic prefix from registers;
// a second, tiny kernel adds the overlap terms.  Any other M: map -> batched DFT -> assemble through scratch.
#include <cmath>
#include <cstdint>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"
#include "host_pipeline.cuh"
#include "fft_device.cuh"
#include "ofdm_common.cuh"
#include "tile_fft.cuh"
#include "quad_fft.cuh"

struct ofdmframegen_s {
    unsigned int M = 0, cp_len = 0, taper_len = 0;
    std::vector<unsigned char> p;
    unsigned int M_null = 0, M_pilot = 0, M_data = 0, M_S0 = 0, M_S1 = 0;
    float g_data = 1.0f;
    std::vector<float> taper;
    std::vector<lc_cf> S0, s0, S1, s1;
    unsigned int pilot_pos = 0;          // bits of the pilot sequence consumed so far, mod 255
    // device side
    int *d_code = nullptr;               // [M] -1 null, -2 data, r >= 0: pilot number (mod 255) of the traversal i -> k = (i + M/2) % M
    float *d_pil = nullptr;              // [255] pilot values +-g_data, from the register's reset state
    float *d_taper = nullptr;            // [taper_len]
    float2 *d_postfix = nullptr;         // [taper_len] first samples of the previous symbol
    float2 *d_s1 = nullptr;              // [M] long PLCP symbol (time)
    lc::FftPlanDev ifft;
    float2 *d_tw_tile = nullptr;         // backward twiddles of the tile (M = 64 ... 512) or quad (M = 1024 ... 4096) transform, or null
    lc::DevBuf stage_in, stage_out, scratch;
    cudaStream_t stream = nullptr;
};

namespace {

using namespace lcofdm;
constexpr int kThreads = 256;
struct MapArgs {
    const int *code;                 // per subcarrier: -1 null, -2 data, r >= 0 pilot number r of the traversal
    const float *pil;
    unsigned int M, npm, pos;        // npm = n_pilot mod 255; pos = pilot bits consumed before the block, mod 255
    float g;
};

// pilot-table position of the first pilot of a symbol: (pos + (s mod 255) n_pilot) mod 255 -- one integer division per
// symbol and thread, not one per subcarrier (the mapping, not the transform, bounded the symbol kernels)
__device__ __forceinline__ unsigned int pilot_base(const MapArgs &a, unsigned int s_mod)
{
    return (s_mod * a.npm + a.pos) % kPilotPeriod;
}

// frequency-domain value of subcarrier k of symbol s of the block
__device__ __forceinline__ float2 map_subcarrier(const MapArgs &a, const float2 *__restrict__ x, unsigned long long s,
                                                 unsigned int base, unsigned int k)
{
    // branch-free: every load is issued unconditionally (x holds M values per symbol, used or not), so the 16 element
    // fetches of a butterfly overlap instead of queueing behind each other's type test
    const int c = __ldg(a.code + k);
    const float2 v = __ldg(x + s * a.M + k);
    unsigned int idx = base + (unsigned int)(c > 0 ? c : 0);      // both terms < 255
    idx = idx >= kPilotPeriod ? idx - kPilotPeriod : idx;
    const float pv = __ldg(a.pil + idx);
    const float2 d = make_float2(v.x * a.g, v.y * a.g);
    return c == -2 ? d : make_float2(c >= 0 ? pv : 0.f, 0.f);
}

struct SymbolSrc {
    MapArgs a;
    const float2 *x;
    unsigned long long s0, s_end;
    unsigned int s0_mod;             // s0 mod 255
    __device__ __forceinline__ float2 operator()(unsigned int f, unsigned int k) const
    {
        const unsigned long long s = s0 + f;
        return s < s_end ? map_subcarrier(a, x, s, pilot_base(a, (s0_mod + f) % kPilotPeriod), k) : make_float2(0.f, 0.f);
    }
};

// sample k of symbol s's IFFT: the body, and -- for the last cp_len samples -- the cyclic prefix, tapered
struct PrefixDst {
    float2 *y;
    const float *taper;
    unsigned int M, cp, tl;
    unsigned long long s0, s_end;
    __device__ __forceinline__ void operator()(unsigned int f, unsigned int k, float2 v) const
    {
        const unsigned long long s = s0 + f;
        if (s >= s_end) return;
        float2 *base = y + s * (M + cp);
        base[cp + k] = v;
        if (k >= M - cp) {
            const unsigned int i = k - (M - cp);
            if (i < tl) {
                const float w = __ldg(taper + i);
                v = make_float2(v.x * w, v.y * w);
            }
            base[i] = v;
        }
    }
};

// M = 64 ... 512: the tile transform of tile_fft.cuh between the subcarrier map and the prefix writer -- row DFTs out of
// register-prefetched subcarrier values, column DFTs in a conflict-free shared tile, the last radix-2 stage in the thread
// that writes body and tapered prefix; one instance barrier per tile of 8 symbols instead of a block barrier per pass
// W32: 32-bit element offsets (every shape but M = 512, where they cost a spill and 10 % of the rate)
template <bool W32>
struct SymbolSrcT {
    MapArgs a;
    const float2 *x;
    // the user's value of the subcarrier, loaded a tile ahead and left alone until it is used
    // (32-bit element offsets: the launcher takes this path only while n (M + cp) < 2^32 -- the 64-bit products were a
    // fifth of the M = 64 kernel's instructions and pushed it into spills)
    __device__ __forceinline__ float2 load(long long s, unsigned int k) const 
    {
        return W32 ? __ldg(x + ((unsigned int)s * a.M + k)) : __ldg(x + s * a.M + k);
    }
    // ofdmframegen.c:279-311 at the point of use: data scaled, pilots from the table, nulls zero (the type and pilot
    // tables sit in L1).  32-bit remainder by a constant (the 64-bit one is a subroutine): n_sym < 2^31 here
    __device__ __forceinline__ float2 finish(long long s, unsigned int k, float2 v) const
    {
        const int c = __ldg(a.code + k);
        unsigned int idx = pilot_base(a, (unsigned int)s % kPilotPeriod) + (unsigned int)(c > 0 ? c : 0);
        idx = idx >= kPilotPeriod ? idx - kPilotPeriod : idx;
        const float pv = __ldg(a.pil + idx);
        return c == -2 ? make_float2(v.x * a.g, v.y * a.g) : make_float2(c >= 0 ? pv : 0.f, 0.f);
    }
};
template <bool W32>
struct PrefixDstT {
    float2 *y;
    const float *taper;
    unsigned int M, cp, tl;
    __device__ __forceinline__ void operator()(long long s, unsigned int k, float2 v) const
    {
        float2 *base = W32 ? y + (unsigned int)s * (M + cp) : y + s * (M + cp);
        __stcs(base + cp + k, v);
        if (k >= M - cp) {
            const unsigned int i = k - (M - cp);
            if (i < tl) {
                const float w = __ldg(taper + i);
                v = make_float2(v.x * w, v.y * w);
            }
            base[i] = v;
        }
    }
};

template <int M>
__global__ void __launch_bounds__(lctile::kThreads, 2)
ofdm_symbols_tile_kernel(MapArgs a, const float2 *__restrict__ x, const float2 *__restrict__ tw, const float *__restrict__ taper,
                         float2 *__restrict__ y, unsigned int cp, unsigned int tl, long long n_sym, int tiles_per_inst)
{
    extern __shared__ __align__(16) unsigned char smem_tile[];
    lctile::run<M, +1>(smem_tile, tw, n_sym, tiles_per_inst, SymbolSrcT<M != 512>{a, x}, PrefixDstT<M != 512>{y, taper, (unsigned int)M, cp, tl});
}

template <int M>
void launch_symbols_tile(const MapArgs &a, const float2 *d_x, const float2 *d_tw, const float *d_taper, float2 *d_y,
                         unsigned int cp, unsigned int tl, unsigned long long n, cudaStream_t st)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(ofdm_symbols_tile_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)lctile::smem_bytes<M>()));
    unsigned int grid = 0;
    int tpi = 1;
    lctile::plan(M, n, grid, tpi);
    ofdm_symbols_tile_kernel<M><<<grid, lctile::kThreads, lctile::smem_bytes<M>(), st>>>(a, d_x, d_tw, d_taper, d_y, cp, tl,
                                                                                         (long long)n, tpi);
}

// M = 1024 / 2048 / 4096: the quad transform of quad_fft.cuh between the same two functors
template <int RA>
__global__ void __launch_bounds__(lcquad::kThreads, 2)
ofdm_symbols_quad_kernel(MapArgs a, const float2 *__restrict__ x, const float2 *__restrict__ tw, const float *__restrict__ taper,
                         float2 *__restrict__ y, unsigned int cp, unsigned int tl, long long n_sym, int tiles_per_inst)
{
    extern __shared__ __align__(16) unsigned char smem_quad[];
    lcquad::run<RA, +1, (RA >= 8)>(smem_quad, tw, n_sym, tiles_per_inst, SymbolSrcT<true>{a, x}, PrefixDstT<true>{y, taper, 256u * RA, cp, tl});
}

template <int RA>
void launch_symbols_quad(const MapArgs &a, const float2 *d_x, const float2 *d_tw, const float *d_taper, float2 *d_y,
                         unsigned int cp, unsigned int tl, unsigned long long n, cudaStream_t st)
{
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(ofdm_symbols_quad_kernel<RA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)lcquad::smem_bytes<RA>()));
    unsigned int grid = 0;
    int tpi = 1;
    lcquad::plan(RA, n, grid, tpi);
    ofdm_symbols_quad_kernel<RA><<<grid, lcquad::kThreads, lcquad::smem_bytes<RA>(), st>>>(a, d_x, d_tw, d_taper, d_y, cp, tl,
                                                                                           (long long)n, tpi);
}

template <bool TWG>
__global__ void __launch_bounds__(lcfft::kFftThreads, 3)
ofdm_symbols_pow2_kernel(MapArgs a, const float2 *__restrict__ x, const float2 *__restrict__ tw, const float *__restrict__ taper,
                         float2 *__restrict__ y, unsigned int log2M, unsigned int per_cta, unsigned int cp, unsigned int tl,
                         unsigned long long n_sym)
{
    extern __shared__ float2 smem_f2[];
    lcfft::Smem sm = lcfft::carve(smem_f2, a.M, per_cta);
    if (TWG) sm.tw = const_cast<float2 *>(tw);
    else if (log2M > 4) lcfft::stage_twiddles(sm, tw, a.M);
    const unsigned long long s0 = (unsigned long long)blockIdx.x * per_cta;
    lcfft::stockham_run<+1, TWG>(sm, a.M, log2M, per_cta, SymbolSrc{a, x, s0, n_sym, (unsigned int)(s0 % kPilotPeriod)},
                            PrefixDst{y, taper, a.M, cp, tl, s0, n_sym});
}

// y[s][i] += post[i] * taper[tl-1-i], post = first samples of the previous symbol: its body in y (unmodified there,
// cp >= tl), or the object's postfix for the block's first symbol.  Separate multiply and add, as the reference.
__global__ void __launch_bounds__(kThreads)
ofdm_overlap_kernel(float2 *__restrict__ y, const float2 *__restrict__ postfix, const float *__restrict__ taper, unsigned int M,
                    unsigned int cp, unsigned int tl, unsigned long long n_sym)
{
    const unsigned long long total = n_sym * tl;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long s = g / tl;
        const unsigned int i = (unsigned int)(g % tl);
        const float2 post = s == 0 ? __ldg(postfix + i) : y[(s - 1) * (M + cp) + cp + i];
        const float w = __ldg(taper + (tl - 1 - i));
        float2 *dst = y + s * (M + cp) + i;
        const float2 v = *dst;
        *dst = make_float2(__fadd_rn(v.x, __fmul_rn(post.x, w)), __fadd_rn(v.y, __fmul_rn(post.y, w)));
    }
}

// ---- any other M: through scratch ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
ofdm_map_kernel(MapArgs a, const float2 *__restrict__ x, float2 *__restrict__ X, unsigned long long n_sym)
{
    const unsigned long long total = n_sym * a.M;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (unsigned long long)gridDim.x * blockDim.x)
    {
        const unsigned long long sy = g / a.M;
        X[g] = map_subcarrier(a, x, sy, pilot_base(a, (unsigned int)(sy % kPilotPeriod)), (unsigned int)(g % a.M));
    }
}

// time-domain symbols t [n][M] -> y [n][M + cp] with prefix, taper and overlap (ofdmframegen.c:345-361)
__global__ void __launch_bounds__(kThreads)
ofdm_assemble_kernel(const float2 *__restrict__ t, const float2 *__restrict__ postfix, const float *__restrict__ taper,
                     float2 *__restrict__ y, unsigned int M, unsigned int cp, unsigned int tl, unsigned long long n_sym)
{
    const unsigned int len = M + cp;
    const unsigned long long total = n_sym * len;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long s = g / len;
        const unsigned int i = (unsigned int)(g % len);
        float2 v = t[s * M + (i + M - cp) % M];
        if (i < tl) {
            const float a = __ldg(taper + i), b = __ldg(taper + (tl - 1 - i));
            const float2 post = s == 0 ? __ldg(postfix + i) : t[(s - 1) * M + i];
            v = make_float2(__fmul_rn(v.x, a), __fmul_rn(v.y, a));
            v = make_float2(__fadd_rn(v.x, __fmul_rn(post.x, b)), __fadd_rn(v.y, __fmul_rn(post.y, b)));
        }
        y[g] = v;
    }
}

inline unsigned int grid_for(unsigned long long total)
{
    unsigned long long b = (total + kThreads - 1) / kThreads;
    const unsigned long long cap = 148ull * 8 * 16;
    return (unsigned int)(b < cap ? (b ? b : 1) : cap);
}

// prefix / taper / overlap of n time-domain symbols that are already on the device; updates the postfix
void assemble_dev(ofdmframegen q, const float2 *d_t, unsigned long long n, float2 *d_y, cudaStream_t st)
{
    ofdm_assemble_kernel<<<grid_for(n * (q->M + q->cp_len)), kThreads, 0, st>>>(d_t, q->d_postfix, q->d_taper, d_y, q->M, q->cp_len,
                                                                               q->taper_len, n);
    LC_LAUNCH_CHECK();
    if (q->taper_len)
        LC_CUDA_CHECK(cudaMemcpyAsync(q->d_postfix, d_t + (n - 1) * q->M, sizeof(float2) * q->taper_len, cudaMemcpyDeviceToDevice, st));
}

void symbols_dev(ofdmframegen q, const float2 *d_x, unsigned long long n, float2 *d_y, cudaStream_t st)
{
    if (n == 0) return;
    const MapArgs a{q->d_code, q->d_pil, q->M, q->M_pilot % kPilotPeriod, q->pilot_pos, q->g_data};
    const unsigned int M = q->M;
    if (q->d_tw_tile && n * (M + q->cp_len) < (1ull << 32) && !getenv("LIQUID_CUDA_OFDM_GENERIC") && !getenv("LIQUID_CUDA_NO_FUSED")) {
        // the pilot position of symbol s is (pos + (s mod 255) npm) mod 255 with s counted from the start of the block
        switch (M) {
        case 64: launch_symbols_tile<64>(a, d_x, q->d_tw_tile, q->d_taper, d_y, q->cp_len, q->taper_len, n, st); break;
        case 128: launch_symbols_tile<128>(a, d_x, q->d_tw_tile, q->d_taper, d_y, q->cp_len, q->taper_len, n, st); break;
        case 256: launch_symbols_tile<256>(a, d_x, q->d_tw_tile, q->d_taper, d_y, q->cp_len, q->taper_len, n, st); break;
        case 512: launch_symbols_tile<512>(a, d_x, q->d_tw_tile, q->d_taper, d_y, q->cp_len, q->taper_len, n, st); break;
        case 1024: launch_symbols_quad<4>(a, d_x, q->d_tw_tile, q->d_taper, d_y, q->cp_len, q->taper_len, n, st); break;
        case 2048: launch_symbols_quad<8>(a, d_x, q->d_tw_tile, q->d_taper, d_y, q->cp_len, q->taper_len, n, st); break;
        default: launch_symbols_quad<16>(a, d_x, q->d_tw_tile, q->d_taper, d_y, q->cp_len, q->taper_len, n, st); break;
        }
        LC_LAUNCH_CHECK();
        if (q->taper_len) {
            ofdm_overlap_kernel<<<grid_for(n * q->taper_len), kThreads, 0, st>>>(d_y, q->d_postfix, q->d_taper, M, q->cp_len, q->taper_len, n);
            LC_LAUNCH_CHECK();
            LC_CUDA_CHECK(cudaMemcpyAsync(q->d_postfix, d_y + (n - 1) * (M + q->cp_len) + q->cp_len, sizeof(float2) * q->taper_len,
                                          cudaMemcpyDeviceToDevice, st));
        }
    } else if (lcfft::pow2_in_smem(M) && !getenv("LIQUID_CUDA_OFDM_GENERIC")) {
        static std::atomic<unsigned long long> configured{0};
        if (lc::first_time_on_device(configured)) {
            LC_CUDA_CHECK(cudaFuncSetAttribute(ofdm_symbols_pow2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            LC_CUDA_CHECK(cudaFuncSetAttribute(ofdm_symbols_pow2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
        const unsigned int per_cta = lcfft::per_cta_for(M, n);
        const unsigned long long ctas = (n + per_cta - 1) / per_cta;
        if (ctas >= (1ull << 31)) lc::fatal("ofdmframegen_writesymbols", "too many symbols for one launch");
        const bool twg = M >= lcfft::kTwGlobalFrom;
        const size_t smem = lcfft::smem_for(M, per_cta, !twg);
        if (twg)
            ofdm_symbols_pow2_kernel<true><<<(unsigned int)ctas, lcfft::kFftThreads, smem, st>>>(
                a, d_x, q->ifft.d_twp, q->d_taper, d_y, lc::ilog2(M), per_cta, q->cp_len, q->taper_len, n);
        else
            ofdm_symbols_pow2_kernel<false><<<(unsigned int)ctas, lcfft::kFftThreads, smem, st>>>(
                a, d_x, q->ifft.d_twp, q->d_taper, d_y, lc::ilog2(M), per_cta, q->cp_len, q->taper_len, n);
        LC_LAUNCH_CHECK();
        if (q->taper_len) {
            ofdm_overlap_kernel<<<grid_for(n * q->taper_len), kThreads, 0, st>>>(d_y, q->d_postfix, q->d_taper, M, q->cp_len, q->taper_len, n);
            LC_LAUNCH_CHECK();
            LC_CUDA_CHECK(cudaMemcpyAsync(q->d_postfix, d_y + (n - 1) * (M + q->cp_len) + q->cp_len, sizeof(float2) * q->taper_len,
                                          cudaMemcpyDeviceToDevice, st));
        }
    } else {
        // pieces of at most ~256 MB of scratch
        unsigned long long lim = (256ull << 20) / (sizeof(float2) * M);
        if (lim == 0) lim = 1;
        const unsigned long long piece = n < lim ? n : lim;
        float2 *d_t = (float2 *)q->scratch.get(sizeof(float2) * piece * M);
        for (unsigned long long s = 0; s < n; s += piece) {
            const unsigned long long ns = (n - s < piece) ? n - s : piece;
            MapArgs ap = a;
            ap.pos = (unsigned int)((a.pos + (s % kPilotPeriod) * (q->M_pilot % kPilotPeriod)) % kPilotPeriod);
            ofdm_map_kernel<<<grid_for(ns * M), kThreads, 0, st>>>(ap, d_x + s * M, d_t, ns);
            LC_LAUNCH_CHECK();
            q->ifft.run(d_t, d_t, ns, st);
            assemble_dev(q, d_t, ns, d_y + s * (M + q->cp_len), st);
        }
    }
    q->pilot_pos = (unsigned int)((q->pilot_pos + (n % kPilotPeriod) * (q->M_pilot % kPilotPeriod)) % kPilotPeriod);
}

}  // namespace

extern "C" {

void ofdmframe_init_default_sctype(unsigned int _M, unsigned char *_p)
{
    if (_M < 6) fprintf(stderr, "warning: ofdmframe_init_default_sctype(), less than 4 subcarriers\n");
    const unsigned int M2 = _M / 2;
    unsigned int G = _M / 10;
    if (G < 2) G = 2;
    const unsigned int P = (_M > 34) ? 8 : 4, P2 = P / 2;
    for (unsigned int i = 0; i < _M; i++) _p[i] = OFDMFRAME_SCTYPE_NULL;
    for (unsigned int i = 1; i + G < M2; i++) {
        const unsigned char t = (((i + P2) % P) == 0) ? OFDMFRAME_SCTYPE_PILOT : OFDMFRAME_SCTYPE_DATA;
        _p[i] = t;                      // upper band
        _p[_M - i] = t;                 // lower band
    }
}

void ofdmframe_validate_sctype(unsigned char *_p, unsigned int _M, unsigned int *_M_null, unsigned int *_M_pilot, unsigned int *_M_data)
{
    unsigned int n0 = 0, n1 = 0, n2 = 0;
    for (unsigned int i = 0; i < _M; i++) {
        if (_p[i] == OFDMFRAME_SCTYPE_NULL) n0++;
        else if (_p[i] == OFDMFRAME_SCTYPE_PILOT) n1++;
        else if (_p[i] == OFDMFRAME_SCTYPE_DATA) n2++;
        else lc::fatal("ofdmframe_validate_sctype", "invalid subcarrier type");
    }
    *_M_null = n0;
    *_M_pilot = n1;
    *_M_data = n2;
}

void ofdmframe_print_sctype(unsigned char *_p, unsigned int _M)
{
    for (unsigned int i = 0; i < _M; i++) {
        const unsigned int k = (i + _M / 2) % _M;
        putchar(_p[k] == OFDMFRAME_SCTYPE_NULL ? '.' : _p[k] == OFDMFRAME_SCTYPE_PILOT ? 'P' : '+');
    }
    printf("\n");
}

ofdmframegen ofdmframegen_create(unsigned int _M, unsigned int _cp_len, unsigned int _taper_len, unsigned char *_p)
{
    if (_M < 2) lc::fatal("ofdmframegen_create", "number of subcarriers must be at least 2");
    if (_M % 2) lc::fatal("ofdmframegen_create", "number of subcarriers must be even");
    if (_cp_len > _M) lc::fatal("ofdmframegen_create", "cyclic prefix cannot exceed symbol length");
    if (_taper_len > _cp_len) lc::fatal("ofdmframegen_create", "taper length cannot exceed cyclic prefix");
    std::vector<unsigned char> p(_M);
    if (_p == nullptr) ofdmframe_init_default_sctype(_M, p.data());
    else std::copy(_p, _p + _M, p.begin());
    unsigned int n0, n1, n2;
    ofdmframe_validate_sctype(p.data(), _M, &n0, &n1, &n2);
    if (n1 + n2 == 0) lc::fatal("ofdmframegen_create", "must have at least one enabled subcarrier");
    if (n2 == 0) lc::fatal("ofdmframegen_create", "must have at least one data subcarriers");
    if (n1 < 2) lc::fatal("ofdmframegen_create", "must have at least two pilot subcarriers");
    lc::ensure_device();
    ofdmframegen q = new ofdmframegen_s();
    q->M = _M;
    q->cp_len = _cp_len;
    q->taper_len = _taper_len;
    q->p = p;
    q->M_null = n0;
    q->M_pilot = n1;
    q->M_data = n2;
    q->g_data = 1.0f / sqrtf((float)(n1 + n2));
    q->taper.resize(_taper_len);
    for (unsigned int i = 0; i < _taper_len; i++) {
        const float t = ((float)i + 0.5f) / (float)_taper_len;
        const float g = sinf((float)(M_PI_2 * (double)t));
        q->taper[i] = g * g;
    }
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    q->ifft.init(_M, +1);
    if (lctile::supported(_M) || lcquad::supported(_M)) {
        std::vector<float2> tw;
        if (lctile::supported(_M)) lctile::make_twiddles(tw, _M, +1);
        else lcquad::make_twiddles(tw, _M, +1);
        LC_CUDA_CHECK(cudaMalloc(&q->d_tw_tile, tw.size() * sizeof(float2)));
        LC_CUDA_CHECK(cudaMemcpy(q->d_tw_tile, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    // pilot traversal order (writesymbol walks i and fills k = (i + M/2) % M) and the 255 pilot values
    std::vector<int> code(_M, -1);
    int r = 0;
    for (unsigned int i = 0; i < _M; i++) {
        const unsigned int k = (i + _M / 2) % _M;
        if (p[k] == OFDMFRAME_SCTYPE_PILOT) code[k] = (int)(r++ % kPilotPeriod);   // the rank is only ever used mod 255
        else if (p[k] == OFDMFRAME_SCTYPE_DATA) code[k] = -2;
    }
    std::vector<float> pil(kPilotPeriod);
    Mseq ms(8);
    for (unsigned int i = 0; i < kPilotPeriod; i++) pil[i] = (ms.advance() ? 1.0f : -1.0f) * q->g_data;
    const size_t tl = _taper_len ? _taper_len : 1;
    LC_CUDA_CHECK(cudaMalloc(&q->d_code, sizeof(int) * _M));
    LC_CUDA_CHECK(cudaMalloc(&q->d_pil, sizeof(float) * kPilotPeriod));
    LC_CUDA_CHECK(cudaMalloc(&q->d_taper, sizeof(float) * tl));
    LC_CUDA_CHECK(cudaMalloc(&q->d_postfix, sizeof(float2) * tl));
    LC_CUDA_CHECK(cudaMalloc(&q->d_s1, sizeof(float2) * _M));
    LC_CUDA_CHECK(cudaMemcpy(q->d_code, code.data(), sizeof(int) * _M, cudaMemcpyHostToDevice));
    LC_CUDA_CHECK(cudaMemcpy(q->d_pil, pil.data(), sizeof(float) * kPilotPeriod, cudaMemcpyHostToDevice));
    if (_taper_len) LC_CUDA_CHECK(cudaMemcpy(q->d_taper, q->taper.data(), sizeof(float) * _taper_len, cudaMemcpyHostToDevice));
    // PLCP symbols: frequency domain on the host, time domain by the device IFFT, scaled by 1/sqrt(enabled)
    float2 *d_tmp = (float2 *)q->scratch.get(sizeof(float2) * 2 * _M);
    for (int which = 0; which < 2; which++) {
        std::vector<lc_cf> &S = which ? q->S1 : q->S0, &s = which ? q->s1 : q->s0;
        const unsigned int enabled = plcp_freq(p, _M, which, S);
        if (enabled == 0) lc::fatal(which ? "ofdmframe_init_S1" : "ofdmframe_init_S0", "no subcarriers enabled; check allocation");
        (which ? q->M_S1 : q->M_S0) = enabled;
        LC_CUDA_CHECK(cudaMemcpyAsync(d_tmp, S.data(), sizeof(float2) * _M, cudaMemcpyHostToDevice, q->stream));
        q->ifft.run(d_tmp, d_tmp + _M, 1, q->stream);
        s.resize(_M);
        LC_CUDA_CHECK(cudaMemcpyAsync(s.data(), d_tmp + _M, sizeof(float2) * _M, cudaMemcpyDeviceToHost, q->stream));
        LC_CUDA_CHECK(cudaStreamSynchronize(q->stream));
        const float g = 1.0f / sqrtf((float)enabled);
        for (unsigned int i = 0; i < _M; i++) { s[i].re *= g; s[i].im *= g; }
    }
    LC_CUDA_CHECK(cudaMemcpy(q->d_s1, q->s1.data(), sizeof(float2) * _M, cudaMemcpyHostToDevice));
    ofdmframegen_reset(q);
    return q;
}

void ofdmframegen_destroy(ofdmframegen _q)
{
    if (!_q) return;
    cudaStreamSynchronize(_q->stream);
    _q->ifft.destroy();
    cudaFree(_q->d_code);
    cudaFree(_q->d_pil);
    cudaFree(_q->d_taper);
    cudaFree(_q->d_postfix);
    cudaFree(_q->d_s1);
    cudaFree(_q->d_tw_tile);
    _q->stage_in.release();
    _q->stage_out.release();
    _q->scratch.release();
    cudaStreamDestroy(_q->stream);
    delete _q;
}

void ofdmframegen_print(ofdmframegen _q)
{
    printf("ofdmframegen (cuda):\n");
    printf("    num subcarriers     :   %-u\n", _q->M);
    printf("      - NULL            :   %-u\n", _q->M_null);
    printf("      - pilot           :   %-u\n", _q->M_pilot);
    printf("      - data            :   %-u\n", _q->M_data);
    printf("    cyclic prefix len   :   %-u\n", _q->cp_len);
    printf("    taper len           :   %-u\n", _q->taper_len);
    printf("    ");
    ofdmframe_print_sctype(_q->p.data(), _q->M);
}

void ofdmframegen_reset(ofdmframegen _q)
{
    _q->pilot_pos = 0;
    if (_q->taper_len) LC_CUDA_CHECK(cudaMemsetAsync(_q->d_postfix, 0, sizeof(float2) * _q->taper_len, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

void ofdmframegen_get_sctype(ofdmframegen _q, unsigned char *_p) { std::copy(_q->p.begin(), _q->p.end(), _p); }

// the two short-sequence symbols are rotations of the constant s0 (computed by the device IFFT at create)
void ofdmframegen_write_S0a(ofdmframegen _q, liquid_float_complex *_y)
{
    lc_cf *y = reinterpret_cast<lc_cf *>(_y);
    for (unsigned int i = 0; i < _q->M + _q->cp_len; i++) y[i] = _q->s0[(i + _q->M - 2 * _q->cp_len) % _q->M];
    for (unsigned int i = 0; i < _q->taper_len; i++) { y[i].re *= _q->taper[i]; y[i].im *= _q->taper[i]; }
}

void ofdmframegen_write_S0b(ofdmframegen _q, liquid_float_complex *_y)
{
    lc_cf *y = reinterpret_cast<lc_cf *>(_y);
    for (unsigned int i = 0; i < _q->M + _q->cp_len; i++) y[i] = _q->s0[(i + _q->M - _q->cp_len) % _q->M];
    if (_q->taper_len) {
        LC_CUDA_CHECK(cudaMemcpyAsync(_q->d_postfix, _q->s0.data(), sizeof(float2) * _q->taper_len, cudaMemcpyHostToDevice, _q->stream));
        LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    }
}

void ofdmframegen_write_S1(ofdmframegen _q, liquid_float_complex *_y)
{
    float2 *d_y = (float2 *)_q->stage_out.get(sizeof(float2) * (_q->M + _q->cp_len));
    assemble_dev(_q, _q->d_s1, 1, d_y, _q->stream);
    LC_CUDA_CHECK(cudaMemcpyAsync(_y, d_y, sizeof(float2) * (_q->M + _q->cp_len), cudaMemcpyDeviceToHost, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
}

// _n data symbols: _x [n][M] subcarrier values (only the data subcarriers are read) -> _y [n][M + cp_len]
void ofdmframegen_writesymbols_dev(ofdmframegen _q, const void *_d_x, unsigned long long _n, void *_d_y, void *_stream)
{
    symbols_dev(_q, (const float2 *)_d_x, _n, (float2 *)_d_y, lc::as_stream(_stream));
}

void ofdmframegen_writesymbols(ofdmframegen _q, liquid_float_complex *_x, unsigned int _n, liquid_float_complex *_y)
{
    lc::host_block(_q->stream, _q->stage_in, _q->stage_out, reinterpret_cast<const lc_cf *>(_x), reinterpret_cast<lc_cf *>(_y),
                   _n, _q->M, _q->M + _q->cp_len,
                   [&](float2 *di, size_t nu, float2 *dob, cudaStream_t st) { symbols_dev(_q, di, nu, dob, st); });
}

void ofdmframegen_writesymbol(ofdmframegen _q, liquid_float_complex *_x, liquid_float_complex *_y)
{
    ofdmframegen_writesymbols(_q, _x, 1, _y);
}

void ofdmframegen_writetail(ofdmframegen _q, liquid_float_complex *_buffer)
{
    if (_q->taper_len == 0) return;
    std::vector<lc_cf> post(_q->taper_len);
    LC_CUDA_CHECK(cudaMemcpyAsync(post.data(), _q->d_postfix, sizeof(float2) * _q->taper_len, cudaMemcpyDeviceToHost, _q->stream));
    LC_CUDA_CHECK(cudaStreamSynchronize(_q->stream));
    lc_cf *y = reinterpret_cast<lc_cf *>(_buffer);
    for (unsigned int i = 0; i < _q->taper_len; i++) {
        const float w = _q->taper[_q->taper_len - i - 1];
        y[i].re = post[i].re * w;
        y[i].im = post[i].im * w;
    }
}

}  // extern "C"
