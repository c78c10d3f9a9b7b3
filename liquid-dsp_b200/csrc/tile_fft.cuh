// tile_fft.cuh -- batched M-point transforms (M = 64 / 128 / 256 / 512) in the tile form of pfbch2_tile.cu, for the callers
// of the channelizers that are a transform with something on either side (ofdmframegen: subcarrier map -> IFFT -> cyclic
// prefix; the fft_* batch API).  The caller supplies a source functor with two halves --  float2 load(long long t, unsigned k)
// issues the global load of input k of transform t a tile AHEAD and must not touch the value (arithmetic on it would
// make the prefetch wait for the memory latency there and then),  float2 finish(long long t, unsigned k, float2 raw)
// turns it into the transform input when it is consumed -- and a destination functor  void dst(long long t, unsigned n,
// float2 v)  (output n of transform t; a thread is handed the same two outputs u and u + M/2 of every transform of its
// instance, in transform order, so a destination passed as an lvalue may keep per-thread state -- spgram.cu folds
// |X|^2 into two registers); the transform never makes a trip through HBM or a block-wide shared-memory pass:
//   an instance = M/2 threads walks tiles of 8 transforms.  Decimation in time, X_p[k] = X[2k + p]:
//     pass B'  R2 = M/32-point register DFTs over the rows X_p[k1 + 16 k2], inputs prefetched into registers a tile ahead
//     pass A'  inter-pass twiddle and 16-point register DFTs down the columns of the shared tile (conflict-free layout
//              of tools/tile_layout_search.py), which leaves E_p[u] = sum_k X_p[k] e^{s j 2 pi u k / (M/2)}
//     last     thread u: out[u] = E_0[u] + W_M^u E_1[u], out[u + M/2] = E_0[u] - W_M^u E_1[u]  -> dst
//   Two tile buffers per instance leave one instance barrier per tile.
#pragma once
#include <cmath>
#include <vector>

#include "fused_common.cuh"

namespace lctile {

using namespace lcf;

constexpr int kThreads = 256;
constexpr int kFr = 8;                          // transforms per tile
constexpr int kTwPitch = 18;                    // float2 per column of the shared twiddle copy (16-byte reads, conflict-free)

// address (8-byte units) = f*FS + p*SS + row*PR + col; WIDE: pass B' takes its rows across the whole instance
template <int M> struct Layout;
template <> struct Layout<64>  { static constexpr int PR = 3,  SS = 56,  FS = 106, WIDE = 1, NBUF = 1; };
template <> struct Layout<128> { static constexpr int PR = 5,  SS = 88,  FS = 172, WIDE = 1, NBUF = 2; };
template <> struct Layout<256> { static constexpr int PR = 9,  SS = 152, FS = 295, WIDE = 0, NBUF = 2; };
template <> struct Layout<512> { static constexpr int PR = 17, SS = 280, FS = 551, WIDE = 0, NBUF = 2; };

inline bool supported(unsigned int M) { return M == 64 || M == 128 || M == 256 || M == 512; }

template <int M>
constexpr size_t smem_bytes()
{
    return sizeof(float2) * ((size_t)(kThreads / (M / 2)) * Layout<M>::NBUF * kFr * Layout<M>::FS + (size_t)(M / 32) * kTwPitch);
}

// host: tw2[u] = e^{sign j 2 pi u / M}, u < M/2; then twp[k1][r2] = e^{sign j 2 pi k1 r2 / (M/2)}, k1 < 16, r2 < M/32
inline void make_twiddles(std::vector<float2> &tw, unsigned int M, int sign)
{
    const unsigned int M2 = M / 2, R2 = M2 / 16;
    tw.resize(M2 + 16 * R2);
    for (unsigned int u = 0; u < M2; u++) {
        const double ang = (double)sign * 2.0 * M_PI * (double)u / (double)M;
        tw[u] = make_float2((float)cos(ang), (float)sin(ang));
    }
    for (unsigned int k1 = 0; k1 < 16; k1++)
        for (unsigned int r2 = 0; r2 < R2; r2++) {
            const double ang = (double)sign * 2.0 * M_PI * (double)(k1 * r2) / (double)M2;
            tw[M2 + k1 * R2 + r2] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

// grid for n transforms: instances of `tiles_per_inst` consecutive tiles
inline void plan(unsigned int M, unsigned long long n, unsigned int &grid, int &tiles_per_inst)
{
    const long long n_tiles = (long long)((n + kFr - 1) / kFr);
    const long long inst_per_cta = kThreads / (M / 2);
    long long tpi = 16;
    while (tpi > 1 && (n_tiles / tpi) < 148ll * 4 * inst_per_cta) tpi >>= 1;
    tiles_per_inst = (int)tpi;
    const long long n_inst = (n_tiles + tpi - 1) / tpi;
    grid = (unsigned int)((n_inst + inst_per_cta - 1) / inst_per_cta);
}

#ifdef __CUDACC__
// the whole kernel body: call from a __global__ function launched with kThreads threads and smem_bytes<M>() of dynamic
// shared memory (passed in as smem_raw)
template <int M, int SIGN, class Src, class Dst>
__device__ __forceinline__ void run(unsigned char *smem_raw, const float2 *__restrict__ tw, long long n_tr, int tiles_per_inst,
                                    Src src, Dst &&dst)
{
    using LY = Layout<M>;
    constexpr int M2 = M / 2, R2 = M2 / 16;
    constexpr int PR = LY::PR, SS = LY::SS, FS = LY::FS, NBUF = LY::NBUF;
    constexpr int RPT = 16 / R2;
    const int inst_l = threadIdx.x / M2, u = threadIdx.x % M2;
    float2 *tile0 = reinterpret_cast<float2 *>(smem_raw) + inst_l * (NBUF * kFr * FS);
    float2 *tws = reinterpret_cast<float2 *>(smem_raw) + (kThreads / M2) * (NBUF * kFr * FS);
    const long long inst = (long long)blockIdx.x * (kThreads / M2) + inst_l;
    const long long n_tiles = (n_tr + kFr - 1) / kFr;
    const long long t0 = inst * tiles_per_inst;
    long long ntl = n_tiles - t0;
    if (ntl > tiles_per_inst) ntl = tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;

    auto inst_barrier = [&]() {
        if (M2 == kThreads) __syncthreads();
        else if (M2 == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + inst_l), "n"(M2) : "memory");
    };
    const int f2 = u / (2 * R2), p2 = u & 1, t2 = (u % (2 * R2)) >> 1;
    float2 nx[16];
    auto fetch = [&](int it) {
        const long long tr0 = (t0 + it) * kFr;
#pragma unroll
        for (int i = 0; i < (LY::WIDE ? 32 * kFr / M2 : RPT); i++) {
            int f, off;
            if (LY::WIDE) { const int j = u + M2 * i; f = j >> 5; off = j & 31; }
            else          { f = f2; off = 2 * (t2 + R2 * i) + p2; }
            const bool ok = it < nt && tr0 + f < n_tr;
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) nx[i * R2 + k2] = ok ? src.load(tr0 + f, (unsigned int)(off + 32 * k2)) : make_float2(0.f, 0.f);
        }
    };
    fetch(0);
    for (int i = u; i < 16 * R2; i += M2) tws[(i % R2) * kTwPitch + i / R2] = __ldg(tw + M2 + i);
    const float2 w2 = __ldg(tw + u);
    const float4 *twc = reinterpret_cast<const float4 *>(tws + t2 * kTwPitch);
    const int mac_off = (u / R2) * PR + (u % R2);
    const int sp_off = f2 * FS + p2 * SS;
    __syncthreads();

    for (int it = 0; it < nt; it++) {
        const long long tr0 = (t0 + it) * kFr;
        float2 *tile = tile0 + (it % NBUF) * (kFr * FS);
        if (LY::WIDE) {
#pragma unroll
            for (int i = 0; i < 32 * kFr / M2; i++) {
                const int j = u + M2 * i;
                float2 r[R2];
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++) r[k2] = src.finish(tr0 + (j >> 5), (unsigned int)((j & 31) + 32 * k2), nx[i * R2 + k2]);
                RegDft<R2, SIGN>::run(r);
                float2 *row = tile + (j >> 5) * FS + (j & 1) * SS + ((j & 31) >> 1) * PR;
#pragma unroll
                for (int c = 0; c < R2; c++) row[c] = r[c];
            }
        } else {
#pragma unroll
            for (int i = 0; i < RPT; i++) {
                float2 r[R2];
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++)
                    r[k2] = src.finish(tr0 + f2, (unsigned int)(2 * (t2 + R2 * i) + p2 + 32 * k2), nx[i * R2 + k2]);
                RegDft<R2, SIGN>::run(r);
                float2 *row = tile + sp_off + (t2 + R2 * i) * PR;
#pragma unroll
                for (int c = 0; c < R2; c++) row[c] = r[c];
            }
        }
        fetch(it + 1);
        if (LY::WIDE) inst_barrier();
        else __syncwarp();
        {
            float2 *sp = tile + sp_off;
            float2 v[16];
#pragma unroll
            for (int k1 = 0; k1 < 16; k1++) v[k1] = sp[k1 * PR + t2];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 t = twc[k];
                if (k > 0) v[2 * k] = cmul2_rot(v[2 * k], make_float2(t.x, t.y));
                v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
            }
            RegDft<16, SIGN>::run(v);
#pragma unroll
            for (int r1 = 0; r1 < 16; r1++) sp[r1 * PR + t2] = v[r1];
        }
        inst_barrier();
        const float2 *srcp = tile + mac_off;
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            if (tr0 + f < n_tr) {
                const float2 e0 = srcp[f * FS];
                const float2 t = cmul2_rot(srcp[f * FS + SS], w2);
                dst(tr0 + f, (unsigned int)u, __fadd2_rn(e0, t));
                dst(tr0 + f, (unsigned int)(u + M2), __fadd2_rn(e0, neg2(t)));
            }
        }
        if (NBUF == 1) inst_barrier();
    }
}

// Forward transform, product with a frequency response, backward transform in one kernel (fftfilt_crcf, 2n = 64 ... 512).
// The forward transform is run<M, -1> above with a destination that keeps out[u] and out[u + M/2] in the owning thread;
// the backward transform is its TRANSPOSE (decimation in frequency), which starts exactly from that pair:
//   a[u] = X[u] + X[u + M/2],  b[u] = (X[u] - X[u + M/2]) W_M^(+u)           in the owning thread, written to the tile
//   16-point DFTs down the columns (u = 16-row index * R2 + column), twiddle W_(M/2)^(column k1), in place
//   R2-point DFTs along the rows -> x[2 (k1 + 16 k2) + p]
// so no exchange is needed between the two transforms.  tw_f / tw_b: forward / backward tables of make_twiddles, Hf[M]
// natural order.  UPPER: only outputs M/2 ... M-1 go to dst.  Same tile buffers as run(): the backward transform of
// tile `it` works in the buffer the forward transform just left (its reads are complete after the instance barrier).
template <int M, bool UPPER, class Src, class Dst>
__device__ __forceinline__ void run_chain(unsigned char *smem_raw, const float2 *__restrict__ tw_f, const float2 *__restrict__ tw_b,
                                          const float2 *__restrict__ Hf, long long n_tr, int tiles_per_inst, Src src, Dst &&dst)
{
    using LY = Layout<M>;
    constexpr int M2 = M / 2, R2 = M2 / 16;
    constexpr int PR = LY::PR, SS = LY::SS, FS = LY::FS, NBUF = LY::NBUF;
    constexpr int RPT = 16 / R2;
    static_assert(!LY::WIDE, "the chained form is written for M = 256 / 512");
    const int inst_l = threadIdx.x / M2, u = threadIdx.x % M2;
    float2 *tile0 = reinterpret_cast<float2 *>(smem_raw) + inst_l * (NBUF * kFr * FS);
    float2 *twsf = reinterpret_cast<float2 *>(smem_raw) + (kThreads / M2) * (NBUF * kFr * FS);
    float2 *twsb = twsf + R2 * kTwPitch;
    const long long inst = (long long)blockIdx.x * (kThreads / M2) + inst_l;
    const long long n_tiles = (n_tr + kFr - 1) / kFr;
    const long long t0 = inst * tiles_per_inst;
    long long ntl = n_tiles - t0;
    if (ntl > tiles_per_inst) ntl = tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;

    auto inst_barrier = [&]() {
        if (M2 == kThreads) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + inst_l), "n"(M2) : "memory");
    };
    const int f2 = u / (2 * R2), p2 = u & 1, t2 = (u % (2 * R2)) >> 1;
    float2 nx[16];
    auto fetch = [&](int it) {
        const long long tr0 = (t0 + it) * kFr;
        const bool ok = it < nt && tr0 + f2 < n_tr;
#pragma unroll
        for (int i = 0; i < RPT; i++)
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++)
                nx[i * R2 + k2] = ok ? src.load(tr0 + f2, (unsigned int)(2 * (t2 + R2 * i) + p2 + 32 * k2)) : make_float2(0.f, 0.f);
    };
    fetch(0);
    for (int i = u; i < 16 * R2; i += M2) {
        twsf[(i % R2) * kTwPitch + i / R2] = __ldg(tw_f + M2 + i);
        twsb[(i % R2) * kTwPitch + i / R2] = __ldg(tw_b + M2 + i);
    }
    const float2 w2f = __ldg(tw_f + u), w2b = __ldg(tw_b + u);
    const float2 h_lo = __ldg(Hf + u), h_hi = __ldg(Hf + u + M2);
    const float4 *twcf = reinterpret_cast<const float4 *>(twsf + t2 * kTwPitch);
    const float4 *twcb = reinterpret_cast<const float4 *>(twsb + t2 * kTwPitch);
    const int mac_off = (u / R2) * PR + (u % R2);
    const int sp_off = f2 * FS + p2 * SS;
    __syncthreads();

    for (int it = 0; it < nt; it++) {
        const long long tr0 = (t0 + it) * kFr;
        float2 *tile = tile0 + (it % NBUF) * (kFr * FS);
        float2 *sp = tile + sp_off;
        // ---- forward: rows, columns (as run<M, -1>)
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            float2 r[R2];
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) r[k2] = src.finish(tr0 + f2, (unsigned int)(2 * (t2 + R2 * i) + p2 + 32 * k2), nx[i * R2 + k2]);
            RegDft<R2, -1>::run(r);
            float2 *row = sp + (t2 + R2 * i) * PR;
#pragma unroll
            for (int c = 0; c < R2; c++) row[c] = r[c];
        }
        fetch(it + 1);
        __syncwarp();
        {
            float2 v[16];
#pragma unroll
            for (int k1 = 0; k1 < 16; k1++) v[k1] = sp[k1 * PR + t2];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 t = twcf[k];
                if (k > 0) v[2 * k] = cmul2_rot(v[2 * k], make_float2(t.x, t.y));
                v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
            }
            RegDft<16, -1>::run(v);
#pragma unroll
            for (int r1 = 0; r1 < 16; r1++) sp[r1 * PR + t2] = v[r1];
        }
        inst_barrier();
        // ---- last forward stage, product with H, first backward stage: all in the thread that owns u and u + M/2
        float2 *own = tile + mac_off;
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            const float2 e0 = own[f * FS];
            const float2 t = cmul2_rot(own[f * FS + SS], w2f);
            const float2 xl = cmul2_rot(__fadd2_rn(e0, t), h_lo);
            const float2 xh = cmul2_rot(__fadd2_rn(e0, neg2(t)), h_hi);
            own[f * FS] = __fadd2_rn(xl, xh);
            own[f * FS + SS] = cmul2_rot(__fadd2_rn(xl, neg2(xh)), w2b);
        }
        inst_barrier();
        // ---- backward: columns (DFT, then twiddle), rows
        {
            float2 v[16];
#pragma unroll
            for (int r1 = 0; r1 < 16; r1++) v[r1] = sp[r1 * PR + t2];
            RegDft<16, +1>::run(v);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 t = twcb[k];
                if (k > 0) v[2 * k] = cmul2_rot(v[2 * k], make_float2(t.x, t.y));
                v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
            }
#pragma unroll
            for (int k1 = 0; k1 < 16; k1++) sp[k1 * PR + t2] = v[k1];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            float2 r[R2];
            const float2 *row = sp + (t2 + R2 * i) * PR;
#pragma unroll
            for (int c = 0; c < R2; c++) r[c] = row[c];
            RegDft<R2, +1>::run(r);
            if (tr0 + f2 < n_tr) {
#pragma unroll
                for (int k2 = UPPER ? R2 / 2 : 0; k2 < R2; k2++) dst(tr0 + f2, (unsigned int)(2 * (t2 + R2 * i + 16 * k2) + p2), r[k2]);
            }
        }
        if (NBUF == 1) inst_barrier();
    }
}
template <int M>
constexpr size_t chain_smem_bytes()
{
    return sizeof(float2) * ((size_t)(kThreads / (M / 2)) * Layout<M>::NBUF * kFr * Layout<M>::FS + (size_t)2 * (M / 32) * kTwPitch);
}
#endif

}  // namespace lctile
