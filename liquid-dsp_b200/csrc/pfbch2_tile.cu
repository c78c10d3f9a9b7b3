// pfbch2_tile.cu -- fused firpfbch2_crcf analysis kernels for M = 64, 128, 256, 512 (sm_100a): the shapes of the
// reference's own benchmark (src/multichannel/bench/firpfbch2_crcf_benchmark.c:70-82) below the M = 1024 kernel.
//
// These sizes are HBM-bound (24 B per input sample against 16m + 10 log2 M flops), so the kernel reads every
// sample once, writes every output once and keeps the transform input in shared memory:
//
//   phase 1  thread u (< M/2) owns the window pair (u, u + M/2).  The pair receives ONE stream sample per frame,
//            s_T = x[T M/2 + M/2-1-u], alternately into its two windows, so with the 2m taps of the two branches
//            (hA = branch u, hB = branch u + M/2) and a register window of the last 4m-1 + 8 samples
//                d1 = sum_n hA[n] s_{T-2n},   d2 = sum_n hB[n] s_{T-1-2n}
//            are the pair's two dot products of frame T -- (X[u], X[u+M/2]) = (d1, d2) on even frames, (d2, d1) on
//            odd ones (firpfbch2.c:253-271).  The thread also does the first radix-2 decimation-in-frequency stage
//            of the M-point backward transform on its own two values (z0 = X[u] + X[u+M/2], z1 = (X[u] - X[u+M/2])
//            W_M^u), which leaves two independent M/2-point problems per frame in the shared tile (8 frames).
//   phase 2  M/16 threads per frame (all M/2 threads busy on a tile of 8 frames): each M/2-point problem is
//            16 x R2 (R2 = M/32): a 16-point register DFT down a column, the inter-pass twiddle, an in-place
//            transposition (rows padded to R2 + 1: every access conflict-free), R2-point register DFTs along the
//            rows, and the outputs Y[2 (k1 + 16 k2) + p] leave as 32-byte (M = 64) to 256-byte (M = 512) runs.
//   An instance (M/2 threads; 256 / (M/2) instances per CTA) walks consecutive tiles carrying its window in registers;
//   the 1/M output scale is folded into the taps (exact: M is a power of two).
//
// Reference semantics: src/multichannel/src/firpfbch2.c:244-282 (see firpfbch2.cu).
#include "pfbch2_tile.cuh"
#include "fused_common.cuh"

#include <cmath>

namespace {

using namespace lcf;

constexpr int kThreads = 256;
constexpr int kFr = 8;                          // frames per tile
constexpr int kTwPitch = 18;                    // float2 per column of the shared twiddle table (144 B: 16-byte reads, conflict-free)

// Tile layout (8-byte units): address = f*FS + p*SS + row*PR + col for frame f, sub-problem p, row < 16, col < R2.
// (PR, SS, FS) come from tools/tile_layout_search.py: with them every shared-memory access pattern of the kernel
// (phase-1 stores, pass-A column reads / writes, pass-B row reads) costs the minimum number of wavefronts, or
// 4 % more (M <= 256); rows padded to R2 + 1 alone leave the sub-problems and frames a multiple of 16 units
// apart, which put the lanes of a warp on 2 - 8 banks (776 instead of 100 wavefronts per tile at M = 64).
// WIDE: pass B takes its rows in the order (frame, row, sub-problem) across the whole instance, so that a warp
// store covers 256 contiguous bytes of one output frame whatever M is (grouped order: 16 R2 bytes).
template <int M> struct TileLayout;
// NBUF: tile buffers per instance.  Two buffers leave ONE instance barrier per tile (phase 1 of the next tile writes
// the other buffer while slower warps are still in pass B); an instance that is a single warp (M = 64) needs none.
// STAGES: input slots per thread: 1 = the next tile's 8 samples are in flight from the end of the MAC phase on,
// 2 = a whole tile ahead.
template <> struct TileLayout<64>  { static constexpr int PR = 3,  SS = 56,  FS = 106, WIDE = 1, NBUF = 1, STAGES = 2; };
template <> struct TileLayout<128> { static constexpr int PR = 5,  SS = 88,  FS = 172, WIDE = 1, NBUF = 2, STAGES = 1; };
template <> struct TileLayout<256> { static constexpr int PR = 9,  SS = 152, FS = 295, WIDE = 0, NBUF = 2, STAGES = 1; };
template <> struct TileLayout<512> { static constexpr int PR = 17, SS = 280, FS = 551, WIDE = 0, NBUF = 2, STAGES = 1; };

struct TileArgs {
    const float2 *hist;        // H samples preceding x[0], or null = zero state
    const float2 *x;
    float2 *y;                 // [n_frames][M]
    const float *taps;         // [2][L][M] analyzer table (firpfbch2.cu), pre-scaled by 1/M; parity 0 half is used
    const float2 *tw;          // tw2[u] = e^{+j 2 pi u / M}, u < M/2; then twp[k1][r2] = e^{+j 2 pi k1 r2 / (M/2)}
    long long n_frames;
    int H;
    int par0;                  // parity of the first frame
    int n_tiles;               // ceil(n_frames / 8): the last tile may be ragged
    int tiles_per_inst;
};

__device__ __forceinline__ float2 ld_hist0(const float2 *__restrict__ hist, const float2 *__restrict__ x, long long idx, int H)
{
    if (idx >= 0) return __ldg(x + idx);
    if (hist != nullptr && H + idx >= 0) return __ldg(hist + (H + idx));
    return make_float2(0.f, 0.f);
}

// 8-byte asynchronous copy global -> shared; n = 0 writes zeros without touching src
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src, uint32_t n)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// analyzers: the next tile's samples are prefetched into registers (true) or into a cp.async slot in shared memory
template <int L> constexpr bool kRegPrefetch = (L <= 4);           // measured: m >= 4 loses 3 - 8 % to register pressure

template <int M, int L>
constexpr size_t tile_smem_bytes()
{
    using LY = TileLayout<M>;
    return sizeof(float2) * ((kRegPrefetch<L> ? (size_t)0 : (size_t)LY::STAGES * kFr * kThreads)
                             + (size_t)(kThreads / (M / 2)) * LY::NBUF * kFr * LY::FS + (size_t)(M / 32) * kTwPitch);
}

// Phase 2 of the analyzers for one tile: the 2 x 8 sub-problems of M/2 = 16 x R2 points in `tile` (phase-1 layout) are
// transformed (e^{SIGN j ...}) and the frames leave for y_str[(f0 + f) M + 2 (k1 + 16 k2) + p].  Called by the whole
// instance after the barrier that follows phase 1.
template <int M, int SIGN, class Barrier>
__device__ __forceinline__ void tile_phase2(float2 *tile, int u, const float4 *twc, float2 *y_str, long long f0,
                                            long long n_frames, Barrier inst_barrier)
{
    using LY = TileLayout<M>;
    constexpr int M2 = M / 2, R2 = M2 / 16;
    constexpr int PR = LY::PR, SS = LY::SS, FS = LY::FS;
    // pass-A role inside the instance: frame f2, sub-problem p2 (fastest), column t2
    const int f2 = u / (2 * R2), p2 = u & 1, t2 = (u % (2 * R2)) >> 1;
    // pass A down column t2 (in place: a thread reads and writes its own column)
    float2 *sp = tile + f2 * FS + p2 * SS;
    float2 v[16];
#pragma unroll
    for (int r1 = 0; r1 < 16; r1++) v[r1] = sp[r1 * PR + t2];
    RegDft<16, SIGN>::run(v);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float4 t = twc[k];                              // twiddles k1 = 2k, 2k + 1
        if (k > 0) v[2 * k] = cmul2_rot(v[2 * k], make_float2(t.x, t.y));
        v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
    }
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) sp[k1 * PR + t2] = v[k1];
    if (LY::WIDE) {
        // pass B over the whole instance: item j = (frame, row k1, sub-problem), sub-problem fastest
        inst_barrier();
#pragma unroll
        for (int i = 0; i < 32 * kFr / M2; i++) {
            const int j = u + M2 * i;
            const int f = j >> 5, k1 = (j & 31) >> 1;
            const float2 *row = tile + f * FS + (j & 1) * SS + k1 * PR;
            float2 r[R2];
#pragma unroll
            for (int c = 0; c < R2; c++) r[c] = row[c];
            RegDft<R2, SIGN>::run(r);
            if (f0 + f < n_frames) {
                float2 *yf = y_str + (f0 + f) * M + (j & 31);
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++) __stcs(yf + 32 * k2, r[k2]);
            }
        }
    } else {
        // pass B inside the 2 R2 lanes of the sub-problem: rows k1 = t2 + R2 i
        __syncwarp();
        float2 *yf = y_str + (f0 + f2) * M + p2;
        const bool live = f0 + f2 < n_frames;
#pragma unroll
        for (int i = 0; i < 16 / R2; i++) {
            const int k1 = t2 + R2 * i;
            float2 r[R2];
#pragma unroll
            for (int c = 0; c < R2; c++) r[c] = sp[k1 * PR + c];
            RegDft<R2, SIGN>::run(r);
            if (live) {
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++) __stcs(yf + 2 * (k1 + 16 * k2), r[k2]);
            }
        }
    }
}

template <int M, int L>
__global__ void __launch_bounds__(kThreads, (L <= 10 ? 2 : 1)) pfbch2_tile_kernel(const TileArgs a)
{
    using LY = TileLayout<M>;
    constexpr int M2 = M / 2;                   // threads per instance = window pairs = sub-problem size
    constexpr int R2 = M2 / 16;                 // second factor of the M/2-point transform
    constexpr int PR = LY::PR, SS = LY::SS, FS = LY::FS, kStages = LY::STAGES, NBUF = LY::NBUF;
    constexpr int HL = 2 * L - 1;               // frames of history a frame looks back at
    static_assert(R2 >= 2 && R2 <= 16, "M must be 64, 128, 256 or 512");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *ring = reinterpret_cast<float2 *>(smem_raw);          // [kStages][kFr][kThreads]: a thread's own samples

    const int inst_l = threadIdx.x / M2, u = threadIdx.x % M2;
    constexpr int kRingUnits = kRegPrefetch<L> ? 0 : kStages * kFr * kThreads;
    float2 *tile0 = ring + kRingUnits + inst_l * (NBUF * kFr * FS);
    float2 *tws = ring + kRingUnits + (kThreads / M2) * (NBUF * kFr * FS); // [R2][kTwPitch]: column t2's 16 twiddles
    const long long inst = (long long)blockIdx.x * (kThreads / M2) + inst_l;
    const long long t0 = inst * a.tiles_per_inst;
    long long ntl = (long long)a.n_tiles - t0;
    if (ntl > a.tiles_per_inst) ntl = a.tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;                        // uniform within an instance (a set of whole warps)

    auto inst_barrier = [&]() {
        if (M2 == kThreads) __syncthreads();
        else if (M2 == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + inst_l), "n"(M2) : "memory");
    };
    // tile `it` of this instance -> the thread's own input slot (8 bytes per frame; frames past the end of the
    // block, and tiles past the instance's range, arrive as zeros)
    const uint32_t ring_u32 = smem_u32(ring + threadIdx.x);
    const float2 *xs = a.x + (M2 - 1 - u);
    // REGPF: the next tile's 8 samples are prefetched into registers instead (no shared-memory round trip); taken
    // where the register budget allows it
    constexpr bool REGPF = kRegPrefetch<L>;
    float2 nx[kFr];
    auto fetch = [&](int it) {
        const long long f0 = (t0 + it) * kFr;
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            const bool ok = it < nt && f0 + f < a.n_frames;
            if (REGPF) nx[f] = ok ? __ldg(xs + (f0 + f) * M2) : make_float2(0.f, 0.f);
            else cp_async8(ring_u32 + ((it % kStages) * kFr + f) * kThreads * (int)sizeof(float2), ok ? (const void *)(xs + (f0 + f) * M2) : (const void *)a.x, ok ? 8u : 0u);
        }
        if (!REGPF) cp_async_commit();
    };
    fetch(0);
    // the instance's copy of the inter-pass twiddles, by column: tws[t2][k1] = e^{+j 2 pi k1 t2 / (M/2)}
    for (int i = u; i < 16 * R2; i += M2) tws[(i % R2) * kTwPitch + i / R2] = __ldg(a.tw + M2 + i);

    float hA[L], hB[L];
#pragma unroll
    for (int n = 0; n < L; n++) {
        hA[n] = __ldg(a.taps + n * M + u);
        hB[n] = __ldg(a.taps + n * M + u + M2);
    }
    const float2 w2 = __ldg(a.tw + u);                            // W_M^u
    float2 win[HL + kFr];
    if (nt > 0) {
        const long long f0 = t0 * kFr;
#pragma unroll
        for (int i = 0; i < HL; i++) win[i] = ld_hist0(a.hist, a.x, (f0 - HL + i) * M2 + (M2 - 1 - u), a.H);
    }
    const float4 *twc = reinterpret_cast<const float4 *>(tws + ((u % (2 * R2)) >> 1) * kTwPitch);   // pass-A column of this thread
    const int mac_off = (u / R2) * PR + (u % R2);                 // where z_p[u] lives: row u / R2, column u % R2
    __syncthreads();                                              // the twiddle copies of all instances

    for (int it = 0; it < nt; it++) {
        const long long f0 = (t0 + it) * kFr;                     // first (local) frame of the tile: even
        float2 *tile = tile0 + (it % NBUF) * (kFr * FS);
        if (REGPF) {
#pragma unroll
            for (int f = 0; f < kFr; f++) win[HL + f] = nx[f];
            fetch(it + 1);
        } else {
            if (kStages == 2) fetch(it + 1);
            cp_async_wait<kStages - 1>();
            const float2 *src = ring + (it % kStages) * kFr * kThreads + threadIdx.x;
#pragma unroll
            for (int f = 0; f < kFr; f++) win[HL + f] = src[f * kThreads];
        }
        // ---- phase 1
        float2 *mac_dst = tile + mac_off;
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            float2 d1 = make_float2(0.f, 0.f), d2 = d1;
#pragma unroll
            for (int n = L - 1; n >= 0; n--) {
                d1 = __ffma2_rn(bc(hA[n]), win[HL + f - 2 * n], d1);
                d2 = __ffma2_rn(bc(hB[n]), win[HL + f - 1 - 2 * n], d2);
            }
            // radix-2 DIF stage: z0 = X[u] + X[u+M/2] = d1 + d2 on every frame; z1 = +-(d1 - d2) W_M^u (+ on even frames)
            const float2 z0 = __fadd2_rn(d1, d2);
            const bool odd = ((a.par0 + f) & 1) != 0;             // f0 is even
            const float2 df = odd ? __fadd2_rn(d2, neg2(d1)) : __fadd2_rn(d1, neg2(d2));
            mac_dst[f * FS] = z0;
            mac_dst[f * FS + SS] = cmul2_rot(df, w2);
        }
        if (!REGPF && kStages == 1) fetch(it + 1);                // the slot's values are consumed: refill it
#pragma unroll
        for (int i = 0; i < HL; i++) win[i] = win[kFr + i];
        // with two buffers the one barrier of a tile: the instance's phase-1 stores are visible, and every warp has
        // left pass B of the tile before (which read the other buffer -- the one phase 1 of the next tile writes)
        inst_barrier();
        // ---- phase 2
        tile_phase2<M, +1>(tile, u, twc, a.y, f0, a.n_frames, inst_barrier);
        if (NBUF == 1) inst_barrier();
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------- firpfbch (critically sampled) analyzer
// firpfbch.c:369-409 (M pushes then run, alignment 0): X_s[u] = sum_n h[(M-1-u) + nM] x[(s-n)M + u], Y_s = forward
// DFT.  Same kernel shape as above: thread u owns the windows u and u + M/2 (two samples per frame, two P-tap dot
// products over sliding register windows), does the first radix-2 decimation-in-frequency stage on its two values,
// and the instance finishes the two M/2-point problems per frame in the shared tile (tile_phase2 with the forward
// sign).  grid.y = stream (independent streams from the zero state, or one stateful stream with its history).
struct PfbTileArgs {
    const float2 *hist;        // H samples preceding x[0] of the (single) stream, or null = zero state
    const float2 *x;           // [n_streams][stream_stride]
    float2 *y;
    const float *taps;         // [P][M]: taps[n][i] = h[i + nM]
    const float2 *tw;          // tw2[u] = e^{-j 2 pi u / M}, u < M/2; then twp[k1][r2] = e^{-j 2 pi k1 r2 / (M/2)}
    long long n_frames;
    long long stream_stride;
    int H;
    int n_tiles;
    int tiles_per_inst;
};

// register prefetch of the next tile's 16 samples where it measured faster (short filters; M = 512 at every length),
// a cp.async slot in shared memory elsewhere (M = 64, p = 10: 0.93 against 0.75 with 160 B of spills)
template <int M, int P> constexpr bool kPfbRegPrefetch = (P <= 4) || (M >= 512);

template <int M, int P, typename TapT>
constexpr size_t pfb_tile_smem_bytes()
{
    using LY = TileLayout<M>;
    return sizeof(float2) * (((kPfbRegPrefetch<M, P> || sizeof(TapT) == 8) ? (size_t)0 : (size_t)2 * kFr * kThreads)
                             + (size_t)(kThreads / (M / 2)) * LY::NBUF * kFr * LY::FS + (size_t)(M / 32) * kTwPitch);
}

// tap types: float (firpfbch_crcf) or float2 (firpfbch_cccf: plain complex product, two packed FFMA2 per tap)
__device__ __forceinline__ float2 tap_mac(float h, float2 s, float2 acc) { return __ffma2_rn(bc(h), s, acc); }
__device__ __forceinline__ float2 tap_mac(float2 h, float2 s, float2 acc)
{
    return __ffma2_rn(make_float2(-h.y, h.y), make_float2(s.y, s.x), __ffma2_rn(bc(h.x), s, acc));
}

template <int M, int P, typename TapT>
__global__ void __launch_bounds__(kThreads, (sizeof(TapT) == 4 || P <= 4) ? 2 : 1) pfbch_tile_kernel(const PfbTileArgs a)
{
    using LY = TileLayout<M>;
    constexpr int M2 = M / 2, R2 = M2 / 16;
    constexpr int PR = LY::PR, SS = LY::SS, FS = LY::FS, NBUF = LY::NBUF;
    constexpr int HL = P - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool REGPF = kPfbRegPrefetch<M, P> || sizeof(TapT) == 8;
    constexpr int kSlotUnits = REGPF ? 0 : 2 * kFr * kThreads;
    float2 *slot = reinterpret_cast<float2 *>(smem_raw);          // [2 kFr][kThreads]: a thread's own 16 samples of a tile
    const int inst_l = threadIdx.x / M2, u = threadIdx.x % M2;
    float2 *tile0 = slot + kSlotUnits + inst_l * (NBUF * kFr * FS);
    float2 *tws = slot + kSlotUnits + (kThreads / M2) * (NBUF * kFr * FS);
    const long long inst = (long long)blockIdx.x * (kThreads / M2) + inst_l;
    const long long t0 = inst * a.tiles_per_inst;
    long long ntl = (long long)a.n_tiles - t0;
    if (ntl > a.tiles_per_inst) ntl = a.tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;
    const float2 *x = a.x + (long long)blockIdx.y * a.stream_stride;
    float2 *y = a.y + (long long)blockIdx.y * a.stream_stride;

    auto inst_barrier = [&]() {
        if (M2 == kThreads) __syncthreads();
        else if (M2 == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + inst_l), "n"(M2) : "memory");
    };
    // the thread's 16 samples of the next tile: into registers, or into its cp.async slot
    float2 nx0[kFr], nx1[kFr];
    const uint32_t slot_u32 = smem_u32(slot + threadIdx.x);
    auto fetch = [&](int it) {
        const long long f0 = (t0 + it) * kFr;
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            const bool ok = it < nt && f0 + f < a.n_frames;
            const float2 *src = x + (f0 + f) * M + u;
            if (REGPF) {
                nx0[f] = ok ? __ldg(src) : make_float2(0.f, 0.f);
                nx1[f] = ok ? __ldg(src + M2) : make_float2(0.f, 0.f);
            } else {
                cp_async8(slot_u32 + (2 * f) * kThreads * (int)sizeof(float2), ok ? (const void *)src : (const void *)a.x, ok ? 8u : 0u);
                cp_async8(slot_u32 + (2 * f + 1) * kThreads * (int)sizeof(float2), ok ? (const void *)(src + M2) : (const void *)a.x, ok ? 8u : 0u);
            }
        }
        if (!REGPF) cp_async_commit();
    };
    fetch(0);
    for (int i = u; i < 16 * R2; i += M2) tws[(i % R2) * kTwPitch + i / R2] = __ldg(a.tw + M2 + i);

    const TapT *taps = reinterpret_cast<const TapT *>(a.taps);
    TapT h0[P], h1[P];
#pragma unroll
    for (int n = 0; n < P; n++) {
        h0[n] = __ldg(taps + n * M + (M - 1 - u));                // branch M-1-u feeds input u
        h1[n] = __ldg(taps + n * M + (M2 - 1 - u));               // branch M-1-(u + M/2)
    }
    const float2 w2 = __ldg(a.tw + u);                            // W_M^u, forward
    float2 w0[HL + kFr], w1[HL + kFr];
    if (nt > 0) {
        const long long f0 = t0 * kFr;
#pragma unroll
        for (int i = 0; i < HL; i++) {
            w0[i] = ld_hist0(a.hist, x, (f0 - HL + i) * M + u, a.H);
            w1[i] = ld_hist0(a.hist, x, (f0 - HL + i) * M + u + M2, a.H);
        }
    }
    const float4 *twc = reinterpret_cast<const float4 *>(tws + ((u % (2 * R2)) >> 1) * kTwPitch);
    const int mac_off = (u / R2) * PR + (u % R2);
    __syncthreads();

    for (int it = 0; it < nt; it++) {
        const long long f0 = (t0 + it) * kFr;
        float2 *tile = tile0 + (it % NBUF) * (kFr * FS);
        if (REGPF) {
#pragma unroll
            for (int f = 0; f < kFr; f++) {
                w0[HL + f] = nx0[f];
                w1[HL + f] = nx1[f];
            }
            fetch(it + 1);
        } else {
            cp_async_wait<0>();
            const float2 *src = slot + threadIdx.x;
#pragma unroll
            for (int f = 0; f < kFr; f++) {
                w0[HL + f] = src[(2 * f) * kThreads];
                w1[HL + f] = src[(2 * f + 1) * kThreads];
            }
        }
        float2 *mac_dst = tile + mac_off;
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            float2 d0 = make_float2(0.f, 0.f), d1 = d0;
#pragma unroll
            for (int n = P - 1; n >= 0; n--) {
                d0 = tap_mac(h0[n], w0[HL + f - n], d0);
                d1 = tap_mac(h1[n], w1[HL + f - n], d1);
            }
            mac_dst[f * FS] = __fadd2_rn(d0, d1);
            mac_dst[f * FS + SS] = cmul2_rot(__fadd2_rn(d0, neg2(d1)), w2);
        }
        if (!REGPF) fetch(it + 1);                                // the slot's values are consumed: refill it
#pragma unroll
        for (int i = 0; i < HL; i++) {
            w0[i] = w0[kFr + i];
            w1[i] = w1[kFr + i];
        }
        inst_barrier();
        tile_phase2<M, -1>(tile, u, twc, y, f0, a.n_frames, inst_barrier);
        if (NBUF == 1) inst_barrier();
    }
    if (!REGPF) cp_async_wait<0>();
}

template <int M, int P, typename TapT>
void launch_pfb_MP(const PfbTileArgs &a, dim3 grid, cudaStream_t st)
{
    constexpr size_t smem = pfb_tile_smem_bytes<M, P, TapT>();
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch_tile_kernel<M, P, TapT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pfbch_tile_kernel<M, P, TapT><<<grid, kThreads, smem, st>>>(a);
}

template <int M, typename TapT>
void launch_pfb_M(const PfbTileArgs &a, unsigned int P, dim3 grid, cudaStream_t st)
{
    switch (P) {
    case 4: launch_pfb_MP<M, 4, TapT>(a, grid, st); break;
    case 6: launch_pfb_MP<M, 6, TapT>(a, grid, st); break;
    case 8: launch_pfb_MP<M, 8, TapT>(a, grid, st); break;
    case 10: launch_pfb_MP<M, 10, TapT>(a, grid, st); break;
    case 12: launch_pfb_MP<M, 12, TapT>(a, grid, st); break;
    case 14: launch_pfb_MP<M, 14, TapT>(a, grid, st); break;
    default: lc::fatal("launch_pfbch_tile", "unsupported tap count");
    }
}

// ---------------------------------------------------------------------------------- synthesizer
// The synthesizer (firpfbch2.c:287-335) is the transposed flow graph of the analyzer, and so is its kernel:
//   pass B'  R2-point register DFTs over the rows X_p[k1 + 16 k2] (prefetched into registers one tile ahead),
//   pass A'  the inter-pass twiddle (times the reference's 1/M * M/2 = 1/2, exact) and a 16-point register DFT down
//            each column: the tile then holds E_p[u] = sum_k X[2k + p] e^{+j 2 pi u k / (M/2)} / 2,
//   FIR      thread u does the last radix-2 stage, v[u] = E_0[u] + W_M^u E_1[u], v[u + M/2] = E_0[u] - W_M^u E_1[u],
//            pushes the two values into register windows of the last 4m rows of v in its two columns and forms the
//            output y[t][u] = sum_q h[u + q M/2] v[t - q][u + (t odd ? M/2 : 0)] with packed FFMA2.
// The rows of v never reach HBM (the two-kernel path writes and re-reads 16 B per output sample).  An instance that
// starts inside the stream first transforms the ceil((4m-1)/8) tiles before its range to rebuild its windows; the
// last 4m rows of v of a call are written out as the object's state.
struct SynthArgs {
    const float2 *hist;        // hist_rows rows of v preceding frame 0, or null = zero state
    const float2 *x;           // [n_frames][M] channel samples
    float2 *y;                 // [n_frames][M/2]
    float2 *hist_out;          // [NT][M]: the rows of v of the last NT frames
    const float *taps;         // [NT][M/2]: taps[q][i] = h[i + q M/2]
    const float2 *tw;          // as TileArgs::tw
    long long n_frames;
    int hist_rows;
    int n_tiles;
    int tiles_per_inst;
};

template <int M>
constexpr size_t tile_synth_smem_bytes()
{
    using LY = TileLayout<M>;
    return sizeof(float2) * ((size_t)(kThreads / (M / 2)) * LY::NBUF * kFr * LY::FS + (size_t)(M / 32) * kTwPitch);
}

template <int M, int NT, int PAR0>
__global__ void __launch_bounds__(kThreads, NT <= 8 ? 2 : 1) pfbch2_tile_synth_kernel(const SynthArgs a)
{
    using LY = TileLayout<M>;
    constexpr int M2 = M / 2, R2 = M2 / 16;
    constexpr int PR = LY::PR, SS = LY::SS, FS = LY::FS, NBUF = LY::NBUF;
    constexpr int HL = NT - 1;                  // rows of history an output looks back at
    constexpr int WT = (HL + kFr - 1) / kFr;    // warm-up tiles of an instance that starts inside the stream
    constexpr int RPT = 16 / R2;                // pass-B' rows per thread (grouped order)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int inst_l = threadIdx.x / M2, u = threadIdx.x % M2;
    float2 *tile0 = reinterpret_cast<float2 *>(smem_raw) + inst_l * (NBUF * kFr * FS);
    float2 *tws = reinterpret_cast<float2 *>(smem_raw) + (kThreads / M2) * (NBUF * kFr * FS);
    const long long inst = (long long)blockIdx.x * (kThreads / M2) + inst_l;
    const long long t0 = inst * a.tiles_per_inst;
    long long ntl = (long long)a.n_tiles - t0;
    if (ntl > a.tiles_per_inst) ntl = a.tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;

    auto inst_barrier = [&]() {
        if (M2 == kThreads) __syncthreads();
        else if (M2 == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + inst_l), "n"(M2) : "memory");
    };
    // pass-A' role: frame f2, sub-problem p2 (fastest), column t2
    const int f2 = u / (2 * R2), p2 = u & 1, t2 = (u % (2 * R2)) >> 1;
    // the thread's 16 inputs of tile `it`: grouped order rows k1 = t2 + R2 i of (f2, p2); wide order items j = u + M2 i.
    // They are prefetched into REGISTERS one tile ahead (issued when pass B' has consumed the previous set, needed at
    // the top of the next tile): staging them through a cp.async slot cost a third of the kernel's shared-memory
    // wavefronts (ncu: mio_throttle 22 % of the stall samples, shared-memory pipe 59 % busy at 0.67 of the roofline).
    float2 nx[16];
    auto fetch = [&](int it) {
        const long long f0 = (t0 + it) * kFr;
#pragma unroll
        for (int i = 0; i < (LY::WIDE ? 32 * kFr / M2 : RPT); i++) {
            int f, off;
            if (LY::WIDE) { const int j = u + M2 * i; f = j >> 5; off = j & 31; }
            else          { f = f2; off = 2 * (t2 + R2 * i) + p2; }
            const bool ok = it < nt && f0 + f < a.n_frames;
            const float2 *src = a.x + (f0 + f) * M + off;
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) nx[i * R2 + k2] = ok ? __ldg(src + 32 * k2) : make_float2(0.f, 0.f);
        }
    };
    const int it0 = (t0 > 0 && nt > 0) ? -WT : 0;
    fetch(it0);
    // the inter-pass twiddles by column, times 1/2: tws[t2][k1] = e^{+j 2 pi k1 t2 / (M/2)} / 2
    for (int i = u; i < 16 * R2; i += M2) {
        const float2 t = __ldg(a.tw + M2 + i);
        tws[(i % R2) * kTwPitch + i / R2] = make_float2(0.5f * t.x, 0.5f * t.y);
    }
    float h[NT];
#pragma unroll
    for (int q = 0; q < NT; q++) h[q] = __ldg(a.taps + q * M2 + u);
    const float2 w2 = __ldg(a.tw + u);                            // W_M^u
    float2 wlo[HL + kFr], whi[HL + kFr];                          // columns u and u + M/2 of v, oldest first
#pragma unroll
    for (int i = 0; i < HL; i++) {
        wlo[i] = whi[i] = make_float2(0.f, 0.f);
        if (t0 == 0 && a.hist != nullptr) {
            wlo[i] = __ldg(a.hist + (size_t)(a.hist_rows - HL + i) * M + u);
            whi[i] = __ldg(a.hist + (size_t)(a.hist_rows - HL + i) * M + u + M2);
        }
    }
    const float4 *twc = reinterpret_cast<const float4 *>(tws + t2 * kTwPitch);
    const int mac_off = (u / R2) * PR + (u % R2);
    const int sp_off = f2 * FS + p2 * SS;
    __syncthreads();

    for (int it = it0; it < nt; it++) {
        const long long f0 = (t0 + it) * kFr;
        float2 *tile = tile0 + ((it - it0) % NBUF) * (kFr * FS);
        // ---- pass B': rows of the sub-problems, out of the input slot
        if (LY::WIDE) {
#pragma unroll
            for (int i = 0; i < 32 * kFr / M2; i++) {
                const int j = u + M2 * i;
                float2 r[R2];
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++) r[k2] = nx[i * R2 + k2];
                RegDft<R2, +1>::run(r);
                float2 *row = tile + (j >> 5) * FS + (j & 1) * SS + ((j & 31) >> 1) * PR;
#pragma unroll
                for (int c = 0; c < R2; c++) row[c] = r[c];
            }
        } else {
#pragma unroll
            for (int i = 0; i < RPT; i++) {
                float2 r[R2];
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++) r[k2] = nx[i * R2 + k2];
                RegDft<R2, +1>::run(r);
                float2 *row = tile + sp_off + (t2 + R2 * i) * PR;
#pragma unroll
                for (int c = 0; c < R2; c++) row[c] = r[c];
            }
        }
        fetch(it + 1);                                            // the values are consumed: load the next tile's
        if (LY::WIDE) inst_barrier();
        else __syncwarp();
        // ---- pass A': twiddle and 16-point DFT down column t2 of (f2, p2), in place
        {
            float2 *sp = tile + sp_off;
            float2 v[16];
#pragma unroll
            for (int k1 = 0; k1 < 16; k1++) v[k1] = sp[k1 * PR + t2];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 t = twc[k];
                v[2 * k] = (k == 0) ? __fmul2_rn(bc(0.5f), v[0]) : cmul2_rot(v[2 * k], make_float2(t.x, t.y));
                v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
            }
            RegDft<16, +1>::run(v);
#pragma unroll
            for (int r1 = 0; r1 < 16; r1++) sp[r1 * PR + t2] = v[r1];
        }
        inst_barrier();
        // ---- FIR: last radix-2 stage, window push, dot products
        const float2 *src = tile + mac_off;
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            const float2 e0 = src[f * FS];
            const float2 t = cmul2_rot(src[f * FS + SS], w2);
            wlo[HL + f] = __fadd2_rn(e0, t);
            whi[HL + f] = __fadd2_rn(e0, neg2(t));
        }
        if (it >= 0) {
#pragma unroll
            for (int f = 0; f < kFr; f++) {
                const bool odd = ((PAR0 + f) & 1) != 0;           // f0 is even
                float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
                for (int q = 0; q < NT; q++) {
                    const float2 sv = odd ? whi[HL + f - q] : wlo[HL + f - q];
                    if (q & 1) a1 = __ffma2_rn(bc(h[q]), sv, a1);
                    else       a0 = __ffma2_rn(bc(h[q]), sv, a0);
                }
                if (f0 + f < a.n_frames) __stcs(a.y + (f0 + f) * M2 + u, __fadd2_rn(a0, a1));
            }
            if (f0 + kFr > a.n_frames - NT) {
                // the rows of v of the last NT frames are the object's state after this call
#pragma unroll
                for (int f = 0; f < kFr; f++) {
                    const long long r = f0 + f - (a.n_frames - NT);
                    if (r >= 0 && f0 + f < a.n_frames) {
                        a.hist_out[r * M + u] = wlo[HL + f];
                        a.hist_out[r * M + u + M2] = whi[HL + f];
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < HL; i++) {
            wlo[i] = wlo[kFr + i];
            whi[i] = whi[kFr + i];
        }
        if (NBUF == 1) inst_barrier();
    }
}

// firpfbch (critically sampled) synthesizer, firpfbch.c:314-336: v_s = unscaled backward DFT of frame s,
// y[sM + i] = sum_n h[i + nM] v_{s-n}[i].  The kernel above without the frame parity: thread u owns the columns u and
// u + M/2 of v and forms both outputs every frame; the object's state (the last p rows of v) is recomputed by the caller.
struct PfbSynthArgs {
    const float2 *hist;        // hist_rows rows of v preceding frame 0 (single-stream stateful use), or null
    const float2 *x;           // [n_streams][stream_stride]
    float2 *y;
    const float *taps;         // [P][M]: taps[n][i] = h[i + nM]
    const float2 *tw;          // backward twiddles, as TileArgs::tw
    long long n_frames;
    long long stream_stride;
    int hist_rows;
    int n_tiles;
    int tiles_per_inst;
};

template <int M, int NT>
__global__ void __launch_bounds__(kThreads, NT <= 4 ? 2 : 1) pfbch_tile_synth_kernel(const PfbSynthArgs a)
{
    using LY = TileLayout<M>;
    constexpr int M2 = M / 2, R2 = M2 / 16;
    constexpr int PR = LY::PR, SS = LY::SS, FS = LY::FS, NBUF = LY::NBUF;
    constexpr int HL = NT - 1;                  // rows of history an output looks back at
    constexpr int WT = (HL + kFr - 1) / kFr;    // warm-up tiles of an instance that starts inside the stream
    constexpr int RPT = 16 / R2;                // pass-B' rows per thread (grouped order)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int inst_l = threadIdx.x / M2, u = threadIdx.x % M2;
    float2 *tile0 = reinterpret_cast<float2 *>(smem_raw) + inst_l * (NBUF * kFr * FS);
    float2 *tws = reinterpret_cast<float2 *>(smem_raw) + (kThreads / M2) * (NBUF * kFr * FS);
    const long long inst = (long long)blockIdx.x * (kThreads / M2) + inst_l;
    const long long t0 = inst * a.tiles_per_inst;
    long long ntl = (long long)a.n_tiles - t0;
    if (ntl > a.tiles_per_inst) ntl = a.tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;
    const float2 *xstr = a.x + (long long)blockIdx.y * a.stream_stride;
    float2 *ystr = a.y + (long long)blockIdx.y * a.stream_stride;

    auto inst_barrier = [&]() {
        if (M2 == kThreads) __syncthreads();
        else if (M2 == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + inst_l), "n"(M2) : "memory");
    };
    // pass-A' role: frame f2, sub-problem p2 (fastest), column t2
    const int f2 = u / (2 * R2), p2 = u & 1, t2 = (u % (2 * R2)) >> 1;
    // the thread's 16 inputs of tile `it`: grouped order rows k1 = t2 + R2 i of (f2, p2); wide order items j = u + M2 i.
    // They are prefetched into REGISTERS one tile ahead (issued when pass B' has consumed the previous set, needed at
    // the top of the next tile): staging them through a cp.async slot cost a third of the kernel's shared-memory
    // wavefronts (ncu: mio_throttle 22 % of the stall samples, shared-memory pipe 59 % busy at 0.67 of the roofline).
    float2 nx[16];
    auto fetch = [&](int it) {
        const long long f0 = (t0 + it) * kFr;
#pragma unroll
        for (int i = 0; i < (LY::WIDE ? 32 * kFr / M2 : RPT); i++) {
            int f, off;
            if (LY::WIDE) { const int j = u + M2 * i; f = j >> 5; off = j & 31; }
            else          { f = f2; off = 2 * (t2 + R2 * i) + p2; }
            const bool ok = it < nt && f0 + f < a.n_frames;
            const float2 *src = xstr + (f0 + f) * M + off;
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) nx[i * R2 + k2] = ok ? __ldg(src + 32 * k2) : make_float2(0.f, 0.f);
        }
    };
    const int it0 = (t0 > 0 && nt > 0) ? -WT : 0;
    fetch(it0);
    // the inter-pass twiddles by column: tws[t2][k1] = e^{+j 2 pi k1 t2 / (M/2)}
    for (int i = u; i < 16 * R2; i += M2) tws[(i % R2) * kTwPitch + i / R2] = __ldg(a.tw + M2 + i);
    float h0[NT], h1[NT];
#pragma unroll
    for (int q = 0; q < NT; q++) {
        h0[q] = __ldg(a.taps + q * M + u);
        h1[q] = __ldg(a.taps + q * M + u + M2);
    }
    const float2 w2 = __ldg(a.tw + u);                            // W_M^u
    float2 wlo[HL + kFr], whi[HL + kFr];                          // columns u and u + M/2 of v, oldest first
#pragma unroll
    for (int i = 0; i < HL; i++) {
        wlo[i] = whi[i] = make_float2(0.f, 0.f);
        if (t0 == 0 && a.hist != nullptr) {
            wlo[i] = __ldg(a.hist + (size_t)(a.hist_rows - HL + i) * M + u);
            whi[i] = __ldg(a.hist + (size_t)(a.hist_rows - HL + i) * M + u + M2);
        }
    }
    const float4 *twc = reinterpret_cast<const float4 *>(tws + t2 * kTwPitch);
    const int mac_off = (u / R2) * PR + (u % R2);
    const int sp_off = f2 * FS + p2 * SS;
    __syncthreads();

    for (int it = it0; it < nt; it++) {
        const long long f0 = (t0 + it) * kFr;
        float2 *tile = tile0 + ((it - it0) % NBUF) * (kFr * FS);
        // ---- pass B': rows of the sub-problems, out of the input slot
        if (LY::WIDE) {
#pragma unroll
            for (int i = 0; i < 32 * kFr / M2; i++) {
                const int j = u + M2 * i;
                float2 r[R2];
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++) r[k2] = nx[i * R2 + k2];
                RegDft<R2, +1>::run(r);
                float2 *row = tile + (j >> 5) * FS + (j & 1) * SS + ((j & 31) >> 1) * PR;
#pragma unroll
                for (int c = 0; c < R2; c++) row[c] = r[c];
            }
        } else {
#pragma unroll
            for (int i = 0; i < RPT; i++) {
                float2 r[R2];
#pragma unroll
                for (int k2 = 0; k2 < R2; k2++) r[k2] = nx[i * R2 + k2];
                RegDft<R2, +1>::run(r);
                float2 *row = tile + sp_off + (t2 + R2 * i) * PR;
#pragma unroll
                for (int c = 0; c < R2; c++) row[c] = r[c];
            }
        }
        fetch(it + 1);                                            // the values are consumed: load the next tile's
        if (LY::WIDE) inst_barrier();
        else __syncwarp();
        // ---- pass A': twiddle and 16-point DFT down column t2 of (f2, p2), in place
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
            RegDft<16, +1>::run(v);
#pragma unroll
            for (int r1 = 0; r1 < 16; r1++) sp[r1 * PR + t2] = v[r1];
        }
        inst_barrier();
        // ---- FIR: last radix-2 stage, window push, dot products
        const float2 *src = tile + mac_off;
#pragma unroll
        for (int f = 0; f < kFr; f++) {
            const float2 e0 = src[f * FS];
            const float2 t = cmul2_rot(src[f * FS + SS], w2);
            wlo[HL + f] = __fadd2_rn(e0, t);
            whi[HL + f] = __fadd2_rn(e0, neg2(t));
        }
        if (it >= 0) {
#pragma unroll
            for (int f = 0; f < kFr; f++) {
                float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
                for (int n = NT - 1; n >= 0; n--) {
                    a0 = __ffma2_rn(bc(h0[n]), wlo[HL + f - n], a0);
                    a1 = __ffma2_rn(bc(h1[n]), whi[HL + f - n], a1);
                }
                if (f0 + f < a.n_frames) {
                    __stcs(ystr + (f0 + f) * M + u, a0);
                    __stcs(ystr + (f0 + f) * M + u + M2, a1);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < HL; i++) {
            wlo[i] = wlo[kFr + i];
            whi[i] = whi[kFr + i];
        }
        if (NBUF == 1) inst_barrier();
    }
}

// ---------------------------------------------------------------------------------- synthesizer, M = 1024
// Same transposed flow with a radix-4 last stage: one instance = one CTA of 256 threads, thread u owns the FOUR columns
// u + 256 q of v.  Per frame four 256-point problems X_p[k] = X[4k + p]: 64 threads per frame, lane = (p fastest, row):
// 16-point DFTs along the rows straight out of the thread's input slot (256-byte runs of the frame), a 64-thread
// named barrier, twiddle (times 1/2) and 16-point DFTs down the columns; then thread u multiplies E_p[u] by W1024^(u p),
// does the radix-4 butterfly v[u + 256 q] = sum_p (+j)^(q p) W^(u p) E_p[u], pushes the four values into register
// windows and forms its two outputs y[t][u], y[t][u + 256] (columns u, u + 256 on even frames, u + 512, u + 768 on
// odd ones).  4 frames per tile, two tile buffers (one block barrier per tile), sub-problems 276 units apart.
constexpr int kQFr = 4;
constexpr int kQSub = 276;
constexpr int kQFrame = 4 * kQSub;

constexpr bool kQuadSynthRegPf = true;
template <int NT>
constexpr size_t quad_synth_smem_bytes()
{
    return sizeof(float2) * ((kQuadSynthRegPf ? (size_t)0 : (size_t)16 * kThreads) + (size_t)2 * kQFr * kQFrame + 16 * kTwPitch);
}

template <int NT, int PAR0>
__global__ void __launch_bounds__(kThreads, (NT <= 8 && !kQuadSynthRegPf) ? 2 : 1) pfbch2_quad_synth_1024_kernel(const SynthArgs a)
{
    constexpr int M = 1024, M2 = 512, HL = NT - 1;
    constexpr int WT = (HL + kQFr - 1) / kQFr;                        // warm-up tiles
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *slot = reinterpret_cast<float2 *>(smem_raw);              // [16][kThreads]: a thread's own 16 inputs of a tile
    float2 *tile0 = slot + (kQuadSynthRegPf ? 0 : 16 * kThreads);
    float2 *tws = tile0 + 2 * kQFr * kQFrame;                         // [c][k1], pitch kTwPitch, times 1/2
    const int u = threadIdx.x;
    const long long t0 = (long long)blockIdx.x * a.tiles_per_inst;
    long long ntl = (long long)a.n_tiles - t0;
    if (ntl > a.tiles_per_inst) ntl = a.tiles_per_inst;
    const int nt = ntl > 0 ? (int)ntl : 0;

    // phase-A' role: frame f2, sub-problem pp (fastest), row / column idx
    const int f2 = u >> 6, pp = u & 3, idx = (u >> 2) & 15;
    const uint32_t slot_u32 = smem_u32(slot + u);
    float2 nx[16];
    auto fetch = [&](int it) {
        const long long f0 = (t0 + it) * kQFr;
        const bool ok = it < nt && f0 + f2 < a.n_frames;
        const float2 *src = a.x + (f0 + f2) * M + 4 * idx + pp;       // X_p[k1 = idx + 16 k2] at + 64 k2
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            if (kQuadSynthRegPf) nx[k2] = ok ? __ldg(src + 64 * k2) : make_float2(0.f, 0.f);
            else cp_async8(slot_u32 + k2 * kThreads * (int)sizeof(float2), ok ? (const void *)(src + 64 * k2) : (const void *)a.x, ok ? 8u : 0u);
        }
        if (!kQuadSynthRegPf) cp_async_commit();
    };
    const int it0 = (t0 > 0 && nt > 0) ? -WT : 0;
    fetch(it0);
    {
        const float2 t = __ldg(a.tw + 768 + u);                       // tw256[k1 = u >> 4][c = u & 15]
        tws[(u & 15) * kTwPitch + (u >> 4)] = make_float2(0.5f * t.x, 0.5f * t.y);
    }
    const float2 t1 = __ldg(a.tw + u), t2 = __ldg(a.tw + 256 + u), t3 = __ldg(a.tw + 512 + u);   // W1024^(u p)
    float h0[NT], h1[NT];
#pragma unroll
    for (int q = 0; q < NT; q++) {
        h0[q] = __ldg(a.taps + q * M2 + u);
        h1[q] = __ldg(a.taps + q * M2 + u + 256);
    }
    float2 w[4][HL + kQFr];                                           // columns u + 256 q of v, oldest first
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int i = 0; i < HL; i++) {
            w[q][i] = make_float2(0.f, 0.f);
            if (t0 == 0 && a.hist != nullptr) w[q][i] = __ldg(a.hist + (size_t)(a.hist_rows - HL + i) * M + u + 256 * q);
        }
    const float4 *twc = reinterpret_cast<const float4 *>(tws + idx * kTwPitch);
    const int z_off = (u >> 4) * 17 + (u & 15);
    __syncthreads();

    for (int it = it0; it < nt; it++) {
        const long long f0 = (t0 + it) * kQFr;
        float2 *tile = tile0 + ((it - it0) & 1) * (kQFr * kQFrame);
        float2 *sp = tile + f2 * kQFrame + pp * kQSub;
        if (!kQuadSynthRegPf) cp_async_wait<0>();
        float2 v[16];
        // ---- pass B': row k1 = idx of sub-problem pp
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) v[k2] = kQuadSynthRegPf ? nx[k2] : slot[k2 * kThreads + u];
        RegDft<16, +1>::run(v);
#pragma unroll
        for (int c = 0; c < 16; c++) sp[idx * 17 + c] = v[c];
        fetch(it + 1);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + f2) : "memory");
        // ---- pass A': twiddle and DFT down column c = idx, in place
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) v[k1] = sp[k1 * 17 + idx];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 t = twc[k];
            v[2 * k] = (k == 0) ? __fmul2_rn(bc(0.5f), v[0]) : cmul2_rot(v[2 * k], make_float2(t.x, t.y));
            v[2 * k + 1] = cmul2_rot(v[2 * k + 1], make_float2(t.z, t.w));
        }
        RegDft<16, +1>::run(v);
#pragma unroll
        for (int r1 = 0; r1 < 16; r1++) sp[r1 * 17 + idx] = v[r1];
        __syncthreads();
        // ---- FIR: radix-4 last stage, window push, dot products
#pragma unroll
        for (int f = 0; f < kQFr; f++) {
            const float2 *zf = tile + f * kQFrame + z_off;
            const float2 b0 = zf[0], b1 = cmul2_rot(zf[kQSub], t1), b2 = cmul2_rot(zf[2 * kQSub], t2), b3 = cmul2_rot(zf[3 * kQSub], t3);
            const float2 s02 = __fadd2_rn(b0, b2), d02 = __fadd2_rn(b0, neg2(b2));
            const float2 s13 = __fadd2_rn(b1, b3), d13 = __fadd2_rn(b1, neg2(b3));
            w[0][HL + f] = __fadd2_rn(s02, s13);
            w[1][HL + f] = __fadd2_rn(d02, rot90(d13));               // W4 = +j (backward transform)
            w[2][HL + f] = __fadd2_rn(s02, neg2(s13));
            w[3][HL + f] = __fadd2_rn(d02, rot270(d13));
        }
        if (it >= 0) {
#pragma unroll
            for (int f = 0; f < kQFr; f++) {
                const bool odd = ((PAR0 + f) & 1) != 0;               // f0 is even
                float2 a0 = make_float2(0.f, 0.f), a1 = a0, c0 = a0, c1 = a0;
#pragma unroll
                for (int q = 0; q < NT; q++) {
                    const float2 s0 = odd ? w[2][HL + f - q] : w[0][HL + f - q];
                    const float2 s1 = odd ? w[3][HL + f - q] : w[1][HL + f - q];
                    if (q & 1) { a1 = __ffma2_rn(bc(h0[q]), s0, a1); c1 = __ffma2_rn(bc(h1[q]), s1, c1); }
                    else       { a0 = __ffma2_rn(bc(h0[q]), s0, a0); c0 = __ffma2_rn(bc(h1[q]), s1, c0); }
                }
                if (f0 + f < a.n_frames) {
                    __stcs(a.y + (f0 + f) * M2 + u, __fadd2_rn(a0, a1));
                    __stcs(a.y + (f0 + f) * M2 + u + 256, __fadd2_rn(c0, c1));
                }
            }
            if (f0 + kQFr > a.n_frames - NT) {
#pragma unroll
                for (int f = 0; f < kQFr; f++) {
                    const long long r = f0 + f - (a.n_frames - NT);
                    if (r >= 0 && f0 + f < a.n_frames) {
#pragma unroll
                        for (int q = 0; q < 4; q++) a.hist_out[r * M + u + 256 * q] = w[q][HL + f];
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int i = 0; i < HL; i++) w[q][i] = w[q][kQFr + i];
    }
    cp_async_wait<0>();
}

template <int NT>
void launch_quad_synth_N(const SynthArgs &a, int par0, unsigned int grid, cudaStream_t st)
{
    constexpr size_t smem = quad_synth_smem_bytes<NT>();
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured)) {
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_quad_synth_1024_kernel<NT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_quad_synth_1024_kernel<NT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (par0) pfbch2_quad_synth_1024_kernel<NT, 1><<<grid, kThreads, smem, st>>>(a);
    else      pfbch2_quad_synth_1024_kernel<NT, 0><<<grid, kThreads, smem, st>>>(a);
}

template <int M, int NT>
void launch_synth_MN(const SynthArgs &a, int par0, unsigned int grid, cudaStream_t st)
{
    constexpr size_t smem = tile_synth_smem_bytes<M>();
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured)) {
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_tile_synth_kernel<M, NT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_tile_synth_kernel<M, NT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (par0) pfbch2_tile_synth_kernel<M, NT, 1><<<grid, kThreads, smem, st>>>(a);
    else      pfbch2_tile_synth_kernel<M, NT, 0><<<grid, kThreads, smem, st>>>(a);
}

template <int M>
void launch_synth_M(const SynthArgs &a, unsigned int NT, int par0, unsigned int grid, cudaStream_t st)
{
    switch (NT) {
    case 8: launch_synth_MN<M, 8>(a, par0, grid, st); break;
    case 12: launch_synth_MN<M, 12>(a, par0, grid, st); break;
    case 16: launch_synth_MN<M, 16>(a, par0, grid, st); break;
    case 20: launch_synth_MN<M, 20>(a, par0, grid, st); break;
    case 24: launch_synth_MN<M, 24>(a, par0, grid, st); break;
    default: lc::fatal("launch_pfbch2_tile_synth", "unsupported tap count");
    }
}

template <int M, int L>
void launch_ML(const TileArgs &a, unsigned int grid, cudaStream_t st)
{
    constexpr size_t smem = tile_smem_bytes<M, L>();
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch2_tile_kernel<M, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pfbch2_tile_kernel<M, L><<<grid, kThreads, smem, st>>>(a);
}

template <int M>
void launch_M(const TileArgs &a, unsigned int L, unsigned int grid, cudaStream_t st)
{
    switch (L) {
    case 4: launch_ML<M, 4>(a, grid, st); break;
    case 6: launch_ML<M, 6>(a, grid, st); break;
    case 8: launch_ML<M, 8>(a, grid, st); break;
    case 10: launch_ML<M, 10>(a, grid, st); break;
    case 12: launch_ML<M, 12>(a, grid, st); break;
    case 14: launch_ML<M, 14>(a, grid, st); break;
    default: lc::fatal("launch_pfbch2_tile", "unsupported tap count");
    }
}

}  // namespace

namespace lc {
namespace pfb {

bool tile_pfbch2_supported(unsigned int M, unsigned int L)
{
    return (M == 64 || M == 128 || M == 256 || M == 512) && (L == 4 || L == 6 || L == 8 || L == 10 || L == 12 || L == 14);
}

void tile_pfbch2_make_twiddles(std::vector<float2> &tw, unsigned int M)
{
    // tw2[u] = e^{+j 2 pi u / M}, u < M/2; then twp[k1][r2] = e^{+j 2 pi k1 r2 / (M/2)}, k1 < 16, r2 < M/32
    const unsigned int M2 = M / 2, R2 = M2 / 16;
    tw.resize(M2 + 16 * R2);
    for (unsigned int u = 0; u < M2; u++) {
        const double ang = 2.0 * M_PI * (double)u / (double)M;
        tw[u] = make_float2((float)cos(ang), (float)sin(ang));
    }
    for (unsigned int k1 = 0; k1 < 16; k1++)
        for (unsigned int r2 = 0; r2 < R2; r2++) {
            const double ang = 2.0 * M_PI * (double)(k1 * r2) / (double)M2;
            tw[M2 + k1 * R2 + r2] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

bool tile_pfbch_supported(unsigned int M, unsigned int p)
{
    // (M = 256 stays on pfbch_fused_256_kernel: this kernel measured 0.91 against 0.90 at p = 4 but 0.74 against 0.87 at p = 10)
    return ((M == 64 || M == 128 || M == 512) && (p == 4 || p == 6 || p == 8 || p == 10 || p == 12 || p == 14)) || (M == 256 && (p == 6 || p == 8 || p == 12 || p == 14));
}

void tile_pfbch_make_twiddles(std::vector<float2> &tw, unsigned int M)
{
    tile_pfbch2_make_twiddles(tw, M);
    for (float2 &t : tw) t.y = -t.y;                              // forward transform
}

namespace {
template <typename TapT>
long long launch_pfbch_tile_T(const float2 *d_hist, const float2 *d_x, int H, const TapT *d_taps, const float2 *d_tw,
                              float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                              unsigned int n_streams, long long stream_stride, cudaStream_t st)
{
    const long long n_tiles = (long long)((n_frames + kFr - 1) / kFr);
    if (n_tiles == 0 || n_tiles > (1ll << 30) || n_streams == 0 || n_streams > 65535) return 0;
    PfbTileArgs a;
    a.hist = d_hist;
    a.x = d_x;
    a.y = d_y;
    a.taps = reinterpret_cast<const float *>(d_taps);
    a.tw = d_tw;
    a.H = H;
    a.n_frames = (long long)n_frames;
    a.stream_stride = stream_stride;
    a.n_tiles = (int)n_tiles;
    const long long inst_per_cta = kThreads / (M / 2);
    long long tpi = 64;
    while (tpi > 1 && (n_tiles / tpi) * n_streams < 148ll * 4 * inst_per_cta) tpi >>= 1;
    a.tiles_per_inst = (int)tpi;
    const long long n_inst = (n_tiles + tpi - 1) / tpi;
    dim3 grid((unsigned int)((n_inst + inst_per_cta - 1) / inst_per_cta), n_streams, 1);
    switch (M) {
    case 64: launch_pfb_M<64, TapT>(a, p, grid, st); break;
    case 128: launch_pfb_M<128, TapT>(a, p, grid, st); break;
    case 256:
        launch_pfb_M<256, TapT>(a, p, grid, st);
        break;
    case 512: launch_pfb_M<512, TapT>(a, p, grid, st); break;
    default: lc::fatal("launch_pfbch_tile", "unsupported channel count");
    }
    LC_LAUNCH_CHECK();
    return (long long)n_frames;
}
}  // namespace

long long launch_pfbch_tile(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps, const float2 *d_tw,
                            float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                            unsigned int n_streams, long long stream_stride, cudaStream_t st)
{
    return launch_pfbch_tile_T<float>(d_hist, d_x, H, d_taps, d_tw, d_y, M, p, n_frames, n_streams, stream_stride, st);
}

bool tile_pfbch_cc_supported(unsigned int M, unsigned int p)
{
    return (M == 64 || M == 128 || M == 256 || M == 512) && (p == 4 || p == 6 || p == 8 || p == 10 || p == 12);
}

long long launch_pfbch_tile_cc(const float2 *d_hist, const float2 *d_x, int H, const float2 *d_taps, const float2 *d_tw,
                               float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames, cudaStream_t st)
{
    return launch_pfbch_tile_T<float2>(d_hist, d_x, H, d_taps, d_tw, d_y, M, p, n_frames, 1, 0, st);
}

bool tile_pfbch_synth_supported(unsigned int M, unsigned int p)
{
    return (M == 64 || M == 128 || M == 512) && (p == 4 || p == 6 || p == 8 || p == 10 || p == 12 || p == 14);
}

template <int M, int P>
void launch_pfb_synth_MP(const PfbSynthArgs &a, dim3 grid, cudaStream_t st)
{
    constexpr size_t smem = tile_synth_smem_bytes<M>();
    static std::atomic<unsigned long long> configured{0};
    if (lc::first_time_on_device(configured))
        LC_CUDA_CHECK(cudaFuncSetAttribute(pfbch_tile_synth_kernel<M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pfbch_tile_synth_kernel<M, P><<<grid, kThreads, smem, st>>>(a);
}

template <int M>
void launch_pfb_synth_M(const PfbSynthArgs &a, unsigned int p, dim3 grid, cudaStream_t st)
{
    switch (p) {
    case 4: launch_pfb_synth_MP<M, 4>(a, grid, st); break;
    case 6: launch_pfb_synth_MP<M, 6>(a, grid, st); break;
    case 8: launch_pfb_synth_MP<M, 8>(a, grid, st); break;
    case 10: launch_pfb_synth_MP<M, 10>(a, grid, st); break;
    case 12: launch_pfb_synth_MP<M, 12>(a, grid, st); break;
    case 14: launch_pfb_synth_MP<M, 14>(a, grid, st); break;
    default: lc::fatal("launch_pfbch_tile_synth", "unsupported tap count");
    }
}

long long launch_pfbch_tile_synth(const float2 *d_hist_rows, int hist_rows, const float2 *d_x, const float *d_taps,
                                  const float2 *d_tw, float2 *d_y, unsigned int M, unsigned int p, unsigned long long n_frames,
                                  unsigned int n_streams, long long stream_stride, cudaStream_t st)
{
    const long long n_tiles = (long long)((n_frames + kFr - 1) / kFr);
    if (n_tiles == 0 || n_tiles > (1ll << 30) || n_streams == 0 || n_streams > 65535) return 0;
    PfbSynthArgs a;
    a.hist = d_hist_rows;
    a.hist_rows = hist_rows;
    a.x = d_x;
    a.y = d_y;
    a.taps = d_taps;
    a.tw = d_tw;
    a.n_frames = (long long)n_frames;
    a.stream_stride = stream_stride;
    a.n_tiles = (int)n_tiles;
    const long long inst_per_cta = kThreads / (M / 2);
    long long tpi = 64;
    while (tpi > 4 && (n_tiles / tpi) * n_streams < 148ll * 4 * inst_per_cta) tpi >>= 1;
    a.tiles_per_inst = (int)tpi;
    const long long n_inst = (n_tiles + tpi - 1) / tpi;
    dim3 grid((unsigned int)((n_inst + inst_per_cta - 1) / inst_per_cta), n_streams, 1);
    switch (M) {
    case 64: launch_pfb_synth_M<64>(a, p, grid, st); break;
    case 128: launch_pfb_synth_M<128>(a, p, grid, st); break;
    case 512: launch_pfb_synth_M<512>(a, p, grid, st); break;
    default: fatal("launch_pfbch_tile_synth", "unsupported channel count");
    }
    LC_LAUNCH_CHECK();
    return (long long)n_frames;
}

bool tile_pfbch2_synth_supported(unsigned int M, unsigned int n_taps)
{
    return ((M == 64 || M == 128 || M == 256 || M == 512) && (n_taps == 8 || n_taps == 12 || n_taps == 16 || n_taps == 20 || n_taps == 24)) ||
           (M == 1024 && (n_taps == 8 || n_taps == 12 || n_taps == 16));
}

void tile_pfbch2_synth_make_twiddles(std::vector<float2> &tw, unsigned int M)
{
    if (M != 1024) {
        tile_pfbch2_make_twiddles(tw, M);
        return;
    }
    // pfbch2_quad_synth_1024_kernel: tw4[p-1][u] = e^{+j 2 pi u p / 1024}, p = 1..3, u < 256; then tw256[k1][c] = e^{+j 2 pi k1 c / 256}
    tw.resize(3 * 256 + 256);
    for (int pp = 1; pp <= 3; pp++)
        for (int u = 0; u < 256; u++) {
            const double ang = 2.0 * M_PI * (double)(u * pp) / 1024.0;
            tw[(pp - 1) * 256 + u] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (int k1 = 0; k1 < 16; k1++)
        for (int c = 0; c < 16; c++) {
            const double ang = 2.0 * M_PI * (double)(k1 * c) / 256.0;
            tw[768 + k1 * 16 + c] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

long long launch_pfbch2_tile_synth(const float2 *d_hist_rows, int hist_rows, const float2 *d_x, const float *d_taps,
                                   const float2 *d_tw, float2 *d_y, float2 *d_hist_out, unsigned int M, unsigned int n_taps,
                                   unsigned long long n_frames, int parity0, cudaStream_t st)
{
    const int fr_per_tile = (M == 1024) ? kQFr : kFr;
    const long long n_tiles = (long long)((n_frames + fr_per_tile - 1) / fr_per_tile);
    if (n_frames < n_taps || n_tiles > (1ll << 30)) return 0;     // the state rows all come from this call
    SynthArgs a;
    a.hist = d_hist_rows;
    a.hist_rows = hist_rows;
    a.x = d_x;
    a.y = d_y;
    a.hist_out = d_hist_out;
    a.taps = d_taps;
    a.tw = d_tw;
    a.n_frames = (long long)n_frames;
    a.n_tiles = (int)n_tiles;
    // 64 tiles per instance keep the warm-up tiles (1 - 3) below 5 %; never fewer than 4 (the warm-up reaches back 3)
    if (M == 1024) {
        long long tpi = 128;                                      // 512 frames per instance; never fewer than 8 tiles (warm-up: 2 - 4)
        while (tpi > 8 && (n_tiles / tpi) < 148ll * 4) tpi >>= 1;
        a.tiles_per_inst = (int)tpi;
        const unsigned int grid = (unsigned int)((n_tiles + tpi - 1) / tpi);
        if (n_taps == 8) launch_quad_synth_N<8>(a, parity0 & 1, grid, st);
        else if (n_taps == 12) launch_quad_synth_N<12>(a, parity0 & 1, grid, st);
        else if (n_taps == 16) launch_quad_synth_N<16>(a, parity0 & 1, grid, st);
        else fatal("launch_pfbch2_tile_synth", "unsupported tap count");
        LC_LAUNCH_CHECK();
        return (long long)n_frames;
    }
    const long long inst_per_cta = kThreads / (M / 2);
    long long tpi = 64;
    while (tpi > 4 && (n_tiles / tpi) < 148ll * 4 * inst_per_cta) tpi >>= 1;
    a.tiles_per_inst = (int)tpi;
    const long long n_inst = (n_tiles + tpi - 1) / tpi;
    const unsigned int grid = (unsigned int)((n_inst + inst_per_cta - 1) / inst_per_cta);
    switch (M) {
    case 64: launch_synth_M<64>(a, n_taps, parity0 & 1, grid, st); break;
    case 128: launch_synth_M<128>(a, n_taps, parity0 & 1, grid, st); break;
    case 256: launch_synth_M<256>(a, n_taps, parity0 & 1, grid, st); break;
    case 512: launch_synth_M<512>(a, n_taps, parity0 & 1, grid, st); break;
    default: fatal("launch_pfbch2_tile_synth", "unsupported channel count");
    }
    LC_LAUNCH_CHECK();
    return (long long)n_frames;
}

long long launch_pfbch2_tile(const float2 *d_hist, const float2 *d_x, int H, const float *d_taps_scaled, const float2 *d_tw,
                             float2 *d_y, unsigned int M, unsigned int L, unsigned long long n_frames, int parity0,
                             cudaStream_t st)
{
    const long long n_tiles = (long long)((n_frames + kFr - 1) / kFr);
    if (n_tiles == 0 || n_tiles > (1ll << 30)) return 0;
    TileArgs a;
    a.hist = d_hist;
    a.x = d_x;
    a.y = d_y;
    a.taps = d_taps_scaled;
    a.tw = d_tw;
    a.H = H;
    a.n_frames = (long long)n_frames;
    a.par0 = parity0 & 1;
    a.n_tiles = (int)n_tiles;
    // ~64 tiles (512 frames) per instance keeps the re-read of the 4m-1 history frames below 4 %, while leaving
    // enough instances to fill the GPU; shrink for short blocks
    const long long inst_per_cta = kThreads / (M / 2);
    long long tpi = 64;
    while (tpi > 1 && (n_tiles / tpi) < 148ll * 4 * inst_per_cta) tpi >>= 1;
    a.tiles_per_inst = (int)tpi;
    const long long n_inst = (n_tiles + tpi - 1) / tpi;
    const unsigned int grid = (unsigned int)((n_inst + inst_per_cta - 1) / inst_per_cta);
    switch (M) {
    case 64: launch_M<64>(a, L, grid, st); break;
    case 128: launch_M<128>(a, L, grid, st); break;
    case 256: launch_M<256>(a, L, grid, st); break;
    case 512: launch_M<512>(a, L, grid, st); break;
    default: fatal("launch_pfbch2_tile", "unsupported channel count");
    }
    LC_LAUNCH_CHECK();
    return (long long)n_frames;
}

}  // namespace pfb
}  // namespace lc
