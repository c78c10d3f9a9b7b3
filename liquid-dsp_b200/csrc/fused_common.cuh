// fused_common.cuh -- device building blocks shared by the fused channelizer kernels:
// mbarrier / bulk-copy PTX wrappers, packed FP32x2 arithmetic, and the register DFT.
#pragma once
#include <cstdint>

#include "lc_common.cuh"

namespace lcf {

// ---------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA engine)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, float2 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// ---------------------------------------------------------------------------------- register DFT
// cos / sin of 2 pi k / 32, k = 0..15
__device__ constexpr float kCos32[16] = {1.0f,
                                         0.98078528040323044913f,
                                         0.92387953251128675613f,
                                         0.83146961230254523708f,
                                         0.70710678118654752440f,
                                         0.55557023301960222474f,
                                         0.38268343236508977173f,
                                         0.19509032201612826785f,
                                         0.0f,
                                         -0.19509032201612826785f,
                                         -0.38268343236508977173f,
                                         -0.55557023301960222474f,
                                         -0.70710678118654752440f,
                                         -0.83146961230254523708f,
                                         -0.92387953251128675613f,
                                         -0.98078528040323044913f};
__device__ constexpr float kSin32[16] = {0.0f,
                                         0.19509032201612826785f,
                                         0.38268343236508977173f,
                                         0.55557023301960222474f,
                                         0.70710678118654752440f,
                                         0.83146961230254523708f,
                                         0.92387953251128675613f,
                                         0.98078528040323044913f,
                                         1.0f,
                                         0.98078528040323044913f,
                                         0.92387953251128675613f,
                                         0.83146961230254523708f,
                                         0.70710678118654752440f,
                                         0.55557023301960222474f,
                                         0.38268343236508977173f,
                                         0.19509032201612826785f};

// packed FP32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2): one instruction works on a {re, im} register
// pair; ptxas folds make_float2(t, t) into a scalar-broadcast operand and make_float2(-a.y, a.x) into a
// swapped, half-negated operand, so none of these cost extra moves.  The FMA pipe spends two cycles on
// a packed instruction (same flop rate as two scalar ones) but it takes a single issue slot.
__device__ __forceinline__ float2 bc(float t) { return make_float2(t, t); }
__device__ __forceinline__ float2 rot90(float2 a) { return make_float2(-a.y, a.x); }      //  j * a
__device__ __forceinline__ float2 rot270(float2 a) { return make_float2(a.y, -a.x); }     // -j * a
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
// complex product a * t: FMUL2 (t.x broadcast) + FFMA2 with the pair {-t.y, t.y} against the swapped a.
// The sign flip is an integer XOR (ALU pipe), so the FMA pipe sees exactly two packed instructions;
// writing it as t.y * {-a.y, a.x} makes ptxas negate a.y with a scalar FADD on the FMA pipe instead.
__device__ __forceinline__ float2 cmul2(float2 a, float2 t)
{
    const float2 ms_s = make_float2(__int_as_float(__float_as_int(t.y) ^ (int)0x80000000), t.y);
    return __ffma2_rn(ms_s, make_float2(a.y, a.x), __fmul2_rn(bc(t.x), a));
}

// the same product as t.x * a + t.y * (j a): the negation inside j a = {-a.y, a.x} costs one scalar FADD on the FMA
// pipe, but no register moves (the form above builds its {-t.y, t.y} pair with two MOVs on the half-rate ALU pipe)
__device__ __forceinline__ float2 cmul2_rot(float2 a, float2 t)
{
    return __ffma2_rn(bc(t.y), rot90(a), __fmul2_rn(bc(t.x), a));
}

// N-point DFT of v[] in registers, natural order in and out, e^{SIGN j 2 pi nk/N}.
// Decimation in time; every index is a compile-time constant, so the even/odd split is renaming.
template <int N, int SIGN>
struct RegDft {
    static __device__ __forceinline__ void run(float2 (&v)[N])
    {
        float2 e[N / 2], o[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; i++) {
            e[i] = v[2 * i];
            o[i] = v[2 * i + 1];
        }
        RegDft<N / 2, SIGN>::run(e);
        RegDft<N / 2, SIGN>::run(o);
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            if (k == 0) {
                v[k] = __fadd2_rn(e[k], o[k]);
                v[k + N / 2] = __fadd2_rn(e[k], neg2(o[k]));
            } else if (4 * k == N) {
                // w = SIGN * j
                const float2 r = (SIGN > 0) ? rot90(o[k]) : rot270(o[k]);
                v[k] = __fadd2_rn(e[k], r);
                v[k + N / 2] = __fadd2_rn(e[k], neg2(r));
            } else {
                const float c = kCos32[k * (32 / N)];
                const float sn = (float)SIGN * kSin32[k * (32 / N)];
                // a + w b = a + c b + sn (j b) in two packed FMAs; a - w b = 2a - (a + w b) in one
                const float2 t = __ffma2_rn(bc(c), o[k], e[k]);
                v[k] = __ffma2_rn(bc(sn), rot90(o[k]), t);
                v[k + N / 2] = __ffma2_rn(bc(2.0f), e[k], neg2(v[k]));
            }
        }
    }
};
template <int SIGN>
struct RegDft<1, SIGN> {
    static __device__ __forceinline__ void run(float2 (&)[1]) {}
};

// ---------------------------------------------------------------------------------- firpfbch2 MAC state
// MAC state of one window pair (w, w+M/2) of the firpfbch2 analyzer: histories and branches, all in
// registers.  The pair receives stream sample F*M/2 + M/2-1-w on frame F; the window pushed on even
// local frames keeps its 2m samples in dA, the one pushed on odd local frames in dB (circular, slot
// index = double frame mod L).
template <int L>
struct PairState {
    float hA[L], hB[L];        // taps of branch w and branch w+M/2 (tap n multiplies the n-th newest sample)
    float2 dA[L], dB[L];       // circular histories of the window pushed on even / odd local frames
};

// taps: analyzer table [2][L][M] (parity, tap, window), f_begin: first (even-local) frame of the range
template <int L, int M>
__device__ __forceinline__ void pair_load(PairState<L> &s, const float *__restrict__ taps,
                                          const float2 *__restrict__ hist, const float2 *__restrict__ x, int H,
                                          int w, long long f_begin)
{
    constexpr int M2 = M / 2;
#pragma unroll
    for (int n = 0; n < L; n++) {
        s.hA[n] = __ldg(taps + (size_t)n * M + w);                // parity-0 table: window w -> branch w
        s.hB[n] = __ldg(taps + (size_t)n * M + w + M2);           // window w+M/2 -> branch w+M/2
    }
    // history: the pair's sample of frame F is stream index F*M/2 + M/2-1 - w
    const long long base = f_begin * M2 + (M2 - 1 - w);
#pragma unroll
    for (int n = 1; n < L; n++) s.dA[L - n] = lcd::ld_stream(hist, x, base - 2LL * n * M2, H);
    s.dA[0] = make_float2(0.f, 0.f);
#pragma unroll
    for (int n = 0; n < L; n++) s.dB[L - 1 - n] = lcd::ld_stream(hist, x, base - (1LL + 2 * n) * M2, H);
}

// the two dot products of a pair on the even (ODD = 0) or odd (ODD = 1) frame of double frame uu:
// o1 belongs to the window pushed on even frames, o2 to the one pushed on odd frames
template <int L, int ODD>
__device__ __forceinline__ void pair_mac(const PairState<L> &s, int uu, float2 &o1, float2 &o2)
{
    o1 = make_float2(0.f, 0.f);
    o2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int n = L - 1; n >= 0; n--) {
        const float2 s1 = s.dA[(uu - n + L) % L];
        const float2 s2 = s.dB[(uu - (ODD ? 0 : 1) - n + 2 * L) % L];
        o1 = __ffma2_rn(bc(ODD ? s.hB[n] : s.hA[n]), s1, o1);
        o2 = __ffma2_rn(bc(ODD ? s.hA[n] : s.hB[n]), s2, o2);
    }
}

// same dot products with two accumulators per output (even / odd taps, summed at the end): four
// independent FFMA2 chains for a thread that owns a single pair
template <int L, int ODD>
__device__ __forceinline__ void pair_mac_split(const PairState<L> &s, int uu, float2 &o1, float2 &o2)
{
    static_assert(L % 2 == 0, "tap count per branch is 2m");
    float2 e1 = make_float2(0.f, 0.f), e2 = e1, f1 = e1, f2 = e1;
#pragma unroll
    for (int n = L - 1; n >= 1; n -= 2) {
        const float2 a1 = s.dA[(uu - n + L) % L], b1 = s.dA[(uu - (n - 1) + L) % L];
        const float2 a2 = s.dB[(uu - (ODD ? 0 : 1) - n + 2 * L) % L], b2 = s.dB[(uu - (ODD ? 0 : 1) - (n - 1) + 2 * L) % L];
        e1 = __ffma2_rn(bc(ODD ? s.hB[n] : s.hA[n]), a1, e1);
        e2 = __ffma2_rn(bc(ODD ? s.hA[n] : s.hB[n]), a2, e2);
        f1 = __ffma2_rn(bc(ODD ? s.hB[n - 1] : s.hA[n - 1]), b1, f1);
        f2 = __ffma2_rn(bc(ODD ? s.hA[n - 1] : s.hB[n - 1]), b2, f2);
    }
    o1 = __fadd2_rn(e1, f1);
    o2 = __fadd2_rn(e2, f2);
}

}  // namespace lcf
