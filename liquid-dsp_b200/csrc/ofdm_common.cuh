// ofdm_common.cuh -- what the OFDM frame generator (ofdmframe.cu) and synchronizer (ofdmframesync.cu) share:
// the default m-sequences (src/sequence/src/msequence.c), the PLCP training symbols in the frequency domain
// (src/multichannel/src/ofdmframe.common.c:41-167) and the subcarrier types.
#pragma once
#include <vector>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"

namespace lcofdm {

constexpr unsigned int kPilotPeriod = 255;

// default m-sequence of register length m: generator (already shifted right by one bit) and start state 1 << (m-1)
struct Mseq {
    unsigned int g, n, v;
    explicit Mseq(unsigned int m)
    {
        static const unsigned int poly[16] = {0, 0, 0x0003, 0x0005, 0x0009, 0x0012, 0x0021, 0x0044, 0x008E,
                                              0x0108, 0x0204, 0x0402, 0x0829, 0x100d, 0x2015, 0x4001};
        g = poly[m];
        n = (1u << m) - 1;
        v = 1u << (m - 1);
    }
    unsigned int advance()
    {
        const unsigned int b = (unsigned int)__builtin_parity(v & g);
        v = ((v << 1) | b) & n;
        return b;
    }
    unsigned int symbol(unsigned int bps)
    {
        unsigned int s = 0;
        for (unsigned int i = 0; i < bps; i++) s = (s << 1) | advance();
        return s;
    }
};

inline unsigned int nextpow2(unsigned int x)
{
    unsigned int n = 0;
    x--;
    while (x > 0) { x >>= 1; n++; }
    return n;
}

// PLCP sequence (which = 0 short: even enabled subcarriers only; 1 long: register one longer)   ofdmframe.common.c:41-167
inline unsigned int plcp_freq(const std::vector<unsigned char> &p, unsigned int M, int which, std::vector<lc_cf> &S)
{
    unsigned int m = nextpow2(M);
    if (m < 4) m = 4;
    else if (m > 8) m = 8;
    if (which) m++;
    Mseq ms(m);
    unsigned int enabled = 0;
    S.assign(M, lc_cf{0.f, 0.f});
    for (unsigned int i = 0; i < M; i++) {
        const unsigned int b = ms.symbol(3) & 1;
        if (p[i] == OFDMFRAME_SCTYPE_NULL || (!which && (i % 2))) continue;
        S[i].re = b ? 1.0f : -1.0f;
        enabled++;
    }
    return enabled;
}

}  // namespace lcofdm
