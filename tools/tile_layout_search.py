#!/usr/bin/env python
"""Shared-memory layout search for pfbch2_tile.cu: the tile of 8 frames x 2 sub-problems x (16 rows x R2 columns) of
float2, address (8-byte units) = f*FS + p*SS + row*PR + col.  For every access pattern of the kernel count the
wavefronts of a 64-bit warp access (half-warps of 16 lanes over 16 8-byte banks) and pick the (PR, SS, FS) with
the fewest, smallest footprint first."""
import itertools
import sys


def degree(addrs):
    # addrs: 16 unit addresses of a half-warp -> wavefronts (same address = broadcast)
    banks = {}
    for a in set(addrs):
        banks.setdefault(a % 16, set()).add(a)
    return max(len(v) for v in banks.values())


def cost(M, PR, SS, FS, wide):
    M2 = M // 2
    R2 = M2 // 16
    A = lambda f, p, r, c: f * FS + p * SS + r * PR + c
    tot, ideal = 0, 0
    lanes_all = range(min(M2, 256))
    def warps(fn):
        nonlocal tot, ideal
        for h0 in range(0, M2, 16):
            tot += degree([fn(u) for u in range(h0, h0 + 16)])
            ideal += 1
    # phase 1 stores (per frame f, sub-problem p)
    for p in range(2):
        warps(lambda u: A(0, p, u // R2, u % R2))
    # pass A (read and write), grouped mapping
    g = lambda u: (u // (2 * R2), (u % (2 * R2)) >> 1, u & 1)
    for r1 in range(16):
        for _ in range(2):
            warps(lambda u: A(g(u)[0], g(u)[2], r1, g(u)[1]))
    # pass B reads
    if wide:
        for i in range(256 // M2):
            for c in range(R2):
                def fn(u):
                    j = u + M2 * i
                    return A(j // 32, j & 1, (j % 32) >> 1, c)
                warps(fn)
    else:
        for i in range(16 // R2):
            for c in range(R2):
                warps(lambda u: A(g(u)[0], g(u)[2], g(u)[1] + R2 * i, c))
    return tot, ideal


for M in (64, 128, 256, 512):
    R2 = M // 32
    for wide in (0, 1):
        best = None
        for PR in range(R2, R2 + 3):
            span = 15 * PR + R2
            for SS in range(span, span + 17):
                for FS in range(SS + span, SS + span + 17):
                    t, i = cost(M, PR, SS, FS, wide)
                    key = (t, FS)
                    if best is None or key < best[0]:
                        best = (key, PR, SS, FS, i)
        print("M=%d wide=%d: PR=%d SS=%d FS=%d wavefronts %d (ideal %d), tile = %d units" % (M, wide, best[1], best[2], best[3], best[0][0], best[4], 8 * best[3]))
    t, i = cost(M, R2 + 1, 16 * (R2 + 1), 32 * (R2 + 1), 0)
    print("   current layout: %d (ideal %d)" % (t, i))
