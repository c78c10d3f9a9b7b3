"""bit-for-bit comparison of two code paths of the fused M=1024 analyzer selected by environment variables
usage: python tools/x_v8check.py VAR=VALUE [m ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import liquid_dsp_b200 as L
from oracle import oracle as O

var, val = sys.argv[1].split("=")
ms = [int(a) for a in sys.argv[2:]] or [7, 2, 4, 5]
ok = True
for m in ms:
    M = 1024
    nfr = 33 + 100 + 3015
    x = O.synth_input(nfr * M // 2, seed=0xF0 + m)
    outs = []
    for setit in (False, True):
        if setit: os.environ[var] = val
        else: os.environ.pop(var, None)
        q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
        got = np.concatenate([q.execute_block(x[:33 * 512]), q.execute_block(x[33 * 512:133 * 512]), q.execute_block(x[133 * 512:])])
        q.destroy()
        outs.append(got)
    same = np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    d = float(np.abs(outs[0] - outs[1]).max())
    print("m=%d: %s identical=%s max|diff|=%.3e finite=%s" % (m, sys.argv[1], same, d, bool(np.isfinite(outs[1]).all())))
    ok &= same
sys.exit(0 if ok else 1)
