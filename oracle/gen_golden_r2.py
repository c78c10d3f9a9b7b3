#!/usr/bin/env python
"""Golden vectors added in round 2, generated HERE by the real reference compiled into oracle/_ref (make -C oracle ref;
the baseline x86-64 build, no FMA contraction) and committed as tests/golden/ref_outputs_r2.npz.  /root/reference does
not exist on the GPU box; this script is the record of how the fixture was made.

  pfbd_*   firpfb_crcf_create_rnyquist / _create_drnyquist(type, M, k, m, beta) (src/filter/src/firpfb.c:146-239): a
           seeded input pushed sample by sample, one output per push from filter index (7 n + 3) % M
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ["LIQUID_ORACLE_REF_VARIANT"] = "_sse"
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402


class _cf(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


def main():
    R = O.ref()
    out = {}
    for fn in ("firpfb_crcf_create_rnyquist", "firpfb_crcf_create_drnyquist"):
        getattr(R, fn).restype = C.c_void_p
        getattr(R, fn).argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_float]
    R.firpfb_crcf_push.argtypes = [C.c_void_p, _cf]
    R.firpfb_crcf_execute.argtypes = [C.c_void_p, C.c_uint, C.POINTER(_cf)]
    R.firpfb_crcf_destroy.argtypes = [C.c_void_p]
    # (ftype, M, k, m, beta): LIQUID_FIRFILT_ARKAISER = 7, RKAISER = 8, RRC = 9 (include/liquid.h:1415-1435)
    for ftype, M, k, m, beta in [(9, 16, 2, 3, 0.3), (7, 32, 2, 4, 0.25), (8, 8, 4, 2, 0.5), (9, 64, 2, 5, 0.2)]:
        x = O.synth_input(300, seed=0x5EEDE000 + M)
        for kind in ("rnyquist", "drnyquist"):
            q = getattr(R, "firpfb_crcf_create_" + kind)(ftype, M, k, m, beta)
            y = np.empty(x.size, np.complex64)
            v = _cf()
            for n in range(x.size):
                R.firpfb_crcf_push(q, _cf(float(x[n].real), float(x[n].imag)))
                R.firpfb_crcf_execute(q, (7 * n + 3) % M, C.byref(v))
                y[n] = complex(v.re, v.im)
            R.firpfb_crcf_destroy(q)
            tag = "pfbd_%s_t%d_M%d" % (kind, ftype, M)
            out[tag + "_args"] = np.array([ftype, M, k, m, beta], np.float64)
            out[tag + "_x"] = x
            out[tag + "_y"] = y
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_outputs_r2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
