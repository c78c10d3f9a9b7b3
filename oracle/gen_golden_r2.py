#!/usr/bin/env python
"""Golden vectors added in round 2, generated HERE by the real reference compiled into oracle/_ref (make -C oracle ref;
the baseline x86-64 build, no FMA contraction) and committed as tests/golden/ref_outputs_r2.npz.  /root/reference does
not exist on the GPU box; this script is the record of how the fixture was made.

  osync_*  ofdmframesync (src/multichannel/src/ofdmframesync.c): frames made by the reference's ofdmframegen (S0a S0b S1, QPSK
           data symbols, tail), passed through a channel (delay, gain, carrier offset, a little seeded noise) and fed to the
           reference's ofdmframesync in uneven chunks; every row the callback received is recorded.  One case has two
           frames in the stream and a callback that ends each frame after its last data symbol (return 1)
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
    # ---- ofdmframesync
    CB = C.CFUNCTYPE(C.c_int, C.POINTER(_cf), C.POINTER(C.c_ubyte), C.c_uint, C.c_void_p)
    R.ofdmframegen_create.restype = C.c_void_p
    R.ofdmframegen_create.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
    R.ofdmframesync_create.restype = C.c_void_p
    R.ofdmframesync_create.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_void_p, CB, C.c_void_p]
    for fn in ("ofdmframegen_write_S0a", "ofdmframegen_write_S0b", "ofdmframegen_write_S1", "ofdmframegen_writetail"):
        getattr(R, fn).argtypes = [C.c_void_p, C.c_void_p]
    R.ofdmframegen_writesymbol.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    R.ofdmframegen_destroy.argtypes = [C.c_void_p]
    R.ofdmframegen_reset.argtypes = [C.c_void_p]
    R.ofdmframesync_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    R.ofdmframesync_destroy.argtypes = [C.c_void_p]
    R.ofdmframesync_get_cfo.restype = C.c_float
    R.ofdmframesync_get_cfo.argtypes = [C.c_void_p]
    rng = np.random.default_rng(20261021)

    def make_frame(M, cp, taper, nsym):
        g = R.ofdmframegen_create(M, cp, taper, None)
        buf = np.empty(M + cp, np.complex64)
        parts = []
        for fn in ("ofdmframegen_write_S0a", "ofdmframegen_write_S0b", "ofdmframegen_write_S1"):
            getattr(R, fn)(g, buf.ctypes.data)
            parts.append(buf.copy())
        for _ in range(nsym):
            X = ((rng.integers(0, 2, M) * 2 - 1) + 1j * (rng.integers(0, 2, M) * 2 - 1)).astype(np.complex64) * np.float32(np.sqrt(0.5))
            R.ofdmframegen_writesymbol(g, X.ctypes.data, buf.ctypes.data)
            parts.append(buf.copy())
        tail = np.zeros(max(taper, 1), np.complex64)
        R.ofdmframegen_writetail(g, tail.ctypes.data)
        parts.append(tail[:taper])
        R.ofdmframegen_destroy(g)
        return np.concatenate(parts)

    def channel(x, delay, gain, dphi, phi, nstd):
        n = np.arange(x.size + delay)
        y = np.concatenate([np.zeros(delay, np.complex64), x]).astype(np.complex128)
        y = gain * y * np.exp(1j * (dphi * n + phi))
        y = y + nstd * (rng.standard_normal(y.size) + 1j * rng.standard_normal(y.size)) / np.sqrt(2)
        return y.astype(np.complex64)

    def run_ref_sync(M, cp, taper, y, chunks, end_after=None):
        rows = []

        def cb(yp, pp, m, ud):
            rows.append(np.array([complex(yp[i].re, yp[i].im) for i in range(m)], np.complex64))
            if end_after is not None and len(rows) % end_after == 0:
                return 1
            return 0
        cbk = CB(cb)
        q = R.ofdmframesync_create(M, cp, taper, None, cbk, None)
        pos = 0
        k = 0
        while pos < y.size:
            c = chunks[k % len(chunks)]
            k += 1
            seg = np.ascontiguousarray(y[pos: pos + c])
            R.ofdmframesync_execute(q, seg.ctypes.data, seg.size)
            pos += c
        cfo = R.ofdmframesync_get_cfo(q)
        R.ofdmframesync_destroy(q)
        return np.array(rows, np.complex64).reshape(-1, M), cfo

    for tag, M, cp, taper, nsym, delay, gain, dphi, nstd, chunks in [
            ("osync_M64", 64, 16, 4, 12, 37, 0.8, 0.02, 1e-3, [1000]),
            ("osync_M256", 256, 32, 8, 9, 301, 1.7, -0.004, 1e-3, [777, 13, 2048]),
            ("osync_M48", 48, 10, 0, 7, 5, 1.0, 0.01, 0.0, [64]),
            ("osync_M1024", 1024, 64, 16, 5, 100, 0.5, 0.001, 1e-3, [4096, 100])]:
        x = make_frame(M, cp, taper, nsym)
        y = channel(x, delay, gain, dphi, 0.7, nstd)
        y = np.concatenate([y, np.zeros(3 * (M + cp), np.complex64)])
        # stop the stream right after the last data symbol has been delivered: what follows a frame is noise to the receiver
        rows, cfo = run_ref_sync(M, cp, taper, y, chunks)
        out[tag + "_args"] = np.array([M, cp, taper, nsym, cfo], np.float64)
        out[tag + "_chunks"] = np.array(chunks, np.int64)
        out[tag + "_y"] = y
        out[tag + "_rows"] = rows[:nsym]
        print(tag, "rows", rows.shape, "cfo", cfo)
    # two frames, the callback ends each after its nsym-th symbol
    M, cp, taper, nsym = 64, 16, 4, 6
    f1, f2 = make_frame(M, cp, taper, nsym), make_frame(M, cp, taper, nsym)
    y = np.concatenate([channel(f1, 50, 1.0, 0.015, 0.3, 1e-3), np.zeros(200, np.complex64), channel(f2, 20, 0.9, -0.01, 1.1, 1e-3),
                        np.zeros(400, np.complex64)])
    rows, cfo = run_ref_sync(M, cp, taper, y, [500, 333], end_after=nsym)
    out["osync_two_args"] = np.array([M, cp, taper, nsym, cfo], np.float64)
    out["osync_two_chunks"] = np.array([500, 333], np.int64)
    out["osync_two_y"] = y
    out["osync_two_rows"] = rows
    print("osync_two rows", rows.shape)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_outputs_r2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
