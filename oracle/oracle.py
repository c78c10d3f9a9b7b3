"""ctypes front-end to the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  * ``liboracle.so``             -- our C restatement (oracle/liquid_oracle.c)
  * ``_ref/libliquid_ref*.so``   -- the real reference compiled from /root/reference by
                                    oracle/Makefile (present when built in the dev container;
                                    travels to the GPU box as a prebuilt .so)

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module, and only as the checker -- never as the thing measured or shipped.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ANALYZER, SYNTHESIZER = 0, 1
FFT_FORWARD, FFT_BACKWARD = +1, -1

c_cf_p = np.ctypeslib.ndpointer(dtype=np.complex64, flags="C_CONTIGUOUS")
c_f_p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


class c_cf(C.Structure):
    """complex float passed by value ({re, im}: same SysV class as C99 `float complex`)"""
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


def build(ref=True):
    """(re)build liboracle.so and, if /root/reference is present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


# ----------------------------------------------------------------------------------------
# our restatement
# ----------------------------------------------------------------------------------------
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.orc_lngammaf.restype = C.c_float; L.orc_lngammaf.argtypes = [C.c_float]
        L.orc_besseli0f.restype = C.c_float; L.orc_besseli0f.argtypes = [C.c_float]
        L.orc_sincf.restype = C.c_float; L.orc_sincf.argtypes = [C.c_float]
        L.orc_kaiser.restype = C.c_float; L.orc_kaiser.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float]
        L.orc_kaiser_beta_As.restype = C.c_float; L.orc_kaiser_beta_As.argtypes = [C.c_float]
        L.orc_msb_index.restype = C.c_uint; L.orc_msb_index.argtypes = [C.c_uint]
        L.orc_firdes_kaiser.restype = C.c_int
        L.orc_firdes_kaiser.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_float, c_f_p]
        L.orc_firdes_rrcos.restype = C.c_int
        L.orc_firdes_rrcos.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float, c_f_p]
        L.orc_firdes_arkaiser.restype = C.c_int
        L.orc_firdes_arkaiser.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float, c_f_p]
        L.orc_firdes_rkaiser.restype = C.c_int
        L.orc_firdes_rkaiser.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float, c_f_p, C.POINTER(C.c_float)]
        L.orc_estimate_req_filter_As.restype = C.c_float
        L.orc_estimate_req_filter_As.argtypes = [C.c_float, C.c_uint]
        L.orc_rkaiser_approximate_rho.restype = C.c_float
        L.orc_rkaiser_approximate_rho.argtypes = [C.c_uint, C.c_float]
        L.orc_filter_isi.restype = None
        L.orc_filter_isi.argtypes = [c_f_p, C.c_uint, C.c_uint, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        for name in ("orc_dotprod_crcf_run", "orc_dotprod_crcf_run4", "orc_dotprod_crcf_run_sse_order"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [c_f_p, c_cf_p, C.c_uint, c_cf_p]
        L.orc_fft_run.restype = None; L.orc_fft_run.argtypes = [C.c_uint, c_cf_p, c_cf_p, C.c_int]
        L.orc_firfilt_create.restype = C.c_void_p; L.orc_firfilt_create.argtypes = [c_f_p, C.c_uint]
        L.orc_firfilt_create_kaiser.restype = C.c_void_p
        L.orc_firfilt_create_kaiser.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_float]
        L.orc_firfilt_destroy.argtypes = [C.c_void_p]; L.orc_firfilt_reset.argtypes = [C.c_void_p]
        L.orc_firfilt_set_scale.argtypes = [C.c_void_p, C.c_float]
        L.orc_firfilt_execute_block.restype = None
        L.orc_firfilt_execute_block.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
        L.orc_firpfb_create.restype = C.c_void_p; L.orc_firpfb_create.argtypes = [C.c_uint, c_f_p, C.c_uint]
        L.orc_firpfb_create_kaiser.restype = C.c_void_p
        L.orc_firpfb_create_kaiser.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float]
        L.orc_firpfb_destroy.argtypes = [C.c_void_p]; L.orc_firpfb_reset.argtypes = [C.c_void_p]
        L.orc_firpfb_set_scale.argtypes = [C.c_void_p, C.c_float]
        L.orc_firpfb_push.argtypes = [C.c_void_p, c_cf]
        L.orc_firpfb_execute.argtypes = [C.c_void_p, C.c_uint, c_cf_p]
        L.orc_firpfbch_create.restype = C.c_void_p
        L.orc_firpfbch_create.argtypes = [C.c_int, C.c_uint, C.c_uint, c_f_p]
        L.orc_firpfbch_create_kaiser.restype = C.c_void_p
        L.orc_firpfbch_create_kaiser.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float]
        L.orc_firpfbch_create_rrc.restype = C.c_void_p
        L.orc_firpfbch_create_rrc.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float]
        L.orc_firpfbch_create_rnyquist.restype = C.c_void_p
        L.orc_firpfbch_create_rnyquist.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float, C.c_int]
        L.orc_firpfbch_cccf_create.restype = C.c_void_p
        L.orc_firpfbch_cccf_create.argtypes = [C.c_int, C.c_uint, C.c_uint, c_cf_p]
        L.orc_firpfbch_cccf_destroy.argtypes = [C.c_void_p]
        L.orc_firpfbch_cccf_execute_frames.restype = None
        L.orc_firpfbch_cccf_execute_frames.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
        for nm in ("firdecim", "firinterp"):
            getattr(L, "orc_%s_create" % nm).restype = C.c_void_p
            getattr(L, "orc_%s_create" % nm).argtypes = [C.c_uint, c_f_p, C.c_uint]
            getattr(L, "orc_%s_create_kaiser" % nm).restype = C.c_void_p
            getattr(L, "orc_%s_create_kaiser" % nm).argtypes = [C.c_uint, C.c_uint, C.c_float]
            getattr(L, "orc_%s_destroy" % nm).argtypes = [C.c_void_p]
            getattr(L, "orc_%s_execute_block" % nm).restype = None
            getattr(L, "orc_%s_execute_block" % nm).argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
        L.orc_firdecim_clear.argtypes = [C.c_void_p]
        L.orc_estimate_req_filter_len.restype = C.c_uint
        L.orc_estimate_req_filter_len.argtypes = [C.c_float, C.c_float]
        L.orc_resamp2_create.restype = C.c_void_p
        L.orc_resamp2_create.argtypes = [C.c_uint, C.c_float, C.c_float]
        L.orc_resamp2_destroy.argtypes = [C.c_void_p]
        L.orc_resamp2_clear.argtypes = [C.c_void_p]
        L.orc_resamp2_get_taps.argtypes = [C.c_void_p, c_f_p]
        L.orc_resamp2_filter_execute.restype = None
        L.orc_resamp2_filter_execute.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p, c_cf_p]
        for nm in ("analyzer", "synthesizer", "decim", "interp"):
            getattr(L, "orc_resamp2_%s_execute" % nm).restype = None
            getattr(L, "orc_resamp2_%s_execute" % nm).argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
        L.orc_msresamp2_create.restype = C.c_void_p
        L.orc_msresamp2_create.argtypes = [C.c_int, C.c_uint, C.c_float, C.c_float, C.c_float]
        L.orc_msresamp2_destroy.argtypes = [C.c_void_p]
        L.orc_msresamp2_reset.argtypes = [C.c_void_p]
        L.orc_msresamp2_stage_m.restype = C.c_uint
        L.orc_msresamp2_stage_m.argtypes = [C.c_void_p, C.c_uint]
        L.orc_msresamp2_get_delay.restype = C.c_float
        L.orc_msresamp2_get_delay.argtypes = [C.c_void_p]
        L.orc_asgramcf_create.restype = C.c_void_p
        L.orc_asgramcf_create.argtypes = [C.c_uint]
        L.orc_asgramcf_destroy.argtypes = [C.c_void_p]
        L.orc_asgramcf_reset.argtypes = [C.c_void_p]
        L.orc_asgramcf_set_scale.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.orc_asgramcf_write.argtypes = [C.c_void_p, c_cf_p, C.c_uint]
        L.orc_asgramcf_execute.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        c_u8_p = np.ctypeslib.ndpointer(np.uint8, flags="C")
        L.orc_ofdmframe_init_default_sctype.argtypes = [C.c_uint, c_u8_p]
        L.orc_ofdmframegen_create.restype = C.c_void_p
        L.orc_ofdmframegen_create.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
        L.orc_ofdmframegen_destroy.argtypes = [C.c_void_p]
        L.orc_ofdmframegen_reset.argtypes = [C.c_void_p]
        L.orc_ofdmframegen_get_sctype.argtypes = [C.c_void_p, c_u8_p]
        for nm in ("write_S0a", "write_S0b", "write_S1", "writetail"):
            getattr(L, "orc_ofdmframegen_" + nm).argtypes = [C.c_void_p, c_cf_p]
        L.orc_ofdmframegen_writesymbols.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
        L.orc_fftfilt_create.restype = C.c_void_p
        L.orc_fftfilt_create.argtypes = [c_f_p, C.c_uint, C.c_uint]
        L.orc_fftfilt_destroy.argtypes = [C.c_void_p]
        L.orc_fftfilt_reset.argtypes = [C.c_void_p]
        L.orc_fftfilt_set_scale.argtypes = [C.c_void_p, C.c_float]
        L.orc_fftfilt_execute_blocks.restype = None
        L.orc_fftfilt_execute_blocks.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
        L.orc_msresamp2_execute_block.restype = None
        L.orc_msresamp2_execute_block.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
        L.orc_firinterp_reset.argtypes = [C.c_void_p]
        L.orc_spgramcf_create.restype = C.c_void_p
        L.orc_spgramcf_create.argtypes = [C.c_uint, c_f_p, C.c_uint]
        L.orc_spgramcf_create_kaiser.restype = C.c_void_p
        L.orc_spgramcf_create_kaiser.argtypes = [C.c_uint, C.c_uint, C.c_float]
        L.orc_spgramcf_create_default.restype = C.c_void_p
        L.orc_spgramcf_create_default.argtypes = [C.c_uint]
        for fn in ("destroy", "reset"):
            getattr(L, "orc_spgramcf_" + fn).argtypes = [C.c_void_p]
        L.orc_spgramcf_push.argtypes = [C.c_void_p, c_cf]
        L.orc_spgramcf_write.argtypes = [C.c_void_p, c_cf_p, C.c_uint]
        L.orc_spgramcf_execute.argtypes = [C.c_void_p, c_cf_p]
        L.orc_spgramcf_execute_psd.argtypes = [C.c_void_p, c_f_p]
        L.orc_spgramcf_accumulate_psd.restype = C.c_int
        L.orc_spgramcf_accumulate_psd.argtypes = [C.c_void_p, c_cf_p, C.c_float, C.c_uint]
        L.orc_spgramcf_write_accumulation.argtypes = [C.c_void_p, c_f_p]
        L.orc_spgramcf_estimate_psd.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_f_p]
        L.orc_spgramcf_num_transforms.restype = C.c_uint
        L.orc_spgramcf_num_transforms.argtypes = [C.c_void_p]
        L.orc_firpfbch_destroy.argtypes = [C.c_void_p]; L.orc_firpfbch_reset.argtypes = [C.c_void_p]
        L.orc_firpfbch_execute_frames.restype = None
        L.orc_firpfbch_execute_frames.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
        L.orc_firpfbch_analyzer_run.restype = None
        L.orc_firpfbch_analyzer_run.argtypes = [C.c_void_p, C.c_uint, c_cf_p]
        L.orc_firpfbch2_create.restype = C.c_void_p
        L.orc_firpfbch2_create.argtypes = [C.c_int, C.c_uint, C.c_uint, c_f_p]
        L.orc_firpfbch2_create_kaiser.restype = C.c_void_p
        L.orc_firpfbch2_create_kaiser.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float]
        L.orc_firpfbch2_destroy.argtypes = [C.c_void_p]; L.orc_firpfbch2_reset.argtypes = [C.c_void_p]
        L.orc_firpfbch2_execute_frames.restype = None
        L.orc_firpfbch2_execute_frames.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
        L.orc_synth_input.restype = None
        L.orc_synth_input.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, c_cf_p]
        L.orc_bench_firpfbch2_analysis.restype = C.c_double
        L.orc_bench_firpfbch2_analysis.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_ulonglong, C.c_uint]
        _lib = L
    return _lib


def _cf(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def firdes_kaiser(n, fc, As, mu=0.0):
    h = np.empty(n, np.float32)
    if lib().orc_firdes_kaiser(n, fc, As, mu, h):
        raise ValueError("liquid_firdes_kaiser: invalid arguments")
    return h


def firdes_rrcos(k, m, beta, dt=0.0):
    h = np.empty(2 * k * m + 1, np.float32)
    if lib().orc_firdes_rrcos(k, m, beta, dt, h):
        raise ValueError("liquid_firdes_rrcos: invalid arguments")
    return h


def firdes_arkaiser(k, m, beta, dt=0.0):
    h = np.empty(2 * k * m + 1, np.float32)
    if lib().orc_firdes_arkaiser(k, m, beta, dt, h):
        raise ValueError("liquid_firdes_arkaiser: invalid arguments")
    return h


def firdes_rkaiser(k, m, beta, dt=0.0, return_rho=False):
    h = np.empty(2 * k * m + 1, np.float32)
    rho = C.c_float(0.0)
    if lib().orc_firdes_rkaiser(k, m, beta, dt, h, C.byref(rho)):
        raise ValueError("liquid_firdes_rkaiser: invalid arguments")
    return (h, rho.value) if return_rho else h


def filter_isi(h, k, m):
    rms, peak = C.c_float(0.0), C.c_float(0.0)
    lib().orc_filter_isi(_f(h), k, m, C.byref(rms), C.byref(peak))
    return rms.value, peak.value


def dotprod_crcf(h, x, order="run"):
    y = np.zeros(1, np.complex64)
    fn = {"run": "orc_dotprod_crcf_run", "run4": "orc_dotprod_crcf_run4",
          "sse": "orc_dotprod_crcf_run_sse_order"}[order]
    getattr(lib(), fn)(_f(h), _cf(x), len(h), y)
    return y[0]


def fft(x, direction=FFT_FORWARD):
    x = _cf(x)
    y = np.empty_like(x)
    lib().orc_fft_run(x.size, x, y, direction)
    return y


def synth_input(n, seed=0x5EED0000, first=0):
    x = np.empty(n, np.complex64)
    lib().orc_synth_input(seed, first, n, x)
    return x


class _Obj:
    _destroy = None

    def __init__(self, handle):
        if not handle:
            raise ValueError("%s: invalid arguments" % type(self).__name__)
        self.h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            getattr(_lib, self._destroy)(self.h)
            self.h = None


class Firfilt(_Obj):
    _destroy = "orc_firfilt_destroy"

    def __init__(self, h=None, kaiser=None):
        super().__init__(lib().orc_firfilt_create(_f(h), len(h)) if kaiser is None
                         else lib().orc_firfilt_create_kaiser(*kaiser))

    def reset(self):
        lib().orc_firfilt_reset(self.h)

    def set_scale(self, s):
        lib().orc_firfilt_set_scale(self.h, s)

    def execute_block(self, x):
        x = _cf(x)
        y = np.empty_like(x)
        lib().orc_firfilt_execute_block(self.h, x, x.size, y)
        return y


class Firpfb(_Obj):
    _destroy = "orc_firpfb_destroy"

    def __init__(self, M, h=None, kaiser=None):
        super().__init__(lib().orc_firpfb_create(M, _f(h), len(h)) if kaiser is None
                         else lib().orc_firpfb_create_kaiser(M, *kaiser))

    def reset(self):
        lib().orc_firpfb_reset(self.h)

    def set_scale(self, s):
        lib().orc_firpfb_set_scale(self.h, s)

    def push(self, v):
        v = complex(v)
        lib().orc_firpfb_push(self.h, c_cf(v.real, v.imag))

    def execute(self, i):
        y = np.zeros(1, np.complex64)
        lib().orc_firpfb_execute(self.h, i, y)
        return y[0]


class Firpfbch(_Obj):
    _destroy = "orc_firpfbch_destroy"

    def __init__(self, type, M, p=None, h=None, kaiser=None, rrc=None, rnyquist=None):
        self.type, self.M = type, M
        if kaiser is not None:
            hd = lib().orc_firpfbch_create_kaiser(type, M, *kaiser)
        elif rnyquist is not None:                      # (m, beta, ftype)
            hd = lib().orc_firpfbch_create_rnyquist(type, M, *rnyquist)
        elif rrc is not None:
            hd = lib().orc_firpfbch_create_rrc(type, M, *rrc)
        else:
            hd = lib().orc_firpfbch_create(type, M, p, _f(h))
        super().__init__(hd)

    def reset(self):
        lib().orc_firpfbch_reset(self.h)

    def execute_frames(self, x):
        x = _cf(x).reshape(-1, self.M)
        y = np.empty_like(x)
        lib().orc_firpfbch_execute_frames(self.h, x, x.shape[0], y)
        return y


class Firdecim(_Obj):
    """firdecim_crcf: n outputs from n*M inputs"""
    _destroy = "orc_firdecim_destroy"

    def __init__(self, M, h=None, kaiser=None):
        self.M = M
        super().__init__(lib().orc_firdecim_create(M, _f(h), len(h)) if kaiser is None
                         else lib().orc_firdecim_create_kaiser(M, *kaiser))

    def clear(self):
        lib().orc_firdecim_clear(self.h)

    def execute_block(self, x):
        x = _cf(x)
        n = len(x) // self.M
        y = np.empty(n, np.complex64)
        lib().orc_firdecim_execute_block(self.h, x, n, y)
        return y


class Firinterp(_Obj):
    """firinterp_crcf: n*M outputs from n inputs"""
    _destroy = "orc_firinterp_destroy"

    def __init__(self, M, h=None, kaiser=None):
        self.M = M
        super().__init__(lib().orc_firinterp_create(M, _f(h), len(h)) if kaiser is None
                         else lib().orc_firinterp_create_kaiser(M, *kaiser))

    def reset(self):
        lib().orc_firinterp_reset(self.h)

    def execute_block(self, x):
        x = _cf(x)
        y = np.empty(len(x) * self.M, np.complex64)
        lib().orc_firinterp_execute_block(self.h, x, len(x), y)
        return y


RESAMP_INTERP, RESAMP_DECIM = 0, 1      # liquid.h:3023-3024


class Resamp2(_Obj):
    """resamp2_crcf: half-band filter / analyzer / synthesizer / decimator / interpolator, block forms"""
    _destroy = "orc_resamp2_destroy"

    def __init__(self, m, f0=0.0, As=60.0):
        self.m = m
        super().__init__(lib().orc_resamp2_create(m, f0, As))

    def clear(self):
        lib().orc_resamp2_clear(self.h)

    def taps(self):
        hb = np.empty(2 * self.m, np.float32)
        lib().orc_resamp2_get_taps(self.h, hb)
        return hb

    def filter_execute(self, x):
        x = _cf(x)
        y0 = np.empty(len(x), np.complex64); y1 = np.empty(len(x), np.complex64)
        lib().orc_resamp2_filter_execute(self.h, x, len(x), y0, y1)
        return y0, y1

    def _run(self, name, x, n, n_out):
        x = _cf(x)
        y = np.empty(n_out, np.complex64)
        getattr(lib(), "orc_resamp2_%s_execute" % name)(self.h, x, n, y)
        return y

    def analyzer_execute(self, x):
        return self._run("analyzer", x, len(x) // 2, len(x) // 2 * 2)

    def synthesizer_execute(self, x):
        return self._run("synthesizer", x, len(x) // 2, len(x) // 2 * 2)

    def decim_execute(self, x):
        return self._run("decim", x, len(x) // 2, len(x) // 2)

    def interp_execute(self, x):
        return self._run("interp", x, len(x), 2 * len(x))


class Msresamp2(_Obj):
    """msresamp2_crcf: cascade of S half-band stages, rate 2^S"""
    _destroy = "orc_msresamp2_destroy"

    def __init__(self, type, S, fc, f0, As):
        self.type, self.S = type, S
        super().__init__(lib().orc_msresamp2_create(type, S, fc, f0, As))

    def reset(self):
        lib().orc_msresamp2_reset(self.h)

    def stage_m(self):
        return [lib().orc_msresamp2_stage_m(self.h, i) for i in range(self.S)]

    def get_delay(self):
        return lib().orc_msresamp2_get_delay(self.h)

    def execute_block(self, x):
        x = _cf(x)
        M = 1 << self.S
        n = len(x) if self.type == RESAMP_INTERP else len(x) // M
        y = np.empty(n * M if self.type == RESAMP_INTERP else n, np.complex64)
        lib().orc_msresamp2_execute_block(self.h, x, n, y)
        return y


class Asgramcf(_Obj):
    """asgramcf: ASCII spectrogram; execute() -> (string, peak dB, peak frequency)"""
    _destroy = "orc_asgramcf_destroy"

    def __init__(self, nfft):
        self.nfft = nfft
        super().__init__(lib().orc_asgramcf_create(nfft))

    def reset(self):
        lib().orc_asgramcf_reset(self.h)

    def set_scale(self, offset, scale):
        lib().orc_asgramcf_set_scale(self.h, offset, scale)

    def write(self, x):
        x = _cf(x)
        lib().orc_asgramcf_write(self.h, x, len(x))

    def execute(self):
        buf = C.create_string_buffer(self.nfft + 1)
        pv, pf = C.c_float(), C.c_float()
        lib().orc_asgramcf_execute(self.h, buf, C.byref(pv), C.byref(pf))
        return buf.raw[:self.nfft].decode("ascii"), pv.value, pf.value


class Ofdmframegen(_Obj):
    """ofdmframegen: S0a / S0b / S1 / n data symbols / tail; p = subcarrier allocation (uint8) or None (default)"""
    _destroy = "orc_ofdmframegen_destroy"

    def __init__(self, M, cp_len, taper_len, p=None):
        self.M, self.cp, self.tl = M, cp_len, taper_len
        self._p = None if p is None else np.ascontiguousarray(p, np.uint8)
        super().__init__(lib().orc_ofdmframegen_create(M, cp_len, taper_len,
                                                       None if self._p is None else self._p.ctypes.data))

    def reset(self):
        lib().orc_ofdmframegen_reset(self.h)

    def sctype(self):
        p = np.empty(self.M, np.uint8)
        lib().orc_ofdmframegen_get_sctype(self.h, p)
        return p

    def _sym(self, name):
        y = np.empty(self.M + self.cp, np.complex64)
        getattr(lib(), "orc_ofdmframegen_" + name)(self.h, y)
        return y

    def write_S0a(self):
        return self._sym("write_S0a")

    def write_S0b(self):
        return self._sym("write_S0b")

    def write_S1(self):
        return self._sym("write_S1")

    def writesymbols(self, x):
        x = _cf(x)
        n = len(x) // self.M
        y = np.empty(n * (self.M + self.cp), np.complex64)
        lib().orc_ofdmframegen_writesymbols(self.h, x, n, y)
        return y

    def writetail(self):
        y = np.empty(max(self.tl, 1), np.complex64)
        lib().orc_ofdmframegen_writetail(self.h, y)
        return y[:self.tl]


def ofdm_frame(q, x):
    """a whole frame from an object with the methods above: S0a | S0b | S1 | data symbols | tail"""
    return np.concatenate([q.write_S0a(), q.write_S0b(), q.write_S1(), q.writesymbols(x), q.writetail()])


class Fftfilt(_Obj):
    """fftfilt_crcf: overlap-add FIR on blocks of n samples"""
    _destroy = "orc_fftfilt_destroy"

    def __init__(self, h, n):
        self.n = n
        super().__init__(lib().orc_fftfilt_create(_f(h), len(h), n))

    def reset(self):
        lib().orc_fftfilt_reset(self.h)

    def set_scale(self, scale):
        lib().orc_fftfilt_set_scale(self.h, scale)

    def execute_blocks(self, x):
        x = _cf(x)
        nb = len(x) // self.n
        y = np.empty(nb * self.n, np.complex64)
        lib().orc_fftfilt_execute_blocks(self.h, x, nb, y)
        return y


class Spgramcf(_Obj):
    """spgramcf: windowed-FFT periodogram; create(nfft, window=...) | (nfft, window_len=, beta=) | (nfft)"""
    _destroy = "orc_spgramcf_destroy"

    def __init__(self, nfft, window=None, window_len=None, beta=None):
        self.nfft = nfft
        if window is not None:
            hd = lib().orc_spgramcf_create(nfft, _f(window), len(window))
        elif window_len is not None:
            hd = lib().orc_spgramcf_create_kaiser(nfft, window_len, beta)
        else:
            hd = lib().orc_spgramcf_create_default(nfft)
        super().__init__(hd)

    def reset(self):
        lib().orc_spgramcf_reset(self.h)

    def push(self, v):
        v = complex(v)
        lib().orc_spgramcf_push(self.h, c_cf(v.real, v.imag))

    def write(self, x):
        x = _cf(x)
        lib().orc_spgramcf_write(self.h, x, len(x))

    def execute(self):
        X = np.empty(self.nfft, np.complex64)
        lib().orc_spgramcf_execute(self.h, X)
        return X

    def execute_psd(self):
        out = np.empty(self.nfft, np.float32)
        lib().orc_spgramcf_execute_psd(self.h, out)
        return out

    def accumulate_psd(self, x, alpha):
        x = _cf(x)
        if lib().orc_spgramcf_accumulate_psd(self.h, x, alpha, len(x)):
            raise ValueError("spgramcf_accumulate_psd: alpha must be in [0,1]")

    def write_accumulation(self):
        out = np.empty(self.nfft, np.float32)
        lib().orc_spgramcf_write_accumulation(self.h, out)
        return out

    def estimate_psd(self, x):
        x = _cf(x)
        out = np.empty(self.nfft, np.float32)
        lib().orc_spgramcf_estimate_psd(self.h, x, len(x), out)
        return out

    def num_transforms(self):
        return lib().orc_spgramcf_num_transforms(self.h)


class FirpfbchCccf(_Obj):
    """firpfbch_cccf: complex taps h[M*p]"""
    _destroy = "orc_firpfbch_cccf_destroy"

    def __init__(self, type, M, p, h):
        self.type, self.M = type, M
        super().__init__(lib().orc_firpfbch_cccf_create(type, M, p, _cf(h)))

    def execute_frames(self, x):
        x = _cf(x).reshape(-1, self.M)
        y = np.empty_like(x)
        lib().orc_firpfbch_cccf_execute_frames(self.h, x, x.shape[0], y)
        return y


class Firpfbch2(_Obj):
    _destroy = "orc_firpfbch2_destroy"

    def __init__(self, type, M, m, h=None, As=None):
        self.type, self.M = type, M
        hd = (lib().orc_firpfbch2_create_kaiser(type, M, m, As) if h is None
              else lib().orc_firpfbch2_create(type, M, m, _f(h)))
        super().__init__(hd)

    def reset(self):
        lib().orc_firpfbch2_reset(self.h)

    def execute_frames(self, x):
        n_in = self.M // 2 if self.type == ANALYZER else self.M
        n_out = self.M if self.type == ANALYZER else self.M // 2
        x = _cf(x).reshape(-1, n_in)
        y = np.empty((x.shape[0], n_out), np.complex64)
        lib().orc_firpfbch2_execute_frames(self.h, x, x.shape[0], y)
        return y


def bench_firpfbch2_analysis(M, m, As, n_samples, n_threads):
    """wall seconds for n_threads independent oracle instances, n_samples each ("port" baseline)."""
    return lib().orc_bench_firpfbch2_analysis(M, m, As, n_samples, n_threads)


# ----------------------------------------------------------------------------------------
# the real reference (oracle/_ref), when it has been built
# ----------------------------------------------------------------------------------------
_ref = None
_drv = None


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libliquid_ref_sse.so"))


def _ref_variant():
    forced = os.environ.get("LIQUID_ORACLE_REF_VARIANT")      # "" (AVX2+FMA build) or "_sse" (baseline x86-64)
    if forced is not None:
        return forced
    flags = _cpu_flags()
    return "" if {"avx2", "fma", "bmi2"} <= flags else "_sse"


def ref():
    """ctypes handle on the compiled reference; raises if oracle/_ref was never built."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libliquid_ref%s.so" % _ref_variant())
        if not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
        R = C.CDLL(path)
        R.liquid_firdes_kaiser.restype = None
        R.liquid_firdes_kaiser.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_float, c_f_p]
        R.liquid_firdes_rrcos.restype = None
        R.liquid_firdes_rrcos.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float, c_f_p]
        for fn in ("liquid_firdes_arkaiser", "liquid_firdes_rkaiser"):
            getattr(R, fn).restype = None
            getattr(R, fn).argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float, c_f_p]
        R.estimate_req_filter_As.restype = C.c_float; R.estimate_req_filter_As.argtypes = [C.c_float, C.c_uint]
        R.rkaiser_approximate_rho.restype = C.c_float; R.rkaiser_approximate_rho.argtypes = [C.c_uint, C.c_float]
        R.liquid_besseli0f.restype = C.c_float; R.liquid_besseli0f.argtypes = [C.c_float]
        R.liquid_lngammaf.restype = C.c_float; R.liquid_lngammaf.argtypes = [C.c_float]
        R.kaiser_beta_As.restype = C.c_float; R.kaiser_beta_As.argtypes = [C.c_float]
        R.kaiser.restype = C.c_float; R.kaiser.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float]
        R.sincf.restype = C.c_float; R.sincf.argtypes = [C.c_float]
        R.liquid_msb_index.restype = C.c_uint; R.liquid_msb_index.argtypes = [C.c_uint]
        R.dotprod_crcf_run.argtypes = [c_f_p, c_cf_p, C.c_uint, c_cf_p]
        R.dotprod_crcf_create.restype = C.c_void_p; R.dotprod_crcf_create.argtypes = [c_f_p, C.c_uint]
        R.dotprod_crcf_execute.argtypes = [C.c_void_p, c_cf_p, c_cf_p]
        R.dotprod_crcf_destroy.argtypes = [C.c_void_p]
        R.fft_run.argtypes = [C.c_uint, c_cf_p, c_cf_p, C.c_int, C.c_int]
        R.firfilt_crcf_create.restype = C.c_void_p; R.firfilt_crcf_create.argtypes = [c_f_p, C.c_uint]
        R.firfilt_crcf_create_kaiser.restype = C.c_void_p
        R.firfilt_crcf_create_kaiser.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_float]
        R.firfilt_crcf_set_scale.argtypes = [C.c_void_p, C.c_float]
        R.firfilt_crcf_execute_block.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
        R.firfilt_crcf_destroy.argtypes = [C.c_void_p]
        R.firpfb_crcf_create.restype = C.c_void_p; R.firpfb_crcf_create.argtypes = [C.c_uint, c_f_p, C.c_uint]
        R.firpfb_crcf_push.argtypes = [C.c_void_p, c_cf]
        R.firpfb_crcf_execute.argtypes = [C.c_void_p, C.c_uint, c_cf_p]
        R.firpfb_crcf_destroy.argtypes = [C.c_void_p]
        R.firpfbch_crcf_create.restype = C.c_void_p
        R.firpfbch_crcf_create.argtypes = [C.c_int, C.c_uint, C.c_uint, c_f_p]
        R.firpfbch_crcf_create_kaiser.restype = C.c_void_p
        R.firpfbch_crcf_create_kaiser.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float]
        R.firpfbch_crcf_create_rnyquist.restype = C.c_void_p
        R.firpfbch_crcf_create_rnyquist.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float, C.c_int]
        R.firpfbch_crcf_analyzer_execute.argtypes = [C.c_void_p, c_cf_p, c_cf_p]
        R.firpfbch_crcf_synthesizer_execute.argtypes = [C.c_void_p, c_cf_p, c_cf_p]
        R.firpfbch_crcf_destroy.argtypes = [C.c_void_p]
        R.firpfbch2_crcf_create.restype = C.c_void_p
        R.firpfbch2_crcf_create.argtypes = [C.c_int, C.c_uint, C.c_uint, c_f_p]
        R.firpfbch2_crcf_create_kaiser.restype = C.c_void_p
        R.firpfbch2_crcf_create_kaiser.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float]
        R.firpfbch2_crcf_execute.argtypes = [C.c_void_p, c_cf_p, c_cf_p]
        R.firpfbch2_crcf_destroy.argtypes = [C.c_void_p]
        R.msequence_create_default.restype = C.c_void_p; R.msequence_create_default.argtypes = [C.c_uint]
        R.msequence_generate_symbol.restype = C.c_uint
        R.msequence_generate_symbol.argtypes = [C.c_void_p, C.c_uint]
        R.msequence_destroy.argtypes = [C.c_void_p]
        _ref = R
    return _ref


def ref_driver():
    global _drv
    if _drv is None:
        path = os.path.join(HERE, "_ref", "libref_driver%s.so" % _ref_variant())
        if not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref driver not built")
        D = C.CDLL(path)
        D.ref_bench.restype = C.c_double
        D.ref_bench.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float, C.c_ulonglong, C.c_uint]
        D.ref_pfbch2_analysis.restype = None
        D.ref_pfbch2_analysis.argtypes = [C.c_uint, C.c_uint, C.c_float, c_cf_p, C.c_uint, c_cf_p]
        _drv = D
    return _drv


def ref_firpfbch2_frames(type, M, m, As, x, h=None):
    R = ref()
    q = R.firpfbch2_crcf_create_kaiser(type, M, m, As) if h is None else R.firpfbch2_crcf_create(type, M, m, _f(h))
    q = C.c_void_p(q)
    n_in = M // 2 if type == ANALYZER else M
    n_out = M if type == ANALYZER else M // 2
    x = _cf(x).reshape(-1, n_in)
    y = np.empty((x.shape[0], n_out), np.complex64)
    for t in range(x.shape[0]):
        R.firpfbch2_crcf_execute(q, x[t], y[t])
    R.firpfbch2_crcf_destroy(q)
    return y


def ref_firpfbch_frames(type, M, x, p=None, h=None, kaiser=None, rnyquist=None):
    R = ref()
    if kaiser is not None:
        q = R.firpfbch_crcf_create_kaiser(type, M, *kaiser)
    elif rnyquist is not None:
        q = R.firpfbch_crcf_create_rnyquist(type, M, *rnyquist)
    else:
        q = R.firpfbch_crcf_create(type, M, p, _f(h))
    q = C.c_void_p(q)
    x = _cf(x).reshape(-1, M)
    y = np.empty_like(x)
    fn = R.firpfbch_crcf_analyzer_execute if type == ANALYZER else R.firpfbch_crcf_synthesizer_execute
    for t in range(x.shape[0]):
        fn(q, x[t], y[t])
    R.firpfbch_crcf_destroy(q)
    return y


def ref_firpfbch_cccf_frames(type, M, p, h, x):
    """the compiled reference's firpfbch_cccf (complex taps h[M*p]) over whole frames"""
    R = ref()
    R.firpfbch_cccf_create.restype = C.c_void_p
    R.firpfbch_cccf_create.argtypes = [C.c_int, C.c_uint, C.c_uint, c_cf_p]
    R.firpfbch_cccf_analyzer_execute.argtypes = [C.c_void_p, c_cf_p, c_cf_p]
    R.firpfbch_cccf_synthesizer_execute.argtypes = [C.c_void_p, c_cf_p, c_cf_p]
    R.firpfbch_cccf_destroy.argtypes = [C.c_void_p]
    q = C.c_void_p(R.firpfbch_cccf_create(type, M, p, _cf(h)))
    x = _cf(x).reshape(-1, M)
    y = np.empty_like(x)
    fn = R.firpfbch_cccf_analyzer_execute if type == ANALYZER else R.firpfbch_cccf_synthesizer_execute
    for t in range(x.shape[0]):
        fn(q, x[t], y[t])
    R.firpfbch_cccf_destroy(q)
    return y


def ref_firdecim_block(M, x, h=None, kaiser=None):
    R = ref()
    R.firdecim_crcf_create.restype = C.c_void_p; R.firdecim_crcf_create.argtypes = [C.c_uint, c_f_p, C.c_uint]
    R.firdecim_crcf_create_kaiser.restype = C.c_void_p; R.firdecim_crcf_create_kaiser.argtypes = [C.c_uint, C.c_uint, C.c_float]
    R.firdecim_crcf_execute_block.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
    R.firdecim_crcf_destroy.argtypes = [C.c_void_p]
    q = C.c_void_p(R.firdecim_crcf_create(M, _f(h), len(h)) if kaiser is None else R.firdecim_crcf_create_kaiser(M, *kaiser))
    x = _cf(x)
    y = np.empty(len(x) // M, np.complex64)
    R.firdecim_crcf_execute_block(q, x, len(y), y)
    R.firdecim_crcf_destroy(q)
    return y


def ref_firinterp_block(M, x, h=None, kaiser=None):
    R = ref()
    R.firinterp_crcf_create.restype = C.c_void_p; R.firinterp_crcf_create.argtypes = [C.c_uint, c_f_p, C.c_uint]
    R.firinterp_crcf_create_kaiser.restype = C.c_void_p; R.firinterp_crcf_create_kaiser.argtypes = [C.c_uint, C.c_uint, C.c_float]
    R.firinterp_crcf_execute_block.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_cf_p]
    R.firinterp_crcf_destroy.argtypes = [C.c_void_p]
    q = C.c_void_p(R.firinterp_crcf_create(M, _f(h), len(h)) if kaiser is None else R.firinterp_crcf_create_kaiser(M, *kaiser))
    x = _cf(x)
    y = np.empty(len(x) * M, np.complex64)
    R.firinterp_crcf_execute_block(q, x, len(x), y)
    R.firinterp_crcf_destroy(q)
    return y


def _c_by_value(v):
    """a C `float complex` argument passed by value: on x86-64 SysV it travels in the low 8 bytes of one XMM
    register, exactly like a double with the same bits (ctypes has no complex type before 3.14)"""
    return C.c_double(np.array([v], np.complex64).view(np.float64)[0])


def ref_resamp2(mode, m, f0, As, x):
    """the compiled reference's resamp2_crcf run sample by sample (it has no block form)"""
    R = ref()
    R.resamp2_crcf_create.restype = C.c_void_p; R.resamp2_crcf_create.argtypes = [C.c_uint, C.c_float, C.c_float]
    R.resamp2_crcf_destroy.argtypes = [C.c_void_p]
    q = C.c_void_p(R.resamp2_crcf_create(m, f0, As))
    x = _cf(x)
    if mode == "filter":
        R.resamp2_crcf_filter_execute.argtypes = [C.c_void_p, C.c_double, c_cf_p, c_cf_p]
        y0 = np.empty(len(x), np.complex64); y1 = np.empty(len(x), np.complex64)
        for i in range(len(x)):
            R.resamp2_crcf_filter_execute(q, _c_by_value(x[i]), y0[i:i + 1], y1[i:i + 1])
        out = (y0, y1)
    elif mode == "interp":
        R.resamp2_crcf_interp_execute.argtypes = [C.c_void_p, C.c_double, c_cf_p]
        out = np.empty(2 * len(x), np.complex64)
        for i in range(len(x)):
            R.resamp2_crcf_interp_execute(q, _c_by_value(x[i]), out[2 * i:2 * i + 2])
    else:
        fn = getattr(R, "resamp2_crcf_%s_execute" % mode)
        fn.argtypes = [C.c_void_p, c_cf_p, c_cf_p]
        n = len(x) // 2
        per = 1 if mode == "decim" else 2
        out = np.empty(n * per, np.complex64)
        for i in range(n):
            fn(q, x[2 * i:2 * i + 2], out[per * i:per * i + per])
    R.resamp2_crcf_destroy(q)
    return out


MIXED_MODES = ("decim", "filter", "interp", "analyzer", "synthesizer", "filter")


def mixed_segments(n):
    """(mode, first sample, samples) of the mixed-mode run over n input samples: segments of 36 + 2 k samples,
    the modes of MIXED_MODES in turn (the odd number of filter calls in some segments leaves the toggle set)"""
    segs, pos, k = [], 0, 0
    while pos < n:
        mode = MIXED_MODES[k % len(MIXED_MODES)]
        ln = min(36 + 2 * k + (1 if mode == "filter" else 0), n - pos)
        if mode not in ("filter", "interp"):
            ln -= ln % 2
        if ln == 0:
            break
        segs.append((mode, pos, ln))
        pos += ln
        k += 1
    return segs


def ref_resamp2_mixed(m, f0, As, x):
    """one reference object driven through all five modes in turn; outputs concatenated (filter: y0 then y1)"""
    R = ref()
    R.resamp2_crcf_create.restype = C.c_void_p; R.resamp2_crcf_create.argtypes = [C.c_uint, C.c_float, C.c_float]
    R.resamp2_crcf_destroy.argtypes = [C.c_void_p]
    R.resamp2_crcf_filter_execute.argtypes = [C.c_void_p, C.c_double, c_cf_p, c_cf_p]
    R.resamp2_crcf_interp_execute.argtypes = [C.c_void_p, C.c_double, c_cf_p]
    for nm in ("analyzer", "synthesizer", "decim"):
        getattr(R, "resamp2_crcf_%s_execute" % nm).argtypes = [C.c_void_p, c_cf_p, c_cf_p]
    q = C.c_void_p(R.resamp2_crcf_create(m, f0, As))
    x = _cf(x)
    parts = []
    for mode, pos, ln in mixed_segments(len(x)):
        seg = x[pos:pos + ln]
        if mode == "filter":
            y0 = np.empty(ln, np.complex64); y1 = np.empty(ln, np.complex64)
            for i in range(ln):
                R.resamp2_crcf_filter_execute(q, _c_by_value(seg[i]), y0[i:i + 1], y1[i:i + 1])
            parts += [y0, y1]
        elif mode == "interp":
            y = np.empty(2 * ln, np.complex64)
            for i in range(ln):
                R.resamp2_crcf_interp_execute(q, _c_by_value(seg[i]), y[2 * i:2 * i + 2])
            parts.append(y)
        else:
            per = 1 if mode == "decim" else 2
            y = np.empty(ln // 2 * per, np.complex64)
            fn = getattr(R, "resamp2_crcf_%s_execute" % mode)
            for i in range(ln // 2):
                fn(q, seg[2 * i:2 * i + 2], y[per * i:per * i + per])
            parts.append(y)
    R.resamp2_crcf_destroy(q)
    return np.concatenate(parts)


def run_resamp2_mixed(q, x):
    """the same mixed-mode run on an object with the block methods (the oracle's Resamp2 or the CUDA mirror)"""
    x = _cf(x)
    parts = []
    for mode, pos, ln in mixed_segments(len(x)):
        fn = getattr(q, mode + "_execute_block", None) or getattr(q, mode + "_execute")
        r = fn(x[pos:pos + ln])
        parts += list(r) if mode == "filter" else [r]
    return np.concatenate(parts)


def ref_msresamp2_block(type, S, fc, f0, As, x):
    R = ref()
    R.msresamp2_crcf_create.restype = C.c_void_p
    R.msresamp2_crcf_create.argtypes = [C.c_int, C.c_uint, C.c_float, C.c_float, C.c_float]
    R.msresamp2_crcf_execute.argtypes = [C.c_void_p, c_cf_p, c_cf_p]
    R.msresamp2_crcf_get_delay.restype = C.c_float; R.msresamp2_crcf_get_delay.argtypes = [C.c_void_p]
    R.msresamp2_crcf_destroy.argtypes = [C.c_void_p]
    q = C.c_void_p(R.msresamp2_crcf_create(type, S, fc, f0, As))
    x = _cf(x)
    M = 1 << S
    if type == RESAMP_INTERP:
        y = np.empty(len(x) * M, np.complex64)
        for i in range(len(x)):
            R.msresamp2_crcf_execute(q, x[i:i + 1], y[i * M:(i + 1) * M])
    else:
        y = np.empty(len(x) // M, np.complex64)
        for i in range(len(y)):
            R.msresamp2_crcf_execute(q, x[i * M:(i + 1) * M], y[i:i + 1])
    delay = R.msresamp2_crcf_get_delay(q)
    R.msresamp2_crcf_destroy(q)
    return y, delay


def ref_asgramcf(nfft, x, offset=None, scale=None):
    """the compiled reference's asgramcf after writing x: (ascii, peak dB, peak frequency)"""
    R = ref()
    R.asgramcf_create.restype = C.c_void_p; R.asgramcf_create.argtypes = [C.c_uint]
    R.asgramcf_set_scale.argtypes = [C.c_void_p, C.c_float, C.c_float]
    R.asgramcf_write.argtypes = [C.c_void_p, c_cf_p, C.c_uint]
    R.asgramcf_execute.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    R.asgramcf_destroy.argtypes = [C.c_void_p]
    q = C.c_void_p(R.asgramcf_create(nfft))
    if offset is not None:
        R.asgramcf_set_scale(q, offset, scale)
    x = _cf(x)
    R.asgramcf_write(q, x, len(x))
    buf = C.create_string_buffer(nfft + 1)
    pv, pf = C.c_float(), C.c_float()
    R.asgramcf_execute(q, buf, C.byref(pv), C.byref(pf))
    R.asgramcf_destroy(q)
    return buf.raw[:nfft].decode("ascii"), pv.value, pf.value


class RefOfdmframegen:
    """the compiled reference's ofdmframegen behind the same methods as Ofdmframegen"""

    def __init__(self, M, cp_len, taper_len, p=None):
        R = ref()
        R.ofdmframegen_create.restype = C.c_void_p
        R.ofdmframegen_create.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
        R.ofdmframegen_destroy.argtypes = [C.c_void_p]
        R.ofdmframegen_reset.argtypes = [C.c_void_p]
        for nm in ("write_S0a", "write_S0b", "write_S1", "writetail"):
            getattr(R, "ofdmframegen_" + nm).argtypes = [C.c_void_p, c_cf_p]
        R.ofdmframegen_writesymbol.argtypes = [C.c_void_p, c_cf_p, c_cf_p]
        self.R, self.M, self.cp, self.tl = R, M, cp_len, taper_len
        self._p = None if p is None else np.ascontiguousarray(p, np.uint8)
        self.q = C.c_void_p(R.ofdmframegen_create(M, cp_len, taper_len, None if self._p is None else self._p.ctypes.data))

    def __del__(self):
        if getattr(self, "q", None):
            self.R.ofdmframegen_destroy(self.q)
            self.q = None

    def reset(self):
        self.R.ofdmframegen_reset(self.q)

    def _sym(self, name):
        y = np.empty(self.M + self.cp, np.complex64)
        getattr(self.R, "ofdmframegen_" + name)(self.q, y)
        return y

    def write_S0a(self):
        return self._sym("write_S0a")

    def write_S0b(self):
        return self._sym("write_S0b")

    def write_S1(self):
        return self._sym("write_S1")

    def writesymbols(self, x):
        x = _cf(x)
        n = len(x) // self.M
        y = np.empty(n * (self.M + self.cp), np.complex64)
        for s in range(n):
            self.R.ofdmframegen_writesymbol(self.q, x[s * self.M:(s + 1) * self.M], y[s * (self.M + self.cp):(s + 1) * (self.M + self.cp)])
        return y

    def writetail(self):
        y = np.empty(max(self.tl, 1), np.complex64)
        self.R.ofdmframegen_writetail(self.q, y)
        return y[:self.tl]


def ref_fftfilt_blocks(h, n, x, scale=None):
    R = ref()
    R.fftfilt_crcf_create.restype = C.c_void_p; R.fftfilt_crcf_create.argtypes = [c_f_p, C.c_uint, C.c_uint]
    R.fftfilt_crcf_set_scale.argtypes = [C.c_void_p, C.c_float]
    R.fftfilt_crcf_execute.argtypes = [C.c_void_p, c_cf_p, c_cf_p]
    R.fftfilt_crcf_destroy.argtypes = [C.c_void_p]
    q = C.c_void_p(R.fftfilt_crcf_create(_f(h), len(h), n))
    if scale is not None:
        R.fftfilt_crcf_set_scale(q, scale)
    x = _cf(x)
    y = np.empty(len(x) // n * n, np.complex64)
    for b in range(len(x) // n):
        R.fftfilt_crcf_execute(q, x[b * n:(b + 1) * n], y[b * n:(b + 1) * n])
    R.fftfilt_crcf_destroy(q)
    return y


class RefSpgramcf:
    """the compiled reference's spgramcf behind the same methods as Spgramcf"""

    def __init__(self, nfft, window=None, window_len=None, beta=None):
        R = ref()
        R.spgramcf_create.restype = C.c_void_p; R.spgramcf_create.argtypes = [C.c_uint, c_f_p, C.c_uint]
        R.spgramcf_create_kaiser.restype = C.c_void_p; R.spgramcf_create_kaiser.argtypes = [C.c_uint, C.c_uint, C.c_float]
        R.spgramcf_create_default.restype = C.c_void_p; R.spgramcf_create_default.argtypes = [C.c_uint]
        for fn in ("destroy", "reset"):
            getattr(R, "spgramcf_" + fn).argtypes = [C.c_void_p]
        R.spgramcf_write.argtypes = [C.c_void_p, c_cf_p, C.c_uint]
        R.spgramcf_execute.argtypes = [C.c_void_p, c_cf_p]
        R.spgramcf_execute_psd.argtypes = [C.c_void_p, c_f_p]
        R.spgramcf_accumulate_psd.argtypes = [C.c_void_p, c_cf_p, C.c_float, C.c_uint]
        R.spgramcf_write_accumulation.argtypes = [C.c_void_p, c_f_p]
        R.spgramcf_estimate_psd.argtypes = [C.c_void_p, c_cf_p, C.c_uint, c_f_p]
        self.R, self.nfft = R, nfft
        if window is not None:
            self.q = C.c_void_p(R.spgramcf_create(nfft, _f(window), len(window)))
        elif window_len is not None:
            self.q = C.c_void_p(R.spgramcf_create_kaiser(nfft, window_len, beta))
        else:
            self.q = C.c_void_p(R.spgramcf_create_default(nfft))

    def __del__(self):
        if getattr(self, "q", None):
            self.R.spgramcf_destroy(self.q)
            self.q = None

    def reset(self):
        self.R.spgramcf_reset(self.q)

    def write(self, x):
        x = _cf(x)
        self.R.spgramcf_write(self.q, x, len(x))

    def execute(self):
        X = np.empty(self.nfft, np.complex64)
        self.R.spgramcf_execute(self.q, X)
        return X

    def execute_psd(self):
        out = np.empty(self.nfft, np.float32)
        self.R.spgramcf_execute_psd(self.q, out)
        return out

    def accumulate_psd(self, x, alpha):
        x = _cf(x)
        self.R.spgramcf_accumulate_psd(self.q, x, alpha, len(x))

    def write_accumulation(self):
        out = np.empty(self.nfft, np.float32)
        self.R.spgramcf_write_accumulation(self.q, out)
        return out

    def estimate_psd(self, x):
        x = _cf(x)
        out = np.empty(self.nfft, np.float32)
        self.R.spgramcf_estimate_psd(self.q, x, len(x), out)
        return out


def ref_firfilt_block(x, h=None, kaiser=None, scale=None):
    R = ref()
    q = C.c_void_p(R.firfilt_crcf_create(_f(h), len(h)) if kaiser is None else R.firfilt_crcf_create_kaiser(*kaiser))
    if scale is not None:
        R.firfilt_crcf_set_scale(q, scale)
    x = _cf(x)
    y = np.empty_like(x)
    R.firfilt_crcf_execute_block(q, x, x.size, y)
    R.firfilt_crcf_destroy(q)
    return y


def ref_fft(x, direction=FFT_FORWARD):
    x = _cf(x)
    y = np.empty_like(x)
    ref().fft_run(x.size, x, y, direction, 0)
    return y


def ref_bench(kind, M, m, As, n_samples, n_threads):
    """wall seconds: n_threads reference instances, n_samples each (kind 0 pfbch2, 1 pfbch, 2 firfilt)."""
    return ref_driver().ref_bench(kind, M, m, As, n_samples, n_threads)
