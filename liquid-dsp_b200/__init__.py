"""liquid_dsp_b200 -- Python host-side mirror of the liquid-dsp channelizer API over libliquid_cuda.so.

The product is the C-ABI shared library (``include/liquid_cuda.h``); this module only binds it
with ctypes and mirrors the reference's object interface (same names, argument meaning and error
behaviour: invalid arguments make the library print ``error: <fn>(), <msg>`` and exit(1), as
liquid-dsp does).  There is no CPU fallback: if the CUDA library is missing, importing fails; if no
B200 is present, the first ``create`` aborts.

Host arrays are numpy ``complex64`` / ``float32``; the ``*_dev`` methods take raw device pointers
(``tensor.data_ptr()``) and a CUDA stream handle, so torch is plumbing only and never imported here.
"""
import ctypes as C
import weakref
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libliquid_cuda.so")

LIQUID_ANALYZER, LIQUID_SYNTHESIZER = 0, 1
LIQUID_FFT_FORWARD, LIQUID_FFT_BACKWARD = +1, -1
LIQUID_FIRFILT_KAISER, LIQUID_FIRFILT_ARKAISER, LIQUID_FIRFILT_RKAISER = 1, 7, 8
LIQUID_FIRFILT_RRC, LIQUID_FIRFILT_hM3 = 9, 10

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "liquid_dsp_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C liquid-dsp_b200/csrc` (there is no CPU fallback)" % LIB_PATH)

lib = C.CDLL(LIB_PATH)

_cf_p = np.ctypeslib.ndpointer(dtype=np.complex64, flags="C_CONTIGUOUS")
_f_p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u_p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


class _cf(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


_DECLARED = []


def _sig(name, restype, *argtypes):
    _DECLARED.append(name)
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_vp, _u, _ull, _i, _fl = C.c_void_p, C.c_uint, C.c_ulonglong, C.c_int, C.c_float

# every symbol include/liquid_cuda.h declares, with its signature
_sig("kaiser_beta_As", _fl, _fl)
_sig("liquid_firdes_kaiser", None, _u, _fl, _fl, _fl, _f_p)
_sig("liquid_firdes_rrcos", None, _u, _u, _fl, _fl, _f_p)
_sig("estimate_req_filter_As", _fl, _fl, _u)
_sig("liquid_filter_isi", None, _f_p, _u, _u, C.POINTER(C.c_float), C.POINTER(C.c_float))
_sig("liquid_firdes_arkaiser", None, _u, _u, _fl, _fl, _f_p)
_sig("liquid_firdes_rkaiser", None, _u, _u, _fl, _fl, _f_p)
_sig("liquid_firdes_hM3", None, _u, _u, _fl, _fl, _f_p)
_sig("firdespm_create", _vp, _u, _u, _f_p, _f_p, _f_p, C.POINTER(C.c_int), _i)
_sig("firdespm_destroy", None, _vp)
_sig("firdespm_print", None, _vp)
_sig("firdespm_execute", None, _vp, _f_p)
for _nm in ("rcos", "fexp", "rfexp", "fsech", "rfsech", "farcsech", "rfarcsech", "gmsktx", "gmskrx"):
    _sig("liquid_firdes_" + _nm, None, _u, _u, _fl, _fl, _f_p)
_sig("firdespm_run", None, _u, _u, _f_p, _f_p, _f_p, C.POINTER(C.c_int), _i, _f_p)
_sig("liquid_firdes_prototype", None, _i, _u, _u, _fl, _fl, _f_p)
_sig("firpfbch2_crcf_create", _vp, _i, _u, _u, _f_p)
_sig("firpfbch2_crcf_create_kaiser", _vp, _i, _u, _u, _fl)
_sig("firpfbch2_crcf_destroy", None, _vp)
_sig("firpfbch2_crcf_reset", None, _vp)
_sig("firpfbch2_crcf_print", None, _vp)
_sig("firpfbch2_crcf_execute", None, _vp, _cf_p, _cf_p)
_sig("firpfbch2_crcf_execute_block", None, _vp, _vp, _u, _vp)
_sig("firpfbch2_crcf_execute_block_dev", None, _vp, _vp, _ull, _vp, _vp)
_sig("firpfbch2_crcf_execute_segment_dev", None, _vp, _vp, _vp, _ull, _ull, _vp, _vp)
_sig("firpfbch2_crcf_get_state_len", _u, _vp)
_sig("firpfbch2_crcf_get_num_channels", _u, _vp)
_sig("liquid_cuda_comm_unique_id_bytes", _i)
_sig("liquid_cuda_comm_get_unique_id", None, _vp)
_sig("liquid_cuda_comm_create", _vp, _i, _i, _vp)
_sig("liquid_cuda_comm_destroy", None, _vp)
_sig("liquid_cuda_comm_rank", _i, _vp)
_sig("liquid_cuda_comm_world", _i, _vp)
_sig("liquid_cuda_comm_exchange_halo_dev", None, _vp, _vp, _vp, _ull, _vp)
_sig("firpfbch2_crcf_execute_sharded_dev", None, _vp, _vp, _vp, _ull, _ull, _vp, _vp)
_sig("firpfbch2_crcf_execute_sharded", None, _vp, _vp, _vp, _ull, _u, _vp)
_sig("liquid_cuda_host_register", _i, _vp, _ull)
_sig("liquid_cuda_host_unregister", None, _vp)
_sig("firpfbch2_crcf_get_state", None, _vp, _cf_p, C.POINTER(C.c_int))
_sig("firpfbch2_crcf_set_state", None, _vp, _cf_p, _i)
_sig("firpfbch_crcf_create", _vp, _i, _u, _u, _f_p)
_sig("firpfbch_crcf_create_kaiser", _vp, _i, _u, _u, _fl)
_sig("firpfbch_crcf_create_rnyquist", _vp, _i, _u, _u, _fl, _i)
_sig("firpfbch_crcf_destroy", None, _vp)
_sig("firpfbch_crcf_reset", None, _vp)
_sig("firpfbch_crcf_print", None, _vp)
_sig("firpfbch_crcf_synthesizer_execute", None, _vp, _cf_p, _cf_p)
_sig("firpfbch_crcf_analyzer_execute", None, _vp, _cf_p, _cf_p)
_sig("firpfbch_crcf_analyzer_push", None, _vp, _cf)
_sig("firpfbch_crcf_analyzer_run", None, _vp, _u, _cf_p)
_sig("firpfbch_crcf_analyzer_execute_block", None, _vp, _vp, _u, _vp)
_sig("firpfbch_crcf_synthesizer_execute_block", None, _vp, _vp, _u, _vp)
_sig("firpfbch_crcf_execute_block_dev", None, _vp, _vp, _ull, _vp, _vp)
_sig("firpfbch_crcf_execute_streams_dev", None, _vp, _vp, _u, _ull, _vp, _vp)
_sig("firpfbch_crcf_get_state_len", _u, _vp)
_sig("firpfbch_crcf_get_state", None, _vp, _cf_p)
_sig("firpfbch_crcf_set_state", None, _vp, _cf_p)
_sig("firdecim_crcf_create", _vp, _u, _f_p, _u)
_sig("firdecim_crcf_create_kaiser", _vp, _u, _u, _fl)
_sig("firdecim_crcf_create_prototype", _vp, _i, _u, _u, _fl, _fl)
_sig("firdecim_crcf_destroy", None, _vp)
_sig("firdecim_crcf_print", None, _vp)
_sig("firdecim_crcf_clear", None, _vp)
_sig("firdecim_crcf_execute", None, _vp, _cf_p, _cf_p)
_sig("firdecim_crcf_execute_block", None, _vp, _vp, _u, _vp)
_sig("firdecim_crcf_execute_block_dev", None, _vp, _vp, _ull, _vp, _vp)
_sig("firinterp_crcf_create", _vp, _u, _f_p, _u)
_sig("firinterp_crcf_create_kaiser", _vp, _u, _u, _fl)
_sig("firinterp_crcf_create_prototype", _vp, _i, _u, _u, _fl, _fl)
_sig("firinterp_crcf_destroy", None, _vp)
_sig("firinterp_crcf_print", None, _vp)
_sig("firinterp_crcf_reset", None, _vp)
_sig("firinterp_crcf_execute", None, _vp, _cf, _cf_p)
_sig("estimate_req_filter_len", _u, _fl, _fl)
_sig("resamp2_crcf_create", _vp, _u, _fl, _fl)
_sig("resamp2_crcf_recreate", _vp, _vp, _u, _fl, _fl)
_sig("resamp2_crcf_destroy", None, _vp)
_sig("resamp2_crcf_print", None, _vp)
_sig("resamp2_crcf_clear", None, _vp)
_sig("resamp2_crcf_get_delay", _u, _vp)
_sig("resamp2_crcf_get_taps", None, _vp, _vp, _vp)
_sig("resamp2_crcf_filter_execute", None, _vp, _cf, _cf_p, _cf_p)
_sig("resamp2_crcf_analyzer_execute", None, _vp, _cf_p, _cf_p)
_sig("resamp2_crcf_synthesizer_execute", None, _vp, _cf_p, _cf_p)
_sig("resamp2_crcf_decim_execute", None, _vp, _cf_p, _cf_p)
_sig("resamp2_crcf_interp_execute", None, _vp, _cf, _cf_p)
_sig("resamp2_crcf_filter_execute_block", None, _vp, _vp, _u, _vp, _vp)
_sig("resamp2_crcf_filter_execute_block_dev", None, _vp, _vp, _ull, _vp, _vp, _vp)
for _nm in ("analyzer", "synthesizer", "decim", "interp"):
    _sig("resamp2_crcf_%s_execute_block" % _nm, None, _vp, _vp, _u, _vp)
    _sig("resamp2_crcf_%s_execute_block_dev" % _nm, None, _vp, _vp, _ull, _vp, _vp)
_sig("asgramcf_create", _vp, _u)
_sig("asgramcf_destroy", None, _vp)
_sig("asgramcf_reset", None, _vp)
_sig("asgramcf_set_scale", None, _vp, _fl, _fl)
_sig("asgramcf_push", None, _vp, _cf)
_sig("asgramcf_write", None, _vp, _cf_p, _u)
_sig("asgramcf_execute", None, _vp, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_float))
_sig("asgramcf_print", None, _vp)
_u8_p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_sig("ofdmframe_init_default_sctype", None, _u, _u8_p)
_sig("ofdmframe_validate_sctype", None, _u8_p, _u, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint))
_sig("ofdmframe_print_sctype", None, _u8_p, _u)
_sig("ofdmframegen_create", _vp, _u, _u, _u, _vp)
_sig("ofdmframegen_destroy", None, _vp)
_sig("ofdmframegen_print", None, _vp)
_sig("ofdmframegen_reset", None, _vp)
_sig("ofdmframegen_get_sctype", None, _vp, _u8_p)
_sig("ofdmframegen_write_S0a", None, _vp, _cf_p)
_sig("ofdmframegen_write_S0b", None, _vp, _cf_p)
_sig("ofdmframegen_write_S1", None, _vp, _cf_p)
_sig("ofdmframegen_writesymbol", None, _vp, _cf_p, _cf_p)
_sig("ofdmframegen_writesymbols", None, _vp, _vp, _u, _vp)
_sig("ofdmframegen_writesymbols_dev", None, _vp, _vp, _ull, _vp, _vp)
_sig("ofdmframegen_writetail", None, _vp, _cf_p)
OFDMFRAMESYNC_CALLBACK = C.CFUNCTYPE(C.c_int, C.POINTER(_cf), C.POINTER(C.c_ubyte), C.c_uint, C.c_void_p)
_sig("ofdmframesync_create", _vp, _u, _u, _u, _vp, OFDMFRAMESYNC_CALLBACK, _vp)
_sig("ofdmframesync_destroy", None, _vp)
_sig("ofdmframesync_print", None, _vp)
_sig("ofdmframesync_reset", None, _vp)
_sig("ofdmframesync_execute", None, _vp, _vp, _u)
_sig("ofdmframesync_get_rssi", _fl, _vp)
_sig("ofdmframesync_get_cfo", _fl, _vp)
_sig("ofdmframesync_debug_enable", None, _vp)
_sig("ofdmframesync_debug_disable", None, _vp)
_sig("ofdmframesync_debug_print", None, _vp, C.c_char_p)
_sig("ofdmframesync_is_frame_open", _i, _vp)
_sig("fftfilt_crcf_create", _vp, _f_p, _u, _u)
_sig("fftfilt_crcf_destroy", None, _vp)
_sig("fftfilt_crcf_reset", None, _vp)
_sig("fftfilt_crcf_print", None, _vp)
_sig("fftfilt_crcf_set_scale", None, _vp, _fl)
_sig("fftfilt_crcf_get_length", _u, _vp)
_sig("fftfilt_crcf_execute", None, _vp, _cf_p, _cf_p)
_sig("fftfilt_crcf_execute_block", None, _vp, _vp, _u, _vp)
_sig("fftfilt_crcf_execute_block_dev", None, _vp, _vp, _ull, _vp, _vp)
_sig("msresamp2_crcf_create", _vp, _i, _u, _fl, _fl, _fl)
_sig("msresamp2_crcf_destroy", None, _vp)
_sig("msresamp2_crcf_print", None, _vp)
_sig("msresamp2_crcf_reset", None, _vp)
_sig("msresamp2_crcf_get_delay", _fl, _vp)
_sig("msresamp2_crcf_get_stage_semilength", _u, _vp, _u)
_sig("msresamp2_crcf_execute", None, _vp, _cf_p, _cf_p)
_sig("msresamp2_crcf_execute_block", None, _vp, _vp, _u, _vp)
_sig("msresamp2_crcf_execute_block_dev", None, _vp, _vp, _ull, _vp, _vp)
_sig("firinterp_crcf_execute_block", None, _vp, _vp, _u, _vp)
_sig("firinterp_crcf_execute_block_dev", None, _vp, _vp, _ull, _vp, _vp)
_sig("spgramcf_create", _vp, _u, _f_p, _u)
_sig("spgramcf_create_kaiser", _vp, _u, _u, _fl)
_sig("spgramcf_create_default", _vp, _u)
_sig("spgramcf_destroy", None, _vp)
_sig("spgramcf_reset", None, _vp)
_sig("spgramcf_push", None, _vp, _cf)
_sig("spgramcf_write", None, _vp, _cf_p, _u)
_sig("spgramcf_execute", None, _vp, _cf_p)
_sig("spgramcf_execute_psd", None, _vp, _f_p)
_sig("spgramcf_accumulate_psd", None, _vp, _cf_p, _fl, _u)
_sig("spgramcf_write_accumulation", None, _vp, _f_p)
_sig("spgramcf_estimate_psd", None, _vp, _cf_p, _u, _f_p)
_sig("spgramcf_accumulate_psd_dev", None, _vp, _vp, _fl, _ull, _vp)
_sig("spgramcf_estimate_psd_dev", None, _vp, _vp, _ull, _f_p, _vp)
_sig("spgramcf_get_num_transforms", _u, _vp)
_sig("firpfbch_cccf_create", _vp, _i, _u, _u, _cf_p)
_sig("firpfbch_cccf_create_kaiser", _vp, _i, _u, _u, _fl)
_sig("firpfbch_cccf_create_rnyquist", _vp, _i, _u, _u, _fl, _i)
_sig("firpfbch_cccf_destroy", None, _vp)
_sig("firpfbch_cccf_reset", None, _vp)
_sig("firpfbch_cccf_print", None, _vp)
_sig("firpfbch_cccf_synthesizer_execute", None, _vp, _cf_p, _cf_p)
_sig("firpfbch_cccf_analyzer_execute", None, _vp, _cf_p, _cf_p)
_sig("firpfbch_cccf_analyzer_push", None, _vp, _cf)
_sig("firpfbch_cccf_analyzer_run", None, _vp, _u, _cf_p)
_sig("firpfbch_cccf_analyzer_execute_block", None, _vp, _vp, _u, _vp)
_sig("firpfbch_cccf_synthesizer_execute_block", None, _vp, _vp, _u, _vp)
_sig("firpfbch_cccf_execute_block_dev", None, _vp, _vp, _ull, _vp, _vp)
_sig("firfilt_crcf_create", _vp, _f_p, _u)
_sig("firfilt_crcf_create_kaiser", _vp, _u, _fl, _fl, _fl)
_sig("firfilt_crcf_create_rnyquist", _vp, _i, _u, _u, _fl, _fl)
_sig("firfilt_crcf_create_rect", _vp, _u)
_sig("firfilt_crcf_recreate", _vp, _vp, _f_p, _u)
_sig("firfilt_crcf_destroy", None, _vp)
_sig("firfilt_crcf_reset", None, _vp)
_sig("firfilt_crcf_print", None, _vp)
_sig("firfilt_crcf_set_scale", None, _vp, _fl)
_sig("firfilt_crcf_push", None, _vp, _cf)
_sig("firfilt_crcf_execute", None, _vp, _cf_p)
_sig("firfilt_crcf_execute_block", None, _vp, _vp, _u, _vp)
_sig("firfilt_crcf_execute_block_dev", None, _vp, _vp, _ull, _vp, _vp)
_sig("firfilt_crcf_execute_streams_dev", None, _vp, _vp, _u, _ull, _vp, _vp)
_sig("firfilt_crcf_get_length", _u, _vp)
_sig("firfilt_crcf_freqresponse", None, _vp, _fl, _cf_p)
_sig("firfilt_crcf_groupdelay", _fl, _vp, _fl)
_sig("firpfb_crcf_create", _vp, _u, _f_p, _u)
_sig("firpfb_crcf_create_kaiser", _vp, _u, _u, _fl, _fl)
_sig("firpfb_crcf_create_rnyquist", _vp, _i, _u, _u, _u, _fl)
_sig("firpfb_crcf_create_drnyquist", _vp, _i, _u, _u, _u, _fl)
_sig("firpfb_crcf_recreate", _vp, _vp, _u, _f_p, _u)
_sig("firpfb_crcf_destroy", None, _vp)
_sig("firpfb_crcf_print", None, _vp)
_sig("firpfb_crcf_set_scale", None, _vp, _fl)
_sig("firpfb_crcf_reset", None, _vp)
_sig("firpfb_crcf_push", None, _vp, _cf)
_sig("firpfb_crcf_execute", None, _vp, _u, _cf_p)
_sig("firpfb_crcf_execute_block", None, _vp, _cf_p, _u_p, _u, _cf_p)
_sig("firpfb_crcf_execute_block_dev", None, _vp, _vp, _vp, _ull, _vp, _vp)
_sig("dotprod_crcf_run", None, _f_p, _cf_p, _u, _cf_p)
_sig("dotprod_crcf_run4", None, _f_p, _cf_p, _u, _cf_p)
_sig("dotprod_crcf_create", _vp, _f_p, _u)
_sig("dotprod_crcf_recreate", _vp, _vp, _f_p, _u)
_sig("dotprod_crcf_destroy", None, _vp)
_sig("dotprod_crcf_print", None, _vp)
_sig("dotprod_crcf_execute", None, _vp, _cf_p, _cf_p)
_sig("dotprod_crcf_execute_batch_dev", None, _vp, _vp, _ull, _ull, _vp, _vp)
_sig("fft_create_plan", _vp, _u, _vp, _vp, _i, _i)
_sig("fft_destroy_plan", None, _vp)
_sig("fft_print_plan", None, _vp)
_sig("fft_execute", None, _vp)
_sig("fft_run", None, _u, _cf_p, _cf_p, _i, _i)
_sig("fft_shift", None, _cf_p, _u)
_sig("fft_execute_batch_dev", None, _vp, _vp, _vp, _ull, _vp)
_sig("fft_execute_batch", None, _vp, _cf_p, _cf_p, _u)
_sig("liquid_cuda_version", C.c_char_p)
_sig("liquid_cuda_device_count", _i)
_sig("liquid_cuda_set_device", None, _i)
_sig("liquid_cuda_malloc", _vp, _ull)
_sig("liquid_cuda_free", None, _vp)
_sig("liquid_cuda_host_alloc", _vp, _ull)
_sig("liquid_cuda_host_free", None, _vp)
_sig("liquid_cuda_memcpy_h2d", None, _vp, _vp, _ull)
_sig("liquid_cuda_memcpy_d2h", None, _vp, _vp, _ull)
_sig("liquid_cuda_sync", None)
_sig("liquid_cuda_launch_count", _ull)
_sig("liquid_cuda_last_chunk_frames", _ull)
_sig("liquid_cuda_set_reserved_sms", None, _i)
_sig("liquid_cuda_get_reserved_sms", _i)
_sig("liquid_cuda_synth_input_dev", None, _vp, _ull, _ull, _ull, _vp)

#: names include/liquid_cuda.h declares (checked against the header and the .so by the CPU tests)
DECLARED_SYMBOLS = sorted(_DECLARED)


def version():
    return lib.liquid_cuda_version().decode()


def set_reserved_sms(n):
    """SMs the persistent kernels leave free for a collective / side-stream launch (default 0)"""
    lib.liquid_cuda_set_reserved_sms(n)


def device_count():
    return lib.liquid_cuda_device_count()


def set_device(dev):
    lib.liquid_cuda_set_device(int(dev))


def launch_count():
    return int(lib.liquid_cuda_launch_count())


class Comm:
    """liquid_cuda_comm: the C multi-GPU driver's communicator (NCCL underneath).  `id_bytes` comes from
    Comm.unique_id() on rank 0 and travels to the other ranks by whatever the launcher offers."""

    def __init__(self, rank, world, id_bytes=None):
        buf = C.create_string_buffer(bytes(id_bytes), len(id_bytes)) if id_bytes is not None else None
        self._h = lib.liquid_cuda_comm_create(rank, world, C.cast(buf, C.c_void_p) if buf is not None else None)
        self.rank, self.world = rank, world

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(lib.liquid_cuda_comm_unique_id_bytes())
        lib.liquid_cuda_comm_get_unique_id(C.cast(buf, C.c_void_p))
        return buf.raw

    def exchange_halo_dev(self, d_head, d_tail, n_samples, stream=0):
        lib.liquid_cuda_comm_exchange_halo_dev(self._h, d_head, d_tail, n_samples, stream)

    def destroy(self):
        if self._h:
            lib.liquid_cuda_comm_destroy(self._h)
            self._h = None


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def liquid_firdes_kaiser(n, fc, As, mu=0.0):
    h = np.empty(n, np.float32)
    lib.liquid_firdes_kaiser(n, fc, As, mu, h)
    return h


def liquid_firdes_rrcos(k, m, beta, dt=0.0):
    h = np.empty(2 * k * m + 1, np.float32)
    lib.liquid_firdes_rrcos(k, m, beta, dt, h)
    return h


def liquid_firdes_arkaiser(k, m, beta, dt=0.0):
    h = np.empty(2 * k * m + 1, np.float32)
    lib.liquid_firdes_arkaiser(k, m, beta, dt, h)
    return h


def liquid_firdes_rkaiser(k, m, beta, dt=0.0):
    h = np.empty(2 * k * m + 1, np.float32)
    lib.liquid_firdes_rkaiser(k, m, beta, dt, h)
    return h


def liquid_firdes_hM3(k, m, beta, dt=0.0):
    h = np.empty(2 * k * m + 1, np.float32)
    lib.liquid_firdes_hM3(k, m, beta, dt, h)
    return h


LIQUID_FIRDESPM_BANDPASS, LIQUID_FIRDESPM_DIFFERENTIATOR, LIQUID_FIRDESPM_HILBERT = 0, 1, 2
LIQUID_FIRDESPM_FLATWEIGHT, LIQUID_FIRDESPM_EXPWEIGHT, LIQUID_FIRDESPM_LINWEIGHT = 0, 1, 2
LIQUID_FIRFILT_PM = 2
LIQUID_FIRFILT_RCOS, LIQUID_FIRFILT_FEXP, LIQUID_FIRFILT_FSECH, LIQUID_FIRFILT_FARCSECH = 3, 4, 5, 6
LIQUID_FIRFILT_GMSKTX, LIQUID_FIRFILT_GMSKRX, LIQUID_FIRFILT_RFEXP, LIQUID_FIRFILT_RFSECH, LIQUID_FIRFILT_RFARCSECH = 11, 12, 13, 14, 15


def firdespm_run(h_len, bands, des, weights=None, wtype=None, btype=LIQUID_FIRDESPM_BANDPASS):
    """Parks-McClellan design (liquid.h:1504): bands = [f0, f1] per band in [0, 0.5], des / weights / wtype per band"""
    bands, des = _f32(bands), _f32(des)
    nb = des.size
    w = np.ones(nb, np.float32) if weights is None else _f32(weights)
    wt = np.zeros(nb, np.int32) if wtype is None else np.ascontiguousarray(wtype, np.int32)
    h = np.empty(h_len, np.float32)
    lib.firdespm_run(h_len, nb, bands, des, w, wt.ctypes.data_as(C.POINTER(C.c_int)), btype, h)
    return h


def liquid_firdes_prototype(ftype, k, m, beta, dt=0.0):
    h = np.empty(2 * k * m + 1, np.float32)
    lib.liquid_firdes_prototype(ftype, k, m, beta, dt, h)
    return h


def liquid_filter_isi(h, k, m):
    rms, peak = C.c_float(0.0), C.c_float(0.0)
    lib.liquid_filter_isi(np.ascontiguousarray(h, np.float32), k, m, C.byref(rms), C.byref(peak))
    return rms.value, peak.value


def kaiser_beta_As(As):
    return lib.kaiser_beta_As(As)


class _PinnedBlock:
    """owner of one cudaHostAlloc'd block: freed by pinned_free or, failing that, when the last array over it goes"""

    def __init__(self, nbytes):
        self.ptr = lib.liquid_cuda_host_alloc(nbytes)
        self.buf = (C.c_char * nbytes).from_address(self.ptr)
        self._fin = weakref.finalize(self, lib.liquid_cuda_host_free, self.ptr)

    def free(self):
        self._fin()                      # idempotent: runs liquid_cuda_host_free at most once


def pinned_empty(n, dtype=np.complex64):
    """numpy array over cudaHostAlloc'd (pinned) memory.  The block lives as long as the array (or any view of it)
    does: the array's base chain holds the owner object, whose finalizer frees the block; pinned_free(arr) frees it
    at once -- the array must not be used afterwards."""
    nbytes = int(n) * np.dtype(dtype).itemsize
    blk = _PinnedBlock(max(nbytes, 1))
    arr = np.frombuffer(blk.buf, dtype=dtype, count=int(n))      # arr.base -> blk.buf (a ctypes array)
    blk.buf._owner = blk                                           # ... which keeps the owner (and the block) alive
    _PINNED[arr.ctypes.data] = weakref.ref(blk)
    return arr


_PINNED = {}


def pinned_free(arr):
    ref = _PINNED.pop(arr.ctypes.data, None)
    blk = ref() if ref is not None else None
    if blk is not None:
        blk.free()


class _Handle:
    _destroy = None

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    def destroy(self):
        if getattr(self, "_h", None):
            getattr(lib, self._destroy)(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class firpfbch2_crcf(_Handle):
    """liquid.h:5754-5793.  analyzer: M/2 samples in, M out per frame; synthesizer: M in, M/2 out."""
    _destroy = "firpfbch2_crcf_destroy"

    def __init__(self, type, M, m, h):
        self.type, self.M, self.m = type, M, m
        super().__init__(lib.firpfbch2_crcf_create(type, M, m, _f32(h)))

    @classmethod
    def create(cls, type, M, m, h):
        return cls(type, M, m, h)

    @classmethod
    def create_kaiser(cls, type, M, m, As):
        q = cls.__new__(cls)
        q.type, q.M, q.m = type, M, m
        _Handle.__init__(q, lib.firpfbch2_crcf_create_kaiser(type, M, m, As))
        return q

    @property
    def n_in(self):
        return self.M // 2 if self.type == LIQUID_ANALYZER else self.M

    @property
    def n_out(self):
        return self.M if self.type == LIQUID_ANALYZER else self.M // 2

    def reset(self):
        lib.firpfbch2_crcf_reset(self._h)

    def print(self):
        lib.firpfbch2_crcf_print(self._h)

    def execute(self, x):
        x = _c64(x)
        assert x.size == self.n_in
        y = np.empty(self.n_out, np.complex64)
        lib.firpfbch2_crcf_execute(self._h, x, y)
        return y

    def execute_block(self, x, out=None):
        x = _c64(x)
        n_frames = x.size // self.n_in
        assert n_frames * self.n_in == x.size
        y = out if out is not None else np.empty(n_frames * self.n_out, np.complex64)
        lib.firpfbch2_crcf_execute_block(self._h, _ptr(x), n_frames, _ptr(y))
        return y.reshape(n_frames, self.n_out) if out is None else y

    def execute_block_dev(self, d_x, n_frames, d_y, stream=0):
        lib.firpfbch2_crcf_execute_block_dev(self._h, d_x, n_frames, d_y, stream)

    def execute_segment_dev(self, d_halo, d_x, first_frame, n_frames, d_y, stream=0):
        lib.firpfbch2_crcf_execute_segment_dev(self._h, d_halo, d_x, first_frame, n_frames, d_y, stream)

    def execute_sharded_dev(self, comm, d_x_with_headroom, first_frame, n_frames, d_y, stream=0):
        """one pass of the time-sharded analyzer through the C multi-GPU driver (csrc/multi.cu)"""
        lib.firpfbch2_crcf_execute_sharded_dev(self._h, comm._h, d_x_with_headroom, first_frame, n_frames, d_y, stream)

    def execute_sharded(self, comm, x, first_frame, out=None):
        """this rank's segment of a time-sharded stream, host arrays (csrc/firpfbch2.cu)"""
        x = _c64(x)
        n_frames = x.size // self.n_in
        y = out if out is not None else np.empty(n_frames * self.n_out, np.complex64)
        lib.firpfbch2_crcf_execute_sharded(self._h, comm._h, _ptr(x), first_frame, n_frames, _ptr(y))
        return y.reshape(n_frames, self.n_out) if out is None else y

    def get_state_len(self):
        return lib.firpfbch2_crcf_get_state_len(self._h)

    def get_state(self):
        hist = np.empty(self.get_state_len(), np.complex64)
        flag = C.c_int(0)
        lib.firpfbch2_crcf_get_state(self._h, hist, C.byref(flag))
        return hist, flag.value

    def set_state(self, hist, flag):
        lib.firpfbch2_crcf_set_state(self._h, _c64(hist), int(flag))


class firpfbch_crcf(_Handle):
    """liquid.h:5667-5728 (+ analyzer_push / analyzer_run, firpfbch.c:369, :383)."""
    _destroy = "firpfbch_crcf_destroy"

    def __init__(self, type, M, p, h):
        self.type, self.M = type, M
        super().__init__(lib.firpfbch_crcf_create(type, M, p, _f32(h)))

    @classmethod
    def create(cls, type, M, p, h):
        return cls(type, M, p, h)

    @classmethod
    def create_kaiser(cls, type, M, m, As):
        q = cls.__new__(cls)
        q.type, q.M = type, M
        _Handle.__init__(q, lib.firpfbch_crcf_create_kaiser(type, M, m, As))
        return q

    @classmethod
    def create_rnyquist(cls, type, M, m, beta, ftype=LIQUID_FIRFILT_RRC):
        q = cls.__new__(cls)
        q.type, q.M = type, M
        _Handle.__init__(q, lib.firpfbch_crcf_create_rnyquist(type, M, m, beta, ftype))
        return q

    def reset(self):
        lib.firpfbch_crcf_reset(self._h)

    def print(self):
        lib.firpfbch_crcf_print(self._h)

    def analyzer_execute(self, x):
        y = np.empty(self.M, np.complex64)
        lib.firpfbch_crcf_analyzer_execute(self._h, _c64(x), y)
        return y

    def synthesizer_execute(self, x):
        y = np.empty(self.M, np.complex64)
        lib.firpfbch_crcf_synthesizer_execute(self._h, _c64(x), y)
        return y

    def analyzer_push(self, v):
        v = complex(v)
        lib.firpfbch_crcf_analyzer_push(self._h, _cf(v.real, v.imag))

    def analyzer_run(self, k):
        y = np.empty(self.M, np.complex64)
        lib.firpfbch_crcf_analyzer_run(self._h, k, y)
        return y

    def execute_block(self, x, out=None):
        x = _c64(x)
        n_frames = x.size // self.M
        assert n_frames * self.M == x.size
        y = out if out is not None else np.empty(x.size, np.complex64)
        fn = (lib.firpfbch_crcf_analyzer_execute_block if self.type == LIQUID_ANALYZER
              else lib.firpfbch_crcf_synthesizer_execute_block)
        fn(self._h, _ptr(x), n_frames, _ptr(y))
        return y.reshape(n_frames, self.M) if out is None else y

    def execute_block_dev(self, d_x, n_frames, d_y, stream=0):
        lib.firpfbch_crcf_execute_block_dev(self._h, d_x, n_frames, d_y, stream)

    def execute_streams_dev(self, d_x, n_streams, n_frames, d_y, stream=0):
        lib.firpfbch_crcf_execute_streams_dev(self._h, d_x, n_streams, n_frames, d_y, stream)

    def get_state(self):
        hist = np.empty(lib.firpfbch_crcf_get_state_len(self._h), np.complex64)
        lib.firpfbch_crcf_get_state(self._h, hist)
        return hist

    def set_state(self, hist):
        lib.firpfbch_crcf_set_state(self._h, _c64(hist))



class firdecim_crcf(_Handle):
    """liquid.h:2665-2722 (src/filter/src/firdecim.c): n outputs from n*M inputs."""
    _destroy = "firdecim_crcf_destroy"

    def __init__(self, M, h):
        self.M = M
        hh = _f32(h)
        super().__init__(lib.firdecim_crcf_create(M, hh, len(hh)))

    @classmethod
    def create(cls, M, h):
        return cls(M, h)

    @classmethod
    def create_kaiser(cls, M, m, As):
        q = cls.__new__(cls)
        q.M = M
        _Handle.__init__(q, lib.firdecim_crcf_create_kaiser(M, m, As))
        return q

    @classmethod
    def create_prototype(cls, ftype, M, m, beta, dt=0.0):
        q = cls.__new__(cls)
        q.M = M
        _Handle.__init__(q, lib.firdecim_crcf_create_prototype(ftype, M, m, beta, dt))
        return q

    def clear(self):
        lib.firdecim_crcf_clear(self._h)

    def execute(self, x):
        y = np.empty(1, np.complex64)
        lib.firdecim_crcf_execute(self._h, _c64(x), y)
        return y[0]

    def execute_block(self, x):
        x = _c64(x)
        n = x.size // self.M
        y = np.empty(n, np.complex64)
        lib.firdecim_crcf_execute_block(self._h, _ptr(x), n, _ptr(y))
        return y

    def execute_block_dev(self, d_x, n, d_y, stream=0):
        lib.firdecim_crcf_execute_block_dev(self._h, d_x, n, d_y, stream)


LIQUID_RESAMP_INTERP, LIQUID_RESAMP_DECIM = 0, 1


class resamp2_crcf(_Handle):
    """liquid.h:2840-2925 (src/filter/src/resamp2.c): half-band filter, analysis / synthesis pair, 2:1 decimator,
    1:2 interpolator.  The *_block methods run n units in one kernel."""
    _destroy = "resamp2_crcf_destroy"

    def __init__(self, m, f0=0.0, As=60.0):
        self.m = m
        super().__init__(lib.resamp2_crcf_create(m, f0, As))

    @classmethod
    def create(cls, m, f0, As):
        return cls(m, f0, As)

    def recreate(self, m, f0, As):
        self._h = lib.resamp2_crcf_recreate(self._h, m, f0, As)
        self.m = m
        return self

    def clear(self):
        lib.resamp2_crcf_clear(self._h)

    def get_delay(self):
        return lib.resamp2_crcf_get_delay(self._h)

    def get_taps(self):
        h = np.empty(4 * self.m + 1, np.float32)
        h1 = np.empty(2 * self.m, np.float32)
        lib.resamp2_crcf_get_taps(self._h, _ptr(h), _ptr(h1))
        return h, h1

    def filter_execute(self, v):
        v = complex(v)
        y0 = np.empty(1, np.complex64); y1 = np.empty(1, np.complex64)
        lib.resamp2_crcf_filter_execute(self._h, _cf(v.real, v.imag), y0, y1)
        return y0[0], y1[0]

    def analyzer_execute(self, x):
        y = np.empty(2, np.complex64)
        lib.resamp2_crcf_analyzer_execute(self._h, _c64(x), y)
        return y

    def synthesizer_execute(self, x):
        y = np.empty(2, np.complex64)
        lib.resamp2_crcf_synthesizer_execute(self._h, _c64(x), y)
        return y

    def decim_execute(self, x):
        y = np.empty(1, np.complex64)
        lib.resamp2_crcf_decim_execute(self._h, _c64(x), y)
        return y[0]

    def interp_execute(self, v):
        v = complex(v)
        y = np.empty(2, np.complex64)
        lib.resamp2_crcf_interp_execute(self._h, _cf(v.real, v.imag), y)
        return y

    def filter_execute_block(self, x):
        x = _c64(x)
        y0 = np.empty(x.size, np.complex64); y1 = np.empty(x.size, np.complex64)
        lib.resamp2_crcf_filter_execute_block(self._h, _ptr(x), x.size, _ptr(y0), _ptr(y1))
        return y0, y1

    def _block(self, name, x, n, n_out):
        y = np.empty(n_out, np.complex64)
        getattr(lib, "resamp2_crcf_%s_execute_block" % name)(self._h, _ptr(x), n, _ptr(y))
        return y

    def analyzer_execute_block(self, x):
        x = _c64(x)
        return self._block("analyzer", x, x.size // 2, x.size // 2 * 2)

    def synthesizer_execute_block(self, x):
        x = _c64(x)
        return self._block("synthesizer", x, x.size // 2, x.size // 2 * 2)

    def decim_execute_block(self, x):
        x = _c64(x)
        return self._block("decim", x, x.size // 2, x.size // 2)

    def interp_execute_block(self, x):
        x = _c64(x)
        return self._block("interp", x, x.size, 2 * x.size)

    def execute_block_dev(self, mode, d_x, n, d_y, stream=0, d_y1=None):
        if mode == "filter":
            lib.resamp2_crcf_filter_execute_block_dev(self._h, d_x, n, d_y, d_y1, stream)
        else:
            getattr(lib, "resamp2_crcf_%s_execute_block_dev" % mode)(self._h, d_x, n, d_y, stream)


class msresamp2_crcf(_Handle):
    """liquid.h:3027-3080 (src/filter/src/msresamp2.c): cascade of half-band stages, rate 2^num_stages."""
    _destroy = "msresamp2_crcf_destroy"

    def __init__(self, type, num_stages, fc, f0, As):
        self.type, self.num_stages, self.M = type, num_stages, 1 << num_stages
        super().__init__(lib.msresamp2_crcf_create(type, num_stages, fc, f0, As))

    @classmethod
    def create(cls, type, num_stages, fc, f0, As):
        return cls(type, num_stages, fc, f0, As)

    def reset(self):
        lib.msresamp2_crcf_reset(self._h)

    def get_delay(self):
        return lib.msresamp2_crcf_get_delay(self._h)

    def stage_semilengths(self):
        return [lib.msresamp2_crcf_get_stage_semilength(self._h, i) for i in range(self.num_stages)]

    def execute(self, x):
        y = np.empty(self.M if self.type == LIQUID_RESAMP_INTERP else 1, np.complex64)
        lib.msresamp2_crcf_execute(self._h, _c64(x), y)
        return y

    def execute_block(self, x):
        x = _c64(x)
        interp = self.type == LIQUID_RESAMP_INTERP
        n = x.size if interp else x.size // self.M
        y = np.empty(n * self.M if interp else n, np.complex64)
        lib.msresamp2_crcf_execute_block(self._h, _ptr(x), n, _ptr(y))
        return y

    def execute_block_dev(self, d_x, n, d_y, stream=0):
        lib.msresamp2_crcf_execute_block_dev(self._h, d_x, n, d_y, stream)


class asgramcf(_Handle):
    """liquid.h:1333-1395 (src/fft/src/asgram.c): ASCII spectrogram; execute() -> (string, peak dB, peak frequency)."""
    _destroy = "asgramcf_destroy"

    def __init__(self, nfft):
        self.nfft = nfft
        super().__init__(lib.asgramcf_create(nfft))

    @classmethod
    def create(cls, nfft):
        return cls(nfft)

    def reset(self):
        lib.asgramcf_reset(self._h)

    def set_scale(self, offset, scale):
        lib.asgramcf_set_scale(self._h, offset, scale)

    def push(self, v):
        v = complex(v)
        lib.asgramcf_push(self._h, _cf(v.real, v.imag))

    def write(self, x):
        x = _c64(x)
        lib.asgramcf_write(self._h, x, x.size)

    def execute(self):
        buf = C.create_string_buffer(self.nfft + 1)
        pv, pf = C.c_float(), C.c_float()
        lib.asgramcf_execute(self._h, buf, C.byref(pv), C.byref(pf))
        return buf.raw[:self.nfft].decode("ascii"), pv.value, pf.value

    def print(self):
        lib.asgramcf_print(self._h)


OFDMFRAME_SCTYPE_NULL, OFDMFRAME_SCTYPE_PILOT, OFDMFRAME_SCTYPE_DATA = 0, 1, 2


def ofdmframe_init_default_sctype(M):
    p = np.empty(M, np.uint8)
    lib.ofdmframe_init_default_sctype(M, p)
    return p


class ofdmframesync(_Handle):
    """liquid.h:5878-5911 (src/multichannel/src/ofdmframesync.c).  callback(row, p) -> int receives every equalised
    symbol as a complex64 array of M subcarriers (a copy) and the allocation; a non-zero return ends the frame."""
    _destroy = "ofdmframesync_destroy"

    def __init__(self, M, cp_len, taper_len, p=None, callback=None):
        self.M, self.cp, self.tl = M, cp_len, taper_len
        self._p = None if p is None else np.ascontiguousarray(p, np.uint8)
        self._user_cb = callback

        def _cb(yp, pp, m, ud):
            if self._user_cb is None:
                return 0
            row = np.ctypeslib.as_array(C.cast(yp, C.POINTER(C.c_float)), shape=(2 * m,)).view(np.complex64).copy()
            alloc = np.ctypeslib.as_array(pp, shape=(m,)).copy()
            return int(self._user_cb(row, alloc) or 0)
        self._cb = OFDMFRAMESYNC_CALLBACK(_cb)         # kept alive with the object
        super().__init__(lib.ofdmframesync_create(M, cp_len, taper_len, None if self._p is None else self._p.ctypes.data,
                                                  self._cb, None))

    def reset(self):
        lib.ofdmframesync_reset(self._h)

    def execute(self, x):
        x = _c64(x)
        lib.ofdmframesync_execute(self._h, _ptr(x), x.size)

    def get_cfo(self):
        return lib.ofdmframesync_get_cfo(self._h)

    def get_rssi(self):
        return lib.ofdmframesync_get_rssi(self._h)

    def is_frame_open(self):
        return bool(lib.ofdmframesync_is_frame_open(self._h))


class ofdmframegen(_Handle):
    """liquid.h:5836-5876 (src/multichannel/src/ofdmframegen.c): OFDM symbol generator; p = subcarrier allocation
    (uint8 array) or None for the default one.  writesymbols() runs n data symbols in one call."""
    _destroy = "ofdmframegen_destroy"

    def __init__(self, M, cp_len, taper_len, p=None):
        self.M, self.cp, self.tl = M, cp_len, taper_len
        self._p = None if p is None else np.ascontiguousarray(p, np.uint8)
        super().__init__(lib.ofdmframegen_create(M, cp_len, taper_len, None if self._p is None else self._p.ctypes.data))

    @classmethod
    def create(cls, M, cp_len, taper_len, p=None):
        return cls(M, cp_len, taper_len, p)

    def reset(self):
        lib.ofdmframegen_reset(self._h)

    def print(self):
        lib.ofdmframegen_print(self._h)

    def sctype(self):
        p = np.empty(self.M, np.uint8)
        lib.ofdmframegen_get_sctype(self._h, p)
        return p

    def _sym(self, fn):
        y = np.empty(self.M + self.cp, np.complex64)
        fn(self._h, y)
        return y

    def write_S0a(self):
        return self._sym(lib.ofdmframegen_write_S0a)

    def write_S0b(self):
        return self._sym(lib.ofdmframegen_write_S0b)

    def write_S1(self):
        return self._sym(lib.ofdmframegen_write_S1)

    def writesymbol(self, x):
        y = np.empty(self.M + self.cp, np.complex64)
        lib.ofdmframegen_writesymbol(self._h, _c64(x), y)
        return y

    def writesymbols(self, x):
        x = _c64(x)
        n = x.size // self.M
        y = np.empty(n * (self.M + self.cp), np.complex64)
        lib.ofdmframegen_writesymbols(self._h, _ptr(x), n, _ptr(y))
        return y

    def writesymbols_dev(self, d_x, n, d_y, stream=0):
        lib.ofdmframegen_writesymbols_dev(self._h, d_x, n, d_y, stream)

    def writetail(self):
        y = np.empty(max(self.tl, 1), np.complex64)
        lib.ofdmframegen_writetail(self._h, y)
        return y[:self.tl]


class fftfilt_crcf(_Handle):
    """liquid.h:2192-2232 (src/filter/src/fftfilt.c): FIR filtering of blocks of n samples through 2n-point FFTs."""
    _destroy = "fftfilt_crcf_destroy"

    def __init__(self, h, n):
        hh = _f32(h)
        self.n = n
        super().__init__(lib.fftfilt_crcf_create(hh, len(hh), n))

    @classmethod
    def create(cls, h, n):
        return cls(h, n)

    def reset(self):
        lib.fftfilt_crcf_reset(self._h)

    def set_scale(self, scale):
        lib.fftfilt_crcf_set_scale(self._h, scale)

    def get_length(self):
        return lib.fftfilt_crcf_get_length(self._h)

    def execute(self, x):
        y = np.empty(self.n, np.complex64)
        lib.fftfilt_crcf_execute(self._h, _c64(x), y)
        return y

    def execute_block(self, x):
        x = _c64(x)
        nb = x.size // self.n
        y = np.empty(nb * self.n, np.complex64)
        lib.fftfilt_crcf_execute_block(self._h, _ptr(x), nb, _ptr(y))
        return y

    def execute_block_dev(self, d_x, n_blocks, d_y, stream=0):
        lib.fftfilt_crcf_execute_block_dev(self._h, d_x, n_blocks, d_y, stream)


class firinterp_crcf(_Handle):
    """liquid.h:2498-2556 (src/filter/src/firinterp.c): n*M outputs from n inputs."""
    _destroy = "firinterp_crcf_destroy"

    def __init__(self, M, h):
        self.M = M
        hh = _f32(h)
        super().__init__(lib.firinterp_crcf_create(M, hh, len(hh)))

    @classmethod
    def create(cls, M, h):
        return cls(M, h)

    @classmethod
    def create_kaiser(cls, M, m, As):
        q = cls.__new__(cls)
        q.M = M
        _Handle.__init__(q, lib.firinterp_crcf_create_kaiser(M, m, As))
        return q

    @classmethod
    def create_prototype(cls, ftype, M, m, beta, dt=0.0):
        q = cls.__new__(cls)
        q.M = M
        _Handle.__init__(q, lib.firinterp_crcf_create_prototype(ftype, M, m, beta, dt))
        return q

    def reset(self):
        lib.firinterp_crcf_reset(self._h)

    def execute(self, v):
        v = complex(v)
        y = np.empty(self.M, np.complex64)
        lib.firinterp_crcf_execute(self._h, _cf(v.real, v.imag), y)
        return y

    def execute_block(self, x):
        x = _c64(x)
        y = np.empty(x.size * self.M, np.complex64)
        lib.firinterp_crcf_execute_block(self._h, _ptr(x), x.size, _ptr(y))
        return y

    def execute_block_dev(self, d_x, n, d_y, stream=0):
        lib.firinterp_crcf_execute_block_dev(self._h, d_x, n, d_y, stream)


class spgramcf(_Handle):
    """liquid.h:1230-1316: windowed-FFT spectral periodogram (src/fft/src/spgram.c)."""
    _destroy = "spgramcf_destroy"

    def __init__(self, nfft, window):
        self.nfft = nfft
        w = _f32(window)
        super().__init__(lib.spgramcf_create(nfft, w, len(w)))

    @classmethod
    def create(cls, nfft, window):
        return cls(nfft, window)

    @classmethod
    def create_kaiser(cls, nfft, window_len, beta):
        q = cls.__new__(cls)
        q.nfft = nfft
        _Handle.__init__(q, lib.spgramcf_create_kaiser(nfft, window_len, beta))
        return q

    @classmethod
    def create_default(cls, nfft):
        q = cls.__new__(cls)
        q.nfft = nfft
        _Handle.__init__(q, lib.spgramcf_create_default(nfft))
        return q

    def reset(self):
        lib.spgramcf_reset(self._h)

    def push(self, v):
        v = complex(v)
        lib.spgramcf_push(self._h, _cf(v.real, v.imag))

    def write(self, x):
        x = _c64(x)
        lib.spgramcf_write(self._h, x, len(x))

    def execute(self):
        X = np.empty(self.nfft, np.complex64)
        lib.spgramcf_execute(self._h, X)
        return X

    def execute_psd(self):
        out = np.empty(self.nfft, np.float32)
        lib.spgramcf_execute_psd(self._h, out)
        return out

    def accumulate_psd(self, x, alpha):
        x = _c64(x)
        lib.spgramcf_accumulate_psd(self._h, x, alpha, len(x))

    def accumulate_psd_dev(self, d_x, alpha, n, stream=0):
        lib.spgramcf_accumulate_psd_dev(self._h, d_x, alpha, n, stream)

    def write_accumulation(self):
        out = np.empty(self.nfft, np.float32)
        lib.spgramcf_write_accumulation(self._h, out)
        return out

    def estimate_psd(self, x):
        x = _c64(x)
        out = np.empty(self.nfft, np.float32)
        lib.spgramcf_estimate_psd(self._h, x, len(x), out)
        return out

    def estimate_psd_dev(self, d_x, n, stream=0):
        out = np.empty(self.nfft, np.float32)
        lib.spgramcf_estimate_psd_dev(self._h, d_x, n, out, stream)
        return out

    def num_transforms(self):
        return lib.spgramcf_get_num_transforms(self._h)


class firpfbch_cccf(_Handle):
    """liquid.h:5660-5728, complex taps (+ analyzer_push / analyzer_run, firpfbch.c:369, :383)."""
    _destroy = "firpfbch_cccf_destroy"

    def __init__(self, type, M, p, h):
        self.type, self.M = type, M
        super().__init__(lib.firpfbch_cccf_create(type, M, p, _c64(h)))

    @classmethod
    def create(cls, type, M, p, h):
        return cls(type, M, p, h)

    @classmethod
    def create_kaiser(cls, type, M, m, As):
        q = cls.__new__(cls)
        q.type, q.M = type, M
        _Handle.__init__(q, lib.firpfbch_cccf_create_kaiser(type, M, m, As))
        return q

    @classmethod
    def create_rnyquist(cls, type, M, m, beta, ftype=LIQUID_FIRFILT_RRC):
        q = cls.__new__(cls)
        q.type, q.M = type, M
        _Handle.__init__(q, lib.firpfbch_cccf_create_rnyquist(type, M, m, beta, ftype))
        return q

    def reset(self):
        lib.firpfbch_cccf_reset(self._h)

    def print(self):
        lib.firpfbch_cccf_print(self._h)

    def analyzer_execute(self, x):
        y = np.empty(self.M, np.complex64)
        lib.firpfbch_cccf_analyzer_execute(self._h, _c64(x), y)
        return y

    def synthesizer_execute(self, x):
        y = np.empty(self.M, np.complex64)
        lib.firpfbch_cccf_synthesizer_execute(self._h, _c64(x), y)
        return y

    def analyzer_push(self, v):
        v = complex(v)
        lib.firpfbch_cccf_analyzer_push(self._h, _cf(v.real, v.imag))

    def analyzer_run(self, k):
        y = np.empty(self.M, np.complex64)
        lib.firpfbch_cccf_analyzer_run(self._h, k, y)
        return y

    def execute_block(self, x, out=None):
        x = _c64(x)
        n_frames = x.size // self.M
        assert n_frames * self.M == x.size
        y = out if out is not None else np.empty(x.size, np.complex64)
        fn = (lib.firpfbch_cccf_analyzer_execute_block if self.type == LIQUID_ANALYZER
              else lib.firpfbch_cccf_synthesizer_execute_block)
        fn(self._h, _ptr(x), n_frames, _ptr(y))
        return y.reshape(n_frames, self.M) if out is None else y

    def execute_block_dev(self, d_x, n_frames, d_y, stream=0):
        lib.firpfbch_cccf_execute_block_dev(self._h, d_x, n_frames, d_y, stream)

class firfilt_crcf(_Handle):
    """liquid.h:1985-2074."""
    _destroy = "firfilt_crcf_destroy"

    def __init__(self, h):
        super().__init__(lib.firfilt_crcf_create(_f32(h), len(h)))

    @classmethod
    def create(cls, h):
        return cls(h)

    @classmethod
    def create_kaiser(cls, n, fc, As, mu=0.0):
        q = cls.__new__(cls)
        _Handle.__init__(q, lib.firfilt_crcf_create_kaiser(n, fc, As, mu))
        return q

    @classmethod
    def create_rnyquist(cls, type, k, m, beta, mu=0.0):
        q = cls.__new__(cls)
        _Handle.__init__(q, lib.firfilt_crcf_create_rnyquist(type, k, m, beta, mu))
        return q

    @classmethod
    def create_rect(cls, n):
        q = cls.__new__(cls)
        _Handle.__init__(q, lib.firfilt_crcf_create_rect(n))
        return q

    def recreate(self, h):
        self._h = C.c_void_p(lib.firfilt_crcf_recreate(self._h, _f32(h), len(h)))
        return self

    def reset(self):
        lib.firfilt_crcf_reset(self._h)

    def print(self):
        lib.firfilt_crcf_print(self._h)

    def set_scale(self, s):
        lib.firfilt_crcf_set_scale(self._h, s)

    def get_length(self):
        return lib.firfilt_crcf_get_length(self._h)

    def push(self, v):
        v = complex(v)
        lib.firfilt_crcf_push(self._h, _cf(v.real, v.imag))

    def execute(self):
        y = np.empty(1, np.complex64)
        lib.firfilt_crcf_execute(self._h, y)
        return y[0]

    def execute_block(self, x, out=None):
        x = _c64(x)
        y = out if out is not None else np.empty_like(x)
        lib.firfilt_crcf_execute_block(self._h, _ptr(x), x.size, _ptr(y))
        return y

    def execute_block_dev(self, d_x, n, d_y, stream=0):
        lib.firfilt_crcf_execute_block_dev(self._h, d_x, n, d_y, stream)

    def execute_streams_dev(self, d_x, n_streams, n, d_y, stream=0):
        lib.firfilt_crcf_execute_streams_dev(self._h, d_x, n_streams, n, d_y, stream)

    def freqresponse(self, fc):
        H = np.empty(1, np.complex64)
        lib.firfilt_crcf_freqresponse(self._h, fc, H)
        return H[0]

    def groupdelay(self, fc):
        return lib.firfilt_crcf_groupdelay(self._h, fc)


class firpfb_crcf(_Handle):
    """liquid.h:2392-2470."""
    _destroy = "firpfb_crcf_destroy"

    def __init__(self, M, h):
        self.M = M
        super().__init__(lib.firpfb_crcf_create(M, _f32(h), len(h)))

    @classmethod
    def create_kaiser(cls, M, m, fc, As):
        q = cls.__new__(cls)
        q.M = M
        _Handle.__init__(q, lib.firpfb_crcf_create_kaiser(M, m, fc, As))
        return q

    @classmethod
    def create_rnyquist(cls, type, M, k, m, beta):
        q = cls.__new__(cls)
        q.M = M
        _Handle.__init__(q, lib.firpfb_crcf_create_rnyquist(type, M, k, m, beta))
        return q

    @classmethod
    def create_drnyquist(cls, type, M, k, m, beta):
        q = cls.__new__(cls)
        q.M = M
        _Handle.__init__(q, lib.firpfb_crcf_create_drnyquist(type, M, k, m, beta))
        return q

    def reset(self):
        lib.firpfb_crcf_reset(self._h)

    def set_scale(self, g):
        lib.firpfb_crcf_set_scale(self._h, g)

    def push(self, v):
        v = complex(v)
        lib.firpfb_crcf_push(self._h, _cf(v.real, v.imag))

    def execute(self, i):
        y = np.empty(1, np.complex64)
        lib.firpfb_crcf_execute(self._h, i, y)
        return y[0]

    def execute_block_dev(self, d_x, d_idx, n, d_y, stream=0):
        lib.firpfb_crcf_execute_block_dev(self._h, d_x, d_idx, n, d_y, stream)

    def execute_block(self, x, idx):
        x = _c64(x)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        y = np.empty_like(x)
        lib.firpfb_crcf_execute_block(self._h, x, idx, x.size, y)
        return y


class dotprod_crcf(_Handle):
    """liquid.h:503-541."""
    _destroy = "dotprod_crcf_destroy"

    def __init__(self, h):
        self.n = len(h)
        super().__init__(lib.dotprod_crcf_create(_f32(h), len(h)))

    def execute(self, x):
        y = np.empty(1, np.complex64)
        lib.dotprod_crcf_execute(self._h, _c64(x), y)
        return y[0]

    def execute_batch_dev(self, d_x, n_vec, stride, d_y, stream=0):
        lib.dotprod_crcf_execute_batch_dev(self._h, d_x, n_vec, stride, d_y, stream)

    @staticmethod
    def run(h, x):
        y = np.empty(1, np.complex64)
        lib.dotprod_crcf_run(_f32(h), _c64(x), len(h), y)
        return y[0]


class fftplan(_Handle):
    """liquid.h:1148-1215: the plan borrows x / y for its lifetime (fft_common.c:35-36)."""
    _destroy = "fft_destroy_plan"

    def __init__(self, n, x, y, direction, flags=0):
        assert x.dtype == np.complex64 and y.dtype == np.complex64 and x.size >= n and y.size >= n
        self.n, self.x, self.y = n, x, y
        super().__init__(lib.fft_create_plan(n, _ptr(x), _ptr(y), direction, flags))

    def execute(self):
        lib.fft_execute(self._h)

    def execute_batch(self, x, y=None):
        x = _c64(x)
        y = np.empty_like(x) if y is None else y
        lib.fft_execute_batch(self._h, x, y, x.size // self.n)
        return y

    def execute_batch_dev(self, d_x, d_y, batch, stream=0):
        lib.fft_execute_batch_dev(self._h, d_x, d_y, batch, stream)


def fft_run(x, direction=LIQUID_FFT_FORWARD):
    x = _c64(x)
    y = np.empty_like(x)
    lib.fft_run(x.size, x, y, direction, 0)
    return y


def synth_input_dev(d_x, n, seed=0x5EED0000, first=0, stream=0):
    lib.liquid_cuda_synth_input_dev(d_x, seed, first, n, stream)
