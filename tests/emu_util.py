"""ctypes binding of spf_b200/csrc/libspf_emu.so: the device kernel bodies executed on the host
(64 host threads per team, pthread barrier for bar.sync).  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "spf_b200", "csrc", "libspf_emu.so")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_c64p = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PATH):
            subprocess.check_call(["make", "-C", ROOT, "spf_b200/csrc/libspf_emu.so"], stdout=subprocess.DEVNULL)
        l = C.CDLL(PATH)
        l.emu_poly_fft.argtypes = [_u64p, _c64p]
        l.emu_poly_ifft.argtypes = [_c64p, _u64p]
        l.emu_import_fft.argtypes = [_c64p, _c64p, C.c_size_t]
        l.emu_export_fft.argtypes = [_c64p, _c64p, C.c_size_t]
        l.emu_cmux.argtypes = [_u64p, C.c_void_p, _u64p, _c64p, C.c_int, C.c_int]
        l.emu_cmux_wide.argtypes = [_u64p, C.c_void_p, _u64p, _c64p, C.c_int, C.c_int]
        l.emu_pbs.argtypes = [_u64p, _u64p, C.c_void_p, _c64p] + [C.c_int] * 5
        l.emu_pbs_quad.argtypes = [_u64p, _u64p, C.c_void_p, _c64p] + [C.c_int] * 5
        l.emu_pbs_variant.argtypes = [_u64p, _u64p, C.c_void_p, _c64p] + [C.c_int] * 7
        l.emu_trace_ss.argtypes = [_u64p, C.c_void_p, C.c_void_p, _c64p, _c64p] + [C.c_int] * 8
        l.emu_f64_to_torus.argtypes = [C.c_double]
        l.emu_f64_to_torus.restype = C.c_uint64
        l.emu_i32_to_f64.argtypes = [C.c_int32]
        l.emu_i32_to_f64.restype = C.c_double
        _lib = l
    return _lib


def to_device_scale(x):
    out = np.zeros_like(x)
    lib().emu_import_fft(np.ascontiguousarray(x), out, x.size // 1024)
    return out


def to_reference_scale(x):
    out = np.zeros_like(x)
    lib().emu_export_fft(np.ascontiguousarray(x), out, x.size // 1024)
    return out
