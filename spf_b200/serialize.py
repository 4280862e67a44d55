"""Serialized key / ciphertext layouts of the reference over the C ABI.

bincode 1.3.3 (Cargo.lock:244-245), fixed-width little-endian: every sunscreen_tfhe entity is one
sequence ``u64 length || elements`` (sunscreen_tfhe/src/dst.rs:31-33); ``ComputeKey`` is
``bs_key || ks_key || ss_key || auto_key`` (parasol_runtime/src/crypto/keys.rs:306-318).
Loading follows ``safe_bincode::deserialize`` (parasol_runtime/src/safe_bincode.rs:16-27): byte limit
``GetSize::get_size``, trailing bytes allowed, lengths checked against the parameter set; a
malformed buffer raises ``SpfError`` where the reference returns ``Err``.  No GPU is needed.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import Params, SpfError, default_128, lib

LWE0, LWE1, GLWE1, GLEV1 = 0, 1, 2, 3  # spf_ct_kind


def _err(rc: int) -> SpfError:
    return SpfError(rc, (lib().spf_b200_last_error(None) or b"").decode())


def _buf(data) -> np.ndarray:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1)
    return np.ascontiguousarray(a)


def compute_key_size(params: Params | None = None) -> int:
    """Exact length of ``bincode::serialize(&ComputeKey)``."""
    p = params or default_128()
    return int(lib().spf_b200_serialized_size_compute_key(C.byref(p)))


def compute_key_limit(params: Params | None = None) -> int:
    """``ComputeKey::get_size`` (keys.rs:326-349)."""
    p = params or default_128()
    return int(lib().spf_b200_serialized_limit_compute_key(C.byref(p)))


def load_compute_key(data, params: Params | None = None):
    """-> (bs_key complex128, ks_key uint64, ss_key complex128, auto_key complex128): zero-copy views."""
    p = params or default_128()
    l = lib()
    b = _buf(data)
    off = (C.c_size_t * 4)()
    rc = l.spf_b200_parse_compute_key(C.byref(p), b.ctypes.data, b.size, C.byref(off))
    if rc:
        raise _err(rc)
    n = (l.spf_b200_len_bsk(C.byref(p)), l.spf_b200_len_ksk(C.byref(p)), l.spf_b200_len_ssk(C.byref(p)),
         l.spf_b200_len_ak(C.byref(p)))
    out = []
    for o, cnt, dt in zip(off, n, (np.complex128, np.uint64, np.complex128, np.complex128)):
        out.append(np.frombuffer(b, dtype=dt, count=cnt, offset=int(o)))
    return tuple(out)


def dump_compute_key(bsk_fft, ksk, ssk_fft, ak_fft, params: Params | None = None) -> bytes:
    p = params or default_128()
    l = lib()
    arrs = [np.ascontiguousarray(a, dtype=dt).reshape(-1)
            for a, dt in ((bsk_fft, np.complex128), (ksk, np.uint64), (ssk_fft, np.complex128), (ak_fft, np.complex128))]
    want = (l.spf_b200_len_bsk(C.byref(p)), l.spf_b200_len_ksk(C.byref(p)), l.spf_b200_len_ssk(C.byref(p)),
            l.spf_b200_len_ak(C.byref(p)))
    for a, n, name in zip(arrs, want, ("bs_key", "ks_key", "ss_key", "auto_key")):
        if a.size != n:
            raise SpfError(-1, f"{name} has {a.size} elements, params require {n}")
    out = np.empty(compute_key_size(p), dtype=np.uint8)
    written = C.c_size_t()
    rc = l.spf_b200_write_compute_key(C.byref(p), out.ctypes.data, out.size, arrs[0].ctypes.data, arrs[1].ctypes.data,
                                      arrs[2].ctypes.data, arrs[3].ctypes.data, C.byref(written))
    if rc:
        raise _err(rc)
    return out[:written.value].tobytes()


def ciphertext_size(kind: int, params: Params | None = None) -> int:
    p = params or default_128()
    return int(lib().spf_b200_serialized_size_ciphertext(C.byref(p), kind))


def load_ciphertext(kind: int, data, params: Params | None = None) -> np.ndarray:
    """``safe_bincode::deserialize::<L0Lwe|L1Lwe|L1Glwe|L1Glev Ciphertext>`` -> uint64 view."""
    p = params or default_128()
    b = _buf(data)
    off = C.c_size_t()
    rc = lib().spf_b200_parse_ciphertext(C.byref(p), kind, b.ctypes.data, b.size, C.byref(off))
    if rc:
        raise _err(rc)
    n = ciphertext_size(kind, p) // 8 - 1
    return np.frombuffer(b, dtype=np.uint64, count=n, offset=int(off.value))


def dump_ciphertext(kind: int, ct, params: Params | None = None) -> bytes:
    p = params or default_128()
    a = np.ascontiguousarray(ct, dtype=np.uint64).reshape(-1)
    size = ciphertext_size(kind, p)
    if size == 0:
        raise SpfError(-1, "unknown ciphertext kind")
    if a.size != size // 8 - 1:
        raise SpfError(-1, f"ciphertext has {a.size} elements, params require {size // 8 - 1}")
    out = np.empty(size, dtype=np.uint8)
    written = C.c_size_t()
    rc = lib().spf_b200_write_ciphertext(C.byref(p), kind, a.ctypes.data, out.ctypes.data, out.size, C.byref(written))
    if rc:
        raise _err(rc)
    return out[:written.value].tobytes()
