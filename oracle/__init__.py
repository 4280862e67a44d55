"""CPU ORACLE for spf_b200 -- TEST INFRASTRUCTURE ONLY.

ctypes binding over ``oracle/libspf_oracle.so`` (built from ``spf_oracle.c`` by
``make -C oracle``), a plain-C restatement of the reference's TFHE circuit-bootstrapping
path (see ``spf_oracle.h`` for the parity status and citations).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package; ``spf_b200`` (the product) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libspf_oracle.so")


class Radix(C.Structure):
    _fields_ = [("radix_log", C.c_uint32), ("count", C.c_uint32)]

    def __repr__(self):
        return f"Radix(log={self.radix_log}, count={self.count})"


class Params(C.Structure):
    """parasol_runtime/src/params.rs:59-91, flattened (orc_params)."""

    _fields_ = [
        ("lwe_n", C.c_uint32),
        ("lwe_std", C.c_double),
        ("glwe_k", C.c_uint32),
        ("glwe_n", C.c_uint32),
        ("glwe_std", C.c_double),
        ("cbs", Radix),
        ("pbs", Radix),
        ("ks", Radix),
        ("pfks", Radix),
        ("ss", Radix),
        ("tr", Radix),
    ]


class Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4), ("have_spare", C.c_int), ("spare", C.c_double)]


def build(force: bool = False) -> str:
    """Compile the oracle library if missing (or stale)."""
    src = os.path.join(_HERE, "spf_oracle.c")
    hdr = os.path.join(_HERE, "spf_oracle.h")
    fast = os.path.join(_HERE, "libspf_oracle_fast.so")
    srcs = (src, hdr, os.path.join(_HERE, "fft_avx2.h"))
    stale = any((not os.path.exists(l)) or any(os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(l) for f in srcs)
                for l in (_LIB_PATH, fast))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_fast_lib = None
_FAST_LIB_PATH = os.path.join(_HERE, "libspf_oracle_fast.so")


def lib(fast: bool = False) -> C.CDLL:
    """The oracle library.  fast=True: the AVX2/FMA build of the same source (-DORC_FAST, see spf_oracle.c header),
    used ONLY as the timed CPU baseline of bench.py (cpu_baseline / --impl reference); every parity check uses the
    strict build."""
    global _lib, _fast_lib
    if fast:
        if _fast_lib is None:
            build()
            _fast_lib = C.CDLL(_FAST_LIB_PATH)
            _declare(_fast_lib)
        return _fast_lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_c64p = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_PP = C.POINTER(Params)


def _declare(l):
    sz = C.c_size_t
    u32 = C.c_uint32
    l.orc_default_128.argtypes = [_PP]
    for name in ("orc_size_glwe", "orc_size_bsk_fft", "orc_size_ksk", "orc_size_ak_fft", "orc_size_ssk_fft"):
        getattr(l, name).argtypes = [_PP]
        getattr(l, name).restype = sz
    for name in ("orc_size_glev", "orc_size_ggsw", "orc_size_ggsw_fft"):
        getattr(l, name).argtypes = [_PP, Radix]
        getattr(l, name).restype = sz
    l.orc_fft_forward.argtypes = [_f64p, _c64p, u32]
    l.orc_fft_reverse.argtypes = [_c64p, _f64p, u32]
    l.orc_poly_fft.argtypes = [_u64p, _c64p, u32]
    l.orc_poly_ifft.argtypes = [_c64p, _u64p, u32]
    l.orc_mod_pow2_q_f64.argtypes = [_u64p, _f64p, sz]
    l.orc_radix_round.argtypes = [C.c_uint64, Radix]
    l.orc_radix_round.restype = C.c_uint64
    l.orc_next_decomp.argtypes = [_u64p, _u64p, sz, u32]
    l.orc_modulus_switch.argtypes = [C.c_uint64, u32, u32, u32]
    l.orc_modulus_switch.restype = C.c_uint64
    l.orc_poly_pow_k.argtypes = [_u64p, _u64p, u32, u32]
    l.orc_shr_round.argtypes = [_u64p, _u64p, sz, u32]
    l.orc_poly_mul_monomial.argtypes = [_u64p, u32, C.c_int64]
    l.orc_generate_lut.argtypes = [_u64p, _u64p, u32, u32, u32]
    l.orc_glwe_ggsw_mad.argtypes = [_c64p, _u64p, _c64p, _PP, Radix]
    l.orc_glwe_fft_ifft.argtypes = [_c64p, _u64p, _PP]
    l.orc_cmux.argtypes = [_u64p, _u64p, _u64p, _c64p, _PP, Radix]
    l.orc_glev_cmux.argtypes = [_u64p, _u64p, _u64p, _c64p, _PP, Radix, Radix]
    l.orc_keyswitch_glwe.argtypes = [_u64p, _u64p, _c64p, _PP, Radix]
    l.orc_trace.argtypes = [_u64p, _u64p, _c64p, _PP]
    l.orc_scheme_switch_fft.argtypes = [_c64p, _u64p, _c64p, _PP]
    l.orc_pbs_generalized.argtypes = [_u64p, _u64p, _u64p, _c64p, u32, u32, _PP]
    l.orc_pbs_univariate.argtypes = [_u64p, _u64p, _u64p, _c64p, _PP]
    l.orc_cbs_lut.argtypes = [_u64p, _PP]
    l.orc_cbs_pbs_stage.argtypes = [_u64p, _u64p, _c64p, _PP]
    l.orc_cbs_trace_stage.argtypes = [_u64p, _u64p, _c64p, _PP]
    l.orc_circuit_bootstrap.argtypes = [_c64p, _u64p, _c64p, _c64p, _c64p, _PP]
    l.orc_keyswitch_lwe.argtypes = [_u64p, _u64p, _u64p, _PP]
    l.orc_sample_extract.argtypes = [_u64p, _u64p, u32, _PP]
    l.orc_glwe_add.argtypes = [_u64p, _u64p, _u64p, _PP]
    l.orc_glwe_not.argtypes = [_u64p, _u64p, _PP]
    l.orc_glwe_mul_xn.argtypes = [_u64p, _u64p, u32, _PP]
    l.orc_multiply_glwe_ggsw.argtypes = [_u64p, _u64p, _c64p, _PP]
    l.orc_circuit_bootstrap_batch.argtypes = [_c64p, _u64p, sz, _c64p, _c64p, _c64p, _PP, C.c_int]
    l.orc_cmux_batch.argtypes = [_u64p, _u64p, _u64p, _c64p, sz, _PP, C.c_int]
    l.orc_keyswitch_lwe_batch.argtypes = [_u64p, _u64p, sz, _u64p, _PP, C.c_int]
    l.orc_cmux_batch_ptrs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, sz, _PP, C.c_int]
    l.orc_cmux_batch_ptrs.restype = None
    l.orc_rng_seed.argtypes = [C.POINTER(Rng), C.c_uint64]
    l.orc_rng_u64.argtypes = [C.POINTER(Rng)]
    l.orc_rng_u64.restype = C.c_uint64
    l.orc_keygen_secret.argtypes = [C.POINTER(Rng), _u64p, _u64p, _PP]
    l.orc_keygen_compute.argtypes = [C.POINTER(Rng), _u64p, _u64p, _c64p, _u64p, _c64p, _c64p, _PP, C.c_int]
    l.orc_encrypt_lwe.argtypes = [C.POINTER(Rng), _u64p, _u64p, u32, C.c_double, C.c_uint64]
    l.orc_decrypt_lwe_raw.argtypes = [_u64p, _u64p, u32]
    l.orc_decrypt_lwe_raw.restype = C.c_uint64
    l.orc_decode.argtypes = [C.c_uint64, u32]
    l.orc_decode.restype = C.c_uint64
    l.orc_encrypt_glwe.argtypes = [C.POINTER(Rng), _u64p, _u64p, _u64p, _PP]
    l.orc_decrypt_glwe_raw.argtypes = [_u64p, _u64p, _u64p, _PP]
    l.orc_encrypt_glev.argtypes = [C.POINTER(Rng), _u64p, _u64p, _u64p, _PP, Radix]
    l.orc_rlwe_generate_public_key.argtypes = [C.POINTER(Rng), _u64p, _u64p, _PP]
    l.orc_rlwe_sample_randomness.argtypes = [C.POINTER(Rng), _u64p, _u64p, _u64p, _PP]
    l.orc_rlwe_encrypt_public.argtypes = [_u64p, _u64p, _u64p, _u64p, _u64p, _u64p, _PP]
    l.orc_encrypt_ggsw.argtypes = [C.POINTER(Rng), _u64p, _u64p, _u64p, _PP, Radix]
    l.orc_ggsw_fft.argtypes = [_c64p, _u64p, _PP, Radix]
    l.orc_ggsw_ifft.argtypes = [_u64p, _c64p, _PP, Radix]
    l.orc_hw_threads.restype = C.c_int
    l.orc_bench_fft_forward.argtypes = [C.c_uint32, C.c_int]
    l.orc_bench_fft_forward.restype = C.c_double


def default_128() -> Params:
    p = Params()
    lib().orc_default_128(C.byref(p))
    return p


def small_params(n: int = 256, lwe_n: int = 64) -> Params:
    """A toy parameter set (NOT secure) with the DEFAULT_128 radix shapes scaled so every op
    still decrypts correctly; used for fast CPU tests of the oracle's own logic."""
    p = default_128()
    p.glwe_n = n
    p.lwe_n = lwe_n
    p.lwe_std = 1e-9
    p.glwe_std = 1e-17
    return p


def hw_threads() -> int:
    return int(lib().orc_hw_threads())


KEY_SEED = 0xB2000001
INPUT_SEED = 0xB2000002


class Keys:
    """SecretKey + ComputeKey (parasol_runtime/src/crypto/keys.rs) generated by the oracle."""

    def __init__(self, params: Params | None = None, seed: int = KEY_SEED, nthreads: int | None = None):
        l = lib()
        self.params = params if params is not None else default_128()
        p = self.params
        self.rng = Rng()
        l.orc_rng_seed(C.byref(self.rng), seed)
        self.lwe0_sk = np.zeros(p.lwe_n, dtype=np.uint64)
        self.glwe1_sk = np.zeros(p.glwe_k * p.glwe_n, dtype=np.uint64)
        l.orc_keygen_secret(C.byref(self.rng), self.lwe0_sk, self.glwe1_sk, C.byref(p))
        self.bsk_fft = np.zeros(l.orc_size_bsk_fft(C.byref(p)), dtype=np.complex128)
        self.ksk = np.zeros(l.orc_size_ksk(C.byref(p)), dtype=np.uint64)
        self.ssk_fft = np.zeros(l.orc_size_ssk_fft(C.byref(p)), dtype=np.complex128)
        self.ak_fft = np.zeros(l.orc_size_ak_fft(C.byref(p)), dtype=np.complex128)
        nt = nthreads if nthreads is not None else hw_threads()
        l.orc_keygen_compute(C.byref(self.rng), self.lwe0_sk, self.glwe1_sk, self.bsk_fft, self.ksk,
                             self.ssk_fft, self.ak_fft, C.byref(p), nt)

    # sizes -------------------------------------------------------------------------------
    @property
    def lwe0_len(self):
        return self.params.lwe_n + 1

    @property
    def lwe1_len(self):
        return self.params.glwe_k * self.params.glwe_n + 1

    @property
    def glwe_len(self):
        return (self.params.glwe_k + 1) * self.params.glwe_n

    @property
    def glev_len(self):
        return self.glwe_len * self.params.cbs.count

    @property
    def ggsw_fft_len(self):
        return int(lib().orc_size_ggsw_fft(C.byref(self.params), self.params.cbs))


class Client:
    """Encryption / decryption (parasol_runtime/src/crypto/encryption.rs:127-452) over the oracle."""

    def __init__(self, keys: Keys, seed: int = INPUT_SEED):
        self.k = keys
        self.p = keys.params
        self.rng = Rng()
        lib().orc_rng_seed(C.byref(self.rng), seed)

    # L0 / L1 LWE -------------------------------------------------------------------------
    def encrypt_lwe_l0(self, bit: int) -> np.ndarray:
        ct = np.zeros(self.k.lwe0_len, dtype=np.uint64)
        lib().orc_encrypt_lwe(C.byref(self.rng), ct, self.k.lwe0_sk, self.p.lwe_n, self.p.lwe_std,
                              (int(bit) & 1) << 63)
        return ct

    def encrypt_lwe_l0_batch(self, bits) -> np.ndarray:
        return np.stack([self.encrypt_lwe_l0(b) for b in bits])

    def trivial_lwe_l0(self, bit: int) -> np.ndarray:
        ct = np.zeros(self.k.lwe0_len, dtype=np.uint64)
        ct[-1] = np.uint64((int(bit) & 1) << 63)
        return ct

    def decrypt_lwe_l0(self, ct: np.ndarray, plaintext_bits: int = 1) -> int:
        raw = lib().orc_decrypt_lwe_raw(np.ascontiguousarray(ct), self.k.lwe0_sk, self.p.lwe_n)
        return int(lib().orc_decode(raw, plaintext_bits))

    def encrypt_lwe_l1(self, bit: int) -> np.ndarray:
        ct = np.zeros(self.k.lwe1_len, dtype=np.uint64)
        lib().orc_encrypt_lwe(C.byref(self.rng), ct, self.k.glwe1_sk, self.p.glwe_k * self.p.glwe_n,
                              self.p.glwe_std, (int(bit) & 1) << 63)
        return ct

    def decrypt_lwe_l1(self, ct: np.ndarray, plaintext_bits: int = 1) -> int:
        raw = lib().orc_decrypt_lwe_raw(np.ascontiguousarray(ct), self.k.glwe1_sk, self.p.glwe_k * self.p.glwe_n)
        return int(lib().orc_decode(raw, plaintext_bits))

    def decrypt_lwe_l1_raw(self, ct: np.ndarray) -> int:
        return int(lib().orc_decrypt_lwe_raw(np.ascontiguousarray(ct), self.k.glwe1_sk, self.p.glwe_k * self.p.glwe_n))

    # L1 GLWE / GLEV / GGSW ---------------------------------------------------------------
    def encrypt_glwe_l1(self, bits, plaintext_bits: int = 1) -> np.ndarray:
        msg = np.zeros(self.p.glwe_n, dtype=np.uint64)
        bits = np.asarray(bits, dtype=np.uint64)
        msg[: len(bits)] = bits << np.uint64(64 - plaintext_bits)
        ct = np.zeros(self.k.glwe_len, dtype=np.uint64)
        lib().orc_encrypt_glwe(C.byref(self.rng), ct, msg, self.k.glwe1_sk, C.byref(self.p))
        return ct

    def trivial_glwe_l1(self, bits) -> np.ndarray:
        ct = np.zeros(self.k.glwe_len, dtype=np.uint64)
        bits = np.asarray(bits, dtype=np.uint64)
        ct[self.p.glwe_k * self.p.glwe_n: self.p.glwe_k * self.p.glwe_n + len(bits)] = bits << np.uint64(63)
        return ct

    def decrypt_glwe_l1_raw(self, ct: np.ndarray) -> np.ndarray:
        msg = np.zeros(self.p.glwe_n, dtype=np.uint64)
        lib().orc_decrypt_glwe_raw(msg, np.ascontiguousarray(ct), self.k.glwe1_sk, C.byref(self.p))
        return msg

    def decrypt_glwe_l1(self, ct: np.ndarray, plaintext_bits: int = 1) -> np.ndarray:
        return decode(self.decrypt_glwe_l1_raw(ct), plaintext_bits)

    # RLWE public key (keys.rs:20-70, ops/encryption/rlwe_encryption.rs) -------------------------
    def generate_public_key(self) -> np.ndarray:
        """PublicKey::generate (keys.rs:60-69): an encryption of the zero polynomial under glwe_1."""
        pk = np.zeros(self.k.glwe_len, dtype=np.uint64)
        lib().orc_rlwe_generate_public_key(C.byref(self.rng), pk, self.k.glwe1_sk, C.byref(self.p))
        return pk

    def rlwe_randomness(self):
        """(u, e0, e1) of rlwe_encrypt_public_impl (rlwe_encryption.rs:140-146) from the harness' seeded PRNG."""
        n = self.p.glwe_n
        u, e0, e1 = (np.zeros(n, dtype=np.uint64) for _ in range(3))
        lib().orc_rlwe_sample_randomness(C.byref(self.rng), u, e0, e1, C.byref(self.p))
        return u, e0, e1

    def encrypt_rlwe_l1(self, bits, pk: np.ndarray, randomness=None, plaintext_bits: int = 1) -> np.ndarray:
        """Encryption::encrypt_rlwe_l1 (encryption.rs:205-215) = rlwe_encode_encrypt_public with PlaintextBits(1)."""
        msg = np.zeros(self.p.glwe_n, dtype=np.uint64)
        bits = np.asarray(bits, dtype=np.uint64)
        msg[: len(bits)] = bits << np.uint64(64 - plaintext_bits)
        u, e0, e1 = randomness if randomness is not None else self.rlwe_randomness()
        return rlwe_encrypt_public(self.p, msg, pk, u, e0, e1)

    def encrypt_glev_l1(self, bits) -> np.ndarray:
        msg = np.zeros(self.p.glwe_n, dtype=np.uint64)
        msg[: len(bits)] = np.asarray(bits, dtype=np.uint64)
        ct = np.zeros(self.k.glev_len, dtype=np.uint64)
        lib().orc_encrypt_glev(C.byref(self.rng), ct, msg, self.k.glwe1_sk, C.byref(self.p), self.p.cbs)
        return ct

    def encrypt_ggsw_l1(self, bit: int) -> np.ndarray:
        """encrypt_ggsw_l1_secret (encryption.rs:225-246): GGSW of the constant poly `bit`, FFT'd."""
        l = lib()
        msg = np.zeros(self.p.glwe_n, dtype=np.uint64)
        msg[0] = int(bit) & 1
        ggsw = np.zeros(l.orc_size_ggsw(C.byref(self.p), self.p.cbs), dtype=np.uint64)
        l.orc_encrypt_ggsw(C.byref(self.rng), ggsw, msg, self.k.glwe1_sk, C.byref(self.p), self.p.cbs)
        out = np.zeros(self.k.ggsw_fft_len, dtype=np.complex128)
        l.orc_ggsw_fft(out, ggsw, C.byref(self.p), self.p.cbs)
        return out

    def ggsw_level_messages(self, ggsw_fft: np.ndarray) -> np.ndarray:
        """IFFT a GGSW-FFT and decrypt every (row, level) GLWE at plaintext_bits=(level+1)*logB,
        exactly as can_circuit_bootstrap_via_trace_ss does (circuit_bootstrapping.rs:777-803).
        Returns array [rows, levels, N] of decoded coefficients."""
        l = lib()
        p = self.p
        ggsw = np.zeros(l.orc_size_ggsw(C.byref(p), p.cbs), dtype=np.uint64)
        l.orc_ggsw_ifft(ggsw, np.ascontiguousarray(ggsw_fft), C.byref(p), p.cbs)
        rows, levels = p.glwe_k + 1, p.cbs.count
        g = ggsw.reshape(rows, levels, self.k.glwe_len)
        out = np.zeros((rows, levels, p.glwe_n), dtype=np.uint64)
        for r in range(rows):
            for lv in range(levels):
                out[r, lv] = self.decrypt_glwe_l1(g[r, lv], (lv + 1) * p.cbs.radix_log)
        return out

    def ggsw_phases(self, ggsw_fft: np.ndarray) -> np.ndarray:
        """Raw phases b - a.s of every (row, level) GLWE of a GGSW-FFT: array [rows, levels, N] of torus
        elements (used to state GPU-vs-oracle distances at the ciphertext level)."""
        l = lib()
        p = self.p
        ggsw = np.zeros(l.orc_size_ggsw(C.byref(p), p.cbs), dtype=np.uint64)
        l.orc_ggsw_ifft(ggsw, np.ascontiguousarray(ggsw_fft), C.byref(p), p.cbs)
        rows, levels = p.glwe_k + 1, p.cbs.count
        g = ggsw.reshape(rows, levels, self.k.glwe_len)
        return np.stack([np.stack([self.decrypt_glwe_l1_raw(g[r, lv]) for lv in range(levels)]) for r in range(rows)])

    def ggsw_expected_messages(self, bit: int) -> np.ndarray:
        """Plaintext every (row, level) GLWE of a fresh GGSW(bit) decodes to: row k: bit at
        coeff 0; rows j<k: -(bit * s_j) -- all at plaintext_bits=(level+1)*logB where the
        gadget factor q/B^(level+1) puts the integer at the LSB of the plaintext window."""
        p = self.p
        rows, levels = p.glwe_k + 1, p.cbs.count
        out = np.zeros((rows, levels, p.glwe_n), dtype=np.uint64)
        for lv in range(levels):
            pb = (lv + 1) * p.cbs.radix_log
            mask = np.uint64((1 << pb) - 1)
            for r in range(p.glwe_k):
                s = self.k.glwe1_sk[r * p.glwe_n:(r + 1) * p.glwe_n]
                out[r, lv] = (np.uint64(0) - s * np.uint64(bit)) & mask
            out[p.glwe_k, lv, 0] = bit
        return out

    def decrypt_ggsw_l1(self, ggsw_fft: np.ndarray) -> int:
        """decrypt_ggsw_l1 (encryption.rs:279-297): last row, first GLWE, coefficient 0."""
        msgs = self.ggsw_level_messages(ggsw_fft)
        return int(msgs[self.p.glwe_k, 0, 0] == 1)


def decode(torus: np.ndarray, plaintext_bits: int) -> np.ndarray:
    """Torus::decode (math/torus.rs:293-300), vectorised."""
    t = np.asarray(torus, dtype=np.uint64)
    rb = (t >> np.uint64(64 - plaintext_bits - 1)) & np.uint64(1)
    return ((t >> np.uint64(64 - plaintext_bits)) + rb) & np.uint64((1 << plaintext_bits) - 1)


def torus_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """|a-b| on the torus as a fraction of q (math/torus.rs:236-249), elementwise."""
    d = (np.asarray(a, dtype=np.uint64) - np.asarray(b, dtype=np.uint64)).astype(np.int64)
    return np.abs(d.astype(np.float64)) / 2.0 ** 64


# ---- thin functional wrappers used by the tests ---------------------------------------------

def rlwe_encrypt_public(params: Params, encoded_msg, pk, u, e0, e1) -> np.ndarray:
    """rlwe_encrypt_public_impl with given randomness (rlwe_encryption.rs:125-160)."""
    ct = np.zeros((params.glwe_k + 1) * params.glwe_n, dtype=np.uint64)
    a = [np.ascontiguousarray(x, dtype=np.uint64) for x in (encoded_msg, pk, u, e0, e1)]
    lib().orc_rlwe_encrypt_public(ct, a[0], a[1], a[2], a[3], a[4], C.byref(params))
    return ct


def poly_fft(p: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(p, dtype=np.uint64)
    out = np.zeros(len(p) // 2, dtype=np.complex128)
    lib().orc_poly_fft(p, out, len(p))
    return out


def poly_ifft(f: np.ndarray) -> np.ndarray:
    f = np.ascontiguousarray(f, dtype=np.complex128)
    out = np.zeros(len(f) * 2, dtype=np.uint64)
    lib().orc_poly_ifft(f, out, len(out))
    return out


def cmux(keys: Keys, d0, d1, ggsw_fft, radix=None) -> np.ndarray:
    p = keys.params
    out = np.zeros(keys.glwe_len, dtype=np.uint64)
    lib().orc_cmux(out, np.ascontiguousarray(d0), np.ascontiguousarray(d1), np.ascontiguousarray(ggsw_fft),
                   C.byref(p), radix if radix is not None else p.cbs)
    return out


def circuit_bootstrap(keys: Keys, lwe0: np.ndarray) -> np.ndarray:
    out = np.zeros(keys.ggsw_fft_len, dtype=np.complex128)
    lib().orc_circuit_bootstrap(out, np.ascontiguousarray(lwe0), keys.bsk_fft, keys.ak_fft, keys.ssk_fft,
                                C.byref(keys.params))
    return out


def circuit_bootstrap_batch(keys: Keys, lwe0: np.ndarray, nthreads: int | None = None, fast: bool = False) -> np.ndarray:
    lwe0 = np.ascontiguousarray(lwe0, dtype=np.uint64)
    b = lwe0.shape[0]
    out = np.zeros((b, keys.ggsw_fft_len), dtype=np.complex128)
    lib(fast).orc_circuit_bootstrap_batch(out, lwe0, b, keys.bsk_fft, keys.ak_fft, keys.ssk_fft,
                                      C.byref(keys.params), nthreads or hw_threads())
    return out


def cbs_pbs_stage(keys: Keys, lwe0: np.ndarray) -> np.ndarray:
    out = np.zeros(keys.glwe_len, dtype=np.uint64)
    lib().orc_cbs_pbs_stage(out, np.ascontiguousarray(lwe0), keys.bsk_fft, C.byref(keys.params))
    return out


def cbs_trace_stage(keys: Keys, glwe: np.ndarray) -> np.ndarray:
    out = np.zeros(keys.glev_len, dtype=np.uint64)
    lib().orc_cbs_trace_stage(out, np.ascontiguousarray(glwe), keys.ak_fft, C.byref(keys.params))
    return out


def scheme_switch(keys: Keys, glev: np.ndarray) -> np.ndarray:
    out = np.zeros(keys.ggsw_fft_len, dtype=np.complex128)
    lib().orc_scheme_switch_fft(out, np.ascontiguousarray(glev), keys.ssk_fft, C.byref(keys.params))
    return out


def pbs_generalized(keys: Keys, lwe0, lut_glwe, log_chi=0, log_v=0) -> np.ndarray:
    out = np.zeros(keys.glwe_len, dtype=np.uint64)
    lib().orc_pbs_generalized(out, np.ascontiguousarray(lwe0), np.ascontiguousarray(lut_glwe), keys.bsk_fft,
                              log_chi, log_v, C.byref(keys.params))
    return out


def generate_lut(params: Params, maps, plaintext_bits: int) -> np.ndarray:
    """UnivariateLookupTable::trivial_from_fn: GLWE with a = 0, b = generate_lut(maps)."""
    p = 1 << plaintext_bits
    table = np.array([[m(x) for x in range(p)] for m in maps], dtype=np.uint64)
    poly = np.zeros(params.glwe_n, dtype=np.uint64)
    lib().orc_generate_lut(poly, table, len(maps), params.glwe_n, plaintext_bits)
    glwe = np.zeros((params.glwe_k + 1) * params.glwe_n, dtype=np.uint64)
    glwe[params.glwe_k * params.glwe_n:] = poly
    return glwe


def keyswitch_lwe(keys: Keys, lwe1: np.ndarray) -> np.ndarray:
    out = np.zeros(keys.lwe0_len, dtype=np.uint64)
    lib().orc_keyswitch_lwe(out, np.ascontiguousarray(lwe1), keys.ksk, C.byref(keys.params))
    return out


def sample_extract(keys: Keys, glwe: np.ndarray, h: int) -> np.ndarray:
    out = np.zeros(keys.lwe1_len, dtype=np.uint64)
    lib().orc_sample_extract(out, np.ascontiguousarray(glwe), h, C.byref(keys.params))
    return out


def trace(keys: Keys, glwe: np.ndarray) -> np.ndarray:
    out = np.zeros(keys.glwe_len, dtype=np.uint64)
    lib().orc_trace(out, np.ascontiguousarray(glwe), keys.ak_fft, C.byref(keys.params))
    return out


def multiply_glwe_ggsw(keys: Keys, glwe, ggsw_fft) -> np.ndarray:
    out = np.zeros(keys.glwe_len, dtype=np.uint64)
    lib().orc_multiply_glwe_ggsw(out, np.ascontiguousarray(glwe), np.ascontiguousarray(ggsw_fft), C.byref(keys.params))
    return out


def glev_cmux(keys: Keys, d0, d1, ggsw_fft) -> np.ndarray:
    p = keys.params
    out = np.zeros(keys.glev_len, dtype=np.uint64)
    lib().orc_glev_cmux(out, np.ascontiguousarray(d0), np.ascontiguousarray(d1), np.ascontiguousarray(ggsw_fft),
                        C.byref(p), p.cbs, p.cbs)
    return out


def glwe_mul_xn(keys: Keys, glwe, n: int) -> np.ndarray:
    out = np.zeros(keys.glwe_len, dtype=np.uint64)
    lib().orc_glwe_mul_xn(out, np.ascontiguousarray(glwe), n, C.byref(keys.params))
    return out


def glwe_not(keys: Keys, glwe) -> np.ndarray:
    out = np.zeros(keys.glwe_len, dtype=np.uint64)
    lib().orc_glwe_not(out, np.ascontiguousarray(glwe), C.byref(keys.params))
    return out


# ---- serialized layouts (numpy restatement; checker for spf_b200.serialize) -------------------------
# bincode 1.3.3, fixint little-endian (Cargo.lock:244-245): a sequence is `u64 len || elements`;
# every sunscreen_tfhe entity is a single `data` sequence (sunscreen_tfhe/src/dst.rs:31-33).

def bincode_seq(a: np.ndarray) -> bytes:
    """bincode::serialize of one entity: Torus<u64> -> u64 LE, Complex<f64> -> (re, im) f64 LE."""
    a = np.ascontiguousarray(a).reshape(-1)
    body = a.astype("<c16") if a.dtype.kind == "c" else a.astype("<u8")
    return np.uint64(a.size).astype("<u8").tobytes() + body.tobytes()


def bincode_compute_key(keys: Keys) -> bytes:
    """ComputeKey field order bs_key, ks_key, ss_key, auto_key (parasol_runtime/src/crypto/keys.rs:306-318)."""
    return b"".join(bincode_seq(a) for a in (keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft))


def bincode_secret_key(keys: Keys) -> bytes:
    """SecretKey = lwe_0 || glwe_1 (keys.rs:100-105)."""
    return bincode_seq(keys.lwe0_sk) + bincode_seq(keys.glwe1_sk)


def compute_key_get_size(p: Params) -> int:
    """ComputeKey::get_size (keys.rs:326-349): all four keys counted as Complex<f64> + 4 length fields."""
    l = lib()
    n = l.orc_size_bsk_fft(C.byref(p)) + l.orc_size_ksk(C.byref(p)) + l.orc_size_ssk_fft(C.byref(p)) + \
        l.orc_size_ak_fft(C.byref(p))
    return int(n) * 16 + 4 * 8
