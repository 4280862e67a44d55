"""The FAST flavour of the CPU oracle (oracle/libspf_oracle_fast.so: -DORC_FAST, AVX2/FMA FFT, vectorised conversions
and complex MADs) against the strict build.  The fast flavour is only ever the *timed CPU baseline* of bench.py
(cpu_baseline / --impl reference); these tests make sure the thing being timed still computes the reference's
algorithm: same call graph, results equal to f64 rounding, bit-exact where the arithmetic is integer."""
import ctypes as C

import numpy as np


def test_fast_fft_matches_strict(oracle):
    rng = np.random.default_rng(5)
    s, f = oracle.lib(), oracle.lib(fast=True)
    for _ in range(4):
        p = rng.integers(0, 1 << 64, 2048, dtype=np.uint64)
        a, b = np.zeros(1024, dtype=np.complex128), np.zeros(1024, dtype=np.complex128)
        s.orc_poly_fft(p, a, 2048)
        f.orc_poly_fft(p, b, 2048)
        assert np.abs(a - b).max() <= 1e-14 * np.abs(a).max()
    # exact round trip on small digits (what the blind rotation transforms)
    d = rng.integers(-(1 << 15), 1 << 15, 2048).astype(np.int64).astype(np.uint64)
    fd = np.zeros(1024, dtype=np.complex128)
    f.orc_poly_fft(d, fd, 2048)
    back = np.zeros(2048, dtype=np.uint64)
    f.orc_poly_ifft(fd, back, 2048)
    assert np.array_equal(back, d)


def test_fast_mod_pow2_matches_strict_including_corners(oracle):
    vals = np.array([0.0, 1.0, -1.0, 2.0 ** 63, -(2.0 ** 63), 3 * 2.0 ** 63, -3 * 2.0 ** 63, 2.0 ** 64, 2.0 ** 64 + 4096,
                     -(2.0 ** 70) - 2.0 ** 20, 1e30, -1e30, 123456789.0, -123456789.0, 2.0 ** 52 + 1, 2.0 ** 53 + 2,
                     2.0 ** 62, -(2.0 ** 62), 2.0 ** 63 - 1024, -(2.0 ** 63) + 1024])
    rng = np.random.default_rng(6)
    rnd = np.round(rng.normal(0, 2.0 ** 80, 2048 - len(vals)))
    x = np.concatenate([vals, rnd])
    a, b = np.zeros(len(x), dtype=np.uint64), np.zeros(len(x), dtype=np.uint64)
    oracle.lib().orc_mod_pow2_q_f64(a, x, len(x))
    oracle.lib(fast=True).orc_mod_pow2_q_f64(b, x, len(x))
    assert np.array_equal(a, b)
    # a vector without corner values takes the vectorised path
    oracle.lib().orc_mod_pow2_q_f64(a[:2028], rnd, 2028)
    oracle.lib(fast=True).orc_mod_pow2_q_f64(b[:2028], rnd, 2028)
    assert np.array_equal(a[:2028], b[:2028])


def test_fast_integer_ops_bit_exact(oracle, keys, client):
    s, f = oracle.lib(), oracle.lib(fast=True)
    rng = np.random.default_rng(7)
    for deg in (1, 5, 2047, 2048, 2049, 4095, -1, -2047, -2048, -3000):
        p = rng.integers(0, 1 << 64, 2048, dtype=np.uint64)
        a, b = p.copy(), p.copy()
        s.orc_poly_mul_monomial(a, 2048, deg)
        f.orc_poly_mul_monomial(b, 2048, deg)
        assert np.array_equal(a, b)
    l1 = client.encrypt_lwe_l1(1)
    a, b = np.zeros(keys.lwe0_len, dtype=np.uint64), np.zeros(keys.lwe0_len, dtype=np.uint64)
    s.orc_keyswitch_lwe(a, l1, keys.ksk, C.byref(keys.params))
    f.orc_keyswitch_lwe(b, l1, keys.ksk, C.byref(keys.params))
    assert np.array_equal(a, b)


def test_fast_cbs_decrypts_and_stays_within_noise(oracle, keys, client):
    bits = [1, 0]
    cts = client.encrypt_lwe_l0_batch(bits)
    strict = oracle.circuit_bootstrap_batch(keys, cts, 2)
    fast = oracle.circuit_bootstrap_batch(keys, cts, 2, fast=True)
    for i, bit in enumerate(bits):
        assert client.decrypt_ggsw_l1(fast[i]) == bit
        assert np.array_equal(client.ggsw_level_messages(fast[i]), client.ggsw_expected_messages(bit))
        d = oracle.torus_distance(client.ggsw_phases(strict[i]), client.ggsw_phases(fast[i]))
        assert d.max() <= 2.0 ** -18  # two FFT implementations: digit-flip noise
