"""The oracle against every known-answer test the reference holds for the hot path
(first row).  Vectors live in tests/golden/reference_kats.json with the
reference file:line each was transcribed from."""
import ctypes as C
import json
import os

import numpy as np
import pytest

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
U64 = (1 << 64) - 1


def u(x):
    return np.uint64(int(x) & U64)


def arr(xs):
    return np.array([int(x) & U64 for x in xs], dtype=np.uint64)


def test_negacyclic_conv_exact(oracle):
    k = KATS["negacyclic_conv"]
    l = oracle.lib()
    x = np.array(k["x"], dtype=np.float64)
    y = np.zeros(k["n"] // 2, dtype=np.complex128)
    l.orc_fft_forward(x, y, k["n"])
    out = np.zeros(k["n"])
    l.orc_fft_reverse(np.ascontiguousarray(y * y), out, k["n"])
    assert out.tolist() == [float(v) for v in k["x_squared"]]


def test_fft_roundtrip(oracle):
    k = KATS["fft_roundtrip"]
    l = oracle.lib()
    x = np.arange(k["n"], dtype=np.float64)
    y = np.zeros(k["n"] // 2, dtype=np.complex128)
    out = np.zeros(k["n"])
    l.orc_fft_forward(x, y, k["n"])
    l.orc_fft_reverse(y, out, k["n"])
    assert np.abs(out - x).max() < k["tol"]


@pytest.mark.parametrize("n", [2, 4, 16, 256, 1024, 2048, 4096])
def test_fft_matches_numpy(oracle, n):
    rng = np.random.default_rng(n)
    p = rng.integers(0, 1 << 64, n, dtype=np.uint64)
    f = oracle.poly_fft(p)
    x = p.astype(np.int64).astype(np.float64)
    j = np.arange(n // 2)
    ref = np.fft.fft((x[: n // 2] + 1j * x[n // 2:]) * np.exp(2j * np.pi * j / (2 * n)))
    assert np.abs(f - ref).max() <= 1e-13 * np.abs(ref).max()


def test_round_values(oracle):
    k = KATS["round_values"]
    r = oracle.Radix(k["radix_log"], k["count"])
    for x, want in k["cases"]:
        assert oracle.lib().orc_radix_round(int(x, 16), r) == int(want, 16)


def _digits(oracle, torus, radix_log, count):
    l = oracle.lib()
    st = arr([l.orc_radix_round(int(t), oracle.Radix(radix_log, count)) for t in torus])
    out = []
    for _ in range(count):
        d = np.zeros_like(st)
        l.orc_next_decomp(st, d, len(st), radix_log)
        out.append(d.astype(np.int64).tolist())
    return out


def test_decompose(oracle):
    for c in KATS["decompose"]["cases"]:
        got = _digits(oracle, [int(c["torus"], 16)], c["radix_log"], c["count"])
        assert [g[0] for g in got] == c["digits"]


def test_decompose_polynomial(oracle):
    k = KATS["decompose_polynomial"]
    got = _digits(oracle, [int(t, 16) for t in k["torus"]], k["radix_log"], k["count"])
    assert got == k["digits"]


def test_decompose_recompose(oracle):
    """math/radix.rs:286-394: sum_j digit_j * q/B^(j+1) == value rounded to l*logB bits."""
    rng = np.random.default_rng(7)
    for radix_log, count in [(16, 2), (4, 4), (2, 6), (7, 6), (3, 15)]:
        x = rng.integers(0, 1 << 64, 64, dtype=np.uint64)
        digs = _digits(oracle, x, radix_log, count)
        shift = 64 - radix_log * count
        rec = np.zeros(64, dtype=np.uint64)
        for t, d in enumerate(digs):
            rec += arr(d) << np.uint64(shift + radix_log * t)
        rounded = ((x >> np.uint64(shift)) + ((x >> np.uint64(shift - 1)) & np.uint64(1))) << np.uint64(shift)
        assert np.array_equal(rec, rounded)


def test_modulus_switch(oracle):
    k = KATS["modulus_switch"]
    for log_chi, log_v, log_mod, want in k["cases"]:
        assert oracle.lib().orc_modulus_switch(int(k["x"], 16), log_chi, log_v, log_mod) == int(want, 2)


def test_polynomial_pow_k(oracle):
    k = KATS["polynomial_pow_k"]
    p = np.zeros(k["n"], dtype=np.uint64)
    for i, v in k["input"].items():
        p[int(i)] = v
    out = np.zeros_like(p)
    oracle.lib().orc_poly_pow_k(out, p, k["n"], k["k"])
    want = np.zeros_like(p)
    for i, v in k["output"].items():
        want[int(i)] = u(v)
    assert np.array_equal(out, want)


def test_polynomial_shift_round(oracle):
    k = KATS["polynomial_shift_round"]
    x = arr(k["input"])
    y = np.zeros_like(x)
    oracle.lib().orc_shr_round(y, x, len(x), k["n"])
    assert y.tolist() == k["output"]


@pytest.mark.parametrize("name,sign", [("mul_by_positive_monomial", 1), ("mul_by_negative_monomial", -1)])
def test_monomial_rotation_goldens(oracle, name, sign):
    k = KATS[name]
    for deg, want in k["cases"].items():
        p = arr(k["input"])
        oracle.lib().orc_poly_mul_monomial(p, len(p), sign * int(deg))
        assert np.array_equal(p, arr(want)), (name, deg)


def test_glwe_rotate_doc_kat(oracle):
    """rotate_glwe_monomial_negacyclic's doc-test (ops/bootstrapping/blind_rotation.rs:60-78): message [1..8] at 4
    plaintext bits rotated by +1 / -1 decodes to [8,1,..,7] / [2,..,8,15] (the wrapped element is negated: -1 = 15 mod 16).
    The rotation is exactly what the blind rotation applies to the accumulator (orc_poly_mul_monomial on every GLWE
    polynomial), checked here on the plaintext side of a trivial encryption."""
    k = KATS["glwe_rotate_doc"]
    pb = k["plaintext_bits"]
    msg = np.array(k["input"], dtype=np.uint64) << np.uint64(64 - pb)
    import oracle as O
    for deg, want in ((1, k["plus1"]), (-1, k["minus1"])):
        p = msg.copy()
        oracle.lib().orc_poly_mul_monomial(p, len(p), deg)
        assert O.decode(p, pb).tolist() == want


def test_lut_matches_formula(oracle):
    """generate_lut + the multi-function layout (programmable_bootstrapping.rs:129-185): single
    identity map, p=8 on N=64: stride 8, first half-stride negated and rotated to the end."""
    p = oracle.small_params(64, 8)
    glwe = oracle.generate_lut(p, [lambda x: x], 3)
    lut = glwe[64:]
    delta = 61
    want = np.zeros(64, dtype=np.uint64)
    for j in range(8):
        want[j * 8:(j + 1) * 8] = np.uint64(j << delta)
    want[:4] = np.uint64(0) - want[:4]
    want = np.roll(want, -4)
    assert np.array_equal(lut, want)
    assert not glwe[:64].any()


def test_mod_pow2_matches_reference_semantics(oracle):
    """vector_mod_pow2_q_f64 (simd/scalar.rs:75-119) incl. the saturating-cast corner."""
    vals = np.array([0.0, 1.0, -1.0, 2.0 ** 63, -(2.0 ** 63), 3 * 2.0 ** 63, -3 * 2.0 ** 63, 2.0 ** 64, 2.0 ** 64 + 4096,
                     -(2.0 ** 70) - 2.0 ** 20, 12345.0, -12345.0], dtype=np.float64)
    out = np.zeros(len(vals), dtype=np.uint64)
    oracle.lib().orc_mod_pow2_q_f64(out, vals, len(vals))
    want = [0, 1, U64, 1 << 63, (1 << 63) - 1, 1 << 63, (1 << 63) - 1, 0, 4096, (-(1 << 20)) & U64, 12345, (-12345) & U64]
    assert out.tolist() == want
