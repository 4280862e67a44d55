"""The CUDA kernel bodies (spf_b200/csrc/team_ops.cuh, fft16.cuh) executed on the HOST by the
emulator and compared with the oracle: validates the 16x16x4 FFT factorisation, the register
slot <-> bin mapping, the exchange-buffer indexing, digit extraction, rotations and the CBS
pipeline without a GPU.  (The GPU parity tests proper are in test_gpu_parity.py.)"""
import numpy as np
import pytest

import emu_util as E


def test_forward_fft_matches_oracle(oracle):
    rng = np.random.default_rng(1)
    for _ in range(3):
        p = rng.integers(0, 1 << 64, 2048, dtype=np.uint64)
        f = np.zeros(1024, dtype=np.complex128)
        E.lib().emu_poly_fft(p, f)
        ref = oracle.poly_fft(p)
        assert np.abs(f - ref).max() <= 1e-14 * np.abs(ref).max()


def test_fft_roundtrip_exact_on_digits():
    rng = np.random.default_rng(2)
    d = rng.integers(-(1 << 15), 1 << 15, 2048).astype(np.int64).astype(np.uint64)
    f = np.zeros(1024, dtype=np.complex128)
    E.lib().emu_poly_fft(d, f)
    back = np.zeros(2048, dtype=np.uint64)
    E.lib().emu_poly_ifft(f, back)
    assert np.array_equal(back, d)


def test_negacyclic_product_kat_via_kernel_fft():
    """can_negacyclic_conv (negacyclic/mod.rs:148-164) embedded in the N=2048 ring:
    (x + 2x^2 + 3x^3)^2 mod X^2048+1 has no wrap, so compare with the integer product, and
    X^2047 * X = -1 exercises the negacyclic wrap."""
    a = np.zeros(2048, dtype=np.uint64)
    a[:4] = [0, 1, 2, 3]
    f = np.zeros(1024, dtype=np.complex128)
    E.lib().emu_poly_fft(a, f)
    out = np.zeros(2048, dtype=np.uint64)
    E.lib().emu_poly_ifft(np.ascontiguousarray(f * f), out)
    want = np.zeros(2048, dtype=np.uint64)
    want[:7] = np.convolve([0, 1, 2, 3], [0, 1, 2, 3])
    assert np.array_equal(out, want)
    x1 = np.zeros(2048, dtype=np.uint64); x1[1] = 1
    xl = np.zeros(2048, dtype=np.uint64); xl[2047] = 1
    f1 = np.zeros(1024, dtype=np.complex128); fl = np.zeros(1024, dtype=np.complex128)
    E.lib().emu_poly_fft(x1, f1); E.lib().emu_poly_fft(xl, fl)
    E.lib().emu_poly_ifft(np.ascontiguousarray(f1 * fl), out)
    want[:] = 0
    want[0] = np.uint64((1 << 64) - 1)
    assert np.array_equal(out, want)


def test_f64_to_torus_matches_reference_rounding(oracle):
    """round() half away from zero, mod 2^64, saturating-cast quirk (simd/scalar.rs:26-35,75-119)."""
    vals = [0.0, 0.5, -0.5, 1.5, 2.5, -2.5, 2.0 ** 63, -(2.0 ** 63), 3 * 2.0 ** 63,
            -3 * 2.0 ** 63, 2.0 ** 64, 2.0 ** 64 + 4096, -(2.0 ** 70) - 2.0 ** 20, 1e30, -1e30, 123456789.5,
            -123456789.5, 2.0 ** 52 + 1, 2.0 ** 53 + 2, float("nan"), float("inf"), 5e-324]
    import math
    for v in vals:
        if v != v or v in (float("inf"),):
            rounded = v
        else:
            rounded = math.copysign(math.floor(abs(v) + 0.5), v) if abs(v) < 2 ** 52 else v
            if abs(v) < 0.5:
                rounded = 0.0
        want = np.zeros(1, dtype=np.uint64)
        oracle.lib().orc_mod_pow2_q_f64(want, np.array([rounded]), 1)
        assert E.lib().emu_f64_to_torus(v) == int(want[0]), v
    for x in (0, 1, -1, 32767, -32768, 2 ** 31 - 1, -(2 ** 31)):
        assert E.lib().emu_i32_to_f64(x) == float(x)


def test_cmux_matches_oracle(oracle, keys, client):
    a = client.encrypt_glwe_l1([0, 1, 1])
    b = client.encrypt_glwe_l1([1, 0, 1])
    for sel in (0, 1):
        ggsw = client.encrypt_ggsw_l1(sel)
        ref = oracle.cmux(keys, a, b, ggsw)
        out = np.zeros_like(ref)
        E.lib().emu_cmux(out, a.ctypes.data, b, E.to_device_scale(ggsw), 4, 4)
        assert oracle.torus_distance(ref, out).max() < 1e-11
        assert client.decrypt_glwe_l1(out)[:3].tolist() == ([1, 0, 1] if sel else [0, 1, 1])


def test_cmux_wide_matches_oracle(oracle, keys, client):
    """The latency-oriented 8-team CMUX (cmux_wide: carry-free digits, all 8 transforms concurrent)
    against the oracle, including the plain external product (d0 = NULL)."""
    a = client.encrypt_glwe_l1([0, 1, 1])
    b = client.encrypt_glwe_l1([1, 0, 1])
    for sel in (0, 1):
        ggsw = client.encrypt_ggsw_l1(sel)
        ref = oracle.cmux(keys, a, b, ggsw)
        out = np.zeros_like(ref)
        E.lib().emu_cmux_wide(out, a.ctypes.data, b, E.to_device_scale(ggsw), 4, 4)
        assert oracle.torus_distance(ref, out).max() < 1e-11
        assert client.decrypt_glwe_l1(out)[:3].tolist() == ([1, 0, 1] if sel else [0, 1, 1])
    ggsw = client.encrypt_ggsw_l1(1)
    ref = oracle.multiply_glwe_ggsw(keys, b, ggsw)
    out = np.zeros_like(ref)
    E.lib().emu_cmux_wide(out, None, b, E.to_device_scale(ggsw), 4, 4)
    assert oracle.torus_distance(ref, out).max() < 1e-11


def test_pbs_first_steps_match_oracle(oracle, keys, client):
    """Blind-rotation steps on identical inputs.  After ONE step nothing has been decomposed that
    carries FFT rounding noise, so kernel body and oracle agree to ~2^-34 of the torus (the f64
    ulp of a 2^88-sized IFFT output).  From the second step on, the 32-bit rounding of
    (acc*X^a - acc) sees that noise and digits flip with probability ~1/8 per coefficient: the
    ciphertexts then differ by whole BSK rows while their PHASE (b - a*s) still agrees to the
    noise level -- which is why PBS/CBS parity is stated on decryptions."""
    import ctypes as C
    bsk_dev = E.to_device_scale(keys.bsk_fft)
    ct = client.encrypt_lwe_l0(1)
    for nsteps in (1, 2):
        p = oracle.default_128()
        p.lwe_n = nsteps
        sub = np.concatenate([ct[:nsteps], ct[-1:]])
        rot = sub.copy(); rot[-1] = np.uint64((int(rot[-1]) + (1 << 62)) & ((1 << 64) - 1))
        lut = np.zeros(4096, dtype=np.uint64)
        oracle.lib().orc_cbs_lut(lut, C.byref(p))
        ref = np.zeros(4096, dtype=np.uint64)
        oracle.lib().orc_pbs_generalized(ref, rot, lut, keys.bsk_fft, 0, 2, C.byref(p))
        out = np.zeros(4096, dtype=np.uint64)
        E.lib().emu_pbs(out, sub, None, bsk_dev, nsteps, 0, 2, 4, 4)
        if nsteps == 1:
            assert oracle.torus_distance(ref, out).max() < 1e-10
        ph = oracle.torus_distance(client.decrypt_glwe_l1_raw(ref), client.decrypt_glwe_l1_raw(out))
        assert ph.max() < 2.0 ** -25, nsteps


def test_pbs_quad_matches_pair_and_oracle(oracle, keys, client):
    """The latency-mode four-team blind rotation (pbs_quad_team): one step against the oracle
    (same tolerance as the pair version), a few steps against the pair version in phase, and a full
    PBS by its multi-function LUT phase pattern."""
    import ctypes as C
    bsk_dev = E.to_device_scale(keys.bsk_fft)
    ct = client.encrypt_lwe_l0(1)
    p = oracle.default_128()
    p.lwe_n = 1
    sub = np.concatenate([ct[:1], ct[-1:]])
    rot = sub.copy(); rot[-1] = np.uint64((int(rot[-1]) + (1 << 62)) & ((1 << 64) - 1))
    lut = np.zeros(4096, dtype=np.uint64)
    oracle.lib().orc_cbs_lut(lut, C.byref(p))
    ref = np.zeros(4096, dtype=np.uint64)
    oracle.lib().orc_pbs_generalized(ref, rot, lut, keys.bsk_fft, 0, 2, C.byref(p))
    out = np.zeros(4096, dtype=np.uint64)
    E.lib().emu_pbs_quad(out, sub, None, bsk_dev, 1, 0, 2, 4, 4)
    assert oracle.torus_distance(ref, out).max() < 1e-10
    sub = np.concatenate([ct[:5], ct[-1:]])
    a, b = np.zeros(4096, dtype=np.uint64), np.zeros(4096, dtype=np.uint64)
    E.lib().emu_pbs(a, sub, None, bsk_dev, 5, 0, 2, 4, 4)
    E.lib().emu_pbs_quad(b, sub, None, bsk_dev, 5, 0, 2, 4, 4)
    assert oracle.torus_distance(client.decrypt_glwe_l1_raw(a), client.decrypt_glwe_l1_raw(b)).max() < 2.0 ** -22  # 5 steps of digit-flip noise
    glwe = np.zeros(4096, dtype=np.uint64)
    E.lib().emu_pbs_quad(glwe, ct, None, bsk_dev, 637, 0, 2, 4, 4)
    ph = client.decrypt_glwe_l1_raw(glwe)
    for i in range(4):
        want = (1 << (64 - (4 * (i + 1) + 1)))
        assert abs(int(np.int64(ph[i])) - want) < want // 8


def test_full_cbs_decrypts_like_fresh_ggsw(oracle, keys, client):
    """Whole CBS through the kernel bodies (PBS -> 4 x (pre-process, trace) -> scheme switch),
    checked as can_circuit_bootstrap_via_trace_ss does (circuit_bootstrapping.rs:777-803)."""
    bsk_dev = E.to_device_scale(keys.bsk_fft)
    ak_dev = E.to_device_scale(keys.ak_fft)
    ssk_dev = E.to_device_scale(keys.ssk_fft)
    bit = 1
    ct = client.encrypt_lwe_l0(bit)
    glwe = np.zeros(4096, dtype=np.uint64)
    E.lib().emu_pbs(glwe, ct, None, bsk_dev, 637, 0, 2, 4, 4)
    # the PBS output phase carries -/+ B^-(i+1)/2 in coefficient i (multi-function LUT)
    ph = client.decrypt_glwe_l1_raw(glwe)
    for i in range(4):
        want = (1 << (64 - (4 * (i + 1) + 1)))
        assert abs(int(np.int64(ph[i])) - want) < want // 8
    ggsw = np.zeros(16 * 1024, dtype=np.complex128)
    for lvl in range(4):
        E.lib().emu_trace_ss(glwe, None, ggsw.ctypes.data, ak_dev, ssk_dev, lvl, 0, 4, 4, 7, 6, 3, 15)
    g = E.to_reference_scale(ggsw)
    assert np.array_equal(client.ggsw_level_messages(g), client.ggsw_expected_messages(bit))


def test_scheme_switch_body_matches_oracle(oracle, keys, client):
    """Scheme switching decomposes exact integers, so FFT-domain outputs agree to f64 rounding."""
    ssk_dev = E.to_device_scale(keys.ssk_fft)
    ak_dev = E.to_device_scale(keys.ak_fft)
    glev = client.encrypt_glev_l1([1])
    ref = oracle.scheme_switch(keys, glev)
    ggsw = np.zeros(16 * 1024, dtype=np.complex128)
    for lvl in range(4):
        E.lib().emu_trace_ss(glev[lvl * 4096:(lvl + 1) * 4096].copy(), None, ggsw.ctypes.data, ak_dev, ssk_dev, lvl, 2,
                             4, 4, 7, 6, 3, 15)
    g = E.to_reference_scale(ggsw)
    assert np.abs(g - ref).max() <= 1e-12 * np.abs(ref).max()
