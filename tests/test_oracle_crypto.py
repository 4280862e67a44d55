"""The oracle's crypto path pinned the way the reference pins it: by decryption
('every crypto test asserts on decrypted plaintext')."""
import numpy as np
import pytest


def test_cbs_small_params_all_levels(oracle, small_keys):
    """can_circuit_bootstrap_via_trace_ss (circuit_bootstrapping.rs:721-805) on a toy ring."""
    c = oracle.Client(small_keys)
    for bit in (0, 1):
        g = oracle.circuit_bootstrap(small_keys, c.encrypt_lwe_l0(bit))
        assert np.array_equal(c.ggsw_level_messages(g), c.ggsw_expected_messages(bit))


def test_cbs_default_128(oracle, keys, client):
    """can_circuit_bootstrap (parasol_runtime/src/crypto/evaluation.rs:277-300) at DEFAULT_128."""
    for bit in (0, 1):
        g = oracle.circuit_bootstrap(keys, client.encrypt_lwe_l0(bit))
        assert np.array_equal(client.ggsw_level_messages(g), client.ggsw_expected_messages(bit))
        assert client.decrypt_ggsw_l1(g) == bit


def test_cmux_truth_table(oracle, keys, client):
    """can_cmux (evaluation.rs:318-352): sel ? b : a."""
    a = client.encrypt_glwe_l1([0, 1, 1, 0])
    b = client.encrypt_glwe_l1([1, 0, 1, 0])
    for sel in (0, 1):
        out = oracle.cmux(keys, a, b, client.encrypt_ggsw_l1(sel))
        assert client.decrypt_glwe_l1(out)[:4].tolist() == ([1, 0, 1, 0] if sel else [0, 1, 1, 0])


def test_sample_extract_and_keyswitch(oracle, keys, client):
    """can_sample_extract / can_lwe_keyswitch (evaluation.rs:302-316,354-380)."""
    bits = [1, 0, 1, 1, 0]
    g = client.encrypt_glwe_l1(bits)
    for h, want in enumerate(bits):
        l1 = oracle.sample_extract(keys, g, h)
        assert client.decrypt_lwe_l1(l1) == want
        assert client.decrypt_lwe_l0(oracle.keyswitch_lwe(keys, l1)) == want


def test_trace_keeps_constant_term_times_n(oracle, keys, client):
    """can_trace (ops/automorphisms/mod.rs:100-136): trace zeroes all but coefficient 0 and
    multiplies it by N; with the message pre-shifted by log2 N the constant term survives."""
    n = keys.params.glwe_n
    msg = np.zeros(n, dtype=np.uint64)
    msg[:8] = np.arange(1, 9, dtype=np.uint64) << np.uint64(64 - 4 - 11)  # 4-bit messages / N
    import ctypes as C
    ct = np.zeros(keys.glwe_len, dtype=np.uint64)
    oracle.lib().orc_encrypt_glwe(C.byref(client.rng), ct, msg, keys.glwe1_sk, C.byref(keys.params))
    out = oracle.trace(keys, ct)
    dec = client.decrypt_glwe_l1(out, 4)
    assert dec[0] == 1 and not dec[1:].any()


def test_pbs_univariate_small(oracle, small_keys):
    """PBS identity and (x+3)%8 maps at 3 plaintext bits + 1 padding bit
    (programmable_bootstrapping.rs:709-789), toy ring, every message."""
    import ctypes as C
    k = small_keys
    p = k.params
    c = oracle.Client(k)
    for fn in (lambda x: x, lambda x: (x + 3) % 8):
        lut = oracle.generate_lut(p, [fn], 3)  # LUT over 3 bits; the input carries one extra padding bit
        for m in range(8):
            ct = np.zeros(k.lwe0_len, dtype=np.uint64)
            oracle.lib().orc_encrypt_lwe(C.byref(c.rng), ct, k.lwe0_sk, p.lwe_n, p.lwe_std, m << 60)
            out = oracle.pbs_generalized(k, ct, lut)
            got = int(oracle.decode(c.decrypt_glwe_l1_raw(out)[:1], 3)[0])
            assert got == fn(m), (m, got)


def test_not_xor_mul_xn(oracle, keys, client):
    """can_not / can_xor / can_mul_xn (evaluation.rs:382-470)."""
    a = client.encrypt_glwe_l1([1, 0, 1, 0])
    b = client.encrypt_glwe_l1([1, 1, 0, 0])
    assert client.decrypt_glwe_l1(oracle.glwe_not(keys, a))[:4].tolist() == [0, 0, 1, 0]
    x = np.zeros_like(a)
    import ctypes as C
    oracle.lib().orc_glwe_add(x, a, b, C.byref(keys.params))
    assert client.decrypt_glwe_l1(x)[:4].tolist() == [0, 1, 1, 0]
    r = oracle.glwe_mul_xn(keys, a, 2)
    assert client.decrypt_glwe_l1(r)[:6].tolist() == [0, 0, 1, 0, 1, 0]


def test_scheme_switch_matches_fresh_ggsw(oracle, keys, client):
    """can_scheme_switch (evaluation.rs:472-507): GLEV(bit) -> GGSW decrypting like a fresh one."""
    for bit in (0, 1):
        g = oracle.scheme_switch(keys, client.encrypt_glev_l1([bit]))
        assert np.array_equal(client.ggsw_level_messages(g), client.ggsw_expected_messages(bit))


def _negacyclic_mul_ref(a, u):
    """Exact product in Z_2^64[X]/(X^n + 1) with Python integers (the definition polynomial_external_mad implements)."""
    n = len(a)
    out = [0] * n
    for i, ui in enumerate(int(x) for x in u):
        if ui == 0:
            continue
        for j, aj in enumerate(int(x) for x in a):
            k = i + j
            if k < n:
                out[k] += aj * ui
            else:
                out[k - n] -= aj * ui
    return np.array([x % (1 << 64) for x in out], dtype=np.uint64)


def test_rlwe_encrypt_public_matches_definition(oracle, small_keys):
    """rlwe_encrypt_public_impl (rlwe_encryption.rs:125-160): ct = (p0 u + e0, p1 u + e1 + m), checked against an
    independent big-integer negacyclic product on the toy ring, with a non-binary multiplier too."""
    p = small_keys.params
    n = p.glwe_n
    rng = np.random.default_rng(11)
    pk = rng.integers(0, 1 << 64, 2 * n, dtype=np.uint64)
    m, e0, e1 = (rng.integers(0, 1 << 64, n, dtype=np.uint64) for _ in range(3))
    for u in (rng.integers(0, 2, n, dtype=np.uint64), rng.integers(0, 1 << 64, n, dtype=np.uint64)):
        ct = oracle.rlwe_encrypt_public(p, m, pk, u, e0, e1)
        assert np.array_equal(ct[:n], _negacyclic_mul_ref(pk[:n], u) + e0)
        assert np.array_equal(ct[n:], _negacyclic_mul_ref(pk[n:], u) + e1 + m)


def test_rlwe_public_key_encrypts_zero(oracle, keys, client):
    """rlwe_public_key_encrypts_zero (rlwe_encryption.rs:170-186) at the L1 parameters of DEFAULT_128."""
    c = oracle.Client(keys, seed=0xA11CE)
    for _ in range(3):
        pk = c.generate_public_key()
        assert not c.decrypt_glwe_l1(pk).any()


def test_can_rlwe_public_key_encrypt(oracle, keys):
    """can_rlwe_public_key_encrypt (rlwe_encryption.rs:188-211): msg_i = i mod 2 over the whole polynomial."""
    c = oracle.Client(keys, seed=0xB0B)
    n = keys.params.glwe_n
    msg = np.arange(n, dtype=np.uint64) % 2
    for _ in range(3):
        pk = c.generate_public_key()
        ct = c.encrypt_rlwe_l1(msg, pk)
        assert np.array_equal(c.decrypt_glwe_l1(ct), msg)
