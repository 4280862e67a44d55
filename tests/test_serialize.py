"""Serialized key / ciphertext layouts: spf_b200.serialize (C ABI) against the
oracle's independent numpy restatement of the reference's bincode layout, plus the reference's own
malformed-input vectors (parasol_runtime/src/safe_bincode.rs:41-129).  No GPU needed except for
the last test."""
import numpy as np
import pytest

import spf_b200
from spf_b200 import serialize as S

# safe_bincode.rs:60-62 / :108-110 -- the reference's "malformed length" vector
MALFORMED = bytes([253, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0x1, 0x2, 0x3, 0x4])


def test_sizes_match_get_size(oracle):
    p = spf_b200.default_128()
    # encryption.rs:454-519: (size + 1) * 8
    assert S.ciphertext_size(S.LWE0, p) == (638 + 1) * 8
    assert S.ciphertext_size(S.LWE1, p) == (2049 + 1) * 8
    assert S.ciphertext_size(S.GLWE1, p) == (4096 + 1) * 8
    assert S.ciphertext_size(S.GLEV1, p) == (16384 + 1) * 8
    assert S.ciphertext_size(7, p) == 0  # L1Ggsw is not serialisable
    # exact ComputeKey size: 83 492 864 + 62 717 952 + 491 520 + 2 162 688 bytes + 4 length fields
    assert S.compute_key_size(p) == 83492864 + 62717952 + 491520 + 2162688 + 32
    assert S.compute_key_limit(p) == oracle.compute_key_get_size(oracle.default_128())
    assert S.compute_key_limit(p) >= S.compute_key_size(p)


def test_ciphertext_roundtrip_and_layout(oracle, client, keys):
    # can_safe_deserialize_ciphertexts (safe_bincode.rs:41-53) with real encryptions
    cases = [(S.LWE0, client.encrypt_lwe_l0(1)), (S.GLWE1, client.encrypt_glwe_l1([1])),
             (S.LWE1, np.arange(keys.lwe1_len, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)),
             (S.GLEV1, np.arange(keys.glev_len, dtype=np.uint64) ^ np.uint64(0xDEADBEEF))]
    for kind, ct in cases:
        ours = S.dump_ciphertext(kind, ct)
        assert ours == oracle.bincode_seq(ct)  # byte-identical to the restated bincode layout
        assert len(ours) == S.ciphertext_size(kind)
        back = S.load_ciphertext(kind, ours)
        assert back.dtype == np.uint64 and np.array_equal(back, np.asarray(ct).reshape(-1))
        # allow_trailing_bytes (safe_bincode.rs:20)
        assert np.array_equal(S.load_ciphertext(kind, ours + b"\x00\x01"), back)


def test_known_bytes_small_vector():
    # hand-built golden: a sequence of 638 torus values 0,1,2,... is `7e 02 00..` + LE words
    ct = np.arange(638, dtype=np.uint64)
    b = S.dump_ciphertext(S.LWE0, ct)
    assert b[:8] == bytes([0x7E, 0x02, 0, 0, 0, 0, 0, 0])
    assert b[8:16] == bytes(8) and b[16:24] == bytes([1, 0, 0, 0, 0, 0, 0, 0])
    assert b[-8:] == (637).to_bytes(8, "little")


def test_rejects_malformed_ciphertexts():
    # rejects_malformed_serialized_ciphertext (safe_bincode.rs:55-74)
    for kind in (S.LWE0, S.LWE1, S.GLWE1, S.GLEV1):
        with pytest.raises(spf_b200.SpfError):
            S.load_ciphertext(kind, MALFORMED)
    good = S.dump_ciphertext(S.LWE0, np.zeros(638, dtype=np.uint64))
    with pytest.raises(spf_b200.SpfError):
        S.load_ciphertext(S.LWE0, good[:-1])        # truncated body
    with pytest.raises(spf_b200.SpfError):
        S.load_ciphertext(S.LWE0, good[:5])         # truncated length field
    with pytest.raises(spf_b200.SpfError):
        S.load_ciphertext(S.LWE1, good)             # valid bincode, wrong entity (check_is_valid)
    with pytest.raises(spf_b200.SpfError):
        S.load_ciphertext(S.LWE0, b"")              # empty input
    with pytest.raises(spf_b200.SpfError):
        S.dump_ciphertext(S.LWE0, np.zeros(637, dtype=np.uint64))


def test_compute_key_roundtrip(oracle, keys):
    ref = oracle.bincode_compute_key(keys)
    assert len(ref) == S.compute_key_size()
    bsk, ksk, ssk, ak = S.load_compute_key(ref)
    assert np.array_equal(bsk, keys.bsk_fft) and np.array_equal(ksk, keys.ksk)
    assert np.array_equal(ssk, keys.ssk_fft) and np.array_equal(ak, keys.ak_fft)
    ours = S.dump_compute_key(keys.bsk_fft, keys.ksk, keys.ssk_fft, keys.ak_fft)
    assert ours == ref
    # field order matters: swapping two keys must be rejected by the length check
    swapped = b"".join(oracle.bincode_seq(a) for a in (keys.ksk, keys.bsk_fft, keys.ssk_fft, keys.ak_fft))
    with pytest.raises(spf_b200.SpfError):
        S.load_compute_key(swapped)


def test_rejects_malformed_keys(oracle, keys):
    # rejects_malformed_keys (safe_bincode.rs:102-129)
    with pytest.raises(spf_b200.SpfError):
        S.load_compute_key(MALFORMED)
    ref = oracle.bincode_compute_key(keys)
    with pytest.raises(spf_b200.SpfError):
        S.load_compute_key(ref[:-16])
    bad = bytearray(ref)
    bad[0] ^= 1  # first length field off by one
    with pytest.raises(spf_b200.SpfError):
        S.load_compute_key(bytes(bad))
    with pytest.raises(spf_b200.SpfError):
        S.dump_compute_key(keys.bsk_fft[:-1], keys.ksk, keys.ssk_fft, keys.ak_fft)


@pytest.mark.gpu
def test_evaluation_from_serialized_key(oracle, keys, client):
    """A ComputeKey in the reference's serialized layout loads unmodified and bootstraps correctly."""
    ev = spf_b200.Evaluation.from_serialized(oracle.bincode_compute_key(keys))
    try:
        bits = [0, 1, 1, 0]
        lwe = np.stack([S.load_ciphertext(S.LWE0, S.dump_ciphertext(S.LWE0, client.encrypt_lwe_l0(b))) for b in bits])
        ggsw = ev.circuit_bootstrap(lwe)
        for b, g in zip(bits, ggsw):
            assert client.decrypt_ggsw_l1(g) == b
    finally:
        ev.close()
