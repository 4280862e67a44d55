"""Measured GPU-vs-oracle distances of the multi-step ops (PBS, trace, CBS), as north_star asks: "output
ciphertexts must match the reference's own f64-FFT implementation to within a STATED max torus-coefficient
difference".  Ciphertext bytes of a multi-step op are not comparable between two FFT implementations (DESIGN.md
section 5: digit flips after the first blind-rotation step), so the distance is stated on the PHASE b - a.s of every
output coefficient -- the quantity decryption sees.  The bars below are 4x the maxima measured on the B200
(recorded by this test in gpurun_out/measured_distances.json and copied to profiles/r2_measured_distances.json).

Reference anchors: programmable_bootstrapping.rs:709-789 (PBS test), circuit_bootstrapping.rs:721-805 (CBS test),
ops/automorphisms/mod.rs:53-85 (trace)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# 4 x the maxima measured on the B200 (profiles/r2_measured_distances.json: PBS 9.5e-7 = 2^-20.0, trace 4.8e-8 = 2^-24.3,
# CBS 6.0e-7 = 2^-20.7); fractions of the torus.  The distances are the inherent 32-bit rounding noise of the blind
# rotation (2^-33 per coefficient and step, x sqrt(N/2) x sqrt(637) ~ 2^-23.5 rms) seen through different digit
# choices, not FFT error: the emulator, which rounds like the GPU bit for bit, sits at the same distance from the oracle.
BAR_PBS = 3.8e-6
BAR_TRACE = 1.9e-7
BAR_CBS = 2.4e-6


def _record(key, value):
    path = os.path.join(ROOT, "gpurun_out", "measured_distances.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[key] = value
        json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def test_pbs_phase_distance_measured(oracle, keys, client, evaluation):
    """32 programmable bootstraps (identity LUT, 3 plaintext bits), every output coefficient's phase against
    the oracle's; also the 150-ciphertext pair-kernel path on 8 of them."""
    p = keys.params
    lut = oracle.generate_lut(p, [lambda x: x], 3)
    rng = np.random.default_rng(41)
    msgs = rng.integers(0, 8, 150)
    cts = np.zeros((150, keys.lwe0_len), dtype=np.uint64)
    for m in range(150):
        oracle.lib().orc_encrypt_lwe(C.byref(client.rng), cts[m], keys.lwe0_sk, p.lwe_n, p.lwe_std, int(msgs[m]) << 60)
    quad = evaluation.programmable_bootstrap(cts[:32], lut)   # latency kernel
    pair = evaluation.programmable_bootstrap(cts, lut)        # throughput kernel
    worst = {"quad": 0.0, "pair": 0.0}
    for i in range(32):
        ref = client.decrypt_glwe_l1_raw(oracle.pbs_generalized(keys, cts[i], lut))
        assert int(oracle.decode(ref[:1], 3)[0]) == msgs[i]
        for name, out in (("quad", quad), ("pair", pair)):
            if name == "pair" and i >= 8:
                continue
            ph = client.decrypt_glwe_l1_raw(out[i])
            assert int(oracle.decode(ph[:1], 3)[0]) == msgs[i]
            worst[name] = max(worst[name], float(oracle.torus_distance(ref, ph).max()))
    _record("pbs_phase_max", worst)
    assert max(worst.values()) <= BAR_PBS, worst


def test_trace_phase_distance_measured(oracle, keys, client, evaluation):
    n = keys.params.glwe_n
    rng = np.random.default_rng(42)
    worst = 0.0
    cts = []
    for _ in range(8):
        msg = np.zeros(n, dtype=np.uint64)
        msg[:8] = rng.integers(0, 16, 8).astype(np.uint64) << np.uint64(64 - 4 - 11)
        ct = np.zeros(keys.glwe_len, dtype=np.uint64)
        oracle.lib().orc_encrypt_glwe(C.byref(client.rng), ct, msg, keys.glwe1_sk, C.byref(keys.params))
        cts.append(ct)
    out = evaluation.trace(np.stack(cts))
    for ct, o in zip(cts, out):
        ref = oracle.trace(keys, ct)
        worst = max(worst, float(oracle.torus_distance(client.decrypt_glwe_l1_raw(ref), client.decrypt_glwe_l1_raw(o)).max()))
    _record("trace_phase_max", worst)
    assert worst <= BAR_TRACE, worst


def test_cbs_phase_distance_measured(oracle, keys, client, evaluation):
    """16 circuit bootstraps: the phase of every coefficient of every (row, level) GLWE of the output GGSW against
    the oracle's CBS of the same input (circuit_bootstrapping.rs:721-805 decrypts the same 8 GLWEs)."""
    bits = [0, 1] * 8
    cts = client.encrypt_lwe_l0_batch(bits)
    out = evaluation.circuit_bootstrap(cts)
    worst = np.zeros((2, 4))
    for i, bit in enumerate(bits):
        ref = oracle.circuit_bootstrap(keys, cts[i])
        assert client.decrypt_ggsw_l1(ref) == bit == client.decrypt_ggsw_l1(out[i])
        d = oracle.torus_distance(client.ggsw_phases(ref), client.ggsw_phases(out[i]))
        worst = np.maximum(worst, d.max(axis=2))
    _record("cbs_phase_max_by_row_level", worst.tolist())
    _record("cbs_phase_max", float(worst.max()))
    assert worst.max() <= BAR_CBS, worst


def test_full_config3_batch_every_wave_decrypts(oracle, keys, client, evaluation):
    """BASELINE config 3 at its full size: 4096 independent LWE inputs in ONE call.  512 outputs spread over every
    PBS wave (444 ciphertexts per wave: 9 full waves + the 100-ciphertext tail) and every pair slot are decrypted."""
    rng = np.random.default_rng(43)
    B = 4096
    bits = rng.integers(0, 2, B)
    from bench import encrypt_lwe0_numpy

    cts = encrypt_lwe0_numpy(keys.lwe0_sk, bits, keys.params.lwe_std, 43)
    out = evaluation.circuit_bootstrap(cts)
    idx = sorted(set(np.linspace(0, B - 1, 500).astype(int).tolist()) | {443, 444, 887, 888, 3995, 3996, 4094, 4095, 1, 147, 148, 295, 296})
    bad = [i for i in idx if client.decrypt_ggsw_l1(out[i]) != bits[i]]
    assert not bad, bad
    _record("config3_full_batch_checked", len(idx))
