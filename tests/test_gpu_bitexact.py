"""GPU vs host-emulator, BIT-EXACT.

compute-sanitizer is closed on this GPU pool, so this is the race / memory-ordering detector of the
kernels: `spf_b200/csrc/emu.cpp` executes the SAME `__host__ __device__` bodies (team_ops.cuh,
fft16.cuh) with one host thread per GPU thread and pthread barriers where the device has named
barriers; every multiply-add of the hot path is an explicit fma and both sides are compiled without
floating-point contraction (`nvcc -fmad=false`, `g++ -ffp-contract=off`), so the doubles agree bit
for bit unless the device code has a hazard the sequentially-consistent emulator does not (a missing
barrier around an in-place exchange, a tensor-memory load overtaking its store, a staged BSK row
read before its bulk copy has landed, a wrong mbarrier phase).  Any such hazard changes at least one
of the 637 x 4096 rounded coefficients of a blind rotation and is then amplified by the digit
decomposition of the next step, so equality of a whole PBS / CBS is a very sharp check.

Reference anchors of what is computed: generalized_programmable_bootstrap
(sunscreen_tfhe/src/ops/bootstrapping/programmable_bootstrapping.rs:342-410),
circuit_bootstrap_via_trace_and_scheme_switch (circuit_bootstrapping.rs:171-258), cmux (ops/fft_ops.rs:149-181).
"""
import ctypes as C

import numpy as np
import pytest

import emu_util as E

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev_keys(keys):
    return {"bsk": E.to_device_scale(keys.bsk_fft), "ak": E.to_device_scale(keys.ak_fft), "ssk": E.to_device_scale(keys.ssk_fft)}


def _emu_pbs(fn, ct, lut, dk, log_v=0):
    out = np.zeros(4096, dtype=np.uint64)
    fn(out, np.ascontiguousarray(ct), lut.ctypes.data if lut is not None else None, dk["bsk"], 637, 0, log_v, 4, 4)
    return out


def test_pbs_pair_kernel_bit_exact_vs_emulator(oracle, keys, client, evaluation, dev_keys):
    """Throughput kernel (pbs_kernel: three pair-teams per CTA, TMEM scratchpad): a 300-ciphertext batch,
    items from the first and the last wave, from pair slots 0, 1 and 2 of a CTA."""
    p = keys.params
    rng = np.random.default_rng(31)
    lut = oracle.generate_lut(p, [lambda x: (5 * x + 1) % 8], 3)
    cts = np.zeros((300, keys.lwe0_len), dtype=np.uint64)
    for m in range(300):
        oracle.lib().orc_encrypt_lwe(C.byref(client.rng), cts[m], keys.lwe0_sk, p.lwe_n, p.lwe_std, int(rng.integers(0, 8)) << 60)
    out = evaluation.programmable_bootstrap(cts, lut)
    for i in (0, 147, 148, 299):
        want = _emu_pbs(E.lib().emu_pbs, cts[i], lut, dev_keys)
        assert np.array_equal(out[i], want), f"pbs_kernel item {i}: {np.count_nonzero(out[i] != want)} coefficients differ"


def test_pbs_quad_kernel_bit_exact_vs_emulator(oracle, keys, client, evaluation, dev_keys):
    """Latency kernel (pbs_quad_kernel: four teams per ciphertext, BSK row staged by cp.async.bulk + mbarrier)."""
    p = keys.params
    lut = oracle.generate_lut(p, [lambda x: x], 3)
    cts = np.zeros((3, keys.lwe0_len), dtype=np.uint64)
    for m in range(3):
        oracle.lib().orc_encrypt_lwe(C.byref(client.rng), cts[m], keys.lwe0_sk, p.lwe_n, p.lwe_std, (2 * m + 1) << 60)
    out = evaluation.programmable_bootstrap(cts, lut)
    for i in range(3):
        want = _emu_pbs(E.lib().emu_pbs_quad, cts[i], lut, dev_keys)
        assert np.array_equal(out[i], want), f"pbs_quad_kernel item {i}"


def _emu_cbs(pbs_fn, ct, dk):
    glwe = _emu_pbs(pbs_fn, ct, None, dk, log_v=2)
    ggsw = np.zeros(16 * 1024, dtype=np.complex128)
    for lvl in range(4):
        E.lib().emu_trace_ss(glwe, None, ggsw.ctypes.data, dk["ak"], dk["ssk"], lvl, 0, 4, 4, 7, 6, 3, 15)
    return E.to_reference_scale(ggsw)


def test_cbs_bit_exact_vs_emulator(keys, client, evaluation, dev_keys):
    """Whole circuit bootstraps (blind rotation -> 4 x (mod-switch, trace) -> scheme switch): every one of the
    16 384 complex GGSW coefficients equal.  Small batch = quad kernel, 300 = pair kernel."""
    bits = [1, 0]
    cts = client.encrypt_lwe_l0_batch(bits)
    out = evaluation.circuit_bootstrap(cts)
    for i in range(2):
        want = _emu_cbs(E.lib().emu_pbs_quad, cts[i], dev_keys)
        assert np.array_equal(out[i].view(np.uint64), want.view(np.uint64)), f"CBS (quad) item {i}"
    rng = np.random.default_rng(32)
    bits = rng.integers(0, 2, 300).tolist()
    cts = client.encrypt_lwe_l0_batch(bits)
    out = evaluation.circuit_bootstrap(cts)
    for i in (1, 298):
        want = _emu_cbs(E.lib().emu_pbs, cts[i], dev_keys)
        assert np.array_equal(out[i].view(np.uint64), want.view(np.uint64)), f"CBS (pair) item {i}"


def test_cmux_kernels_bit_exact_vs_emulator(keys, client, evaluation):
    """cmux_kernel (one team per output) and cmux_wide_kernel (8 teams per output, TMA-staged inputs)."""
    a = client.encrypt_glwe_l1([0, 1, 1, 0])
    b = client.encrypt_glwe_l1([1, 0, 1, 0])
    ggsw = np.stack([client.encrypt_ggsw_l1(s) for s in (0, 1)])
    wide = evaluation.cmux(ggsw, np.stack([a] * 2), np.stack([b] * 2))          # <= 2 per SM: wide kernel
    n = 2 * 148 + 1
    bulk = evaluation.cmux(np.stack([ggsw[i % 2] for i in range(n)]), np.stack([a] * n), np.stack([b] * n))  # bulk kernel
    for i in range(2):
        gd = E.to_device_scale(ggsw[i])
        w = np.zeros(4096, dtype=np.uint64)
        E.lib().emu_cmux_wide(w, a.ctypes.data, b, gd, 4, 4)
        assert np.array_equal(wide[i], w), f"cmux_wide_kernel {i}"
        t = np.zeros(4096, dtype=np.uint64)
        E.lib().emu_cmux(t, a.ctypes.data, b, gd, 4, 4)
        assert np.array_equal(bulk[i], t), f"cmux_kernel {i}"
        assert np.array_equal(bulk[294 + i], t), f"cmux_kernel tail item {294 + i}"


def test_trace_and_scheme_switch_bit_exact_vs_emulator(keys, client, evaluation, dev_keys):
    glev = client.encrypt_glev_l1([1])
    out = evaluation.scheme_switch(glev)[0]
    ggsw = np.zeros(16 * 1024, dtype=np.complex128)
    for lvl in range(4):
        E.lib().emu_trace_ss(glev[lvl * 4096:(lvl + 1) * 4096].copy(), None, ggsw.ctypes.data, dev_keys["ak"], dev_keys["ssk"], lvl, 2,
                             4, 4, 7, 6, 3, 15)
    assert np.array_equal(out.view(np.uint64), E.to_reference_scale(ggsw).view(np.uint64))
    ct = client.encrypt_glwe_l1([1, 0, 1])
    got = evaluation.trace(ct)[0]
    want = np.zeros(4096, dtype=np.uint64)
    E.lib().emu_trace_ss(ct, want.ctypes.data, None, dev_keys["ak"], dev_keys["ssk"], 0, 1, 4, 4, 7, 6, 3, 15)
    assert np.array_equal(got, want)
