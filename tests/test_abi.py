"""The C-ABI library loads and exports every symbol include/spf_b200.h declares (no compute
calls: this runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "spf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(spf_b200_\w+)\s*\(", hdr)))


def test_header_symbols_exported():
    import spf_b200

    lib = C.CDLL(spf_b200.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/spf_b200.h but not exported"
    assert sorted(spf_b200.ABI) == syms, "spf_b200.ABI and include/spf_b200.h disagree"


def test_sizes_match_reference_layouts():
    """Entity sizes of SURVEY.md section 8 at DEFAULT_128."""
    import spf_b200

    l = spf_b200.lib()
    p = spf_b200.default_128()
    r = C.byref(p)
    assert (p.lwe_n, p.glwe_k, p.glwe_n) == (637, 1, 2048)
    assert [(x.radix_log, x.count) for x in (p.cbs, p.pbs, p.ks, p.ss, p.tr)] == [(4, 4), (16, 2), (2, 6), (3, 15), (7, 6)]
    assert l.spf_b200_len_lwe_l0(r) == 638
    assert l.spf_b200_len_lwe_l1(r) == 2049
    assert l.spf_b200_len_glwe_l1(r) == 4096
    assert l.spf_b200_len_glev_l1(r) == 16384
    assert l.spf_b200_len_ggsw_l1(r) == 16384
    assert l.spf_b200_len_bsk(r) * 16 == 83492864
    assert l.spf_b200_len_ksk(r) * 8 == 62717952
    assert l.spf_b200_len_ak(r) * 16 == 2162688
    assert l.spf_b200_len_ssk(r) * 16 == 491520


def test_sizes_agree_with_oracle(oracle):
    import spf_b200

    l = spf_b200.lib()
    p = spf_b200.default_128()
    op = oracle.default_128()
    ol = oracle.lib()
    assert l.spf_b200_len_bsk(C.byref(p)) == ol.orc_size_bsk_fft(C.byref(op))
    assert l.spf_b200_len_ksk(C.byref(p)) == ol.orc_size_ksk(C.byref(op))
    assert l.spf_b200_len_ak(C.byref(p)) == ol.orc_size_ak_fft(C.byref(op))
    assert l.spf_b200_len_ssk(C.byref(p)) == ol.orc_size_ssk_fft(C.byref(op))


def test_create_fails_loudly_without_gpu_or_with_bad_args(oracle):
    """No CPU fallback: without a CUDA device create() must fail; bad params fail everywhere."""
    import numpy as np
    import spf_b200

    l = spf_b200.lib()
    p = spf_b200.default_128()
    p.glwe_n = 1024
    h = C.c_void_p()
    rc = l.spf_b200_create(C.byref(p), None, 0, None, 0, None, 0, None, 0, 0, C.byref(h))
    assert rc == -2 and b"specialised" in l.spf_b200_last_error(None)
    p = spf_b200.default_128()
    dummy = np.zeros(8, dtype=np.uint64)
    rc = l.spf_b200_create(C.byref(p), dummy.ctypes.data, 8, dummy.ctypes.data, 8, dummy.ctypes.data, 8,
                           dummy.ctypes.data, 8, 0, C.byref(h))
    assert rc == -1 and b"key length" in l.spf_b200_last_error(None)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        n = [l.spf_b200_len_bsk(C.byref(p)), l.spf_b200_len_ksk(C.byref(p)), l.spf_b200_len_ssk(C.byref(p)),
             l.spf_b200_len_ak(C.byref(p))]
        rc = l.spf_b200_create(C.byref(p), dummy.ctypes.data, n[0], dummy.ctypes.data, n[1], dummy.ctypes.data, n[2],
                               dummy.ctypes.data, n[3], 0, C.byref(h))
        assert rc == -3 and b"no CPU fallback" in l.spf_b200_last_error(None)


def test_plain_c_consumer(tmp_path):
    """The boundary is a C ABI: a plain C11 program includes include/spf_b200.h, links libspf_b200.so and uses the
    host-only entry points (MUX-circuit generator, graph planner, error reporting) without Python or a GPU."""
    import shutil
    import subprocess

    import spf_b200

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    lib_dir = os.path.dirname(spf_b200.LIB_PATH)
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call([cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", lib_dir,
                           "-l:" + os.path.basename(spf_b200.LIB_PATH), "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "c abi ok: 3228 multiplexers" in out.stdout


def test_cpp_host_mirror(tmp_path):
    """include/spf_b200.hpp (the C++ host side above the C ABI: Evaluation / FheCircuit / CircuitProcessor /
    MuxCircuit with the reference's names) compiles with -Wall -Wextra -Werror, links, and its host-only parts run."""
    import shutil
    import subprocess

    import spf_b200

    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    lib_dir = os.path.dirname(spf_b200.LIB_PATH)
    exe = str(tmp_path / "hpp_smoke")
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "hpp_smoke.cpp"), "-o", exe, "-L", lib_dir,
                           "-l:" + os.path.basename(spf_b200.LIB_PATH), "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "hpp ok: 2 + 7 = 9" in out.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_gpu_path(tmp_path):
    """The same program with the argument "gpu": Evaluation::circuit_bootstrap / cmux, CircuitProcessor::
    run_graph_blocking and a CompiledGraph through the C++ wrapper on cuda:0 (all-zero keys: the calls, not the
    cryptography)."""
    import shutil
    import subprocess

    import spf_b200

    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    lib_dir = os.path.dirname(spf_b200.LIB_PATH)
    exe = str(tmp_path / "hpp_smoke")
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "hpp_smoke.cpp"),
                           "-o", exe, "-L", lib_dir, "-l:" + os.path.basename(spf_b200.LIB_PATH), "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "gpu path ok" in out.stdout
