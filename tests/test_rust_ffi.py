"""The Rust FFI crate (rust/spf_b200_sys) ships as source -- there is no Rust toolchain in this image -- so this test is
what keeps it honest: the extern block must declare every function of include/spf_b200.h with the same number of
arguments and compatible types, the generated file must be up to date, and the safe layer (rust/spf_b200_runtime) may only
call functions the sys crate declares."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def rust_functions():
    src = open(os.path.join(ROOT, "rust", "spf_b200_sys", "src", "lib.rs")).read()
    fns = {}
    for m in re.finditer(r"pub fn (spf_b200_\w+)\(([^)]*)\)\s*(->\s*([^;]+))?;", src):
        args = [a.strip() for a in m.group(2).split(",") if a.strip()]
        fns[m.group(1)] = ([a.split(":", 1)[1].strip() for a in args], (m.group(4) or "()").strip())
    return fns


def test_generated_file_is_up_to_date():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_every_header_function_is_declared_with_matching_signature():
    import gen_rust_sys as G

    hdr = open(G.HDR).read()
    c_fns = {name: (ret, params) for name, ret, params in G.parse_functions(hdr)}
    rs = rust_functions()
    # the same symbol set as the header (tests/test_abi.py checks header == exported symbols == python ABI)
    from test_abi import declared_symbols

    assert sorted(c_fns) == declared_symbols() == sorted(rs)
    for name, (ret, params) in c_fns.items():
        r_args, r_ret = rs[name]
        assert len(r_args) == len(params), name
        for (pname, ctype), rtype in zip(params, r_args):
            assert rtype == G.rust_type(ctype), (name, pname, ctype, rtype)
            # pointer constness survives the translation
            if "*" in ctype and not ctype.strip() in G.FNPTR:
                assert rtype.startswith("*const") == ctype.strip().startswith("const") or "*const *" in rtype, (name, pname)
        assert r_ret == ("()" if ret == "void" else G.rust_type(ret)), name


def test_struct_layouts_match_the_c_header():
    """repr(C) field order of the three structs that cross the boundary (sizes are pinned by tests/test_graph_plan.py and
    tests/test_abi.py on the C side)."""
    src = open(os.path.join(ROOT, "rust", "spf_b200_sys", "src", "lib.rs")).read()
    def fields(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, src, flags=re.S).group(1)
        return [(f.strip().split(":")[0].replace("pub ", "").replace("r#", "").strip(), f.split(":")[1].strip()) for f in body.split(",") if ":" in f]
    assert fields("spf_radix") == [("radix_log", "u32"), ("count", "u32")]
    assert [f for f, _ in fields("spf_params")] == ["lwe_n", "lwe_std", "glwe_k", "glwe_n", "glwe_std", "cbs", "pbs", "ks", "pfks", "ss", "tr"]
    assert fields("spf_node") == [("op", "u32"), ("arg", "u32"), ("in", "[i32; 3]"), ("io", "*mut c_void")]
    assert fields("spf_mux_node") == [("op", "u32"), ("arg", "u32"), ("sel", "i32"), ("low", "i32"), ("high", "i32")]
    # enum constants carry the header's values
    hdr = open(os.path.join(ROOT, "include", "spf_b200.h")).read()
    for const, val in (("SPF_E_GRAPH", -4), ("SPF_OP_MUL_XN", 29), ("SPF_OP_CIRCUIT_BOOTSTRAP", 17), ("SPF_MUX_BITSHIFT", 9)):
        assert re.search(r"pub const %s: \w+ = %d;" % (const, val), src), const
        assert const in hdr


def test_safe_layer_only_calls_declared_functions():
    rs = rust_functions()
    src = open(os.path.join(ROOT, "rust", "spf_b200_runtime", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(spf_b200_\w+)", src)) - {"spf_b200_ctx", "spf_b200_graph"}  # the two opaque types
    assert used and used <= set(rs), used - set(rs)
    # argument counts of the calls (a cheap arity check in place of rustc)
    for m in re.finditer(r"sys::(spf_b200_\w+)\(", src):
        if m.group(1) in ("spf_b200_ctx", "spf_b200_graph"):
            continue
        name, i, depth, args, cur = m.group(1), m.end(), 1, 0, ""
        while depth:
            ch = src[i]
            depth += ch in "([{"
            depth -= ch in ")]}"
            if depth == 1 and ch == ",":
                args += 1
            elif depth >= 1:
                cur += ch
            i += 1
        n_args = 0 if not cur.strip() else args + 1
        assert n_args == len(rs[name][0]), (name, n_args, len(rs[name][0]))
    used_consts = set(re.findall(r"sys::(SPF_\w+)", src))
    sys_src = open(os.path.join(ROOT, "rust", "spf_b200_sys", "src", "lib.rs")).read()
    for c in used_consts:
        assert re.search(r"pub const %s:" % c, sys_src), c
