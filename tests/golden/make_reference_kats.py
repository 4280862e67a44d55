"""Writes tests/golden/reference_kats.json: the known-answer vectors the reference's own unit tests and doc-tests hold for
the bootstrapping path, transcribed by hand into the table below.  When the reference checkout is present
(/root/reference, read-only) every entry is VERIFIED against the cited source lines: each `evidence` regex must match
inside the cited range, so a vector cannot drift from the reference text without this script failing.  The JSON travels
to machines without the checkout; tests/test_oracle_kat.py pins the oracle on it.

usage: python tests/golden/make_reference_kats.py [--check]   (--check: verify + compare with the committed JSON, write nothing)"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/sunscreen_tfhe/src"
OUT = os.path.join(HERE, "reference_kats.json")

KATS = {
    "_comment": "Known-answer vectors transcribed from the reference's own unit tests (paths relative to /root/reference/sunscreen_tfhe/src). "
                "Written by tests/golden/make_reference_kats.py.",
    "negacyclic_conv": {"src": "math/fft/negacyclic/mod.rs:148-164", "n": 4, "x": [0, 1, 2, 3], "x_squared": [-10, -12, -8, 4]},
    "fft_roundtrip": {"src": "math/fft/negacyclic/mod.rs:130-145", "n": 8, "tol": 1e-12},
    "round_values": {"src": "math/radix.rs:178-200", "radix_log": 4, "count": 4,
                     "cases": [["0x12348FFFFFFFFFFF", "0x1235"], ["0x12347FFFFFFFFFFF", "0x1234"]]},
    "decompose": {"src": "math/radix.rs:203-247", "cases": [
        {"value": "encode(7,4)", "torus": "0x7000000000000000", "radix_log": 2, "count": 2, "digits": [-1, -2]},
        {"value": "encode(1,1)", "torus": "0x8000000000000000", "radix_log": 4, "count": 3, "digits": [0, 0, -8]}]},
    "decompose_polynomial": {"src": "math/radix.rs:250-283", "radix_log": 2, "count": 2,
                             "torus": ["0x1000000000000000", "0x2000000000000000", "0x3000000000000000", "0x4000000000000000"],
                             "digits": [[1, -2, -1, 0], [0, 1, 1, 1]]},
    "modulus_switch": {"src": "ops/ciphertext/lwe_ciphertext_ops.rs:149-163", "x": "0xDEADBEEFBEEFDEAD",
                       "cases": [[0, 0, 10, "0b1101111011"], [2, 0, 10, "0b0111101011"], [0, 3, 10, "0b1101111000"], [2, 3, 10, "0b0111101000"]]},
    "polynomial_pow_k": {"src": "ops/polynomial/mod.rs:138-160", "n": 128, "k": 33, "input": {"0": 17, "6": 19, "26": 52, "93": 45},
                         "output": {"0": 17, "70": -19, "90": 52, "125": -45}},
    "polynomial_shift_round": {"src": "ops/polynomial/mod.rs:162-170", "n": 2, "input": [0, 1, 2, 3, 4, 5, 6, 7], "output": [0, 0, 1, 1, 1, 1, 2, 2]},
    "mul_by_positive_monomial": {"src": "entities/polynomial.rs:407-520", "input": [1, 2, 3, 4], "cases": {
        "0": [1, 2, 3, 4], "1": [-4, 1, 2, 3], "2": [-3, -4, 1, 2], "3": [-2, -3, -4, 1], "4": [-1, -2, -3, -4], "5": [4, -1, -2, -3],
        "6": [3, 4, -1, -2], "7": [2, 3, 4, -1], "8": [1, 2, 3, 4]}},
    "glwe_rotate_doc": {"src": "ops/bootstrapping/blind_rotation.rs:60-78", "input": [1, 2, 3, 4, 5, 6, 7, 8], "plaintext_bits": 4,
                        "plus1": [8, 1, 2, 3, 4, 5, 6, 7], "minus1": [2, 3, 4, 5, 6, 7, 8, 15]},
    "mul_by_negative_monomial": {"src": "entities/polynomial.rs:489-604", "input": [1, 2, 3, 4], "cases": {
        "0": [1, 2, 3, 4], "1": [2, 3, 4, -1], "2": [3, 4, -1, -2], "3": [4, -1, -2, -3], "4": [-1, -2, -3, -4], "5": [-2, -3, -4, 1],
        "6": [-3, -4, 1, 2], "7": [-4, 1, 2, 3], "8": [1, 2, 3, 4]}},
}

# literal fragments of the reference text that carry each vector (regexes, matched inside the cited line range +- 8 lines)
EVIDENCE = {
    "negacyclic_conv": [r"let n = 4;", r"vec!\[-10\.0, -12\.0, -8\.0, 4\.0\]"],
    "fft_roundtrip": [r"let n = 8", r"reverse"],
    "round_values": [r"0x12348FFFFFFFFFFFu64", r"0x1235", r"0x12347FFFFFFFFFFFu64", r"0x1234", r"RadixLog\(4\)", r"RadixCount\(4\)"],
    "decompose": [r"encode\(7u64, PlaintextBits\(4\)\)", r"RadixLog\(2\)", r"RadixCount\(2\)", r"wrapping_sub\(1\)", r"wrapping_sub\(2\)",
                  r"encode\(1u64, PlaintextBits\(1\)\)", r"let radix_log = 4;", r"RadixCount\(3\)", r"wrapping_sub\(1 << \(radix_log - 1\)\)"],
    "decompose_polynomial": [r"RadixLog\(2\)", r"RadixCount\(2\)", r"\[1u64, 0u64\.wrapping_sub\(2\), 0u64\.wrapping_sub\(1\), 0\]", r"\[0, 1, 1, 1\]"],
    "modulus_switch": [r"0xDEADBEEFBEEFDEAD", r"0b1101111011", r"0b0111101011", r"0b1101111000", r"0b0111101000"],
    "polynomial_pow_k": [r"zero\(128\)", r"\[0\] = 17", r"\[6\] = 19", r"\[26\] = 52", r"\[93\] = 45", r"&polynomial, 33\)", r"70 => 0\.wrapping_sub\(&19\)",
                         r"90 => 52", r"125 => 0\.wrapping_sub\(&45\)"],
    "polynomial_shift_round": [r"\[0, 1, 2, 3, 4, 5, 6, 7, 8\]", r"&poly, 2\)", r"\[0, 0, 1, 1, 1, 1, 2, 2\]"],
    "glwe_rotate_doc": [r"\[1, 2, 3, 4, 5, 6, 7, 8\]", r"\[8, 1, 2, 3, 4, 5, 6, 7\]", r"\[2, 3, 4, 5, 6, 7, 8, 15\]"],
}


def monomial_goldens(fn_name):
    """Parse the `expected_K` polynomials of the reference's golden tests for monomial multiplication
    (entities/polynomial.rs: `Torus::from(4u64.wrapping_neg())` = -4, `Torus::from(1)` = 1, `original.clone()` = input)."""
    src = open(os.path.join(REF, "entities/polynomial.rs")).read()
    body = src[src.index("fn " + fn_name):]
    body = body[:body.index("#[test]")] if "#[test]" in body else body
    out = {}
    for m in re.finditer(r"let expected_(\d+) = (original\.clone\(\)|Polynomial::new\(&\[(.*?)\]\))", body, flags=re.S):
        if m.group(3) is None:
            out[m.group(1)] = [1, 2, 3, 4]
            continue
        vals = []
        for item in re.findall(r"Torus::from\(([^()]*(?:\(\))?)\)", m.group(3)):
            neg = "wrapping_neg" in item
            vals.append((-1 if neg else 1) * int(re.match(r"\d+", item).group(0)))
        out[m.group(1)] = vals
    return out


def verify():
    if not os.path.isdir(REF):
        print("reference checkout absent: vectors not re-verified", file=sys.stderr)
        return None
    checked = 0
    for key, pats in EVIDENCE.items():
        path, rng = KATS[key]["src"].split(":")
        lo, hi = (int(x) for x in rng.split("-"))
        lines = open(os.path.join(REF, path)).read().splitlines()
        text = "\n".join(lines[max(0, lo - 9):hi + 8])
        # whitespace-insensitive match: rustfmt wraps long literals
        flat = re.sub(r"\s+", " ", text)
        flat = re.sub(r"(?<=[0-9A-Fa-fxb])_(?=[0-9A-Fa-f])", "", flat)  # Rust digit separators
        for pat in pats:
            if not re.search(pat, flat) and not re.search(pat, flat.replace(", ", ",").replace(",", ", ")):
                raise SystemExit(f"{key}: reference text at {KATS[key]['src']} does not contain /{pat}/")
            checked += 1
    for key, fn in (("mul_by_positive_monomial", "can_multiply_by_positive_monomial_negacyclic"),
                    ("mul_by_negative_monomial", "can_multiply_by_negative_monomial_negacyclic")):
        got = monomial_goldens(fn)
        if got != KATS[key]["cases"]:
            raise SystemExit(f"{key}: golden polynomials parsed from the reference differ from the table: {got}")
        checked += len(got)
    return checked


if __name__ == "__main__":
    n = verify()
    text = json.dumps(KATS, indent=1) + "\n"
    if "--check" in sys.argv:
        cur = json.load(open(OUT))
        if cur != KATS:
            raise SystemExit("tests/golden/reference_kats.json differs from the table in make_reference_kats.py")
        print(f"reference_kats.json matches the table; {n if n is not None else 0} fragments verified against the reference text")
    else:
        open(OUT, "w").write(text)
        print(f"wrote {OUT}; {n if n is not None else 0} fragments verified against the reference text")
