"""The Haskell shim cannot be compiled here (no GHC), so its FFI declarations are checked against
the C header mechanically: every `foreign import ccall` in hs/Qubism/Backend/FFI.hs must name a
function that include/qubism_sv.h declares and libqubism_sv.so exports, with the same number of
arguments and, position by position, a Haskell FFI type that marshals to the C type
(CInt <-> int, CDouble <-> double, Word64 <-> uint64_t, Ptr _ <-> pointer, IO CInt <-> int ...).
A renamed entry point, a dropped argument or an int / double mix-up fails here instead of at a
maintainer's link step."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_prototypes():
    text = open(os.path.join(ROOT, "include", "qubism_sv.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"^\s*#[^\n]*", " ", text, flags=re.M)  # preprocessor lines
    protos = {}
    for m in re.finditer(r"\b([A-Za-z_][\w\s\*]*?)\b(qb_\w+)\s*\(([^()]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or not ret:
            continue
        protos[name] = (ret, [] if args in ("", "void") else [a.strip() for a in args.split(",")])
    return protos


def _c_kind(decl: str) -> str:
    if "*" in decl or "[" in decl:
        return "ptr"
    base = re.sub(r"\b(const|unsigned)\b", "", decl).split()
    t = base[0]
    return {"int": "int", "double": "double", "uint64_t": "u64", "int64_t": "i64", "void": "void", "qb_c64": "struct"}.get(t, t)


def _hs_kind(t: str) -> str:
    t = t.strip()
    if t.startswith("Ptr") or t == "CString" or t.startswith("FinalizerPtr"):
        return "ptr"
    return {"CInt": "int", "CDouble": "double", "Word64": "u64", "CLLong": "i64", "Int64": "i64", "()": "void"}.get(t, t)


def _split_arrows(sig: str):
    parts, depth, cur = [], 0, ""
    i = 0
    while i < len(sig):
        ch = sig[i]
        depth += ch == "("
        depth -= ch == ")"
        if depth == 0 and sig.startswith("->", i):
            parts.append(cur.strip())
            cur = ""
            i += 2
            continue
        cur += ch
        i += 1
    parts.append(cur.strip())
    return parts


def test_every_foreign_import_matches_the_header_and_the_library():
    from qubism_b200 import capi
    lib = capi.lib()
    protos = _c_prototypes()
    assert len(protos) >= 40, "header parse went wrong"
    src = open(os.path.join(ROOT, "hs", "Qubism", "Backend", "FFI.hs")).read()
    imports = re.findall(r'foreign import ccall\s+(?:safe|unsafe)?\s*"(&?)(qb_\w+)"\s+\w+\s*::\s*(.+)', src)
    assert len(imports) >= 25
    for amp, name, sig in imports:
        assert name in protos, f"{name}: not declared in include/qubism_sv.h"
        assert hasattr(lib, name), f"{name}: not exported by libqubism_sv.so"
        ret, args = protos[name]
        if amp:  # address import (the finalizer): void f(T *)
            assert sig.strip().startswith("FinalizerPtr") and _c_kind(ret) == "void" and len(args) == 1 and _c_kind(args[0]) == "ptr"
            continue
        parts = _split_arrows(sig.strip())
        hs_args, hs_ret = parts[:-1], parts[-1]
        assert hs_ret.startswith("IO "), f"{name}: result must be in IO"
        assert len(hs_args) == len(args), f"{name}: {len(hs_args)} Haskell arguments, {len(args)} in C ({args})"
        for k, (h, c) in enumerate(zip(hs_args, args)):
            assert _c_kind(c) != "struct", f"{name}: argument {k} is a struct by value -- not importable, use the _ri spelling"
            assert _hs_kind(h) == _c_kind(c), f"{name}: argument {k}: {h} vs {c}"
        assert _hs_kind(hs_ret[3:]) == _c_kind(ret), f"{name}: result {hs_ret} vs {ret}"


def test_shim_modules_export_what_the_reference_modules_export():
    """hs/Qubism/{StateVec,QGate}.hs keep the export lists of the modules they replace
    (StateVec.hs:14-25, QGate.hs:14-31): the names the interpreter and the DSL import are all there.
    The reference's lists are restated here (the reference tree does not travel to the GPU box)."""
    want = {"StateVec": ["StateVec", "mkStateVec", "collapse", "measureQubit", "measure", "tensor"],
            "QGate": ["QGate", "(#>)", "gate", "ident", "pauliX", "pauliY", "pauliZ", "hadamard", "unitary", "cnot", "onJust", "onEvery",
                      "onRange", "ifBit", "controlled", "kronecker"]}
    for mod, names in want.items():
        text = open(os.path.join(ROOT, "hs", "Qubism", mod + ".hs")).read()
        head = text[text.index("module Qubism." + mod):]
        head = head[:head.index(") where")]
        for nm in names:
            assert re.search(r"[\s,(]" + re.escape(nm) + r"[\s,(]", head + " "), f"Qubism.{mod} does not export {nm}"
