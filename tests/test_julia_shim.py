"""The Julia shim (julia/WignerD.jl) cannot be executed here -- no Julia in the image.  What CAN be checked without it:
every ``ccall`` in the shim is compared, argument by argument, with the prototype in include/wignerd_b200.h (name, arity,
scalar-vs-pointer kind and width of every argument, return type), and the shim must cover the reference's whole public
API (src/WignerD.jl:9-12 exports plus the qualified wignerdjmn / wignerDjmn)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Julia ccall type -> (kind, bits)
JL = {"Cint": ("int", 32), "Int32": ("int", 32), "Int64": ("int", 64), "Float64": ("float", 64), "Cstring": ("ptr", 64)}
# C type -> (kind, bits)
C = {"int": ("int", 32), "int32_t": ("int", 32), "int64_t": ("int", 64), "double": ("float", 64), "wd_handle_t": ("ptr", 64)}


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        if ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _jl_type(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")):
        return ("ptr", 64)
    return JL[t]


def _c_type(t):
    t = t.strip()
    if "*" in t:
        return ("ptr", 64)
    t = t.replace("const", "").strip()
    return C[t]


def header_prototypes():
    hdr = open(os.path.join(ROOT, "include", "wignerd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|int64_t|const char \*)\s*(wd_[a-zA-Z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = [] if args.strip() in ("", "void") else _split_top(args)
        kinds = []
        for a in args:
            a = a.strip()
            typ = a[: a.rfind(" ")] if "*" not in a else a          # drop the parameter name of scalars
            kinds.append(_c_type(typ))
        protos[name] = (("ptr", 64) if "*" in ret else _c_type(ret), kinds)
    return protos


def shim_ccalls():
    src = open(os.path.join(ROOT, "julia", "WignerD.jl")).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(wd_[a-zA-Z0-9_]+), libwd\),\s*([A-Za-z0-9_{}]+),\s*\(", src):
        name, ret = m.group(1), m.group(2)
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        args = src[m.end(): i - 1]
        calls.append((name, _jl_type(ret), [_jl_type(a) for a in _split_top(args)]))
    return calls


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = shim_ccalls()
    assert len(calls) >= 14
    for name, ret, args in calls:
        assert name in protos, f"{name} is not declared in include/wignerd_b200.h"
        pret, pargs = protos[name]
        assert ret == pret, (name, "return", ret, pret)
        assert len(args) == len(pargs), (name, "arity", len(args), len(pargs))
        for k, (a, b) in enumerate(zip(args, pargs)):
            assert a == b, (name, f"argument {k}", a, b)


def test_shim_covers_the_reference_api():
    src = open(os.path.join(ROOT, "julia", "WignerD.jl")).read()
    assert re.search(r"^export wignerd!, wignerd, wignerD!, wignerD\s*$", src, re.M)          # src/WignerD.jl:9-12
    for fn in ("wignerdjmn", "wignerDjmn", "wignerd!", "wignerd", "wignerD!", "wignerD", "_sincos", "_offsetmatrix"):
        assert re.search(r"^(function\s+)?" + re.escape(fn) + r"\(", src, re.M), fn
    for typ in ("Equator", "NorthPole", "SouthPole", "SpecialCoLatitudes", "WignerMatrix"):
        assert re.search(r"\b(struct|abstract type)\s+" + typ + r"\b", src), typ
    used = {c[0] for c in shim_ccalls()}
    for name in ("wd_create", "wd_create_multi", "wd_destroy", "wd_last_error", "wd_d_batch", "wd_D_batch", "wd_d_matrix", "wd_D_matrix",
                 "wd_djmn_batch", "wd_Djmn_batch", "wd_rotate_sh", "wd_tables_save", "wd_tables_load"):
        assert name in used, name
