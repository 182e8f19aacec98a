"""CPU-only: the Julia `ccall` binding (julia/CoverageCUDA.jl) cannot be executed here (no Julia in the image or on
the GPU boxes), so its foreign-call signatures are checked STATICALLY against include/coverage_cuda.h: every symbol
it calls is declared, with the same number of arguments, and every argument / return type is the Julia spelling of
the C type.  A mismatch here would be memory corruption at the first call on a machine that does have Julia."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "coverage_cuda.h")
JULIA = os.path.join(ROOT, "maximumareacoverageoptimization.jl_b200", "julia", "CoverageCUDA.jl")

# C parameter type (normalised: no const, no names, single spaces) -> Julia types a ccall may spell it with
C_TO_JULIA = {
    "int": {"Cint", "Int32"},
    "int32_t": {"Int32", "Cint"},
    "int64_t": {"Int64"},
    "uint64_t": {"UInt64"},
    "double": {"Float64"},
    "cov_handle *": {"Ptr{Cvoid}"},
    "cov_multi *": {"Ptr{Cvoid}"},
    "void *": {"Ptr{Cvoid}"},
    "cov_handle **": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "cov_multi **": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "void **": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "double *": {"Ptr{Float64}", "Ref{Float64}"},
    "int64_t *": {"Ptr{Int64}", "Ref{Int64}"},
    "uint8_t *": {"Ptr{UInt8}"},
    "uint32_t *": {"Ptr{UInt32}"},
    "int *": {"Ptr{Cint}", "Ref{Cint}"},
}
C_RET_TO_JULIA = {"int": {"Cint"}, "void": {"Cvoid"}, "char *": {"Cstring"}, "int64_t": {"Int64"}, "double": {"Float64"},
                  "void *": {"Ptr{Cvoid}"}, "cov_handle *": {"Ptr{Cvoid}"}}


def _strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def _norm_c_type(param):
    """'const double *r_max' -> 'double *';  'int64_t B' -> 'int64_t'."""
    p = re.sub(r"\bconst\b", " ", param).strip()
    stars = p.count("*")
    p = p.replace("*", " ")
    words = p.split()
    if len(words) > 1 and words[-1] not in ("int", "double", "void", "char"):  # drop the parameter name
        words = words[:-1]
    return " ".join(words) + (" " + "*" * stars if stars else "")


def header_prototypes():
    text = _strip_comments(open(HEADER).read())
    protos = {}
    for m in re.finditer(r"COV_API\s+([\w\s\*]+?)\b(cov_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1), m.group(2), m.group(3)
        params = [p.strip() for p in params.replace("\n", " ").split(",")]
        if params == ["void"]:
            params = []
        protos[name] = (_norm_c_type(ret + " x").strip(), [_norm_c_type(p) for p in params])
    return protos


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    text = open(JULIA).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*LIB\),\s*(\w+(?:\{[^}]*\})?),\s*\(", text):
        start = m.end()
        depth, k = 1, start
        while depth:
            depth += {"(": 1, ")": -1}.get(text[k], 0)
            k += 1
        types = _split_top(text[start:k - 1])
        # the values handed over: everything between the type tuple and the ccall's closing parenthesis
        depth, j = 1, k
        while depth:
            depth += {"(": 1, ")": -1, "[": 1, "]": -1}.get(text[j], 0)
            j += 1
        values = _split_top(text[k:j - 1].lstrip().lstrip(","))
        calls.append((m.group(1), m.group(2), types, text.count("\n", 0, m.start()) + 1, values))
    return calls


def test_header_parses_every_declared_symbol():
    protos = header_prototypes()
    declared = set(re.findall(r"COV_API[^;(]*?\b(cov_\w+)\s*\(", _strip_comments(open(HEADER).read())))
    assert declared and declared == set(protos)
    assert protos["cov_eval_batch"] == ("int", ["cov_handle *", "double *", "int64_t", "double *", "int64_t *", "uint8_t *"])
    assert protos["cov_last_error"][0] == "char *" and protos["cov_abi_version"][1] == []


def test_julia_ccalls_match_the_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 25
    seen = set()
    for name, ret, args, line, values in calls:
        assert len(values) == len(args), f"CoverageCUDA.jl:{line}: {name}: {len(args)} argument types, {len(values)} values"
        assert name in protos, f"CoverageCUDA.jl:{line}: {name} is not declared in coverage_cuda.h"
        c_ret, c_args = protos[name]
        assert ret in C_RET_TO_JULIA[c_ret], f"CoverageCUDA.jl:{line}: {name} returns {c_ret}, ccall says {ret}"
        assert len(args) == len(c_args), f"CoverageCUDA.jl:{line}: {name} takes {len(c_args)} arguments, ccall passes {len(args)}"
        for k, (ja, ca) in enumerate(zip(args, c_args)):
            assert ja in C_TO_JULIA[ca], f"CoverageCUDA.jl:{line}: {name} argument {k + 1} is `{ca}`, ccall says {ja}"
        seen.add(name)
    # the path's entry points are all bound
    for must in ("cov_create", "cov_destroy", "cov_set_points", "cov_set_params", "cov_eval_one", "cov_eval_batch",
                 "cov_eval_batch_best", "cov_eval_batch_packed", "cov_eval_batch_ex", "cov_argmin", "cov_remove_covered",
                 "cov_add_points", "cov_mads_solve", "cov_host_alloc", "cov_host_free"):
        assert must in seen, must


def test_python_ctypes_table_matches_the_header():
    """_lib.SIGNATURES (what the tests and bench.py call through) has the header's arity for every symbol."""
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import coverage_b200 as cov
    protos = header_prototypes()
    assert set(cov._lib.SIGNATURES) == set(protos)
    for name, (res, args) in cov._lib.SIGNATURES.items():
        assert len(args) == len(protos[name][1]), name
