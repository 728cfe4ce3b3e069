"""CPU tests of the C-ABI boundary: the library builds, loads and exports every symbol that
include/ls_cuda.h declares; the host mirror validates arguments like the reference would.
No compute call is made here (no GPU in the build container)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol(built_lib):
    import fast_solver_lippmann_schwinger_b200 as ls
    L = ls.lib()
    declared = ls.declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    # every declared symbol has a ctypes signature in the binding (no unchecked call)
    assert sorted(L._signatures) == declared
    assert L.ls_version() >= 100
    # dynamic symbol table really is extern "C" (unmangled)
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared) <= exported


def test_header_has_no_cxx_or_torch_types():
    import re
    text = open(os.path.join(ROOT, "include", "ls_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # declarations only, comments stripped
    for bad in ("std::", "torch", "at::Tensor", "template", "class "):
        assert bad not in text
    assert 'extern "C"' in text and "ls_cdouble" in text


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fast_solver_lippmann_schwinger_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_error_paths_without_gpu(built_lib):
    import fast_solver_lippmann_schwinger_b200 as ls
    from fast_solver_lippmann_schwinger_b200._lib import lib
    L = lib()
    # argument validation happens before any CUDA call
    h = C.c_void_p()
    nu = np.zeros(4)
    g = np.zeros(16, complex)
    rc = L.ls_op2d_create(C.byref(h), 2, 2, 9, 8, nu.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), 1.0, 1, 0)
    assert rc == -1 and b"ne = 4n" in L.ls_last_error()
    rc = L.ls_op2d_create(C.byref(h), 3, 3, 6, 5, nu.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), 1.0, 0, 0)
    assert rc == -1 and b"trapezoidal needs ne = 2n-1" in L.ls_last_error()
    # valid upstream, too large for the general-size GPU path: refused loudly, before any CUDA call
    rc = L.ls_op2d_create(C.byref(h), 1000, 1000, 4000, 4000, nu.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), 1.0, 1, 0)
    assert rc == -2 and b"4096" in L.ls_last_error()
    rc = L.ls_op2d_create(C.byref(h), 2049, 2049, 4097, 4097, nu.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), 1.0, 0, 0)
    assert rc == -2
    rc = L.ls_op2d_apply(None, None, None, 0, 0)
    assert rc == -1
    assert L.ls_destroy(None) == 0
    with pytest.raises(ValueError):
        ls.FastM(np.zeros((8, 8), complex), np.zeros(5), 8, 8, 2, 2, 1.0, quadRule="Greengard_Vico")
    with pytest.raises(ValueError):
        ls.FastM(np.zeros((8, 8), complex), np.zeros(4), 8, 8, 2, 2, 1.0, quadRule="simpson")


def test_spm_argument_validation(built_lib):
    from fast_solver_lippmann_schwinger_b200._lib import lib
    L = lib()
    h = C.c_void_p()
    colptr = np.array([0, 1, 2], dtype=np.int64)          # 0-based: rejected (Julia arrays are 1-based)
    rowval = np.array([1, 2], dtype=np.int64)
    nz = np.ones(2, complex)
    rc = L.ls_spm_create(C.byref(h), 2, 2, colptr.ctypes.data_as(C.c_void_p), rowval.ctypes.data_as(C.c_void_p),
                         nz.ctypes.data_as(C.c_void_p))
    assert rc == -1 and b"1-based" in L.ls_last_error()
    colptr = np.array([1, 2, 3], dtype=np.int64)
    rowval = np.array([1, 5], dtype=np.int64)             # row out of range
    rc = L.ls_spm_create(C.byref(h), 2, 2, colptr.ctypes.data_as(C.c_void_p), rowval.ctypes.data_as(C.c_void_p),
                         nz.ctypes.data_as(C.c_void_p))
    assert rc == -1 and b"out of range" in L.ls_last_error()


def _split_top_level(s):
    """Split a comma-separated list, ignoring commas nested in (), {} or []."""
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


def test_julia_binding_matches_the_header():
    """julia/LSCuda.jl cannot be executed here (no Julia in the image): check statically that every ccall names a
    function declared in include/ls_cuda.h and passes as many arguments as the C prototype takes."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "ls_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    protos = {}
    for m in re.finditer(r"\b(ls_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_top_level(args))
    jl = open(os.path.join(root, "julia", "LSCuda.jl")).read()
    seen = 0
    for m in re.finditer(r"ccall\(\(:(\w+),\s*libls\),\s*\w+,\s*\(", jl):
        name = m.group(1)
        i, depth = m.end(), 1
        while depth:                      # the argument-type tuple
            depth += {"(": 1, ")": -1}.get(jl[i], 0)
            i += 1
        types = jl[m.end():i - 1].strip().rstrip(",")
        nargs = 0 if not types else len(_split_top_level(types))
        assert name in protos, "LSCuda.jl calls %s, which include/ls_cuda.h does not declare" % name
        assert nargs == protos[name], "%s: LSCuda.jl passes %d arguments, the header declares %d" % (name, nargs, protos[name])
        seen += 1
    assert seen >= 10
