"""The C-ABI library loads without a GPU, exports every symbol include/rrtqx_b200.h declares,
and refuses to work (loudly) without a CUDA device -- there is no CPU fallback."""
import ctypes as C
import os
import re
import subprocess

import pytest

from rrtqx_3d_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rrtqx_b200.h")


def header_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"RRTQX_API[^;(]*?\b(rrtqx_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_declare_the_same_symbols():
    hs = header_symbols()
    assert len(hs) >= 35
    assert hs == sorted(A.SIGNATURES.keys())


def test_julia_module_binds_only_declared_symbols():
    """julia/RRTQXGpu.jl cannot be executed here (no Julia in the image); at least every symbol it `ccall`s must be
    one the header declares, and its block structure must be balanced."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "julia", "RRTQXGpu.jl")).read()
    bound = set(re.findall(r":(rrtqx_[a-z0-9_]+)", src))
    assert len(bound) >= 35
    assert bound <= set(header_symbols()), sorted(bound - set(header_symbols()))
    for name in ("RRTQXGpu.jl", "extend_gpu.jl", "make_reference_vectors.jl"):
        src = open(os.path.join(root, "julia", name)).read()
        depth = 0
        for line in src.splitlines():
            code = re.sub(r'"(\\.|[^"\\])*"', '""', line).split("#")[0]
            for _ in range(6):                                 # a[end], comprehensions: innermost brackets first
                code = re.sub(r"\[[^\[\]]*\]", "_", code)
            for tok in re.findall(r"(?<![\w.:])(function|if|for|while|let|struct|module|begin|do|try|quote|macro|end)\b", code):
                depth += -1 if tok == "end" else 1
                assert depth >= 0, (name, line)
        assert depth == 0, name


def test_library_exports_every_declared_symbol():
    L = A.lib()
    out = subprocess.run(["nm", "-D", "--defined-only", A.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (rrtqx_[a-z0-9_]+)", out))
    for name in header_symbols():
        assert name in exported, name
        assert getattr(L, name) is not None
    # nothing but the C ABI leaks out of the shared object
    assert all(s.startswith("rrtqx_") for s in re.findall(r" T (\S+)", out))
    assert L.rrtqx_version().decode().startswith("rrtqx-b200")


def test_library_is_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", A.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on the CPU box")
    L = A.lib()
    h = A.vp()
    st = L.rrtqx_ctx_create(0, None, C.byref(h))
    assert st == A.ERR_CUDA and not h.value
    msg = L.rrtqx_last_error(None).decode()
    assert "no CPU fallback" in msg
    from rrtqx_3d_b200.device import Context
    with pytest.raises(A.RRTQXError):
        Context(0)


def test_null_handles_are_rejected_not_dereferenced():
    L = A.lib()
    n = A.i64(0)
    assert L.rrtqx_tree_size(None, C.byref(n)) == A.ERR_INVALID
    assert L.rrtqx_ctx_sync(None) == A.ERR_INVALID
    assert L.rrtqx_tree_destroy(None) == A.OK          # destroying NULL is a no-op, like free()
    assert L.rrtqx_range_result_destroy(None) == A.OK


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "rrtqx_3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "rrtqx_oracle" not in txt and "orc_" not in txt, f
