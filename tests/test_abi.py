"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/resnmtf_b200.h
declares; compute entry points fail loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from resnmtf_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "resnmtf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(resnmtf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(L.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = L.load()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.resnmtf_version()
    assert lib.resnmtf_comm_id_size() >= 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "resnmtf_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.replace("# oracle", ""), f"{fn} mentions the oracle"


@pytest.mark.skipif(L.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_a_gpu():
    lib = L.load()
    h = C.c_void_p()
    rc = lib.resnmtf_ctx_create(-1, C.byref(h))
    assert rc == L.E_CUDA
    assert b"no CPU fallback" in lib.resnmtf_last_error()
    with pytest.raises(RuntimeError):
        from resnmtf_b200.device import Context

        Context()


def test_r_shim_compiles_against_the_r_api_declarations():
    """No R in this image: the .Call shim is at least compiled (syntax, types, argument counts against R's C API as
    declared in tests/r_stub/, and against include/resnmtf_b200.h) with warnings as errors."""
    import subprocess

    cmd = ["gcc", "-std=gnu99", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-fsyntax-only",
           "-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "resnmtf_b200", "r", "r_shim.c")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
