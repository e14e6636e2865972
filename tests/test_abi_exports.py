"""CPU-side checks of the drop-in boundary: libarrowgpu.so loads, exports every symbol that
include/arrowgpu.h declares, validates arguments before touching the device, and fails loudly
(never falls back to a CPU path) when no GPU is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from arrow_go_b200 import _native as N


def test_library_exports_every_declared_symbol():
    lib = N.raw()
    declared = N.declared_symbols()
    assert len(declared) >= 80
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_table_matches_header():
    declared = set(N.declared_symbols())
    bound = set(N._SIGS) | set(N._SPECIAL)
    assert declared == bound, (declared - bound, bound - declared)


def test_go_package_calls_only_declared_entry_points():
    """go/arrowgpu cannot be compiled here (no Go toolchain): at least every C.ag_* function and C.AG_* constant the cgo files
    use must exist in include/arrowgpu.h with the number of arguments the call passes."""
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = re.sub(r"/\*.*?\*/", "", open(N.HEADER_PATH).read(), flags=re.S)
    declared = set(N.declared_symbols())
    arity = {}
    for m in re.finditer(r"ag_status\s+(ag_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        arity[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    consts = set(re.findall(r"#define\s+(AG_\w+)", text)) | set(re.findall(r"\b(AG_\w+)\s*=", text))
    files = sorted(glob.glob(os.path.join(root, "go", "arrowgpu", "*.go")))
    assert len(files) >= 7
    calls = 0
    for f in files:
        src = re.sub(r"//[^\n]*", "", open(f).read())
        for m in re.finditer(r"\bC\.(ag_\w+)\(", src):
            name = m.group(1)
            assert name in declared, (os.path.basename(f), name)
            # count the arguments of the call (top-level commas up to the matching parenthesis)
            depth, i, commas, seen = 1, m.end(), 0, False
            while depth:
                ch = src[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                elif ch == "," and depth == 1:
                    commas += 1
                if depth and not ch.isspace():
                    seen = True
                i += 1
            if name in arity:
                assert (commas + 1 if seen else 0) == arity[name], (os.path.basename(f), name, commas + 1, arity[name])
            calls += 1
        for name in re.findall(r"\bC\.(AG_\w+)", src):
            assert name in consts, (os.path.basename(f), name)
    assert calls >= 30


def test_header_cites_reference_for_every_block():
    text = open(N.HEADER_PATH).read()
    for needle in ("arrow/math/_lib/float64.c:20-26", "_lib/base_arithmetic.cc:465-483", "_lib/scalar_comparison.cc:210-256",
                   "arrow/bitutil/_lib/bitmap_ops.c:24-46", "vector_selection.go:449-520", "vector_selection.go:1162-1192",
                   "helpers.go:929-981", "arrow/errors.go:21-28"):
        assert needle in text, needle
    assert "torch" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def test_enums_match_reference_values():
    # arrow/datatype.go:36-72, kernels/base_arithmetic.go:37-82, kernels/types.go:62-71
    assert (N.BOOL, N.UINT8, N.INT8, N.UINT16, N.INT16, N.UINT32, N.INT32, N.UINT64, N.INT64, N.FLOAT32, N.FLOAT64) == (1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12)
    assert (N.OP_ADD, N.OP_SUB, N.OP_MUL, N.OP_DIV, N.OP_ABS, N.OP_NEGATE, N.OP_SIGN) == (0, 1, 2, 3, 4, 5, 20)
    assert (N.OP_ADD_CHECKED, N.OP_SUB_CHECKED, N.OP_MUL_CHECKED, N.OP_DIV_CHECKED, N.OP_ABS_CHECKED, N.OP_NEGATE_CHECKED) == (21, 22, 23, 24, 25, 26)
    assert (N.CMP_EQ, N.CMP_NE, N.CMP_GT, N.CMP_GE, N.CMP_LT, N.CMP_LE) == (0, 1, 2, 3, 4, 5)
    text = open(N.HEADER_PATH).read()
    for name, val in (("AG_TYPE_FLOAT64", 12), ("AG_TYPE_INT64", 9), ("AG_OP_ADD_CHECKED", 21), ("AG_OP_SIGN", 20), ("AG_EMIT_NULLS", 1)):
        assert re.search(rf"#define {name}\s+{val}\b", text), name


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback_without_gpu():
    """Without a device every compute entry point must fail with AG_ERR_CUDA — never compute."""
    x = np.arange(16, dtype=np.float64)
    out = np.full(16, -1.0)
    r = C.c_double(-1.0)
    st, msg = N.call_status("ag_sum_f64", x.ctypes.data, 16, C.byref(r))
    assert st == N.AG_ERR_CUDA and "no CPU fallback" in msg and r.value == -1.0
    st, _ = N.call_status("ag_arith_binary", N.FLOAT64, N.OP_ADD, x.ctypes.data, x.ctypes.data, out.ctypes.data, 16)
    assert st == N.AG_ERR_CUDA and (out == -1.0).all()
    st, _ = N.call_status("ag_compare", N.FLOAT64, N.CMP_GT, N.SHAPE_AA, x.ctypes.data, x.ctypes.data, out.ctypes.data, 16, 0)
    assert st == N.AG_ERR_CUDA
    ln = C.c_int64()
    st, _ = N.call_status("ag_filter_primitive", 64, x.ctypes.data, None, 0, x.ctypes.data, None, 0, 16, 0, out.ctypes.data, None, C.byref(ln), None)
    assert st == N.AG_ERR_CUDA
    st, _ = N.call_status("ag_take_primitive", 64, x.ctypes.data, None, 0, 16, 32, 1, x.ctypes.data, None, 0, 4, 1, out.ctypes.data, None, None, None, None)
    assert st == N.AG_ERR_CUDA


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under arrow_go_b200/ may reference it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "arrow_go_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "cpu_ref" not in text and "libarrowgo_ref" not in text, os.path.join(dirpath, f)
                for line in text.splitlines():
                    if re.search(r"^\s*(from|import)\s+oracle", line):
                        raise AssertionError(f"{f}: {line}")
