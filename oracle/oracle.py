"""ctypes loader for the ORACLE libraries (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under arrow_go_b200/ does.

  cpu()  -> oracle/libcpu_ref.so          our C restatement (cpu_ref.c)
  ref()  -> oracle/_ref/libarrowgo_ref.so the reference's own AVX2/SSE4 instruction stream,
            assembled from /root/reference by oracle/Makefile (None when it was never built)
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_cpu = None
_ref = None
_ref_tried = False

c_p = C.c_void_p
i64 = C.c_int64
NO_ERROR_POS = (1 << 63) - 1


def build(quiet=True):
    """Compile the restatement (and oracle/_ref when /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


_bench = None


def bench():
    """oracle/libbench_cpu.so: the C timing harness of bench.py's CPU legs (pinning, span loop, clock_gettime)."""
    global _bench
    if _bench is None:
        path = os.path.join(_HERE, "libbench_cpu.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.bench_pin_to_core.argtypes = [C.c_int]
        lib.bench_add_f64_chunked.restype = i64
        lib.bench_add_f64_chunked.argtypes = [c_p, C.c_int, C.c_int, c_p, c_p, c_p, i64, i64, i64, i64, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _bench = lib
    return _bench


def cpu():
    global _cpu
    if _cpu is None:
        path = os.path.join(_HERE, "libcpu_ref.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.ref_sum_f64_avx2_order.restype = C.c_double
        lib.ref_sum_f64_avx2_order.argtypes = [c_p, C.c_size_t]
        lib.ref_sum_f64_sequential.restype = C.c_double
        lib.ref_sum_f64_sequential.argtypes = [c_p, C.c_size_t]
        lib.ref_sum_i64.restype = i64
        lib.ref_sum_i64.argtypes = [c_p, C.c_size_t]
        lib.ref_sum_u64.restype = C.c_uint64
        lib.ref_sum_u64.argtypes = [c_p, C.c_size_t]
        lib.ref_arith_binary.argtypes = [C.c_int, C.c_int, C.c_int, c_p, c_p, c_p, i64]
        lib.ref_arith_unary_same.argtypes = [C.c_int, C.c_int, c_p, c_p, i64]
        lib.ref_arith_unary_diff.argtypes = [C.c_int, C.c_int, C.c_int, c_p, c_p, i64]
        lib.ref_arith_checked.argtypes = [C.c_int, C.c_int, C.c_int, c_p, c_p, i64, c_p, c_p, i64, c_p, i64, C.POINTER(i64)]
        lib.ref_arith_unary_checked.argtypes = [C.c_int, C.c_int, c_p, c_p, i64, C.POINTER(i64)]
        lib.ref_cumulative_sum.argtypes = [C.c_int, c_p, c_p, i64, i64, C.c_int, C.c_int, c_p, c_p, i64, c_p, C.POINTER(i64), C.POINTER(i64)]
        lib.ref_min_max.argtypes = [C.c_int, c_p, i64, c_p, c_p]
        lib.ref_cast_numeric.argtypes = [C.c_int, C.c_int, c_p, c_p, i64, c_p, i64, C.c_int, C.c_int, C.POINTER(i64)]
        lib.ref_compare.argtypes = [C.c_int, C.c_int, C.c_int, c_p, c_p, c_p, i64, C.c_int]
        lib.ref_bitmap_op.argtypes = [C.c_int, c_p, i64, c_p, i64, c_p, i64, i64]
        lib.ref_bitmap_copy.restype = None
        lib.ref_bitmap_copy.argtypes = [c_p, i64, i64, c_p, i64]
        lib.ref_bitmap_invert.restype = None
        lib.ref_bitmap_invert.argtypes = [c_p, i64, i64, c_p, i64]
        lib.ref_bitmap_set.restype = None
        lib.ref_bitmap_set.argtypes = [c_p, i64, i64, C.c_int]
        lib.ref_bitmap_popcount.restype = i64
        lib.ref_bitmap_popcount.argtypes = [c_p, i64, i64]
        lib.ref_kleene.argtypes = [C.c_int, c_p, c_p, i64, c_p, c_p, i64, c_p, c_p, i64, i64]
        lib.ref_filter_output_size.restype = i64
        lib.ref_filter_output_size.argtypes = [c_p, c_p, i64, i64, C.c_int]
        lib.ref_filter_primitive.argtypes = [C.c_int, c_p, c_p, i64, c_p, c_p, i64, i64, C.c_int, c_p, c_p,
                                             C.POINTER(i64), C.POINTER(i64)]
        lib.ref_take_indices.argtypes = [C.c_int, c_p, c_p, i64, i64, C.c_int, c_p, c_p, C.POINTER(i64)]
        lib.ref_take_primitive.argtypes = [C.c_int, c_p, c_p, i64, i64, C.c_int, C.c_int, c_p, c_p, i64, i64,
                                           C.c_int, c_p, c_p, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
        lib.ref_is_in.argtypes = [C.c_int, c_p, c_p, i64, i64, c_p, c_p, i64, i64, C.c_int, c_p, c_p, C.POINTER(i64)]
        lib.ref_unique.argtypes = [C.c_int, c_p, c_p, i64, i64, c_p, c_p, C.POINTER(i64), C.POINTER(i64)]
        lib.ref_sort_indices.argtypes = [C.c_int, c_p, c_p, i64, i64, C.c_int, C.c_int, c_p, C.POINTER(i64), C.POINTER(i64)]
        lib.ref_checksum64.restype = C.c_uint64
        lib.ref_checksum64.argtypes = [c_p, C.c_size_t]
        lib.ref_generate.restype = None
        lib.ref_generate.argtypes = [C.c_int, C.c_uint64, i64, i64, c_p, C.c_size_t]
        _cpu = lib
    return _cpu


def ref():
    """The reference's own SIMD loops (SysV ABI, all return void).  None if never built."""
    global _ref, _ref_tried
    if _ref is None and not _ref_tried:
        _ref_tried = True
        path = os.path.join(_HERE, "_ref", "libarrowgo_ref.so")
        if not os.path.exists(path) and os.path.isdir("/root/reference"):
            build()
        if os.path.exists(path):
            lib = C.CDLL(path)
            for isa in ("avx2", "sse4"):
                for t in ("float64", "int64", "uint64"):
                    f = getattr(lib, f"sum_{t}_{isa}")
                    f.restype = None
                    f.argtypes = [c_p, C.c_size_t, c_p]
                for nm in ("binary", "arr_scalar", "scalar_arr"):
                    f = getattr(lib, f"arithmetic_{nm}_{isa}")
                    f.restype = None
                    f.argtypes = [C.c_int, C.c_int8, c_p, c_p, c_p, C.c_int]
                f = getattr(lib, f"arithmetic_unary_same_types_{isa}")
                f.restype = None
                f.argtypes = [C.c_int, C.c_int8, c_p, c_p, C.c_int]
                f = getattr(lib, f"arithmetic_unary_diff_type_{isa}")
                f.restype = None
                f.argtypes = [C.c_int, C.c_int, C.c_int8, c_p, c_p, C.c_int]
                for t in ("int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64"):
                    f = getattr(lib, f"{t}_max_min_{isa}", None)  # (values, len, minout, maxout)
                    if f is not None:
                        f.restype = None
                        f.argtypes = [c_p, C.c_int, c_p, c_p]
                f = getattr(lib, f"cast_type_numeric_{isa}", None)  # (itype, otype, in, out, len)
                if f is not None:
                    f.restype = None
                    f.argtypes = [C.c_int, C.c_int, c_p, c_p, C.c_int]
                for op in ("equal", "not_equal", "greater", "greater_equal"):
                    for sh in ("arr_arr", "arr_scalar", "scalar_arr"):
                        f = getattr(lib, f"comparison_{op}_{sh}_{isa}")
                        f.restype = None
                        f.argtypes = [C.c_int, c_p, c_p, c_p, i64, C.c_int]
                for op in ("and", "or", "and_not", "xor"):
                    f = getattr(lib, f"bitmap_aligned_{op}_{isa}")
                    f.restype = None
                    f.argtypes = [c_p, c_p, c_p, i64]
            _ref = lib
    return _ref


def host_isa():
    """ISA the reference would pick on this host (arrow/math/math_amd64.go:26-34)."""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return "sse4"
    return "avx2" if " avx2" in flags else "sse4"
