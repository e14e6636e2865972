"""Host-side mirror (arrow_go_b200/host) — the pieces that are pure metadata and therefore run
without a GPU: span iteration (restating ExecSpanItrSuite, arrow/compute/exec_internals_test.go:
394-582), registry contents (function names of Appendix A of SURVEY.md), exact-type dispatch
(functions.go:204-217)."""
import pytest

from arrow_go_b200 import compute as pc


def sizes(spans):
    return [s[1] for s in spans]


def test_iterate_basics():
    # TestBasics: two arrays of 100 + a scalar
    assert sizes(pc.iterate_exec_spans([[100], [100], []], [False, False, False])) == [100]
    assert sizes(pc.iterate_exec_spans([[100], [100], []], [False, False, False], 16)) == [16, 16, 16, 16, 16, 16, 4]


def test_iterate_input_validation():
    # TestInputValidation: lengths 10 vs 9 -> ErrInvalid either way round; a single array is fine
    for lens in ([[10], [9]], [[9], [10]]):
        with pytest.raises(pc.ArrowError) as e:
            pc.iterate_exec_spans(lens, [False, False])
        assert e.value.sentinel == "ErrInvalid"
    assert sizes(pc.iterate_exec_spans([[10]], [False])) == [10]


def test_iterate_chunked_arrays():
    # TestChunkedArrays: chunks {0,20,10} x {15,15} x array(30) x 2 scalars
    args = [[0, 20, 10], [15, 15], [30], [], []]
    chunked = [True, True, False, False, False]
    assert sizes(pc.iterate_exec_spans(args, chunked, 10)) == [10, 5, 5, 10]
    assert sizes(pc.iterate_exec_spans(args, chunked, 20)) == [15, 5, 10]
    assert sizes(pc.iterate_exec_spans(args, chunked, 30)) == [15, 5, 10]
    spans = pc.iterate_exec_spans(args, chunked, 10)
    assert [s[0] for s in spans] == [0, 10, 15, 20]
    assert [s[2][0] for s in spans] == [1, 1, 1, 2]  # zero-length first chunk is skipped
    assert [s[2][1] for s in spans] == [0, 0, 1, 1]


def test_iterate_zero_length():
    # TestZeroLengthInput
    assert pc.iterate_exec_spans([[]], [True]) == []
    assert pc.iterate_exec_spans([[0]], [False]) == []
    assert pc.iterate_exec_spans([[0]], [True]) == []


def test_bench_config_spans():
    # config 2 layout: 1M-row chunks against 999,983-row chunks
    spans = pc.iterate_exec_spans([[1_000_000] * 100, [999_983] * 100 + [1700]], [True, True])
    assert sum(sizes(spans)) == 100_000_000
    assert all(l > 0 for l in sizes(spans))
    assert len(spans) == 200


def test_registry_has_the_reference_function_names():
    names = set(pc.function_names())
    for n in ("add", "add_unchecked", "sub", "sub_unchecked", "subtract", "subtract_unchecked", "multiply", "multiply_unchecked",
              "abs", "abs_unchecked", "negate", "negate_unchecked", "sign", "is_null", "is_not_null", "is_nan", "equal", "not_equal", "greater", "greater_equal", "less", "less_equal",
              "and", "or", "xor", "and_not", "and_kleene", "or_kleene", "and_not_kleene", "not",
              "filter", "array_filter", "take", "array_take", "cast", "cast_int8", "cast_int16", "cast_int32", "cast_int64",
              "cast_uint8", "cast_uint16", "cast_uint32", "cast_uint64", "cast_float", "cast_double",
              "cumulative_sum", "cumulative_sum_checked"):
        assert n in names, n


def test_exact_type_dispatch():
    pc.dispatch("add", [pc.FLOAT64, pc.FLOAT64])
    pc.dispatch("greater", [pc.INT64, pc.INT64])
    pc.dispatch("and_kleene", [pc.BOOL, pc.BOOL])
    with pytest.raises(pc.ArrowError) as e:
        pc.dispatch("add", [pc.INT32, pc.FLOAT64])  # DispatchExact alone does not promote
    assert e.value.sentinel == "ErrNotImplemented"
    with pytest.raises(pc.ArrowError):
        pc.dispatch("no_such_function", [pc.INT32])


def test_binary_arithmetic_dispatch_best():
    """arrow/compute/arithmetic_test.go:715-753 TestBinaryArithmeticDispatchBest (numeric rows) and the
    compare functions' DispatchBest (scalar_compare.go:37-65)."""
    rows = [(pc.INT32, pc.INT32, pc.INT32), (pc.INT32, pc.INT8, pc.INT32), (pc.INT32, pc.INT16, pc.INT32),
            (pc.INT32, pc.INT64, pc.INT64), (pc.INT32, pc.UINT8, pc.INT32), (pc.INT32, pc.UINT16, pc.INT32),
            (pc.INT32, pc.UINT32, pc.INT64), (pc.INT32, pc.UINT64, pc.INT64), (pc.UINT8, pc.UINT8, pc.UINT8),
            (pc.UINT8, pc.UINT16, pc.UINT16), (pc.INT32, pc.FLOAT32, pc.FLOAT32), (pc.FLOAT32, pc.INT64, pc.FLOAT32),
            (pc.FLOAT64, pc.INT32, pc.FLOAT64)]
    for name in ("add", "sub", "multiply"):
        for suffix in ("", "_unchecked"):
            for l, r, want in rows:
                assert pc.dispatch_best(name + suffix, [l, r]) == [want, want], (name, l, r)
    for name in ("equal", "not_equal", "greater", "greater_equal", "less", "less_equal"):
        for l, r, want in rows:
            assert pc.dispatch_best(name, [l, r]) == [want, want]
    # unary functions and non-numeric arguments never promote
    with pytest.raises(pc.ArrowError):
        pc.dispatch_best("add", [pc.BOOL, pc.INT32])
    with pytest.raises(pc.ArrowError):
        pc.dispatch_best("and", [pc.INT8, pc.INT32])


def test_common_numeric_table():
    """commonNumeric (utils.go:178-240) against an independent model: floats win, otherwise the
    narrowest integer type that holds both ranges, saturating at 64 bits (uint64 + signed -> int64)."""
    ints = {pc.UINT8: (8, 0), pc.INT8: (8, 1), pc.UINT16: (16, 0), pc.INT16: (16, 1),
            pc.UINT32: (32, 0), pc.INT32: (32, 1), pc.UINT64: (64, 0), pc.INT64: (64, 1)}
    by = {v: k for k, v in ints.items()}
    allt = list(ints) + [pc.FLOAT32, pc.FLOAT64]
    for a in allt:
        for b in allt:
            if pc.FLOAT64 in (a, b):
                want = pc.FLOAT64
            elif pc.FLOAT32 in (a, b):
                want = pc.FLOAT32
            else:
                (wa, sa), (wb, sb) = ints[a], ints[b]
                if sa == sb:
                    want = by[(max(wa, wb), sa)]
                else:
                    ws, wu = (wa, wb) if sa else (wb, wa)
                    want = by[(min(64, ws if ws > wu else 2 * wu), 1)]
            assert pc.common_numeric([a, b]) == want, (a, b)
    assert pc.common_numeric([pc.BOOL, pc.INT8]) is None


def _model_spans(arg_lens, chunked, max_chunk):
    """Plain-Python model of iterateExecSpans (executor.go:757-863) to fuzz the C++ restatement."""
    totals = [sum(l) for l, c in zip(arg_lens, chunked) if l or c]
    length = totals[0] if totals else 1
    max_chunk = min(length, max_chunk)
    idx = [0] * len(arg_lens)
    posn = [0] * len(arg_lens)
    out, pos = [], 0
    while pos != length:
        it = min(length - pos, max_chunk)
        for i, (lens, c) in enumerate(zip(arg_lens, chunked)):
            if not c:
                continue
            while posn[i] == lens[idx[i]]:
                idx[i] += 1
                posn[i] = 0
            it = min(it, lens[idx[i]] - posn[i])
        out.append((pos, it, list(idx)))
        for i, (lens, c) in enumerate(zip(arg_lens, chunked)):
            if lens or c:
                posn[i] += it
        pos += it
    return out


def test_iterate_exec_spans_fuzz():
    import random
    rng = random.Random(0x0FF1CE)
    for _ in range(300):
        total = rng.randint(1, 500)
        nargs = rng.randint(1, 4)
        arg_lens, chunked = [], []
        for _a in range(nargs):
            kind = rng.choice(["array", "chunked", "scalar"])
            if kind == "scalar":
                arg_lens.append([]); chunked.append(False)
            elif kind == "array":
                arg_lens.append([total]); chunked.append(False)
            else:
                cuts = sorted(rng.sample(range(0, total + 1), rng.randint(0, min(6, total))))
                bounds = [0] + cuts + [total]
                lens = [b - a for a, b in zip(bounds, bounds[1:])]  # may contain zero-length chunks
                arg_lens.append(lens); chunked.append(True)
        if all(not l and not c for l, c in zip(arg_lens, chunked)):
            continue
        mc = rng.choice([1 << 62, 7, 64, total])
        got = pc.iterate_exec_spans(arg_lens, chunked, mc)
        want = _model_spans(arg_lens, chunked, mc)
        assert [(p, l) for p, l, _ in got] == [(p, l) for p, l, _ in want], (arg_lens, chunked, mc)
        for (_, _, gi), (_, _, wi) in zip(got, want):
            for a in range(nargs):
                if chunked[a]:
                    assert gi[a] == wi[a]


def test_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: rank 0 prints ONE JSON line on stdout (everything else goes to stderr).  The
    reference arm runs on the CPU, so the contract can be checked here."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--rows", "1000000", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, AG_BENCH_REF_ROWS="1000000"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0
