#!/usr/bin/env python
"""bench.py — headline benchmark of the arrow-go compute hot path on B200.

Workload (BASELINE.json configs[1]): compute.Add(float64, float64) on a 100M-row CHUNKED array
per GPU (left: 100 chunks of 1M rows; right: chunks of 999,983 rows so the executor's span
iteration sees misaligned chunk boundaries, arrow/compute/executor.go:757-863), one contiguous
preallocated output (executor.go:598-623).  A "step" is one Add over the whole column.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ONE JSON line on rank 0 (contract in the task statement):
  value     rows/s, whole job (all ranks), kernels timed with CUDA events on device-resident
            buffers (1.6 GB in + 0.8 GB out per step: far larger than the 126 MB L2, so no flush)
  e2e       same metric through the HOST-pointer C ABI (ag_arith_binary on pinned host buffers):
            H2D of both inputs and D2H of the result inside the timed region
  roofline  achieved HBM GB/s of the Add kernel = 24 B/row x rows / CUDA-event time, against
            MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own AVX2 loop (oracle/_ref, assembled from the reference's
            checked-in assembly) on a bounded sample of the same workload, 1 thread — what a
            reference CallFunction uses (arrow/compute/exec.go:164-170) — plus an all-cores
            row-sharded figure for context
  others    kernel-only numbers for Sum / Greater / Filter / fused / Take at 100M rows
`--impl reference` times the reference's CPU path alone (same metric, same config).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS = 100_000_000
L_CHUNK = 1_000_000
R_CHUNK = 999_983
METRIC = "rows/sec, compute.Add(float64,float64) on a 100M-row chunked array per GPU"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the Add kernel from the committed
    `ncu --set full` capture (profiles/r1/ncu_full_binary_spans_kernel.csv), or None."""
    import csv
    path = os.path.join(ROOT, "profiles", "r1", "ncu_full_binary_spans_kernel.csv")
    try:
        rows = list(csv.reader(open(path)))
        h = rows[0]
        ri = [i for i, c in enumerate(h) if c.startswith("dram__bytes_read.sum")][0]
        wi = [i for i, c in enumerate(h) if c.startswith("dram__bytes_write.sum")][0]
        scale = lambda c: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[c[c.index("[") + 1:c.index("]")]]
        vals = [float(r[ri]) * scale(h[ri]) + float(r[wi]) * scale(h[wi]) for r in rows[1:]]
        return sum(vals) / len(vals)
    except Exception:
        return None


def spans_for(n, lc, rc):
    """iterateExecSpans (executor.go:757-863): span = min(remaining of each arg's current chunk)."""
    out, pos = [], 0
    while pos < n:
        l_rem = lc - pos % lc
        r_rem = rc - pos % rc
        ln = min(l_rem, r_rem, n - pos)
        out.append((pos, ln))
        pos += ln
    return out


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML is
    polled from a thread every few ms (nvidia-smi's 100 ms loop is too coarse for a 10-40 ms
    timed region); nvidia-smi is the fallback when pynvml is unavailable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device):
        self.device = device
        self.samples, self.max_mhz, self.reasons = [], None, set()
        self._stop = threading.Event()
        self._t = None
        self._nvml = None

    def _loop(self):
        nv, h = self._nvml, self._h
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nvml = nv
            self._h = nv.nvmlDeviceGetHandleByIndex(self.device)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        except Exception:
            self._nvml = None

    def stop(self):
        if self._nvml is None:
            return self._smi_once()
        self._stop.set()
        self._t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_min_mhz": float(min(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": "nvml, 2 ms poll during the timed regions"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.device}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1, "source": "nvidia-smi (idle snapshot)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}


# ------------------------------------------------------------------ reference / CPU arm -----
def cpu_reference_add(rows_sample, reps, threads):
    """The reference's own inner loop for this config: _arithmetic_binary_avx2(FLOAT64, OpAddChecked,
    l, r, out, len) (base_arithmetic_avx2_amd64.go:35-39), called once per span like
    executeSpans does (executor.go:598-623).  Returns rows/s."""
    from oracle import oracle
    ref = oracle.ref()
    isa = oracle.host_isa()
    if ref is not None:
        fn = getattr(ref, f"arithmetic_binary_{isa}")
        kind = "reference"

        def run(l, r, o, n):
            fn(12, 21, l, r, o, n)
    else:
        cpu = oracle.cpu()
        kind = "port"

        def run(l, r, o, n):
            cpu.ref_arith_binary(12, 21, 0, l, r, o, n)
    rng = np.random.default_rng(0x94378165)
    a = rng.integers(-(1 << 20), 1 << 20, rows_sample).astype(np.float64)
    b = rng.integers(-(1 << 20), 1 << 20, rows_sample).astype(np.float64)
    out = np.empty(rows_sample)
    spans = spans_for(rows_sample, L_CHUNK, R_CHUNK)

    def shard(lo, hi):
        for pos, ln in spans:
            s, e = max(pos, lo), min(pos + ln, hi)
            if e > s:
                run(a.ctypes.data + 8 * s, b.ctypes.data + 8 * s, out.ctypes.data + 8 * s, e - s)

    def step():
        if threads == 1:
            shard(0, rows_sample)
        else:
            cuts = np.linspace(0, rows_sample, threads + 1).astype(np.int64)
            list(pool.map(lambda i: shard(int(cuts[i]), int(cuts[i + 1])), range(threads)))

    pool = ThreadPoolExecutor(threads) if threads > 1 else None
    step()  # warm-up (page faults)
    best = float("inf")
    tot = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        best = min(best, dt)
        tot += dt
    if pool:
        pool.shutdown()
    assert np.array_equal(out[:1000], a[:1000] + b[:1000])
    return rows_sample * reps / tot, rows_sample / best, kind, isa


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = int(os.environ.get("AG_BENCH_REF_SAMPLE", "16000000"))  # rows per step of the bounded sample (tests shrink it)
    t0 = time.perf_counter()
    for _ in range(max(args.warmup, 1) - 1):
        pass
    one, one_best, kind, isa = cpu_reference_add(sample, max(args.steps, 3), 1)
    allc, allc_best, _, _ = cpu_reference_add(sample * 4, max(args.steps, 3), cores)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": one, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ROWS / one * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "compute.Add(float64,float64), 100M-row chunked array (1M-row x 999,983-row chunks), 1 call = 1 goroutine",
                   "rows": ROWS, "timing": "wall clock, host resident"},
        "cpu_baseline": {"value": one, "unit": "rows/s", "cores": 1, "kind": kind, "isa": isa,
                         "sample": f"{sample} rows x {max(args.steps, 3)} steps of the same chunked Add (the reference executes a CallFunction's spans on one goroutine, exec.go:164-170)",
                         "all_cores": {"value": allc, "unit": "rows/s", "cores": cores, "note": "row-range sharded over every host core; NOT a reference feature, context only"}},
        "e2e": {"value": one, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    _emit(line)


# ------------------------------------------------------------------ our arm ---------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS)
    ap.add_argument("--no-others", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL writes its version banner / debug lines to stdout by
        # default; send them to stderr instead (the level the caller asked for is left alone)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from arrow_go_b200 import _native as N
    from arrow_go_b200.device import DeviceBuffer, Event, PinnedArray

    N.call("ag_init", local_rank)
    rows = args.rows
    W, K = max(args.warmup, 3), max(args.steps, 1)
    peak, peak_kind = peaks()

    def barrier():
        N.call("ag_stream_sync", None)
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident record batch (seeds SURVEY §8d; each rank its own row range) ----
    dl, dr, dout = DeviceBuffer(rows * 8), DeviceBuffer(rows * 8), DeviceBuffer(rows * 8)
    N.call("ag_generate_dev", 3, 0x94378165 + rank * rows, -(1 << 20), 1 << 20, dl.ptr, rows, None)
    N.call("ag_generate_dev", 3, 0x94378166 + rank * rows, -(1 << 20), 1 << 20, dr.ptr, rows, None)
    spans = spans_for(rows, L_CHUNK, R_CHUNK)
    launches0 = N.raw().ag_kernel_launch_count()

    # the whole chunked call is ONE launch: the span table goes to the batched entry point a
    # compute.Function that sees ChunkedDatums binds (include/arrowgpu.h: ag_arith_binary_spans_dev)
    table = N.span_table([(dl.ptr + 8 * pos, dr.ptr + 8 * pos, dout.ptr + 8 * pos, ln) for pos, ln in spans])

    def add_step():
        N.call("ag_arith_binary_spans_dev", N.FLOAT64, N.OP_ADD_CHECKED, N.SHAPE_AA, table, len(spans), None)

    def add_step_per_span():  # what a per-span exec.ArrayKernelExec binding would do (context only)
        for pos, ln in spans:
            N.call("ag_arith_binary_dev", N.FLOAT64, N.OP_ADD_CHECKED, N.SHAPE_AA, dl.ptr + 8 * pos, dr.ptr + 8 * pos, dout.ptr + 8 * pos, ln, None)

    def timed(fn, warm, steps):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = Event(), Event()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        e1.sync()
        barrier()
        return max_over_ranks(e0.elapsed_ms(e1)) / steps

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = N.raw().ag_kernel_launch_count()
    ms_chunked = timed(add_step, W, K)
    launches_timed = (N.raw().ag_kernel_launch_count() - l0) * K // (W + K)
    ms_per_span = timed(add_step_per_span, 3, max(3, K // 4))
    ms_contig = timed(lambda: N.call("ag_arith_binary_dev", N.FLOAT64, N.OP_ADD_CHECKED, N.SHAPE_AA, dl.ptr, dr.ptr, dout.ptr, rows, None), W, K)

    # parity spot check inside the bench (oracle = checker only): first 64K rows of the last step
    if rank == 0:
        from oracle import oracle
        a = dl.to_numpy(np.float64, 1 << 16); b = dr.to_numpy(np.float64, 1 << 16)
        want = np.empty(1 << 16)
        oracle.cpu().ref_arith_binary(12, 21, 0, a.ctypes.data, b.ctypes.data, want.ctypes.data, 1 << 16)
        assert dout.to_numpy(np.float64, 1 << 16).tobytes() == want.tobytes(), "bench output differs from the oracle"

    value = world * rows / (ms_chunked * 1e-3)
    algo_bytes = 24.0 * rows
    achieved = algo_bytes / (ms_chunked * 1e-3) / 1e9

    # ---- other kernels of the path (kernel-only, device resident) ----
    others = {}
    if not args.no_others:
        scal = DeviceBuffer(64)
        ms = timed(lambda: N.call("ag_sum_f64_dev", dl.ptr, rows, scal.ptr, None), W, K)
        others["sum_f64"] = {"rows_per_s": world * rows / ms * 1e3, "gbs_per_gpu": 8.0 * rows / ms / 1e6, "frac": 8.0 * rows / ms / 1e6 / peak, "ms": ms}
        ms = timed(lambda: N.call("ag_sum_i64_dev", dl.ptr, rows, scal.ptr, None), W, K)
        others["sum_i64"] = {"rows_per_s": world * rows / ms * 1e3, "gbs_per_gpu": 8.0 * rows / ms / 1e6, "frac": 8.0 * rows / ms / 1e6 / peak, "ms": ms}
        if dist is not None:
            import torch
            t = torch.zeros(1, dtype=torch.int64, device="cuda")

            def sum_allreduce():
                N.call("ag_sum_i64_dev", dl.ptr, rows, scal.ptr, None)
                N.call("ag_copy_dev", t.data_ptr(), scal.ptr, 8, None)
                N.call("ag_stream_sync", None)
                dist.all_reduce(t)
            for _ in range(W):
                sum_allreduce()
            barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                sum_allreduce()
            torch.cuda.synchronize()
            ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
            others["sum_i64_global_nccl"] = {"rows_per_s": world * rows / ms * 1e3, "ms": ms, "note": "per-GPU Sum + 8-byte NCCL all-reduce, wall clock"}
        # Greater(int64, 89) -> mask ; Filter ; fused ; Take
        vi = dr  # reuse: regenerate as int64 uniform [0,100)
        N.call("ag_generate_dev", 1, 0x0FF1CE + rank * rows, 0, 99, vi.ptr, rows, None)
        sc = np.array([89], dtype=np.int64)
        mask = DeviceBuffer(rows // 8 + 64)
        ms = timed(lambda: N.call("ag_compare_dev", N.INT64, N.CMP_GT, N.SHAPE_AS, vi.ptr, sc.ctypes.data, mask.ptr, rows, 0, None), W, K)
        others["greater_i64_scalar"] = {"rows_per_s": world * rows / ms * 1e3, "gbs_per_gpu": 8.125 * rows / ms / 1e6, "frac": 8.125 * rows / ms / 1e6 / peak, "ms": ms}
        N.call("ag_filter_output_size_dev", mask.ptr, None, 0, rows, 0, scal.ptr, None)
        cnt = int(scal.to_numpy(np.int64, 1)[0])
        sel = cnt / rows
        ms = timed(lambda: N.call("ag_filter_primitive_dev", 64, vi.ptr, None, 0, mask.ptr, None, 0, rows, 0, dout.ptr, None, cnt, scal.ptr + 8, None), W, K)
        fb = (8 + 0.125 + 8 * sel) * rows
        others["filter_i64"] = {"rows_per_s": world * rows / ms * 1e3, "gbs_per_gpu": fb / ms / 1e6, "frac": fb / ms / 1e6 / peak, "ms": ms, "selectivity": sel}
        ms = timed(lambda: N.call("ag_filter_compare_scalar_dev", N.INT64, N.CMP_GT, vi.ptr, sc.ctypes.data, rows, dout.ptr, cnt, scal.ptr + 8, None), W, K)
        fb = (8 + 8 * sel) * rows
        others["fused_greater_filter_i64"] = {"rows_per_s": world * rows / ms * 1e3, "gbs_per_gpu": fb / ms / 1e6, "frac": fb / ms / 1e6 / peak, "ms": ms}
        idx = mask = None
        idx = DeviceBuffer(rows * 4)
        N.call("ag_generate_dev", 2, 0x0FF1CE + 7 + rank * rows, 0, rows - 1, idx.ptr, rows, None)
        bad = DeviceBuffer(64)
        N.call("ag_error_word_reset_dev", bad.ptr, None)
        ms = timed(lambda: N.call("ag_take_primitive_dev", 64, vi.ptr, None, 0, rows, 32, 1, idx.ptr, None, 0, rows, 1, dout.ptr, None, bad.ptr, None), W, K)
        others["take_i64_i32idx_random"] = {"rows_per_s": world * rows / ms * 1e3, "gbs_per_gpu": 20.0 * rows / ms / 1e6, "frac": 20.0 * rows / ms / 1e6 / peak, "ms": ms,
                                            "note": "algorithmic 20 B/row; random 8-byte gathers move 32-byte sectors"}
        # rows SURVEY §8(f) marks "next", same device-resident columns: promotion cast, min/max, cumulative sum
        ms = timed(lambda: N.call("ag_cast_numeric_dev", N.INT32, N.INT64, idx.ptr, dout.ptr, rows, None), W, K)
        others["cast_i32_to_i64"] = {"rows_per_s": world * rows / ms * 1e3, "gbs_per_gpu": 12.0 * rows / ms / 1e6, "frac": 12.0 * rows / ms / 1e6 / peak, "ms": ms}
        N.call("ag_error_word_reset_dev", bad.ptr, None)
        ms = timed(lambda: N.call("ag_cast_numeric_checked_dev", N.INT64, N.FLOAT64, vi.ptr, None, 0, dout.ptr, rows, 0, 0, bad.ptr, None), W, K)
        others["cast_i64_to_f64_safe"] = {"rows_per_s": world * rows / ms * 1e3, "gbs_per_gpu": 16.0 * rows / ms / 1e6, "frac": 16.0 * rows / ms / 1e6 / peak, "ms": ms}
        ms = timed(lambda: N.call("ag_min_max_dev", N.INT64, vi.ptr, rows, scal.ptr, None), W, K)
        others["min_max_i64"] = {"rows_per_s": world * rows / ms * 1e3, "gbs_per_gpu": 8.0 * rows / ms / 1e6, "frac": 8.0 * rows / ms / 1e6 / peak, "ms": ms}
        cstate = DeviceBuffer(64)

        def cumsum():
            N.call("ag_cumulative_sum_state_init_dev", cstate.ptr, N.INT64, None, None)
            N.call("ag_cumulative_sum_dev", N.INT64, vi.ptr, None, 0, rows, 0, 0, dout.ptr, None, 0, cstate.ptr, bad.ptr, None)
        ms = timed(cumsum, W, K)
        others["cumulative_sum_i64"] = {"rows_per_s": world * rows / ms * 1e3, "gbs_per_gpu": 16.0 * rows / ms / 1e6, "frac": 16.0 * rows / ms / 1e6 / peak, "ms": ms,
                                        "note": "single-pass scan, read once + write once (16 B/row)"}
        cstate.free(); idx.free(); bad.free(); scal.free()

    launches_total = N.raw().ag_kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: HOST buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e_rows = rows
        ha, hb, ho = PinnedArray(e_rows, np.float64), PinnedArray(e_rows, np.float64), PinnedArray(e_rows, np.float64)
        N.call("ag_download", ha.ptr, dl.ptr, e_rows * 8, None)
        N.call("ag_generate_dev", 3, 0x94378166 + rank * rows, -(1 << 20), 1 << 20, dr.ptr, rows, None)
        N.call("ag_download", hb.ptr, dr.ptr, e_rows * 8, None)
        N.call("ag_stream_sync", None)

        htable = N.span_table([(ha.ptr + 8 * pos, hb.ptr + 8 * pos, ho.ptr + 8 * pos, ln) for pos, ln in spans])

        def e2e_step():
            N.call("ag_arith_binary_spans", N.FLOAT64, N.OP_ADD_CHECKED, N.SHAPE_AA, htable, len(spans))
        ke = max(3, min(K, 10))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            e2e_step()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        assert np.array_equal(ho.array[:4096], ha.array[:4096] + hb.array[:4096])
        e2e = {"value": world * e_rows * ke / dt, "unit": "rows/s", "h2d_bytes_per_step": 16 * e_rows, "d2h_bytes_per_step": 8 * e_rows,
               "ms_per_step": dt / ke * 1e3, "link_gbs": 24.0 * e_rows * ke / dt / 1e9,
               "how": "ag_arith_binary_spans(host ptrs) over the same 200-span chunked layout on ag_host_alloc (pinned) buffers; synchronous API timed by wall clock, max over ranks"}
        ha.free(); hb.free(); ho.free()

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        one, _, kind, isa = cpu_reference_add(16_000_000, 5, 1)
        allc, _, _, _ = cpu_reference_add(64_000_000, 5, cores)
        cpu_baseline = {"value": one, "unit": "rows/s", "cores": 1, "kind": kind, "isa": isa,
                        "sample": "16M rows x 5 steps of the same chunked Add through the reference's arithmetic_binary_avx2 (1 goroutine per CallFunction, exec.go:164-170)",
                        "all_cores": {"value": allc, "unit": "rows/s", "cores": cores, "note": "row-range sharded over all host cores; not a reference feature"}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_chunked,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "compute.Add(float64,float64) on a 100M-row chunked array per GPU (BASELINE.json configs[1])",
                       "rows_per_gpu": rows, "chunks": f"left {L_CHUNK}-row chunks, right {R_CHUNK}-row chunks -> {len(spans)} spans, one contiguous output",
                       "l2": "inputs (1.6 GB) + output (0.8 GB) per step exceed the 126 MB L2; no flush needed", "parallelism": f"row-range x{world}",
                       "contiguous_ms_per_step": ms_contig, "per_span_launch_ms_per_step": ms_per_span},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                         "traffic_note": "bytes per launch, dram__bytes_read+write from profiles/r1/ncu_full_binary_spans_kernel.csv (same kernel, same shape)",
                         "peak_kind": peak_kind, "algorithmic_bytes_per_row": 24, "kernel": "binary_spans_kernel<double,OpAdd,AA>",
                         "contiguous_frac": 24.0 * rows / (ms_contig * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches_timed), "gpu_launches_total": int(launches_total),
            "clocks": clocks, "others": others,
        }
        _emit(line)
    if dist is not None:
        dist.destroy_process_group()


def _emit(line):
    """The ONE JSON line, on the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _claim_stdout():
    """Everything any library prints (NCCL's version banner, torch notices) goes to stderr: file
    descriptor 1 is pointed at stderr for the whole run and only _emit() writes to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


if __name__ == "__main__":
    main()
