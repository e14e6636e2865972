"""arrow_go_b200 — B200 (sm_100a) implementation of arrow-go's vectorised columnar compute
hot path: arrow/compute scalar arithmetic / comparison / boolean kernels, filter / take, and
the arrow/math Sum reductions, behind the C ABI of include/arrowgpu.h.

Layout
  csrc/        hand-written CUDA kernels + the C ABI (libarrowgpu.so, built in-tree into lib/)
  _native.py   ctypes FFI table of the C ABI (no logic)
  device.py    device buffers / streams / events for tests and bench.py
  compute.py   host-side mirror of the reference's compute API (CallFunction, Add, Filter,
               Take, ... over Datum / ArraySpan) driving the device kernels

There is no CPU fallback anywhere in this package: without libarrowgpu.so and an sm_100
device every compute call raises.
"""
from . import _native  # noqa: F401

__all__ = ["_native"]
