// common.cuh — runtime internals shared by every translation unit of libarrowgpu.so.
// Not part of the public ABI (see include/arrowgpu.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <atomic>
#include <mutex>

#include "../../include/arrowgpu.h"

namespace ag {

// ---------------------------------------------------------------- errors ----------
void set_error(const char* fmt, ...);
ag_status cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define AG_CUDA_TRY(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return ::ag::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define AG_TRY(expr)                     \
  do {                                   \
    ag_status _s = (expr);               \
    if (_s != AG_OK) return _s;          \
  } while (0)

#define AG_FAIL(code, ...)               \
  do {                                   \
    ::ag::set_error(__VA_ARGS__);        \
    return (code);                       \
  } while (0)

// ---------------------------------------------------------------- runtime ---------
// Per-stream scratch: kernels on one stream are serialised by the stream, so one
// workspace per stream is race-free without stream-ordered allocation per call.
struct Workspace {
  // fixed small area
  void* partials;        // kMaxPartials * 16 bytes (sum partials, per-block counters)
  unsigned* ticket;      // last-block-done tickets; kernels leave them at 0
  int64_t* scalars;      // 16 x int64 device scalars (counts, error words)
  // growable area (decoupled look-back tile status for filter / take_indices)
  unsigned long long* tile_status;
  size_t tile_status_cap;  // in elements
  // pinned host mirror of `scalars` for cheap readback
  int64_t* h_scalars;
  // Held by an entry point from get_workspace until its last launch is enqueued: two threads that share a stream
  // (e.g. both pass NULL) may interleave their CALLS but never the launches of one call with another's, so ticket /
  // scalar slots / tile status are always used by one kernel sequence at a time (the stream orders the rest).
  std::mutex* seq_mu;
};
struct WorkspaceLock {
  std::unique_lock<std::mutex> lk;
  explicit WorkspaceLock(Workspace* ws) : lk(*ws->seq_mu) {}
};

constexpr int kMaxPartials = 4096;

ag_status ensure_init();
int sm_count();
cudaStream_t resolve_stream(ag_stream_t s);
int current_device();
int device_count();
ag_status get_workspace(cudaStream_t s, Workspace** ws);
ag_status ensure_tile_status(Workspace* ws, size_t n_tiles, cudaStream_t s);
void count_launch(int n = 1);
ag_status check_launch(const char* what);

// Pooled streams for the synchronous host-pointer entry points (each call owns its streams
// for its duration, so concurrent cgo callers never share a stream or a workspace).
ag_status acquire_call_stream(cudaStream_t* st);
void release_call_stream(cudaStream_t st);
struct CallStream {
  cudaStream_t st = nullptr;
  ag_status acquire() { return acquire_call_stream(&st); }
  ~CallStream() { if (st) release_call_stream(st); }
  operator cudaStream_t() const { return st; }
};

// Device temp allocation (stream-ordered pool).
ag_status dev_alloc_async(void** p, size_t nbytes, cudaStream_t s);
ag_status dev_free_async(void* p, cudaStream_t s);

// ---------------------------------------------------------------- cross-GPU exchange (comm.cu, reduce.cu) ----
struct alignas(32) MailSlot { unsigned long long flag; unsigned long long pad; unsigned long long s; unsigned long long c; };
struct SumExchange {
  MailSlot* const* peers;   // device array [world]: base of every rank's mailbox as addressable from this device
  MailSlot* local;          // this rank's mailbox
  int world, rank;
  unsigned long long epoch; // >= 1, same value on every rank for one collective
};
ag_status comm_next_exchange(ag_comm_t comm, SumExchange* x);   // bumps the communicator's epoch
ag_status comm_nccl_allreduce_sum_i64(ag_comm_t comm, void* d_buf, size_t count, cudaStream_t st);

// ---------------------------------------------------------------- type helpers ----
inline int type_width(int type) {
  switch (type) {
    case AG_TYPE_UINT8: case AG_TYPE_INT8: return 1;
    case AG_TYPE_UINT16: case AG_TYPE_INT16: return 2;
    case AG_TYPE_UINT32: case AG_TYPE_INT32: case AG_TYPE_FLOAT32: return 4;
    case AG_TYPE_UINT64: case AG_TYPE_INT64: case AG_TYPE_FLOAT64: return 8;
    default: return 0;
  }
}
inline bool type_is_signed_int(int type) {
  return type == AG_TYPE_INT8 || type == AG_TYPE_INT16 || type == AG_TYPE_INT32 || type == AG_TYPE_INT64;
}
inline bool type_is_unsigned_int(int type) {
  return type == AG_TYPE_UINT8 || type == AG_TYPE_UINT16 || type == AG_TYPE_UINT32 || type == AG_TYPE_UINT64;
}
inline bool type_is_float(int type) { return type == AG_TYPE_FLOAT32 || type == AG_TYPE_FLOAT64; }

inline int64_t bytes_for_bits(int64_t nbits) { return (nbits + 7) >> 3; }

int blocks_per_sm(const void* kernel, int threads);
// Opt a kernel into `bytes` of dynamic shared memory on the calling thread's device, once per device
// (`done` = the caller's per-instantiation bit set; the attribute is per function per context).
inline ag_status ensure_dynamic_smem(const void* kernel, int bytes, std::atomic<unsigned>* done) {
  const unsigned bit = 1u << (current_device() & 31);
  if (done->load(std::memory_order_acquire) & bit) return AG_OK;
  AG_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done->fetch_or(bit, std::memory_order_release);
  return AG_OK;
}
// One-wave grid for a grid-stride kernel: min(blocks needed, SMs x resident blocks per SM).
template <typename K>
inline int grid_one_wave(K kernel, int threads, int64_t blocks_needed) {
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm(reinterpret_cast<const void*>(kernel), threads);
  if (blocks_needed < 1) blocks_needed = 1;
  return (int)(blocks_needed < cap ? blocks_needed : cap);
}

// Grid sizing: persistent-style grids are multiples of the SM count.
inline int grid_for(int64_t work_items, int items_per_block, int blocks_per_sm) {
  int64_t want = (work_items + items_per_block - 1) / items_per_block;
  int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// ---------------------------------------------------------------- device helpers --
#ifdef __CUDACC__

// Read 32 bits of an LSB-first bitmap starting at absolute bit position `bit`
// (relative to byte pointer `base`), touching only bytes in [lo_byte, hi_byte).
// Bits that fall outside are returned as 0.  Interior windows are two aligned 32-bit loads and
// a funnel shift; windows that touch the ends of the range go through the (out-of-line) byte-wise
// path so that no byte outside the caller's buffer is ever read.
static __device__ __noinline__ uint32_t bitmap_load32_edge(const uint8_t* abase, int64_t w, int sh, int64_t lo, int64_t hi) {
  uint32_t words[2] = {0u, 0u};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int64_t b0 = (w + j) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t b = b0 + k;
      if (b >= lo && b < hi) words[j] |= (uint32_t)abase[b] << (8 * k);
    }
  }
  return sh ? __funnelshift_r(words[0], words[1], sh) : words[0];
}

__device__ __forceinline__ uint32_t bitmap_load32(const uint8_t* __restrict__ base, int64_t bit,
                                                  int64_t lo_byte, int64_t hi_byte) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(base);
  const int64_t mis = (int64_t)(addr & 3);       // base = abase + mis, abase 4-byte aligned
  const uint8_t* abase = base - mis;
  const int64_t abit = bit + mis * 8;            // bit position relative to abase (floor semantics for negatives)
  const int64_t w = abit >> 5;
  const int sh = (int)(abit & 31);
  const int64_t lo = lo_byte + mis, hi = hi_byte + mis;  // valid byte range relative to abase
  const int64_t b0 = w * 4;
  if (b0 >= lo && b0 + 8 <= hi) {
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(abase + b0);
    const uint32_t w0 = wp[0];
    return sh ? __funnelshift_r(w0, wp[1], sh) : w0;
  }
  return bitmap_load32_edge(abase, w, sh, lo, hi);
}

// mask with bits [a, b) set, 0 <= a <= b <= 32
__device__ __forceinline__ uint32_t bit_range_mask(int a, int b) {
  const uint32_t hi = (b >= 32) ? 0xffffffffu : ((1u << b) - 1u);
  const uint32_t lo = (a >= 32) ? 0xffffffffu : ((1u << a) - 1u);
  return hi & ~lo;
}

// Store the bits of `value` selected by `mask` into the aligned 32-bit word at `wp`,
// preserving every other bit and never touching a byte whose 8 mask bits are all zero
// (edge words of a bitmap may extend past the caller's buffer).
__device__ __forceinline__ void bitmap_store32_masked(uint32_t* wp, uint32_t value, uint32_t mask) {
  if (mask == 0xffffffffu) { *wp = value; return; }
  uint8_t* bp = reinterpret_cast<uint8_t*>(wp);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t m = (mask >> (8 * k)) & 0xffu;
    if (m == 0) continue;
    const uint32_t v = (value >> (8 * k)) & 0xffu;
    if (m == 0xffu) bp[k] = (uint8_t)v;
    else bp[k] = (uint8_t)((bp[k] & ~m) | (v & m));
  }
}

__device__ __forceinline__ bool bit_is_set(const uint8_t* __restrict__ bits, int64_t i) {
  return (bits[i >> 3] >> (i & 7)) & 1;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

#endif  // __CUDACC__

}  // namespace ag
