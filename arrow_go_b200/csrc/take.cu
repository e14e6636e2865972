// take.cu — gather (Take) on sm_100a.
//
// Replaces PrimitiveTake (arrow/compute/internal/kernels/vector_selection.go:1162-1192) =
// checkIndexBounds (kernels/helpers.go:929-981) + primitiveTakeImpl (:878-988):
//     out[i]       = values[idx[i]]                         for slots valid in idx and in values
//     out_valid[i] = idxValid[i] ∧ valuesValid[idx[i]]
//     other slots keep the allocator's zero (we write 0)
// Indices of width 1/2/4/8 bytes are reinterpreted as unsigned after the bounds check
// (:1147-1159); only VALID index slots are bounds-checked (helpers.go:942); the error names the
// first offender in row order (:944-955) — here: atomicMin over the offending rows.
// The reference's sorted / reverse-sorted detection (:897-911) only changes its loop shape,
// never the result, and has no counterpart here.
//
// Roofline: HBM, sector-granular.  Algorithmic 4 + 8 + 8 = 20 B/row for int64 values and int32
// indices; a random 8-byte read moves a 32-byte sector, so DRAM traffic is up to 44 B/row when
// the values table exceeds L2 (126 MB).  Bounds check is fused into the gather: one pass over
// the indices instead of the reference's two.
//
// Layout: a block iteration covers 1024 consecutive rows; thread t handles rows t, t+256,
// t+512, t+768 so index loads and output stores are coalesced and 4 independent gathers are in
// flight per thread.  A warp's rows in one step are 32 consecutive rows starting at a multiple
// of 32, so the output validity word is one __ballot_sync.
#include "common.cuh"

#include <stdlib.h>

namespace ag {

constexpr int kTkThreads = 256;
constexpr int kTkUnroll = 4;

struct TakeParams {
  const void* vals;       // element 0 of the values buffer
  const uint8_t* vvalid;
  int64_t voff;
  unsigned long long vlen;
  const void* idx;
  const uint8_t* ivalid;
  int64_t ioff;
  int64_t n;
  int idx_signed;
  int bounds_check;
  void* out;
  uint32_t* out_valid;    // 4-byte aligned, offset 0 (may be NULL)
  long long* bad_pos;     // lowered with atomicMin (may be NULL when !bounds_check)
  int gather_mode;        // experiments (AG_TAKE_GATHER): cache / L2-prefetch-size hint of the random load
};

// The random 8-byte read is the whole cost of take (DRAM moves ~118 B per gathered row with the default
// load): these variants ask L2 for a smaller prefetch granule or keep the line out of L1.
template <typename V>
__device__ __forceinline__ V gather_load(const V* p, int mode) {
  if constexpr (sizeof(V) == 8) {
    unsigned long long r;
    switch (mode) {
      case 1: asm volatile("ld.global.L2::64B.u64 %0, [%1];" : "=l"(r) : "l"(p)); break;
      case 2: asm volatile("ld.global.L2::128B.u64 %0, [%1];" : "=l"(r) : "l"(p)); break;
      case 3: asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p)); break;
      case 4: asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(r) : "l"(p)); break;
      case 5: asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(r) : "l"(p)); break;
      case 6: asm volatile("ld.global.L2::256B.u64 %0, [%1];" : "=l"(r) : "l"(p)); break;
      default: return *p;
    }
    return *reinterpret_cast<V*>(&r);
  } else {
    return *p;
  }
}

template <typename V, typename I>
__global__ void __launch_bounds__(kTkThreads)
take_kernel(const TakeParams p) {
  const V* __restrict__ vals = reinterpret_cast<const V*>(p.vals) + p.voff;
  const I* __restrict__ idx = reinterpret_cast<const I*>(p.idx);
  V* __restrict__ out = reinterpret_cast<V*>(p.out);
  constexpr I kSignBit = (I)((I)1 << (sizeof(I) * 8 - 1));
  const int lane = threadIdx.x & 31;
  const int64_t step = (int64_t)gridDim.x * kTkThreads * kTkUnroll;
  long long my_bad = AG_NO_ERROR_POS;
  for (int64_t base = (int64_t)blockIdx.x * kTkThreads * kTkUnroll; base < p.n; base += step) {
    I ix[kTkUnroll];
    bool ok[kTkUnroll];
#pragma unroll
    for (int k = 0; k < kTkUnroll; ++k) {
      const int64_t i = base + k * kTkThreads + threadIdx.x;
      ok[k] = i < p.n;
      ix[k] = 0;
      if (ok[k]) {
        ix[k] = __ldcs(idx + i);
        if (p.ivalid) ok[k] = bit_is_set(p.ivalid, p.ioff + i);
      }
    }
    V v[kTkUnroll];
#pragma unroll
    for (int k = 0; k < kTkUnroll; ++k) {
      v[k] = V(0);
      if (ok[k]) {
        const bool neg = p.idx_signed && (ix[k] & kSignBit);
        const bool oob = neg || (unsigned long long)ix[k] >= p.vlen;
        if (oob) {
          ok[k] = false;  // never dereference; with bounds_check the whole call fails anyway
          if (p.bounds_check) {
            const long long i = base + k * kTkThreads + threadIdx.x;
            if (i < my_bad) my_bad = i;
          }
        } else {
          if (p.vvalid) ok[k] = bit_is_set(p.vvalid, p.voff + (int64_t)ix[k]);
          if (ok[k]) v[k] = gather_load(vals + ix[k], p.gather_mode);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kTkUnroll; ++k) {
      const int64_t i = base + k * kTkThreads + threadIdx.x;
      if (i < p.n) __stcs(out + i, v[k]);
      if (p.out_valid) {
        const int64_t w = (base + k * kTkThreads + (threadIdx.x & ~31)) >> 5;  // warp-uniform
        const uint32_t bits = __ballot_sync(0xffffffffu, ok[k]);
        const int64_t rem = p.n - (w << 5);
        if (lane == 0 && rem > 0) {
          if (rem >= 32) p.out_valid[w] = bits;
          else bitmap_store32_masked(p.out_valid + w, bits, bit_range_mask(0, (int)rem));
        }
      }
    }
  }
  if (p.bounds_check) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      const long long o = __shfl_xor_sync(0xffffffffu, my_bad, m);
      my_bad = o < my_bad ? o : my_bad;
    }
    if (lane == 0 && my_bad != AG_NO_ERROR_POS) atomicMin(p.bad_pos, my_bad);
  }
}

// booleanTakeImpl (vector_selection.go:990-1074): out bit i = values bit [voff + idx[i]]; same
// validity and bounds rules.  A warp's 32 rows of one step form one output word (ballot).
template <typename I>
__global__ void __launch_bounds__(kTkThreads)
take_bool_kernel(const TakeParams p) {
  const uint8_t* __restrict__ vals = reinterpret_cast<const uint8_t*>(p.vals);
  const I* __restrict__ idx = reinterpret_cast<const I*>(p.idx);
  uint32_t* __restrict__ out = reinterpret_cast<uint32_t*>(p.out);
  constexpr I kSignBit = (I)((I)1 << (sizeof(I) * 8 - 1));
  const int lane = threadIdx.x & 31;
  const int64_t step = (int64_t)gridDim.x * kTkThreads;
  long long my_bad = AG_NO_ERROR_POS;
  for (int64_t base = (int64_t)blockIdx.x * kTkThreads; base < p.n; base += step) {
    const int64_t i = base + threadIdx.x;
    bool ok = i < p.n, bit = false;
    if (ok) {
      const I ix = idx[i];
      if (p.ivalid) ok = bit_is_set(p.ivalid, p.ioff + i);
      if (ok) {
        const bool oob = (p.idx_signed && (ix & kSignBit)) || (unsigned long long)ix >= p.vlen;
        if (oob) {
          ok = false;
          if (p.bounds_check && (long long)i < my_bad) my_bad = (long long)i;
        } else {
          if (p.vvalid) ok = bit_is_set(p.vvalid, p.voff + (int64_t)ix);
          if (ok) bit = bit_is_set(vals, p.voff + (int64_t)ix);
        }
      }
    }
    const int64_t w = (base + (threadIdx.x & ~31)) >> 5;
    const uint32_t dbits = __ballot_sync(0xffffffffu, bit);
    const uint32_t vbits = __ballot_sync(0xffffffffu, ok);
    const int64_t rem = p.n - (w << 5);
    if (lane == 0 && rem > 0) {
      const uint32_t m = rem >= 32 ? 0xffffffffu : bit_range_mask(0, (int)rem);
      bitmap_store32_masked(out + w, dbits, m);
      if (p.out_valid) bitmap_store32_masked(p.out_valid + w, vbits, m);
    }
  }
  if (p.bounds_check) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      const long long o = __shfl_xor_sync(0xffffffffu, my_bad, m);
      my_bad = o < my_bad ? o : my_bad;
    }
    if (lane == 0 && my_bad != AG_NO_ERROR_POS) atomicMin(p.bad_pos, my_bad);
  }
}

static ag_status launch_take_bool(int idx_width, const TakeParams& p, cudaStream_t st) {
  const int64_t need = (p.n + kTkThreads - 1) / kTkThreads;
  switch (idx_width) {
    case 8: take_bool_kernel<uint8_t><<<grid_one_wave(take_bool_kernel<uint8_t>, kTkThreads, need), kTkThreads, 0, st>>>(p); break;
    case 16: take_bool_kernel<uint16_t><<<grid_one_wave(take_bool_kernel<uint16_t>, kTkThreads, need), kTkThreads, 0, st>>>(p); break;
    case 32: take_bool_kernel<uint32_t><<<grid_one_wave(take_bool_kernel<uint32_t>, kTkThreads, need), kTkThreads, 0, st>>>(p); break;
    case 64: take_bool_kernel<unsigned long long><<<grid_one_wave(take_bool_kernel<unsigned long long>, kTkThreads, need), kTkThreads, 0, st>>>(p); break;
    default: AG_FAIL(AG_ERR_INDEX, "take: invalid indices byte width");
  }
  return check_launch("take_bool_kernel");
}

template <typename V>
static ag_status launch_take_v(int idx_width, const TakeParams& p, cudaStream_t st) {
  const int64_t need = (p.n + kTkThreads * kTkUnroll - 1) / (kTkThreads * kTkUnroll);
  switch (idx_width) {
    case 8: take_kernel<V, uint8_t><<<grid_one_wave(take_kernel<V, uint8_t>, kTkThreads, need), kTkThreads, 0, st>>>(p); break;
    case 16: take_kernel<V, uint16_t><<<grid_one_wave(take_kernel<V, uint16_t>, kTkThreads, need), kTkThreads, 0, st>>>(p); break;
    case 32: take_kernel<V, uint32_t><<<grid_one_wave(take_kernel<V, uint32_t>, kTkThreads, need), kTkThreads, 0, st>>>(p); break;
    case 64: take_kernel<V, unsigned long long><<<grid_one_wave(take_kernel<V, unsigned long long>, kTkThreads, need), kTkThreads, 0, st>>>(p); break;
    default: AG_FAIL(AG_ERR_INDEX, "take: invalid indices byte width");  // vector_selection.go:1157
  }
  return check_launch("take_kernel");
}

ag_status take_primitive_dev(int bit_width, const void* vals, const uint8_t* vvalid, int64_t voff, int64_t vlen,
                             int idx_width, int idx_signed, const void* idx, const uint8_t* ivalid, int64_t ioff,
                             int64_t n, int bounds_check, void* out, uint8_t* out_valid, int64_t* d_bad_pos, cudaStream_t st) {
  if (n < 0 || voff < 0 || ioff < 0 || vlen < 0) AG_FAIL(AG_ERR_INVALID, "take: negative length or offset");
  if (n == 0) return AG_OK;
  if (!idx || !out) AG_FAIL(AG_ERR_INVALID, "take: NULL indices/output");
  if (!vals && vlen > 0) AG_FAIL(AG_ERR_INVALID, "take: NULL values");
  if (bounds_check && !d_bad_pos) AG_FAIL(AG_ERR_INVALID, "take: bounds_check needs an error word");
  if (out_valid && (reinterpret_cast<uintptr_t>(out_valid) & 3)) AG_FAIL(AG_ERR_INVALID, "take: out_valid must be 4-byte aligned");
  if ((vvalid || ivalid) && !out_valid)
    AG_FAIL(AG_ERR_INVALID, "take: inputs may have nulls but no output validity buffer was given (vector_selection.go:1175)");
  TakeParams p{};
  p.vals = vals; p.vvalid = vvalid; p.voff = voff; p.vlen = (unsigned long long)vlen;
  p.idx = idx; p.ivalid = ivalid; p.ioff = ioff; p.n = n;
  p.idx_signed = idx_signed; p.bounds_check = bounds_check;
  p.out = out; p.out_valid = reinterpret_cast<uint32_t*>(out_valid); p.bad_pos = reinterpret_cast<long long*>(d_bad_pos);
  { static int mode = -1; if (mode < 0) { const char* e = getenv("AG_TAKE_GATHER"); mode = e ? atoi(e) : 0; } p.gather_mode = mode; }
  switch (bit_width) {
    case 8: return launch_take_v<uint8_t>(idx_width, p, st);
    case 16: return launch_take_v<uint16_t>(idx_width, p, st);
    case 32: return launch_take_v<uint32_t>(idx_width, p, st);
    case 64: return launch_take_v<unsigned long long>(idx_width, p, st);
    case 1:
      if (reinterpret_cast<uintptr_t>(out) & 3) AG_FAIL(AG_ERR_INVALID, "take: boolean output bitmap must be 4-byte aligned");
      return launch_take_bool(idx_width, p, st);
    default: AG_FAIL(AG_ERR_INVALID, "take: invalid values byte width for take");  // vector_selection.go:1190
  }
}

}  // namespace ag

using namespace ag;

extern "C" ag_status ag_take_primitive_dev(int bit_width, const void* vals, const uint8_t* vvalid, int64_t voff, int64_t vlen,
                                           int idx_width, int idx_signed, const void* idx, const uint8_t* ivalid, int64_t ioff,
                                           int64_t n, int bounds_check, void* out, uint8_t* out_valid,
                                           int64_t* d_bad_pos, ag_stream_t s) {
  AG_TRY(ensure_init());
  return take_primitive_dev(bit_width, vals, vvalid, voff, vlen, idx_width, idx_signed, idx, ivalid, ioff, n,
                            bounds_check, out, out_valid, d_bad_pos, resolve_stream(s));
}
