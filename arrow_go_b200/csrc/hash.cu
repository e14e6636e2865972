// hash.cu — is_in and unique for fixed-width columns on sm_100a.
//
// Replaces, for 1/2/4/8-byte values, the memo-table kernels of the reference:
//   is_in   arrow/compute/internal/kernels/scalar_set_lookup.go:112-413 (SetLookupState.Init builds a memo table of
//           the value set keyed by the RAW BYTES of a value — floats compare by bit pattern: NaN == NaN of the same
//           payload, -0.0 != +0.0; isInKernelExec :373-413 writes a data bit and a validity bit per row by the
//           NullMatchingBehavior table restated in is_in_kernel below);
//   unique  arrow/compute/internal/kernels/vector_hash.go (the distinct values in order of FIRST appearance, a null
//           kept once at the position of its first appearance).
// Both sit on the same open-addressing table in HBM: 64-bit keys (the value's bytes, zero-extended), linear probing,
// load factor <= 0.5, inserted with atomicCAS; every slot also keeps the lowest row that carried its key (atomicMin),
// which is all `unique` needs: a row is emitted iff it is the first row of its key, and the emitted rows are compacted
// in row order by the filter kernel (filter.cu) — so the output order is the reference's and does not depend on the
// order in which threads reached the table.
// Roofline: the probe is take's access pattern (one random 8-byte read per row into a table that is L2-resident for
// value sets up to a few million entries); `unique` on n rows sizes its table by n (2n slots rounded up to a power of
// two, 12 bytes per slot).
#include "common.cuh"

namespace ag {

ag_status filter_primitive_dev(int bit_width, const void* vals, const uint8_t* vvalid, int64_t voff, const uint8_t* mask,
                               const uint8_t* mvalid, int64_t moff, int64_t n, int null_selection, void* out, uint8_t* out_valid,
                               int64_t capacity, int64_t* d_out_len, cudaStream_t st);

constexpr int kHThreads = 256;
constexpr unsigned long long kEmptyKey = 0xffffffffffffffffull;
constexpr unsigned kNoRow = 0xffffffffu;

struct HashTable {
  unsigned long long* keys;   // [slots], kEmptyKey = free
  unsigned* first;            // [slots] lowest row that carried the key
  unsigned long long mask;    // slots - 1
  unsigned* special;          // [2]: lowest row whose key is the all-ones pattern (cannot live in `keys`), lowest NULL row
};

template <typename V>
__device__ __forceinline__ unsigned long long raw_key(const V* __restrict__ vals, int64_t i) {
  return (unsigned long long)vals[i];
}

__device__ __forceinline__ void table_insert(const HashTable& t, unsigned long long key, unsigned row) {
  if (key == kEmptyKey) { atomicMin(&t.special[0], row); return; }
  unsigned long long h = mix64(key) & t.mask;
  while (true) {
    unsigned long long cur = t.keys[h];
    if (cur == kEmptyKey) cur = atomicCAS(&t.keys[h], kEmptyKey, key);
    if (cur == kEmptyKey || cur == key) {
      // rows arrive roughly in increasing order, so after a key's first few rows the slot already holds a lower row:
      // a plain load then replaces the atomic (100M rows over 100 distinct keys would otherwise serialise 100M
      // atomicMin on 100 addresses)
      if (*reinterpret_cast<volatile unsigned*>(&t.first[h]) > row) atomicMin(&t.first[h], row);
      return;
    }
    h = (h + 1) & t.mask;
  }
}
// lowest row that carried `key`, kNoRow when absent
__device__ __forceinline__ unsigned table_first_row(const HashTable& t, unsigned long long key) {
  if (key == kEmptyKey) return t.special[0];
  unsigned long long h = mix64(key) & t.mask;
  while (true) {
    const unsigned long long cur = t.keys[h];
    if (cur == key) return t.first[h];
    if (cur == kEmptyKey) return kNoRow;
    h = (h + 1) & t.mask;
  }
}

template <typename V>
__global__ void __launch_bounds__(kHThreads)
hash_insert_kernel(const V* __restrict__ vals, const uint8_t* __restrict__ valid, int64_t off, int64_t n, const HashTable t) {
  for (int64_t i = (int64_t)blockIdx.x * kHThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kHThreads) {
    if (valid && !bit_is_set(valid, off + i)) atomicMin(&t.special[1], (unsigned)i);
    else table_insert(t, raw_key(vals + off, i), (unsigned)i);
  }
}

// isInKernelExec (scalar_set_lookup.go:373-413), per row:
//   valid value : in the set -> (true, valid); else INCONCLUSIVE with a null in the set -> (false, null); else (false, valid)
//   null        : MATCH with a null in the set -> (true, valid); SKIP, or MATCH without one -> (false, valid); else (false, null)
constexpr int kIsInSmemSlots = 4096;   // 32 KB of keys: value sets up to 2048 entries are probed in shared memory
constexpr int kIsInUnroll = 4;

template <typename V, bool kSmem>
__global__ void __launch_bounds__(kHThreads)
is_in_kernel(const V* __restrict__ vals, const uint8_t* __restrict__ valid, int64_t off, int64_t n, const HashTable t, int null_behavior,
             uint32_t* __restrict__ out_data, uint32_t* __restrict__ out_valid, unsigned long long* __restrict__ null_count) {
  __shared__ unsigned long long s_keys[kSmem ? kIsInSmemSlots : 1];
  if (kSmem) {
    // the column streams through L1 and would keep evicting a table that lives in global memory
    for (unsigned i = threadIdx.x; i <= (unsigned)t.mask; i += kHThreads) s_keys[i] = t.keys[i];
    __syncthreads();
  }
  const bool set_has_null = t.special[1] != kNoRow;
  const bool set_has_ones = t.special[0] != kNoRow;
  auto member = [&](unsigned long long key) -> bool {
    if (key == kEmptyKey) return set_has_ones;
    unsigned long long h = mix64(key) & t.mask;
    while (true) {
      const unsigned long long cur = kSmem ? s_keys[h] : t.keys[h];
      if (cur == key) return true;
      if (cur == kEmptyKey) return false;
      h = (h + 1) & t.mask;
    }
  };
  const int lane = threadIdx.x & 31;
  const int64_t n_words = (n + 31) >> 5;
  const int64_t warp0 = ((int64_t)blockIdx.x * kHThreads + threadIdx.x) >> 5;
  const int64_t warps = ((int64_t)gridDim.x * kHThreads) >> 5;
  unsigned long long nulls = 0;
  for (int64_t w0 = warp0 * kIsInUnroll; w0 < n_words; w0 += warps * kIsInUnroll) {
    V x[kIsInUnroll];
    bool have[kIsInUnroll], isnull[kIsInUnroll];
#pragma unroll
    for (int k = 0; k < kIsInUnroll; ++k) {   // independent loads first: 4 rows in flight per lane
      const int64_t i = ((w0 + k) << 5) + lane;
      have[k] = i < n;
      isnull[k] = have[k] && valid && !bit_is_set(valid, off + i);
      x[k] = (have[k] && !isnull[k]) ? __ldcs(vals + off + i) : V(0);
    }
#pragma unroll
    for (int k = 0; k < kIsInUnroll; ++k) {
      const int64_t w = w0 + k;
      if (w >= n_words) break;   // warp-uniform
      bool d = false, v = true;
      if (have[k]) {
        if (isnull[k]) {
          if (null_behavior == AG_NULL_MATCH && set_has_null) d = true;
          else if (null_behavior == AG_NULL_SKIP || (null_behavior == AG_NULL_MATCH && !set_has_null)) d = false;
          else v = false;
        } else if (member((unsigned long long)x[k])) {
          d = true;
        } else if (null_behavior == AG_NULL_INCONCLUSIVE && set_has_null) {
          v = false;
        }
      }
      const uint32_t dbits = __ballot_sync(0xffffffffu, d);
      const uint32_t vbits = __ballot_sync(0xffffffffu, v);
      if (lane == 0) {
        const int64_t rem = n - (w << 5);
        const uint32_t m = rem >= 32 ? 0xffffffffu : bit_range_mask(0, (int)rem);
        bitmap_store32_masked(out_data + w, dbits, m);
        if (out_valid) bitmap_store32_masked(out_valid + w, vbits, m);
        nulls += __popc(~vbits & m);
      }
    }
  }
  if (lane == 0 && nulls && null_count) atomicAdd(null_count, nulls);
}

// unique: row i is kept iff it is the first row of its key (or the first null row)
template <typename V>
__global__ void __launch_bounds__(kHThreads)
unique_mark_kernel(const V* __restrict__ vals, const uint8_t* __restrict__ valid, int64_t off, int64_t n, const HashTable t,
                   uint32_t* __restrict__ keep) {
  const int lane = threadIdx.x & 31;
  const int64_t n_words = (n + 31) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * kHThreads + threadIdx.x) >> 5; w < n_words; w += ((int64_t)gridDim.x * kHThreads) >> 5) {
    const int64_t i = (w << 5) + lane;
    bool k = false;
    if (i < n) {
      if (valid && !bit_is_set(valid, off + i)) k = t.special[1] == (unsigned)i;
      else k = table_first_row(t, raw_key(vals + off, i)) == (unsigned)i;
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, k);
    if (lane == 0) keep[w] = bits;
  }
}

struct TableMem {
  void* block = nullptr;
  HashTable t{};
  ag_status alloc(int64_t entries, cudaStream_t st) {
    unsigned long long slots = 1024;
    while (slots < (unsigned long long)entries * 2) slots <<= 1;
    const size_t kb = slots * 8, fb = slots * 4;
    AG_TRY(dev_alloc_async(&block, kb + fb + 64, st));
    t.keys = reinterpret_cast<unsigned long long*>(block);
    t.first = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(block) + kb);
    t.special = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(block) + kb + fb);
    t.mask = slots - 1;
    AG_CUDA_TRY(cudaMemsetAsync(block, 0xff, kb + fb + 64, st));   // every key free, every first row = none
    return AG_OK;
  }
  void release(cudaStream_t st) { if (block) cudaFreeAsync(block, st); block = nullptr; }
};

template <typename V>
static ag_status is_in_t(const void* vals, const uint8_t* valid, int64_t off, int64_t n, const void* set, const uint8_t* set_valid, int64_t set_off,
                         int64_t set_n, int null_behavior, uint8_t* out_data, uint8_t* out_valid, int64_t* d_null_count, cudaStream_t st) {
  TableMem tm;
  AG_TRY(tm.alloc(set_n, st));
  ag_status rc = AG_OK;
  do {
    if (set_n > 0) {
      hash_insert_kernel<V><<<grid_for(set_n, kHThreads * 4, 8), kHThreads, 0, st>>>(reinterpret_cast<const V*>(set), set_valid, set_off, set_n, tm.t);
      if ((rc = check_launch("hash_insert_kernel")) != AG_OK) break;
    }
    if (d_null_count && cudaMemsetAsync(d_null_count, 0, 8, st) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "memset", __FILE__, __LINE__); break; }
    if (tm.t.mask < (unsigned long long)kIsInSmemSlots)
      is_in_kernel<V, true><<<grid_for(n, kHThreads * 4, 6), kHThreads, 0, st>>>(reinterpret_cast<const V*>(vals), valid, off, n, tm.t, null_behavior,
                                                                                 reinterpret_cast<uint32_t*>(out_data), reinterpret_cast<uint32_t*>(out_valid),
                                                                                 reinterpret_cast<unsigned long long*>(d_null_count));
    else
      is_in_kernel<V, false><<<grid_for(n, kHThreads * 4, 8), kHThreads, 0, st>>>(reinterpret_cast<const V*>(vals), valid, off, n, tm.t, null_behavior,
                                                                                  reinterpret_cast<uint32_t*>(out_data), reinterpret_cast<uint32_t*>(out_valid),
                                                                                  reinterpret_cast<unsigned long long*>(d_null_count));
    rc = check_launch("is_in_kernel");
  } while (0);
  tm.release(st);
  return rc;
}

ag_status is_in_dev(int bit_width, const void* vals, const uint8_t* valid, int64_t off, int64_t n, const void* set, const uint8_t* set_valid,
                    int64_t set_off, int64_t set_n, int null_behavior, uint8_t* out_data, uint8_t* out_valid, int64_t* d_null_count, cudaStream_t st) {
  if (n < 0 || set_n < 0 || off < 0 || set_off < 0) AG_FAIL(AG_ERR_INVALID, "is_in: negative length or offset");
  if (null_behavior < 0 || null_behavior > 3) AG_FAIL(AG_ERR_INVALID, "is_in: bad null matching behaviour %d", null_behavior);
  if (n == 0) return AG_OK;
  if (!vals || !out_data || (set_n > 0 && !set)) AG_FAIL(AG_ERR_INVALID, "is_in: NULL values / value set / output");
  if ((reinterpret_cast<uintptr_t>(out_data) & 3) || (reinterpret_cast<uintptr_t>(out_valid) & 3)) AG_FAIL(AG_ERR_INVALID, "is_in: output bitmaps must be 4-byte aligned");
  if (set_n >= (1ll << 31)) AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "is_in: value set too large");
  switch (bit_width) {
    case 8: return is_in_t<uint8_t>(vals, valid, off, n, set, set_valid, set_off, set_n, null_behavior, out_data, out_valid, d_null_count, st);
    case 16: return is_in_t<uint16_t>(vals, valid, off, n, set, set_valid, set_off, set_n, null_behavior, out_data, out_valid, d_null_count, st);
    case 32: return is_in_t<uint32_t>(vals, valid, off, n, set, set_valid, set_off, set_n, null_behavior, out_data, out_valid, d_null_count, st);
    case 64: return is_in_t<unsigned long long>(vals, valid, off, n, set, set_valid, set_off, set_n, null_behavior, out_data, out_valid, d_null_count, st);
    default: AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "is_in: fixed-width values of 1/2/4/8 bytes only (got %d bits)", bit_width);
  }
}

template <typename V>
static ag_status unique_t(int bit_width, const void* vals, const uint8_t* valid, int64_t off, int64_t n, void* out, uint8_t* out_valid,
                          int64_t capacity, int64_t* d_out_len, cudaStream_t st) {
  TableMem tm;
  AG_TRY(tm.alloc(n, st));
  uint32_t* keep = nullptr;
  ag_status rc = dev_alloc_async((void**)&keep, (size_t)((n + 31) / 32) * 4 + 64, st);
  if (rc == AG_OK) {
    do {
      hash_insert_kernel<V><<<grid_for(n, kHThreads * 4, 8), kHThreads, 0, st>>>(reinterpret_cast<const V*>(vals), valid, off, n, tm.t);
      if ((rc = check_launch("hash_insert_kernel")) != AG_OK) break;
      unique_mark_kernel<V><<<grid_for(n, kHThreads * 4, 8), kHThreads, 0, st>>>(reinterpret_cast<const V*>(vals), valid, off, n, tm.t, keep);
      if ((rc = check_launch("unique_mark_kernel")) != AG_OK) break;
      if (out_valid) {
        if (cudaMemsetAsync(out_valid, 0, (size_t)((capacity + 31) / 32) * 4, st) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "memset", __FILE__, __LINE__); break; }
      }
      rc = filter_primitive_dev(bit_width, vals, valid, off, reinterpret_cast<const uint8_t*>(keep), nullptr, 0, n, AG_DROP_NULLS, out, out_valid,
                                capacity, d_out_len, st);
    } while (0);
    cudaFreeAsync(keep, st);
  }
  tm.release(st);
  return rc;
}

ag_status unique_dev(int bit_width, const void* vals, const uint8_t* valid, int64_t off, int64_t n, void* out, uint8_t* out_valid, int64_t capacity,
                     int64_t* d_out_len, cudaStream_t st) {
  if (n < 0 || off < 0 || capacity < 0) AG_FAIL(AG_ERR_INVALID, "unique: negative length or offset");
  if (!d_out_len) AG_FAIL(AG_ERR_INVALID, "unique: NULL length word");
  if (n == 0) { AG_CUDA_TRY(cudaMemsetAsync(d_out_len, 0, sizeof(int64_t), st)); return AG_OK; }
  if (!vals || (!out && capacity > 0)) AG_FAIL(AG_ERR_INVALID, "unique: NULL values/output");
  if (valid && !out_valid && capacity > 0) AG_FAIL(AG_ERR_INVALID, "unique: an input with a validity bitmap needs an output validity bitmap");
  if (n >= (1ll << 31)) AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "unique: more than 2^31-1 rows per call");
  switch (bit_width) {
    case 8: return unique_t<uint8_t>(bit_width, vals, valid, off, n, out, out_valid, capacity, d_out_len, st);
    case 16: return unique_t<uint16_t>(bit_width, vals, valid, off, n, out, out_valid, capacity, d_out_len, st);
    case 32: return unique_t<uint32_t>(bit_width, vals, valid, off, n, out, out_valid, capacity, d_out_len, st);
    case 64: return unique_t<unsigned long long>(bit_width, vals, valid, off, n, out, out_valid, capacity, d_out_len, st);
    default: AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "unique: fixed-width values of 1/2/4/8 bytes only (got %d bits)", bit_width);
  }
}

}  // namespace ag

using namespace ag;

extern "C" ag_status ag_is_in_dev(int bit_width, const void* d_vals, const uint8_t* d_valid, int64_t offset, int64_t n, const void* d_set,
                                  const uint8_t* d_set_valid, int64_t set_offset, int64_t set_n, int null_behavior, uint8_t* d_out_data,
                                  uint8_t* d_out_valid, int64_t* d_null_count, ag_stream_t s) {
  AG_TRY(ensure_init());
  return is_in_dev(bit_width, d_vals, d_valid, offset, n, d_set, d_set_valid, set_offset, set_n, null_behavior, d_out_data, d_out_valid, d_null_count,
                   resolve_stream(s));
}

extern "C" ag_status ag_unique_dev(int bit_width, const void* d_vals, const uint8_t* d_valid, int64_t offset, int64_t n, void* d_out,
                                   uint8_t* d_out_valid, int64_t capacity, int64_t* d_out_len, ag_stream_t s) {
  AG_TRY(ensure_init());
  return unique_dev(bit_width, d_vals, d_valid, offset, n, d_out, d_out_valid, capacity, d_out_len, resolve_stream(s));
}
