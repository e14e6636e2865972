// filter.cu — stream compaction (Filter, GetTakeIndices) on sm_100a.
//
// Replaces PrimitiveFilter (arrow/compute/internal/kernels/vector_selection.go:449-520) =
// getFilterOutputSize (:57-81) + primitiveFilterImpl (:267-395) + filterWriter (:397-421), and
// GetTakeIndices (:102-236).  Row rule restated from the reference's default branch
// (:321-392; the block-level fast paths :303-320 are shortcuts of it):
//     selected  = mask ∧ maskValid            -> emit value, validity = valuesValid (or 1)
//     EmitNulls ∧ ¬maskValid                  -> emit null  (value 0 :417-421, validity 0)
//     anything else                           -> dropped
// Output is stable (row order preserved), so it is bit-identical to the reference's.
//
// Roofline: HBM.  Algorithmic bytes/row for int64 at selectivity s: 8 + 1/8 + 8s.  Values are
// fetched only for emitted rows (predicated 8-byte loads; DRAM moves the touched 32-byte
// sectors), so at low selectivity the kernel can move FEWER bytes than that figure.
//
// Single pass, WARP-granular decoupled look-back (no block barriers, no shared memory):
//   * a tile is 1024 rows = one warp, one 32-bit mask word per lane; tiles are claimed with an
//     atomic ticket (a tile only ever waits on tiles that are already running) and the next
//     ticket is requested while the current tile's loads are in flight;
//   * popcount per lane -> warp scan -> tile aggregate published in a 64-bit status word
//     (2 flag bits + 62-bit count, one store => a reader never sees a flag without its value);
//   * the warp looks back 32 tiles at a time until it meets an inclusive prefix; the other
//     warps of the SM are in their load/store phases meanwhile, which is what hides the chain;
//   * compaction is warp-cooperative: for each of the 32 words, lane j owns row 32k+j, so loads
//     are coalesced; it stores to out[base_k + rank], rank = popc of the lower set bits, so the
//     writes of one step are contiguous.  Sparse tiles load only the selected rows (DRAM moves
//     the touched sectors); dense tiles load all 32 x 256 B rows;
//   * output validity bits are compacted with __reduce_or_sync and merged into pre-zeroed
//     words with at most two atomicOr per step.
#include "common.cuh"

#include <stdlib.h>
#include <type_traits>

namespace ag {

constexpr int kFThreads = 256;
constexpr int kFWarps = kFThreads / 32;
constexpr int kFTileRows = 1024;  // one WARP tile: 32 lanes x one 32-bit mask word
constexpr int kFBlocksPerSM = 4;

constexpr unsigned long long kFlagShift = 62;
constexpr unsigned long long kFlagAgg = 1ull << kFlagShift;
constexpr unsigned long long kFlagIncl = 2ull << kFlagShift;
constexpr unsigned long long kValMask = (1ull << kFlagShift) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

struct FilterParams {
  const void* vals;          // element 0 of the values buffer (NULL in index mode)
  const uint8_t* vvalid;     // values validity (may be NULL)
  int64_t voff;
  const uint8_t* mask;
  const uint8_t* mvalid;     // mask validity (may be NULL)
  int64_t moff;
  int64_t n;
  int emit_nulls;
  void* out;
  uint32_t* out_valid;       // 4-byte aligned, pre-zeroed for ceil(capacity/32) words (may be NULL)
  int64_t capacity;          // rows the output buffers can hold
  unsigned long long* status;  // [0] = tile ticket, [1..] = tile status words
  long long* out_len;
  int64_t n_tiles;
  int dense_threshold;       // tiles emitting at least this many rows load their 1024 values coalesced
};

// Whole warp.  Returns the exclusive prefix of `tile` (rows emitted by all earlier tiles) and
// publishes this tile's inclusive prefix.  One 64-bit word carries flag + count, so a reader
// never sees a flag without its value.
__device__ __forceinline__ unsigned long long lookback(unsigned long long* st, int64_t tile, unsigned long long total, int lane) {
  if (tile == 0) {
    if (lane == 0) st_status(st, kFlagIncl | total);
    return 0;
  }
  if (lane == 0) st_status(st + tile, kFlagAgg | total);
  unsigned long long running = 0;
  int64_t look = tile - 1;
  while (true) {
    const int64_t idx = look - lane;
    unsigned long long s;
    unsigned flag;
    do {
      s = (idx >= 0) ? ld_status(st + idx) : kFlagIncl;  // before tile 0: inclusive prefix 0
      flag = (unsigned)(s >> kFlagShift);
    } while (__any_sync(0xffffffffu, flag == 0));
    const unsigned incl = __ballot_sync(0xffffffffu, flag == 2);
    unsigned long long v = s & kValMask;
    if (incl) {
      const int first = __ffs(incl) - 1;  // closest tile that already knows its inclusive prefix
      if (lane > first) v = 0;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    running += v;
    if (incl) break;
    look -= 32;
  }
  if (lane == 0) st_status(st + tile, kFlagIncl | (running + total));
  return running;
}

// Warp-wide exclusive scan of the per-lane counts + look-back.  Returns the global output slot
// of this lane's first emitted row; *tile_total = rows the warp tile emits.
__device__ __forceinline__ unsigned long long warp_tile_scan(unsigned cnt, unsigned long long* status, int64_t tile, int64_t n_tiles,
                                                             long long* out_len, unsigned* tile_total_out) {
  const int lane = threadIdx.x & 31;
  unsigned incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
  const unsigned long long excl = lookback(status, tile, total, lane);
  if (lane == 0 && tile == n_tiles - 1) *out_len = (long long)(excl + total);
  *tile_total_out = total;
  return excl + (incl - cnt);
}

// Warp tiles are claimed with an atomic ticket (a tile only waits on tiles that are already
// running => forward progress without assuming co-residency); the ticket for the NEXT tile is
// requested while the current tile's loads are in flight.
__device__ __forceinline__ long long claim_tile(unsigned long long* ticket, int lane) {
  unsigned long long t = 0;
  if (lane == 0) t = atomicAdd(ticket, 1ull);
  return (long long)__shfl_sync(0xffffffffu, t, 0);
}

// kMode 0: copy values of type V.  kMode 1: write row indices as V (GetTakeIndices).
template <typename V, int kMode, bool kValidity>
__global__ void __launch_bounds__(kFThreads)
filter_kernel(const FilterParams p) {
  const int lane = threadIdx.x & 31;
  const V* __restrict__ vals = reinterpret_cast<const V*>(p.vals) + (kMode == 0 ? p.voff : 0);
  V* __restrict__ out = reinterpret_cast<V*>(p.out);
  const int64_t m_lo = p.moff >> 3, m_hi = (p.moff + p.n + 7) >> 3;
  const long long cap_words = (p.capacity + 31) >> 5;

  long long tile = claim_tile(p.status, lane);
  while (tile < p.n_tiles) {
    const int64_t wrow0 = tile * kFTileRows;
    const int64_t row0 = wrow0 + (int64_t)lane * 32;  // first row of this lane's mask word
    uint32_t sel = 0, nul = 0;
    if (row0 < p.n) {
      const int64_t rem = p.n - row0;
      const uint32_t range = rem >= 32 ? 0xffffffffu : bit_range_mask(0, (int)rem);
      const uint32_t m = bitmap_load32(p.mask, p.moff + row0, m_lo, m_hi);
      uint32_t mv = 0xffffffffu;
      if (p.mvalid) mv = bitmap_load32(p.mvalid, p.moff + row0, m_lo, m_hi);
      sel = m & mv & range;
      if (p.emit_nulls) nul = ~mv & range;
    }
    const long long next_tile = claim_tile(p.status, lane);
    const uint32_t emit = sel | nul;
    unsigned tile_total;
    const unsigned long long my_base = warp_tile_scan(__popc(emit), p.status + 1, tile, p.n_tiles, p.out_len, &tile_total);

    if (tile_total != 0) {
      const bool dense = kMode == 0 && (int)tile_total >= p.dense_threshold && wrow0 + kFTileRows <= p.n;
#pragma unroll
      for (int kb = 0; kb < 32; kb += 8) {
        // phase 1: 8 independent loads in flight per lane (all 8 x 256 B rows when the tile is dense,
        // only the selected rows' sectors when it is sparse)
        V v[8];
        uint32_t w_emit[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          w_emit[u] = __shfl_sync(0xffffffffu, emit, kb + u);
          const uint32_t w_sel = __shfl_sync(0xffffffffu, sel, kb + u);
          v[u] = V(0);
          if (kMode == 0) {
            if (dense || ((w_sel >> lane) & 1)) v[u] = vals[wrow0 + (kb + u) * 32 + lane];
            if (!((w_sel >> lane) & 1)) v[u] = V(0);
          } else {
            if ((w_sel >> lane) & 1) v[u] = (V)(wrow0 + (kb + u) * 32 + lane);
          }
        }
        // phase 2: compacting stores
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (w_emit[u] == 0) continue;  // warp-uniform
          const unsigned long long b = __shfl_sync(0xffffffffu, my_base, kb + u);
          const bool e = (w_emit[u] >> lane) & 1;
          const unsigned rank = __popc(w_emit[u] & ((1u << lane) - 1u));
          const unsigned long long pos = b + rank;
          if (e && (long long)pos < p.capacity) out[pos] = v[u];
          if (kValidity) {
            const uint32_t w_sel = __shfl_sync(0xffffffffu, sel, kb + u);
            const int64_t row = wrow0 + (kb + u) * 32 + lane;
            uint32_t vb = 0;
            if (e && ((w_sel >> lane) & 1)) vb = (kMode == 0 && p.vvalid) ? (uint32_t)bit_is_set(p.vvalid, p.voff + row) : 1u;
            const uint32_t pattern = __reduce_or_sync(0xffffffffu, vb << rank);  // rank < 32 whenever vb != 0
            if (lane == 0 && pattern) {
              const unsigned sh = (unsigned)(b & 31);
              const unsigned long long wi = b >> 5;
              if ((long long)wi < cap_words) atomicOr(p.out_valid + wi, pattern << sh);
              if (sh && (pattern >> (32 - sh)) && (long long)(wi + 1) < cap_words) atomicOr(p.out_valid + wi + 1, pattern >> (32 - sh));
            }
          }
        }
      }
    }
    tile = next_tile;
  }
}

static int filter_dense_threshold() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("AG_FILTER_DENSE_THRESHOLD");
    cached = e ? atoi(e) : 96;  // rows emitted per 1024-row tile above which coalesced loads win
  }
  return cached;
}

template <typename V, int kMode>
static ag_status launch_filter_t(FilterParams& p, cudaStream_t st) {
  Workspace* ws;
  AG_TRY(get_workspace(st, &ws));
  p.n_tiles = (p.n + kFTileRows - 1) / kFTileRows;
  p.dense_threshold = filter_dense_threshold();
  AG_TRY(ensure_tile_status(ws, (size_t)p.n_tiles + 1, st));
  p.status = ws->tile_status;
  AG_CUDA_TRY(cudaMemsetAsync(p.status, 0, ((size_t)p.n_tiles + 1) * sizeof(unsigned long long), st));
  if (p.out_valid) AG_CUDA_TRY(cudaMemsetAsync(p.out_valid, 0, (size_t)((p.capacity + 31) >> 5) * 4, st));
  const int64_t blocks_needed = (p.n_tiles + kFWarps - 1) / kFWarps;
  if (p.out_valid) filter_kernel<V, kMode, true><<<grid_one_wave(filter_kernel<V, kMode, true>, kFThreads, blocks_needed), kFThreads, 0, st>>>(p);
  else filter_kernel<V, kMode, false><<<grid_one_wave(filter_kernel<V, kMode, false>, kFThreads, blocks_needed), kFThreads, 0, st>>>(p);
  return check_launch("filter_kernel");
}

static ag_status check_filter_args(const FilterParams& p, const char* who) {
  if (p.n < 0 || p.moff < 0 || p.voff < 0 || p.capacity < 0) AG_FAIL(AG_ERR_INVALID, "%s: negative length or offset", who);
  if (!p.out_len) AG_FAIL(AG_ERR_INVALID, "%s: NULL out_len", who);
  if (p.n > 0 && !p.mask) AG_FAIL(AG_ERR_INVALID, "%s: NULL mask", who);
  if (p.out_valid && (reinterpret_cast<uintptr_t>(p.out_valid) & 3)) AG_FAIL(AG_ERR_INVALID, "%s: out_valid must be 4-byte aligned", who);
  return AG_OK;
}

ag_status filter_primitive_dev(int bit_width, const void* vals, const uint8_t* vvalid, int64_t voff,
                               const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n,
                               int null_selection, void* out, uint8_t* out_valid, int64_t capacity,
                               int64_t* d_out_len, cudaStream_t st) {
  FilterParams p{};
  p.vals = vals; p.vvalid = vvalid; p.voff = voff; p.mask = mask; p.mvalid = mvalid; p.moff = moff; p.n = n;
  p.emit_nulls = (null_selection == AG_EMIT_NULLS) && mvalid != nullptr;
  p.out = out; p.out_valid = reinterpret_cast<uint32_t*>(out_valid); p.capacity = capacity;
  p.out_len = reinterpret_cast<long long*>(d_out_len);
  AG_TRY(check_filter_args(p, "filter"));
  if (null_selection != AG_DROP_NULLS && null_selection != AG_EMIT_NULLS) AG_FAIL(AG_ERR_INVALID, "filter: bad null_selection %d", null_selection);
  if ((vvalid || p.emit_nulls) && !out_valid && capacity > 0)
    AG_FAIL(AG_ERR_INVALID, "filter: the output can contain nulls but no output validity buffer was given (vector_selection.go:473)");
  if (n == 0) { AG_CUDA_TRY(cudaMemsetAsync(d_out_len, 0, sizeof(int64_t), st)); return AG_OK; }
  if (!vals || (!out && capacity > 0)) AG_FAIL(AG_ERR_INVALID, "filter: NULL values/output");
  const uintptr_t am = (uintptr_t)(bit_width / 8 - 1);
  switch (bit_width) {
    case 8: case 16: case 32: case 64: break;
    case 1: AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "filter: boolean values are not implemented (see DESIGN.md: boolFilterWriter.WriteValue, vector_selection.go:433-436)");
    default: AG_FAIL(AG_ERR_TYPE, "filter: invalid values bit width %d", bit_width);
  }
  if (((uintptr_t)vals & am) || ((uintptr_t)out & am)) AG_FAIL(AG_ERR_INVALID, "filter: buffers not aligned to the element width");
  switch (bit_width) {
    case 8: return launch_filter_t<uint8_t, 0>(p, st);
    case 16: return launch_filter_t<uint16_t, 0>(p, st);
    case 32: return launch_filter_t<uint32_t, 0>(p, st);
    default: return launch_filter_t<unsigned long long, 0>(p, st);
  }
}

ag_status take_indices_dev(int index_width, const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n,
                           int null_selection, void* out_idx, uint8_t* out_valid, int64_t capacity,
                           int64_t* d_out_len, cudaStream_t st) {
  FilterParams p{};
  p.mask = mask; p.mvalid = mvalid; p.moff = moff; p.n = n;
  p.emit_nulls = (null_selection == AG_EMIT_NULLS) && mvalid != nullptr;
  p.out = out_idx; p.out_valid = reinterpret_cast<uint32_t*>(out_valid); p.capacity = capacity;
  p.out_len = reinterpret_cast<long long*>(d_out_len);
  AG_TRY(check_filter_args(p, "take_indices"));
  if (n == 0) { AG_CUDA_TRY(cudaMemsetAsync(d_out_len, 0, sizeof(int64_t), st)); return AG_OK; }
  if (index_width == 16) {
    if (n >= 65535) AG_FAIL(AG_ERR_INVALID, "take_indices: uint16 indices need n < 65535 (vector_selection.go:229-231)");
    return launch_filter_t<uint16_t, 1>(p, st);
  }
  if (index_width == 32) {
    if (n >= 4294967295ll) AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "take_indices: filter length exceeds UINT32_MAX (vector_selection.go:233-235)");
    return launch_filter_t<uint32_t, 1>(p, st);
  }
  AG_FAIL(AG_ERR_TYPE, "take_indices: index width must be 16 or 32");
}

// ---------------------------------------------------------------- fused compare + filter ----
// Greater/…(values, scalar) -> Filter in ONE pass over `values` (config 3 of BASELINE.json):
// no intermediate mask, 8 + 8s bytes/row.  A warp keeps its 1024 rows in registers (lane holds
// rows 32k+lane, k = 0..31 — the layout both the ballot and the compaction loop want), so the
// values are read from HBM exactly once: 32 coalesced 256-byte loads per warp, all issued
// before the first vote.  Same result as compare_dev + filter_primitive_dev.
struct FCmpEq { template <typename T> static __device__ __forceinline__ bool apply(T a, T b) { return a == b; } };
struct FCmpNe { template <typename T> static __device__ __forceinline__ bool apply(T a, T b) { return a != b; } };
struct FCmpGt { template <typename T> static __device__ __forceinline__ bool apply(T a, T b) { return a > b; } };
struct FCmpGe { template <typename T> static __device__ __forceinline__ bool apply(T a, T b) { return a >= b; } };
struct FCmpLt { template <typename T> static __device__ __forceinline__ bool apply(T a, T b) { return b > a; } };
struct FCmpLe { template <typename T> static __device__ __forceinline__ bool apply(T a, T b) { return b >= a; } };

template <typename T, typename Cmp>
__global__ void __launch_bounds__(kFThreads)
fused_cmp_filter_kernel(const T* __restrict__ vals, T scalar, int64_t n, T* __restrict__ out, int64_t capacity,
                        unsigned long long* status, long long* out_len, int64_t n_tiles) {
  const int lane = threadIdx.x & 31;
  long long tile = claim_tile(status, lane);
  while (tile < n_tiles) {
    const int64_t wrow0 = tile * kFTileRows;
    T v[32];
    if (wrow0 + kFTileRows <= n) {
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __ldcs(vals + wrow0 + k * 32 + lane);
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int64_t row = wrow0 + k * 32 + lane;
        v[k] = (row < n) ? __ldcs(vals + row) : scalar;
      }
    }
    const long long next_tile = claim_tile(status, lane);  // latency hidden behind the loads above
    __syncwarp();                                          // keep every load ahead of the first vote
    uint32_t emit = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int64_t row = wrow0 + k * 32 + lane;
      const uint32_t bits = __ballot_sync(0xffffffffu, row < n && Cmp::template apply<T>(v[k], scalar));
      if (lane == k) emit = bits;
    }
    unsigned tile_total;
    const unsigned long long my_base = warp_tile_scan(__popc(emit), status + 1, tile, n_tiles, out_len, &tile_total);
    if (tile_total != 0) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const uint32_t w_emit = __shfl_sync(0xffffffffu, emit, k);
        const unsigned long long b = __shfl_sync(0xffffffffu, my_base, k);
        const unsigned rank = __popc(w_emit & ((1u << lane) - 1u));
        const unsigned long long pos = b + rank;
        if (((w_emit >> lane) & 1) && (long long)pos < capacity) out[pos] = v[k];
      }
    }
    tile = next_tile;
  }
}

template <typename T, typename Cmp>
static ag_status launch_fused_t(const void* vals, const void* scalar_host, int64_t n, void* out, int64_t capacity,
                                int64_t* d_out_len, cudaStream_t st) {
  Workspace* ws;
  AG_TRY(get_workspace(st, &ws));
  const int64_t n_tiles = (n + kFTileRows - 1) / kFTileRows;
  AG_TRY(ensure_tile_status(ws, (size_t)n_tiles + 1, st));
  AG_CUDA_TRY(cudaMemsetAsync(ws->tile_status, 0, ((size_t)n_tiles + 1) * sizeof(unsigned long long), st));
  const int grid = grid_one_wave(fused_cmp_filter_kernel<T, Cmp>, kFThreads, (n_tiles + kFWarps - 1) / kFWarps);
  fused_cmp_filter_kernel<T, Cmp><<<grid, kFThreads, 0, st>>>((const T*)vals, *(const T*)scalar_host, n, (T*)out, capacity,
                                                             ws->tile_status, (long long*)d_out_len, n_tiles);
  return check_launch("fused_cmp_filter_kernel");
}

template <typename T>
static ag_status launch_fused_cmp(int cmp, const void* vals, const void* scalar_host, int64_t n, void* out, int64_t capacity,
                                  int64_t* d_out_len, cudaStream_t st) {
  switch (cmp) {
    case AG_CMP_EQ: return launch_fused_t<T, FCmpEq>(vals, scalar_host, n, out, capacity, d_out_len, st);
    case AG_CMP_NE: return launch_fused_t<T, FCmpNe>(vals, scalar_host, n, out, capacity, d_out_len, st);
    case AG_CMP_GT: return launch_fused_t<T, FCmpGt>(vals, scalar_host, n, out, capacity, d_out_len, st);
    case AG_CMP_GE: return launch_fused_t<T, FCmpGe>(vals, scalar_host, n, out, capacity, d_out_len, st);
    case AG_CMP_LT: return launch_fused_t<T, FCmpLt>(vals, scalar_host, n, out, capacity, d_out_len, st);
    case AG_CMP_LE: return launch_fused_t<T, FCmpLe>(vals, scalar_host, n, out, capacity, d_out_len, st);
    default: AG_FAIL(AG_ERR_INVALID, "filter_compare: bad operator %d", cmp);
  }
}

ag_status filter_compare_scalar_dev(int type, int cmp, const void* vals, const void* scalar_host, int64_t n,
                                    void* out, int64_t capacity, int64_t* d_out_len, cudaStream_t st) {
  if (n < 0 || capacity < 0) AG_FAIL(AG_ERR_INVALID, "filter_compare: negative length");
  if (!d_out_len || !scalar_host) AG_FAIL(AG_ERR_INVALID, "filter_compare: NULL argument");
  if (n == 0) { AG_CUDA_TRY(cudaMemsetAsync(d_out_len, 0, sizeof(int64_t), st)); return AG_OK; }
  if (!vals || (!out && capacity > 0)) AG_FAIL(AG_ERR_INVALID, "filter_compare: NULL values/output");
  switch (type) {
    case AG_TYPE_INT32: return launch_fused_cmp<int32_t>(cmp, vals, scalar_host, n, out, capacity, d_out_len, st);
    case AG_TYPE_UINT32: return launch_fused_cmp<uint32_t>(cmp, vals, scalar_host, n, out, capacity, d_out_len, st);
    case AG_TYPE_INT64: return launch_fused_cmp<long long>(cmp, vals, scalar_host, n, out, capacity, d_out_len, st);
    case AG_TYPE_UINT64: return launch_fused_cmp<unsigned long long>(cmp, vals, scalar_host, n, out, capacity, d_out_len, st);
    case AG_TYPE_FLOAT32: return launch_fused_cmp<float>(cmp, vals, scalar_host, n, out, capacity, d_out_len, st);
    case AG_TYPE_FLOAT64: return launch_fused_cmp<double>(cmp, vals, scalar_host, n, out, capacity, d_out_len, st);
    default: AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "filter_compare: fused path covers 32/64-bit types; use compare + filter for type id %d", type);
  }
}

// ---------------------------------------------------------------- output size ------
// getFilterOutputSize (vector_selection.go:57-81): popcount(mask ∧ valid) | popcount(mask ∨ ¬valid)
__global__ void __launch_bounds__(kFThreads)
filter_count_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ mvalid, int64_t moff, int64_t n,
                    int emit_nulls, unsigned long long* __restrict__ count) {
  const int64_t n_words = (n + 31) >> 5;
  const int64_t lo = moff >> 3, hi = (moff + n + 7) >> 3;
  const int64_t stride = (int64_t)gridDim.x * kFThreads;
  unsigned long long c = 0;
  for (int64_t w = (int64_t)blockIdx.x * kFThreads + threadIdx.x; w < n_words; w += stride) {
    const int64_t rem = n - (w << 5);
    const uint32_t range = rem >= 32 ? 0xffffffffu : bit_range_mask(0, (int)rem);
    const uint32_t m = bitmap_load32(mask, moff + (w << 5), lo, hi);
    uint32_t mv = 0xffffffffu;
    if (mvalid) mv = bitmap_load32(mvalid, moff + (w << 5), lo, hi);
    const uint32_t v = emit_nulls ? (m | ~mv) : (m & mv);
    c += __popc(v & range);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
  __shared__ unsigned long long sm[kFWarps];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < kFWarps; ++i) t += sm[i];
    if (t) atomicAdd(count, t);
  }
}

ag_status filter_output_size_dev(const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n,
                                 int null_selection, int64_t* d_out_len, cudaStream_t st) {
  if (n < 0 || moff < 0) AG_FAIL(AG_ERR_INVALID, "filter_output_size: negative length or offset");
  if (!d_out_len) AG_FAIL(AG_ERR_INVALID, "filter_output_size: NULL result");
  AG_CUDA_TRY(cudaMemsetAsync(d_out_len, 0, sizeof(int64_t), st));
  if (n == 0) return AG_OK;
  if (!mask) AG_FAIL(AG_ERR_INVALID, "filter_output_size: NULL mask");
  const int grid = grid_for((n + 31) >> 5, kFThreads * 4, 4);
  filter_count_kernel<<<grid, kFThreads, 0, st>>>(mask, mvalid, moff, n, null_selection == AG_EMIT_NULLS, (unsigned long long*)d_out_len);
  return check_launch("filter_count_kernel");
}

}  // namespace ag

using namespace ag;

extern "C" {

ag_status ag_filter_output_size_dev(const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n,
                                    int null_selection, int64_t* d_out_len, ag_stream_t s) {
  AG_TRY(ensure_init());
  return filter_output_size_dev(mask, mvalid, moff, n, null_selection, d_out_len, resolve_stream(s));
}
ag_status ag_filter_primitive_dev(int bit_width, const void* vals, const uint8_t* vvalid, int64_t voff,
                                  const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n,
                                  int null_selection, void* out, uint8_t* out_valid, int64_t out_capacity,
                                  int64_t* d_out_len, ag_stream_t s) {
  AG_TRY(ensure_init());
  return filter_primitive_dev(bit_width, vals, vvalid, voff, mask, mvalid, moff, n, null_selection, out, out_valid,
                              out_capacity, d_out_len, resolve_stream(s));
}
ag_status ag_filter_compare_scalar_dev(int type, int cmp, const void* d_vals, const void* scalar_host, int64_t n,
                                       void* d_out, int64_t out_capacity, int64_t* d_out_len, ag_stream_t s) {
  AG_TRY(ensure_init());
  return filter_compare_scalar_dev(type, cmp, d_vals, scalar_host, n, d_out, out_capacity, d_out_len, resolve_stream(s));
}
ag_status ag_take_indices_dev(int index_width, const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n,
                              int null_selection, void* out_idx, uint8_t* out_valid, int64_t out_capacity,
                              int64_t* d_out_len, ag_stream_t s) {
  AG_TRY(ensure_init());
  return take_indices_dev(index_width, mask, mvalid, moff, n, null_selection, out_idx, out_valid, out_capacity,
                          d_out_len, resolve_stream(s));
}

}  // extern "C"
