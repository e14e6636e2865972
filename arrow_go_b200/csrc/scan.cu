// scan.cu — cumulative_sum / cumulative_sum_checked in ONE pass over HBM.
//
// Replaces cumulativeSumNoNulls / WithNulls (+Checked) of
// arrow/compute/internal/kernels/vector_cumulative.go:228-330 (driver cumulativeSumSpans :332-368,
// state :100-145, checked adders :147-206):
//   out[i] = start + sum of the VALID in[j], j <= i;  a null input gives a null output slot (value
//   left at 0, like the reference's freshly allocated buffer); unless SkipNulls, every slot after the
//   first null is null as well (state.encounteredNull) and nothing accumulates past it; the checked
//   flavour fails with "overflow" at the first running sum that leaves the type's range.
// Chunked input is one logical sequence (cumulativeSumExecChunked :385-410): the running value and
// the encountered-null flag live in a small device-resident state block that every call reads and
// updates, so a chunked column is a chain of launches on one stream with no host round trip.
//
// Roofline: HBM, 2 x sizeof(T) algorithmic bytes per row (read once, write once).
// Algorithm: single-pass scan with decoupled look-back over 32 KB tiles.
//   * a block = 8 warps; warp w of a tile owns a contiguous 4 KB segment and reads it as 8 rows of
//     32 16-byte vectors (fully coalesced); the prefix inside a segment is 8 warp scans over the
//     vector totals, carried row to row — no shared-memory transpose, no strided access;
//   * tile status, three levels: every tile publishes its AGGREGATE; the last tile of a group of 128
//     publishes the GROUP aggregate (it needs only its own group's tile aggregates, so it never waits
//     on earlier groups); the last tile of a super-group of 32 groups publishes the INCLUSIVE prefix.
//     Tile i gathers <= 127 tile aggregates (four per lane), <= 31 group aggregates (one per lane) and
//     one inclusive prefix, and folds them in a FIXED order — the association of a floating-point sum
//     is a function of (data, n) only, never of timing.  The only serial hand-off happens once per
//     4096 tiles (128 MB of input); a first version with one inclusive prefix per 128 tiles spent
//     ~5 us per hand-off and ran at 0.25 of the roofline, every resident block stalled behind it;
//   * every status word carries its own flag and is written exactly once, so there is no torn
//     flag/value pair to guard against (a 64-bit accumulator is split over 32-bit payloads);
//   * static tile order + cooperative launch (one resident wave): a tile only ever waits on lower
//     tiles, which resident blocks reach in increasing order.
// Integers accumulate in an exact 96-bit value (64-bit wrap + carry count): the wrapped low word is
// the unchecked result, and the checked flavour compares the exact running sum with the type's
// range per element.  Floats accumulate in their own type; results are bit-exact with the
// reference's left-to-right loop whenever every partial sum is exactly representable, and differ by
// rounding of the same order as the reference's own otherwise.
#include "common.cuh"

#include <stdlib.h>
#include <limits>

namespace ag {
namespace {

constexpr int kScThreads = 256;
constexpr int kScWarps = kScThreads / 32;
constexpr int kScRows = 8;                              // 16-byte vectors per lane per tile
constexpr int kScTileBytes = kScThreads * kScRows * 16;  // 32 KB of input
constexpr int kScGroup = 128;                           // tiles per look-back group
constexpr int kScSuper = 32;                            // groups per super-group (the serial chain steps once per 4096 tiles)
constexpr unsigned long long kScFlag = 1ull << 32;      // payload in bits 0..31, flag in bit 32

struct CumsumState {  // device-resident, 32 bytes: include/arrowgpu.h ag_cumsum_state
  unsigned long long lo;  // running value: wrapped 64-bit integer sum, or the float / double bit pattern
  long long hi;           // carries of the exact integer sum (0 for floats)
  long long encountered_null;
  long long null_count;   // null slots written so far (all chunks of the sequence)
};

// ---- accumulators ---------------------------------------------------------------------------
struct AccX { unsigned long long lo; int hi; };  // exact: value = hi * 2^64 + lo

template <typename T> struct IsFp { static constexpr bool v = false; };
template <> struct IsFp<float> { static constexpr bool v = true; };
template <> struct IsFp<double> { static constexpr bool v = true; };

template <typename T, bool kFp = IsFp<T>::v> struct Pol;

template <typename T>
struct Pol<T, false> {  // integers
  using A = AccX;
  static constexpr int K = 3;
  static constexpr bool kSigned = T(-1) < T(0);
  static __device__ __forceinline__ A zero() { return A{0ull, 0}; }
  static __device__ __forceinline__ A add(A a, A b) {
    A r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1 : 0);
    return r;
  }
  static __device__ __forceinline__ A add_elem(A a, T x) {
    A b;
    b.lo = (unsigned long long)(long long)x;  // sign / zero extension to 64 bits
    if (!kSigned) b.lo = (unsigned long long)x;
    b.hi = (kSigned && x < T(0)) ? -1 : 0;
    return add(a, b);
  }
  static __device__ __forceinline__ T value(A a) { return (T)a.lo; }
  static __device__ __forceinline__ bool out_of_range(A a) {
    if (sizeof(T) == 8) {
      if (kSigned) return a.hi != (((long long)a.lo < 0) ? -1 : 0);
      return a.hi != 0;
    }
    // exact value fits 64 bits comfortably for narrower types (|sum| < 2^32 * 2^31)
    const long long v = (long long)a.lo;
    if (a.hi != (v < 0 ? -1 : 0)) return true;
    return v < (long long)std::numeric_limits<T>::lowest() || v > (long long)std::numeric_limits<T>::max();
  }
  static __device__ __forceinline__ A shfl_up(A a, int d) {
    A r;
    r.lo = __shfl_up_sync(0xffffffffu, a.lo, d);
    r.hi = __shfl_up_sync(0xffffffffu, a.hi, d);
    return r;
  }
  static __device__ __forceinline__ A shfl(A a, int src) {
    A r;
    r.lo = __shfl_sync(0xffffffffu, a.lo, src);
    r.hi = __shfl_sync(0xffffffffu, a.hi, src);
    return r;
  }
  static __device__ __forceinline__ A shfl_xor(A a, int m) {
    A r;
    r.lo = __shfl_xor_sync(0xffffffffu, a.lo, m);
    r.hi = __shfl_xor_sync(0xffffffffu, a.hi, m);
    return r;
  }
  static __device__ __forceinline__ void to_words(A a, unsigned (&w)[3]) { w[0] = (unsigned)a.lo; w[1] = (unsigned)(a.lo >> 32); w[2] = (unsigned)a.hi; }
  static __device__ __forceinline__ A from_words(const unsigned (&w)[3]) { return A{(unsigned long long)w[0] | ((unsigned long long)w[1] << 32), (int)w[2]}; }
  static __device__ __forceinline__ A from_state(const CumsumState& s) { return A{s.lo, (int)s.hi}; }
  static __device__ __forceinline__ void to_state(A a, CumsumState* s) { s->lo = a.lo; s->hi = a.hi; }
};

template <typename T>
struct PolWrap {  // unchecked integers: the sum modulo 2^64 is all the result needs
  using A = unsigned long long;
  static constexpr int K = 2;
  static constexpr bool kSigned = T(-1) < T(0);
  static __device__ __forceinline__ A zero() { return 0ull; }
  static __device__ __forceinline__ A add(A a, A b) { return a + b; }
  static __device__ __forceinline__ A add_elem(A a, T x) { return a + (kSigned ? (A)(long long)x : (A)x); }
  static __device__ __forceinline__ T value(A a) { return (T)a; }
  static __device__ __forceinline__ T offset_add(T off, T local) { return (T)((A)off + (A)local); }  // modulo 2^w
  static __device__ __forceinline__ bool out_of_range(A) { return false; }
  static __device__ __forceinline__ A shfl_up(A a, int d) { return __shfl_up_sync(0xffffffffu, a, d); }
  static __device__ __forceinline__ A shfl(A a, int src) { return __shfl_sync(0xffffffffu, a, src); }
  static __device__ __forceinline__ void to_words(A a, unsigned (&w)[3]) { w[0] = (unsigned)a; w[1] = (unsigned)(a >> 32); }
  static __device__ __forceinline__ A from_words(const unsigned (&w)[3]) { return (A)w[0] | ((A)w[1] << 32); }
  static __device__ __forceinline__ A from_state(const CumsumState& s) { return s.lo; }
  static __device__ __forceinline__ void to_state(A a, CumsumState* s) {
    // keep the exact-sum convention of the state block: hi = sign extension of the wrapped low word is
    // NOT the exact carry count, so an unchecked call leaves hi consistent with "value = lo" only
    s->lo = a;
    s->hi = (kSigned && (long long)a < 0) ? -1 : 0;
  }
};

template <typename T, bool kChecked> struct PolSel { using type = Pol<T>; };
template <> struct PolSel<uint8_t, false> { using type = PolWrap<uint8_t>; };
template <> struct PolSel<int8_t, false> { using type = PolWrap<int8_t>; };
template <> struct PolSel<uint16_t, false> { using type = PolWrap<uint16_t>; };
template <> struct PolSel<int16_t, false> { using type = PolWrap<int16_t>; };
template <> struct PolSel<uint32_t, false> { using type = PolWrap<uint32_t>; };
template <> struct PolSel<int32_t, false> { using type = PolWrap<int32_t>; };
template <> struct PolSel<unsigned long long, false> { using type = PolWrap<unsigned long long>; };
template <> struct PolSel<long long, false> { using type = PolWrap<long long>; };

template <typename T>
struct Pol<T, true> {  // float / double: accumulate in the value type
  using A = T;
  static constexpr int K = sizeof(T) == 8 ? 2 : 1;
  static __device__ __forceinline__ A zero() { return T(0); }
  static __device__ __forceinline__ A add(A a, A b) {
    if constexpr (sizeof(T) == 8) return __dadd_rn(a, b); else return __fadd_rn(a, b);
  }
  static __device__ __forceinline__ A add_elem(A a, T x) { return add(a, x); }
  static __device__ __forceinline__ T value(A a) { return a; }
  static __device__ __forceinline__ T offset_add(T off, T local) { return add(off, local); }
  static __device__ __forceinline__ bool out_of_range(A) { return false; }  // checkedAdder default: plain add
  static __device__ __forceinline__ A shfl_up(A a, int d) { return __shfl_up_sync(0xffffffffu, a, d); }
  static __device__ __forceinline__ A shfl(A a, int src) { return __shfl_sync(0xffffffffu, a, src); }
  static __device__ __forceinline__ A shfl_xor(A a, int m) { return __shfl_xor_sync(0xffffffffu, a, m); }
  static __device__ __forceinline__ void to_words(A a, unsigned (&w)[3]) {
    if constexpr (sizeof(T) == 8) { const unsigned long long b = (unsigned long long)__double_as_longlong(a); w[0] = (unsigned)b; w[1] = (unsigned)(b >> 32); }
    else w[0] = __float_as_uint(a);
  }
  static __device__ __forceinline__ A from_words(const unsigned (&w)[3]) {
    if constexpr (sizeof(T) == 8) return __longlong_as_double((long long)((unsigned long long)w[0] | ((unsigned long long)w[1] << 32)));
    else return __uint_as_float(w[0]);
  }
  static __device__ __forceinline__ A from_state(const CumsumState& s) {
    if constexpr (sizeof(T) == 8) return __longlong_as_double((long long)s.lo); else return __uint_as_float((unsigned)s.lo);
  }
  static __device__ __forceinline__ void to_state(A a, CumsumState* s) {
    if constexpr (sizeof(T) == 8) s->lo = (unsigned long long)__double_as_longlong(a); else s->lo = __float_as_uint(a);
    s->hi = 0;
  }
};

__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
__device__ __forceinline__ void st_word(unsigned long long* p, unsigned v) { *reinterpret_cast<volatile unsigned long long*>(p) = kScFlag | v; }

struct CumsumParams {
  const void* in;
  void* out;
  const uint8_t* valid;   // may be NULL
  int64_t voff;
  int64_t n;
  int skip_nulls;
  int checked;
  const long long* first_null;   // device: first null row of this call (n when none); NULL when valid == NULL
  CumsumState* state;
  long long* first_bad;          // may be NULL when !checked
  uint8_t* out_valid;            // may be NULL (then there are no nulls anywhere in the sequence)
  int64_t ooff;
  unsigned long long* agg;       // [n_tiles][K]
  unsigned long long* gagg;      // [n_groups][K]   aggregate of a whole group
  unsigned long long* sincl;     // [n_super][K]    inclusive prefix at the end of a super-group
  int64_t n_tiles;
};

// Wait for K flagged words and return the value they carry.
template <typename P>
__device__ __forceinline__ typename P::A poll_value(const unsigned long long* words) {
  unsigned w[3] = {0u, 0u, 0u};
#pragma unroll
  for (int k = 0; k < P::K; ++k) {
    unsigned long long s;
    do { s = ld_word(words + k); } while (!(s & kScFlag));
    w[k] = (unsigned)s;
  }
  return P::from_words(w);
}

// The look-back of one tile, executed by a whole warp.  Returns the tile's exclusive prefix (all lanes)
// after publishing whatever this tile owes the others (its group aggregate, the super-group prefix).
template <typename P>
__device__ __forceinline__ typename P::A scan_lookback(const CumsumParams& p, int64_t tile, typename P::A tile_total,
                                                       typename P::A start, int lane) {
  using A = typename P::A;
    const int64_t g = tile / kScGroup;
    const int qpos = (int)(tile - g * kScGroup);
    // Everything this tile needs from earlier tiles, fetched as ONE batch of independent polls
    // (re-issued only for words whose flag is not up yet): slots 0..3 = aggregates of the earlier
    // tiles of this group (four per lane), slot 4 = aggregates of the earlier groups of this
    // super-group (one per lane), slot 5 = the previous super-group's inclusive prefix.
    const int64_t sg = g / kScSuper;
    const int gq = (int)(g - sg * kScSuper);
    constexpr int kSlots = kScGroup / 32 + 2;
    const unsigned long long* src[kSlots];
    unsigned need = 0;
#pragma unroll
    for (int r = 0; r < kScGroup / 32; ++r) {
      const int pos = r * 32 + lane;
      src[r] = p.agg + (g * kScGroup + pos) * P::K;
      if (pos < qpos) need |= ((1u << P::K) - 1u) << (r * P::K);
    }
    src[kSlots - 2] = p.gagg + (sg * kScSuper + lane) * P::K;
    if (lane < gq) need |= ((1u << P::K) - 1u) << ((kSlots - 2) * P::K);
    src[kSlots - 1] = p.sincl + (sg > 0 ? sg - 1 : 0) * P::K;
    if (sg > 0) need |= ((1u << P::K) - 1u) << ((kSlots - 1) * P::K);
    unsigned val[kSlots][3];
#pragma unroll
    for (int r = 0; r < kSlots; ++r) { val[r][0] = 0u; val[r][1] = 0u; val[r][2] = 0u; }
    unsigned got = 0;
    constexpr unsigned kAggMask = (1u << ((kScGroup / 32) * P::K)) - 1u;
    bool folded = false;
    A part = P::zero();
    while (true) {  // warp-uniform loop: the fold below uses shuffles
      unsigned long long sw[kSlots][3];
#pragma unroll
      for (int r = 0; r < kSlots; ++r)
#pragma unroll
        for (int k = 0; k < P::K; ++k)
          sw[r][k] = (((need & ~got) >> (r * P::K + k)) & 1u) ? ld_word(src[r] + k) : 0ull;
#pragma unroll
      for (int r = 0; r < kSlots; ++r)
#pragma unroll
        for (int k = 0; k < P::K; ++k)
          if (sw[r][k] & kScFlag) { val[r][k] = (unsigned)sw[r][k]; got |= 1u << (r * P::K + k); }
      if (!folded && __all_sync(0xffffffffu, ((got ^ need) & kAggMask) == 0u)) {
        // tile aggregates are in: fold them in a fixed order (lower tiles on the left) and, for the
        // last tile of a group, publish the group aggregate WITHOUT waiting for anything older
        folded = true;
#pragma unroll
        for (int r = 0; r < kScGroup / 32; ++r) {
          A a = (r * 32 + lane < qpos) ? P::from_words(val[r]) : P::zero();
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const A up = P::shfl_up(a, d);
            if (lane >= d) a = P::add(up, a);
          }
          part = P::add(part, P::shfl(a, 31));
        }
        if (qpos == kScGroup - 1 && lane == 0) {
          unsigned wg[3] = {0u, 0u, 0u};
          P::to_words(P::add(part, tile_total), wg);
#pragma unroll
          for (int k = 0; k < P::K; ++k) st_word(p.gagg + g * P::K + k, wg[k]);
        }
      }
      if (__all_sync(0xffffffffu, got == need)) break;
    }
    A gpart = (lane < gq) ? P::from_words(val[kSlots - 2]) : P::zero();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const A up = P::shfl_up(gpart, d);
      if (lane >= d) gpart = P::add(up, gpart);
    }
    gpart = P::shfl(gpart, 31);
    const A base = (sg > 0) ? P::from_words(val[kSlots - 1]) : start;
    const A excl = P::add(P::add(base, gpart), part);
    if (lane == 0 && qpos == kScGroup - 1 && gq == kScSuper - 1) {
      unsigned wi[3] = {0u, 0u, 0u};
      P::to_words(P::add(excl, tile_total), wi);
#pragma unroll
      for (int k = 0; k < P::K; ++k) st_word(p.sincl + sg * P::K + k, wi[k]);
    }
    return excl;
}

template <typename T, bool kVec, bool kHasValid, bool kChecked>
__global__ void __launch_bounds__(kScThreads, 2)
cumsum_kernel(const CumsumParams p) {
  using P = typename PolSel<T, kChecked>::type;
  using A = typename P::A;
  constexpr int N = 16 / sizeof(T);            // elements per 16-byte vector
  constexpr int kTileRows = kScTileBytes / sizeof(T);
  constexpr int kSegRows = kTileRows / kScWarps;  // rows of one warp segment
  constexpr int E = kScRows * N;                  // rows owned by one lane (128 contiguous bytes)
  __shared__ uint4 s_tile[kScThreads * kScRows];  // 32 KB: the tile, transposed warp by warp
  __shared__ unsigned long long s_warp_lo[kScWarps];
  __shared__ int s_warp_hi[kScWarps];
  __shared__ unsigned long long s_excl_lo;
  __shared__ int s_excl_hi;
  __shared__ CumsumState s_state;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* __restrict__ in = reinterpret_cast<const T*>(p.in);
  T* __restrict__ out = reinterpret_cast<T*>(p.out);

  // every block reads the carried state BEFORE it publishes anything; the last tile rewrites the state
  // only after all earlier tiles have published, i.e. after every block has passed this point
  if (threadIdx.x == 0) s_state = *p.state;
  __syncthreads();
  const bool dead = !p.skip_nulls && s_state.encountered_null != 0;  // a previous chunk hit a null
  int64_t limit = p.n;                                               // rows >= limit are null
  if (dead) limit = 0;
  else if (kHasValid && !p.skip_nulls) limit = *p.first_null;
  const A start = P::from_state(s_state);
  long long my_bad = AG_NO_ERROR_POS;
  const int64_t vlo = p.voff >> 3, vhi = (p.voff + p.n + 7) >> 3;   // byte range of the validity bitmap

  // Coalesced tile load: row k of a warp's 4 KB segment is 32 consecutive 16-byte vectors.  The
  // loads of tile t+G are issued while tile t is still waiting on its look-back and being written
  // out, so HBM stays busy across the only serial part of the algorithm.
  uint4 q[kScRows];
  auto load_tile = [&](int64_t tile) {
    const int64_t r0 = tile * kTileRows + (int64_t)warp * kSegRows;
#pragma unroll
    for (int k = 0; k < kScRows; ++k) {
      const int64_t e0 = r0 + ((int64_t)k * 32 + lane) * N;
      if (kVec && e0 + N <= p.n) {
        q[k] = __ldcs(reinterpret_cast<const uint4*>(in + e0));
      } else {
        __align__(16) T tmp[N];
#pragma unroll
        for (int j = 0; j < N; ++j) tmp[j] = (e0 + j < p.n) ? in[e0 + j] : T(0);
        q[k] = *reinterpret_cast<const uint4*>(tmp);
      }
    }
  };
  if ((int64_t)blockIdx.x < p.n_tiles) load_tile(blockIdx.x);

  for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTileRows + (int64_t)warp * kSegRows;  // first row of this warp's segment
    // ---- load (coalesced): row k of the warp's 4 KB segment is 32 consecutive 16-byte vectors -----
    // and transpose through shared memory so that lane l OWNS the 128 contiguous bytes l of the
    // segment.  Vector vi = k*32 + lane belongs to owner vi>>3, chunk vi&7; it is parked at
    // chunk ^ (owner & 7) of the owner's 128-byte row: both the striped writes and the blocked reads
    // touch every bank once per quarter warp.
    uint4* seg = s_tile + warp * (kScRows * 32);
#pragma unroll
    for (int k = 0; k < kScRows; ++k) {  // q[] was loaded one tile ahead (see the prefetch below)
      const int vi = k * 32 + lane;
      const int owner = vi >> 3, c = vi & 7;
      seg[owner * 8 + (c ^ (owner & 7))] = q[k];
    }
    __syncwarp();
    uint4 raw[kScRows];  // this lane's E = 8*N consecutive elements, still packed
#pragma unroll
    for (int c = 0; c < kScRows; ++c) raw[c] = seg[lane * 8 + (c ^ (lane & 7))];
    const int64_t t0 = row0 + (int64_t)lane * E;  // first row owned by this lane
    // validity of the E rows, 32 at a time; rows at or past `limit` (and past n) contribute nothing
    unsigned vbits[(E + 31) / 32];
#pragma unroll
    for (int w = 0; w < (E + 31) / 32; ++w) {
      const int64_t e0 = t0 + w * 32;
      unsigned m = 0xffffffffu;
      if (kHasValid) m = (e0 < p.n) ? bitmap_load32(p.valid, p.voff + e0, vlo, vhi) : 0u;
      const int64_t room = limit - e0;
      if (room < 32) m &= (room <= 0) ? 0u : ((1u << (int)room) - 1u);
      if (E < 32) m &= (1u << (E & 31)) - 1u;
      vbits[w] = m;
    }
    // ---- lane total, ONE warp scan, segment total ---------------------------------------------
    // all E rows valid (the common case): no per-element bit tests — the kernel is issue-bound, not
    // byte-bound (ncu: 7800 warp instructions per 32 KB tile at IPC 1.6)
    bool dense = true;
#pragma unroll
    for (int w = 0; w < (E + 31) / 32; ++w) dense = dense && vbits[w] == ((E >= 32) ? 0xffffffffu : ((1u << (E & 31)) - 1u));
    A tot = P::zero();
    if (dense) {
#pragma unroll
      for (int i = 0; i < E; ++i) tot = P::add_elem(tot, reinterpret_cast<const T*>(raw)[i]);
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i)
        if ((vbits[i >> 5] >> (i & 31)) & 1u) tot = P::add_elem(tot, reinterpret_cast<const T*>(raw)[i]);
    }
    A incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const A up = P::shfl_up(incl, d);
      if (lane >= d) incl = P::add(up, incl);
    }
    A lane_excl = P::shfl_up(incl, 1);
    if (lane == 0) lane_excl = P::zero();
    const A carry = P::shfl(incl, 31);
    // ---- block: exclusive offsets of the 8 warp segments, tile total ----------------------------
    {
      unsigned w[3] = {0u, 0u, 0u};
      P::to_words(carry, w);
      if (lane == 0) { s_warp_lo[warp] = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32); s_warp_hi[warp] = (int)w[2]; }
    }
    __syncthreads();
    A warp_excl = P::zero(), tile_total = P::zero();
#pragma unroll
    for (int wi = 0; wi < kScWarps; ++wi) {
      const unsigned w[3] = {(unsigned)s_warp_lo[wi], (unsigned)(s_warp_lo[wi] >> 32), (unsigned)s_warp_hi[wi]};
      const A t = P::from_words(w);
      if (wi == warp) warp_excl = tile_total;
      tile_total = P::add(tile_total, t);
    }
    // ---- look-back (warp 0): publish the aggregate, gather the exclusive prefix of the tile -----
    if (warp == 0 && lane == 0) {
      unsigned w[3] = {0u, 0u, 0u};
      P::to_words(tile_total, w);
#pragma unroll
      for (int k = 0; k < P::K; ++k) st_word(p.agg + tile * P::K + k, w[k]);
    }
    if (tile + gridDim.x < p.n_tiles) load_tile(tile + gridDim.x);  // prefetch: in flight during the wait below
    // Unchecked flavours: the running sums RELATIVE to the tile start need nothing from other tiles, so
    // warps 1..7 compute them while warp 0 is in the look-back; after the barrier only the tile's
    // exclusive prefix is added.  (Checked integers keep the exact 96-bit pass after the barrier.)
    auto local_pass = [&]() {
      A run = P::add(warp_excl, lane_excl);
      T* o = reinterpret_cast<T*>(raw);  // results replace the inputs in place
      if (dense) {
#pragma unroll
        for (int i = 0; i < E; ++i) { run = P::add_elem(run, o[i]); o[i] = P::value(run); }
      } else {
#pragma unroll
        for (int i = 0; i < E; ++i) {
          if ((vbits[i >> 5] >> (i & 31)) & 1u) { run = P::add_elem(run, o[i]); o[i] = P::value(run); }
          else o[i] = T(0);
        }
      }
    };
    if (!kChecked && warp != 0) local_pass();
    if (warp == 0) {
      const A excl = scan_lookback<P>(p, tile, tile_total, start, lane);
      if (lane == 0) {
        unsigned w[3] = {0u, 0u, 0u};
        P::to_words(excl, w);
        s_excl_lo = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
        s_excl_hi = (int)w[2];
        const A incl = P::add(excl, tile_total);
        if (tile == p.n_tiles - 1) {  // carry to the next chunk (null_count was advanced by cumsum_validity_kernel)
          CumsumState ns = s_state;
          P::to_state(incl, &ns);
          if (kHasValid && *p.first_null < p.n) ns.encountered_null = 1;
          *p.state = ns;
        }
      }
    }
    __syncthreads();
    A tile_excl;
    {
      const unsigned w[3] = {(unsigned)s_excl_lo, (unsigned)(s_excl_lo >> 32), (unsigned)s_excl_hi};
      tile_excl = P::from_words(w);
    }
    // ---- running sums of the valid slots (0 in null slots), back through shared memory, store --
    if constexpr (kChecked) {
      A run = P::add(P::add(tile_excl, warp_excl), lane_excl);
      T* o = reinterpret_cast<T*>(raw);
#pragma unroll
      for (int i = 0; i < E; ++i) {
        if ((vbits[i >> 5] >> (i & 31)) & 1u) {
          run = P::add_elem(run, o[i]);
          o[i] = P::value(run);
          if (P::out_of_range(run) && t0 + i < my_bad) my_bad = t0 + i;
        } else {
          o[i] = T(0);
        }
      }
    } else {
      if (warp == 0) local_pass();
      const T off = P::value(tile_excl);  // integers: the sum modulo 2^w; floats: the prefix itself
      T* o = reinterpret_cast<T*>(raw);
      if (dense) {
#pragma unroll
        for (int i = 0; i < E; ++i) o[i] = P::offset_add(off, o[i]);
      } else {
#pragma unroll
        for (int i = 0; i < E; ++i)
          if ((vbits[i >> 5] >> (i & 31)) & 1u) o[i] = P::offset_add(off, o[i]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < kScRows; ++c) seg[lane * 8 + (c ^ (lane & 7))] = raw[c];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kScRows; ++k) {
      const int vi = k * 32 + lane;
      const int64_t e0 = row0 + (int64_t)vi * N;
      if (e0 >= p.n) continue;
      const int owner = vi >> 3, c = vi & 7;
      const uint4 q = seg[owner * 8 + (c ^ (owner & 7))];
      if (kVec && e0 + N <= p.n) {
        __stcs(reinterpret_cast<uint4*>(out + e0), q);
      } else {
        const T* o = reinterpret_cast<const T*>(&q);
#pragma unroll
        for (int j = 0; j < N; ++j) if (e0 + j < p.n) out[e0 + j] = o[j];
      }
    }
    __syncthreads();  // s_warp_* / s_excl_* are rewritten by the next tile
  }
  if (kChecked) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const long long o = __shfl_xor_sync(0xffffffffu, my_bad, d);
      my_bad = o < my_bad ? o : my_bad;
    }
    if (lane == 0 && my_bad != AG_NO_ERROR_POS) atomicMin(p.first_bad, my_bad);
  }
}

// ---------------------------------------------------------------- TMA-fed variant -----------
// Same scan, but tiles travel through a ring of shared-memory stages moved by the TMA engine
// (cp.async.bulk, mbarrier complete_tx; SASS UBLKCP): loads are issued kTmaStages-1 tiles ahead
// of the tile being scanned, results go back into the stage in place and leave with ONE bulk
// store per tile.  Nothing the threads do — the scans, the look-back wait, the barrier — stops
// HBM traffic any more, and no registers are spent on staging.  The stage holds the tile exactly
// as it lies in memory; lane l owns bytes [l*128, l*128+128) of its warp's 4 KB segment and reads
// the eight 16-byte chunks in the rotated order (c + l) & 7, which touches every bank once per
// quarter warp without a swizzled copy; chunk sums are put back in order with 8x8 predicated adds.
// 4- and 8-byte types with 16-byte aligned operands; everything else uses cumsum_kernel.
// Opt-in (AG_SCAN_TMA=1): it measured the same as the register-prefetch kernel (see launch_cumsum).
constexpr int kTmaScanStages = 2;   // stage t+1 is refilled while tile t is in its look-back
constexpr int kTmaScanBlocksPerSM = 3;  // 64 KB of ring per block

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)), "l"(gsrc),
               "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T, bool kHasValid, bool kChecked>
__global__ void __launch_bounds__(kScThreads, kTmaScanBlocksPerSM)
cumsum_tma_kernel(const CumsumParams p) {
  using P = typename PolSel<T, kChecked>::type;
  using A = typename P::A;
  constexpr int N = 16 / sizeof(T);
  constexpr int kTileRows = kScTileBytes / sizeof(T);
  constexpr int kSegRows = kTileRows / kScWarps;
  constexpr int E = kScRows * N;  // <= 32 for the types routed here
  static_assert(E <= 32, "the TMA variant keeps a lane's validity bits in one word");
  extern __shared__ __align__(128) unsigned char s_ring[];  // kTmaScanStages x 32 KB
  __shared__ unsigned long long s_bar[kTmaScanStages];
  __shared__ unsigned long long s_warp_lo[kScWarps];
  __shared__ int s_warp_hi[kScWarps];
  __shared__ unsigned long long s_excl_lo;
  __shared__ int s_excl_hi;
  __shared__ CumsumState s_state;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* __restrict__ in = reinterpret_cast<const T*>(p.in);
  T* __restrict__ out = reinterpret_cast<T*>(p.out);

  if (threadIdx.x == 0) {
    s_state = *p.state;
#pragma unroll
    for (int s = 0; s < kTmaScanStages; ++s) mbar_init(&s_bar[s], 1);
    fence_async_smem();
  }
  __syncthreads();
  const bool dead = !p.skip_nulls && s_state.encountered_null != 0;
  int64_t limit = p.n;
  if (dead) limit = 0;
  else if (kHasValid && !p.skip_nulls) limit = *p.first_null;
  const A start = P::from_state(s_state);
  long long my_bad = AG_NO_ERROR_POS;
  const int64_t vlo = p.voff >> 3, vhi = (p.voff + p.n + 7) >> 3;
  const int64_t full_tiles = p.n / kTileRows;  // tiles below this index are complete: TMA moves them

  auto issue_load = [&](int64_t it) {  // thread 0 only
    const int64_t tile = blockIdx.x + it * (int64_t)gridDim.x;
    if (tile < full_tiles) {
      const int s = (int)(it % kTmaScanStages);
      mbar_expect_tx(&s_bar[s], kScTileBytes);
      tma_load_1d(s_ring + (size_t)s * kScTileBytes, in + tile * kTileRows, kScTileBytes, &s_bar[s]);
    }
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < kTmaScanStages - 1; ++j) issue_load(j);
  }

  int64_t it = 0;
  for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
    const int s = (int)(it % kTmaScanStages);
    unsigned char* stage = s_ring + (size_t)s * kScTileBytes;
    const bool full = tile < full_tiles;
    const int64_t row0 = tile * kTileRows + (int64_t)warp * kSegRows;
    const int64_t t0 = row0 + (int64_t)lane * E;
    if (full) {
      mbar_wait(&s_bar[s], (unsigned)((it / kTmaScanStages) & 1));
    } else {  // the last, partial tile: plain loads into the stage (zero-padded)
      T* st = reinterpret_cast<T*>(stage);
      for (int i = threadIdx.x; i < kTileRows; i += kScThreads) {
        const int64_t r = tile * kTileRows + i;
        st[i] = r < p.n ? in[r] : T(0);
      }
      __syncthreads();
    }
    // ---- this lane's 128 bytes, chunk (c + lane) & 7 at step c: conflict-free on the linear tile ----
    uint4* seg = reinterpret_cast<uint4*>(stage) + warp * (kScRows * 32);
    uint4 raw[kScRows];
#pragma unroll
    for (int c = 0; c < kScRows; ++c) raw[c] = seg[lane * 8 + ((c + lane) & 7)];
    unsigned vb = 0xffffffffu;
    if (kHasValid) vb = (t0 < p.n) ? bitmap_load32(p.valid, p.voff + t0, vlo, vhi) : 0u;
    {
      const int64_t room = limit - t0;
      if (room < 32) vb &= (room <= 0) ? 0u : ((1u << (int)room) - 1u);
      if (E < 32) vb &= (1u << (E & 31)) - 1u;
    }
    // ---- chunk sums (rotated order), lane total, one warp scan ------------------------------------
    A csum[kScRows];
    A tot = P::zero();
#pragma unroll
    for (int c = 0; c < kScRows; ++c) {
      const unsigned cb = (vb >> (((c + lane) & 7) * N)) & ((1u << N) - 1u);
      A a = P::zero();
#pragma unroll
      for (int e = 0; e < N; ++e) if ((cb >> e) & 1u) a = P::add_elem(a, reinterpret_cast<const T*>(&raw[c])[e]);
      csum[c] = a;
      tot = P::add(tot, a);
    }
    A incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const A up = P::shfl_up(incl, d);
      if (lane >= d) incl = P::add(up, incl);
    }
    A lane_excl = P::shfl_up(incl, 1);
    if (lane == 0) lane_excl = P::zero();
    const A carry = P::shfl(incl, 31);
    {
      unsigned w[3] = {0u, 0u, 0u};
      P::to_words(carry, w);
      if (lane == 0) { s_warp_lo[warp] = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32); s_warp_hi[warp] = (int)w[2]; }
    }
    __syncthreads();
    A warp_excl = P::zero(), tile_total = P::zero();
#pragma unroll
    for (int wi = 0; wi < kScWarps; ++wi) {
      const unsigned w[3] = {(unsigned)s_warp_lo[wi], (unsigned)(s_warp_lo[wi] >> 32), (unsigned)s_warp_hi[wi]};
      const A t = P::from_words(w);
      if (wi == warp) warp_excl = tile_total;
      tile_total = P::add(tile_total, t);
    }
    if (warp == 0 && lane == 0) {
      unsigned w[3] = {0u, 0u, 0u};
      P::to_words(tile_total, w);
#pragma unroll
      for (int k = 0; k < P::K; ++k) st_word(p.agg + tile * P::K + k, w[k]);
    }
    // refill the ring now — the bulk store of the previous tile has long released its stage — so the
    // next tile's bytes travel while this one sits in its look-back
    if (threadIdx.x == 0) {
      tma_store_wait_read<0>();
      issue_load(it + kTmaScanStages - 1);
    }
    // exclusive offset of each chunk inside the lane: the sums of the chunks that precede it in memory
    A coff[kScRows];
#pragma unroll
    for (int c = 0; c < kScRows; ++c) {
      A a = P::zero();
#pragma unroll
      for (int c2 = 0; c2 < kScRows; ++c2)
        if (((c2 + lane) & 7) < ((c + lane) & 7)) a = P::add(a, csum[c2]);
      coff[c] = a;
    }
    // running sums relative to the tile start (unchecked) while warp 0 is in the look-back
    auto local_pass = [&]() {
      const A lbase = P::add(warp_excl, lane_excl);
#pragma unroll
      for (int c = 0; c < kScRows; ++c) {
        const unsigned cb = (vb >> (((c + lane) & 7) * N)) & ((1u << N) - 1u);
        A run = P::add(lbase, coff[c]);
        T* o = reinterpret_cast<T*>(&raw[c]);
#pragma unroll
        for (int e = 0; e < N; ++e) {
          if ((cb >> e) & 1u) { run = P::add_elem(run, o[e]); o[e] = P::value(run); }
          else o[e] = T(0);
        }
      }
    };
    if (!kChecked && warp != 0) local_pass();
    if (warp == 0) {
      const A excl = scan_lookback<P>(p, tile, tile_total, start, lane);
      if (lane == 0) {
        unsigned w[3] = {0u, 0u, 0u};
        P::to_words(excl, w);
        s_excl_lo = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
        s_excl_hi = (int)w[2];
        if (tile == p.n_tiles - 1) {
          CumsumState ns = s_state;
          P::to_state(P::add(excl, tile_total), &ns);
          if (kHasValid && *p.first_null < p.n) ns.encountered_null = 1;
          *p.state = ns;
        }
      }
    }
    __syncthreads();
    A tile_excl;
    {
      const unsigned w[3] = {(unsigned)s_excl_lo, (unsigned)(s_excl_lo >> 32), (unsigned)s_excl_hi};
      tile_excl = P::from_words(w);
    }
    if constexpr (kChecked) {
      const A lbase = P::add(P::add(tile_excl, warp_excl), lane_excl);
#pragma unroll
      for (int c = 0; c < kScRows; ++c) {
        const int j = (c + lane) & 7;
        const unsigned cb = (vb >> (j * N)) & ((1u << N) - 1u);
        A run = P::add(lbase, coff[c]);
        T* o = reinterpret_cast<T*>(&raw[c]);
#pragma unroll
        for (int e = 0; e < N; ++e) {
          if ((cb >> e) & 1u) {
            run = P::add_elem(run, o[e]);
            o[e] = P::value(run);
            const long long row = t0 + j * N + e;
            if (P::out_of_range(run) && row < my_bad) my_bad = row;
          } else {
            o[e] = T(0);
          }
        }
      }
    } else {
      if (warp == 0) local_pass();
      const T off = P::value(tile_excl);
#pragma unroll
      for (int c = 0; c < kScRows; ++c) {
        const unsigned cb = (vb >> (((c + lane) & 7) * N)) & ((1u << N) - 1u);
        T* o = reinterpret_cast<T*>(&raw[c]);
#pragma unroll
        for (int e = 0; e < N; ++e) if ((cb >> e) & 1u) o[e] = P::offset_add(off, o[e]);
      }
    }
    // ---- results back into the stage (same places), one bulk store for the tile -------------------
#pragma unroll
    for (int c = 0; c < kScRows; ++c) seg[lane * 8 + ((c + lane) & 7)] = raw[c];
    if (full) {
      fence_async_smem();  // make the generic-proxy writes visible to the TMA engine
      __syncthreads();
      if (threadIdx.x == 0) {
        tma_store_1d(out + tile * kTileRows, stage, kScTileBytes);
      }
    } else {
      __syncthreads();
      const T* st = reinterpret_cast<const T*>(stage);
      for (int i = threadIdx.x; i < kTileRows; i += kScThreads) {
        const int64_t r = tile * kTileRows + i;
        if (r < p.n) out[r] = st[i];
      }
    }
  }
  if (threadIdx.x == 0) tma_store_wait_read<0>();  // shared memory must outlive the last bulk store
  if (kChecked) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const long long o = __shfl_xor_sync(0xffffffffu, my_bad, d);
      my_bad = o < my_bad ? o : my_bad;
    }
    if (lane == 0 && my_bad != AG_NO_ERROR_POS) atomicMin(p.first_bad, my_bad);
  }
}

// ---------------------------------------------------------------- two-phase pipelined variant --
// cumsum_tma_kernel with each tile's work split in two phases that a block interleaves ACROSS tiles:
//   A(k): wait for the tile, per-lane / warp / block totals, publish the tile aggregate, remember the
//         lane's offset inside the tile;
//   B(k): look-back (the aggregates it needs were published a whole A-phase ago, so the first batch of
//         polls normally finds them all), running sums, bulk store.
// The block runs A(k), refills the ring, then B(k-1): the publish -> poll round trip through L2 — the
// one thing every tile has to sit through — is covered by the next tile's A phase instead of a barrier
// with seven idle warps (ncu on the single-phase kernels: 45-50 % of stall samples there).  The tile
// stays in its shared-memory stage between the phases; B re-derives the chunk sums from it.
constexpr int kPipeStages = 3;

template <typename T, bool kHasValid, bool kChecked>
__global__ void __launch_bounds__(kScThreads, 2)
cumsum_pipe_kernel(const CumsumParams p) {
  using P = typename PolSel<T, kChecked>::type;
  using A = typename P::A;
  constexpr int N = 16 / sizeof(T);
  constexpr int kTileRows = kScTileBytes / sizeof(T);
  constexpr int kSegRows = kTileRows / kScWarps;
  constexpr int E = kScRows * N;
  static_assert(E <= 32, "a lane's validity bits live in one word");
  extern __shared__ __align__(128) unsigned char s_ring[];  // kPipeStages x 32 KB
  __shared__ unsigned long long s_bar[kPipeStages];
  __shared__ unsigned long long s_base_lo[kPipeStages][kScThreads];  // lane offset inside the tile (phase A -> B)
  __shared__ int s_base_hi[kPipeStages][kScThreads];
  __shared__ unsigned s_vb[kPipeStages][kScThreads];
  __shared__ unsigned long long s_tot_lo[kPipeStages];
  __shared__ int s_tot_hi[kPipeStages];
  __shared__ unsigned long long s_warp_lo[kScWarps];
  __shared__ int s_warp_hi[kScWarps];
  __shared__ unsigned long long s_excl_lo;
  __shared__ int s_excl_hi;
  __shared__ CumsumState s_state;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* __restrict__ in = reinterpret_cast<const T*>(p.in);
  T* __restrict__ out = reinterpret_cast<T*>(p.out);

  if (threadIdx.x == 0) {
    s_state = *p.state;
#pragma unroll
    for (int s = 0; s < kPipeStages; ++s) mbar_init(&s_bar[s], 1);
    fence_async_smem();
  }
  __syncthreads();
  const bool dead = !p.skip_nulls && s_state.encountered_null != 0;
  int64_t limit = p.n;
  if (dead) limit = 0;
  else if (kHasValid && !p.skip_nulls) limit = *p.first_null;
  const A start = P::from_state(s_state);
  long long my_bad = AG_NO_ERROR_POS;
  const int64_t vlo = p.voff >> 3, vhi = (p.voff + p.n + 7) >> 3;
  const int64_t full_tiles = p.n / kTileRows;
  const int64_t n_my = (p.n_tiles - 1 - (int64_t)blockIdx.x) / (int64_t)gridDim.x + 1;  // tiles of this block (>= 1)

  auto issue_load = [&](int64_t k) {  // thread 0 only
    const int64_t tile = blockIdx.x + k * (int64_t)gridDim.x;
    if (k < n_my && tile < full_tiles) {
      const int s = (int)(k % kPipeStages);
      mbar_expect_tx(&s_bar[s], kScTileBytes);
      tma_load_1d(s_ring + (size_t)s * kScTileBytes, in + tile * kTileRows, kScTileBytes, &s_bar[s]);
    }
  };
  if (threadIdx.x == 0) issue_load(0);

  for (int64_t k = 0; k <= n_my; ++k) {
    // ================================ phase A of tile k ===========================================
    if (k < n_my) {
      const int64_t tile = blockIdx.x + k * (int64_t)gridDim.x;
      const int s = (int)(k % kPipeStages);
      unsigned char* stage = s_ring + (size_t)s * kScTileBytes;
      const int64_t t0 = tile * kTileRows + (int64_t)warp * kSegRows + (int64_t)lane * E;
      if (tile < full_tiles) {
        mbar_wait(&s_bar[s], (unsigned)((k / kPipeStages) & 1));
      } else {
        T* st = reinterpret_cast<T*>(stage);
        for (int i = threadIdx.x; i < kTileRows; i += kScThreads) {
          const int64_t r = tile * kTileRows + i;
          st[i] = r < p.n ? in[r] : T(0);
        }
        __syncthreads();
      }
      const uint4* seg = reinterpret_cast<const uint4*>(stage) + warp * (kScRows * 32);
      unsigned vb = 0xffffffffu;
      if (kHasValid) vb = (t0 < p.n) ? bitmap_load32(p.valid, p.voff + t0, vlo, vhi) : 0u;
      {
        const int64_t room = limit - t0;
        if (room < 32) vb &= (room <= 0) ? 0u : ((1u << (int)room) - 1u);
        if (E < 32) vb &= (1u << (E & 31)) - 1u;
      }
      A tot = P::zero();
#pragma unroll
      for (int c = 0; c < kScRows; ++c) {
        const uint4 q = seg[lane * 8 + ((c + lane) & 7)];
        const unsigned cb = (vb >> (((c + lane) & 7) * N)) & ((1u << N) - 1u);
#pragma unroll
        for (int e = 0; e < N; ++e) if ((cb >> e) & 1u) tot = P::add_elem(tot, reinterpret_cast<const T*>(&q)[e]);
      }
      A incl = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const A up = P::shfl_up(incl, d);
        if (lane >= d) incl = P::add(up, incl);
      }
      A lane_excl = P::shfl_up(incl, 1);
      if (lane == 0) lane_excl = P::zero();
      {
        unsigned w[3] = {0u, 0u, 0u};
        P::to_words(P::shfl(incl, 31), w);
        if (lane == 0) { s_warp_lo[warp] = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32); s_warp_hi[warp] = (int)w[2]; }
      }
      __syncthreads();
      A warp_excl = P::zero(), tile_total = P::zero();
#pragma unroll
      for (int wi = 0; wi < kScWarps; ++wi) {
        const unsigned w[3] = {(unsigned)s_warp_lo[wi], (unsigned)(s_warp_lo[wi] >> 32), (unsigned)s_warp_hi[wi]};
        const A t = P::from_words(w);
        if (wi == warp) warp_excl = tile_total;
        tile_total = P::add(tile_total, t);
      }
      {
        unsigned w[3] = {0u, 0u, 0u};
        P::to_words(tile_total, w);
        if (threadIdx.x == 0) {
#pragma unroll
          for (int kk = 0; kk < P::K; ++kk) st_word(p.agg + tile * P::K + kk, w[kk]);
          s_tot_lo[s] = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
          s_tot_hi[s] = (int)w[2];
        }
        unsigned wb[3] = {0u, 0u, 0u};
        P::to_words(P::add(warp_excl, lane_excl), wb);
        s_base_lo[s][threadIdx.x] = (unsigned long long)wb[0] | ((unsigned long long)wb[1] << 32);
        s_base_hi[s][threadIdx.x] = (int)wb[2];
        s_vb[s][threadIdx.x] = vb;
      }
      __syncthreads();  // s_warp_* may be rewritten by the next phase A; s_tot_* is read by warp 0 in phase B
    }
    // ================================ refill the ring ==============================================
    // tile k+1 goes into the stage tile k-2 used; its bulk store was issued a whole phase A ago
    if (threadIdx.x == 0) {
      tma_store_wait_read<0>();
      issue_load(k + 1);
    }
    // ================================ phase B of tile k-1 ==========================================
    if (k >= 1) {
      const int64_t j = k - 1;
      const int64_t tile = blockIdx.x + j * (int64_t)gridDim.x;
      const int s = (int)(j % kPipeStages);
      unsigned char* stage = s_ring + (size_t)s * kScTileBytes;
      const bool full = tile < full_tiles;
      const int64_t t0 = tile * kTileRows + (int64_t)warp * kSegRows + (int64_t)lane * E;
      A tile_total;
      {
        const unsigned w[3] = {(unsigned)s_tot_lo[s], (unsigned)(s_tot_lo[s] >> 32), (unsigned)s_tot_hi[s]};
        tile_total = P::from_words(w);
      }
      if (warp == 0) {
        const A excl = scan_lookback<P>(p, tile, tile_total, start, lane);
        if (lane == 0) {
          unsigned w[3] = {0u, 0u, 0u};
          P::to_words(excl, w);
          s_excl_lo = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
          s_excl_hi = (int)w[2];
          if (tile == p.n_tiles - 1) {
            CumsumState ns = s_state;
            P::to_state(P::add(excl, tile_total), &ns);
            if (kHasValid && *p.first_null < p.n) ns.encountered_null = 1;
            *p.state = ns;
          }
        }
      }
      __syncthreads();
      A tile_excl, lbase;
      {
        const unsigned w[3] = {(unsigned)s_excl_lo, (unsigned)(s_excl_lo >> 32), (unsigned)s_excl_hi};
        tile_excl = P::from_words(w);
        const unsigned wb[3] = {(unsigned)s_base_lo[s][threadIdx.x], (unsigned)(s_base_lo[s][threadIdx.x] >> 32), (unsigned)s_base_hi[s][threadIdx.x]};
        lbase = P::add(tile_excl, P::from_words(wb));
      }
      const unsigned vb = s_vb[s][threadIdx.x];
      uint4* seg = reinterpret_cast<uint4*>(stage) + warp * (kScRows * 32);
      uint4 raw[kScRows];
      A csum[kScRows];
#pragma unroll
      for (int c = 0; c < kScRows; ++c) {
        raw[c] = seg[lane * 8 + ((c + lane) & 7)];
        const unsigned cb = (vb >> (((c + lane) & 7) * N)) & ((1u << N) - 1u);
        A a = P::zero();
#pragma unroll
        for (int e = 0; e < N; ++e) if ((cb >> e) & 1u) a = P::add_elem(a, reinterpret_cast<const T*>(&raw[c])[e]);
        csum[c] = a;
      }
#pragma unroll
      for (int c = 0; c < kScRows; ++c) {
        const int jj = (c + lane) & 7;
        A run = lbase;  // + the sums of the chunks that precede chunk jj in memory
#pragma unroll
        for (int c2 = 0; c2 < kScRows; ++c2)
          if (((c2 + lane) & 7) < jj) run = P::add(run, csum[c2]);
        const unsigned cb = (vb >> (jj * N)) & ((1u << N) - 1u);
        T* o = reinterpret_cast<T*>(&raw[c]);
#pragma unroll
        for (int e = 0; e < N; ++e) {
          if ((cb >> e) & 1u) {
            run = P::add_elem(run, o[e]);
            o[e] = P::value(run);
            if (kChecked) { const long long row = t0 + jj * N + e; if (P::out_of_range(run) && row < my_bad) my_bad = row; }
          } else {
            o[e] = T(0);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kScRows; ++c) seg[lane * 8 + ((c + lane) & 7)] = raw[c];
      if (full) {
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) tma_store_1d(out + tile * kTileRows, stage, kScTileBytes);
      } else {
        __syncthreads();
        const T* st = reinterpret_cast<const T*>(stage);
        for (int i = threadIdx.x; i < kTileRows; i += kScThreads) {
          const int64_t r = tile * kTileRows + i;
          if (r < p.n) out[r] = st[i];
        }
        __syncthreads();
      }
    }
  }
  if (threadIdx.x == 0) tma_store_wait_read<0>();
  if (kChecked) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const long long o = __shfl_xor_sync(0xffffffffu, my_bad, d);
      my_bad = o < my_bad ? o : my_bad;
    }
    if (lane == 0 && my_bad != AG_NO_ERROR_POS) atomicMin(p.first_bad, my_bad);
  }
}

// first 0 bit of validity[voff, voff+n): one thread per 32 rows, atomicMin into *first_null (pre-set to n)
__global__ void __launch_bounds__(256)
first_null_kernel(const uint8_t* __restrict__ valid, int64_t voff, int64_t n, long long* first_null) {
  const int64_t vlo = voff >> 3, vhi = (voff + n + 7) >> 3;
  const int64_t n_words = (n + 31) >> 5;
  long long mine = n;
  for (int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * 256) {
    uint32_t bits = bitmap_load32(valid, voff + w * 32, vlo, vhi);
    const int64_t room = n - w * 32;
    if (room < 32) bits |= ~((1u << (int)room) - 1u);  // rows past n count as valid
    if (bits != 0xffffffffu) { const long long r = w * 32 + (__ffs(~bits) - 1); if (r < mine) mine = r; break; }  // rows only grow with w
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const long long o = __shfl_xor_sync(0xffffffffu, mine, d);
    mine = o < mine ? o : mine;
  }
  if ((threadIdx.x & 31) == 0 && mine < n) atomicMin(first_null, mine);
}

__global__ void set_i64_kernel(long long* p, long long v) { *p = v; }

// Output validity of one call, one thread per aligned 32-bit output word:
//   SkipNulls: the input validity;  otherwise: valid up to the first null, null from there on
//   (vector_cumulative.go:277-289), everything null once an earlier chunk met a null.
// Runs BEFORE cumsum_kernel on the same stream (it needs the state as the call found it) and adds
// the number of null slots it wrote to state->null_count.
__global__ void __launch_bounds__(256)
cumsum_validity_kernel(const uint8_t* __restrict__ valid, int64_t voff, int64_t n, int skip_nulls, const long long* first_null,
                       CumsumState* state, uint32_t* __restrict__ out_words, int shift, int64_t n_words) {
  const bool dead = !skip_nulls && state->encountered_null != 0;
  const int64_t limit = dead ? 0 : ((valid && !skip_nulls) ? *first_null : n);
  const int64_t vlo = voff >> 3, vhi = (voff + n + 7) >> 3;
  long long nulls = 0;
  for (int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * 256) {
    const int64_t e0 = (w << 5) - shift;
    uint32_t v = 0xffffffffu;
    if (valid) v = bitmap_load32(valid, voff + e0, vlo, vhi);
    const int64_t room = limit - e0;  // rows e0+j with j >= room are null
    if (room < 32) v &= (room <= 0) ? 0u : ((1u << (int)room) - 1u);
    const int64_t lo64 = -e0, hi64 = n - e0;
    const int lo = lo64 > 0 ? (int)lo64 : 0;
    const int hi = hi64 < 32 ? (int)hi64 : 32;
    if (hi > lo) {
      const uint32_t m = bit_range_mask(lo, hi);
      bitmap_store32_masked(out_words + w, v, m);
      nulls += __popc(~v & m);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) nulls += __shfl_xor_sync(0xffffffffu, nulls, d);
  if ((threadIdx.x & 31) == 0 && nulls) atomicAdd(reinterpret_cast<unsigned long long*>(&state->null_count), (unsigned long long)nulls);
}

template <typename T>
ag_status launch_cumsum(CumsumParams& p, cudaStream_t st) {
  constexpr int kStatusK = 3;  // sized for the widest accumulator (exact 96-bit)
  constexpr int kTileRows = kScTileBytes / sizeof(T);
  Workspace* ws;
  AG_TRY(get_workspace(st, &ws));
  p.n_tiles = (p.n + kTileRows - 1) / kTileRows;
  const int64_t n_groups = (p.n_tiles + kScGroup - 1) / kScGroup;
  const int64_t n_super = (n_groups + kScSuper - 1) / kScSuper;
  const size_t words = (size_t)(p.n_tiles + n_groups + n_super) * kStatusK;
  AG_TRY(ensure_tile_status(ws, words, st));
  p.agg = ws->tile_status;
  p.gagg = p.agg + (size_t)p.n_tiles * kStatusK;
  p.sincl = p.gagg + (size_t)n_groups * kStatusK;
  AG_CUDA_TRY(cudaMemsetAsync(ws->tile_status, 0, words * sizeof(unsigned long long), st));
  if (p.valid) {
    long long* fn = reinterpret_cast<long long*>(ws->scalars) + 8;  // slot 8 of the per-stream scalars
    set_i64_kernel<<<1, 1, 0, st>>>(fn, (long long)p.n);
    const int grid = grid_for((p.n + 31) >> 5, 256 * 4, 8);
    first_null_kernel<<<grid, 256, 0, st>>>(p.valid, p.voff, p.n, fn);
    AG_TRY(check_launch("first_null_kernel"));
    p.first_null = fn;
  }
  if (p.out_valid) {
    uint8_t* first = p.out_valid + (p.ooff >> 3);
    const uintptr_t a = reinterpret_cast<uintptr_t>(first);
    uint32_t* words = reinterpret_cast<uint32_t*>(a & ~(uintptr_t)3);
    const int shift = (int)(a & 3) * 8 + (int)(p.ooff & 7);
    const int64_t n_words = (p.n + shift + 31) >> 5;
    cumsum_validity_kernel<<<grid_for(n_words, 256, 8), 256, 0, st>>>(p.valid, p.voff, p.n, p.skip_nulls, p.first_null, p.state, words, shift, n_words);
    AG_TRY(check_launch("cumsum_validity_kernel"));
  }
  const bool vec = (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
  void* args[] = {(void*)&p};
  const void* fn;
  const bool chk = p.checked && !IsFp<T>::v;  // floats: the checked adder is the plain one
  if (chk) {
    if (vec) fn = p.valid ? (const void*)cumsum_kernel<T, true, true, !IsFp<T>::v> : (const void*)cumsum_kernel<T, true, false, !IsFp<T>::v>;
    else fn = p.valid ? (const void*)cumsum_kernel<T, false, true, !IsFp<T>::v> : (const void*)cumsum_kernel<T, false, false, !IsFp<T>::v>;
  } else {
    if (vec) fn = p.valid ? (const void*)cumsum_kernel<T, true, true, false> : (const void*)cumsum_kernel<T, true, false, false>;
    else fn = p.valid ? (const void*)cumsum_kernel<T, false, true, false> : (const void*)cumsum_kernel<T, false, false, false>;
  }
  size_t dyn_smem = 0;
  if constexpr (sizeof(T) >= 4) {
    // opt-in (AG_SCAN_TMA=1, read per call so a test can flip it): measured 433 us vs 437 us for the
    // register-prefetch kernel at 100M int64 rows — the scan is bound by the look-back dependency between
    // concurrently running tiles, not by the load path, so the simpler kernel stays the default
    const char* e = getenv("AG_SCAN_TMA");
    const bool use_tma = e && e[0] == '1';
    const bool use_pipe = e && e[0] == '2';
    if (vec && use_pipe) {
      if (chk) fn = p.valid ? (const void*)cumsum_pipe_kernel<T, true, !IsFp<T>::v> : (const void*)cumsum_pipe_kernel<T, false, !IsFp<T>::v>;
      else fn = p.valid ? (const void*)cumsum_pipe_kernel<T, true, false> : (const void*)cumsum_pipe_kernel<T, false, false>;
      dyn_smem = (size_t)kPipeStages * kScTileBytes;
      AG_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    } else if (vec && use_tma) {
      if (chk) fn = p.valid ? (const void*)cumsum_tma_kernel<T, true, !IsFp<T>::v> : (const void*)cumsum_tma_kernel<T, false, !IsFp<T>::v>;
      else fn = p.valid ? (const void*)cumsum_tma_kernel<T, true, false> : (const void*)cumsum_tma_kernel<T, false, false>;
      dyn_smem = (size_t)kTmaScanStages * kScTileBytes;
      AG_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    }
  }
  int per_sm = 0;
  if (dyn_smem) {
    AG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kScThreads, dyn_smem));
    if (per_sm < 1) AG_FAIL(AG_ERR_CUDA, "cumulative_sum: the TMA scan kernel does not fit on an SM");
  } else {
    per_sm = blocks_per_sm(fn, kScThreads);
  }
  const int64_t cap = (int64_t)sm_count() * per_sm;
  const int grid = (int)(p.n_tiles < cap ? p.n_tiles : cap);
  AG_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kScThreads), args, dyn_smem, st));
  return check_launch("cumsum_kernel");
}

__global__ void cumsum_state_init_kernel(CumsumState* s, unsigned long long lo, long long hi) {
  s->lo = lo; s->hi = hi; s->encountered_null = 0; s->null_count = 0;
}

}  // namespace

ag_status cumulative_sum_dev(int type, const void* in, const uint8_t* valid, int64_t voff, int64_t n, int skip_nulls, int checked,
                             void* out, uint8_t* out_valid, int64_t ooff, void* d_state, int64_t* d_first_bad, cudaStream_t st) {
  if (n < 0 || voff < 0 || ooff < 0) AG_FAIL(AG_ERR_INVALID, "cumulative_sum: negative length or offset");
  if (type_width(type) == 0) AG_FAIL(AG_ERR_TYPE, "cumulative_sum: input type must be numeric, got type id %d", type);
  if (!d_state) AG_FAIL(AG_ERR_INVALID, "cumulative_sum: NULL state");
  if (checked && !d_first_bad) AG_FAIL(AG_ERR_INVALID, "cumulative_sum: the checked flavour needs an error word");
  if (n == 0) return AG_OK;
  if (!in || !out) AG_FAIL(AG_ERR_INVALID, "cumulative_sum: NULL operand");
  if (valid && !out_valid) AG_FAIL(AG_ERR_INVALID, "cumulative_sum: an input with a validity bitmap needs an output validity bitmap");
  CumsumParams p{};
  p.out_valid = out_valid; p.ooff = ooff;
  p.in = in; p.out = out; p.valid = valid; p.voff = voff; p.n = n; p.skip_nulls = skip_nulls; p.checked = checked;
  p.first_null = nullptr; p.state = static_cast<CumsumState*>(d_state); p.first_bad = (long long*)d_first_bad;
  switch (type) {
    case AG_TYPE_UINT8: return launch_cumsum<uint8_t>(p, st);
    case AG_TYPE_INT8: return launch_cumsum<int8_t>(p, st);
    case AG_TYPE_UINT16: return launch_cumsum<uint16_t>(p, st);
    case AG_TYPE_INT16: return launch_cumsum<int16_t>(p, st);
    case AG_TYPE_UINT32: return launch_cumsum<uint32_t>(p, st);
    case AG_TYPE_INT32: return launch_cumsum<int32_t>(p, st);
    case AG_TYPE_UINT64: return launch_cumsum<unsigned long long>(p, st);
    case AG_TYPE_INT64: return launch_cumsum<long long>(p, st);
    case AG_TYPE_FLOAT32: return launch_cumsum<float>(p, st);
    default: return launch_cumsum<double>(p, st);
  }
}

ag_status cumulative_sum_state_init(void* d_state, int type, const void* start_host, cudaStream_t st) {
  if (!d_state) AG_FAIL(AG_ERR_INVALID, "cumulative_sum: NULL state");
  const int w = type_width(type);
  if (w == 0) AG_FAIL(AG_ERR_TYPE, "cumulative_sum: input type must be numeric, got type id %d", type);
  unsigned long long lo = 0;
  long long hi = 0;
  if (start_host) {
    if (type_is_float(type)) memcpy(&lo, start_host, (size_t)w);  // bit pattern (float in the low 32 bits)
    else if (type_is_signed_int(type)) {
      long long v = 0;
      switch (w) { case 1: v = *(const int8_t*)start_host; break; case 2: v = *(const int16_t*)start_host; break;
                   case 4: v = *(const int32_t*)start_host; break; default: v = *(const int64_t*)start_host; break; }
      lo = (unsigned long long)v; hi = v < 0 ? -1 : 0;
    } else {
      switch (w) { case 1: lo = *(const uint8_t*)start_host; break; case 2: lo = *(const uint16_t*)start_host; break;
                   case 4: lo = *(const uint32_t*)start_host; break; default: lo = *(const uint64_t*)start_host; break; }
    }
  }
  cumsum_state_init_kernel<<<1, 1, 0, st>>>(static_cast<CumsumState*>(d_state), lo, hi);
  return check_launch("cumsum_state_init_kernel");
}

}  // namespace ag

using namespace ag;

extern "C" ag_status ag_cumulative_sum_state_init_dev(void* d_state, int type, const void* start_host, ag_stream_t s) {
  AG_TRY(ensure_init());
  return cumulative_sum_state_init(d_state, type, start_host, resolve_stream(s));
}

extern "C" ag_status ag_cumulative_sum_dev(int type, const void* d_in, const uint8_t* d_valid, int64_t valid_offset, int64_t n,
                                           int skip_nulls, int checked, void* d_out, uint8_t* d_out_valid, int64_t out_valid_offset,
                                           void* d_state, int64_t* d_first_bad, ag_stream_t s) {
  AG_TRY(ensure_init());
  return cumulative_sum_dev(type, d_in, d_valid, valid_offset, n, skip_nulls, checked, d_out, d_out_valid, out_valid_offset,
                            d_state, d_first_bad, resolve_stream(s));
}
