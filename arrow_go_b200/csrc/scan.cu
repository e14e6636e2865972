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

// ---- streaming flavour: no validity bitmap, unchecked, 16-byte aligned operands ------------------
// ncu on the block-synchronous shape (general kernel above, and a first streaming kernel with the same
// schedule; profiles/r2/scan_stream_history.txt): 67 % of all warp samples sit at the barrier behind the
// look-back and warp 0 spends 92 % of its time inside it (half waiting for flags of tiles that other
// blocks are reducing at the same moment, half in ~600 dependent instructions of one warp) — every
// resident block is in the same phase, so nothing on the SM covers that wait: 7.2 us per 32 KB tile per
// block against 4.4 us of HBM time.  This kernel takes the look-back off the critical path:
//   * a block = 8 compute warps + 1 LOOK-BACK warp.  In iteration j the compute warps reduce tile j
//     and finalize tile j-1 while the look-back warp resolves the prefix of tile j-1 (whose aggregate,
//     and those of its predecessors in other blocks, were published an iteration ago: the polls hit);
//     one block barrier per tile joins them;
//   * tiles arrive by cp.async (16 bytes per thread, no staging registers) straight into the SWIZZLED
//     position of a 3-stage shared-memory ring (loading / being reduced / waiting for its prefix); a
//     warp only ever touches its own 4 KB segment of a stage, so the ring needs __syncwarp only;
//   * the look-back reads a whole 16-byte status entry per load, issues all its polls as one batch
//     and folds them with interleaved warp scans: ~1 L2 round trip + ~150 instructions;
//   * interior tiles run a branch-free body (base pointer + immediate offsets).
// Same tiles, same status words, same fold order as cumsum_kernel: results are bit-identical to it
// (floats included), so the two kernels can be mixed inside one chunked sequence.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
// 4- / 8-byte pieces for an input that is only element-aligned (an Arrow slice): same destination layout, more copies
template <int kB>
__device__ __forceinline__ void cp_async_small_zfill(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(dst), "l"(src), "n"(kB), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// One status entry = K flagged 64-bit words, fetched with ONE volatile load (K = 2: 128 bits; every word carries its own
// flag, so it does not matter whether the two halves are observed at the same instant).
template <int K> struct StatusEntry;
template <> struct StatusEntry<1> {
  unsigned long long w0;
  __device__ __forceinline__ void load(const unsigned long long* p) { asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w0) : "l"(p) : "memory"); }
  __device__ __forceinline__ bool ready() const { return (w0 & kScFlag) != 0; }
  __device__ __forceinline__ void words(unsigned (&w)[3]) const { w[0] = (unsigned)w0; w[1] = 0u; w[2] = 0u; }
  __device__ __forceinline__ void set_ready_zero() { w0 = kScFlag; }
};
template <> struct StatusEntry<2> {
  unsigned long long w0, w1;
  __device__ __forceinline__ void load(const unsigned long long* p) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
  }
  __device__ __forceinline__ bool ready() const { return ((w0 & w1) & kScFlag) != 0; }
  __device__ __forceinline__ void words(unsigned (&w)[3]) const { w[0] = (unsigned)w0; w[1] = (unsigned)w1; w[2] = 0u; }
  __device__ __forceinline__ void set_ready_zero() { w0 = kScFlag; w1 = kScFlag; }
};

// scan_lookback with the same three levels, the same words and the same fold order, built for latency: whole entries per
// load, every poll of a level in one batch, the four tile-aggregate scans interleaved.  Entries must be 16-byte aligned.
template <typename P>
__device__ __forceinline__ typename P::A scan_lookback_lean(const CumsumParams& p, int64_t tile, typename P::A tile_total,
                                                            typename P::A start, int lane) {
  using A = typename P::A;
  constexpr int K = P::K;
  static_assert(K <= 2, "lean look-back: float / double / wrapped 64-bit accumulators");
  constexpr int R = kScGroup / 32;
  const int64_t g = tile / kScGroup;
  const int qpos = (int)(tile - g * kScGroup);
  const int64_t sg = g / kScSuper;
  const int gq = (int)(g - sg * kScSuper);
  const unsigned long long* a_src = p.agg + (g * kScGroup + lane) * K;
  const unsigned long long* g_src = p.gagg + (sg * kScSuper + lane) * K;
  const unsigned long long* s_src = p.sincl + (sg > 0 ? sg - 1 : 0) * K;
  StatusEntry<K> ea[R], eg, es;
  bool need_a[R];
#pragma unroll
  for (int r = 0; r < R; ++r) need_a[r] = r * 32 + lane < qpos;
  const bool need_g = lane < gq, need_s = sg > 0;
  // first batch: everything at once (one L2 round trip when all flags are up, the common case)
#pragma unroll
  for (int r = 0; r < R; ++r) { if (need_a[r]) ea[r].load(a_src + r * 32 * K); else ea[r].set_ready_zero(); }
  if (need_g) eg.load(g_src); else eg.set_ready_zero();
  if (need_s) es.load(s_src); else es.set_ready_zero();
  while (true) {
    bool ok = true;
#pragma unroll
    for (int r = 0; r < R; ++r) ok = ok && ea[r].ready();
    if (__all_sync(0xffffffffu, ok)) break;
#pragma unroll
    for (int r = 0; r < R; ++r) if (need_a[r] && !ea[r].ready()) ea[r].load(a_src + r * 32 * K);
  }
  A x[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    unsigned w[3];
    ea[r].words(w);
    x[r] = need_a[r] ? P::from_words(w) : P::zero();
  }
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const A up = P::shfl_up(x[r], d);
      if (lane >= d) x[r] = P::add(up, x[r]);
    }
  }
  A part = P::zero();
#pragma unroll
  for (int r = 0; r < R; ++r) part = P::add(part, P::shfl(x[r], 31));
  if (qpos == kScGroup - 1 && lane == 0) {  // group aggregate: needs nothing older than this group
    unsigned wg[3] = {0u, 0u, 0u};
    P::to_words(P::add(part, tile_total), wg);
#pragma unroll
    for (int k = 0; k < K; ++k) st_word(p.gagg + g * K + k, wg[k]);
  }
  while (!__all_sync(0xffffffffu, eg.ready() && es.ready())) {
    if (need_g && !eg.ready()) eg.load(g_src);
    if (need_s && !es.ready()) es.load(s_src);
  }
  unsigned w[3];
  eg.words(w);
  A gpart = need_g ? P::from_words(w) : P::zero();
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const A up = P::shfl_up(gpart, d);
    if (lane >= d) gpart = P::add(up, gpart);
  }
  gpart = P::shfl(gpart, 31);
  es.words(w);
  const A base = need_s ? P::from_words(w) : start;
  const A excl = P::add(P::add(base, gpart), part);
  if (lane == 0 && qpos == kScGroup - 1 && gq == kScSuper - 1) {
    unsigned wi[3] = {0u, 0u, 0u};
    P::to_words(P::add(excl, tile_total), wi);
#pragma unroll
    for (int k = 0; k < K; ++k) st_word(p.sincl + sg * K + k, wi[k]);
  }
  return excl;
}

constexpr int kScStreamThreads = kScThreads + 32;  // 8 compute warps + the look-back warp
constexpr int kScStreamStages = 3;
constexpr int kScStreamBlocksPerSm = 2;            // 2 x (3 x 32 KB) of the SM's 227 KB

template <typename T, bool kHasValid>
__global__ void __launch_bounds__(kScStreamThreads, kScStreamBlocksPerSm)
cumsum_stream_kernel(const CumsumParams p) {
  using P = typename PolSel<T, false>::type;
  using A = typename P::A;
  constexpr int N = 16 / sizeof(T);
  constexpr int E = kScRows * N;                  // rows owned by one lane (128 contiguous bytes)
  constexpr int kTileRows = kScTileBytes / sizeof(T);
  constexpr int kSegRows = kTileRows / kScWarps;
  constexpr int VW = (E + 31) / 32;               // validity words per lane
  extern __shared__ __align__(128) unsigned char s_ring[];  // 3 stages x 32 KB
  __shared__ A s_warp[2][kScWarps];
  __shared__ A s_excl[2];
  __shared__ unsigned s_arrived[2];
  __shared__ CumsumState s_state;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool looker = warp == kScWarps;
  if (threadIdx.x < 2) s_arrived[threadIdx.x] = 0u;
  const unsigned char* __restrict__ in = reinterpret_cast<const unsigned char*>(p.in);
  unsigned char* __restrict__ out = reinterpret_cast<unsigned char*>(p.out);
  const int64_t n_bytes = p.n * (int64_t)sizeof(T);

  if (threadIdx.x == 0) s_state = *p.state;
  __syncthreads();
  const A start = P::from_state(s_state);
  if (!p.skip_nulls && s_state.encountered_null != 0) {
    // an earlier chunk of the sequence met a null: every slot of this call is null (value 0) and the
    // carried state stays as it is (cumsum_kernel's limit = 0 case)
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t b = ((int64_t)blockIdx.x * kScStreamThreads + threadIdx.x) * 16; b < n_bytes; b += (int64_t)gridDim.x * kScStreamThreads * 16) {
      if (b + 16 <= n_bytes) __stcs(reinterpret_cast<uint4*>(out + b), z);
      else for (int64_t j = b; j < n_bytes; ++j) out[j] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { CumsumState ns = s_state; P::to_state(start, &ns); *p.state = ns; }
    return;
  }

  // validity of the E rows a lane owns, 32 at a time: rows at or past `limit` (the first null when nulls are not
  // skipped) and past n contribute nothing — cumsum_kernel's rule.  Returns whether every row counts (the common case).
  int64_t limit = p.n;
  if (kHasValid && !p.skip_nulls) limit = *p.first_null;
  const int64_t vlo = p.voff >> 3, vhi = (p.voff + p.n + 7) >> 3;
  auto lane_validity = [&](int64_t tile, unsigned (&vbits)[VW]) -> bool {
    if (!kHasValid) return true;
    const int64_t t0 = tile * kTileRows + (int64_t)(warp & (kScWarps - 1)) * kSegRows + (int64_t)lane * E;
    bool dense = true;
#pragma unroll
    for (int w = 0; w < VW; ++w) {
      const int64_t e0 = t0 + w * 32;
      unsigned m = (e0 < p.n) ? bitmap_load32(p.valid, p.voff + e0, vlo, vhi) : 0u;
      const int64_t room = limit - e0;
      if (room < 32) m &= (room <= 0) ? 0u : ((1u << (int)room) - 1u);
      constexpr unsigned kFull = (E >= 32) ? 0xffffffffu : ((1u << (E & 31)) - 1u);
      m &= kFull;
      vbits[w] = m;
      dense = dense && m == kFull;
    }
    return dense;
  };

  // byte offsets of this lane inside a 4 KB warp segment
  //   striped (coalesced) vector k:  logical (k*32 + lane) * 16, parked at its owner's row, chunk ^ (owner & 7)
  //   blocked chunk c of row `lane`: lane*128 + ((c ^ (lane & 7)) << 4)
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(s_ring);
  const uint32_t seg_off = (uint32_t)(warp & (kScWarps - 1)) * 4096u;
  const uint32_t str_even = (uint32_t)(((lane & ~7) | ((lane & 7) ^ (lane >> 3))) << 4);  // k even; k odd: ^ 64; + k*512
  const uint32_t blk_row = (uint32_t)lane * 128u;
  const uint32_t blk_x = (uint32_t)(lane & 7) << 4;

  const int in_mis = (int)(reinterpret_cast<uintptr_t>(in) & 15);   // 0, or a multiple of sizeof(T) >= 4 (block-uniform)
  auto prefetch = [&](int64_t tile, int stage) {
    const uint32_t dst0 = ring + (uint32_t)stage * kScTileBytes + seg_off;
    const int64_t g0 = tile * kScTileBytes + seg_off + (int64_t)lane * 16;
    if (in_mis) {
      // element-aligned input: the 16-byte slots of the ring are filled by two 8-byte or four 4-byte copies
#pragma unroll
      for (int k = 0; k < kScRows; ++k) {
        const uint32_t d = dst0 + (str_even ^ ((k & 1) << 6)) + k * 512;
        const int64_t g = g0 + k * 512;
        if ((in_mis & 7) == 0) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int64_t room = n_bytes - (g + q * 8);
            const int bytes = room >= 8 ? 8 : (room > 0 ? (int)room : 0);
            cp_async_small_zfill<8>(d + q * 8, bytes > 0 ? in + g + q * 8 : in, bytes);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int64_t room = n_bytes - (g + q * 4);
            const int bytes = room >= 4 ? 4 : (room > 0 ? (int)room : 0);
            cp_async_small_zfill<4>(d + q * 4, bytes > 0 ? in + g + q * 4 : in, bytes);
          }
        }
      }
    } else if (g0 - (int64_t)lane * 16 + 4096 <= n_bytes) {
#pragma unroll
      for (int k = 0; k < kScRows; ++k) cp_async16(dst0 + (str_even ^ ((k & 1) << 6)) + k * 512, in + g0 + k * 512);
    } else {  // the segment crosses the end of the input: zero-fill what is not there
#pragma unroll
      for (int k = 0; k < kScRows; ++k) {
        const int64_t g = g0 + k * 512;
        const int64_t room = n_bytes - g;
        const int bytes = room >= 16 ? 16 : (room > 0 ? (int)room : 0);
        cp_async16_zfill(dst0 + (str_even ^ ((k & 1) << 6)) + k * 512, bytes > 0 ? in + g : in, bytes);
      }
    }
  };

  const int64_t G = gridDim.x;
  const int64_t mine = ((int64_t)blockIdx.x < p.n_tiles) ? (p.n_tiles - blockIdx.x + G - 1) / G : 0;  // tiles of this block
  if (!looker) { if (mine > 0) prefetch(blockIdx.x, 0); cp_async_commit(); }
  A lane_excl_cur = P::zero(), warp_excl_cur = P::zero();
  // validity words of the rows this lane owns: fetched once per tile, in its reduction, and carried to its finalization
  // one iteration later; `dense` is made warp-uniform so a warp runs ONE of the two loops, not both under divergence
  unsigned vb_cur[VW], vb_prev[VW];
  bool dense_cur = true, dense_prev = true;
#pragma unroll
  for (int w = 0; w < VW; ++w) { vb_cur[w] = 0u; vb_prev[w] = 0u; }
  int st_r = 0, st_f = kScStreamStages - 1;   // stage of the tile being reduced (j % 3) / finalized ((j-1) % 3)
  // iteration j: reduce tile j (j < mine), look-back + finalize tile j-1 (j >= 1)
  for (int64_t j = 0; j <= mine; ++j) {
    const int64_t tile_r = blockIdx.x + j * G, tile_f = tile_r - G;
    const int br = (int)(j & 1), bf = br ^ 1;
    A lane_excl_next = P::zero();
    if (!looker) {
      const int st_n = st_r + 1 == kScStreamStages ? 0 : st_r + 1;
      if (j + 1 < mine) prefetch(tile_r + G, st_n);  // that stage held tile j-2: finalized, by this warp, an iteration ago
      cp_async_commit();
      if (j < mine) {
        cp_async_wait<1>();
        __syncwarp();
        const unsigned char* seg = s_ring + st_r * kScTileBytes + seg_off;
        uint4 raw[kScRows];
#pragma unroll
        for (int c = 0; c < kScRows; ++c) raw[c] = *reinterpret_cast<const uint4*>(seg + blk_row + (((uint32_t)c << 4) ^ blk_x));
        const T* o = reinterpret_cast<const T*>(raw);
        A tot = P::zero();
        dense_cur = __all_sync(0xffffffffu, lane_validity(tile_r, vb_cur));
        if (dense_cur) {
#pragma unroll
          for (int i = 0; i < E; ++i) tot = P::add_elem(tot, o[i]);
        } else {
#pragma unroll
          for (int i = 0; i < E; ++i) if ((vb_cur[i >> 5] >> (i & 31)) & 1u) tot = P::add_elem(tot, o[i]);
        }
        A incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const A up = P::shfl_up(incl, d);
          if (lane >= d) incl = P::add(up, incl);
        }
        lane_excl_next = P::shfl_up(incl, 1);
        if (lane == 0) lane_excl_next = P::zero();
        if (lane == 31) {
          // the LAST warp to deliver its total publishes the tile aggregate: no barrier between a tile's reduction
          // and the moment other blocks can see it (their look-backs wait on exactly this word)
          s_warp[br][warp] = incl;
          __threadfence_block();
          if (atomicAdd(&s_arrived[br], 1u) == kScWarps - 1) {
            __threadfence_block();
            s_arrived[br] = 0u;
            A tile_total = P::zero();
#pragma unroll
            for (int wi = 0; wi < kScWarps; ++wi) tile_total = P::add(tile_total, reinterpret_cast<volatile A*>(s_warp[br])[wi]);
            unsigned w[3] = {0u, 0u, 0u};
            P::to_words(tile_total, w);
#pragma unroll
            for (int k = 0; k < P::K; ++k) st_word(p.agg + tile_r * P::K + k, w[k]);
          }
        }
      }
    } else if (j >= 1) {
      A tile_total = P::zero();
#pragma unroll
      for (int wi = 0; wi < kScWarps; ++wi) tile_total = P::add(tile_total, s_warp[bf][wi]);
      const A excl = scan_lookback_lean<P>(p, tile_f, tile_total, start, lane);
      if (lane == 0) {
        s_excl[bf] = excl;
        if (tile_f == p.n_tiles - 1) {
          CumsumState ns = s_state;
          P::to_state(P::add(excl, tile_total), &ns);
          if (kHasValid && *p.first_null < p.n) ns.encountered_null = 1;
          *p.state = ns;
        }
      }
    }
    __syncthreads();
    if (!looker) {
      A warp_excl_next = P::zero();
      if (j < mine) {
        A tile_total = P::zero();
#pragma unroll
        for (int wi = 0; wi < kScWarps; ++wi) {
          if (wi == warp) warp_excl_next = tile_total;
          tile_total = P::add(tile_total, s_warp[br][wi]);
        }
      }
      if (j >= 1) {
        unsigned char* seg = s_ring + st_f * kScTileBytes + seg_off;
        uint4 raw[kScRows];
#pragma unroll
        for (int c = 0; c < kScRows; ++c) raw[c] = *reinterpret_cast<const uint4*>(seg + blk_row + (((uint32_t)c << 4) ^ blk_x));
        T* o = reinterpret_cast<T*>(raw);
        A run = P::add(warp_excl_cur, lane_excl_cur);
        const T off = P::value(s_excl[bf]);
        if (dense_prev) {
#pragma unroll
          for (int i = 0; i < E; ++i) { run = P::add_elem(run, o[i]); o[i] = P::offset_add(off, P::value(run)); }
        } else {
#pragma unroll
          for (int i = 0; i < E; ++i) {
            if ((vb_prev[i >> 5] >> (i & 31)) & 1u) { run = P::add_elem(run, o[i]); o[i] = P::offset_add(off, P::value(run)); }
            else o[i] = T(0);
          }
        }
#pragma unroll
        for (int c = 0; c < kScRows; ++c) *reinterpret_cast<uint4*>(seg + blk_row + (((uint32_t)c << 4) ^ blk_x)) = raw[c];
        __syncwarp();
        const int64_t g0 = tile_f * kScTileBytes + seg_off + (int64_t)lane * 16;
        if (g0 - (int64_t)lane * 16 + 4096 <= n_bytes) {
#pragma unroll
          for (int k = 0; k < kScRows; ++k)
            __stcs(reinterpret_cast<uint4*>(out + g0 + k * 512), *reinterpret_cast<const uint4*>(seg + (str_even ^ ((k & 1) << 6)) + k * 512));
        } else {
#pragma unroll
          for (int k = 0; k < kScRows; ++k) {
            const int64_t g = g0 + k * 512;
            const uint4 v = *reinterpret_cast<const uint4*>(seg + (str_even ^ ((k & 1) << 6)) + k * 512);
            if (g + 16 <= n_bytes) __stcs(reinterpret_cast<uint4*>(out + g), v);
            else {
              const unsigned char* vb = reinterpret_cast<const unsigned char*>(&v);
              for (int q = 0; q < 16; ++q) if (g + q < n_bytes) out[g + q] = vb[q];
            }
          }
        }
        __syncwarp();  // the segment is refilled (cp.async) two iterations from now, by this warp
      }
      lane_excl_cur = lane_excl_next;
      warp_excl_cur = warp_excl_next;
      dense_prev = dense_cur;
#pragma unroll
      for (int w = 0; w < VW; ++w) vb_prev[w] = vb_cur[w];
    }
    st_f = st_r;
    st_r = st_r + 1 == kScStreamStages ? 0 : st_r + 1;
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
  WorkspaceLock ws_lock(ws);
  p.n_tiles = (p.n + kTileRows - 1) / kTileRows;
  const int64_t n_groups = (p.n_tiles + kScGroup - 1) / kScGroup;
  const int64_t n_super = (n_groups + kScSuper - 1) / kScSuper;
  // every level starts on a 16-byte boundary (the streaming kernel fetches whole entries with 128-bit loads)
  const size_t agg_words = ((size_t)p.n_tiles * kStatusK + 1) & ~(size_t)1, gagg_words = ((size_t)n_groups * kStatusK + 1) & ~(size_t)1;
  const size_t words = agg_words + gagg_words + (size_t)n_super * kStatusK;
  AG_TRY(ensure_tile_status(ws, words, st));
  p.agg = ws->tile_status;
  p.gagg = p.agg + agg_words;
  p.sincl = p.gagg + gagg_words;
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
  // the streaming kernel wants a 16-byte aligned OUTPUT; an input that is only element-aligned (a slice) is fetched in
  // 4- / 8-byte pieces, which needs elements of at least 4 bytes
  const bool vec = ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0 || sizeof(T) >= 4) && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
  void* args[] = {(void*)&p};
  const void* fn;
  const bool chk = p.checked && !IsFp<T>::v;  // floats: the checked adder is the plain one
  if (vec && !chk) {
    constexpr int kRing = kScStreamStages * kScTileBytes;
    static std::atomic<unsigned> attr_set{0u};  // per instantiation, one bit per device
    const void* sfn = p.valid ? (const void*)cumsum_stream_kernel<T, true> : (const void*)cumsum_stream_kernel<T, false>;
    static std::atomic<unsigned> attr_set_v{0u};
    AG_TRY(ensure_dynamic_smem(sfn, kRing, p.valid ? &attr_set_v : &attr_set));
    int per_sm = 0;
    AG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sfn, kScStreamThreads, kRing));
    if (per_sm < 1) per_sm = 1;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    const int grid = (int)(p.n_tiles < cap ? p.n_tiles : cap);
    AG_CUDA_TRY(cudaLaunchCooperativeKernel(sfn, dim3(grid), dim3(kScStreamThreads), args, kRing, st));
    return check_launch("cumsum_stream_kernel");
  }
  if (chk) {
    if (vec) fn = p.valid ? (const void*)cumsum_kernel<T, true, true, !IsFp<T>::v> : (const void*)cumsum_kernel<T, true, false, !IsFp<T>::v>;
    else fn = p.valid ? (const void*)cumsum_kernel<T, false, true, !IsFp<T>::v> : (const void*)cumsum_kernel<T, false, false, !IsFp<T>::v>;
  } else {
    if (vec) fn = p.valid ? (const void*)cumsum_kernel<T, true, true, false> : (const void*)cumsum_kernel<T, true, false, false>;
    else fn = p.valid ? (const void*)cumsum_kernel<T, false, true, false> : (const void*)cumsum_kernel<T, false, false, false>;
  }
  const int per_sm = blocks_per_sm(fn, kScThreads);
  const int64_t cap = (int64_t)sm_count() * per_sm;
  const int grid = (int)(p.n_tiles < cap ? p.n_tiles : cap);
  AG_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kScThreads), args, 0, st));
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
