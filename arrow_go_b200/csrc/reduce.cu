// reduce.cu — arrow/math Sum on sm_100a.
//
// Replaces sum_{float64,int64,uint64}_{avx2,sse4,neon} (arrow/math/_lib/float64.c:20-26,
// int64.c:21-27, uint64.c; Go entry Float64Funcs.Sum arrow/math/float64.go:34-39).
// Validity bitmaps are ignored, exactly like the reference (it sums values[offset:offset+len]).
//
// Roofline: HBM.  8 algorithmic bytes per row, ~0.125 flop/byte.
//
// Shape of the reduction (fixed => the float64 result is a pure function of (data, n)):
//   * the input is viewed as PAIRS (x[2j], x[2j+1]); an odd last element is folded in at the end;
//   * G = sum_grid(n) blocks of 256 threads (at most 148 x 8, a constant); block b streams the
//     contiguous 2048-pair (32 KB) tiles b, b+G, ...; in a tile thread t owns pairs t, t+256, ...
//     with 8 pair-loads in flight, load k feeding accumulator pair k % 4 (8 chains per thread);
//   * thread sum -> warp xor-shuffle tree -> block tree (fixed) -> partials[G];
//   * the last block to finish (ticket) reduces partials[] with the same fixed tree.
// 16-byte aligned inputs use 128-bit ld.global.cs; 8-byte aligned ones use two 64-bit loads
// with the SAME pair->thread mapping, so alignment never changes the result.
#include "common.cuh"

#include <stdlib.h>
#include <cuda/barrier>

namespace ag {

constexpr int kSumThreads = 256;
constexpr int kSumLoads = 8;                                  // 16-byte pair loads in flight per thread
constexpr int kSumAcc = 4;                                    // accumulator pairs per thread (slot = load % 4)
constexpr int kSumTilePairs = kSumThreads * kSumLoads;        // 2048 pairs = 32 KB per block iteration
constexpr int kSumBlocksPerSM = 8;                            // measured at 100M rows: 296..888 blocks 126-135 us, 1184 blocks 123.6 us
constexpr int kSumMaxBlocks = 148 * kSumBlocksPerSM;          // fixed (not queried): the result must not depend on the device

static inline int sum_grid(size_t n_pairs) {
  // small inputs get few blocks (latency), big ones the fixed one-wave grid
  size_t want = (n_pairs + kSumTilePairs - 1) / kSumTilePairs;
  if (want < 1) want = 1;
  static int cap = 0;
  if (cap == 0) {
    const char* e = getenv("AG_SUM_BLOCKS");  // experiments only
    cap = (e && atoi(e) > 0) ? atoi(e) : kSumMaxBlocks;
    if (cap > kMaxPartials) cap = kMaxPartials;
  }
  if (want > (size_t)cap) want = cap;
  return (int)want;
}

template <typename T> struct Pair { T x, y; };

template <typename T, bool kAligned>
__device__ __forceinline__ Pair<T> load_pair(const T* __restrict__ in, size_t j) {
  Pair<T> p;
  if (kAligned) {
    const ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2*>(in) + j);
    p.x = *reinterpret_cast<const T*>(&v.x);
    p.y = *reinterpret_cast<const T*>(&v.y);
  } else {
    p.x = __ldcs(in + 2 * j);
    p.y = __ldcs(in + 2 * j + 1);
  }
  return p;
}

template <typename T>
__device__ __forceinline__ T shfl_xor_t(T v, int m) {
  unsigned long long u = *reinterpret_cast<unsigned long long*>(&v);
  u = __shfl_xor_sync(0xffffffffu, u, m);
  return *reinterpret_cast<T*>(&u);
}

// fixed tree over the 256 threads of a block; result valid in thread 0
template <typename T>
__device__ __forceinline__ T block_tree_sum(T v, T* smem /* >= 8 */) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = v + shfl_xor_t(v, m);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  T r = T(0);
  if (warp == 0) {
    r = (lane < kSumThreads / 32) ? smem[lane] : T(0);
#pragma unroll
    for (int m = 4; m >= 1; m >>= 1) r = r + shfl_xor_t(r, m);
  }
  __syncthreads();
  return r;
}

template <typename T, bool kAligned>
__global__ void __launch_bounds__(kSumThreads)
sum_kernel(const T* __restrict__ in, size_t n, T* __restrict__ partials, unsigned* __restrict__ ticket,
           T* __restrict__ out) {
  __shared__ T smem[8];
  __shared__ bool is_last;
  const size_t n_pairs = n >> 1;
  const size_t n_tiles = (n_pairs + kSumTilePairs - 1) / kSumTilePairs;
  T ax[kSumAcc], ay[kSumAcc];
#pragma unroll
  for (int k = 0; k < kSumAcc; ++k) { ax[k] = T(0); ay[k] = T(0); }
  // Block b streams the contiguous 32 KB tiles b, b+G, b+2G, ...; inside a tile thread t owns
  // pairs t, t+256, ... (8 independent 16-byte loads in flight), load k feeds accumulator k % 4.
  for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t j0 = tile * kSumTilePairs + threadIdx.x;
    if (j0 + (size_t)(kSumLoads - 1) * kSumThreads < n_pairs) {
      Pair<T> p[kSumLoads];
#pragma unroll
      for (int k = 0; k < kSumLoads; ++k) p[k] = load_pair<T, kAligned>(in, j0 + (size_t)k * kSumThreads);
#pragma unroll
      for (int k = 0; k < kSumLoads; ++k) { ax[k % kSumAcc] = ax[k % kSumAcc] + p[k].x; ay[k % kSumAcc] = ay[k % kSumAcc] + p[k].y; }
    } else {
#pragma unroll
      for (int k = 0; k < kSumLoads; ++k) {
        const size_t jj = j0 + (size_t)k * kSumThreads;
        if (jj < n_pairs) {
          const Pair<T> p = load_pair<T, kAligned>(in, jj);
          ax[k % kSumAcc] = ax[k % kSumAcc] + p.x; ay[k % kSumAcc] = ay[k % kSumAcc] + p.y;
        }
      }
    }
  }
  T v = ((ax[0] + ay[0]) + (ax[1] + ay[1])) + ((ax[2] + ay[2]) + (ax[3] + ay[3]));
  v = block_tree_sum(v, smem);

  if (gridDim.x == 1) {
    if (threadIdx.x == 0) { if (n & 1) v = v + in[n - 1]; *out = v; }
    return;
  }
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = v;
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // fixed-order reduction of the partials: thread t takes t, t+256, ... then the block tree
  T acc = T(0);
  for (unsigned i = threadIdx.x; i < gridDim.x; i += kSumThreads) acc = acc + __ldcg(partials + i);
  acc = block_tree_sum(acc, smem);
  if (threadIdx.x == 0) {
    if (n & 1) acc = acc + in[n - 1];
    *out = acc;
    *ticket = 0;  // leave the workspace ready for the next launch on this stream
  }
}


// ---------------------------------------------------------------- TMA-staged variant ---------
// Same reduction, but the input is staged through shared memory by the TMA engine instead of
// per-thread loads: one elected thread issues cp.async.bulk (1-D bulk copy, SASS UBLKCP) of a
// 16 KB chunk into a 4-deep ring of shared-memory stages, completion is signalled on an mbarrier
// (expect-tx bytes), and the 256 threads read the stage with conflict-free LDS.128.  Bytes in
// flight per SM no longer depend on registers or occupancy: 3 resident blocks x 4 stages x 16 KB
// = 192 KB.  Requires a 16-byte aligned input; the (< 2048-element) tail is folded in by the last
// block.  Fixed shape (444 blocks, static chunk -> block map) => deterministic like sum_kernel.
namespace cde = cuda::device::experimental;
using block_barrier = cuda::barrier<cuda::thread_scope_block>;

constexpr int kTmaStages = 4;
constexpr int kTmaChunkElems = 2048;  // 16 KB of 8-byte elements
constexpr int kTmaBlocks = 148 * 3;

template <typename T>
__global__ void __launch_bounds__(kSumThreads)
sum_tma_kernel(const T* __restrict__ in, size_t n, T* __restrict__ partials, unsigned* __restrict__ ticket, T* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char tma_smem[];
  T(*buf)[kTmaChunkElems] = reinterpret_cast<T(*)[kTmaChunkElems]>(tma_smem);
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ block_barrier bar[kTmaStages];
  __shared__ T smem[8];
  __shared__ bool is_last;
  const size_t n_chunks = n / kTmaChunkElems;
  constexpr unsigned kBytes = kTmaChunkElems * sizeof(T);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) init(&bar[s], kSumThreads);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  block_barrier::arrival_token token[kTmaStages];
  auto issue = [&](int s, size_t chunk) {
    if (threadIdx.x == 0) {
      cde::cp_async_bulk_global_to_shared(buf[s], in + chunk * kTmaChunkElems, kBytes, bar[s]);
      token[s] = cuda::device::barrier_arrive_tx(bar[s], 1, kBytes);
    } else {
      token[s] = bar[s].arrive();
    }
  };
#pragma unroll
  for (int s = 0; s < kTmaStages; ++s) {
    const size_t chunk = blockIdx.x + (size_t)s * gridDim.x;
    if (chunk < n_chunks) issue(s, chunk);
  }
  T ax[kSumAcc], ay[kSumAcc];
#pragma unroll
  for (int k = 0; k < kSumAcc; ++k) { ax[k] = T(0); ay[k] = T(0); }
  for (size_t base = blockIdx.x; base < n_chunks; base += (size_t)kTmaStages * gridDim.x) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) {
      const size_t chunk = base + (size_t)s * gridDim.x;
      if (chunk >= n_chunks) break;
      bar[s].wait(std::move(token[s]));
      const ulonglong2* v2 = reinterpret_cast<const ulonglong2*>(buf[s]);
#pragma unroll
      for (int k = 0; k < kTmaChunkElems / 2 / kSumThreads; ++k) {  // 4 pairs per thread
        const ulonglong2 v = v2[k * kSumThreads + threadIdx.x];
        ax[k % kSumAcc] = ax[k % kSumAcc] + *reinterpret_cast<const T*>(&v.x);
        ay[k % kSumAcc] = ay[k % kSumAcc] + *reinterpret_cast<const T*>(&v.y);
      }
      __syncthreads();  // every thread is done with stage s before it is refilled
      const size_t next = chunk + (size_t)kTmaStages * gridDim.x;
      if (next < n_chunks) issue(s, next);
    }
  }
  T v = ((ax[0] + ay[0]) + (ax[1] + ay[1])) + ((ax[2] + ay[2]) + (ax[3] + ay[3]));
  v = block_tree_sum(v, smem);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = v;
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  T acc = T(0);
  for (unsigned i = threadIdx.x; i < gridDim.x; i += kSumThreads) acc = acc + __ldcg(partials + i);
  acc = block_tree_sum(acc, smem);
  // tail (< one chunk): thread t takes elements t, t+256, ... of the remainder, fixed tree again
  T tail = T(0);
  for (size_t i = n_chunks * kTmaChunkElems + threadIdx.x; i < n; i += kSumThreads) tail = tail + in[i];
  tail = block_tree_sum(tail, smem);
  if (threadIdx.x == 0) {
    *out = acc + tail;
    *ticket = 0;
  }
}

static bool sum_use_tma() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AG_SUM_TMA"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

// Reference association order (float64_avx2_amd64.s:36-43,86-174): 32 interleaved serial
// chains over the first n&~31 elements — lane j of ONE warp owns chain j, so every load is a
// fully coalesced 256-byte row — combined as (y0+y4)+(y2+y6) + (y1+y5)+(y3+y7) per ymm lane,
// lo128+hi128, hadd, then the n&31 tail sequentially.  n < 32: purely sequential.
// Latency-bound by design (n/32 dependent DADDs); bit-exact with the reference on any input.
__global__ void __launch_bounds__(32) sum_f64_reforder_kernel(const double* __restrict__ in, size_t n, double* out) {
  const int lane = threadIdx.x;
  const size_t body = (n > 31) ? (n & ~(size_t)31) : 0;
  double acc = 0.0;
  size_t k = 0;
  // 8 independent loads in flight per lane; the adds stay in chain order
  for (; k + 8 * 32 <= body; k += 8 * 32) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(in + k + u * 32 + lane);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, v[u]);
  }
  for (; k < body; k += 32) acc = __dadd_rn(acc, __ldcs(in + k + lane));
  double s = 0.0;
  if (body) {
    // lane j = 4*y + q  (ymm register y, 64-bit lane q)
    // t_y = acc[y] + acc[y+4]  for y in 0..3            -> xor 16
    double t = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 16));
    // u0 = t0 + t2, u1 = t1 + t3                         -> xor 8 (valid in y<2)
    double u = __dadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 8));
    // v = u0 + u1                                         -> xor 4 (valid in y==0, lanes q=0..3)
    double v = __dadd_rn(u, __shfl_xor_sync(0xffffffffu, u, 4));
    // w0 = v[0]+v[2], w1 = v[1]+v[3]                      -> xor 2
    double w = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
    // hadd: w0 + w1                                       -> xor 1
    s = __dadd_rn(w, __shfl_xor_sync(0xffffffffu, w, 1));
  }
  if (lane == 0) {
    for (size_t i = body; i < n; ++i) s = __dadd_rn(s, in[i]);
    *out = s;
  }
}

template <typename T>
static ag_status launch_sum(const T* d_in, size_t n, T* d_res, cudaStream_t st) {
  if (n == 0) {
    AG_CUDA_TRY(cudaMemsetAsync(d_res, 0, sizeof(T), st));
    return AG_OK;
  }
  if ((reinterpret_cast<uintptr_t>(d_in) & 7) != 0) AG_FAIL(AG_ERR_INVALID, "sum: input is not 8-byte aligned");
  Workspace* ws;
  AG_TRY(get_workspace(st, &ws));
  const bool aligned = (reinterpret_cast<uintptr_t>(d_in) & 15) == 0;
  if (sum_use_tma() && aligned && n >= (size_t)kTmaChunkElems * kTmaBlocks) {
    constexpr size_t smem = (size_t)kTmaStages * kTmaChunkElems * sizeof(T);
    static std::atomic<bool> attr{false};
    if (!attr.load()) {
      AG_CUDA_TRY(cudaFuncSetAttribute((const void*)sum_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr.store(true);
    }
    sum_tma_kernel<T><<<kTmaBlocks, kSumThreads, smem, st>>>(d_in, n, (T*)ws->partials, ws->ticket, d_res);
    return check_launch("sum_tma_kernel");
  }
  const int grid = sum_grid(n >> 1);
  if (aligned)
    sum_kernel<T, true><<<grid, kSumThreads, 0, st>>>(d_in, n, (T*)ws->partials, ws->ticket, d_res);
  else
    sum_kernel<T, false><<<grid, kSumThreads, 0, st>>>(d_in, n, (T*)ws->partials, ws->ticket, d_res);
  return check_launch("sum_kernel");
}

// Host-pointer flavour: stage the whole column to HBM with chunked async copies (the copy is
// >100x the kernel time, so there is nothing to gain from overlapping the reduction), then run
// exactly the same kernel as the device flavour => identical bits.
template <typename T, typename F>
static ag_status host_sum(const T* buf, size_t n, T* res, F&& launch) {
  if (!res) AG_FAIL(AG_ERR_INVALID, "sum: NULL result pointer");
  AG_TRY(ensure_init());
  if (n == 0) { *res = T(0); return AG_OK; }
  if (!buf) AG_FAIL(AG_ERR_INVALID, "sum: NULL buffer");
  CallStream cs;  // a private pooled stream per call keeps concurrent host calls independent
  AG_TRY(cs.acquire());
  T* d_in = nullptr;
  T* d_res = nullptr;
  ag_status rc = AG_OK;
  do {
    if ((rc = dev_alloc_async((void**)&d_in, n * sizeof(T) + sizeof(T), cs)) != AG_OK) break;
    d_res = d_in + n;
    const size_t chunk = (size_t)8 << 20;  // elements (64 MiB)
    cudaError_t e = cudaSuccess;
    for (size_t off = 0; off < n && e == cudaSuccess; off += chunk) {
      const size_t len = (n - off < chunk) ? (n - off) : chunk;
      e = cudaMemcpyAsync(d_in + off, buf + off, len * sizeof(T), cudaMemcpyHostToDevice, cs);
    }
    if (e != cudaSuccess) { rc = cuda_fail(e, "H2D", __FILE__, __LINE__); break; }
    if ((rc = launch(d_in, n, d_res, cs)) != AG_OK) break;
    e = cudaMemcpyAsync(res, d_res, sizeof(T), cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
    if (e != cudaSuccess) { rc = cuda_fail(e, "D2H", __FILE__, __LINE__); break; }
  } while (0);
  if (d_in) cudaFreeAsync(d_in, cs);
  cudaStreamSynchronize(cs);
  return rc;
}

static ag_status launch_reforder(const double* d_in, size_t n, double* d_res, cudaStream_t st) {
  if (n == 0) { AG_CUDA_TRY(cudaMemsetAsync(d_res, 0, sizeof(double), st)); return AG_OK; }
  sum_f64_reforder_kernel<<<1, 32, 0, st>>>(d_in, n, d_res);
  return check_launch("sum_f64_reforder_kernel");
}

}  // namespace ag

using namespace ag;

extern "C" {

ag_status ag_sum_f64_dev(const double* d, size_t n, double* d_res, ag_stream_t s) {
  AG_TRY(ensure_init());
  return launch_sum<double>(d, n, d_res, resolve_stream(s));
}
ag_status ag_sum_i64_dev(const int64_t* d, size_t n, int64_t* d_res, ag_stream_t s) {
  AG_TRY(ensure_init());
  // wrapping two's-complement sum == unsigned sum (int64.c:21-27)
  return launch_sum<unsigned long long>((const unsigned long long*)d, n, (unsigned long long*)d_res, resolve_stream(s));
}
ag_status ag_sum_u64_dev(const uint64_t* d, size_t n, uint64_t* d_res, ag_stream_t s) {
  AG_TRY(ensure_init());
  return launch_sum<unsigned long long>((const unsigned long long*)d, n, (unsigned long long*)d_res, resolve_stream(s));
}
ag_status ag_sum_f64_reforder_dev(const double* d, size_t n, double* d_res, ag_stream_t s) {
  AG_TRY(ensure_init());
  return launch_reforder(d, n, d_res, resolve_stream(s));
}

ag_status ag_sum_f64(const double* buf, size_t n, double* res) {
  return host_sum<double>(buf, n, res, [](const double* d, size_t m, double* r, cudaStream_t st) { return launch_sum<double>(d, m, r, st); });
}
ag_status ag_sum_i64(const int64_t* buf, size_t n, int64_t* res) {
  return host_sum<unsigned long long>((const unsigned long long*)buf, n, (unsigned long long*)res,
      [](const unsigned long long* d, size_t m, unsigned long long* r, cudaStream_t st) { return launch_sum<unsigned long long>(d, m, r, st); });
}
ag_status ag_sum_u64(const uint64_t* buf, size_t n, uint64_t* res) {
  return host_sum<unsigned long long>((const unsigned long long*)buf, n, (unsigned long long*)res,
      [](const unsigned long long* d, size_t m, unsigned long long* r, cudaStream_t st) { return launch_sum<unsigned long long>(d, m, r, st); });
}
ag_status ag_sum_f64_reforder(const double* buf, size_t n, double* res) {
  return host_sum<double>(buf, n, res, [](const double* d, size_t m, double* r, cudaStream_t st) { return launch_reforder(d, m, r, st); });
}

}  // extern "C"
