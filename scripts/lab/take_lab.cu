// take_lab.cu — standalone experiment bench for the Take (random gather) redesign.
// Not part of the product; the winning shape moves into arrow_go_b200/csrc/take.cu.
//   direct gather with U loads in flight, under different cudaLimitMaxL2FetchGranularity values
//   windowed gather: A (tile-local bucket partition) -> B (bucket-major gather) -> C (tile un-permute)
// usage: take_lab <vlen> <n> [window_mb] [reps]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __host__ inline uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void fill_vals(uint64_t* v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = mix64((uint64_t)i);
}
__global__ void fill_idx(uint32_t* ix, int64_t n, uint64_t vlen, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    ix[i] = (uint32_t)(mix64((uint64_t)i * 0x9e3779b97f4a7c15ull + seed) % vlen);
}
__global__ void check_out(const uint64_t* out, const uint32_t* ix, int64_t n, unsigned long long* bad) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (out[i] != mix64((uint64_t)ix[i])) atomicAdd(bad, 1ull);
}
__global__ void flush_l2(uint4* p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = make_uint4(1, 2, 3, 4);
}

// ---------------------------------------------------------------- direct gather
template <int U, int THREADS>
__global__ void __launch_bounds__(THREADS) direct_kernel(const uint64_t* __restrict__ vals, const uint32_t* __restrict__ idx,
                                                         uint64_t* __restrict__ out, int64_t n, uint64_t vlen) {
  const int64_t step = (int64_t)gridDim.x * THREADS * U;
  for (int64_t base = (int64_t)blockIdx.x * THREADS * U; base < n; base += step) {
    uint32_t ix[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int64_t i = base + k * THREADS + threadIdx.x;
      ix[k] = i < n ? __ldcs(idx + i) : 0u;
    }
    uint64_t v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = ix[k] < vlen ? vals[ix[k]] : 0ull;
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int64_t i = base + k * THREADS + threadIdx.x;
      if (i < n) __stcs(out + i, v[k]);
    }
  }
}

// ---------------------------------------------------------------- windowed gather
constexpr int T = 8192;          // rows per tile
constexpr int A_THREADS = 512;   // 16 rows per thread
constexpr int A_PER = T / A_THREADS;
constexpr int MAXB = 1024;

// Pass A: tile-local partition by table window.  sorted[tile*T + pos] = idx & (W-1); perm[row] = pos;
// off_t[b * ntiles + tile] = first position of bucket b in the tile (b in [0, nb]; row nb = tile length).
__global__ void __launch_bounds__(A_THREADS) passA(const uint32_t* __restrict__ idx, int64_t n, uint64_t vlen, int shift, int nb,
                                                   uint32_t* __restrict__ sorted, uint16_t* __restrict__ perm,
                                                   uint16_t* __restrict__ off_t, int64_t ntiles) {
  __shared__ uint32_t hist[MAXB + 1];
  __shared__ uint32_t s_sorted[T];
  __shared__ uint32_t wsum[A_THREADS / 32];
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * T;
    const int len = (int)min((int64_t)T, n - base);
    for (int b = threadIdx.x; b <= nb; b += A_THREADS) hist[b] = 0;
    __syncthreads();
    uint32_t ix[A_PER];
    uint32_t rk[A_PER];
#pragma unroll
    for (int k = 0; k < A_PER; ++k) {
      const int r = k * A_THREADS + threadIdx.x;
      ix[k] = r < len ? __ldcs(idx + base + r) : 0xffffffffu;
    }
#pragma unroll
    for (int k = 0; k < A_PER; ++k) {
      const int r = k * A_THREADS + threadIdx.x;
      if (r < len) {
        const uint32_t b = ix[k] < vlen ? (ix[k] >> shift) : (uint32_t)nb;  // out-of-range -> no bucket (tail)
        rk[k] = atomicAdd(&hist[b], 1u);
      }
    }
    __syncthreads();
    // exclusive scan of hist[0..nb] (nb+1 entries) by the whole block: each thread owns ceil((nb+1)/A_THREADS) entries (<= 3)
    {
      constexpr int PER = (MAXB + 1 + A_THREADS - 1) / A_THREADS;
      uint32_t loc[PER];
      uint32_t s = 0;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int b = threadIdx.x * PER + j;
        loc[j] = b <= nb ? hist[b] : 0u;
        s += loc[j];
      }
      uint32_t inc = s;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if ((threadIdx.x & 31) >= d) inc += o;
      }
      if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
      __syncthreads();
      uint32_t wbase = 0;
      for (int w = 0; w < (threadIdx.x >> 5); ++w) wbase += wsum[w];
      uint32_t run = wbase + inc - s;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int b = threadIdx.x * PER + j;
        if (b <= nb) {
          hist[b] = run;
          off_t[(int64_t)b * ntiles + tile] = (uint16_t)run;
        }
        run += loc[j];
      }
    }
    __syncthreads();
    const uint32_t wmask = (1u << shift) - 1u;
#pragma unroll
    for (int k = 0; k < A_PER; ++k) {
      const int r = k * A_THREADS + threadIdx.x;
      if (r < len) {
        const uint32_t b = ix[k] < vlen ? (ix[k] >> shift) : (uint32_t)nb;
        const uint32_t pos = hist[b] + rk[k];
        s_sorted[pos] = ix[k] & wmask;
        perm[base + r] = (uint16_t)pos;
      }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < len; r += A_THREADS) sorted[base + r] = s_sorted[r];
    __syncthreads();
  }
}

// Pass B: bucket-major sweep.  Work item = (bucket b, tile): the run sorted[tile*T + off[b][tile] .. off[b+1][tile]).
// Items are dealt round-robin to warps in (b, tile) order so every warp of the grid is inside the same table
// window at the same time -> the window stays L2-resident.
template <int WPB>
__global__ void __launch_bounds__(WPB * 32) passB(const uint64_t* __restrict__ vals, const uint32_t* __restrict__ sorted,
                                                  const uint16_t* __restrict__ off_t, int64_t ntiles, int nb, int shift,
                                                  uint64_t* __restrict__ gathered) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * WPB + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WPB;
  const int64_t items = (int64_t)nb * ntiles;
  for (int64_t it = warp; it < items; it += nwarps) {
    const int64_t b = it / ntiles;
    const int64_t tile = it - b * ntiles;
    const int o0 = off_t[it];
    const int o1 = off_t[it + ntiles];
    const uint64_t* __restrict__ win = vals + ((uint64_t)b << shift);
    const int64_t tb = tile * T;
    for (int j = o0 + lane; j < o1; j += 128) {
      uint32_t l0, l1 = 0, l2 = 0, l3 = 0;
      l0 = sorted[tb + j];
      const bool h1 = j + 32 < o1, h2 = j + 64 < o1, h3 = j + 96 < o1;
      if (h1) l1 = sorted[tb + j + 32];
      if (h2) l2 = sorted[tb + j + 64];
      if (h3) l3 = sorted[tb + j + 96];
      const uint64_t v0 = win[l0];
      const uint64_t v1 = h1 ? win[l1] : 0;
      const uint64_t v2 = h2 ? win[l2] : 0;
      const uint64_t v3 = h3 ? win[l3] : 0;
      gathered[tb + j] = v0;
      if (h1) gathered[tb + j + 32] = v1;
      if (h2) gathered[tb + j + 64] = v2;
      if (h3) gathered[tb + j + 96] = v3;
    }
  }
}

// Pass B, dynamic: blocks draw chunks of CH consecutive items from a global counter, so the in-flight items are
// always one contiguous range of the (bucket, tile) sequence.
template <int WPB, int CH>
__global__ void __launch_bounds__(WPB * 32) passB_dyn(const uint64_t* __restrict__ vals, const uint32_t* __restrict__ sorted,
                                                      const uint16_t* __restrict__ off_t, int64_t ntiles, int nb, int shift,
                                                      uint64_t* __restrict__ gathered, unsigned long long* counter) {
  __shared__ unsigned long long s_next;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t items = (int64_t)nb * ntiles;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_next = atomicAdd(counter, (unsigned long long)(WPB * CH));
    __syncthreads();
    const int64_t first = (int64_t)s_next;
    if (first >= items) break;
#pragma unroll 1
    for (int c = 0; c < CH; ++c) {
      const int64_t it = first + c * WPB + wib;
      if (it >= items) break;
      const int64_t b = it / ntiles;
      const int64_t tile = it - b * ntiles;
      const int o0 = off_t[it];
      const int o1 = off_t[it + ntiles];
      const uint64_t* __restrict__ win = vals + ((uint64_t)b << shift);
      const int64_t tb = tile * T;
      for (int j = o0 + lane; j < o1; j += 128) {
        uint32_t l0, l1 = 0, l2 = 0, l3 = 0;
        l0 = sorted[tb + j];
        const bool h1 = j + 32 < o1, h2 = j + 64 < o1, h3 = j + 96 < o1;
        if (h1) l1 = sorted[tb + j + 32];
        if (h2) l2 = sorted[tb + j + 64];
        if (h3) l3 = sorted[tb + j + 96];
        const uint64_t v0 = win[l0];
        const uint64_t v1 = h1 ? win[l1] : 0;
        const uint64_t v2 = h2 ? win[l2] : 0;
        const uint64_t v3 = h3 ? win[l3] : 0;
        gathered[tb + j] = v0;
        if (h1) gathered[tb + j + 32] = v1;
        if (h2) gathered[tb + j + 64] = v2;
        if (h3) gathered[tb + j + 96] = v3;
      }
    }
  }
}

// Pass C: un-permute each tile in place: out[row] = gathered[perm[row]].
constexpr int C_THREADS = 512;
__global__ void __launch_bounds__(C_THREADS) passC(uint64_t* __restrict__ out, const uint16_t* __restrict__ perm, int64_t n, int64_t ntiles) {
  extern __shared__ uint64_t s_vals[];
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * T;
    const int len = (int)min((int64_t)T, n - base);
    uint16_t pm[T / C_THREADS];
#pragma unroll
    for (int k = 0; k < T / C_THREADS; ++k) {
      const int r = k * C_THREADS + threadIdx.x;
      pm[k] = r < len ? perm[base + r] : 0;
    }
    for (int r = threadIdx.x; r < len; r += C_THREADS) s_vals[r] = __ldcs(out + base + r);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < T / C_THREADS; ++k) {
      const int r = k * C_THREADS + threadIdx.x;
      if (r < len) __stcs(out + base + r, s_vals[pm[k]]);
    }
    __syncthreads();
  }
}

static float time_it(cudaStream_t st, int reps, void (*fn)(void*), void* arg, uint4* flush, int64_t flush_n) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps + 1; ++r) {
    flush_l2<<<1184, 256, 0, st>>>(flush, flush_n);
    CK(cudaEventRecord(e0, st));
    fn(arg);
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0) best = std::min(best, ms);
  }
  return best;
}

struct Ctx {
  const uint64_t* vals; const uint32_t* idx; uint64_t* out; int64_t n; uint64_t vlen;
  uint32_t* sorted; uint16_t* perm; uint16_t* off_t; int64_t ntiles; int nb; int shift; int sms; int which; int bgrid; unsigned long long* counter;
};

template <int U, int TH> static void run_direct(void* a) {
  Ctx* c = (Ctx*)a;
  int bps = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, direct_kernel<U, TH>, TH, 0));
  direct_kernel<U, TH><<<c->sms * bps, TH>>>(c->vals, c->idx, c->out, c->n, c->vlen);
}
static void run_A(void* a) { Ctx* c = (Ctx*)a; passA<<<c->sms * 2, A_THREADS>>>(c->idx, c->n, c->vlen, c->shift, c->nb, c->sorted, c->perm, c->off_t, c->ntiles); }
static void run_B(void* a) { Ctx* c = (Ctx*)a; passB<8><<<c->bgrid, 256>>>(c->vals, c->sorted, c->off_t, c->ntiles, c->nb, c->shift, c->out); }
static void run_Bd(void* a) { Ctx* c = (Ctx*)a; cudaMemsetAsync(c->counter, 0, 8, 0); passB_dyn<8, 4><<<c->bgrid, 256>>>(c->vals, c->sorted, c->off_t, c->ntiles, c->nb, c->shift, c->out, c->counter); }
static void run_C(void* a) { Ctx* c = (Ctx*)a; passC<<<c->sms * 3, C_THREADS, T * 8>>>(c->out, c->perm, c->n, c->ntiles); }
static void run_ABC(void* a) { run_A(a); run_B(a); run_C(a); }
static void run_ABdC(void* a) { run_A(a); run_Bd(a); run_C(a); }

int main(int argc, char** argv) {
  const uint64_t vlen = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
  const int64_t n = argc > 2 ? atoll(argv[2]) : 100000000ll;
  const int reps = argc > 4 ? atoi(argv[4]) : 3;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, L2 %d MB; vlen=%llu n=%lld\n", prop.name, sms, prop.l2CacheSize >> 20, (unsigned long long)vlen, (long long)n);
  size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("default MaxL2FetchGranularity = %zu\n", lim);
  uint64_t *vals, *out; uint32_t *idx, *sorted; uint16_t *perm, *off_t; unsigned long long* bad; uint4* flush;
  const int64_t flush_n = (256ll << 20) / 16;
  CK(cudaMalloc(&vals, vlen * 8)); CK(cudaMalloc(&out, n * 8)); CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&sorted, n * 4));
  CK(cudaMalloc(&perm, n * 2)); CK(cudaMalloc(&bad, 8)); CK(cudaMalloc(&flush, flush_n * 16));
  const int64_t ntiles = (n + T - 1) / T;
  CK(cudaMalloc(&off_t, (size_t)(MAXB + 2) * ntiles * 2));
  CK(cudaFuncSetAttribute(passC, cudaFuncAttributeMaxDynamicSharedMemorySize, T * 8));
  fill_vals<<<1184, 256>>>(vals, (int64_t)vlen);
  fill_idx<<<1184, 256>>>(idx, n, vlen, 0x0ff1ce);
  CK(cudaDeviceSynchronize());
  unsigned long long* counter; CK(cudaMalloc(&counter, 8));
  Ctx c{vals, idx, out, n, vlen, sorted, perm, off_t, ntiles, 0, 0, sms, 0, 0, counter};
  const bool sweep = getenv("SWEEP") != nullptr;
  if (sweep) {
    for (size_t g : {64, 32}) {
      cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
      for (uint64_t mb : {2ull, 4ull, 8ull, 16ull, 32ull, 48ull, 64ull, 96ull, 128ull, 192ull, 256ull, 384ull, 512ull, 768ull}) {
        const uint64_t vl = (mb << 20) / 8;
        if (vl > vlen) break;
        fill_idx<<<1184, 256>>>(idx, n, vl, 0x0ff1ce);
        c.vlen = vl;
        float ms = time_it(0, reps, run_direct<8, 256>, &c, flush, flush_n);
        printf("gran %zu table %5llu MB: %8.3f ms %7.2f Grows/s\n", g, (unsigned long long)mb, ms, n / ms / 1e6);
      }
    }
    c.vlen = vlen;
    fill_idx<<<1184, 256>>>(idx, n, vlen, 0x0ff1ce);
    CK(cudaDeviceSynchronize());
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 64);
  }
  auto verify = [&](const char* what) {
    CK(cudaMemset(bad, 0, 8));
    check_out<<<1184, 256>>>(out, idx, n, bad);
    unsigned long long h; CK(cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost));
    if (h) printf("   !! %s: %llu mismatches\n", what, h);
  };
  auto report = [&](const char* name, float ms) {
    printf("%-44s %8.3f ms  %7.2f Grows/s  alg %7.1f GB/s  frac %.3f\n", name, ms, n / ms / 1e6, 20.0 * n / ms / 1e6, 20.0 * n / ms / 1e6 / 6586.4);
  };
  const size_t grans[] = {0, 32};
  if (!sweep) for (size_t g : grans) {
    if (g) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("-- set gran %zu -> %s, now %zu\n", g, cudaGetErrorString(e), lim); }
    CK(cudaMemset(out, 0, n * 8));
    report("direct U=4 256thr", time_it(0, reps, run_direct<4, 256>, &c, flush, flush_n)); verify("direct4");
    report("direct U=8 256thr", time_it(0, reps, run_direct<8, 256>, &c, flush, flush_n));
    report("direct U=16 256thr", time_it(0, reps, run_direct<16, 256>, &c, flush, flush_n)); verify("direct16");
    report("direct U=8 512thr", time_it(0, reps, run_direct<8, 512>, &c, flush, flush_n));
  }
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 64);
  const int wmbs_default[] = {2, 4, 8, 16, 32, 64};
  std::vector<int> wmbs;
  if (argc > 3 && atoi(argv[3]) > 0) wmbs.push_back(atoi(argv[3])); else wmbs.assign(wmbs_default, wmbs_default + 6);
  for (int wmb : wmbs) {
    int shift = 0; while ((8ull << shift) < ((uint64_t)wmb << 20)) ++shift;
    int nb = (int)((vlen + (1ull << shift) - 1) >> shift);
    while (nb > MAXB) { ++shift; nb = (int)((vlen + (1ull << shift) - 1) >> shift); }
    c.shift = shift; c.nb = nb;
    printf("== window %d MB -> shift %d, %d buckets, avg run %.1f rows\n", (int)((8ull << shift) >> 20), shift, nb, (double)T / nb);
    CK(cudaMemset(out, 0, n * 8));
    for (int bg : {2, 4, 8}) {
      c.bgrid = sms * bg;
      float a = time_it(0, reps, run_A, &c, flush, flush_n);
      float b = time_it(0, reps, run_B, &c, flush, flush_n);
      float cc = time_it(0, reps, run_C, &c, flush, flush_n);
      // run_C permutes in place, so rebuild before the full pipeline timing
      float all = time_it(0, reps, run_ABC, &c, flush, flush_n);
      verify("windowed");
      printf("   bgrid=%d/SM  A %.3f  B %.3f  C %.3f ms\n", bg, a, b, cc);
      char nm[64]; snprintf(nm, sizeof nm, "windowed %dMB bgrid %d", wmb, bg);
      report(nm, all);
      float bd = time_it(0, reps, run_Bd, &c, flush, flush_n);
      all = time_it(0, reps, run_ABdC, &c, flush, flush_n);
      verify("windowed-dyn");
      printf("   dynamic B %.3f ms\n", bd);
      snprintf(nm, sizeof nm, "windowed-dyn %dMB bgrid %d", wmb, bg);
      report(nm, all);
    }
  }
  return 0;
}
