// take_lab2.cu — second lab iteration of the windowed Take: load-balanced pass B.
//   A  : tile-local partition by table window -> sorted local indices, perm (u16), off[tile][b]
//   T  : transpose off -> offT[b][tile]
//   B  : blocks claim chunks (bucket b, CT consecutive tiles) in order from a global counter; the runs of a chunk are
//        walked as ONE concatenated sequence, so every lane is busy whatever the run length
//   C  : un-permute each tile in place
// usage: take_lab2 <vlen> <n> [reps]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

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

constexpr int MAXB = 1024;

// ------------------------------------------------------------------ pass A
template <int T, int THREADS>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS) passA(const uint32_t* __restrict__ idx, int64_t n, uint64_t vlen, int shift, int nb,
                                                 uint32_t* __restrict__ sorted, uint16_t* __restrict__ perm,
                                                 uint16_t* __restrict__ off, int64_t ntiles) {
  constexpr int PER = T / THREADS;
  __shared__ uint32_t hist[MAXB + 2];
  __shared__ uint32_t s_sorted[T];
  __shared__ uint32_t wsum[THREADS / 32];
  const int nslots = nb + 2;  // buckets 0..nb-1, bucket nb = no-gather rows, slot nb+1 = tile length
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * T;
    const int len = (int)min((int64_t)T, n - base);
    for (int b = threadIdx.x; b < nslots; b += THREADS) hist[b] = 0;
    __syncthreads();
    uint32_t ix[PER];
    uint32_t rk[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int r = k * THREADS + threadIdx.x;
      ix[k] = r < len ? __ldcs(idx + base + r) : 0xffffffffu;
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int r = k * THREADS + threadIdx.x;
      if (r < len) {
        const uint32_t b = ix[k] < vlen ? (ix[k] >> shift) : (uint32_t)nb;
        rk[k] = atomicAdd(&hist[b], 1u);
      }
    }
    __syncthreads();
    {
      constexpr int SPER = (MAXB + 2 + THREADS - 1) / THREADS;
      uint32_t loc[SPER];
      uint32_t s = 0;
#pragma unroll
      for (int j = 0; j < SPER; ++j) {
        const int b = threadIdx.x * SPER + j;
        loc[j] = b < nslots ? hist[b] : 0u;
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
      for (int j = 0; j < SPER; ++j) {
        const int b = threadIdx.x * SPER + j;
        if (b < nslots) hist[b] = run;
        run += loc[j];
      }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nslots; b += THREADS) off[tile * nslots + b] = (uint16_t)hist[b];
    const uint32_t wmask = (1u << shift) - 1u;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int r = k * THREADS + threadIdx.x;
      if (r < len) {
        const uint32_t b = ix[k] < vlen ? (ix[k] >> shift) : (uint32_t)nb;
        const uint32_t pos = hist[b] + rk[k];
        s_sorted[pos] = ix[k] & wmask;
        perm[base + r] = (uint16_t)pos;
      }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < len; r += THREADS) sorted[base + r] = s_sorted[r];
    __syncthreads();
  }
}

// ------------------------------------------------------------------ transpose off[tile][slot] -> offT[slot][tile_pad]
__global__ void transpose_off(const uint16_t* __restrict__ off, uint16_t* __restrict__ offT, int64_t ntiles, int nslots, int64_t tpad) {
  __shared__ uint16_t tile[32][33];
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  const int s0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t t = t0 + j; const int s = s0 + threadIdx.x;
    tile[j][threadIdx.x] = (t < ntiles && s < nslots) ? off[t * nslots + s] : (uint16_t)0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int s = s0 + j; const int64_t t = t0 + threadIdx.x;
    if (s < nslots && t < tpad) offT[(int64_t)s * tpad + t] = t < ntiles ? tile[threadIdx.x][j] : (uint16_t)0;
  }
}

// ------------------------------------------------------------------ pass B (load-balanced)
template <int T, int THREADS, int CT>
__global__ void __launch_bounds__(THREADS) passB(const uint64_t* __restrict__ vals, const uint32_t* __restrict__ sorted,
                                                 const uint16_t* __restrict__ offT, int64_t tpad, int nb, int shift,
                                                 uint64_t* __restrict__ gathered, unsigned long long* counter) {
  static_assert(CT <= THREADS, "one descriptor per thread");
  __shared__ uint32_t s_start[CT + 1];
  __shared__ long long s_src[CT];
  __shared__ uint32_t s_w[THREADS / 32];
  __shared__ unsigned long long s_claim[2];
  const int64_t cpb = tpad / CT;                 // chunks per bucket (tpad is a multiple of CT)
  const int64_t nchunks = cpb * nb;
  if (threadIdx.x == 0) s_claim[0] = atomicAdd(counter, 1ull);
  __syncthreads();
  for (int it = 0;; ++it) {
    const int64_t chunk = (int64_t)s_claim[it & 1];
    if (chunk >= nchunks) break;
    if (threadIdx.x == 0) s_claim[(it + 1) & 1] = atomicAdd(counter, 1ull);  // claim ahead; read after the next barriers
    const int64_t b = chunk / cpb;
    const int64_t t0 = (chunk - b * cpb) * CT;
    uint32_t len = 0, o0 = 0;
    if (threadIdx.x < CT) {
      o0 = offT[b * tpad + t0 + threadIdx.x];
      len = (uint32_t)offT[(b + 1) * tpad + t0 + threadIdx.x] - o0;
    }
    uint32_t inc = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if ((threadIdx.x & 31) >= d) inc += o;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) wbase += s_w[w];
    const uint32_t start = wbase + inc - len;
    if (threadIdx.x < CT) {
      s_start[threadIdx.x] = start;
      s_src[threadIdx.x] = (long long)(t0 + threadIdx.x) * T + o0 - start;
      if (threadIdx.x == CT - 1) s_start[CT] = start + len;
    }
    __syncthreads();
    const uint32_t total = s_start[CT];
    const uint64_t* __restrict__ win = vals + ((uint64_t)b << shift);
    // first element of this thread: binary search for its run, then walk forward
    int r = 0;
    {
      int lo = 0, hi = CT;  // largest r with s_start[r] <= tid
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_start[mid] <= threadIdx.x) lo = mid; else hi = mid; }
      r = lo;
    }
    for (uint32_t e = threadIdx.x; e < total; e += 4 * THREADS) {
      long long a[4]; uint32_t l[4]; bool h[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t ee = e + k * THREADS;
        h[k] = ee < total;
        if (h[k]) {
          while (s_start[r + 1] <= ee) ++r;
          a[k] = s_src[r] + ee;
          l[k] = sorted[a[k]];
        }
      }
      uint64_t v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) if (h[k]) v[k] = win[l[k]];
#pragma unroll
      for (int k = 0; k < 4; ++k) if (h[k]) gathered[a[k]] = v[k];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ pass C
template <int T, int THREADS>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS) passC(uint64_t* __restrict__ out, const uint16_t* __restrict__ perm, int64_t n, int64_t ntiles) {
  extern __shared__ uint64_t s_vals[];
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * T;
    const int len = (int)min((int64_t)T, n - base);
    uint16_t pm[T / THREADS];
#pragma unroll
    for (int k = 0; k < T / THREADS; ++k) {
      const int r = k * THREADS + threadIdx.x;
      pm[k] = r < len ? perm[base + r] : 0;
    }
    if (len == T) {
      const uint4* src = reinterpret_cast<const uint4*>(out + base);
      uint4* dst = reinterpret_cast<uint4*>(s_vals);
#pragma unroll
      for (int k = 0; k < T / 2 / THREADS; ++k) dst[k * THREADS + threadIdx.x] = __ldcs(src + k * THREADS + threadIdx.x);
    } else {
      for (int r = threadIdx.x; r < len; r += THREADS) s_vals[r] = __ldcs(out + base + r);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < T / THREADS; ++k) {
      const int r = k * THREADS + threadIdx.x;
      if (r < len) __stcs(out + base + r, s_vals[pm[k]]);
    }
    __syncthreads();
  }
}

struct Ctx {
  const uint64_t* vals; const uint32_t* idx; uint64_t* out; int64_t n; uint64_t vlen;
  uint32_t* sorted; uint16_t* perm; uint16_t* off; uint16_t* offT; int64_t ntiles; int64_t tpad; int nb; int shift; int sms;
  unsigned long long* counter; int bgrid; int ct; int T; int ablocks; int cblocks;
};

template <int T> static void run_A(void* a) {
  Ctx* c = (Ctx*)a;
  passA<T, 1024><<<c->sms * c->ablocks, 1024>>>(c->idx, c->n, c->vlen, c->shift, c->nb, c->sorted, c->perm, c->off, c->ntiles);
  dim3 g((unsigned)((c->tpad + 31) / 32), (unsigned)((c->nb + 2 + 31) / 32));
  transpose_off<<<g, dim3(32, 8)>>>(c->off, c->offT, c->ntiles, c->nb + 2, c->tpad);
}
template <int T> static void run_B(void* a) {
  Ctx* c = (Ctx*)a;
  cudaMemsetAsync(c->counter, 0, 8, 0);
  switch (c->ct) {
    case 32: passB<T, 256, 32><<<c->bgrid, 256>>>(c->vals, c->sorted, c->offT, c->tpad, c->nb, c->shift, c->out, c->counter); break;
    case 64: passB<T, 256, 64><<<c->bgrid, 256>>>(c->vals, c->sorted, c->offT, c->tpad, c->nb, c->shift, c->out, c->counter); break;
    case 128: passB<T, 256, 128><<<c->bgrid, 256>>>(c->vals, c->sorted, c->offT, c->tpad, c->nb, c->shift, c->out, c->counter); break;
    default: passB<T, 256, 256><<<c->bgrid, 256>>>(c->vals, c->sorted, c->offT, c->tpad, c->nb, c->shift, c->out, c->counter); break;
  }
}
template <int T> static void run_C(void* a) {
  Ctx* c = (Ctx*)a;
  passC<T, 1024><<<c->sms * c->cblocks, 1024, T * 8>>>(c->out, c->perm, c->n, c->ntiles);
}
template <int T> static void run_ABC(void* a) { run_A<T>(a); run_B<T>(a); run_C<T>(a); }

static float time_it(int reps, void (*fn)(void*), void* arg, uint4* flush, int64_t flush_n) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps + 1; ++r) {
    flush_l2<<<1184, 256>>>(flush, flush_n);
    CK(cudaEventRecord(e0, 0));
    fn(arg);
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0) best = std::min(best, ms);
  }
  CK(cudaGetLastError());
  return best;
}

template <int T>
static void experiment(Ctx c, int reps, uint4* flush, int64_t flush_n, unsigned long long* bad, const std::vector<int>& wmbs) {
  const int64_t n = c.n;
  c.T = T;
  c.ntiles = (n + T - 1) / T;
  c.tpad = (c.ntiles + 255) / 256 * 256;
  CK(cudaFuncSetAttribute(passC<T, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, T * 8));
  for (int wmb : wmbs) {
    int shift = 0; while ((8ull << shift) < ((uint64_t)wmb << 20)) ++shift;
    int nb = (int)((c.vlen + (1ull << shift) - 1) >> shift);
    if (nb > MAXB) continue;
    c.shift = shift; c.nb = nb;
    printf("== T=%d window %d MB -> %d buckets, avg run %.1f rows\n", T, wmb, nb, (double)T / nb);
    CK(cudaMemset(c.out, 0, n * 8));
    c.ablocks = T == 8192 ? 2 : 2; c.cblocks = 2;
    float a = time_it(reps, run_A<T>, &c, flush, flush_n);
    printf("   A %.3f ms\n", a);
    float bestB = 1e9; int bct = 0, bbg = 0;
    for (int ct : {32, 64, 128, 256}) {
      for (int bg : {2, 4, 8}) {
        c.ct = ct; c.bgrid = c.sms * bg;
        float b = time_it(reps, run_B<T>, &c, flush, flush_n);
        printf("   B ct=%3d bgrid=%d/SM  %.3f ms\n", ct, bg, b);
        if (b < bestB) { bestB = b; bct = ct; bbg = bg; }
      }
    }
    c.ct = bct; c.bgrid = c.sms * bbg;
    float cc = time_it(reps, run_C<T>, &c, flush, flush_n);
    float all = time_it(reps, run_ABC<T>, &c, flush, flush_n);
    CK(cudaMemset(bad, 0, 8));
    check_out<<<1184, 256>>>(c.out, c.idx, n, bad);
    unsigned long long h; CK(cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost));
    printf("   C %.3f ms;  best B ct=%d bgrid=%d: %.3f;  A+B+C = %.3f ms  %.2f Grows/s  frac %.3f  %s\n", cc, bct, bbg, bestB, all, n / all / 1e6,
           20.0 * n / all / 1e6 / 6586.4, h ? "MISMATCH" : "ok");
  }
}

int main(int argc, char** argv) {
  const uint64_t vlen = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
  const int64_t n = argc > 2 ? atoll(argv[2]) : 100000000ll;
  const int reps = argc > 3 ? atoi(argv[3]) : 3;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs; vlen=%llu n=%lld\n", prop.name, sms, (unsigned long long)vlen, (long long)n);
  uint64_t *vals, *out; uint32_t *idx, *sorted; uint16_t *perm, *off, *offT; unsigned long long *bad, *counter; uint4* flush;
  const int64_t flush_n = (256ll << 20) / 16;
  CK(cudaMalloc(&vals, vlen * 8)); CK(cudaMalloc(&out, n * 8)); CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&sorted, n * 4));
  CK(cudaMalloc(&perm, n * 2)); CK(cudaMalloc(&bad, 8)); CK(cudaMalloc(&counter, 8)); CK(cudaMalloc(&flush, flush_n * 16));
  const int64_t ntiles_max = (n + 4095) / 4096 + 256;
  CK(cudaMalloc(&off, (size_t)(MAXB + 2) * ntiles_max * 2)); CK(cudaMalloc(&offT, (size_t)(MAXB + 2) * ntiles_max * 2));
  fill_vals<<<1184, 256>>>(vals, (int64_t)vlen);
  fill_idx<<<1184, 256>>>(idx, n, vlen, 0x0ff1ce);
  CK(cudaDeviceSynchronize());
  Ctx c{}; c.vals = vals; c.idx = idx; c.out = out; c.n = n; c.vlen = vlen; c.sorted = sorted; c.perm = perm; c.off = off; c.offT = offT;
  c.sms = sms; c.counter = counter;
  std::vector<int> wmbs = {4, 8, 16, 32};
  if (getenv("WMB")) { wmbs.clear(); wmbs.push_back(atoi(getenv("WMB"))); }
  experiment<8192>(c, reps, flush, flush_n, bad, wmbs);
  experiment<4096>(c, reps, flush, flush_n, bad, wmbs);
  return 0;
}
