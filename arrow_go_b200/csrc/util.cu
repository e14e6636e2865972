// util.cu — parity helpers for inputs too large to bring back to the host:
// an order-sensitive 64-bit checksum and a counter-based generator (the test suite keeps a CPU
// twin of both definitions).  Not part of the reference's surface.
#include "common.cuh"

namespace ag {

constexpr int kUtThreads = 256;

__global__ void __launch_bounds__(kUtThreads)
checksum64_kernel(const unsigned long long* __restrict__ buf, size_t n_words, unsigned long long* __restrict__ res) {
  const size_t stride = (size_t)gridDim.x * kUtThreads;
  unsigned long long acc = 0;
  for (size_t i = (size_t)blockIdx.x * kUtThreads + threadIdx.x; i < n_words; i += stride) acc += mix64((uint64_t)i) * __ldcs(buf + i);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  __shared__ unsigned long long sm[kUtThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < kUtThreads / 32; ++i) t += sm[i];
    atomicAdd(res, t);  // wrapping integer adds: order-free
  }
}

// one thread per 8 elements so bitmap bytes (kind 4) are written whole
__global__ void __launch_bounds__(kUtThreads)
generate_kernel(int kind, unsigned long long seed, long long lo, long long hi, void* out, size_t n) {
  const unsigned long long span = (unsigned long long)(hi - lo) + 1ull;
  const size_t stride = (size_t)gridDim.x * kUtThreads;
  for (size_t g = (size_t)blockIdx.x * kUtThreads + threadIdx.x; g * 8 < n; g += stride) {
    unsigned byte = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const size_t i = g * 8 + e;
      if (i >= n) break;
      const unsigned long long z = mix64(seed + (unsigned long long)i);
      const unsigned long long k = ((z >> 32) * span) >> 32;
      switch (kind) {
        case 0: ((unsigned long long*)out)[i] = z; break;
        case 1: ((long long*)out)[i] = lo + (long long)k; break;
        case 2: ((int*)out)[i] = (int)(lo + (long long)k); break;
        case 3: ((double*)out)[i] = (double)(lo + (long long)k); break;
        case 4: byte |= (unsigned)((((z >> 32) * (unsigned long long)hi) >> 32) < (unsigned long long)lo) << e; break;
        case 5: ((int*)out)[i] = (int)(lo + (long long)(((unsigned __int128)i * span) / n)); break;              // sorted ramp over [lo, hi]
        case 6: ((int*)out)[i] = (int)(lo + (long long)(((unsigned __int128)(n - 1 - i) * span) / n)); break;  // reverse-sorted
        default: break;
      }
    }
    if (kind == 4) ((uint8_t*)out)[g] = (uint8_t)byte;
  }
}

}  // namespace ag

using namespace ag;

extern "C" {

ag_status ag_checksum64_dev(const void* d_buf, size_t n_words, uint64_t* d_res, ag_stream_t s) {
  AG_TRY(ensure_init());
  cudaStream_t st = resolve_stream(s);
  if (!d_res) AG_FAIL(AG_ERR_INVALID, "checksum64: NULL result");
  AG_CUDA_TRY(cudaMemsetAsync(d_res, 0, sizeof(uint64_t), st));
  if (n_words == 0) return AG_OK;
  const int grid = grid_for((int64_t)n_words, kUtThreads * 8, 8);
  checksum64_kernel<<<grid, kUtThreads, 0, st>>>((const unsigned long long*)d_buf, n_words, (unsigned long long*)d_res);
  return check_launch("checksum64_kernel");
}

ag_status ag_generate_dev(int kind, uint64_t seed, int64_t lo, int64_t hi, void* d_out, size_t n, ag_stream_t s) {
  AG_TRY(ensure_init());
  if (kind < 0 || kind > 6) AG_FAIL(AG_ERR_INVALID, "generate: bad kind %d", kind);
  if (kind != 0 && (hi < lo || (kind != 4 && (uint64_t)(hi - lo) >= (1ull << 32)))) AG_FAIL(AG_ERR_INVALID, "generate: need lo <= hi and hi-lo < 2^32");
  if (n == 0) return AG_OK;
  const int grid = grid_for((int64_t)((n + 7) / 8), kUtThreads * 2, 8);
  generate_kernel<<<grid, kUtThreads, 0, resolve_stream(s)>>>(kind, seed, lo, hi, d_out, n);
  return check_launch("generate_kernel");
}

}  // extern "C"
