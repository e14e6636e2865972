// runtime.cu — device binding, streams, per-stream workspaces, pinned/device memory,
// events.  The residency half of the C ABI (include/arrowgpu.h, "Runtime" block).
//
// Reference hooks this backs: a pinned memory.Allocator (arrow/memory/allocator.go:23-27,
// shape of arrow/memory/internal/cgoalloc/allocator.h:13-18) and "DMA once per record
// batch" residency (SURVEY.md §7 hard-part 2).
#include "common.cuh"

#include <ctype.h>
#include <sched.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace ag {

// ---------------------------------------------------------------- errors ----------
static thread_local char tls_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_err, sizeof(tls_err), fmt, ap);
  va_end(ap);
}

ag_status cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  cudaGetLastError();  // clear sticky-less errors
  return (e == cudaErrorMemoryAllocation) ? AG_ERR_OOM : AG_ERR_CUDA;
}

// ---------------------------------------------------------------- state -----------
struct Runtime {
  bool ready = false;
  int device = -1;
  int sms = 0;
  cudaStream_t default_stream = nullptr;
  std::mutex mu;
  std::unordered_map<cudaStream_t, Workspace*> workspaces;
  std::vector<cudaStream_t> free_streams;   // pooled streams for the host-pointer entry points
  std::vector<cudaStream_t> all_streams;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
};
static Runtime g_rt;
static std::mutex g_init_mu;
static std::atomic<uint64_t> g_launches{0};

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

ag_status check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("kernel launch failed (%s): %s", what, cudaGetErrorString(e));
    return AG_ERR_CUDA;
  }
  count_launch();
  return AG_OK;
}

static ag_status init_locked(int device) {
  if (g_rt.ready) {
    // one process per GPU: re-binding the calling thread is all that is needed
    AG_CUDA_TRY(cudaSetDevice(g_rt.device));
    return AG_OK;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    AG_FAIL(AG_ERR_CUDA, "no CUDA device available (%s); libarrowgpu has no CPU fallback",
            e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0) {
    const char* lr = getenv("LOCAL_RANK");
    device = lr ? atoi(lr) : 0;
  }
  if (device >= count) AG_FAIL(AG_ERR_INVALID, "device %d out of range (have %d)", device, count);
  AG_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  AG_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    AG_FAIL(AG_ERR_CUDA, "device %d is sm_%d%d; libarrowgpu is built for sm_100a only", device, prop.major, prop.minor);
  g_rt.device = device;
  g_rt.sms = prop.multiProcessorCount;
  AG_CUDA_TRY(cudaStreamCreateWithFlags(&g_rt.default_stream, cudaStreamNonBlocking));
  // keep freed blocks cached in the stream-ordered pool (temp buffers of the host entry points)
  cudaMemPool_t pool;
  AG_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thresh = UINT64_MAX;
  AG_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  g_rt.ready = true;
  return AG_OK;
}

ag_status ensure_init() {
  if (g_rt.ready) {
    // cgo calls arrive on arbitrary OS threads: bind the device for this thread.
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != g_rt.device) AG_CUDA_TRY(cudaSetDevice(g_rt.device));
    return AG_OK;
  }
  std::lock_guard<std::mutex> lk(g_init_mu);
  return init_locked(-1);
}

int sm_count() { return g_rt.sms > 0 ? g_rt.sms : 148; }

cudaStream_t resolve_stream(ag_stream_t s) { return s ? (cudaStream_t)s : g_rt.default_stream; }

ag_status get_workspace(cudaStream_t s, Workspace** out) {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  auto it = g_rt.workspaces.find(s);
  if (it != g_rt.workspaces.end()) { *out = it->second; return AG_OK; }
  Workspace* ws = new Workspace();
  memset(ws, 0, sizeof(*ws));
  AG_CUDA_TRY(cudaMalloc(&ws->partials, (size_t)kMaxPartials * 16));
  // Zero the tickets ON THE OWNING STREAM: `s` is a non-blocking stream, so a cudaMemset on the
  // legacy default stream would not be ordered before the first kernel that reads the ticket.
  AG_CUDA_TRY(cudaMalloc((void**)&ws->ticket, 64 * sizeof(unsigned)));
  AG_CUDA_TRY(cudaMemsetAsync(ws->ticket, 0, 64 * sizeof(unsigned), s));
  AG_CUDA_TRY(cudaMalloc((void**)&ws->scalars, 16 * sizeof(int64_t)));
  AG_CUDA_TRY(cudaMemsetAsync(ws->scalars, 0, 16 * sizeof(int64_t), s));
  AG_CUDA_TRY(cudaHostAlloc((void**)&ws->h_scalars, 16 * sizeof(int64_t), cudaHostAllocDefault));
  ws->tile_status = nullptr;
  ws->tile_status_cap = 0;
  g_rt.workspaces[s] = ws;
  *out = ws;
  return AG_OK;
}

ag_status ensure_tile_status(Workspace* ws, size_t n_tiles, cudaStream_t s) {
  if (ws->tile_status_cap >= n_tiles) return AG_OK;
  size_t cap = ws->tile_status_cap ? ws->tile_status_cap : 4096;
  while (cap < n_tiles) cap *= 2;
  if (ws->tile_status) {
    // earlier kernels on this stream may still read the old buffer
    AG_CUDA_TRY(cudaStreamSynchronize(s));
    AG_CUDA_TRY(cudaFree(ws->tile_status));
    ws->tile_status = nullptr;
    ws->tile_status_cap = 0;
  }
  AG_CUDA_TRY(cudaMalloc((void**)&ws->tile_status, cap * sizeof(unsigned long long)));
  ws->tile_status_cap = cap;
  return AG_OK;
}

ag_status acquire_call_stream(cudaStream_t* out) {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  if (!g_rt.free_streams.empty()) {
    *out = g_rt.free_streams.back();
    g_rt.free_streams.pop_back();
    return AG_OK;
  }
  cudaStream_t st;
  AG_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  g_rt.all_streams.push_back(st);
  *out = st;
  return AG_OK;
}
void release_call_stream(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_rt.mu);
  g_rt.free_streams.push_back(st);
}

// Resident blocks per SM for a kernel (cudaOccupancyMaxActiveBlocksPerMultiprocessor), cached per
// entry point.  Persistent / grid-stride kernels size their grid as SMs x this, so the whole
// grid is one wave and no tail wave runs at partial occupancy.
int blocks_per_sm(const void* kernel, int threads) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(kernel);
  if (it != cache.end()) return it->second;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0) != cudaSuccess || n < 1) { cudaGetLastError(); n = 1; }
  cache[kernel] = n;
  return n;
}

ag_status dev_alloc_async(void** p, size_t nbytes, cudaStream_t s) {
  if (nbytes == 0) nbytes = 16;
  AG_CUDA_TRY(cudaMallocAsync(p, nbytes, s));
  return AG_OK;
}
ag_status dev_free_async(void* p, cudaStream_t s) {
  if (p) AG_CUDA_TRY(cudaFreeAsync(p, s));
  return AG_OK;
}

}  // namespace ag

using namespace ag;

// ================================================================ C ABI =============
extern "C" {

ag_status ag_init(int device) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (g_rt.ready && device >= 0 && device != g_rt.device)
    AG_FAIL(AG_ERR_INVALID, "ag_init(%d): process already bound to device %d (one process per GPU)", device, g_rt.device);
  return init_locked(device);
}

ag_status ag_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (!g_rt.ready) return AG_OK;
  cudaDeviceSynchronize();
  {
    std::lock_guard<std::mutex> lk2(g_rt.mu);
    for (auto& kv : g_rt.workspaces) {
      Workspace* ws = kv.second;
      cudaFree(ws->partials); cudaFree(ws->ticket); cudaFree(ws->scalars);
      if (ws->tile_status) cudaFree(ws->tile_status);
      cudaFreeHost(ws->h_scalars);
      delete ws;
    }
    g_rt.workspaces.clear();
    for (cudaStream_t st : g_rt.all_streams) cudaStreamDestroy(st);
    g_rt.all_streams.clear();
    g_rt.free_streams.clear();
    if (g_rt.flush_buf) { cudaFree(g_rt.flush_buf); g_rt.flush_buf = nullptr; }
  }
  cudaStreamDestroy(g_rt.default_stream);
  g_rt.default_stream = nullptr;
  g_rt.ready = false;
  return AG_OK;
}

ag_status ag_device_count(int* count) {
  if (!count) AG_FAIL(AG_ERR_INVALID, "ag_device_count: NULL argument");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
  return AG_OK;
}

ag_status ag_device_info(int* device, int* sms, size_t* hbm_bytes, int* cc_major, int* cc_minor) {
  AG_TRY(ensure_init());
  cudaDeviceProp prop;
  AG_CUDA_TRY(cudaGetDeviceProperties(&prop, g_rt.device));
  if (device) *device = g_rt.device;
  if (sms) *sms = prop.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = prop.totalGlobalMem;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return AG_OK;
}

void ag_last_error(char* buf, size_t buflen) {
  if (!buf || buflen == 0) return;
  strncpy(buf, tls_err, buflen - 1);
  buf[buflen - 1] = 0;
}

const char* ag_version(void) { return "arrowgpu 0.1 (sm_100a)"; }
uint64_t ag_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- pinned host memory ---------------------------------------------------------
// CPUs of the NUMA node the GPU hangs off (sysfs); empty when unknown.  Pinned buffers are
// allocated and first-touched from one of those CPUs so DMA does not cross the socket link.
static bool gpu_node_cpus(cpu_set_t* set) {
  static int state = 0;  // 0 unknown, 1 have set, -1 unavailable
  static cpu_set_t cached;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (state == 0) {
    state = -1;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), g_rt.device) == cudaSuccess) {
      for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
      char path[128];
      snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
      int node = -1;
      if (FILE* f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
      if (node >= 0) {
        snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
        if (FILE* f = fopen(path, "r")) {
          CPU_ZERO(&cached);
          int a, b; char sep;
          bool any = false;
          while (fscanf(f, "%d", &a) == 1) {
            b = a;
            if (fscanf(f, "%c", &sep) == 1 && sep == '-') { if (fscanf(f, "%d", &b) != 1) b = a; if (fscanf(f, "%c", &sep) != 1) sep = 0; }
            for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &cached); any = true; }
            if (sep != ',') break;
          }
          fclose(f);
          if (any) state = 1;
        }
      }
    } else {
      cudaGetLastError();
    }
  }
  if (state == 1) { *set = cached; return true; }
  return false;
}

ag_status ag_host_alloc(void** ptr, size_t nbytes) {
  if (!ptr) AG_FAIL(AG_ERR_INVALID, "ag_host_alloc: NULL argument");
  AG_TRY(ensure_init());
  size_t sz = nbytes ? ((nbytes + 63) & ~(size_t)63) : 64;
  cpu_set_t node_cpus, saved;
  const bool pin = gpu_node_cpus(&node_cpus) && sched_getaffinity(0, sizeof(saved), &saved) == 0 &&
                   sched_setaffinity(0, sizeof(node_cpus), &node_cpus) == 0;
  cudaError_t e = cudaHostAlloc(ptr, sz, cudaHostAllocDefault);  // page aligned >= 64 B
  if (e == cudaSuccess) memset(*ptr, 0, sz);                     // zero-initialised like GoAllocator / calloc
  if (pin) sched_setaffinity(0, sizeof(saved), &saved);
  if (e != cudaSuccess) { *ptr = nullptr; return cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__); }
  return AG_OK;
}

ag_status ag_host_realloc(void** ptr, size_t old_nbytes, size_t new_nbytes) {
  if (!ptr) AG_FAIL(AG_ERR_INVALID, "ag_host_realloc: NULL argument");
  void* np = nullptr;
  AG_TRY(ag_host_alloc(&np, new_nbytes));
  if (*ptr) {
    memcpy(np, *ptr, old_nbytes < new_nbytes ? old_nbytes : new_nbytes);
    cudaFreeHost(*ptr);
  }
  *ptr = np;
  return AG_OK;
}

ag_status ag_host_free(void* ptr) {
  if (!ptr) return AG_OK;
  AG_TRY(ensure_init());
  AG_CUDA_TRY(cudaFreeHost(ptr));
  return AG_OK;
}

ag_status ag_host_register(void* ptr, size_t nbytes) {
  AG_TRY(ensure_init());
  AG_CUDA_TRY(cudaHostRegister(ptr, nbytes, cudaHostRegisterDefault));
  return AG_OK;
}
ag_status ag_host_unregister(void* ptr) {
  AG_TRY(ensure_init());
  AG_CUDA_TRY(cudaHostUnregister(ptr));
  return AG_OK;
}

// ---- device memory ---------------------------------------------------------------
ag_status ag_dev_alloc(void** dptr, size_t nbytes) {
  if (!dptr) AG_FAIL(AG_ERR_INVALID, "ag_dev_alloc: NULL argument");
  AG_TRY(ensure_init());
  size_t sz = nbytes ? ((nbytes + 63) & ~(size_t)63) : 64;  // Arrow padding: 64-byte multiples
  cudaError_t e = cudaMalloc(dptr, sz);
  if (e != cudaSuccess) { *dptr = nullptr; return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
  AG_CUDA_TRY(cudaMemsetAsync(*dptr, 0, sz, g_rt.default_stream));
  AG_CUDA_TRY(cudaStreamSynchronize(g_rt.default_stream));
  return AG_OK;
}
ag_status ag_dev_free(void* dptr) {
  if (!dptr) return AG_OK;
  AG_TRY(ensure_init());
  AG_CUDA_TRY(cudaFree(dptr));
  return AG_OK;
}
ag_status ag_dev_memset(void* dptr, int byte, size_t nbytes, ag_stream_t s) {
  AG_TRY(ensure_init());
  AG_CUDA_TRY(cudaMemsetAsync(dptr, byte, nbytes, resolve_stream(s)));
  return AG_OK;
}
ag_status ag_upload(void* dst, const void* src, size_t nbytes, ag_stream_t s) {
  AG_TRY(ensure_init());
  if (nbytes) AG_CUDA_TRY(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, resolve_stream(s)));
  return AG_OK;
}
ag_status ag_download(void* dst, const void* src, size_t nbytes, ag_stream_t s) {
  AG_TRY(ensure_init());
  if (nbytes) AG_CUDA_TRY(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, resolve_stream(s)));
  return AG_OK;
}
ag_status ag_copy_dev(void* dst, const void* src, size_t nbytes, ag_stream_t s) {
  AG_TRY(ensure_init());
  if (nbytes) AG_CUDA_TRY(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, resolve_stream(s)));
  return AG_OK;
}

// ---- streams & events --------------------------------------------------------------
ag_status ag_stream_create(ag_stream_t* s) {
  if (!s) AG_FAIL(AG_ERR_INVALID, "ag_stream_create: NULL argument");
  AG_TRY(ensure_init());
  cudaStream_t st;
  AG_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = (ag_stream_t)st;
  return AG_OK;
}
ag_status ag_stream_destroy(ag_stream_t s) {
  if (!s) return AG_OK;
  AG_TRY(ensure_init());
  cudaStream_t st = (cudaStream_t)s;
  AG_CUDA_TRY(cudaStreamSynchronize(st));
  {
    std::lock_guard<std::mutex> lk(g_rt.mu);
    auto it = g_rt.workspaces.find(st);
    if (it != g_rt.workspaces.end()) {
      Workspace* ws = it->second;
      cudaFree(ws->partials); cudaFree(ws->ticket); cudaFree(ws->scalars);
      if (ws->tile_status) cudaFree(ws->tile_status);
      cudaFreeHost(ws->h_scalars);
      delete ws;
      g_rt.workspaces.erase(it);
    }
  }
  AG_CUDA_TRY(cudaStreamDestroy(st));
  return AG_OK;
}
ag_status ag_stream_sync(ag_stream_t s) {
  AG_TRY(ensure_init());
  AG_CUDA_TRY(cudaStreamSynchronize(resolve_stream(s)));
  return AG_OK;
}
ag_status ag_event_create(ag_event_t* e) {
  if (!e) AG_FAIL(AG_ERR_INVALID, "ag_event_create: NULL argument");
  AG_TRY(ensure_init());
  cudaEvent_t ev;
  AG_CUDA_TRY(cudaEventCreate(&ev));
  *e = (ag_event_t)ev;
  return AG_OK;
}
ag_status ag_event_destroy(ag_event_t e) {
  if (e) AG_CUDA_TRY(cudaEventDestroy((cudaEvent_t)e));
  return AG_OK;
}
ag_status ag_event_record(ag_event_t e, ag_stream_t s) {
  AG_TRY(ensure_init());
  AG_CUDA_TRY(cudaEventRecord((cudaEvent_t)e, resolve_stream(s)));
  return AG_OK;
}
ag_status ag_event_sync(ag_event_t e) {
  AG_CUDA_TRY(cudaEventSynchronize((cudaEvent_t)e));
  return AG_OK;
}
ag_status ag_event_elapsed_ms(ag_event_t a, ag_event_t b, float* ms) {
  AG_CUDA_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return AG_OK;
}

ag_status ag_flush_l2(ag_stream_t s) {
  AG_TRY(ensure_init());
  {
    std::lock_guard<std::mutex> lk(g_rt.mu);
    if (!g_rt.flush_buf) {
      g_rt.flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
      AG_CUDA_TRY(cudaMalloc(&g_rt.flush_buf, g_rt.flush_bytes));
    }
  }
  AG_CUDA_TRY(cudaMemsetAsync(g_rt.flush_buf, 0x5a, g_rt.flush_bytes, resolve_stream(s)));
  return AG_OK;
}

}  // extern "C"
