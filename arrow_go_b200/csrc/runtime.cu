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
// One Runtime per device.  A thread works on its CURRENT device: the one it chose with ag_set_device, else the
// process default (first ag_init / LOCAL_RANK / device 0).  One process per GPU (torchrun) uses exactly one entry;
// a single process that drives the whole box (a Go program: one process, many goroutines) calls ag_init_all and
// then ag_set_device per worker thread / per shard.  Streams remember their device, so an entry point that is
// handed a stream always runs on that stream's device whatever the thread's current device is.
constexpr int kMaxDevices = 16;
struct Runtime {
  bool ready = false;
  int device = -1;
  int sms = 0;
  cudaStream_t default_stream = nullptr;
  std::mutex mu;
  std::vector<cudaStream_t> free_streams;   // pooled streams for the host-pointer entry points
  std::vector<cudaStream_t> all_streams;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  int numa_state = 0;                        // 0 unknown, 1 have cpu set, -1 unavailable
  cpu_set_t numa_cpus;
};
static Runtime g_dev[kMaxDevices];
static int g_default_device = -1;            // set by the first successful init
static int g_device_count = -1;
static thread_local int tls_device = -1;     // ag_set_device; -1 = process default
static std::mutex g_init_mu;
static std::mutex g_ws_mu;                   // workspaces + stream -> device registry (streams are unique process-wide)
static std::unordered_map<cudaStream_t, Workspace*> g_workspaces;
static std::unordered_map<cudaStream_t, int> g_stream_device;
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

static int current_device_index() { return tls_device >= 0 ? tls_device : g_default_device; }
static Runtime& cur() { return g_dev[current_device_index()]; }

static void register_stream(cudaStream_t st, int device) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  g_stream_device[st] = device;
}

// g_init_mu held
static ag_status init_device_locked(int device) {
  if (g_device_count < 0) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
      cudaGetLastError();
      AG_FAIL(AG_ERR_CUDA, "no CUDA device available (%s); libarrowgpu has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    g_device_count = count < kMaxDevices ? count : kMaxDevices;
  }
  if (device < 0) {
    const char* lr = getenv("LOCAL_RANK");   // torchrun: one process per GPU
    device = lr ? atoi(lr) : 0;
  }
  if (device >= g_device_count) AG_FAIL(AG_ERR_INVALID, "device %d out of range (have %d)", device, g_device_count);
  Runtime& rt = g_dev[device];
  AG_CUDA_TRY(cudaSetDevice(device));
  if (rt.ready) return AG_OK;
  cudaDeviceProp prop;
  AG_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    AG_FAIL(AG_ERR_CUDA, "device %d is sm_%d%d; libarrowgpu is built for sm_100a only", device, prop.major, prop.minor);
  rt.device = device;
  rt.sms = prop.multiProcessorCount;
  AG_CUDA_TRY(cudaStreamCreateWithFlags(&rt.default_stream, cudaStreamNonBlocking));
  register_stream(rt.default_stream, device);
  // keep freed blocks cached in the stream-ordered pool (temp buffers of the host entry points, take's scratch)
  cudaMemPool_t pool;
  AG_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thresh = UINT64_MAX;
  AG_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  rt.ready = true;
  if (g_default_device < 0) g_default_device = device;
  return AG_OK;
}

ag_status ensure_init() {
  const int want = current_device_index();
  if (want >= 0 && g_dev[want].ready) {
    // cgo calls arrive on arbitrary OS threads: bind the device for this thread.
    int c = -1;
    if (cudaGetDevice(&c) != cudaSuccess || c != want) AG_CUDA_TRY(cudaSetDevice(want));
    return AG_OK;
  }
  std::lock_guard<std::mutex> lk(g_init_mu);
  return init_device_locked(want);
}

int sm_count() {
  const int d = current_device_index();
  return (d >= 0 && g_dev[d].sms > 0) ? g_dev[d].sms : 148;
}
int current_device() { return current_device_index(); }

int device_count() { return g_device_count; }

// NULL -> the current device's default stream.  A stream created by this library runs on ITS device: the calling
// thread is switched to it (kernel launches and stream-ordered allocations follow the current device).
cudaStream_t resolve_stream(ag_stream_t s) {
  if (!s) return cur().default_stream;
  cudaStream_t st = (cudaStream_t)s;
  int dev = -1;
  {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    auto it = g_stream_device.find(st);
    if (it != g_stream_device.end()) dev = it->second;
  }
  if (dev >= 0 && dev != current_device_index()) {
    tls_device = dev;
    cudaSetDevice(dev);
  }
  return st;
}

ag_status get_workspace(cudaStream_t s, Workspace** out) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  auto it = g_workspaces.find(s);
  if (it != g_workspaces.end()) { *out = it->second; return AG_OK; }
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
  ws->seq_mu = new std::mutex();
  g_workspaces[s] = ws;
  *out = ws;
  return AG_OK;
}

// Caller holds WorkspaceLock(ws): growth cannot race with another call's launches on the same stream.
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
  Runtime& rt = cur();
  {
    std::lock_guard<std::mutex> lk(rt.mu);
    if (!rt.free_streams.empty()) {
      *out = rt.free_streams.back();
      rt.free_streams.pop_back();
      return AG_OK;
    }
  }
  cudaStream_t st;
  AG_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  register_stream(st, rt.device);
  {
    std::lock_guard<std::mutex> lk(rt.mu);
    rt.all_streams.push_back(st);
  }
  *out = st;
  return AG_OK;
}
void release_call_stream(cudaStream_t st) {
  int dev = -1;
  {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    auto it = g_stream_device.find(st);
    if (it != g_stream_device.end()) dev = it->second;
  }
  Runtime& rt = dev >= 0 ? g_dev[dev] : cur();
  std::lock_guard<std::mutex> lk(rt.mu);
  rt.free_streams.push_back(st);
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
  AG_TRY(init_device_locked(device));
  int d = -1;
  cudaGetDevice(&d);
  // the first ag_init names the process default; a later ag_init(other) initialises that device too and makes it
  // the calling thread's current device (same as ag_set_device)
  if (d >= 0 && d != g_default_device) tls_device = d;
  return AG_OK;
}

ag_status ag_init_all(int* n_devices) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  AG_TRY(init_device_locked(g_default_device >= 0 ? g_default_device : 0));
  for (int d = 0; d < g_device_count; ++d) AG_TRY(init_device_locked(d));
  // peer access between every pair: the sharded reductions read the other devices' partial results directly
  for (int a = 0; a < g_device_count; ++a) {
    AG_CUDA_TRY(cudaSetDevice(a));
    for (int b = 0; b < g_device_count; ++b) {
      if (a == b) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
        cudaGetLastError();
      }
    }
  }
  AG_CUDA_TRY(cudaSetDevice(current_device_index()));
  if (n_devices) *n_devices = g_device_count;
  return AG_OK;
}

ag_status ag_set_device(int device) {
  {
    std::lock_guard<std::mutex> lk(g_init_mu);
    AG_TRY(init_device_locked(device));
  }
  int d = -1;
  AG_CUDA_TRY(cudaGetDevice(&d));
  tls_device = d;
  return AG_OK;
}

ag_status ag_get_device(int* device) {
  if (!device) AG_FAIL(AG_ERR_INVALID, "ag_get_device: NULL argument");
  AG_TRY(ensure_init());
  *device = current_device_index();
  return AG_OK;
}

ag_status ag_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (g_default_device < 0) return AG_OK;
  for (int d = 0; d < kMaxDevices; ++d) {
    if (!g_dev[d].ready) continue;
    cudaSetDevice(d);
    cudaDeviceSynchronize();
  }
  {
    std::lock_guard<std::mutex> lk2(g_ws_mu);
    for (auto& kv : g_workspaces) {
      Workspace* ws = kv.second;
      auto it = g_stream_device.find(kv.first);
      if (it != g_stream_device.end()) cudaSetDevice(it->second);
      cudaFree(ws->partials); cudaFree(ws->ticket); cudaFree(ws->scalars);
      if (ws->tile_status) cudaFree(ws->tile_status);
      cudaFreeHost(ws->h_scalars);
      delete ws->seq_mu;
      delete ws;
    }
    g_workspaces.clear();
    g_stream_device.clear();
  }
  for (int d = 0; d < kMaxDevices; ++d) {
    Runtime& rt = g_dev[d];
    if (!rt.ready) continue;
    cudaSetDevice(d);
    {
      std::lock_guard<std::mutex> lk2(rt.mu);
      for (cudaStream_t st : rt.all_streams) cudaStreamDestroy(st);
      rt.all_streams.clear();
      rt.free_streams.clear();
      if (rt.flush_buf) { cudaFree(rt.flush_buf); rt.flush_buf = nullptr; }
    }
    cudaStreamDestroy(rt.default_stream);
    rt.default_stream = nullptr;
    rt.ready = false;
  }
  g_default_device = -1;
  tls_device = -1;
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
  AG_CUDA_TRY(cudaGetDeviceProperties(&prop, cur().device));
  if (device) *device = cur().device;
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
  Runtime& rt = cur();
  std::lock_guard<std::mutex> lk(rt.mu);
  if (rt.numa_state == 0) {
    rt.numa_state = -1;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), rt.device) == cudaSuccess) {
      for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
      char path[128];
      snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
      int node = -1;
      if (FILE* f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
      if (node >= 0) {
        snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
        if (FILE* f = fopen(path, "r")) {
          CPU_ZERO(&rt.numa_cpus);
          int a, b; char sep;
          bool any = false;
          while (fscanf(f, "%d", &a) == 1) {
            b = a;
            if (fscanf(f, "%c", &sep) == 1 && sep == '-') { if (fscanf(f, "%d", &b) != 1) b = a; if (fscanf(f, "%c", &sep) != 1) sep = 0; }
            for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &rt.numa_cpus); any = true; }
            if (sep != ',') break;
          }
          fclose(f);
          if (any) rt.numa_state = 1;
        }
      }
    } else {
      cudaGetLastError();
    }
  }
  if (rt.numa_state == 1) { *set = rt.numa_cpus; return true; }
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
  AG_CUDA_TRY(cudaMemsetAsync(*dptr, 0, sz, cur().default_stream));
  AG_CUDA_TRY(cudaStreamSynchronize(cur().default_stream));
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
  register_stream(st, cur().device);
  *s = (ag_stream_t)st;
  return AG_OK;
}
ag_status ag_stream_destroy(ag_stream_t s) {
  if (!s) return AG_OK;
  AG_TRY(ensure_init());
  cudaStream_t st = resolve_stream(s);
  AG_CUDA_TRY(cudaStreamSynchronize(st));
  {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    auto it = g_workspaces.find(st);
    if (it != g_workspaces.end()) {
      Workspace* ws = it->second;
      cudaFree(ws->partials); cudaFree(ws->ticket); cudaFree(ws->scalars);
      if (ws->tile_status) cudaFree(ws->tile_status);
      cudaFreeHost(ws->h_scalars);
      delete ws->seq_mu;
      delete ws;
      g_workspaces.erase(it);
    }
    g_stream_device.erase(st);
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
  cudaStream_t st = resolve_stream(s);
  Runtime& rt = cur();
  {
    std::lock_guard<std::mutex> lk(rt.mu);
    if (!rt.flush_buf) {
      rt.flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
      AG_CUDA_TRY(cudaMalloc(&rt.flush_buf, rt.flush_bytes));
    }
  }
  AG_CUDA_TRY(cudaMemsetAsync(rt.flush_buf, 0x5a, rt.flush_bytes, st));
  return AG_OK;
}

}  // extern "C"
