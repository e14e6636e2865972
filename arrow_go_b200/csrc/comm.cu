// comm.cu — the multi-GPU plumbing of the C ABI: row-range sharding and the communicator behind the
// global Sum.  The reference has no distributed layer; the fan-out a binding replaces is the per-chunk errgroup of
// arrow/compute/selection.go:127-150 (one process, NumParallel goroutines).  Here:
//
//   * one process per GPU (torchrun): every rank creates a mailbox (ag_comm_local_handle), the caller's own plumbing
//     all-gathers the 64-byte IPC handles, ag_comm_create maps the peers' mailboxes (cudaIpcOpenMemHandle);
//   * one process, all GPUs (a Go program): ag_init_all + ag_comm_create_local — peer access, no IPC;
//   * ag_sum_*_global_dev (reduce.cu) then runs the local reduction and the fold over the ranks as ONE kernel that
//     stores into / polls HBM mailboxes over NVLink (reduce.cu: exchange_sum);
//   * NCCL is optional plumbing: ag_comm_attach_nccl dlopens libnccl.so.2 (whatever copy the process already has) for
//     ag_sum_i64_global_nccl_dev — per-GPU Sum + ncclAllReduce on the same stream, the north_star's literal form.
#include "common.cuh"

#include <dlfcn.h>
#include <string.h>
#include <mutex>
#include <vector>

namespace ag {

struct Comm {
  int world = 1, rank = 0, device = 0;
  MailSlot* local = nullptr;            // 2 x world slots in this rank's HBM
  bool owns_local = true;
  std::vector<void*> opened;            // peers' mailboxes mapped by cudaIpcOpenMemHandle
  MailSlot** d_peers = nullptr;         // device array [world]
  std::atomic<unsigned long long> epoch{0};
  void* nccl = nullptr;                 // ncclComm_t when attached
};

// a mailbox created by ag_comm_local_handle and not yet owned by a communicator (per device)
static std::mutex g_pending_mu;
static MailSlot* g_pending[16] = {nullptr};
static int g_pending_world[16] = {0};

static ag_status alloc_mailbox(int world, MailSlot** out) {
  const size_t bytes = sizeof(MailSlot) * 2 * (size_t)world;
  AG_CUDA_TRY(cudaMalloc((void**)out, bytes));   // cudaMalloc (not the pool): IPC-exportable
  AG_CUDA_TRY(cudaMemset(*out, 0, bytes));
  AG_CUDA_TRY(cudaDeviceSynchronize());
  return AG_OK;
}

ag_status comm_next_exchange(ag_comm_t c, SumExchange* x) {
  Comm* comm = reinterpret_cast<Comm*>(c);
  if (!comm) AG_FAIL(AG_ERR_INVALID, "global sum: NULL communicator");
  if (comm->device != current_device()) AG_FAIL(AG_ERR_INVALID, "global sum: communicator belongs to device %d, the call runs on device %d", comm->device, current_device());
  x->peers = comm->d_peers;
  x->local = comm->local;
  x->world = comm->world;
  x->rank = comm->rank;
  x->epoch = comm->epoch.fetch_add(1) + 1;
  return AG_OK;
}

// ---- NCCL through dlopen (no link-time dependency; the process may already hold a copy, e.g. torch's) -------------
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, const void*, int) = nullptr;   // ncclUniqueId is passed BY VALUE in the real ABI: see below
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId { char bytes[128]; };   // ncclUniqueId: 128 opaque bytes, passed by value
typedef int (*nccl_init_rank_fn)(void**, int, NcclId, int);
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static ag_status load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return AG_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "NCCL is not available in this process (dlopen libnccl.so.2: %s)", dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, const void*, int))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    AG_FAIL(AG_ERR_NOT_IMPLEMENTED, "libnccl.so.2 lacks an expected symbol");
  g_nccl.lib = h;
  return AG_OK;
}
#define AG_NCCL_TRY(expr)                                                                                   \
  do {                                                                                                      \
    int _r = (expr);                                                                                        \
    if (_r != 0) AG_FAIL(AG_ERR_CUDA, "NCCL error %d (%s): %s", _r, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?", #expr); \
  } while (0)

ag_status comm_nccl_allreduce_sum_i64(ag_comm_t c, void* d_buf, size_t count, cudaStream_t st) {
  Comm* comm = reinterpret_cast<Comm*>(c);
  if (!comm || !comm->nccl) AG_FAIL(AG_ERR_INVALID, "communicator has no NCCL attached (ag_comm_attach_nccl)");
  AG_NCCL_TRY(g_nccl.AllReduce(d_buf, d_buf, count, /*ncclInt64*/ 4, /*ncclSum*/ 0, comm->nccl, st));
  return AG_OK;
}

}  // namespace ag

using namespace ag;

extern "C" {

// [start, stop) of shard `shard` of n_rows: ceil-balanced, cut points at multiples of 64 rows so no two shards share
// a 64-bit bitmap word (the last shard takes the tail).  Same rule as arrow_go_b200/sharding.py.
ag_status ag_shard_range(int64_t n_rows, int shard, int n_shards, int64_t* start, int64_t* stop) {
  if (n_rows < 0 || n_shards < 1 || shard < 0 || shard >= n_shards || !start || !stop) AG_FAIL(AG_ERR_INVALID, "ag_shard_range: bad arguments");
  int64_t per = (n_rows + n_shards - 1) / n_shards;
  per = (per + 63) / 64 * 64;
  const int64_t a = (int64_t)shard * per < n_rows ? (int64_t)shard * per : n_rows;
  *start = a;
  *stop = a + per < n_rows ? a + per : n_rows;
  return AG_OK;
}

ag_status ag_comm_local_handle(int world, void* handle64) {
  if (world < 1 || world > 64 || !handle64) AG_FAIL(AG_ERR_INVALID, "ag_comm_local_handle: bad arguments");
  AG_TRY(ensure_init());
  const int dev = current_device();
  std::lock_guard<std::mutex> lk(g_pending_mu);
  if (g_pending[dev]) { cudaFree(g_pending[dev]); g_pending[dev] = nullptr; }
  AG_TRY(alloc_mailbox(world, &g_pending[dev]));
  g_pending_world[dev] = world;
  cudaIpcMemHandle_t h;
  AG_CUDA_TRY(cudaIpcGetMemHandle(&h, g_pending[dev]));
  static_assert(sizeof(h) == AG_COMM_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, sizeof(h));
  return AG_OK;
}

ag_status ag_comm_create(ag_comm_t* out, int world, int rank, const void* all_handles) {
  if (!out || world < 1 || rank < 0 || rank >= world) AG_FAIL(AG_ERR_INVALID, "ag_comm_create: bad arguments");
  AG_TRY(ensure_init());
  const int dev = current_device();
  Comm* c = new Comm();
  c->world = world; c->rank = rank; c->device = dev;
  {
    std::lock_guard<std::mutex> lk(g_pending_mu);
    if (g_pending[dev] && g_pending_world[dev] == world) { c->local = g_pending[dev]; g_pending[dev] = nullptr; }
  }
  if (!c->local) {
    if (world > 1) { delete c; AG_FAIL(AG_ERR_INVALID, "ag_comm_create: call ag_comm_local_handle(world, ...) first and all-gather the handles"); }
    ag_status rc = alloc_mailbox(world, &c->local);
    if (rc != AG_OK) { delete c; return rc; }
  }
  std::vector<MailSlot*> peers((size_t)world, nullptr);
  for (int r = 0; r < world; ++r) {
    if (r == rank) { peers[r] = c->local; continue; }
    if (!all_handles) { delete c; AG_FAIL(AG_ERR_INVALID, "ag_comm_create: NULL handle table"); }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)all_handles + (size_t)r * AG_COMM_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaIpcOpenMemHandle (peer mailbox)", __FILE__, __LINE__); }
    c->opened.push_back(p);
    peers[r] = reinterpret_cast<MailSlot*>(p);
  }
  AG_CUDA_TRY(cudaMalloc((void**)&c->d_peers, sizeof(MailSlot*) * (size_t)world));
  AG_CUDA_TRY(cudaMemcpy(c->d_peers, peers.data(), sizeof(MailSlot*) * (size_t)world, cudaMemcpyHostToDevice));
  *out = reinterpret_cast<ag_comm_t>(c);
  return AG_OK;
}

// One process driving n devices: comms[k] is the communicator of devices[k] (rank k).  Needs ag_init_all (peer access).
ag_status ag_comm_create_local(ag_comm_t* comms, int n, const int* devices) {
  if (!comms || n < 1 || n > 16 || !devices) AG_FAIL(AG_ERR_INVALID, "ag_comm_create_local: bad arguments");
  int saved = -1;
  AG_TRY(ag_get_device(&saved));
  std::vector<Comm*> cs((size_t)n, nullptr);
  std::vector<MailSlot*> boxes((size_t)n, nullptr);
  ag_status rc = AG_OK;
  for (int k = 0; k < n && rc == AG_OK; ++k) {
    rc = ag_set_device(devices[k]);
    if (rc == AG_OK) rc = alloc_mailbox(n, &boxes[k]);
  }
  for (int k = 0; k < n && rc == AG_OK; ++k) {
    rc = ag_set_device(devices[k]);
    if (rc != AG_OK) break;
    Comm* c = new Comm();
    c->world = n; c->rank = k; c->device = devices[k]; c->local = boxes[k];
    if (cudaMalloc((void**)&c->d_peers, sizeof(MailSlot*) * (size_t)n) != cudaSuccess ||
        cudaMemcpy(c->d_peers, boxes.data(), sizeof(MailSlot*) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) {
      rc = cuda_fail(cudaGetLastError(), "peer table", __FILE__, __LINE__);
      delete c;
      break;
    }
    cs[k] = c;
    comms[k] = reinterpret_cast<ag_comm_t>(c);
  }
  ag_set_device(saved);
  return rc;
}

ag_status ag_comm_unique_id(void* id128) {
  if (!id128) AG_FAIL(AG_ERR_INVALID, "ag_comm_unique_id: NULL argument");
  AG_TRY(load_nccl());
  AG_NCCL_TRY(g_nccl.GetUniqueId(id128));
  return AG_OK;
}

ag_status ag_comm_attach_nccl(ag_comm_t c, const void* id128) {
  Comm* comm = reinterpret_cast<Comm*>(c);
  if (!comm || !id128) AG_FAIL(AG_ERR_INVALID, "ag_comm_attach_nccl: NULL argument");
  AG_TRY(ensure_init());
  AG_TRY(load_nccl());
  NcclId id;
  memcpy(id.bytes, id128, sizeof(id.bytes));
  void* nc = nullptr;
  AG_NCCL_TRY(reinterpret_cast<nccl_init_rank_fn>(g_nccl.CommInitRank)(&nc, comm->world, id, comm->rank));
  comm->nccl = nc;
  return AG_OK;
}

ag_status ag_comm_info(ag_comm_t c, int* world, int* rank, int* device) {
  Comm* comm = reinterpret_cast<Comm*>(c);
  if (!comm) AG_FAIL(AG_ERR_INVALID, "ag_comm_info: NULL communicator");
  if (world) *world = comm->world;
  if (rank) *rank = comm->rank;
  if (device) *device = comm->device;
  return AG_OK;
}

ag_status ag_comm_destroy(ag_comm_t c) {
  Comm* comm = reinterpret_cast<Comm*>(c);
  if (!comm) return AG_OK;
  int saved = -1;
  ag_get_device(&saved);
  ag_set_device(comm->device);
  cudaDeviceSynchronize();
  if (comm->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(comm->nccl);
  for (void* p : comm->opened) cudaIpcCloseMemHandle(p);
  if (comm->d_peers) cudaFree(comm->d_peers);
  if (comm->local && comm->owns_local) cudaFree(comm->local);
  delete comm;
  if (saved >= 0) ag_set_device(saved);
  return AG_OK;
}

}  // extern "C"
