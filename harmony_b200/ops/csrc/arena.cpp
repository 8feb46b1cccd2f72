// Peer arena: the HBM side of the Elastic Table.
//
// Every executor process owns ONE GPU and allocates table slabs from a cudaMalloc'ed arena that
// is exported with CUDA IPC; every other executor opens the handle so that the whole table is
// addressable from every GPU (loads / stores / reductions travel over NVLink 5 / NVSwitch).
// This replaces the reference's executor-to-executor Netty/NCS transport
// (services/et/.../common/impl/NetworkConnectionImpl.java:72-95).
//
// The arena is a bump allocator with a free list (block slabs are large and few); all
// allocations are 512-byte aligned so that 128-bit vector accesses and bulk copies are legal.
#include <torch/extension.h>
#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace hb {

#define HB_CUDA_CHECK(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    TORCH_CHECK(_e == cudaSuccess, "CUDA error ", cudaGetErrorString(_e), " at " #expr);         \
  } while (0)

// vmm.cpp
std::vector<int64_t> vmm_create(int64_t device, int64_t bytes, int64_t align);
void vmm_release(int64_t base);

struct Arena {
  int kind = 0;           // 0 = cudaMalloc segment exported with CUDA IPC, 1 = VMM segment (cuMemCreate, fd shareable)
  char* base = nullptr;
  size_t capacity = 0;
  size_t bump = 0;
  int device = 0;
  std::multimap<size_t, size_t> free_by_size;  // size -> offset
  std::map<size_t, size_t> live;               // offset -> size
  std::mutex mu;
};

// Switch the calling thread to `device` for one scope and put it back afterwards: these entry points are called
// from threads that drive OTHER GPUs too (in-process executors, tests), and a leaked cudaSetDevice silently
// redirects the caller's later `.cuda()` allocations and raw-pointer launches to the wrong GPU.
struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int device) {
    cudaGetDevice(&prev);
    if (prev != device) HB_CUDA_CHECK(cudaSetDevice(device));
    else prev = -1;
  }
  ~DeviceScope() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static std::map<int64_t, Arena*> g_arenas;
static std::mutex g_mu;

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- per-thread SM budget (see hb_common.cuh) ----
static thread_local double t_sm_share = 1.0;
void set_sm_share(double share) { t_sm_share = share <= 0.0 ? 1.0 : (share > 1.0 ? 1.0 : share); }
double get_sm_share() { return t_sm_share; }
int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
      sms = n;
    else { cudaGetLastError(); sms = 148; }
  }
  const int n = (int)(sms * t_sm_share + 0.5);
  return n < 1 ? 1 : n;
}

int64_t arena_create(int64_t device, int64_t bytes, int64_t kind, int64_t align) {
  auto* a = new Arena();
  a->device = (int)device;
  a->kind = (int)kind;
  DeviceScope scope((int)device);
  if (kind == 1) {
    auto r = vmm_create(device, bytes, align);
    a->base = reinterpret_cast<char*>(r[0]);
    a->capacity = (size_t)r[1];
  } else {
    a->capacity = align_up((size_t)bytes, 2u << 20);
    HB_CUDA_CHECK(cudaMalloc((void**)&a->base, a->capacity));
  }
  std::lock_guard<std::mutex> g(g_mu);
  int64_t id = reinterpret_cast<int64_t>(a->base);
  g_arenas[id] = a;
  return id;
}

static Arena* get_arena(int64_t id) {
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_arenas.find(id);
  TORCH_CHECK(it != g_arenas.end(), "unknown arena");
  return it->second;
}

// returns the device pointer of the allocation (0 on exhaustion)
int64_t arena_alloc(int64_t id, int64_t bytes) {
  Arena* a = get_arena(id);
  std::lock_guard<std::mutex> g(a->mu);
  size_t need = align_up((size_t)std::max<int64_t>(bytes, 1), 512);
  auto it = a->free_by_size.lower_bound(need);
  size_t off;
  if (it != a->free_by_size.end()) {
    off = it->second;
    size_t sz = it->first;
    a->free_by_size.erase(it);
    if (sz > need) a->free_by_size.emplace(sz - need, off + need);
  } else {
    if (a->bump + need > a->capacity) return 0;
    off = a->bump;
    a->bump += need;
  }
  a->live[off] = need;
  return reinterpret_cast<int64_t>(a->base + off);
}

void arena_free(int64_t id, int64_t ptr) {
  Arena* a = get_arena(id);
  std::lock_guard<std::mutex> g(a->mu);
  size_t off = (size_t)(reinterpret_cast<char*>(ptr) - a->base);
  auto it = a->live.find(off);
  TORCH_CHECK(it != a->live.end(), "arena_free of unknown pointer");
  size_t sz = it->second;
  a->live.erase(it);
  // coalesce with the bump frontier when possible, else keep on the free list
  if (off + sz == a->bump) a->bump = off;
  else a->free_by_size.emplace(sz, off);
}

std::vector<int64_t> arena_stats(int64_t id) {
  Arena* a = get_arena(id);
  std::lock_guard<std::mutex> g(a->mu);
  size_t live = 0;
  for (auto& kv : a->live) live += kv.second;
  return {(int64_t)a->capacity, (int64_t)a->bump, (int64_t)live};
}

void arena_destroy(int64_t id) {
  Arena* a;
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_arenas.find(id);
    if (it == g_arenas.end()) return;
    a = it->second;
    g_arenas.erase(it);
  }
  if (a->kind == 1) {
    vmm_release(reinterpret_cast<int64_t>(a->base));
  } else {
    DeviceScope scope(a->device);
    cudaFree(a->base);
  }
  delete a;
}

// ---- CUDA IPC --------------------------------------------------------------------------------
py::bytes ipc_export(int64_t ptr) {
  cudaIpcMemHandle_t h;
  HB_CUDA_CHECK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(ptr)));
  return py::bytes(reinterpret_cast<const char*>(&h), sizeof(h));
}

int64_t ipc_open(std::string handle_bytes, int64_t device) {
  TORCH_CHECK(handle_bytes.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_bytes.data(), sizeof(h));
  DeviceScope scope((int)device);
  void* p = nullptr;
  HB_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return reinterpret_cast<int64_t>(p);
}

void ipc_close(int64_t ptr) { cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr)); }

// Same-process multi-GPU (tests / single-process launcher): enable direct peer access.
bool enable_peer_access(int64_t device, int64_t peer) {
  if (device == peer) return true;
  int can = 0;
  HB_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, (int)device, (int)peer));
  if (!can) return false;
  DeviceScope scope((int)device);
  cudaError_t e = cudaDeviceEnablePeerAccess((int)peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return true; }
  HB_CUDA_CHECK(e);
  return true;
}

// View raw device memory as a tensor (no ownership).
at::Tensor tensor_from_ptr(int64_t ptr, std::vector<int64_t> sizes, std::vector<int64_t> strides,
                           int64_t dtype_code, int64_t device) {
  at::ScalarType st = dtype_code == 0 ? at::kFloat : (dtype_code == 1 ? at::kInt : (dtype_code == 2 ? at::kLong : at::kByte));
  auto opts = at::TensorOptions().dtype(st).device(at::kCUDA, (int)device);
  return at::from_blob(reinterpret_cast<void*>(ptr), sizes, strides, [](void*) {}, opts);
}

}  // namespace hb
