// VMM + NVLS plumbing of the peer arena (driver API, resolved at run time so that the extension still
// loads on a CPU-only box).
//
// A VMM segment is physical HBM created with cuMemCreate (POSIX-fd shareable), mapped into a reserved
// VA range of the owning process and -- through the exported fd -- into every peer process, so loads /
// stores / reductions on the mapped range travel over NVLink 5 exactly like the CUDA-IPC segments of
// arena.cpp.  What VMM adds is the multicast object: cuMulticastCreate -> every rank adds its device
// and binds the SAME offset range of its own segment -> the object is mapped at `mc_va`.  A
// `multimem.st` on mc_va + off lands in every GPU's segment at `off` (in-switch broadcast) and a
// `multimem.ld_reduce` returns the in-switch sum of the 8 replicas: the NVSwitch (NVLS) data plane that
// replaces the reference's per-block message fan-out (TableImpl.java:366-408,460-501).
#include <torch/extension.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace hb {

namespace {

struct DriverApi {
  bool ok = false;
  std::string why;
  decltype(&cuMemCreate) MemCreate = nullptr;
  decltype(&cuMemRelease) MemRelease = nullptr;
  decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
  decltype(&cuMemAddressFree) MemAddressFree = nullptr;
  decltype(&cuMemMap) MemMap = nullptr;
  decltype(&cuMemUnmap) MemUnmap = nullptr;
  decltype(&cuMemSetAccess) MemSetAccess = nullptr;
  decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
  decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
  decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
  decltype(&cuMulticastCreate) MulticastCreate = nullptr;
  decltype(&cuMulticastAddDevice) MulticastAddDevice = nullptr;
  decltype(&cuMulticastBindMem) MulticastBindMem = nullptr;
  decltype(&cuMulticastUnbind) MulticastUnbind = nullptr;
  decltype(&cuMulticastGetGranularity) MulticastGetGranularity = nullptr;
  decltype(&cuDeviceGet) DeviceGet = nullptr;
  decltype(&cuDeviceGetAttribute) DeviceGetAttribute = nullptr;
  decltype(&cuGetErrorString) GetErrorString = nullptr;
};

template <typename F>
bool resolve(const char* name, F& fn, std::string& why) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || p == nullptr) {
    cudaGetLastError();
    why = std::string("driver entry point ") + name + " is not available";
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

DriverApi& api() {
  static DriverApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    bool ok = true;
#define HB_RESOLVE(N) ok = ok && resolve("cu" #N, a.N, a.why)
    HB_RESOLVE(MemCreate);
    HB_RESOLVE(MemRelease);
    HB_RESOLVE(MemAddressReserve);
    HB_RESOLVE(MemAddressFree);
    HB_RESOLVE(MemMap);
    HB_RESOLVE(MemUnmap);
    HB_RESOLVE(MemSetAccess);
    HB_RESOLVE(MemGetAllocationGranularity);
    HB_RESOLVE(MemExportToShareableHandle);
    HB_RESOLVE(MemImportFromShareableHandle);
    HB_RESOLVE(MulticastCreate);
    HB_RESOLVE(MulticastAddDevice);
    HB_RESOLVE(MulticastBindMem);
    HB_RESOLVE(MulticastUnbind);
    HB_RESOLVE(MulticastGetGranularity);
    HB_RESOLVE(DeviceGet);
    HB_RESOLVE(DeviceGetAttribute);
    HB_RESOLVE(GetErrorString);
#undef HB_RESOLVE
    a.ok = ok;
  });
  return a;
}

std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (api().GetErrorString && api().GetErrorString(r, &s) == CUDA_SUCCESS && s) return std::string(s);
  return "CUresult " + std::to_string((int)r);
}

#define HB_CU_CHECK(expr)                                                               \
  do {                                                                                  \
    CUresult _r = (expr);                                                               \
    TORCH_CHECK(_r == CUDA_SUCCESS, "driver error '", cu_err(_r), "' at " #expr);       \
  } while (0)

struct Scope {   // make `device`'s primary context current for one scope (see arena.cpp DeviceScope)
  int prev = -1;
  explicit Scope(int device) {
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device); else prev = -1;
    cudaFree(nullptr);   // force context creation
  }
  ~Scope() { if (prev >= 0) cudaSetDevice(prev); }
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

CUmemAllocationProp seg_prop(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

void set_access(CUdeviceptr va, size_t bytes, int device) {
  CUmemAccessDesc d{};
  d.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  d.location.id = device;
  d.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  HB_CU_CHECK(api().MemSetAccess(va, bytes, &d, 1));
}

struct Mapping {          // one mapped range (own segment, imported peer segment, or multicast object)
  CUmemGenericAllocationHandle handle = 0;
  CUdeviceptr va = 0;
  size_t bytes = 0;
  int device = 0;
  bool owns_physical = false;
};

std::mutex g_mu;
std::map<int64_t, Mapping> g_maps;          // base VA -> mapping
struct McObject { CUmemGenericAllocationHandle handle = 0; size_t bytes = 0; int world = 0; };
std::map<int64_t, McObject> g_mcs;          // id -> multicast object
int64_t g_next_mc = 1;

}  // namespace

// (posix-fd shareable VMM supported, multicast supported) on `device`; (false, false) without a driver.
std::vector<bool> vmm_supported(int64_t device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device >= n) { cudaGetLastError(); return {false, false}; }
  if (!api().ok) return {false, false};
  CUdevice dev;
  if (api().DeviceGet(&dev, (int)device) != CUDA_SUCCESS) return {false, false};
  int fd_ok = 0, mc_ok = 0, vmm_ok = 0;
  api().DeviceGetAttribute(&vmm_ok, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
  api().DeviceGetAttribute(&fd_ok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
  api().DeviceGetAttribute(&mc_ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
  return {vmm_ok != 0 && fd_ok != 0, mc_ok != 0};
}

int64_t vmm_granularity(int64_t device) {
  TORCH_CHECK(api().ok, api().why);
  CUmemAllocationProp p = seg_prop((int)device);
  size_t g = 0;
  HB_CU_CHECK(api().MemGetAllocationGranularity(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  return (int64_t)g;
}

// Physical segment on `device`, mapped read-write for `device`.  Returns (base VA, actual bytes).
std::vector<int64_t> vmm_create(int64_t device, int64_t bytes, int64_t align) {
  TORCH_CHECK(api().ok, api().why);
  Scope scope((int)device);
  CUmemAllocationProp p = seg_prop((int)device);
  size_t gran = 0;
  HB_CU_CHECK(api().MemGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  gran = std::max(gran, (size_t)std::max<int64_t>(align, 1));
  Mapping m;
  m.bytes = align_up((size_t)bytes, gran);
  m.device = (int)device;
  m.owns_physical = true;
  HB_CU_CHECK(api().MemCreate(&m.handle, m.bytes, &p, 0));
  HB_CU_CHECK(api().MemAddressReserve(&m.va, m.bytes, gran, 0, 0));
  HB_CU_CHECK(api().MemMap(m.va, m.bytes, 0, m.handle, 0));
  set_access(m.va, m.bytes, (int)device);
  std::lock_guard<std::mutex> g(g_mu);
  g_maps[(int64_t)m.va] = m;
  return {(int64_t)m.va, (int64_t)m.bytes};
}

static Mapping get_map(int64_t base) {
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_maps.find(base);
  TORCH_CHECK(it != g_maps.end(), "unknown VMM mapping");
  return it->second;
}

// POSIX fd of the segment's physical allocation (the caller passes it over a unix socket and closes it).
int64_t vmm_export_fd(int64_t base) {
  Mapping m = get_map(base);
  int fd = -1;
  HB_CU_CHECK(api().MemExportToShareableHandle(&fd, m.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  return fd;
}

// Map a peer's segment (received fd) for `device`; returns the base VA in this process.
int64_t vmm_import_fd(int64_t fd, int64_t bytes, int64_t device) {
  TORCH_CHECK(api().ok, api().why);
  Scope scope((int)device);
  Mapping m;
  m.bytes = (size_t)bytes;
  m.device = (int)device;
  HB_CU_CHECK(api().MemImportFromShareableHandle(&m.handle, (void*)(uintptr_t)fd,
                                                  CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  size_t gran = (size_t)vmm_granularity(device);
  HB_CU_CHECK(api().MemAddressReserve(&m.va, m.bytes, gran, 0, 0));
  HB_CU_CHECK(api().MemMap(m.va, m.bytes, 0, m.handle, 0));
  set_access(m.va, m.bytes, (int)device);
  std::lock_guard<std::mutex> g(g_mu);
  g_maps[(int64_t)m.va] = m;
  return (int64_t)m.va;
}

// Same-process executors driving another GPU: the VA is process-wide, grant that device access.
void vmm_grant(int64_t base, int64_t device) {
  Mapping m = get_map(base);
  set_access(m.va, m.bytes, (int)device);
}

void vmm_release(int64_t base) {
  Mapping m;
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_maps.find(base);
    if (it == g_maps.end()) return;
    m = it->second;
    g_maps.erase(it);
  }
  Scope scope(m.device);
  api().MemUnmap(m.va, m.bytes);
  api().MemAddressFree(m.va, m.bytes);
  api().MemRelease(m.handle);
}

// ---- multicast (NVLS) ---------------------------------------------------------------------------
int64_t mc_granularity(int64_t world, int64_t bytes, bool recommended) {
  TORCH_CHECK(api().ok, api().why);
  CUmulticastObjectProp p{};
  p.numDevices = (unsigned)world;
  p.size = (size_t)bytes;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t g = 0;
  HB_CU_CHECK(api().MulticastGetGranularity(&g, &p, recommended ? CU_MULTICAST_GRANULARITY_RECOMMENDED
                                                                 : CU_MULTICAST_GRANULARITY_MINIMUM));
  return (int64_t)g;
}

// Rank 0: create the object for `world` devices; returns (mc id, fd to hand to the other ranks).
std::vector<int64_t> mc_create(int64_t device, int64_t world, int64_t bytes) {
  TORCH_CHECK(api().ok, api().why);
  Scope scope((int)device);
  CUmulticastObjectProp p{};
  p.numDevices = (unsigned)world;
  p.size = (size_t)bytes;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  McObject o;
  o.bytes = (size_t)bytes;
  o.world = (int)world;
  HB_CU_CHECK(api().MulticastCreate(&o.handle, &p));
  int fd = -1;
  HB_CU_CHECK(api().MemExportToShareableHandle(&fd, o.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  std::lock_guard<std::mutex> g(g_mu);
  int64_t id = g_next_mc++;
  g_mcs[id] = o;
  return {id, (int64_t)fd};
}

int64_t mc_import(int64_t fd, int64_t device, int64_t world, int64_t bytes) {
  TORCH_CHECK(api().ok, api().why);
  Scope scope((int)device);
  McObject o;
  o.bytes = (size_t)bytes;
  o.world = (int)world;
  HB_CU_CHECK(api().MemImportFromShareableHandle(&o.handle, (void*)(uintptr_t)fd,
                                                  CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  std::lock_guard<std::mutex> g(g_mu);
  int64_t id = g_next_mc++;
  g_mcs[id] = o;
  return id;
}

static McObject get_mc(int64_t id) {
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_mcs.find(id);
  TORCH_CHECK(it != g_mcs.end(), "unknown multicast object");
  return it->second;
}

void mc_add_device(int64_t id, int64_t device) {
  McObject o = get_mc(id);
  Scope scope((int)device);
  CUdevice dev;
  HB_CU_CHECK(api().DeviceGet(&dev, (int)device));
  HB_CU_CHECK(api().MulticastAddDevice(o.handle, dev));
}

// Bind [mem_offset, mem_offset + bytes) of the local VMM segment `seg_base` at `mc_offset` of the object.
// Blocks until every rank has added its device.
void mc_bind(int64_t id, int64_t mc_offset, int64_t seg_base, int64_t mem_offset, int64_t bytes) {
  McObject o = get_mc(id);
  Mapping m = get_map(seg_base);
  Scope scope(m.device);
  HB_CU_CHECK(api().MulticastBindMem(o.handle, (size_t)mc_offset, m.handle, (size_t)mem_offset, (size_t)bytes, 0));
}

// Map the whole object for `device`; returns the multicast VA.
int64_t mc_map(int64_t id, int64_t device) {
  McObject o = get_mc(id);
  Scope scope((int)device);
  Mapping m;
  m.bytes = o.bytes;
  m.device = (int)device;
  m.handle = o.handle;
  size_t gran = (size_t)mc_granularity(o.world, (int64_t)o.bytes, true);
  HB_CU_CHECK(api().MemAddressReserve(&m.va, m.bytes, gran, 0, 0));
  HB_CU_CHECK(api().MemMap(m.va, m.bytes, 0, m.handle, 0));
  set_access(m.va, m.bytes, (int)device);
  return (int64_t)m.va;
}

void mc_unmap(int64_t va, int64_t bytes) {
  if (!api().ok) return;
  api().MemUnmap((CUdeviceptr)va, (size_t)bytes);
  api().MemAddressFree((CUdeviceptr)va, (size_t)bytes);
}

void mc_destroy(int64_t id, int64_t device) {
  McObject o;
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_mcs.find(id);
    if (it == g_mcs.end()) return;
    o = it->second;
    g_mcs.erase(it);
  }
  Scope scope((int)device);
  CUdevice dev;
  if (api().DeviceGet(&dev, (int)device) == CUDA_SUCCESS) api().MulticastUnbind(o.handle, dev, 0, o.bytes);
  api().MemRelease(o.handle);
}

void close_fd(int64_t fd) { if (fd >= 0) ::close((int)fd); }

}  // namespace hb
