// Common device helpers for the harmony_b200 data plane (sm_100a only).
//
// The central abstraction is TableView: the device-side image of an Elastic Table's
// ownership cache.  A table is `num_blocks` blocks; block b lives in ONE GPU's HBM and
// `block_ptrs[b]` holds the virtual address of its first row *as mapped in this process*
// (a local cudaMalloc pointer or a peer mapping opened through CUDA IPC).  Kernels resolve
// key -> (block, slot) -> address and then issue plain ld/st/red on that address; when the
// address is a peer mapping the access travels over NVLink 5 / NVSwitch with no NCCL call.
//
// Reference parity: this replaces OwnershipCache.resolveExecutor + RemoteAccessOpSender
// (services/et/.../evaluator/impl/OwnershipCache.java:123-153, TableImpl.java:366-408).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

// Host side: the number of SMs a launch of the CALLING THREAD may fill (arena.cpp).  A tasklet thread whose job
// shares the GPU with other jobs gets a fraction of the 148 SMs (`set_sm_share`): its grid-stride kernels then take
// that share of the CTA slots and the co-located jobs' kernels run next to them instead of behind them -- the GPU
// form of the reference's resource-typed task-unit slots (LocalTaskUnitScheduler.java:36-37: CPU x1, NET x2).
int sm_count();

enum PartitionMode : int { PART_ORDERED = 0, PART_HASHED = 1 };

struct TableView {
  // [num_blocks + 2]: VA of each block's row 0, then two guard words maintained by row_ptr():
  //   [num_blocks] = number of out-of-range keys seen, [num_blocks + 1] = the last offending key
  const long long* __restrict__ block_ptrs;
  int num_blocks;
  int rows_per_block;   // capacity (rows) of every block slab
  int row_stride;       // elements between consecutive rows (padded to 32 elements)
  int row_dim;          // logical elements per row
  int mode;             // PartitionMode
};

// key -> element pointer of the row. Works for any 4-byte element type.
// An out-of-range key would become a wild HBM (possibly peer) address: it is counted in the guard words of the
// pointer table and clamped to key 0, so the access stays inside the table and DeviceTable.check() raises later
// (no device sync on the hot path; cost = one 64-bit compare per row).
template <typename T>
__device__ __forceinline__ T* row_ptr(const TableView& t, long long key) {
  if ((unsigned long long)key >= (unsigned long long)t.num_blocks * (unsigned long long)t.rows_per_block) {
    long long* guard = const_cast<long long*>(t.block_ptrs) + t.num_blocks;
    atomicAdd(reinterpret_cast<unsigned long long*>(guard), 1ULL);
    guard[1] = key;
    key = 0;
  }
  int blk, slot;
  if (t.mode == PART_ORDERED) {
    blk = (int)(key / t.rows_per_block);
    slot = (int)(key - (long long)blk * t.rows_per_block);
  } else {
    blk = (int)(key % t.num_blocks);
    slot = (int)(key / t.num_blocks);
  }
  return reinterpret_cast<T*>(t.block_ptrs[blk]) + (size_t)slot * t.row_stride;
}

// ---- vector memory ops ------------------------------------------------------------------
// Weak (non-.nc) 128-bit load that does not allocate in L1: table rows may be concurrently
// updated by other workers' pushes (async PS), so the read-only .nc path is not legal.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ld_stream_i4(const int* p) {
  int4 r;
  asm("ld.global.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// ---- L2 residency hints -------------------------------------------------------------------
// Small, re-read buffers (the [B, C] error matrix of MLR) are tagged evict_last so the streaming
// row traffic (GBs per step) does not push them out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_hint_f4(const float* p, uint64_t pol) {
  float4 r;
  asm("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_hint_f4(float* p, const float4& v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_f4_sys_hint(float* p, const float4& v, uint64_t pol) {
  asm volatile("red.relaxed.sys.global.add.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// 128-bit fire-and-forget float reduction (sm_90+). System scope so that it is valid on a
// peer-mapped address: this is the PUSH primitive (reduce-write into the owner's HBM).
__device__ __forceinline__ void red_add_f4_sys(float* p, const float4& v) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_f32_sys(float* p, float v) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float atom_add_f32_sys(float* p, float v) {
  float old;
  asm volatile("atom.relaxed.sys.global.add.f32 %0, [%1], %2;" : "=f"(old) : "l"(p), "f"(v) : "memory");
  return old;
}
__device__ __forceinline__ int atom_add_s32_sys(int* p, int v) {
  int old;
  asm volatile("atom.relaxed.sys.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned atom_cas_b32_sys(unsigned* p, unsigned cmp, unsigned val) {
  unsigned old;
  asm volatile("atom.relaxed.sys.global.cas.b32 %0, [%1], %2, %3;"
               : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ void st_release_sys_s32(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

static constexpr long long HB_SPIN_TIMEOUT = 4000000000LL;   // ~2 s of SM clocks: raise, never hang the GPU

// ---- multimem (NVLS) primitives: the address must be a multicast mapping -------------------------------
__device__ __forceinline__ float4 multimem_ld_reduce_add_f4(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_release_sys_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Bounded spin until *p >= target (wrap-safe).  Returns false on timeout after raising `err`.
// The poll is a RELAXED system-scope load (served by L2, never stale): an acquire load per poll would invalidate
// the SM's L1 every few hundred nanoseconds and slow the kernels that share the SM with a resident spinner
// (measured: +22% on K1 with K2 resident).  One acquire fence follows the successful poll.
__device__ __forceinline__ unsigned ld_relaxed_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool spin_until_ge(const unsigned* p, unsigned target, int* err, int code) {
  const long long t0 = clock64();
  bool ok = true;
  unsigned ns = 32;
  while ((int)(ld_relaxed_sys_u32(p) - target) < 0) {
    if (clock64() - t0 > HB_SPIN_TIMEOUT) { atomicExch(err, code); ok = false; break; }
    __nanosleep(ns);
    if (ns < 256) ns <<= 1;
  }
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  return ok;
}

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide reductions for blockDim.x <= 1024, `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  __syncthreads();
  return r;
}

// ---- Philox-lite counter RNG (for table initialisers and the LDA sampler) -------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t rng_u32(uint64_t seed, uint64_t ctr) {
  uint32_t a = mix32((uint32_t)(ctr) ^ (uint32_t)(seed));
  uint32_t b = mix32((uint32_t)(ctr >> 32) ^ (uint32_t)(seed >> 32) ^ 0x9e3779b9U);
  return mix32(a ^ (b * 0x85ebca6bU) ^ 0xc2b2ae35U);
}
__device__ __forceinline__ float rng_uniform(uint64_t seed, uint64_t ctr) {
  return (rng_u32(seed, ctr) >> 8) * (1.0f / 16777216.0f);  // [0,1)
}
__device__ __forceinline__ float rng_normal(uint64_t seed, uint64_t ctr) {
  float u1 = fmaxf(rng_uniform(seed, 2 * ctr), 1e-7f), u2 = rng_uniform(seed, 2 * ctr + 1);
  return sqrtf(-2.f * __logf(u1)) * __cosf(6.28318530718f * u2);
}

}  // namespace hb
