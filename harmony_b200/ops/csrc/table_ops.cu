// Generic Elastic-Table data plane kernels (K1 PULL, K2 PUSH fused with K3/K4/K5 update
// functions, K10 block migration, table initialisers, device clocks).
//
// Reference parity (behaviour, not code):
//   PULL  = Table.multiGetOrInit  -> services/et/.../evaluator/impl/TableImpl.java:366-408
//   PUSH  = Table.multiUpdate     -> TableImpl.java:460-501 + BlockImpl.update (BlockImpl.java:65-76)
//   update functions              -> MLRETModelUpdateFunction.java:51-62 (ADD),
//                                    NMFETModelUpdateFunction.java:43-52 (SGD + clamp),
//                                    LDAETModelUpdateFunction.java:38-64 (sparse int add, clamp>=0, nnz)
//   migration                     -> MigrationExecutor.java:338-414 (chunked block copy)
// The B200 design replaces per-block Avro messages with ONE kernel whose loads / reductions
// land directly in the owner GPU's HBM through peer mappings (NVLink 5 / NVSwitch).
#include "hb_common.cuh"
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

namespace hb {

enum UpdateFn : int {
  UPD_ADD = 0,          // w += d                                   (MLR / AddVector / Lasso)
  UPD_SGD_CLAMP = 1,    // w = clamp(w - step*d, lo, hi)           (NMF)
  UPD_ADAGRAD = 2,      // h += d^2 ; w -= step*d/sqrt(h+eps)      (north-star option)
  UPD_ASSIGN = 3,       // w = d                                   (put)
};

struct UpdateParams {
  float step, lo, hi, eps;
};

// ------------------------------------------------------------------------------------------
// PULL: out[i, :] = table[keys[i], :]      (one warp per row, 128-bit loads)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pull_rows_kernel(TableView t, const long long* __restrict__ keys,
                                                        int n, float* __restrict__ out, int out_stride,
                                                        int out_width, int vec_ok) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int vec = t.row_stride >> 2;  // float4 per row (stride is a multiple of 32 floats)
  for (int i = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_cta) {
    const float* src = row_ptr<float>(t, keys[i]);
    float* dst = out + (size_t)i * out_stride;
#pragma unroll 4
    for (int v = lane; v < vec; v += 32) {
      float4 x = ld_stream_f4(src + 4 * v);
      if (vec_ok && 4 * v + 3 < out_width) {
        *reinterpret_cast<float4*>(dst + 4 * v) = x;
      } else {
        float xs[4] = {x.x, x.y, x.z, x.w};
        for (int e = 0; e < 4; ++e) if (4 * v + e < out_width) dst[4 * v + e] = xs[e];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// PUSH fused with the update function, applied in the owner's HBM.
//   ADD        : 128-bit red.add (no return value, fully fire-and-forget)
//   SGD_CLAMP  : atom.add returning the old value; the rare out-of-range result is projected
//                back with a CAS loop, so the row is never left outside [lo, hi]
//   ADAGRAD    : table row = [w(0..D) | h(0..D)] : h via atom.add (returns old), w via red.add
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void clamp_fixup(float* p, float lo, float hi) {
  unsigned* up = reinterpret_cast<unsigned*>(p);
  unsigned cur = *reinterpret_cast<volatile unsigned*>(up);
  for (int it = 0; it < 64; ++it) {
    float f = __uint_as_float(cur);
    float c = fminf(fmaxf(f, lo), hi);
    if (c == f) break;
    unsigned prev = atom_cas_b32_sys(up, cur, __float_as_uint(c));
    if (prev == cur) break;
    cur = prev;
  }
}

__global__ void __launch_bounds__(256) push_rows_kernel(TableView t, const long long* __restrict__ keys, int n,
                                                        const float* __restrict__ deltas, int d_stride,
                                                        int fn, UpdateParams up, int vec_ok) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = t.row_dim;
  for (int i = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_cta) {
    float* dst = row_ptr<float>(t, keys[i]);
    const float* src = deltas + (size_t)i * d_stride;
    if (fn == UPD_ADD) {
      const int vec = vec_ok ? (D >> 2) : 0;
      for (int v = lane; v < vec; v += 32) {
        float4 x = *reinterpret_cast<const float4*>(src + 4 * v);
        red_add_f4_sys(dst + 4 * v, x);
      }
      for (int e = (vec << 2) + lane; e < D; e += 32) red_add_f32_sys(dst + e, src[e]);
    } else if (fn == UPD_SGD_CLAMP) {
      for (int e = lane; e < D; e += 32) {
        float d = -up.step * src[e];
        float old = atom_add_f32_sys(dst + e, d);
        float nv = old + d;
        if (nv < up.lo || nv > up.hi) clamp_fixup(dst + e, up.lo, up.hi);
      }
    } else if (fn == UPD_ADAGRAD) {
      // second half of the (2*D wide) row holds the accumulated squared gradients
      float* h = dst + (t.row_stride >> 1);
      for (int e = lane; e < D; e += 32) {
        float g = src[e];
        float hold = atom_add_f32_sys(h + e, g * g);
        red_add_f32_sys(dst + e, -up.step * g * rsqrtf(hold + g * g + up.eps));
      }
    } else {  // ASSIGN
      for (int e = lane; e < D; e += 32) dst[e] = src[e];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Integer tables (LDA): pull rows; push sparse (row, topic, delta) triples.
// Row layout int[K+1], last element = number of non-zero topics.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pull_rows_i32_kernel(TableView t, const long long* __restrict__ keys,
                                                            int n, int* __restrict__ out, int out_stride,
                                                            int out_width) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int vec = t.row_stride >> 2;
  for (int i = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_cta) {
    const int* src = row_ptr<int>(t, keys[i]);
    int* dst = out + (size_t)i * out_stride;
    for (int v = lane; v < vec; v += 32) {
      int4 x = ld_stream_i4(src + 4 * v);
      int xs[4] = {x.x, x.y, x.z, x.w};
      for (int e = 0; e < 4; ++e) if (4 * v + e < out_width) dst[4 * v + e] = xs[e];
    }
  }
}

// One thread per (key, topic, delta) triple. atom.add returns the old count so that the
// clamp-at-zero and the nnz bookkeeping of the reference update function are preserved:
// a decrement below zero is compensated by adding the deficit back.
__global__ void __launch_bounds__(256) push_sparse_i32_kernel(TableView t, const long long* __restrict__ keys,
                                                              const int* __restrict__ topics,
                                                              const int* __restrict__ deltas, int n, int K,
                                                              int bitmap_off) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = deltas[i];
  if (d == 0) return;
  int* row = row_ptr<int>(t, keys[i]);
  int old = atom_add_s32_sys(row + topics[i], d);
  int nv = old + d;
  if (nv < 0) {  // clamp >= 0 : give the deficit back
    atom_add_s32_sys(row + topics[i], -nv);
    nv = 0;
  }
  if (old <= 0 && nv > 0) {
    atom_add_s32_sys(row + K, 1);
    // topic bitmap of the sparse PULL: bits are only ever SET here (set / clear do not commute under concurrent
    // +1 / -1 of different workers); stale bits cost a 16 B chunk read and are dropped by lda_rebuild_bitmaps
    if (bitmap_off > 0) {
      unsigned* w = reinterpret_cast<unsigned*>(row + bitmap_off) + (topics[i] >> 5);
      asm volatile("red.relaxed.sys.global.or.b32 [%0], %1;" :: "l"(w), "r"(1u << (topics[i] & 31)) : "memory");
    }
  } else if (old > 0 && nv == 0) {
    atom_add_s32_sys(row + K, -1);
  }
}

// Sparse PULL of LDA count rows (SparseArrayCodec.java:58-101 sends nnz-prefixed (index, value) pairs; here the
// row's topic BITMAP -- 128 B -- is read first and then only the 16 B chunks that hold a non-zero count, so a
// word with 10 topics costs ~300 B of NVLink / HBM traffic instead of 4 KB).  The output row is the dense
// staging row the sampler expects (zero-filled locally).  One warp per row; lane l owns topics [32 l, 32 l + 32).
__global__ void __launch_bounds__(256) pull_rows_lda_sparse_kernel(TableView t, const long long* __restrict__ keys,
                                                                   int n, int* __restrict__ out, int out_stride,
                                                                   int K, int bitmap_off,
                                                                   unsigned long long* __restrict__ chunk_count) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  unsigned long long mine = 0;
  for (int i = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_cta) {
    const int* src = row_ptr<int>(t, keys[i]);
    int* dst = out + (size_t)i * out_stride;
    const unsigned word = (unsigned)src[bitmap_off + lane];          // 128 B: the whole bitmap in one transaction
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int ch = 32 * q + lane;                                  // chunk (4 topics): coalesced across the warp
      const unsigned w = __shfl_sync(0xffffffffu, word, ch >> 3);    // bitmap word of topics [32 (ch / 8), +32)
      const int c = 4 * ch;
      if (c < out_stride) {
        int4 x = make_int4(0, 0, 0, 0);
        if ((w >> (4 * (ch & 7))) & 0xFu) { x = ld_stream_i4(src + c); ++mine; }
        *reinterpret_cast<int4*>(dst + c) = x;                       // the staging row is dense: zeros elsewhere
      }
    }
    __syncwarp();
    if (lane == 0) dst[K] = src[K];                                  // the row's non-zero counter rides along
  }
  if (chunk_count != nullptr && mine) atomicAdd(chunk_count, mine);
}

// Owner-side: recompute the bitmaps of a list of LOCAL slabs from the counts (drops stale bits).
__global__ void __launch_bounds__(256) lda_rebuild_bitmaps_kernel(const long long* __restrict__ slab_ptrs, int n_blocks,
                                                                  int rows_per_block, int row_stride, int K,
                                                                  int bitmap_off) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)n_blocks * rows_per_block;
  for (long long r = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); r < total; r += (long long)gridDim.x * warps_per_cta) {
    int* row = reinterpret_cast<int*>(slab_ptrs[r / rows_per_block]) + (size_t)(r % rows_per_block) * row_stride;
    unsigned word = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = 32 * lane + 4 * q;
      if (c < K) {
        const int4 x = *reinterpret_cast<const int4*>(row + c);
        const int xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c + e < K && xs[e] > 0) word |= 1u << (4 * q + e);
      }
    }
    reinterpret_cast<unsigned*>(row + bitmap_off)[lane] = word;
  }
}

// Integer scatter-reduce of one value per key (Pregel message combiners, K13): op 0 = add, 1 = min.
// The combine runs in the owner's memory system: messages to the same vertex from any GPU meet there.
__global__ void __launch_bounds__(256) push_scalar_i32_kernel(TableView t, const long long* __restrict__ keys,
                                                              const int* __restrict__ values, int n, int op) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int* p = row_ptr<int>(t, keys[i]);
  if (op == 0) asm volatile("red.relaxed.sys.global.add.s32 [%0], %1;" :: "l"(p), "r"(values[i]) : "memory");
  else asm volatile("red.relaxed.sys.global.min.s32 [%0], %1;" :: "l"(p), "r"(values[i]) : "memory");
}

// ------------------------------------------------------------------------------------------
// Initialisers: eager initValue for every row of a block slab (lazy init in the reference).
// ------------------------------------------------------------------------------------------
enum InitKind : int { INIT_ZERO = 0, INIT_GAUSSIAN = 1, INIT_UNIFORM = 2, INIT_CONST = 3 };

__global__ void init_slab_kernel(float* slab, long long rows, int row_stride, int row_dim, int kind,
                                 float a, float b, unsigned long long seed, long long row_key_base,
                                 long long key_stride) {
  long long total = rows * row_stride;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / row_stride;
    int c = (int)(idx - r * row_stride);
    float v = 0.f;
    if (c < row_dim) {
      // counter derived from the GLOBAL key so that init is independent of block placement
      unsigned long long ctr = (unsigned long long)(row_key_base + r * key_stride) * (unsigned long long)row_dim + c;
      if (kind == INIT_GAUSSIAN) v = a + b * rng_normal(seed, ctr);        // mean a, std b
      else if (kind == INIT_UNIFORM) v = a + (b - a) * rng_uniform(seed, ctr);
      else if (kind == INIT_CONST) v = a;
    }
    slab[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------
// Block migration (K10): bulk copy of a block slab from (peer) src to (local or peer) dst.
// 128-bit grid-stride copy with 4 loads in flight per thread; run on a side stream.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) migrate_copy_kernel(const int4* __restrict__ src, int4* __restrict__ dst,
                                                           long long n_vec) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n_vec; i += 4 * stride) {
    int4 a = ld_stream_i4(reinterpret_cast<const int*>(src + i));
    int4 b = ld_stream_i4(reinterpret_cast<const int*>(src + i + stride));
    int4 c = ld_stream_i4(reinterpret_cast<const int*>(src + i + 2 * stride));
    int4 d = ld_stream_i4(reinterpret_cast<const int*>(src + i + 3 * stride));
    dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
  }
  for (; i < n_vec; i += stride) dst[i] = ld_stream_i4(reinterpret_cast<const int*>(src + i));
}


// ------------------------------------------------------------------------------------------
// Live block migration, second generation (K10): a LIST of blocks moves in ONE launch with the
// bulk-copy engine (cp.async.bulk = TMA 1-D, SASS UBLKCP), run by the SENDER on its side stream:
//     old slab (local HBM) --TMA--> shared-memory stage --TMA--> new slab (peer HBM over NVLink)
//                                                      \--TMA--> snapshot (local HBM)
// One elected thread per CTA drives a ring of MIG_STAGES x MIG_CHUNK bytes (4 x 48 KB in flight per
// SM); the other threads idle.  Accessors are NOT stopped: pushes that still land in the old slab
// after its bytes were read (stale pointer, in-flight kernels) are forwarded later by
// migrate_drain_kernel as  new += old - snapshot  -- exact for the additive update functions.
// Reference: MigrationExecutor.java:134-295 (ownership-first move; redirected stale accesses).
// ------------------------------------------------------------------------------------------
static constexpr int MIG_CHUNK = 48 * 1024;
static constexpr int MIG_STAGES = 4;

__device__ __forceinline__ void mig_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mig_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mig_mbar_wait(uint32_t bar, uint32_t parity) {
  for (long long it = 0; it < 400000000LL; ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;     // bounded: a lost transaction must not hang the GPU
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}

// work item = (block, chunk); item i -> block i / chunks_per_block.  snap_ptrs may be null (plain copy).
__global__ void __launch_bounds__(128) migrate_bulk_kernel(const long long* __restrict__ src_ptrs,
                                                           const long long* __restrict__ dst_ptrs,
                                                           const long long* __restrict__ snap_ptrs, int n_blocks,
                                                           long long block_bytes, int* __restrict__ err) {
  extern __shared__ __align__(128) unsigned char mig_smem[];
  __shared__ __align__(8) unsigned long long mig_bar[MIG_STAGES];
  if (threadIdx.x != 0) return;
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(mig_smem);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(mig_bar);
  for (int s = 0; s < MIG_STAGES; ++s) mig_mbar_init(bar0 + 8 * s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const long long chunks_per_block = (block_bytes + MIG_CHUNK - 1) / MIG_CHUNK;
  const long long n_items = chunks_per_block * n_blocks;
  // software pipeline over this CTA's items: issue load(i + STAGES - 1) before storing item i
  long long issued = 0, done = 0;
  long long next_item = blockIdx.x;
  long long item_of[MIG_STAGES];
  auto item_geom = [&](long long it, const char*& src, char*& dst, char*& snap, uint32_t& bytes) {
    const long long b = it / chunks_per_block, c = it - b * chunks_per_block;
    const long long off = c * MIG_CHUNK;
    bytes = (uint32_t)min((long long)MIG_CHUNK, block_bytes - off);
    src = reinterpret_cast<const char*>(src_ptrs[b]) + off;
    dst = reinterpret_cast<char*>(dst_ptrs[b]) + off;
    snap = snap_ptrs ? reinterpret_cast<char*>(snap_ptrs[b]) + off : nullptr;
  };
  while (true) {
    // fill the ring
    while (next_item < n_items && issued - done < MIG_STAGES) {
      const int s = (int)(issued % MIG_STAGES);
      if (issued >= MIG_STAGES) {
        // the stage's previous stores must have finished READING shared memory before it is overwritten
        asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(MIG_STAGES - 1) : "memory");
      }
      const char* src; char* dst; char* snap; uint32_t bytes;
      item_geom(next_item, src, dst, snap, bytes);
      mig_mbar_expect_tx(bar0 + 8 * s, bytes);
      bulk_g2s(smem0 + s * MIG_CHUNK, src, bytes, bar0 + 8 * s);
      item_of[s] = next_item;
      ++issued;
      next_item += gridDim.x;
    }
    if (done == issued) break;
    const int s = (int)(done % MIG_STAGES);
    if (!mig_mbar_wait(bar0 + 8 * s, (uint32_t)((done / MIG_STAGES) & 1))) { atomicExch(err, 4); return; }
    const char* src; char* dst; char* snap; uint32_t bytes;
    item_geom(item_of[s], src, dst, snap, bytes);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    bulk_s2g(dst, smem0 + s * MIG_CHUNK, bytes);
    if (snap) bulk_s2g(snap, smem0 + s * MIG_CHUNK, bytes);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    ++done;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete (not just read) before exit
  __threadfence_system();
}

// new += old - snapshot for every element that changed after it was copied.  IS_INT selects s32 arithmetic.
// CLAMP (float tables with a bounded update function, e.g. NMF's SGD + clamp): the forwarded delta is applied with
// the returning atomic and the result is projected back into [lo, hi], like the push kernel does.
template <bool IS_INT>
__global__ void __launch_bounds__(256) migrate_drain_kernel(const long long* __restrict__ old_ptrs,
                                                            const long long* __restrict__ snap_ptrs,
                                                            const long long* __restrict__ new_ptrs, int n_blocks,
                                                            long long vec_per_block,
                                                            unsigned long long* __restrict__ forwarded,
                                                            int do_clamp, float lo, float hi) {
  const long long total = vec_per_block * n_blocks;
  unsigned long long mine = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / vec_per_block, v = i - b * vec_per_block;
    const int4 o = ld_stream_i4(reinterpret_cast<const int*>(old_ptrs[b]) + 4 * v);
    const int4 s = ld_stream_i4(reinterpret_cast<const int*>(snap_ptrs[b]) + 4 * v);
    if (o.x == s.x && o.y == s.y && o.z == s.z && o.w == s.w) continue;
    const int oo[4] = {o.x, o.y, o.z, o.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (oo[e] == ss[e]) continue;
      ++mine;
      if (IS_INT) {
        int* p = reinterpret_cast<int*>(new_ptrs[b]) + 4 * v + e;
        asm volatile("red.relaxed.sys.global.add.s32 [%0], %1;" :: "l"(p), "r"(oo[e] - ss[e]) : "memory");
      } else {
        float* p = reinterpret_cast<float*>(new_ptrs[b]) + 4 * v + e;
        const float d = __int_as_float(oo[e]) - __int_as_float(ss[e]);
        if (do_clamp) {
            const float nv = atom_add_f32_sys(p, d) + d;
            if (nv < lo || nv > hi) clamp_fixup(p, lo, hi);
        } else {
            red_add_f32_sys(p, d);
        }
      }
    }
  }
  if (mine) atomicAdd(forwarded, mine);
  __threadfence_system();
}

// Ownership flip of a LIST of blocks in THIS GPU's pointer table (one launch, release semantics).
__global__ void flip_blocks_kernel(long long* __restrict__ ptr_table, const int* __restrict__ blocks,
                                   const long long* __restrict__ new_ptrs, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    asm volatile("st.release.sys.global.b64 [%0], %1;" :: "l"(ptr_table + blocks[i]), "l"(new_ptrs[i]) : "memory");
  }
}

// Owner-side decay of a list of local slabs in one launch (dense L2 term; the lazy scale avoids it on BSP paths).
__global__ void __launch_bounds__(256) scale_blocks_kernel(const long long* __restrict__ ptrs, int n_blocks,
                                                           long long vec_per_block, float factor) {
  const long long total = vec_per_block * n_blocks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / vec_per_block, v = i - b * vec_per_block;
    float4* p = reinterpret_cast<float4*>(ptrs[b]) + v;
    float4 x = *p;
    x.x *= factor; x.y *= factor; x.z *= factor; x.w *= factor;
    *p = x;
  }
}

// ASSIGN for int32 tables (checkpoint restore of the LDA model): bit-exact copy, no float conversion.
__global__ void __launch_bounds__(256) put_rows_i32_kernel(TableView t, const long long* __restrict__ keys, int n,
                                                           const int* __restrict__ values, int v_stride,
                                                           int bitmap_off) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_cta) {
    int* dst = row_ptr<int>(t, keys[i]);
    const int* src = values + (size_t)i * v_stride;
    for (int e0 = 0; e0 < t.row_dim; e0 += 32) {
      const int e = e0 + lane;
      const int v = (e < t.row_dim) ? src[e] : 0;
      if (e < t.row_dim) dst[e] = v;
      if (bitmap_off > 0) {                       // LDA rows: word e0 / 32 of the topic bitmap (last column = nnz counter)
        const unsigned m = __ballot_sync(0xffffffffu, e < t.row_dim - 1 && v > 0);
        if (lane == 0 && (e0 >> 5) < 32) reinterpret_cast<unsigned*>(dst + bitmap_off)[e0 >> 5] = m;
      }
    }
  }
}

// Ownership flip: write the new base pointer of `block` into every GPU's pointer table
// (ptr_tables[g] is the VA of GPU g's table as mapped here) and publish with a system fence.
__global__ void flip_owner_kernel(const long long* __restrict__ ptr_tables, int n_tables, int block,
                                  const long long* __restrict__ new_ptrs) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_tables) {
    long long* tbl = reinterpret_cast<long long*>(ptr_tables[g]);
    tbl[block] = new_ptrs[g];
    __threadfence_system();
  }
}

// ------------------------------------------------------------------------------------------
// Device clocks: SSP / barrier without host round trips.  clock_ptrs[g] = VA of GPU g's
// int[world] clock array.  publish: write my clock into slot `rank` of every GPU.
// wait: spin (bounded) until every slot of MY array >= min_clock.
// ------------------------------------------------------------------------------------------
__global__ void clock_publish_kernel(const long long* __restrict__ clock_ptrs, int world, int rank, int value) {
  int g = threadIdx.x;
  if (g < world) {
    __threadfence_system();  // all prior pushes of this GPU are visible before the clock moves
    int* arr = reinterpret_cast<int*>(clock_ptrs[g]);
    st_release_sys_s32(arr + rank, value);
  }
}

__global__ void clock_wait_kernel(const int* __restrict__ my_clocks, int world, int min_value,
                                  long long timeout_cycles, int* __restrict__ err_flag) {
  int g = threadIdx.x;
  if (g < world) {
    long long t0 = clock64();
    while (ld_acquire_sys_s32(my_clocks + g) < min_value) {
      if (clock64() - t0 > timeout_cycles) { atomicExch(err_flag, 1 + g); break; }
      __nanosleep(64);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host wrappers
// ------------------------------------------------------------------------------------------
static TableView make_view(const at::Tensor& block_ptrs, int64_t rows_per_block, int64_t row_stride,
                           int64_t row_dim, int64_t mode) {
  TORCH_CHECK(block_ptrs.is_cuda() && block_ptrs.scalar_type() == at::kLong, "block_ptrs must be cuda int64");
  TableView v;
  v.block_ptrs = reinterpret_cast<const long long*>(block_ptrs.data_ptr<int64_t>());
  v.num_blocks = (int)block_ptrs.numel() - 2;   // two guard words follow the pointers
  v.rows_per_block = (int)rows_per_block;
  v.row_stride = (int)row_stride;
  v.row_dim = (int)row_dim;
  v.mode = (int)mode;
  return v;
}

static int grid_for_rows(int64_t n, int warps_per_cta) {
  int64_t g = (n + warps_per_cta - 1) / warps_per_cta;
  int64_t cap = hb::sm_count() * 16;
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, cap));
}

void table_pull(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim, int64_t mode,
                at::Tensor keys, at::Tensor out) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView v = make_view(block_ptrs, rows_per_block, row_stride, row_dim, mode);
  int64_t n = keys.numel();
  if (n == 0) return;
  TORCH_CHECK(keys.scalar_type() == at::kLong && keys.is_cuda());
  TORCH_CHECK(out.is_cuda() && out.dim() == 2 && out.size(0) == n && out.stride(1) == 1);
  TORCH_CHECK(out.size(1) <= row_stride, "pull: output rows are wider than the table rows");
  if (out.scalar_type() == at::kFloat) {
    pull_rows_kernel<<<grid_for_rows(n, 8), 256, 0, stream>>>(
        v, reinterpret_cast<const long long*>(keys.data_ptr<int64_t>()), (int)n, out.data_ptr<float>(),
        (int)out.stride(0), (int)out.size(1),
        (int)((out.stride(0) % 4 == 0) && (reinterpret_cast<uintptr_t>(out.data_ptr<float>()) % 16 == 0)));
  } else {
    TORCH_CHECK(out.scalar_type() == at::kInt);
    pull_rows_i32_kernel<<<grid_for_rows(n, 8), 256, 0, stream>>>(
        v, reinterpret_cast<const long long*>(keys.data_ptr<int64_t>()), (int)n, out.data_ptr<int>(),
        (int)out.stride(0), (int)out.size(1));
  }
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void table_push(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim, int64_t mode,
                at::Tensor keys, at::Tensor deltas, int64_t fn, double step, double lo, double hi, double eps) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView v = make_view(block_ptrs, rows_per_block, row_stride, row_dim, mode);
  int64_t n = keys.numel();
  if (n == 0) return;
  TORCH_CHECK(keys.scalar_type() == at::kLong && keys.is_cuda());
  TORCH_CHECK(deltas.is_cuda() && deltas.scalar_type() == at::kFloat && deltas.dim() == 2 && deltas.size(0) == n);
  TORCH_CHECK(deltas.stride(1) == 1, "deltas rows must be contiguous");
  const int vec_ok = (deltas.stride(0) % 4 == 0) && (reinterpret_cast<uintptr_t>(deltas.data_ptr<float>()) % 16 == 0);
  UpdateParams up{(float)step, (float)lo, (float)hi, (float)eps};
  push_rows_kernel<<<grid_for_rows(n, 8), 256, 0, stream>>>(
      v, reinterpret_cast<const long long*>(keys.data_ptr<int64_t>()), (int)n, deltas.data_ptr<float>(),
      (int)deltas.stride(0), (int)fn, up, vec_ok);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void table_push_sparse_i32(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim,
                           int64_t mode, at::Tensor keys, at::Tensor topics, at::Tensor deltas, int64_t K,
                           int64_t bitmap_off) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView v = make_view(block_ptrs, rows_per_block, row_stride, row_dim, mode);
  int64_t n = keys.numel();
  if (n == 0) return;
  TORCH_CHECK(keys.scalar_type() == at::kLong && topics.scalar_type() == at::kInt && deltas.scalar_type() == at::kInt);
  push_sparse_i32_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(
      v, reinterpret_cast<const long long*>(keys.data_ptr<int64_t>()), topics.data_ptr<int>(),
      deltas.data_ptr<int>(), (int)n, (int)K, (int)bitmap_off);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void table_pull_lda_sparse(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim,
                           int64_t mode, at::Tensor keys, at::Tensor out, int64_t K, int64_t bitmap_off,
                           c10::optional<at::Tensor> chunk_count) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView v = make_view(block_ptrs, rows_per_block, row_stride, row_dim, mode);
  int64_t n = keys.numel();
  if (n == 0) return;
  TORCH_CHECK(keys.scalar_type() == at::kLong && out.scalar_type() == at::kInt && out.dim() == 2 && out.size(0) >= n);
  TORCH_CHECK(out.stride(1) == 1 && out.stride(0) % 4 == 0 && out.stride(0) <= row_stride && out.stride(0) > K &&
              reinterpret_cast<uintptr_t>(out.data_ptr<int>()) % 16 == 0, "sparse pull: bad staging layout");
  TORCH_CHECK(bitmap_off > 0 && bitmap_off + 32 <= row_stride && K <= 1024);
  pull_rows_lda_sparse_kernel<<<grid_for_rows(n, 8), 256, 0, stream>>>(
      v, reinterpret_cast<const long long*>(keys.data_ptr<int64_t>()), (int)n, out.data_ptr<int>(), (int)out.stride(0),
      (int)K, (int)bitmap_off,
      chunk_count.has_value() ? reinterpret_cast<unsigned long long*>(chunk_count->data_ptr<int64_t>()) : nullptr);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void lda_rebuild_bitmaps(at::Tensor slab_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t K, int64_t bitmap_off) {
  c10::cuda::CUDAGuard guard(slab_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int nb = (int)slab_ptrs.numel();
  if (nb == 0) return;
  const long long rows = (long long)nb * rows_per_block;
  lda_rebuild_bitmaps_kernel<<<(int)std::min<long long>((rows + 7) / 8, 148 * 16), 256, 0, stream>>>(
      reinterpret_cast<const long long*>(slab_ptrs.data_ptr<int64_t>()), nb, (int)rows_per_block, (int)row_stride, (int)K,
      (int)bitmap_off);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void table_push_scalar_i32(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim,
                           int64_t mode, at::Tensor keys, at::Tensor values, int64_t op) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView v = make_view(block_ptrs, rows_per_block, row_stride, row_dim, mode);
  int64_t n = keys.numel();
  if (n == 0) return;
  TORCH_CHECK(keys.scalar_type() == at::kLong && values.scalar_type() == at::kInt && values.numel() == n);
  push_scalar_i32_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(
      v, reinterpret_cast<const long long*>(keys.data_ptr<int64_t>()), values.data_ptr<int>(), (int)n, (int)op);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// Raw-pointer entry points launch on the GPU that OWNS the memory, not on whatever device the calling thread
// happens to be on (a peer GPU could write it over NVLink, but un-ordered with the owner's streams).
static c10::DeviceIndex device_of_ptr(int64_t ptr) {
  cudaPointerAttributes attr;
  C10_CUDA_CHECK(cudaPointerGetAttributes(&attr, reinterpret_cast<const void*>(ptr)));
  TORCH_CHECK(attr.type == cudaMemoryTypeDevice, "expected a device pointer");
  return (c10::DeviceIndex)attr.device;
}

void init_slab(int64_t slab_ptr, int64_t rows, int64_t row_stride, int64_t row_dim, int64_t kind, double a,
               double b, int64_t seed, int64_t row_key_base, int64_t key_stride) {
  c10::cuda::CUDAGuard guard(device_of_ptr(slab_ptr));
  auto stream = at::cuda::getCurrentCUDAStream();
  long long total = rows * row_stride;
  if (total == 0) return;
  int grid = (int)std::min<long long>((total + 255) / 256, 148 * 32);
  init_slab_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<float*>(slab_ptr), rows, (int)row_stride,
                                             (int)row_dim, (int)kind, (float)a, (float)b,
                                             (unsigned long long)seed, row_key_base, key_stride);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void migrate_copy(int64_t src_ptr, int64_t dst_ptr, int64_t nbytes) {
  c10::cuda::CUDAGuard guard(device_of_ptr(dst_ptr));      // the receiver pulls the block over NVLink
  auto stream = at::cuda::getCurrentCUDAStream();
  TORCH_CHECK(nbytes % 16 == 0 && src_ptr % 16 == 0 && dst_ptr % 16 == 0, "migration needs 16B alignment");
  long long n_vec = nbytes / 16;
  if (n_vec == 0) return;
  int grid = (int)std::min<long long>((n_vec + 2047) / 2048, 148 * 4);
  migrate_copy_kernel<<<grid, 512, 0, stream>>>(reinterpret_cast<const int4*>(src_ptr),
                                                reinterpret_cast<int4*>(dst_ptr), n_vec);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void migrate_bulk(at::Tensor src_ptrs, at::Tensor dst_ptrs, c10::optional<at::Tensor> snap_ptrs, int64_t block_bytes,
                  at::Tensor err) {
  c10::cuda::CUDAGuard guard(src_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int n = (int)src_ptrs.numel();
  if (n == 0) return;
  TORCH_CHECK(block_bytes % 16 == 0, "bulk copies move multiples of 16 bytes");
  TORCH_CHECK(dst_ptrs.numel() == n && (!snap_ptrs.has_value() || snap_ptrs->numel() == n));
  const int smem = MIG_STAGES * MIG_CHUNK;
  // per device (executors of one process drive different GPUs), so set it on every call: it is a cheap host call
  C10_CUDA_CHECK(cudaFuncSetAttribute(migrate_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long items = ((block_bytes + MIG_CHUNK - 1) / MIG_CHUNK) * n;
  const int grid = (int)std::min<long long>(items, 148);
  migrate_bulk_kernel<<<grid, 128, smem, stream>>>(
      reinterpret_cast<const long long*>(src_ptrs.data_ptr<int64_t>()),
      reinterpret_cast<const long long*>(dst_ptrs.data_ptr<int64_t>()),
      snap_ptrs.has_value() ? reinterpret_cast<const long long*>(snap_ptrs->data_ptr<int64_t>()) : nullptr, n,
      (long long)block_bytes, err.data_ptr<int>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void migrate_drain(at::Tensor old_ptrs, at::Tensor snap_ptrs, at::Tensor new_ptrs, int64_t block_bytes, bool is_int,
                   at::Tensor forwarded, bool do_clamp, double lo, double hi) {
  c10::cuda::CUDAGuard guard(old_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int n = (int)old_ptrs.numel();
  if (n == 0) return;
  TORCH_CHECK(forwarded.scalar_type() == at::kLong && block_bytes % 16 == 0);
  const long long vpb = block_bytes / 16;
  const int grid = (int)std::min<long long>((vpb * n + 255) / 256, 148 * 16);
  auto* fw = reinterpret_cast<unsigned long long*>(forwarded.data_ptr<int64_t>());
  auto* o = reinterpret_cast<const long long*>(old_ptrs.data_ptr<int64_t>());
  auto* sn = reinterpret_cast<const long long*>(snap_ptrs.data_ptr<int64_t>());
  auto* nw = reinterpret_cast<const long long*>(new_ptrs.data_ptr<int64_t>());
  if (is_int) migrate_drain_kernel<true><<<grid, 256, 0, stream>>>(o, sn, nw, n, vpb, fw, 0, 0.f, 0.f);
  else migrate_drain_kernel<false><<<grid, 256, 0, stream>>>(o, sn, nw, n, vpb, fw, do_clamp ? 1 : 0, (float)lo, (float)hi);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void flip_blocks(at::Tensor ptr_table, at::Tensor blocks, at::Tensor new_ptrs) {
  c10::cuda::CUDAGuard guard(ptr_table.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int n = (int)blocks.numel();
  if (n == 0) return;
  TORCH_CHECK(blocks.scalar_type() == at::kInt && new_ptrs.scalar_type() == at::kLong && new_ptrs.numel() == n);
  flip_blocks_kernel<<<(n + 127) / 128, 128, 0, stream>>>(
      reinterpret_cast<long long*>(ptr_table.data_ptr<int64_t>()), blocks.data_ptr<int>(),
      reinterpret_cast<const long long*>(new_ptrs.data_ptr<int64_t>()), n);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void scale_blocks(at::Tensor ptrs, int64_t block_bytes, double factor) {
  c10::cuda::CUDAGuard guard(ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int n = (int)ptrs.numel();
  if (n == 0) return;
  const long long vpb = block_bytes / 16;
  const int grid = (int)std::min<long long>((vpb * n + 255) / 256, 148 * 16);
  scale_blocks_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const long long*>(ptrs.data_ptr<int64_t>()), n, vpb,
                                                (float)factor);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void table_put_i32(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim, int64_t mode,
                   at::Tensor keys, at::Tensor values, int64_t bitmap_off) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView v = make_view(block_ptrs, rows_per_block, row_stride, row_dim, mode);
  int64_t n = keys.numel();
  if (n == 0) return;
  TORCH_CHECK(keys.scalar_type() == at::kLong && values.scalar_type() == at::kInt && values.dim() == 2 &&
              values.size(0) == n && values.stride(1) == 1 && values.size(1) >= row_dim);
  put_rows_i32_kernel<<<grid_for_rows(n, 8), 256, 0, stream>>>(
      v, reinterpret_cast<const long long*>(keys.data_ptr<int64_t>()), (int)n, values.data_ptr<int>(),
      (int)values.stride(0), (int)bitmap_off);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void flip_owner(at::Tensor ptr_tables, int64_t block, at::Tensor new_ptrs) {
  c10::cuda::CUDAGuard guard(ptr_tables.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  int n = (int)ptr_tables.numel();
  flip_owner_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const long long*>(ptr_tables.data_ptr<int64_t>()), n,
                                          (int)block,
                                          reinterpret_cast<const long long*>(new_ptrs.data_ptr<int64_t>()));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void clock_publish(at::Tensor clock_ptrs, int64_t rank, int64_t value) {
  c10::cuda::CUDAGuard guard(clock_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  int world = (int)clock_ptrs.numel();
  clock_publish_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const long long*>(clock_ptrs.data_ptr<int64_t>()),
                                             world, (int)rank, (int)value);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void clock_wait(at::Tensor my_clocks, int64_t min_value, int64_t timeout_cycles, at::Tensor err_flag) {
  c10::cuda::CUDAGuard guard(my_clocks.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  int world = (int)my_clocks.numel();
  clock_wait_kernel<<<1, 32, 0, stream>>>(my_clocks.data_ptr<int>(), world, (int)min_value,
                                          (long long)timeout_cycles, err_flag.data_ptr<int>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace hb
