// Dense-input MLR forward on the 5th-gen tensor cores: logits[B, C] = X[B, F] . W[F, C].
//
// This is the one GEMM-shaped op of the framework (dense instances; the sparse 1M-feature path is
// bandwidth-bound and lives in mlr_ops.cu).  The PULL of the model is fused into the GEMM: the B
// operand tiles are TMA-loaded (cp.async.bulk.tensor) straight out of the OWNER's block slabs --
// local HBM or an NVLink peer mapping -- through one tensor map per table block, so the model is
// never staged or converted: the fp32 master weights are consumed as tf32 (kind::tf32).
//
//   warp 0      TMA producer  (A = X tile [128 x 32] K-major, B = W tile [32 x 128] MN-major, SW128)
//   warp 1      TMEM alloc + single-thread tcgen05.mma issue, tcgen05.commit -> mbarriers
//   warps 2..5  epilogue: tcgen05.ld accumulators (TMEM) -> registers -> fp32 logits in global
// 4-stage smem ring (32 KB / stage), accumulator 128 x 128 fp32 = 128 TMEM columns.
// Reference op: MLRTrainer.predict (dolphin/mlapps/mlr/MLRTrainer.java:475-481, C sparse dots per
// instance through Breeze) -- here one tensor-core GEMM per mini-batch.
#include "hb_common.cuh"
#include <cuda.h>
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

namespace hb {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 4, TC_UMMA_K = 8;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;          // 16 KB
constexpr int TC_B_BYTES = TC_BK * TC_BN * 4;          // 16 KB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 128;

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > 200000000u) __trap();      // a protocol bug must fault, never hang the GPU
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_dst), "l"(tmap), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor), SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;   // LayoutType: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

// A_MN = false : A tile is K-major  (forward:  A = X[B, F],  M = instances, K = features)
// A_MN = true  : A tile is MN-major (backward: A = X^T,      M = features,  K = instances)
// PUSH = false : epilogue stores fp32 logits
// PUSH = true  : epilogue reduce-adds the gradient tile straight into the OWNERS' table rows
//                (GEMM fused with the PUSH + `w += delta` update, K6-bwd + K2 + K3)
template <bool A_MN, bool PUSH>
__global__ void __launch_bounds__(TC_THREADS, 1) dense_gemm_tf32_kernel(
    const __grid_constant__ CUtensorMap tmA, const CUtensorMap* __restrict__ tmB, float* __restrict__ out,
    int out_stride, int M, int N, int K, int rows_per_block, TableView Wtab) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B needs 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + TC_STAGES * TC_STAGE_BYTES;      // full[4], empty[4], tmem_full, tmem_slot
  const uint32_t full0 = bars, empty0 = bars + 8 * TC_STAGES, tmem_full = bars + 16 * TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TC_STAGES * TC_STAGE_BYTES + 16 * TC_STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
  const int KT = K / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation is a warp-wide instruction; the address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % TC_STAGES;
        const uint32_t par = ((kt / TC_STAGES) & 1) ^ 1;
        mbar_wait(empty0 + 8 * s, par);
        const uint32_t sa = base + s * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
        mbar_expect_tx(full0 + 8 * s, TC_STAGE_BYTES);
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < TC_BM / 32; ++j)      // A^T: 4 column groups of 32 features x BK instances
            tma_load_2d(sa + j * (TC_BK * 128), &tmA, full0 + 8 * s, m0 + 32 * j, kt * TC_BK);
        } else {
          tma_load_2d(sa, &tmA, full0 + 8 * s, kt * TC_BK, m0);
        }
        const int f0 = kt * TC_BK;
        const int blk = f0 / rows_per_block, y = f0 - blk * rows_per_block;
        const CUtensorMap* mb = tmB + blk;          // tensor map of the OWNER's slab (local or peer)
#pragma unroll
        for (int j = 0; j < TC_BN / 32; ++j)        // 4 column groups of 128 B (SW128 inner box limit)
          tma_load_2d(sb + j * (TC_BK * 128), mb, full0 + 8 * s, n0 + 32 * j, y);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer (one thread) =====
      // instruction descriptor (cute InstrDescriptor): D=f32, A=B=tf32, A K-major, B MN-major, N=128, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | (1u << 16) |
                             ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % TC_STAGES;
        mbar_wait(full0 + 8 * s, (kt / TC_STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = base + s * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
#pragma unroll
        for (int i = 0; i < TC_BK / TC_UMMA_K; ++i) {
          // A: K-major SW128, 8-row groups 1024 B apart, K-step = 8 tf32 = 32 B inside the swizzle row
          const uint64_t ad = A_MN ? umma_desc(sa + i * 1024, TC_BK * 128, 512, 1)
                                   : umma_desc(sa + i * (TC_UMMA_K * 4), 16, 1024);
          // B: MN-major SW128, 128 B column groups (BK x 128 B = 4096 B apart), 8 K-rows = 1024 B
          const uint64_t bd = umma_desc(sb + i * 1024, TC_BK * 128, 512, 1);
          tc_mma_tf32(tmem_base, ad, bd, idesc, (kt > 0 || i > 0) ? 1u : 0u);
        }
        tc_commit(empty0 + 8 * s);                  // smem slot is free once these MMAs retire
      }
      tc_commit(tmem_full);                         // accumulator complete
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> global =====
    const int q = warp & 3;                         // TMEM lane quadrant this warp may access
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int row = m0 + 32 * q + lane;
#pragma unroll
    for (int c = 0; c < TC_BN / 32; ++c) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (PUSH) {
        if (row < M) {
          float* o = row_ptr<float>(Wtab, (long long)row) + n0 + c * 32;   // owner's row (local or peer)
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            if (n0 + c * 32 + e < N)
              red_add_f4_sys(o + e, make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])));
        }
      } else if (row < M) {
        float* o = out + (size_t)row * out_stride + n0 + c * 32;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          if (n0 + c * 32 + e < N)
            *reinterpret_cast<float4*>(o + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                            __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
  }
}


// ===================================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  Each CTA
// stages its own 128 rows of A and HALF of the B tile (128 of the 256 columns), the leader issues
// tcgen05.mma.cta_group::2 (M = 256), both CTAs keep 128 x 256 fp32 accumulators in their TMEM.
// Per CTA the staged bytes per K-step are unchanged (32 KB) while the math doubles: 64 FLOP per
// staged byte instead of 32, which lifts the L2-bandwidth ceiling that bounds the 1-CTA kernel.
// ===================================================================================================
constexpr int T2_BN = 256, T2_STAGES = 6, T2_TMEM_COLS = 256;
constexpr int T2_SMEM_BYTES = T2_STAGES * TC_STAGE_BYTES + 1024 + 256;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const void* tmap, uint32_t leader_bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_dst), "l"(tmap), "r"(leader_bar & kPeerBitMask), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool A_MN, bool PUSH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) dense_gemm_tf32_2cta_kernel(
    const __grid_constant__ CUtensorMap tmA, const CUtensorMap* __restrict__ tmB, float* __restrict__ out,
    int out_stride, int M, int N, int K, int rows_per_block, TableView Wtab) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + T2_STAGES * TC_STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * T2_STAGES, tmem_full = bars + 16 * T2_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T2_STAGES * TC_STAGE_BYTES + 16 * T2_STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int m0 = (blockIdx.x >> 1) * 256 + (int)cta_rank * TC_BM;   // this CTA's 128 rows of the 256-row tile
  const int n0 = blockIdx.y * T2_BN;
  const int nb0 = n0 + (int)cta_rank * 128;                          // this CTA's half of the B tile
  const int KT = K / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < T2_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "n"(T2_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs; transaction bytes are credited to the LEADER's barrier) =====
      for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % T2_STAGES;
        mbar_wait(empty0 + 8 * s, ((kt / T2_STAGES) & 1) ^ 1);
        const uint32_t sa = base + s * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
        if (leader) mbar_expect_tx(full0 + 8 * s, 2 * TC_STAGE_BYTES);
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < TC_BM / 32; ++j)
            tma_load_2d_2cta(sa + j * (TC_BK * 128), &tmA, full0 + 8 * s, m0 + 32 * j, kt * TC_BK);
        } else {
          tma_load_2d_2cta(sa, &tmA, full0 + 8 * s, kt * TC_BK, m0);
        }
        const int f0 = kt * TC_BK;
        const int blk = f0 / rows_per_block, y = f0 - blk * rows_per_block;
        const CUtensorMap* mb = tmB + blk;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          tma_load_2d_2cta(sb + j * (TC_BK * 128), mb, full0 + 8 * s, nb0 + 32 * j, y);
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // ===== MMA issuer: leader CTA only, M = 256 across the pair, N = 256 =====
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | (1u << 16) |
                             ((uint32_t)(T2_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % T2_STAGES;
        mbar_wait(full0 + 8 * s, (kt / T2_STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = base + s * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
#pragma unroll
        for (int i = 0; i < TC_BK / TC_UMMA_K; ++i) {
          const uint64_t ad = A_MN ? umma_desc(sa + i * 1024, TC_BK * 128, 512, 1)
                                   : umma_desc(sa + i * (TC_UMMA_K * 4), 16, 1024);
          const uint64_t bd = umma_desc(sb + i * 1024, TC_BK * 128, 512, 1);
          tc_mma_tf32_2cta(tmem_base, ad, bd, idesc, (kt > 0 || i > 0) ? 1u : 0u);
        }
        tc_commit_2cta(empty0 + 8 * s);             // frees the stage in BOTH CTAs
      }
      tc_commit_2cta(tmem_full);                    // accumulators complete in BOTH CTAs
    }
  } else {
    // ===== epilogue (each CTA drains its own 128 x 256 accumulator) =====
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int row = m0 + 32 * q + lane;
#pragma unroll 1
    for (int c = 0; c < T2_BN / 32; ++c) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int col0 = n0 + c * 32;
      if (PUSH) {
        if (row < M) {
          float* o = row_ptr<float>(Wtab, (long long)row) + col0;
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            if (col0 + e < N)
              red_add_f4_sys(o + e, make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])));
        }
      } else if (row < M) {
        float* o = out + (size_t)row * out_stride + col0;
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          if (col0 + e < N)
            *reinterpret_cast<float4*>(o + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                            __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                               // both CTAs are done with TMEM / peer smem
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(T2_TMEM_COLS) : "memory");
  }
}


// ===================================================================================================
// Persistent "data-parallel + stream-K tail" variant of the 2-CTA kernel (the default).  The grid is ONE
// 2-CTA cluster per TPC (74 on a B200) whatever the problem.  Whole waves of 74 tiles are handed out
// strided (cluster c computes tiles c, c + 74, ...: at any moment the machine works on 74 CONSECUTIVE
// tiles, m fastest, which share their B panel in L2).  The last, incomplete wave of R < 74 tiles is cut
// into 74 equal ranges of (tile, k-step) so every SM computes the same number of MMAs (plain tiling
// leaves 10 of 74 clusters idle at 4096 x 1024 = 64 tiles).  A range that crosses a tile border is run
// HEAD FIRST (k = 0.. of the later tile), then the tail of the earlier tile: all clusters then sit at the
// same k (plus one common offset in the tail phase), which keeps the A / B k-slices shared in L2 -- with
// the natural order the clusters are staggered over all k and the kernel becomes HBM bound (measured:
// 407 instead of 640 TFLOP/s).  A partial range contributes with red.add (the output is zeroed first; the
// PUSH epilogue is a red.add anyway).
// TMEM is double buffered (2 x 256 columns = all 512): the epilogue of work unit u drains buffer
// u & 1 while the MMAs of unit u + 1 fill the other; the TMA producer never stops at unit borders.
//   barriers (per CTA): full[6] / empty[6] smem stages; tmem_full[2] (multicast commit);
//                       tmem_empty[2] on the LEADER, 8 arrivals = 4 epilogue warps x 2 CTAs.
// ===================================================================================================
constexpr int SK_TMEM_COLS = 512;
constexpr int SK_SMEM_BYTES = T2_STAGES * TC_STAGE_BYTES + 1024 + 256;

struct SkPlan {            // identical in every role of the cluster (producer / MMA / epilogue)
  int nc, cluster, KT, waves, n_units;
  int h_tile, h_kb, h_ke;  // remainder segment run first  (head of the later tile, or the only segment)
  int t_tile, t_kb, t_ke;  // remainder segment run second (tail of the earlier tile); t_tile < 0: none
};
__device__ __forceinline__ SkPlan sk_plan(int tiles, int KT, int nc, int cluster) {
  SkPlan p;
  p.nc = nc; p.cluster = cluster; p.KT = KT;
  p.waves = tiles / nc;
  const int R = tiles - p.waves * nc;
  p.h_tile = p.t_tile = -1;
  p.h_kb = p.h_ke = p.t_kb = p.t_ke = 0;
  int extra = 0;
  if (R > 0) {
    const long long rit = (long long)R * KT;
    const long long rb = rit * cluster / nc, re = rit * (cluster + 1) / nc;
    if (re > rb) {
      const int t0 = (int)(rb / KT), k0 = (int)(rb - (long long)t0 * KT);
      const int len = (int)(re - rb);
      if (k0 + len <= KT) {
        p.h_tile = p.waves * nc + t0; p.h_kb = k0; p.h_ke = k0 + len; extra = 1;
      } else {
        p.t_tile = p.waves * nc + t0; p.t_kb = k0; p.t_ke = KT;
        p.h_tile = p.waves * nc + t0 + 1; p.h_kb = 0; p.h_ke = k0 + len - KT; extra = 2;
      }
    }
  }
  p.n_units = p.waves + extra;
  return p;
}
__device__ __forceinline__ void sk_unit(const SkPlan& p, int u, int& tile, int& kb, int& ke) {
  if (u < p.waves) { tile = u * p.nc + p.cluster; kb = 0; ke = p.KT; }
  else if (u == p.waves) { tile = p.h_tile; kb = p.h_kb; ke = p.h_ke; }
  else { tile = p.t_tile; kb = p.t_kb; ke = p.t_ke; }
}

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t target_cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(target_cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
}

template <bool A_MN, bool PUSH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) dense_gemm_tf32_sk_kernel(
    const __grid_constant__ CUtensorMap tmA, const CUtensorMap* __restrict__ tmB, float* __restrict__ out,
    int out_stride, int M, int N, int K, int rows_per_block, TableView Wtab) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + T2_STAGES * TC_STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * T2_STAGES;
  const uint32_t tfull0 = bars + 16 * T2_STAGES, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T2_STAGES * TC_STAGE_BYTES + 16 * T2_STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int KT = K / TC_BK;
  const int MT = (M + 255) / 256, NT = (N + T2_BN - 1) / T2_BN;
  const SkPlan plan = sk_plan(MT * NT, KT, (int)(gridDim.x >> 1), (int)(blockIdx.x >> 1));

  if (threadIdx.x == 0) {
    for (int s = 0; s < T2_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "n"(SK_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: one continuous stream of k-steps across the work units =====
      long long g = 0;                                   // global stage counter of this CTA
      for (int u = 0; u < plan.n_units; ++u) {
        int tile, kb, ke;
        sk_unit(plan, u, tile, kb, ke);
        const int nt = tile / MT, mt = tile - nt * MT;   // m fastest: concurrent tiles share their B panel
        const int m0 = mt * 256 + (int)cta_rank * TC_BM, nb0 = nt * T2_BN + (int)cta_rank * 128;
        for (int kt = kb; kt < ke; ++kt, ++g) {
          const int s = (int)(g % T2_STAGES);
          mbar_wait(empty0 + 8 * s, (uint32_t)(((g / T2_STAGES) & 1) ^ 1));
          const uint32_t sa = base + s * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
          if (leader) mbar_expect_tx(full0 + 8 * s, 2 * TC_STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < TC_BM / 32; ++j)
              tma_load_2d_2cta(sa + j * (TC_BK * 128), &tmA, full0 + 8 * s, m0 + 32 * j, kt * TC_BK);
          } else {
            tma_load_2d_2cta(sa, &tmA, full0 + 8 * s, kt * TC_BK, m0);
          }
          const int f0 = kt * TC_BK;
          const int blk = f0 / rows_per_block, y = f0 - blk * rows_per_block;
          const CUtensorMap* mb = tmB + blk;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tma_load_2d_2cta(sb + j * (TC_BK * 128), mb, full0 + 8 * s, nb0 + 32 * j, y);
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // ===== MMA issuer (leader only): unit u accumulates into TMEM buffer u & 1 =====
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | (1u << 16) |
                             ((uint32_t)(T2_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      long long g = 0;
      for (int u = 0; u < plan.n_units; ++u) {
        int tile, kb, ke;
        sk_unit(plan, u, tile, kb, ke);
        const int a = u & 1;
        mbar_wait(tempty0 + 8 * a, (uint32_t)(((u >> 1) & 1) ^ 1));     // both CTAs drained this buffer
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(a * T2_BN);
        for (int kt = kb; kt < ke; ++kt, ++g) {
          const int s = (int)(g % T2_STAGES);
          mbar_wait(full0 + 8 * s, (uint32_t)((g / T2_STAGES) & 1));
          tc_fence_after();
          const uint32_t sa = base + s * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
#pragma unroll
          for (int i = 0; i < TC_BK / TC_UMMA_K; ++i) {
            const uint64_t ad = A_MN ? umma_desc(sa + i * 1024, TC_BK * 128, 512, 1)
                                     : umma_desc(sa + i * (TC_UMMA_K * 4), 16, 1024);
            const uint64_t bd = umma_desc(sb + i * 1024, TC_BK * 128, 512, 1);
            tc_mma_tf32_2cta(acc, ad, bd, idesc, (kt > kb || i > 0) ? 1u : 0u);
          }
          tc_commit_2cta(empty0 + 8 * s);
        }
        tc_commit_2cta(tfull0 + 8 * a);                    // this unit's accumulators are complete in both CTAs
      }
    }
  } else {
    // ===== epilogue warps (both CTAs): drain buffer u & 1, then hand it back to the leader's MMA warp =====
    const int q = warp & 3;
    for (int u = 0; u < plan.n_units; ++u) {
      int tile, kb, ke;
      sk_unit(plan, u, tile, kb, ke);
      const int nt = tile / MT, mt = tile - nt * MT;
      const int m0 = mt * 256 + (int)cta_rank * TC_BM, n0 = nt * T2_BN;
      const bool partial = PUSH || kb > 0 || ke < KT;      // another cluster owns the rest of this tile's K
      const int a = u & 1;
      mbar_wait(tfull0 + 8 * a, (uint32_t)((u >> 1) & 1));
      tc_fence_after();
      const int row = m0 + 32 * q + lane;
#pragma unroll 1
      for (int c = 0; c < T2_BN / 32; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(a * T2_BN + c * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int col0 = n0 + c * 32;
        if (row < M) {
          float* o = PUSH ? row_ptr<float>(Wtab, (long long)row) + col0 : out + (size_t)row * out_stride + col0;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            if (col0 + e < N) {
              const float4 f = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                           __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
              if (partial) red_add_f4_sys(o + e, f);
              else *reinterpret_cast<float4*>(o + e) = f;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * a, 0);   // 8 arrivals (4 warps x 2 CTAs) free the buffer
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(SK_TMEM_COLS) : "memory");
  }
}

// Launch helper: one cluster per TPC; `zero_me` (forward) is zeroed when some tile is split across clusters.
template <bool A_MN, bool PUSH>
static void launch_stream_k(const CUtensorMap& tmA, const CUtensorMap* tmB, float* out, int out_stride, int M, int N,
                            int K, int rows_per_block, const TableView& W, at::Tensor* zero_me, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int MT = (M + 255) / 256, NT = (N + T2_BN - 1) / T2_BN, KT = K / TC_BK;
  const long long tiles = (long long)MT * NT;
  int n_clusters = (int)std::min<long long>(sms / 2, tiles * KT);
  if (tiles % n_clusters != 0 && zero_me != nullptr)     // the last wave is cut by k: its writers all red.add
    C10_CUDA_CHECK(cudaMemsetAsync(zero_me->data_ptr(), 0, (size_t)zero_me->size(0) * zero_me->stride(0) * 4, stream));
  cudaFuncSetAttribute(dense_gemm_tf32_sk_kernel<A_MN, PUSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES);
  dense_gemm_tf32_sk_kernel<A_MN, PUSH><<<2 * n_clusters, TC_THREADS, SK_SMEM_BYTES, stream>>>(
      tmA, tmB, out, out_stride, M, N, K, rows_per_block, W);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// ---- host: tensor maps ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    TORCH_CHECK(e == cudaSuccess && p != nullptr, "cuTensorMapEncodeTiled is not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static void encode_2d_f32(CUtensorMap* tm, void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                          uint32_t box_inner, uint32_t box_outer,
                          CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_elems * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code ", (int)r);
}

// One tensor map per table block: [rows_per_block, row_stride] fp32 at the block's (local or peer) VA.
at::Tensor make_block_tensormaps(at::Tensor block_ptrs_cpu, int64_t rows_per_block, int64_t row_stride,
                                 int64_t device, int64_t swizzle_mode) {
  TORCH_CHECK(!block_ptrs_cpu.is_cuda() && block_ptrs_cpu.scalar_type() == at::kLong);
  const int64_t nb = block_ptrs_cpu.numel();
  std::vector<CUtensorMap> maps(nb);
  const int64_t* p = block_ptrs_cpu.data_ptr<int64_t>();
  for (int64_t b = 0; b < nb; ++b)
    encode_2d_f32(&maps[b], reinterpret_cast<void*>(p[b]), (uint64_t)row_stride, (uint64_t)rows_per_block,
                  (uint64_t)row_stride, 32, TC_BK,
                  swizzle_mode == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
  auto out = at::empty({nb, (int64_t)sizeof(CUtensorMap)},
                       at::TensorOptions().dtype(at::kByte).device(at::kCUDA, (int)device));
  cudaMemcpy(out.data_ptr(), maps.data(), nb * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  return out;
}

// logits[M, >=N] = X[M, K] . W[K, N]   (W rows live in the table blocks described by `bmaps`)
void dense_logits_tf32(at::Tensor X, at::Tensor bmaps, int64_t rows_per_block, int64_t n_cols, at::Tensor out,
                       bool use_2cta, bool stream_k) {
  c10::cuda::CUDAGuard guard(X.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TORCH_CHECK(X.is_cuda() && X.scalar_type() == at::kFloat && X.dim() == 2 && X.stride(1) == 1);
  TORCH_CHECK(out.is_cuda() && out.scalar_type() == at::kFloat && out.stride(1) == 1 && out.stride(0) % 4 == 0);
  const int M = (int)X.size(0), K = (int)X.size(1), N = (int)n_cols;
  TORCH_CHECK(K % TC_BK == 0 && rows_per_block % TC_BK == 0, "features and rows_per_block must be multiples of 32");
  TORCH_CHECK(X.stride(0) % 4 == 0 && (reinterpret_cast<uintptr_t>(X.data_ptr<float>()) % 16 == 0));
  TORCH_CHECK(out.size(0) >= M && out.size(1) >= N);
  CUtensorMap tmA;
  encode_2d_f32(&tmA, X.data_ptr<float>(), (uint64_t)K, (uint64_t)M, (uint64_t)X.stride(0), TC_BK, TC_BM);
  // per-device attribute (several executors of one process may drive different GPUs): set it every launch
  cudaFuncSetAttribute(dense_gemm_tf32_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  TableView none{};
  if (use_2cta && stream_k) {
    launch_stream_k<false, false>(tmA, reinterpret_cast<const CUtensorMap*>(bmaps.data_ptr()), out.data_ptr<float>(),
                                  (int)out.stride(0), M, N, K, (int)rows_per_block, none, &out, stream);
    return;
  }
  if (use_2cta) {
    cudaFuncSetAttribute(dense_gemm_tf32_2cta_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES);
    dim3 grid2(2 * ((M + 255) / 256), (N + T2_BN - 1) / T2_BN);
    dense_gemm_tf32_2cta_kernel<false, false><<<grid2, TC_THREADS, T2_SMEM_BYTES, stream>>>(
        tmA, reinterpret_cast<const CUtensorMap*>(bmaps.data_ptr()), out.data_ptr<float>(), (int)out.stride(0), M, N,
        K, (int)rows_per_block, none);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return;
  }
  dim3 grid((N + TC_BN - 1) / TC_BN, (M + TC_BM - 1) / TC_BM);
  dense_gemm_tf32_kernel<false, false><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(
      tmA, reinterpret_cast<const CUtensorMap*>(bmaps.data_ptr()), out.data_ptr<float>(), (int)out.stride(0), M, N,
      K, (int)rows_per_block, none);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// W[f, :] += sum_b X[b, f] * D[b, :]   (D is pre-scaled by -step).  GEMM M = features, N = classes, K = batch;
// the epilogue reduce-adds each 128 x 128 gradient tile into the owners' rows: no gradient buffer exists.
void dense_grad_push_tf32(at::Tensor X, at::Tensor D, at::Tensor block_ptrs, int64_t rows_per_block,
                          int64_t row_stride, int64_t n_classes, int64_t mode, bool use_2cta, bool stream_k) {
  c10::cuda::CUDAGuard guard(X.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TORCH_CHECK(X.is_cuda() && X.scalar_type() == at::kFloat && X.dim() == 2 && X.stride(1) == 1);
  TORCH_CHECK(D.is_cuda() && D.scalar_type() == at::kFloat && D.dim() == 2 && D.stride(1) == 1);
  const int Bk = (int)X.size(0), F = (int)X.size(1);
  TORCH_CHECK(Bk % TC_BK == 0, "batch must be a multiple of 32 for the tensor-core backward");
  TORCH_CHECK(D.size(0) >= Bk && D.stride(0) == row_stride);
  CUtensorMap tmA, tmD;
  // A^T tiles come from X [B, F] (features contiguous): box = 32 features x 32 instances, 32B-atom swizzle
  encode_2d_f32(&tmA, X.data_ptr<float>(), (uint64_t)F, (uint64_t)Bk, (uint64_t)X.stride(0), 32, TC_BK,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  encode_2d_f32(&tmD, D.data_ptr<float>(), (uint64_t)row_stride, (uint64_t)Bk, (uint64_t)D.stride(0), 32, TC_BK,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  auto dmap = at::empty({(int64_t)sizeof(CUtensorMap)}, at::TensorOptions().dtype(at::kByte).device(X.device()));
  cudaMemcpyAsync(dmap.data_ptr(), &tmD, sizeof(CUtensorMap), cudaMemcpyHostToDevice, stream);
  cudaFuncSetAttribute(dense_gemm_tf32_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  TableView W;
  W.block_ptrs = reinterpret_cast<const long long*>(block_ptrs.data_ptr<int64_t>());
  W.num_blocks = (int)block_ptrs.numel() - 2;   // two guard words follow the pointers
  W.rows_per_block = (int)rows_per_block;
  W.row_stride = (int)row_stride;
  W.row_dim = (int)n_classes;
  W.mode = (int)mode;
  const int N = (int)row_stride;
  if (use_2cta && stream_k) {
    launch_stream_k<true, true>(tmA, reinterpret_cast<const CUtensorMap*>(dmap.data_ptr()), nullptr, 0, F, N, Bk, 1 << 30,
                                W, nullptr, stream);
    return;
  }
  if (use_2cta) {
    cudaFuncSetAttribute(dense_gemm_tf32_2cta_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES);
    dim3 grid2(2 * ((F + 255) / 256), (N + T2_BN - 1) / T2_BN);
    dense_gemm_tf32_2cta_kernel<true, true><<<grid2, TC_THREADS, T2_SMEM_BYTES, stream>>>(
        tmA, reinterpret_cast<const CUtensorMap*>(dmap.data_ptr()), nullptr, 0, F, N, Bk, 1 << 30, W);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return;
  }
  dim3 grid((N + TC_BN - 1) / TC_BN, (F + TC_BM - 1) / TC_BM);
  dense_gemm_tf32_kernel<true, true><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(
      tmA, reinterpret_cast<const CUtensorMap*>(dmap.data_ptr()), nullptr, 0, F, N, Bk, 1 << 30, W);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace hb
