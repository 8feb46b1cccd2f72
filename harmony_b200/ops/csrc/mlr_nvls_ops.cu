// MLR ship mode, second generation: NVLS (multimem) collectives fused into the compute kernels and
// tile-granular release/acquire flags instead of device-wide barriers.
//
//   K1 ship_fwd_kernel      owner g: partial logits of EVERY worker's instances over the features g owns
//                           (row gathers are local HBM) -> plain stores into g's symmetric buffer P_g.
//                           Rows are visited tile-major across workers; when the last row of tile (w, t) is
//                           done the CTA that finished it bumps worker w's flag pflag_w[t] over NVLink
//                           (red.release.sys).
//   K2 ship_softmax_kernel  worker w, one CTA per tile, resident next to K1 on a high-priority stream: waits
//                           until all owners have flagged tile t, then
//                             reduce-scatter : multimem.ld_reduce.add.v4.f32 on the multicast address of P
//                                              (the NVSwitch sums the 8 replicas; 16 MB ingress instead of 112)
//                             softmax / loss / error row D = -step (p - y) in registers
//                             all-gather     : multimem.st.v4.f32 on the multicast address of D (one store,
//                                              the switch replicates it into every owner's buffer)
//                           and bumps every owner's dcount (red.release.sys) once per tile.
//   K3 (mlr_ops.cu)         owner g: waits for dcount == all tiles of all workers (acquire), then folds the
//                           error rows into its own model rows (local HBM).
// No barrier kernel, no NCCL: 3 launches per step, K2 overlaps K1 tile by tile.  Without a multicast mapping
// (1 GPU, in-process executors, fabric without NVLS) K2 falls back to unicast peer loads / stores -- same
// kernel, template flag.  Every spin is bounded (error flag, no hang).
// Arithmetic: MLRTrainer.java:374-398,475-505; replaces the message fan-out of TableImpl.java:366-408,460-501.
#include "hb_common.cuh"
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <map>
#include <mutex>

namespace hb {

static constexpr int NV_THREADS = 256;
static constexpr float NV_ZERO_THRESHOLD = 1e-7f;
// ---------------------------------------------------------------------------------------------------------
// K1: rows = world * B (row id = worker * B + b), CSR restricted to the features this GPU owns.
// ---------------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(NV_THREADS) ship_fwd_kernel(
    TableView W, const int* __restrict__ csr_ptr, const int* __restrict__ csr_col,
    const float* __restrict__ csr_val, int world, int B, int tile_rows, int n_tiles,
    float* __restrict__ P, int p_stride, unsigned* __restrict__ tile_done,
    const long long* __restrict__ pflag_ptrs, unsigned epoch) {
  __shared__ const float* s_row[NV_THREADS];
  __shared__ float s_val[NV_THREADS];
  const int tid = threadIdx.x;
  const int rows_per_round = world * tile_rows;          // one tile of every worker
  const int n_virtual = n_tiles * rows_per_round;
  for (int v = blockIdx.x; v < n_virtual; v += gridDim.x) {
    const int t = v / rows_per_round;
    const int rem = v - t * rows_per_round;
    const int w_id = rem / tile_rows;
    const int b = t * tile_rows + (rem - w_id * tile_rows);
    if (b >= B) continue;                                 // ragged last tile (uniform across the CTA)
    const int r = w_id * B + b;
    const int beg = csr_ptr[r], end = csr_ptr[r + 1];
    float4 acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += NV_THREADS) {
      const int cnt = min(NV_THREADS, end - base);
      __syncthreads();
      if (tid < cnt) {
        s_row[tid] = row_ptr<float>(W, (long long)csr_col[base + tid]);
        s_val[tid] = csr_val[base + tid];
      }
      __syncthreads();
      constexpr int U = (VEC == 1) ? 8 : (VEC == 2 ? 4 : 2);
      int j = 0;
      for (; j + U <= cnt; j += U) {
        float4 wv[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const int c0 = 4 * (tid + k * NV_THREADS);
            wv[u][k] = (c0 < W.row_stride) ? ld_stream_f4(s_row[j + u] + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float x = s_val[j + u];
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            acc[k].x = fmaf(x, wv[u][k].x, acc[k].x); acc[k].y = fmaf(x, wv[u][k].y, acc[k].y);
            acc[k].z = fmaf(x, wv[u][k].z, acc[k].z); acc[k].w = fmaf(x, wv[u][k].w, acc[k].w);
          }
        }
      }
      for (; j < cnt; ++j) {
        const float x = s_val[j];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const int c0 = 4 * (tid + k * NV_THREADS);
          if (c0 < W.row_stride) {
            const float4 w4 = ld_stream_f4(s_row[j] + c0);
            acc[k].x = fmaf(x, w4.x, acc[k].x); acc[k].y = fmaf(x, w4.y, acc[k].y);
            acc[k].z = fmaf(x, w4.z, acc[k].z); acc[k].w = fmaf(x, w4.w, acc[k].w);
          }
        }
      }
    }
    // the row is stored even when this owner holds none of its features: P needs no zeroing pass
    float* dst = P + (size_t)r * p_stride;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c0 = 4 * (tid + k * NV_THREADS);
      if (c0 < p_stride) *reinterpret_cast<float4*>(dst + c0) = acc[k];
    }
    if (tile_done != nullptr) {
      __syncthreads();                                    // every thread's stores precede the fence below
      if (tid == 0) {
        __threadfence();
        const int cnt_rows = min(tile_rows, B - t * tile_rows);
        const unsigned old = atomicAdd(tile_done + w_id * n_tiles + t, 1u);
        if (old + 1u == epoch * (unsigned)cnt_rows) {     // last row of tile (w, t) on this owner
          __threadfence_system();
          red_release_sys_add_u32(reinterpret_cast<unsigned*>(pflag_ptrs[w_id]) + t, 1u);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K1, third generation (the default): PUSH-based reduce-scatter.  Same gather loop as above, but the partial
// logits of instance (w, b) leave the SM as 128-bit red.add straight into worker w's logits buffer (local or
// NVLink peer mapping): fire-and-forget, so the transfer hides behind the next rows' gathers, and no P buffer
// exists (-256 MB of HBM traffic per step at 8 GPUs).
// A CTA owns a whole TILE (w, t) of `tile_rows` instances: it issues the reductions of all of them, fences
// ONCE at system scope (one NVLink round trip per tile, not per instance -- per-instance fences cost 2x on real
// links) and raises worker w's flag itself: no tile counter.  Tiles are handed out tile-major by an atomic
// work counter, so the SMs stay evenly loaded and the workers' tiles complete in order for K2.
// ---------------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(NV_THREADS) ship3_fwd_kernel(
    TableView W, const int* __restrict__ csr_ptr, const int* __restrict__ csr_col,
    const float* __restrict__ csr_val, int world, int B, int tile_rows, int n_tiles,
    const long long* __restrict__ logit_ptrs, int l_stride, unsigned* __restrict__ work_ctr,
    const long long* __restrict__ pflag_ptrs, unsigned epoch) {
  __shared__ const float* s_row[NV_THREADS];
  __shared__ float s_val[NV_THREADS];
  __shared__ int s_vt;
  const int tid = threadIdx.x;
  const int n_vtiles = n_tiles * world;
  unsigned* ctr = work_ctr + (epoch & 1u);
  if (blockIdx.x == 0 && tid == 0) work_ctr[(epoch + 1u) & 1u] = 0u;     // re-arm the NEXT launch's counter
  while (true) {
    __syncthreads();                                        // s_vt / s_row of the previous tile are consumed
    if (tid == 0) s_vt = (int)atomicAdd(ctr, 1u);
    __syncthreads();
    const int vt = s_vt;
    if (vt >= n_vtiles) break;
    const int t = vt / world, w_id = vt - t * world;
    const int b_end = min(B, (t + 1) * tile_rows);
    for (int b = t * tile_rows; b < b_end; ++b) {
      const int r = w_id * B + b;
      const int beg = csr_ptr[r], end = csr_ptr[r + 1];
      if (beg == end) continue;                             // (uniform across the CTA)
      float4 acc[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int base = beg; base < end; base += NV_THREADS) {
        const int cnt = min(NV_THREADS, end - base);
        __syncthreads();
        if (tid < cnt) {
          s_row[tid] = row_ptr<float>(W, (long long)csr_col[base + tid]);
          s_val[tid] = csr_val[base + tid];
        }
        __syncthreads();
        constexpr int U = (VEC == 1) ? 8 : (VEC == 2 ? 4 : 2);
        for (int j = 0; j < cnt; j += U) {                  // the tail group is padded: every load precedes the FMAs
          float4 wv[U][VEC];
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              const int c0 = 4 * (tid + k * NV_THREADS);
              wv[u][k] = (j + u < cnt && c0 < W.row_stride) ? ld_stream_f4(s_row[j + u] + c0)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const float x = (j + u < cnt) ? s_val[j + u] : 0.f;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              acc[k].x = fmaf(x, wv[u][k].x, acc[k].x); acc[k].y = fmaf(x, wv[u][k].y, acc[k].y);
              acc[k].z = fmaf(x, wv[u][k].z, acc[k].z); acc[k].w = fmaf(x, wv[u][k].w, acc[k].w);
            }
          }
        }
      }
      // reduce-scatter epilogue: add the partial logits into the owning worker's buffer (peer memory)
      float* dst = reinterpret_cast<float*>(logit_ptrs[w_id]) + (size_t)b * l_stride;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const int c0 = 4 * (tid + k * NV_THREADS);
        if (c0 < l_stride) red_add_f4_sys(dst + c0, acc[k]);
      }
    }
    if (pflag_ptrs != nullptr) {
      __threadfence_system();                               // MY reductions of this tile are performed system-wide
      __syncthreads();                                      // ... and so are those of every thread of the CTA
      if (tid == 0) red_release_sys_add_u32(reinterpret_cast<unsigned*>(pflag_ptrs[w_id]) + t, 1u);
    }
  }
}

// Warp-per-instance K1 (row_stride <= 1024 floats).  With the model sharded over N GPUs an instance has only
// nnz / N of its features at this owner (12 at N = 8): a CTA per instance would spend its time in barriers and in a
// serial tail.  Here every warp owns one instance: lane j fetches non-zero j (column, value), the pair is broadcast
// with shuffles, and each lane keeps 2 rows x NV 128-bit loads in flight -- no shared memory, no block barrier.
static constexpr int FW_THREADS = 128;
template <int NV>
__global__ void __launch_bounds__(FW_THREADS) ship_fwd_warp_kernel(
    TableView W, const int* __restrict__ csr_ptr, const int* __restrict__ csr_col,
    const float* __restrict__ csr_val, int world, int B, int tile_rows, int n_tiles,
    float* __restrict__ P, int p_stride, unsigned* __restrict__ tile_done,
    const long long* __restrict__ pflag_ptrs, unsigned epoch) {
  const int lane = threadIdx.x & 31;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int rows_per_round = world * tile_rows;
  const int n_virtual = n_tiles * rows_per_round;
  for (int v = wglobal; v < n_virtual; v += nwarps) {
    const int t = v / rows_per_round;
    const int rem = v - t * rows_per_round;
    const int w_id = rem / tile_rows;
    const int b = t * tile_rows + (rem - w_id * tile_rows);
    if (b >= B) continue;
    const int r = w_id * B + b;
    const int beg = csr_ptr[r], end = csr_ptr[r + 1];
    float4 acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      const int my_c = (lane < cnt) ? csr_col[base + lane] : 0;
      const float my_x = (lane < cnt) ? csr_val[base + lane] : 0.f;
      int j = 0;
      for (; j + 2 <= cnt; j += 2) {
        const float* r0 = row_ptr<float>(W, (long long)__shfl_sync(0xffffffffu, my_c, j));
        const float* r1 = row_ptr<float>(W, (long long)__shfl_sync(0xffffffffu, my_c, j + 1));
        const float x0 = __shfl_sync(0xffffffffu, my_x, j), x1 = __shfl_sync(0xffffffffu, my_x, j + 1);
        float4 a[NV], c[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int c0 = 4 * (lane + 32 * k);
          a[k] = (c0 < W.row_stride) ? ld_stream_f4(r0 + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
          c[k] = (c0 < W.row_stride) ? ld_stream_f4(r1 + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          acc[k].x = fmaf(x0, a[k].x, fmaf(x1, c[k].x, acc[k].x)); acc[k].y = fmaf(x0, a[k].y, fmaf(x1, c[k].y, acc[k].y));
          acc[k].z = fmaf(x0, a[k].z, fmaf(x1, c[k].z, acc[k].z)); acc[k].w = fmaf(x0, a[k].w, fmaf(x1, c[k].w, acc[k].w));
        }
      }
      if (j < cnt) {
        const float* r0 = row_ptr<float>(W, (long long)__shfl_sync(0xffffffffu, my_c, j));
        const float x0 = __shfl_sync(0xffffffffu, my_x, j);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int c0 = 4 * (lane + 32 * k);
          if (c0 < W.row_stride) {
            const float4 a = ld_stream_f4(r0 + c0);
            acc[k].x = fmaf(x0, a.x, acc[k].x); acc[k].y = fmaf(x0, a.y, acc[k].y);
            acc[k].z = fmaf(x0, a.z, acc[k].z); acc[k].w = fmaf(x0, a.w, acc[k].w);
          }
        }
      }
    }
    float* dst = P + (size_t)r * p_stride;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane + 32 * k);
      if (c0 < p_stride) *reinterpret_cast<float4*>(dst + c0) = acc[k];
    }
    if (tile_done != nullptr) {
      __syncwarp();                                       // every lane's stores precede the fence below
      if (lane == 0) {
        __threadfence();
        const int cnt_rows = min(tile_rows, B - t * tile_rows);
        const unsigned old = atomicAdd(tile_done + w_id * n_tiles + t, 1u);
        if (old + 1u == epoch * (unsigned)cnt_rows) {     // last row of tile (w, t) on this owner
          __threadfence_system();
          red_release_sys_add_u32(reinterpret_cast<unsigned*>(pflag_ptrs[w_id]) + t, 1u);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2: one CTA per tile, one warp per instance (row_stride <= 1024 floats: NV float4 chunks per lane).
//   USE_MC : src_mc / dst_mc are multicast addresses of P / D.
//   !USE_MC: src_ptrs[n_src] are unicast buffers to sum, dst_ptrs[n_dst] are unicast destinations.
// Source row = src_row0 + b, destination row = dst_row0 + b.
// ---------------------------------------------------------------------------------------------------------
static constexpr int SM_THREADS = 128;      // 4 warps x 2 instances per tile; 10 K registers: fits next to the K1 CTAs
template <int NV, bool SRC_MC, bool DST_MC>
__global__ void __launch_bounds__(SM_THREADS) ship_softmax_kernel(
    const float* __restrict__ src_mc, const long long* __restrict__ src_ptrs, int n_src, int src_row0,
    float* __restrict__ dst_mc, const long long* __restrict__ dst_ptrs, int n_dst, int dst_row0,
    int stride, const int* __restrict__ labels, int B, int C, float step, float lscale, int tile_rows, int n_tiles,
    const unsigned* __restrict__ pflag, unsigned pflag_target,
    const long long* __restrict__ dcount_ptrs, float* __restrict__ stats, int* __restrict__ err, int zero_src) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __shared__ int s_ok;
  float wloss = 0.f, whit = 0.f;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    if (pflag != nullptr) {
      if (threadIdx.x == 0) s_ok = spin_until_ge(pflag + t, pflag_target, err, 2) ? 1 : 0;
      __syncthreads();
    }
    const int b_end = min(B, (t + 1) * tile_rows);
    for (int b = t * tile_rows + wid; b < b_end; b += SM_THREADS / 32) {
      float4 v[NV];
      float tmax = -INFINITY;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane + 32 * k);
        v[k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (c0 < stride) {
          const size_t off = (size_t)(src_row0 + b) * stride + c0;
          if (SRC_MC) {
            v[k] = multimem_ld_reduce_add_f4(src_mc + off);
          } else {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int g = 0; g < n_src; ++g) {
              float* sp = reinterpret_cast<float*>(src_ptrs[g]) + off;
              const float4 a = ld_stream_f4(sp);
              s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
              // push-based reduce-scatter (v3): the owners red.add into this buffer, the reader re-arms it
              if (zero_src) *reinterpret_cast<float4*>(sp) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            v[k] = s;
          }
          v[k].x *= lscale; v[k].y *= lscale; v[k].z *= lscale; v[k].w *= lscale;   // lazy L2: w = s * v
          if (c0 + 0 >= C) v[k].x = -INFINITY;
          if (c0 + 1 >= C) v[k].y = -INFINITY;
          if (c0 + 2 >= C) v[k].z = -INFINITY;
          if (c0 + 3 >= C) v[k].w = -INFINITY;
        }
        tmax = fmaxf(tmax, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
      }
      const float m = warp_max(tmax);
      float tsum = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        v[k].x = __expf(v[k].x - m); v[k].y = __expf(v[k].y - m);
        v[k].z = __expf(v[k].z - m); v[k].w = __expf(v[k].w - m);
        tsum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      }
      const float inv = 1.f / warp_sum(tsum);
      const int y = labels[b];
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane + 32 * k);
        if (c0 >= stride) continue;
        const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
        float d[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = c0 + q;
          float p = fminf(fmaxf(e[q] * inv, NV_ZERO_THRESHOLD), 1.f - NV_ZERO_THRESHOLD);
          if (c == y) { wloss += -__logf(p); whit += (e[q] >= 1.f) ? 1.f : 0.f; p -= 1.f; }
          d[q] = (c < C && y >= 0) ? -step * p : 0.f;     // label -1 = padding batch: contributes nothing
        }
        const float4 dv = make_float4(d[0], d[1], d[2], d[3]);
        const size_t off = (size_t)(dst_row0 + b) * stride + c0;
        if (DST_MC) {
          multimem_st_f4(dst_mc + off, dv);
        } else {
          for (int g = 0; g < n_dst; ++g) st_stream_f4(reinterpret_cast<float*>(dst_ptrs[g]) + off, dv);
        }
      }
    }
    if (dcount_ptrs != nullptr) {
      __syncthreads();                                    // the whole tile has been stored
      if (wid == 0 && lane < n_dst) {
        __threadfence_system();                           // ... and is visible system-wide before the count moves
        red_release_sys_add_u32(reinterpret_cast<unsigned*>(dcount_ptrs[lane]), 1u);
      }
    }
  }
  wloss = warp_sum(wloss);
  whit = warp_sum(whit);
  if (lane == 0 && stats != nullptr && (wloss != 0.f || whit != 0.f)) {
    atomicAdd(stats + 0, wloss);
    atomicAdd(stats + 1, whit);
  }
}

// Tiny stream-ordered wait used where a whole kernel must follow remote progress (K3 of the wide-row path).
__global__ void wait_count_kernel(const unsigned* __restrict__ p, unsigned target, int* __restrict__ err) {
  if (threadIdx.x == 0) spin_until_ge(p, target, err, 3);
}

static TableView nv_view(const at::Tensor& block_ptrs, int64_t rows_per_block, int64_t row_stride,
                         int64_t row_dim, int64_t mode) {
  TableView v;
  v.block_ptrs = reinterpret_cast<const long long*>(block_ptrs.data_ptr<int64_t>());
  v.num_blocks = (int)block_ptrs.numel() - 2;   // two guard words follow the pointers
  v.rows_per_block = (int)rows_per_block;
  v.row_stride = (int)row_stride;
  v.row_dim = (int)row_dim;
  v.mode = (int)mode;
  return v;
}

// Leaves room for the resident K2 CTAs: K1 takes (occupancy - 1) CTAs per SM.
template <typename K>
static int fwd_grid(K kernel, int threads = NV_THREADS) {
  static std::mutex mu;
  static std::map<const void*, int> cache;              // keyed by the kernel (instantiations share one K type)
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(reinterpret_cast<const void*>(kernel));
  if (it != cache.end()) return it->second;
  int occ = 0, sms = 148, dev = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0);
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * std::max(1, occ - 1);
  cache[reinterpret_cast<const void*>(kernel)] = grid;
  return grid;
}

// CUDA loads a kernel lazily at its FIRST launch, and that load can synchronise with running kernels: K2 spinning
// on flags that only a not-yet-loaded K1 can raise would wait for its own timeout.  Touching the function attributes
// loads every instantiation up front (called once per device from MLRShipModel.__init__).
void mlr_preload_bwd();   // mlr_ops.cu
void ship_preload(int64_t device) {
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  cudaFuncAttributes a;
#define HB_TOUCH(K) C10_CUDA_CHECK(cudaFuncGetAttributes(&a, K))
  HB_TOUCH(ship_fwd_kernel<1>); HB_TOUCH(ship_fwd_kernel<2>); HB_TOUCH(ship_fwd_kernel<4>);
  HB_TOUCH(ship_fwd_warp_kernel<1>); HB_TOUCH(ship_fwd_warp_kernel<2>); HB_TOUCH(ship_fwd_warp_kernel<4>);
  HB_TOUCH(ship_fwd_warp_kernel<8>);
  HB_TOUCH((ship_softmax_kernel<1, true, true>)); HB_TOUCH((ship_softmax_kernel<2, true, true>));
  HB_TOUCH((ship_softmax_kernel<4, true, true>)); HB_TOUCH((ship_softmax_kernel<8, true, true>));
  HB_TOUCH((ship_softmax_kernel<1, false, true>)); HB_TOUCH((ship_softmax_kernel<2, false, true>));
  HB_TOUCH((ship_softmax_kernel<4, false, true>)); HB_TOUCH((ship_softmax_kernel<8, false, true>));
  HB_TOUCH((ship_softmax_kernel<1, false, false>)); HB_TOUCH((ship_softmax_kernel<2, false, false>));
  HB_TOUCH((ship_softmax_kernel<4, false, false>)); HB_TOUCH((ship_softmax_kernel<8, false, false>));
  HB_TOUCH(ship3_fwd_kernel<1>); HB_TOUCH(ship3_fwd_kernel<2>); HB_TOUCH(ship3_fwd_kernel<4>);
  HB_TOUCH(wait_count_kernel);
#undef HB_TOUCH
  mlr_preload_bwd();
}

// P: [world*B, stride] local buffer; tile_done: uint32 [world*n_tiles] (0 => no signalling: the NCCL arm)
void ship_forward(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t num_classes,
                  int64_t mode, at::Tensor csr_ptr, at::Tensor csr_col, at::Tensor csr_val, int64_t world, int64_t B,
                  int64_t tile_rows, int64_t p_ptr, int64_t tile_done_ptr, c10::optional<at::Tensor> pflag_ptrs,
                  int64_t epoch) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView W = nv_view(block_ptrs, rows_per_block, row_stride, num_classes, mode);
  TORCH_CHECK(csr_ptr.numel() - 1 == world * B, "routed CSR must have world*B rows");
  TORCH_CHECK(row_stride % 32 == 0 && row_stride <= 4096);
  TORCH_CHECK(csr_ptr.scalar_type() == at::kInt && csr_col.scalar_type() == at::kInt &&
              csr_val.scalar_type() == at::kFloat);
  const int n_tiles = (int)((B + tile_rows - 1) / tile_rows);
  const int vec = (int)((row_stride / 4 + NV_THREADS - 1) / NV_THREADS);
  const long long* pf = pflag_ptrs.has_value()
      ? reinterpret_cast<const long long*>(pflag_ptrs->data_ptr<int64_t>()) : nullptr;
  unsigned* td = pf ? reinterpret_cast<unsigned*>(tile_done_ptr) : nullptr;
#define LAUNCH_F(V)                                                                                     \
  {                                                                                                     \
    const int grid = std::min((int)(n_tiles * world * tile_rows), fwd_grid(ship_fwd_kernel<V>));        \
    ship_fwd_kernel<V><<<grid, NV_THREADS, 0, stream>>>(                                                \
        W, csr_ptr.data_ptr<int>(), csr_col.data_ptr<int>(), csr_val.data_ptr<float>(), (int)world,     \
        (int)B, (int)tile_rows, n_tiles, reinterpret_cast<float*>(p_ptr), (int)row_stride, td, pf,      \
        (unsigned)epoch);                                                                               \
  }
#define LAUNCH_FW(NVv)                                                                                  \
  {                                                                                                     \
    const int want = (int)((n_tiles * world * tile_rows + FW_THREADS / 32 - 1) / (FW_THREADS / 32));    \
    const int grid = std::min(want, fwd_grid(ship_fwd_warp_kernel<NVv>, FW_THREADS));                   \
    ship_fwd_warp_kernel<NVv><<<grid, FW_THREADS, 0, stream>>>(                                         \
        W, csr_ptr.data_ptr<int>(), csr_col.data_ptr<int>(), csr_val.data_ptr<float>(), (int)world,     \
        (int)B, (int)tile_rows, n_tiles, reinterpret_cast<float*>(p_ptr), (int)row_stride, td, pf,      \
        (unsigned)epoch);                                                                               \
  }
  // warp per instance pays off for SHORT rows (model sharded over many owners); long rows keep the CTA-per-instance
  // kernel with 8 rows in flight per thread.  HB_SHIP_K1=cta|warp overrides (A/B runs).
  const double avg_nnz = (double)csr_col.numel() / (double)std::max<int64_t>(1, world * B);
  static const char* k1_env = getenv("HB_SHIP_K1");
  const bool use_warp = row_stride <= 1024 &&
      (k1_env && k1_env[0] == 'w' ? true : (k1_env && k1_env[0] == 'c' ? false : avg_nnz < 32.0));
  if (use_warp) {
    const int nv = (int)((row_stride / 4 + 31) / 32);
    if (nv <= 1) LAUNCH_FW(1) else if (nv <= 2) LAUNCH_FW(2) else if (nv <= 4) LAUNCH_FW(4) else LAUNCH_FW(8)
  } else if (vec <= 1) LAUNCH_F(1) else if (vec == 2) LAUNCH_F(2) else LAUNCH_F(4)
#undef LAUNCH_FW
#undef LAUNCH_F
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// v3 forward: logit_ptrs[w] = worker w's [B, stride] logits buffer as mapped here; tile signalling as in v2.
void ship3_forward(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t num_classes,
                   int64_t mode, at::Tensor csr_ptr, at::Tensor csr_col, at::Tensor csr_val, int64_t world, int64_t B,
                   int64_t tile_rows, at::Tensor logit_ptrs, int64_t tile_done_ptr,
                   c10::optional<at::Tensor> pflag_ptrs, int64_t epoch) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView W = nv_view(block_ptrs, rows_per_block, row_stride, num_classes, mode);
  TORCH_CHECK(csr_ptr.numel() - 1 == world * B, "routed CSR must have world*B rows");
  TORCH_CHECK(row_stride % 32 == 0 && row_stride <= 4096 && logit_ptrs.numel() == world);
  const int n_tiles = (int)((B + tile_rows - 1) / tile_rows);
  const int vec = (int)((row_stride / 4 + NV_THREADS - 1) / NV_THREADS);
  const long long* pf = pflag_ptrs.has_value()
      ? reinterpret_cast<const long long*>(pflag_ptrs->data_ptr<int64_t>()) : nullptr;
  unsigned* td = reinterpret_cast<unsigned*>(tile_done_ptr);      // uint32[2]: the per-launch work counters
  TORCH_CHECK(td != nullptr, "ship3_forward needs the two work counters");
  // persistent CTAs fetch tiles from the work counter; one CTA slot per SM is left to the resident K2
  static const int spare = getenv("HB_SHIP_K1_SPARE") ? atoi(getenv("HB_SHIP_K1_SPARE")) : 1;   // CTA slots per SM left to K2
  const int grid = (int)std::min<int64_t>((int64_t)n_tiles * world, fwd_grid(ship3_fwd_kernel<1>) + (1 - spare) * 148);
#define LAUNCH_F3(V)                                                                                    \
  ship3_fwd_kernel<V><<<grid, NV_THREADS, 0, stream>>>(                                                 \
      W, csr_ptr.data_ptr<int>(), csr_col.data_ptr<int>(), csr_val.data_ptr<float>(), (int)world, (int)B, \
      (int)tile_rows, n_tiles, reinterpret_cast<const long long*>(logit_ptrs.data_ptr<int64_t>()),      \
      (int)row_stride, td, pf, (unsigned)epoch);
  if (vec <= 1) { LAUNCH_F3(1) } else if (vec == 2) { LAUNCH_F3(2) } else { LAUNCH_F3(4) }
#undef LAUNCH_F3
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void ship_softmax(int64_t src_mc, c10::optional<at::Tensor> src_ptrs, int64_t src_row0, int64_t dst_mc,
                  c10::optional<at::Tensor> dst_ptrs, int64_t dst_row0, int64_t stride, at::Tensor labels,
                  int64_t num_classes, double step, double logit_scale, int64_t tile_rows, int64_t pflag_ptr,
                  int64_t pflag_target, c10::optional<at::Tensor> dcount_ptrs, at::Tensor stats, at::Tensor err,
                  bool zero_src) {
  c10::cuda::CUDAGuard guard(labels.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int B = (int)labels.numel();
  if (B == 0) return;
  TORCH_CHECK(stride % 32 == 0 && stride <= 1024, "ship_softmax: row stride must be <= 1024 floats");
  TORCH_CHECK(labels.scalar_type() == at::kInt && err.scalar_type() == at::kInt);
  const bool src_is_mc = src_mc != 0, dst_is_mc = dst_mc != 0;
  TORCH_CHECK(src_is_mc || src_ptrs.has_value(), "unicast source needs a pointer list");
  TORCH_CHECK(dst_is_mc || dst_ptrs.has_value(), "unicast destination needs a pointer list");
  TORCH_CHECK(!zero_src || (!src_is_mc && src_ptrs->numel() == 1), "zero_src re-arms ONE local source buffer");
  const int n_tiles = (int)((B + tile_rows - 1) / tile_rows);
  const int nv = (int)((stride / 4 + 31) / 32);
  const long long* sp = src_ptrs.has_value() ? reinterpret_cast<const long long*>(src_ptrs->data_ptr<int64_t>()) : nullptr;
  const long long* dp = dst_ptrs.has_value() ? reinterpret_cast<const long long*>(dst_ptrs->data_ptr<int64_t>()) : nullptr;
  const int n_src = src_ptrs.has_value() ? (int)src_ptrs->numel() : 0;
  const int n_dst = dcount_ptrs.has_value() ? (int)dcount_ptrs->numel() : (dst_ptrs.has_value() ? (int)dst_ptrs->numel() : 0);
  TORCH_CHECK(n_dst <= 32);
  const long long* dc = dcount_ptrs.has_value() ? reinterpret_cast<const long long*>(dcount_ptrs->data_ptr<int64_t>()) : nullptr;
  const int grid = std::min(n_tiles, 148);
#define LAUNCH_S(NVv, SMC, DMC)                                                                          \
  ship_softmax_kernel<NVv, SMC, DMC><<<grid, SM_THREADS, 0, stream>>>(                                         \
      reinterpret_cast<const float*>(src_mc), sp, n_src, (int)src_row0, reinterpret_cast<float*>(dst_mc), \
      dp, dp ? (int)dst_ptrs->numel() : n_dst, (int)dst_row0, (int)stride, labels.data_ptr<int>(), B,    \
      (int)num_classes, (float)step, (float)logit_scale, (int)tile_rows, n_tiles,                       \
      reinterpret_cast<const unsigned*>(pflag_ptr),                                                     \
      (unsigned)pflag_target, dc, stats.data_ptr<float>(), err.data_ptr<int>(), zero_src ? 1 : 0);
#define DISPATCH_S(SMC, DMC)                                                                             \
  if (nv <= 1) { LAUNCH_S(1, SMC, DMC) } else if (nv <= 2) { LAUNCH_S(2, SMC, DMC) }                     \
  else if (nv <= 4) { LAUNCH_S(4, SMC, DMC) } else { LAUNCH_S(8, SMC, DMC) }
  if (src_is_mc && dst_is_mc) { DISPATCH_S(true, true) }
  else if (!src_is_mc && dst_is_mc) { DISPATCH_S(false, true) }
  else if (!src_is_mc && !dst_is_mc) { DISPATCH_S(false, false) }
  else { TORCH_CHECK(false, "multicast source with unicast destinations is not instantiated"); }
#undef DISPATCH_S
#undef LAUNCH_S
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void wait_count(int64_t ptr, int64_t target, at::Tensor err) {
  c10::cuda::CUDAGuard guard(err.device());
  wait_count_kernel<<<1, 32, 0, at::cuda::getCurrentCUDAStream()>>>(reinterpret_cast<const unsigned*>(ptr),
                                                                    (unsigned)target, err.data_ptr<int>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace hb
