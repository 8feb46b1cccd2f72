// MLR "ship the activations" mode: the sparse dot runs on the GPU that OWNS the feature rows and
// only [B, C] activations cross NVLink (12x fewer bytes than moving 4 KB rows at 8 GPUs).
//
//   phase 1  mlr_partial_fwd_kernel   owner g: partial logits of EVERY worker's instances over the
//                                     features g owns (all row reads are local HBM), reduce-added
//                                     straight into the worker's logits buffer over NVLink
//                                     = compute fused with a reduce-scatter over peer memory
//   phase 2  mlr_softmax_bcast_kernel worker w: softmax / loss of its own instances, the pre-scaled
//                                     error rows D = -step (p - y) are stored into EVERY owner's D
//                                     buffer = compute fused with an all-gather over peer memory
//   phase 3  mlr_bwd_push_kernel      owner g: gradient rows of its own features from all workers'
//                                     D (local), w += delta (local)
// Device clocks (parallel/device_clock.py) separate the phases.  The arithmetic is identical to the
// pull/push mode (MLRTrainer.java:374-398,475-505); only the placement of the dot changes -- in the
// reference's co-located mode (DolphinJobEntity.java:80-82) every executor is server AND worker.
#include "hb_common.cuh"
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

namespace hb {

static constexpr int SHIP_THREADS = 256;
static constexpr float SHIP_ZERO_THRESHOLD = 1e-7f;

__device__ __forceinline__ float4 ship_ld_row(const float* r, int c0, int width) {
  return (c0 < width) ? ld_stream_f4(r + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// rows = N*B instances (row id = worker * B + b); CSR restricted to the features this GPU owns.
template <int VEC>
__global__ void __launch_bounds__(SHIP_THREADS) mlr_partial_fwd_kernel(
    TableView W, const int* __restrict__ csr_ptr, const int* __restrict__ csr_col,
    const float* __restrict__ csr_val, int n_rows, int B, const long long* __restrict__ logit_ptrs,
    int l_stride) {
  __shared__ const float* s_row[SHIP_THREADS];
  __shared__ float s_val[SHIP_THREADS];
  const int tid = threadIdx.x;
  for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const int beg = csr_ptr[r], end = csr_ptr[r + 1];
    if (beg == end) continue;                      // uniform across the CTA
    float4 acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += SHIP_THREADS) {
      const int cnt = min(SHIP_THREADS, end - base);
      __syncthreads();
      if (tid < cnt) {
        s_row[tid] = row_ptr<float>(W, (long long)csr_col[base + tid]);
        s_val[tid] = csr_val[base + tid];
      }
      __syncthreads();
      constexpr int U = (VEC == 1) ? 8 : (VEC == 2 ? 4 : 2);
      int j = 0;
      for (; j + U <= cnt; j += U) {
        float4 w[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int k = 0; k < VEC; ++k) w[u][k] = ship_ld_row(s_row[j + u], 4 * (tid + k * SHIP_THREADS), W.row_stride);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float x = s_val[j + u];
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            acc[k].x = fmaf(x, w[u][k].x, acc[k].x); acc[k].y = fmaf(x, w[u][k].y, acc[k].y);
            acc[k].z = fmaf(x, w[u][k].z, acc[k].z); acc[k].w = fmaf(x, w[u][k].w, acc[k].w);
          }
        }
      }
      for (; j < cnt; ++j) {
        const float x = s_val[j];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          float4 w = ship_ld_row(s_row[j], 4 * (tid + k * SHIP_THREADS), W.row_stride);
          acc[k].x = fmaf(x, w.x, acc[k].x); acc[k].y = fmaf(x, w.y, acc[k].y);
          acc[k].z = fmaf(x, w.z, acc[k].z); acc[k].w = fmaf(x, w.w, acc[k].w);
        }
      }
    }
    // reduce-scatter epilogue: add the partial logits into the owning worker's buffer (peer memory)
    const int w_id = r / B, b = r - w_id * B;
    float* dst = reinterpret_cast<float*>(logit_ptrs[w_id]) + (size_t)b * l_stride;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c0 = 4 * (tid + k * SHIP_THREADS);
      if (c0 < l_stride) red_add_f4_sys(dst + c0, acc[k]);
    }
  }
}

// One CTA per local instance: softmax of the fully reduced logits, loss/accuracy, error rows
// broadcast into every owner's D buffer at row (rank * B + b); the logits row is re-zeroed.
template <int VEC>
__global__ void __launch_bounds__(SHIP_THREADS) mlr_softmax_bcast_kernel(
    float* __restrict__ L, int l_stride, const int* __restrict__ labels, int B, int C, float step,
    const long long* __restrict__ d_ptrs, int n_owners, int rank, int d_stride, float* __restrict__ stats) {
  __shared__ float s_red[32];
  const int tid = threadIdx.x;
  float cta_loss = 0.f, cta_hit = 0.f;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float* Lrow = L + (size_t)b * l_stride;
    float l[VEC][4];
    float tmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c0 = 4 * (tid + k * SHIP_THREADS);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 < l_stride) {
        v = *reinterpret_cast<const float4*>(Lrow + c0);
        *reinterpret_cast<float4*>(Lrow + c0) = make_float4(0.f, 0.f, 0.f, 0.f);   // ready for the next step
      }
      l[k][0] = v.x; l[k][1] = v.y; l[k][2] = v.z; l[k][3] = v.w;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (c0 + e >= C) l[k][e] = -INFINITY;
        tmax = fmaxf(tmax, l[k][e]);
      }
    }
    const float m = block_max(tmax, s_red);
    float tsum = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) { l[k][e] = __expf(l[k][e] - m); tsum += l[k][e]; }
    const float inv = 1.f / block_sum(tsum, s_red);
    const int y = labels[b];
    float my_loss = 0.f, my_hit = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c0 = 4 * (tid + k * SHIP_THREADS);
      float d[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = c0 + e;
        float p = fminf(fmaxf(l[k][e] * inv, SHIP_ZERO_THRESHOLD), 1.f - SHIP_ZERO_THRESHOLD);
        if (c == y) { my_loss = -__logf(p); my_hit = (l[k][e] >= 1.f) ? 1.f : 0.f; p -= 1.f; }
        d[e] = (c < C) ? -step * p : 0.f;
      }
      if (c0 < d_stride) {
        const float4 dv = make_float4(d[0], d[1], d[2], d[3]);
        for (int g = 0; g < n_owners; ++g) {       // all-gather epilogue: one store per owner GPU
          float* dst = reinterpret_cast<float*>(d_ptrs[g]) + ((size_t)rank * B + b) * d_stride + c0;
          st_stream_f4(dst, dv);
        }
      }
    }
    cta_loss += my_loss;
    cta_hit += my_hit;
  }
  const float tl = block_sum(cta_loss, s_red);
  const float tc = block_sum(cta_hit, s_red);
  if (tid == 0 && stats != nullptr) { atomicAdd(stats + 0, tl); atomicAdd(stats + 1, tc); }
}

// Warp-per-instance variant for row_stride <= 1024 floats: 8 float4 per lane, warp shuffles only
// (no block barriers), 8 instances per CTA.  NV = float4 chunks per lane.
template <int NV>
__global__ void __launch_bounds__(256) mlr_softmax_bcast_warp_kernel(
    float* __restrict__ L, int l_stride, const int* __restrict__ labels, int B, int C, float step,
    const long long* __restrict__ d_ptrs, int n_owners, int rank, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float wloss = 0.f, whit = 0.f;
  for (int b = wglobal; b < B; b += nwarps) {
    float* Lrow = L + (size_t)b * l_stride;
    float4 v[NV];
    float tmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane + 32 * k);
      v[k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (c0 < l_stride) {
        v[k] = *reinterpret_cast<const float4*>(Lrow + c0);
        *reinterpret_cast<float4*>(Lrow + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
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
      if (c0 >= l_stride) continue;
      float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
      float d[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c0 + q;
        float p = fminf(fmaxf(e[q] * inv, SHIP_ZERO_THRESHOLD), 1.f - SHIP_ZERO_THRESHOLD);
        if (c == y) { wloss += -__logf(p); whit += (e[q] >= 1.f) ? 1.f : 0.f; p -= 1.f; }
        d[q] = (c < C) ? -step * p : 0.f;
      }
      const float4 dv = make_float4(d[0], d[1], d[2], d[3]);
      for (int g = 0; g < n_owners; ++g) {
        float* dst = reinterpret_cast<float*>(d_ptrs[g]) + ((size_t)rank * B + b) * l_stride + c0;
        st_stream_f4(dst, dv);
      }
    }
  }
  wloss = warp_sum(wloss);
  whit = warp_sum(whit);
  if (lane == 0 && stats != nullptr && (wloss != 0.f || whit != 0.f)) { atomicAdd(stats + 0, wloss); atomicAdd(stats + 1, whit); }
}

static TableView ship_view(const at::Tensor& block_ptrs, int64_t rows_per_block, int64_t row_stride,
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

void mlr_partial_forward(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t num_classes,
                         int64_t mode, at::Tensor csr_ptr, at::Tensor csr_col, at::Tensor csr_val, int64_t B,
                         at::Tensor logit_ptrs) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView W = ship_view(block_ptrs, rows_per_block, row_stride, num_classes, mode);
  const int n_rows = (int)csr_ptr.numel() - 1;
  if (n_rows <= 0) return;
  TORCH_CHECK(row_stride % 32 == 0 && row_stride <= 4096);
  const int vec = (int)((row_stride / 4 + SHIP_THREADS - 1) / SHIP_THREADS);
  const int grid = std::min(n_rows, 148 * 64);
#define LAUNCH_PF(V)                                                                                 \
  mlr_partial_fwd_kernel<V><<<grid, SHIP_THREADS, 0, stream>>>(                                      \
      W, csr_ptr.data_ptr<int>(), csr_col.data_ptr<int>(), csr_val.data_ptr<float>(), n_rows, (int)B, \
      reinterpret_cast<const long long*>(logit_ptrs.data_ptr<int64_t>()), (int)row_stride);
  if (vec <= 1) { LAUNCH_PF(1) } else if (vec == 2) { LAUNCH_PF(2) } else { LAUNCH_PF(4) }
#undef LAUNCH_PF
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void mlr_softmax_bcast(at::Tensor L, at::Tensor labels, int64_t num_classes, double step, at::Tensor d_ptrs,
                       int64_t rank, at::Tensor stats) {
  c10::cuda::CUDAGuard guard(L.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int B = (int)labels.numel();
  if (B == 0) return;
  const int stride = (int)L.stride(0);
  TORCH_CHECK(stride % 32 == 0 && stride <= 4096 && L.stride(1) == 1);
  if (stride <= 1024) {       // warp-per-instance fast path
    const int nv = (stride / 4 + 31) / 32;
    const int gridw = std::min((B + 7) / 8, 148 * 8);
#define LAUNCH_SW(NVv)                                                                                  \
    mlr_softmax_bcast_warp_kernel<NVv><<<gridw, 256, 0, stream>>>(                                      \
        L.data_ptr<float>(), stride, labels.data_ptr<int>(), B, (int)num_classes, (float)step,          \
        reinterpret_cast<const long long*>(d_ptrs.data_ptr<int64_t>()), (int)d_ptrs.numel(), (int)rank, \
        stats.data_ptr<float>());
    if (nv <= 1) { LAUNCH_SW(1) } else if (nv <= 2) { LAUNCH_SW(2) } else if (nv <= 4) { LAUNCH_SW(4) } else { LAUNCH_SW(8) }
#undef LAUNCH_SW
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return;
  }
  const int vec = (stride / 4 + SHIP_THREADS - 1) / SHIP_THREADS;
  const int grid = std::min(B, 148 * 16);
#define LAUNCH_SB(V)                                                                                  \
  mlr_softmax_bcast_kernel<V><<<grid, SHIP_THREADS, 0, stream>>>(                                     \
      L.data_ptr<float>(), stride, labels.data_ptr<int>(), B, (int)num_classes, (float)step,          \
      reinterpret_cast<const long long*>(d_ptrs.data_ptr<int64_t>()), (int)d_ptrs.numel(), (int)rank, \
      stride, stats.data_ptr<float>());
  if (vec <= 1) { LAUNCH_SB(1) } else if (vec == 2) { LAUNCH_SB(2) } else { LAUNCH_SB(4) }
#undef LAUNCH_SB
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace hb
