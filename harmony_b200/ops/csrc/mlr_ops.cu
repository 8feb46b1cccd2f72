// Fused multinomial-logistic-regression kernels for a feature-major sharded model table.
//
// Model table row f = weights of feature f for all C classes (row_stride = C padded to 32).
// With sparse inputs the PULL of the reference (MLRTrainer.pullModel -> multiGetOrInit of
// every key, jobserver/.../mlapps/mlr/MLRTrainer.java:185-187) collapses to "read the rows of
// the features present in the mini-batch", and the PUSH (pushUpdate -> multiUpdate,
// MLRTrainer.java:303-307 + MLRETModelUpdateFunction.java:51-62 `w += delta`) collapses to
// "reduce-add one row per distinct feature of the mini-batch".
//
//   mlr_fwd_kernel  : K1 + K6-forward fused.  One CTA per instance; rows are read straight
//                     from the owner GPU's HBM (local or NVLink peer mapping), accumulated in
//                     registers, followed by the log-sum-exp softmax of MLRTrainer.predict
//                     (:475-505), the loss/accuracy of computeLoss (:404-470) and the
//                     pre-scaled error  D[b,:] = -step * (p - y)  of updateGradient (:374-398).
//   mlr_bwd_push_kernel : K6-backward + K2 + K3 fused.  One CTA per distinct feature:
//                     delta_f = sum_b x_bf * D[b,:]  (D is L2 resident) is reduce-added with
//                     128-bit red.add directly into the owner's row: the update function
//                     `w += delta` executes inside the owner's memory system, no owner kernel.
#include "hb_common.cuh"
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

namespace hb {

static constexpr int MLR_THREADS = 256;
static constexpr float MLR_ZERO_THRESHOLD = 1e-7f;

__device__ __forceinline__ float4 ld_row_hint_f4(const float* r, int c0, int width, uint64_t pol) {
  return (c0 < width) ? ld_hint_f4(r + c0, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// Guarded 128-bit row load: threads beyond the (padded) row width contribute zeros.
__device__ __forceinline__ float4 ld_row_f4(const float* r, int c0, int width) {
  return (c0 < width) ? ld_stream_f4(r + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int VEC, bool WRITE_D>
__global__ void __launch_bounds__(MLR_THREADS) mlr_fwd_kernel(
    TableView W, const int* __restrict__ csr_ptr, const int* __restrict__ csr_col,
    const float* __restrict__ csr_val, const int* __restrict__ labels, int B, int C, float step,
    float* __restrict__ D, int d_stride, float* __restrict__ stats, float lscale) {
  __shared__ const float* s_row[MLR_THREADS];
  __shared__ float s_val[MLR_THREADS];
  __shared__ float s_red[32];
  const int tid = threadIdx.x;
  float cta_loss = 0.f, cta_correct = 0.f;
  const uint64_t pol_keep = l2_policy_evict_last();

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float4 acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int beg = csr_ptr[b], end = csr_ptr[b + 1];
    for (int base = beg; base < end; base += MLR_THREADS) {
      const int cnt = min(MLR_THREADS, end - base);
      __syncthreads();  // previous chunk fully consumed
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
        for (int u = 0; u < U; ++u) {
          const float* r = s_row[j + u];
#pragma unroll
          for (int k = 0; k < VEC; ++k) w[u][k] = ld_row_f4(r, 4 * (tid + k * MLR_THREADS), W.row_stride);
        }
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
        const float* r = s_row[j];
        const float x = s_val[j];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          float4 w = ld_row_f4(r, 4 * (tid + k * MLR_THREADS), W.row_stride);
          acc[k].x = fmaf(x, w.x, acc[k].x); acc[k].y = fmaf(x, w.y, acc[k].y);
          acc[k].z = fmaf(x, w.z, acc[k].z); acc[k].w = fmaf(x, w.w, acc[k].w);
        }
      }
    }
    // ---- softmax over the C valid classes -------------------------------------------------
    float l[VEC][4];
    float tmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c0 = 4 * (tid + k * MLR_THREADS);
      // lazy L2 decay: the table stores v with w = lscale * v (models/mlr.py LazyScale)
      l[k][0] = acc[k].x * lscale; l[k][1] = acc[k].y * lscale; l[k][2] = acc[k].z * lscale; l[k][3] = acc[k].w * lscale;
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
      for (int e = 0; e < 4; ++e) {
        l[k][e] = __expf(l[k][e] - m);  // exp(-inf)=0 for padded classes
        tsum += l[k][e];
      }
    const float denom = block_sum(tsum, s_red);
    const float inv = 1.f / denom;
    const int y = labels[b];
    float my_loss = 0.f, my_hit = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c0 = 4 * (tid + k * MLR_THREADS);
      float d[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = c0 + e;
        float p = fminf(fmaxf(l[k][e] * inv, MLR_ZERO_THRESHOLD), 1.f - MLR_ZERO_THRESHOLD);
        if (c == y) {
          my_loss = -__logf(p);
          // argmax == label  <=>  exp(l_y - m) == 1 (ties resolved in favour of the label)
          my_hit = (l[k][e] >= 1.f) ? 1.f : 0.f;
          p -= 1.f;
        }
        d[e] = (c < C) ? -step * p : 0.f;
      }
      if (WRITE_D && c0 < d_stride)
        st_hint_f4(D + (size_t)b * d_stride + c0, make_float4(d[0], d[1], d[2], d[3]), pol_keep);
    }
    cta_loss += my_loss;      // exactly one thread of the CTA owns column y
    cta_correct += my_hit;
  }
  const float tl = block_sum(cta_loss, s_red);
  const float tc = block_sum(cta_correct, s_red);
  if (tid == 0 && stats != nullptr) {
    atomicAdd(stats + 0, tl);
    atomicAdd(stats + 1, tc);
  }
}

// LOCAL_RMW = false : PUSH with 128-bit red.add (several workers may update the same row concurrently)
// LOCAL_RMW = true  : every row is owned by THIS GPU and touched by exactly one CTA of this launch
//                     (ship mode: the owner folds all workers' error rows), so a plain vector
//                     load-add-store replaces the L2 atomic units
template <int VEC, bool LOCAL_RMW>
__global__ void __launch_bounds__(MLR_THREADS) mlr_bwd_push_kernel(
    TableView W, const int* __restrict__ uniq_feat, const int* __restrict__ feat_ptr,
    const int* __restrict__ samp, const float* __restrict__ xval, int U, const float* __restrict__ D,
    int d_stride, const unsigned* __restrict__ wait_ptr, unsigned wait_target, int* __restrict__ err) {
  const int tid = threadIdx.x;
  if (wait_ptr != nullptr) {       // ship mode: every worker's error rows have landed in D (acquire)
    if (tid == 0) spin_until_ge(wait_ptr, wait_target, err, 3);
    __syncthreads();
  }
  const uint64_t pol_keep = l2_policy_evict_last();    // D is re-read ~nnz/B times: keep it in L2
  const uint64_t pol_stream = l2_policy_evict_first(); // model rows are touched once per step
  for (int u = blockIdx.x; u < U; u += gridDim.x) {
    const int beg = feat_ptr[u], end = feat_ptr[u + 1];
    float* dst = row_ptr<float>(W, (long long)uniq_feat[u]);
    float4 acc[VEC];
    float4 wrow[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      // LOCAL_RMW: issue the read of the model row FIRST so its DRAM latency overlaps the D gathers
      if (LOCAL_RMW) wrow[k] = ld_row_hint_f4(dst, 4 * (tid + k * MLR_THREADS), W.row_stride, pol_stream);
    }
    int j = beg;
    for (; j + 4 <= end; j += 4) {
      float4 d[4][VEC];
      float x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float* drow = D + (size_t)samp[j + q] * d_stride;
        x[q] = xval[j + q];
#pragma unroll
        for (int k = 0; k < VEC; ++k) d[q][k] = ld_row_hint_f4(drow, 4 * (tid + k * MLR_THREADS), d_stride, pol_keep);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          acc[k].x = fmaf(x[q], d[q][k].x, acc[k].x); acc[k].y = fmaf(x[q], d[q][k].y, acc[k].y);
          acc[k].z = fmaf(x[q], d[q][k].z, acc[k].z); acc[k].w = fmaf(x[q], d[q][k].w, acc[k].w);
        }
    }
    for (; j < end; ++j) {
      const float* drow = D + (size_t)samp[j] * d_stride;
      const float x = xval[j];
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float4 d = ld_row_hint_f4(drow, 4 * (tid + k * MLR_THREADS), d_stride, pol_keep);
        acc[k].x = fmaf(x, d.x, acc[k].x); acc[k].y = fmaf(x, d.y, acc[k].y);
        acc[k].z = fmaf(x, d.z, acc[k].z); acc[k].w = fmaf(x, d.w, acc[k].w);
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c0 = 4 * (tid + k * MLR_THREADS);
      if (c0 < W.row_stride) {
        if (LOCAL_RMW) {
          float4 w = wrow[k];
          w.x += acc[k].x; w.y += acc[k].y; w.z += acc[k].z; w.w += acc[k].w;
          st_hint_f4(dst + c0, w, pol_stream);
        } else {
          red_add_f4_sys_hint(dst + c0, acc[k], pol_stream);   // PUSH + `w += delta`
        }
      }
    }
  }
}

// Warp-per-feature variant (row_stride <= 1024 floats): every warp owns one distinct feature, 8 x
// 128-bit chunks per lane, no block barriers => ~3x more rows in flight per SM than the CTA-per-feature
// kernel, which is what this latency-bound gather -> reduce-add chain needs.
template <int NV>
__global__ void __launch_bounds__(256) mlr_bwd_push_warp_kernel(
    TableView W, const int* __restrict__ uniq_feat, const int* __restrict__ feat_ptr,
    const int* __restrict__ samp, const float* __restrict__ xval, int U, const float* __restrict__ D,
    int d_stride, const unsigned* __restrict__ wait_ptr, unsigned wait_target, int* __restrict__ err) {
  if (wait_ptr != nullptr) {       // ship mode: every worker's error rows have landed in D (acquire)
    if (threadIdx.x == 0) spin_until_ge(wait_ptr, wait_target, err, 3);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint64_t pol_keep = l2_policy_evict_last();
  const uint64_t pol_stream = l2_policy_evict_first();
  for (int u = wglobal; u < U; u += nwarps) {
    const int beg = feat_ptr[u], end = feat_ptr[u + 1];
    float* dst = row_ptr<float>(W, (long long)uniq_feat[u]);
    float4 acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = beg;
    for (; j + 2 <= end; j += 2) {
      const float* d0 = D + (size_t)samp[j] * d_stride;
      const float* d1 = D + (size_t)samp[j + 1] * d_stride;
      const float x0 = xval[j], x1 = xval[j + 1];
      float4 a[NV], b[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        a[k] = ld_row_hint_f4(d0, 4 * (lane + 32 * k), d_stride, pol_keep);
        b[k] = ld_row_hint_f4(d1, 4 * (lane + 32 * k), d_stride, pol_keep);
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        acc[k].x = fmaf(x0, a[k].x, fmaf(x1, b[k].x, acc[k].x)); acc[k].y = fmaf(x0, a[k].y, fmaf(x1, b[k].y, acc[k].y));
        acc[k].z = fmaf(x0, a[k].z, fmaf(x1, b[k].z, acc[k].z)); acc[k].w = fmaf(x0, a[k].w, fmaf(x1, b[k].w, acc[k].w));
      }
    }
    if (j < end) {
      const float* d0 = D + (size_t)samp[j] * d_stride;
      const float x0 = xval[j];
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float4 a = ld_row_hint_f4(d0, 4 * (lane + 32 * k), d_stride, pol_keep);
        acc[k].x = fmaf(x0, a.x, acc[k].x); acc[k].y = fmaf(x0, a.y, acc[k].y);
        acc[k].z = fmaf(x0, a.z, acc[k].z); acc[k].w = fmaf(x0, a.w, acc[k].w);
      }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane + 32 * k);
      if (c0 < W.row_stride) red_add_f4_sys_hint(dst + c0, acc[k], pol_stream);
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------
// Force-load the backward kernels (see ship_preload in mlr_nvls_ops.cu: lazy loading vs. spinning kernels).
void mlr_preload_bwd() {
  cudaFuncAttributes a;
#define HB_TOUCH(K) C10_CUDA_CHECK(cudaFuncGetAttributes(&a, K))
  HB_TOUCH(mlr_bwd_push_warp_kernel<1>); HB_TOUCH(mlr_bwd_push_warp_kernel<2>);
  HB_TOUCH(mlr_bwd_push_warp_kernel<4>); HB_TOUCH(mlr_bwd_push_warp_kernel<8>);
  HB_TOUCH((mlr_bwd_push_kernel<1, false>)); HB_TOUCH((mlr_bwd_push_kernel<2, false>));
  HB_TOUCH((mlr_bwd_push_kernel<4, false>));
#undef HB_TOUCH
}

static TableView mlr_view(const at::Tensor& block_ptrs, int64_t rows_per_block, int64_t row_stride,
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

// D: [B, d_stride] float32 scratch (d_stride == row_stride); stats: float32[2] (loss, #correct)
void mlr_forward(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t num_classes,
                 int64_t mode, at::Tensor csr_ptr, at::Tensor csr_col, at::Tensor csr_val, at::Tensor labels,
                 double step, c10::optional<at::Tensor> D, at::Tensor stats, double logit_scale) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView W = mlr_view(block_ptrs, rows_per_block, row_stride, num_classes, mode);
  const int B = (int)labels.numel();
  if (B == 0) return;
  TORCH_CHECK(csr_ptr.scalar_type() == at::kInt && csr_col.scalar_type() == at::kInt &&
              csr_val.scalar_type() == at::kFloat && labels.scalar_type() == at::kInt);
  TORCH_CHECK(row_stride % 32 == 0 && row_stride <= 4 * MLR_THREADS * 4, "row_stride must be <= 4096 and % 32 == 0");
  float* dptr = nullptr;
  int d_stride = (int)row_stride;
  if (D.has_value()) {
    TORCH_CHECK(D->scalar_type() == at::kFloat && D->dim() == 2 && D->size(0) >= B && D->stride(1) == 1);
    TORCH_CHECK(D->stride(0) == row_stride, "D row stride must equal the table row stride");
    dptr = D->data_ptr<float>();
  }
  const int vec = (int)((row_stride / 4 + MLR_THREADS - 1) / MLR_THREADS);
  const int grid = std::min(B, hb::sm_count() * 32);
#define LAUNCH_FWD(V)                                                                                      \
  if (dptr) mlr_fwd_kernel<V, true><<<grid, MLR_THREADS, 0, stream>>>(                                     \
        W, csr_ptr.data_ptr<int>(), csr_col.data_ptr<int>(), csr_val.data_ptr<float>(),                    \
        labels.data_ptr<int>(), B, (int)num_classes, (float)step, dptr, d_stride, stats.data_ptr<float>(), \
        (float)logit_scale);                                                                               \
  else mlr_fwd_kernel<V, false><<<grid, MLR_THREADS, 0, stream>>>(                                         \
        W, csr_ptr.data_ptr<int>(), csr_col.data_ptr<int>(), csr_val.data_ptr<float>(),                    \
        labels.data_ptr<int>(), B, (int)num_classes, (float)step, nullptr, d_stride, stats.data_ptr<float>(), \
        (float)logit_scale);
  if (vec <= 1) { LAUNCH_FWD(1) } else if (vec == 2) { LAUNCH_FWD(2) } else { LAUNCH_FWD(4) }
#undef LAUNCH_FWD
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void mlr_backward_push(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t num_classes,
                       int64_t mode, at::Tensor uniq_feat, at::Tensor feat_ptr, at::Tensor samp,
                       at::Tensor xval, at::Tensor D, bool local_rmw, int64_t wait_ptr, int64_t wait_target,
                       c10::optional<at::Tensor> err) {
  c10::cuda::CUDAGuard guard(block_ptrs.device());
  const unsigned* wp = reinterpret_cast<const unsigned*>(wait_ptr);
  int* ep = err.has_value() ? err->data_ptr<int>() : nullptr;
  TORCH_CHECK(wp == nullptr || ep != nullptr, "a wait needs an error flag");
  auto stream = at::cuda::getCurrentCUDAStream();
  TableView W = mlr_view(block_ptrs, rows_per_block, row_stride, num_classes, mode);
  const int U = (int)uniq_feat.numel();
  if (U == 0) return;
  TORCH_CHECK(uniq_feat.scalar_type() == at::kInt && feat_ptr.scalar_type() == at::kInt &&
              samp.scalar_type() == at::kInt && xval.scalar_type() == at::kFloat);
  TORCH_CHECK(D.scalar_type() == at::kFloat && D.stride(0) == row_stride && D.stride(1) == 1);
  if (!local_rmw && row_stride <= 1024) {     // warp-per-feature fast path
    const int nv = (int)((row_stride / 4 + 31) / 32);
    const int gridw = std::min((U + 7) / 8, hb::sm_count() * 16);
#define LAUNCH_BW(NVv)                                                                             \
    mlr_bwd_push_warp_kernel<NVv><<<gridw, 256, 0, stream>>>(                                      \
        W, uniq_feat.data_ptr<int>(), feat_ptr.data_ptr<int>(), samp.data_ptr<int>(),               \
        xval.data_ptr<float>(), U, D.data_ptr<float>(), (int)row_stride, wp, (unsigned)wait_target, ep);
    if (nv <= 1) { LAUNCH_BW(1) } else if (nv <= 2) { LAUNCH_BW(2) } else if (nv <= 4) { LAUNCH_BW(4) } else { LAUNCH_BW(8) }
#undef LAUNCH_BW
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return;
  }
  const int vec = (int)((row_stride / 4 + MLR_THREADS - 1) / MLR_THREADS);
  const int grid = std::min(U, hb::sm_count() * 8);
#define LAUNCH_BWD(V)                                                                               \
  if (local_rmw) mlr_bwd_push_kernel<V, true><<<grid, MLR_THREADS, 0, stream>>>(                    \
      W, uniq_feat.data_ptr<int>(), feat_ptr.data_ptr<int>(), samp.data_ptr<int>(),                 \
      xval.data_ptr<float>(), U, D.data_ptr<float>(), (int)row_stride, wp, (unsigned)wait_target, ep);                             \
  else mlr_bwd_push_kernel<V, false><<<grid, MLR_THREADS, 0, stream>>>(                             \
      W, uniq_feat.data_ptr<int>(), feat_ptr.data_ptr<int>(), samp.data_ptr<int>(),                 \
      xval.data_ptr<float>(), U, D.data_ptr<float>(), (int)row_stride, wp, (unsigned)wait_target, ep);
  if (vec <= 1) { LAUNCH_BWD(1) } else if (vec == 2) { LAUNCH_BWD(2) } else { LAUNCH_BWD(4) }
#undef LAUNCH_BWD
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace hb
