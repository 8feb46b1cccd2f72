// K12 / K9 kernels: Lasso coordinate descent fused with the model pull/push, gradient-boosted-tree split
// finding / prediction, and the log-gamma reductions of the LDA likelihood.
//
// Reference parity (behaviour, not code):
//   * LassoTrainer.java:170-224 (cyclic coordinate descent with soft-threshold, weights ~0 skipped),
//     :292-322 (push step*(new-old) per model partition)
//   * GBTTrainer.java:327-531 (exact greedy split finding over sorted feature values / categories ordered by
//     mean gradient, gain = GL^2/(2nL+lambda) + GR^2/(2nR+lambda) - G^2/(2n+lambda)), :244-320 (residuals)
//   * LDAStatCalculator.java:103-184 (sums of logGamma over the word-topic / doc-topic counts)
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "hb_common.cuh"

namespace hb {

// ------------------------------------------------------------------------------------------
// Lasso: one CTA runs exact cyclic CD on a dense mini-batch.  The prediction vector lives in shared memory, a
// coordinate step is ONE pass over the feature's column (norm and correlation reduced together); the model is
// read through the pointer table at the start (pull) and the scaled delta is reduce-added into the owners' rows
// at the end (push) -- no host round trip between pull, compute and push.
// ------------------------------------------------------------------------------------------
constexpr int LASSO_THREADS = 1024;

__device__ __forceinline__ double2 block_sum2(double a, double b, double* red) {
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();                       // red[] may still be read from the previous call
  if (l == 0) { red[2 * w] = a; red[2 * w + 1] = b; }
  __syncthreads();
  a = l < nw ? red[2 * l] : 0.0;
  b = l < nw ? red[2 * l + 1] : 0.0;
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  return make_double2(a, b);
}

__global__ void __launch_bounds__(LASSO_THREADS, 1)
lasso_cd_kernel(TableView model, const float* __restrict__ Xt, const float* __restrict__ y, int B, int F, int fpp,
                float lambda, float step, float* __restrict__ w_new, float* __restrict__ w_old,
                float* __restrict__ stats) {
  extern __shared__ float pre[];          // [B] current prediction X w
  __shared__ double red[2 * (LASSO_THREADS / 32)];
  const int t = threadIdx.x;
  // pull: w_new[j] = model[j / fpp][j % fpp]
  for (int j = t; j < F; j += LASSO_THREADS) w_old[j] = w_new[j] = row_ptr<float>(model, j / fpp)[j % fpp];
  __syncthreads();
  for (int b = t; b < B; b += LASSO_THREADS) {
    float acc = 0.f;
    for (int j = 0; j < F; ++j) acc = fmaf(Xt[(size_t)j * B + b], w_new[j], acc);
    pre[b] = acc;
  }
  __syncthreads();
  for (int j = 0; j < F; ++j) {
    const float wj = w_new[j];
    if (fabsf(wj) < 1e-9f) continue;      // uniform across the CTA
    const float* col = Xt + (size_t)j * B;
    double nrm = 0.0, dot = 0.0;
    for (int b = t; b < B; b += LASSO_THREADS) {
      float c = col[b];
      float p = pre[b] - c * wj;
      nrm += (double)c * c;
      dot += (double)c * (y[b] - p);
    }
    double2 r = block_sum2(nrm, dot, red);
    if (r.x < 1e-9) continue;
    float x = (float)(r.y / r.x), th = lambda / (float)r.x;
    float nw = fabsf(x) <= th ? 0.f : (x >= 0.f ? x - th : x + th);
    for (int b = t; b < B; b += LASSO_THREADS) pre[b] += col[b] * (nw - wj);
    if (t == 0) w_new[j] = nw;
    __syncthreads();
  }
  // push: model += step * (new - old); loss of the batch for the summary
  for (int j = t; j < F; j += LASSO_THREADS) {
    float* p = row_ptr<float>(model, j / fpp) + (j % fpp);
    float d = step * (w_new[j] - w_old[j]);
    if (d != 0.f) red_add_f32_sys(p, d);
  }
  double se = 0.0;
  for (int b = t; b < B; b += LASSO_THREADS) { double e = (double)pre[b] - y[b]; se += e * e; }
  double2 s = block_sum2(se, 0.0, red);
  if (t == 0) stats[0] = (float)(s.x / B);
}

void lasso_cd(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim, int64_t mode,
              at::Tensor Xt, at::Tensor y, double lambda, double step, at::Tensor w_new, at::Tensor w_old,
              at::Tensor stats) {
  c10::cuda::CUDAGuard guard(Xt.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TORCH_CHECK(Xt.is_cuda() && Xt.scalar_type() == at::kFloat && Xt.dim() == 2 && Xt.is_contiguous());
  int F = (int)Xt.size(0), B = (int)Xt.size(1);
  TORCH_CHECK(y.numel() == B && y.scalar_type() == at::kFloat && w_new.numel() == F && w_old.numel() == F &&
              stats.numel() >= 1);
  TORCH_CHECK(B * 4 <= 200 * 1024, "lasso mini-batch too large for the shared-memory prediction vector");
  TableView v;
  v.block_ptrs = reinterpret_cast<const long long*>(block_ptrs.data_ptr<int64_t>());
  v.num_blocks = (int)block_ptrs.numel() - 2; /* two guard words follow the pointers */ v.rows_per_block = (int)rows_per_block; v.row_stride = (int)row_stride;
  v.row_dim = (int)row_dim; v.mode = (int)mode;
  size_t smem = (size_t)B * 4;
  C10_CUDA_CHECK(cudaFuncSetAttribute(lasso_cd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lasso_cd_kernel<<<1, LASSO_THREADS, smem, stream>>>(v, Xt.data_ptr<float>(), y.data_ptr<float>(), B, F,
                                                      (int)row_dim, (float)lambda, (float)step,
                                                      w_new.data_ptr<float>(), w_old.data_ptr<float>(),
                                                      stats.data_ptr<float>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// GBT: level-wise exact greedy split search.  One warp per (feature, active node) walks the feature's sorted
// sample order, keeps the node members with a ballot, and turns the members' gradients into running prefix sums
// with a warp scan; a split candidate sits between two consecutive members with different values.
// ------------------------------------------------------------------------------------------
constexpr int GBT_MAX_CAT = 63;       // bit 63 of the category mask flags a categorical split

__global__ void gbt_node_stats_kernel(const int* __restrict__ node_of, const double* __restrict__ g, int n,
                                      int node_base, int num_nodes, double* __restrict__ node_G,
                                      int* __restrict__ node_n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int k = node_of[i] - node_base;
  if (k < 0 || k >= num_nodes) return;
  atomicAdd(&node_G[k], g[i]);
  atomicAdd(&node_n[k], 1);
}

__device__ __forceinline__ double gbt_gain(double gl, double gr, double nl, double nr, double lam, double total) {
  return gl * gl / (2.0 * nl + lam) + gr * gr / (2.0 * nr + lam) - total;
}

// grid (F, num_nodes), block 32.  out_*[node * F + f]
__global__ void __launch_bounds__(32)
gbt_split_kernel(const float* __restrict__ X, int n, int F, const int* __restrict__ order,
                 const int* __restrict__ ftype, const int* __restrict__ node_of, const double* __restrict__ g,
                 int node_base, const double* __restrict__ node_G, const int* __restrict__ node_n, int leaf_min,
                 double lam, double gamma, double* __restrict__ out_gain, float* __restrict__ out_thr,
                 unsigned long long* __restrict__ out_mask) {
  const int f = blockIdx.x, k = blockIdx.y, lane = threadIdx.x;
  const int node = node_base + k;
  const int nn = node_n[k];
  const double G = node_G[k];
  double best = gamma; float best_thr = 0.f; unsigned long long best_mask = 0ull; bool found = false;
  const size_t o = (size_t)k * F + f;
  if (nn < 2 * leaf_min) {
    if (lane == 0) { out_gain[o] = gamma; out_thr[o] = 0.f; out_mask[o] = 0ull; }
    return;
  }
  const double total = G * G / (2.0 * nn + lam);
  if (ftype[f] == 0) {
    const int* ord = order + (size_t)f * n;
    double carry_g = 0.0; int carry_n = 0; float carry_v = 0.f;     // over the members seen so far
    for (int base = 0; base < n; base += 32) {
      int i = base + lane;
      int s = i < n ? ord[i] : -1;
      bool mem = s >= 0 && node_of[s] == node;
      float v = mem ? X[(size_t)s * F + f] : 0.f;
      double gi = mem ? g[s] : 0.0;
      unsigned bal = __ballot_sync(0xffffffffu, mem);
      // exclusive prefix over the members of this chunk
      double ps = gi;
      for (int d = 1; d < 32; d <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, ps, d);
        if (lane >= d) ps += t;
      }
      double gl = carry_g + ps - gi;
      int nl = carry_n + __popc(bal & ((1u << lane) - 1u));
      unsigned lower = bal & ((1u << lane) - 1u);
      int pl = lower ? 31 - __clz(lower) : -1;
      float pv_in = __shfl_sync(0xffffffffu, v, pl < 0 ? 0 : pl);
      float pv = pl < 0 ? carry_v : pv_in;
      if (mem && nl >= 1 && nl >= leaf_min && nn - nl >= leaf_min && pv != v) {
        double gain = gbt_gain(gl, G - gl, nl, nn - nl, lam, total);
        if (gain > best) { best = gain; best_thr = 0.5f * (pv + v); found = true; }   // earlier i wins ties (strict >)
      }
      if (bal) {
        int last = 31 - __clz(bal);
        carry_g += __shfl_sync(0xffffffffu, ps, 31);
        carry_n += __popc(bal);
        carry_v = __shfl_sync(0xffffffffu, v, last);
      }
    }
    // warp arg-max; ties -> the candidate met first in scan order.  Within a lane candidates are met in order and
    // only a strictly larger gain replaces, but across lanes order interleaves by chunk; break ties by threshold
    // position = smaller threshold (sorted order is ascending in value).
    for (int d = 16; d > 0; d >>= 1) {
      double ob = __shfl_xor_sync(0xffffffffu, best, d);
      float ot = __shfl_xor_sync(0xffffffffu, best_thr, d);
      int of = __shfl_xor_sync(0xffffffffu, (int)found, d);
      bool take = of && (!found || ob > best || (ob == best && ot < best_thr));
      if (take) { best = ob; best_thr = ot; found = true; }
    }
    if (lane == 0) { out_gain[o] = found ? best : gamma; out_thr[o] = best_thr; out_mask[o] = 0ull; }
  } else {
    __shared__ double cg[GBT_MAX_CAT];
    __shared__ int cn[GBT_MAX_CAT];
    __shared__ int cid[GBT_MAX_CAT];
    for (int c = lane; c < GBT_MAX_CAT; c += 32) { cg[c] = 0.0; cn[c] = 0; }
    __syncwarp();
    for (int s = lane; s < n; s += 32) {
      if (node_of[s] != node) continue;
      int c = (int)X[(size_t)s * F + f];
      c = c < 0 ? 0 : (c >= GBT_MAX_CAT ? GBT_MAX_CAT - 1 : c);
      atomicAdd(&cg[c], g[s]);
      atomicAdd(&cn[c], 1);
    }
    __syncwarp();
    if (lane == 0) {
      int m = 0;
      for (int c = 0; c < GBT_MAX_CAT; ++c) if (cn[c] > 0) cid[m++] = c;
      for (int a = 1; a < m; ++a) {            // insertion sort by (mean gradient, category)
        int c = cid[a]; double mc = cg[c] / cn[c]; int b = a - 1;
        while (b >= 0) {
          int cb = cid[b]; double mb = cg[cb] / cn[cb];
          if (mb > mc || (mb == mc && cb > c)) { cid[b + 1] = cb; --b; } else break;
        }
        cid[b + 1] = c;
      }
      double gl = 0.0; int nl = 0; unsigned long long mask = 0ull;
      for (int a = 0; a + 1 < m; ++a) {
        int c = cid[a];
        gl += cg[c]; nl += cn[c]; mask |= 1ull << c;
        if (nl < leaf_min || nn - nl < leaf_min) continue;
        double gain = gbt_gain(gl, G - gl, nl, nn - nl, lam, total);
        if (gain > best) { best = gain; best_mask = mask; found = true; }
      }
      out_gain[o] = found ? best : gamma; out_thr[o] = 0.f; out_mask[o] = found ? best_mask : 0ull;
    }
  }
}

// one warp per node: arg-max over features (lowest feature wins ties), write the tree arrays, or make a leaf.
// tree arrays are indexed by the complete-tree node id.
__global__ void __launch_bounds__(32)
gbt_pick_kernel(int F, int node_base, int num_nodes, const double* __restrict__ gains,
                const float* __restrict__ thrs, const unsigned long long* __restrict__ masks,
                const double* __restrict__ node_G, const int* __restrict__ node_n, const int* __restrict__ ftype,
                double gamma, double lam, double step, int allow_split, int* __restrict__ t_feature,
                float* __restrict__ t_thr, unsigned long long* __restrict__ t_mask, float* __restrict__ t_value,
                int* __restrict__ t_leaf) {
  const int k = blockIdx.x, lane = threadIdx.x, node = node_base + k;
  if (node_n[k] == 0) return;                    // node absent at this level
  double best = gamma; int bf = -1;
  if (allow_split) {
    for (int f = lane; f < F; f += 32) {
      double gn = gains[(size_t)k * F + f];
      if (gn > best) { best = gn; bf = f; }
    }
    for (int d = 16; d > 0; d >>= 1) {
      double ob = __shfl_xor_sync(0xffffffffu, best, d);
      int of = __shfl_xor_sync(0xffffffffu, bf, d);
      if (of >= 0 && (bf < 0 || ob > best || (ob == best && of < bf))) { best = ob; bf = of; }
    }
  }
  if (lane == 0) {
    if (bf < 0) {
      t_leaf[node] = 1; t_feature[node] = -1;
      t_value[node] = (float)(step * node_G[k] / ((double)node_n[k] + lam));
    } else {
      t_leaf[node] = 0; t_feature[node] = bf;
      t_thr[node] = thrs[(size_t)k * F + bf];
      t_mask[node] = ftype[bf] ? (masks[(size_t)k * F + bf] | (1ull << 63)) : 0ull;   // bit 63 = categorical flag
    }
  }
}

// route every sample of a split node to its child
__global__ void gbt_route_kernel(const float* __restrict__ X, int n, int F, int* __restrict__ node_of,
                                 int node_base, int num_nodes, const int* __restrict__ t_feature,
                                 const float* __restrict__ t_thr, const unsigned long long* __restrict__ t_mask,
                                 const int* __restrict__ t_leaf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int node = node_of[i];
  if (node < node_base || node >= node_base + num_nodes) return;
  if (t_leaf[node]) { node_of[i] = -1 - node; return; }       // parked at a finished leaf
  int f = t_feature[node];
  float v = X[(size_t)i * F + f];
  unsigned long long m = t_mask[node];
  bool left;
  if (m >> 63) { int c = (int)v; left = c >= 0 && c < 63 && ((m >> c) & 1ull); }
  else left = v <= t_thr[node];
  node_of[i] = left ? 2 * node + 1 : 2 * node + 2;
}

// forest prediction: out[i] += sum over trees of the reached leaf value.  Trees are [T, nodes] arrays.
__global__ void gbt_predict_kernel(const float* __restrict__ X, int n, int F, int T, int nodes,
                                   const int* __restrict__ t_feature, const float* __restrict__ t_thr,
                                   const unsigned long long* __restrict__ t_mask,
                                   const float* __restrict__ t_value, const int* __restrict__ t_leaf,
                                   double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* x = X + (size_t)i * F;
  double acc = 0.0;
  for (int t = 0; t < T; ++t) {
    const size_t o = (size_t)t * nodes;
    int node = 0;
    while (node < nodes && !t_leaf[o + node]) {
      int f = t_feature[o + node];
      if (f < 0) { node = nodes; break; }       // absent node -> contributes 0
      unsigned long long m = t_mask[o + node];
      bool left;
      if (m >> 63) { int c = (int)x[f]; left = c >= 0 && c < 63 && ((m >> c) & 1ull); }
      else left = x[f] <= t_thr[o + node];
      node = left ? 2 * node + 1 : 2 * node + 2;
    }
    if (node < nodes) acc += (double)t_value[o + node];
  }
  out[i] += acc;
}

#define U64P(t) reinterpret_cast<unsigned long long*>((t).data_ptr<int64_t>())

// Build one level of the tree: nodes [node_base, node_base + num_nodes).  `scratch_*` are [num_nodes * F].
void gbt_level(at::Tensor X, at::Tensor order, at::Tensor ftype, at::Tensor node_of, at::Tensor g,
               int64_t node_base, int64_t num_nodes, int64_t leaf_min, double lam, double gamma, double step,
               bool allow_split, at::Tensor t_feature, at::Tensor t_thr, at::Tensor t_mask, at::Tensor t_value,
               at::Tensor t_leaf) {
  c10::cuda::CUDAGuard guard(X.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TORCH_CHECK(X.is_cuda() && X.scalar_type() == at::kFloat && X.dim() == 2 && X.is_contiguous());
  int n = (int)X.size(0), F = (int)X.size(1);
  TORCH_CHECK(order.scalar_type() == at::kInt && order.numel() == (int64_t)n * F && order.is_contiguous());
  TORCH_CHECK(g.scalar_type() == at::kDouble && g.numel() == n && node_of.scalar_type() == at::kInt);
  TORCH_CHECK(t_mask.scalar_type() == at::kLong && t_leaf.scalar_type() == at::kInt);
  auto optd = X.options().dtype(at::kDouble);
  auto opti = X.options().dtype(at::kInt);
  at::Tensor node_G = at::zeros({num_nodes}, optd), node_n = at::zeros({num_nodes}, opti);
  at::Tensor gains = at::empty({num_nodes * F}, optd), thrs = at::empty({num_nodes * F}, X.options());
  at::Tensor masks = at::empty({num_nodes * F}, X.options().dtype(at::kLong));
  gbt_node_stats_kernel<<<(n + 255) / 256, 256, 0, stream>>>(node_of.data_ptr<int>(), g.data_ptr<double>(), n,
                                                              (int)node_base, (int)num_nodes,
                                                              node_G.data_ptr<double>(), node_n.data_ptr<int>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  if (allow_split) {
    dim3 grid(F, (unsigned)num_nodes);
    gbt_split_kernel<<<grid, 32, 0, stream>>>(X.data_ptr<float>(), n, F, order.data_ptr<int>(),
                                               ftype.data_ptr<int>(), node_of.data_ptr<int>(), g.data_ptr<double>(),
                                               (int)node_base, node_G.data_ptr<double>(), node_n.data_ptr<int>(),
                                               (int)leaf_min, lam, gamma, gains.data_ptr<double>(),
                                               thrs.data_ptr<float>(), U64P(masks));
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
  gbt_pick_kernel<<<(unsigned)num_nodes, 32, 0, stream>>>(
      F, (int)node_base, (int)num_nodes, gains.data_ptr<double>(), thrs.data_ptr<float>(), U64P(masks),
      node_G.data_ptr<double>(), node_n.data_ptr<int>(), ftype.data_ptr<int>(), gamma, lam, step,
      allow_split ? 1 : 0, t_feature.data_ptr<int>(), t_thr.data_ptr<float>(), U64P(t_mask),
      t_value.data_ptr<float>(), t_leaf.data_ptr<int>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  gbt_route_kernel<<<(n + 255) / 256, 256, 0, stream>>>(X.data_ptr<float>(), n, F, node_of.data_ptr<int>(),
                                                         (int)node_base, (int)num_nodes, t_feature.data_ptr<int>(),
                                                         t_thr.data_ptr<float>(), U64P(t_mask),
                                                         t_leaf.data_ptr<int>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void gbt_predict(at::Tensor X, at::Tensor t_feature, at::Tensor t_thr, at::Tensor t_mask, at::Tensor t_value,
                 at::Tensor t_leaf, at::Tensor out) {
  c10::cuda::CUDAGuard guard(X.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TORCH_CHECK(X.is_cuda() && X.scalar_type() == at::kFloat && X.dim() == 2 && X.is_contiguous());
  int n = (int)X.size(0), F = (int)X.size(1);
  if (t_feature.numel() == 0 || n == 0) return;
  TORCH_CHECK(t_feature.dim() == 2 && out.scalar_type() == at::kDouble && out.numel() == n);
  int T = (int)t_feature.size(0), nodes = (int)t_feature.size(1);
  gbt_predict_kernel<<<(n + 127) / 128, 128, 0, stream>>>(X.data_ptr<float>(), n, F, T, nodes,
                                                           t_feature.data_ptr<int>(), t_thr.data_ptr<float>(),
                                                           U64P(t_mask), t_value.data_ptr<float>(),
                                                           t_leaf.data_ptr<int>(), out.data_ptr<double>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// K9: sum over a strided int32 matrix of lgamma(max(x, 0) + c)  (LDA word / doc log-likelihood terms)
// ------------------------------------------------------------------------------------------
__global__ void lgamma_sum_kernel(const int* __restrict__ x, long long rows, int cols, int stride, double c,
                                  double* __restrict__ out) {
  double acc = 0.0;
  long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i / cols; int k = (int)(i - r * cols);
    int v = x[r * stride + k];
    acc += lgamma((double)(v > 0 ? v : 0) + c);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

double lgamma_sum_i32(at::Tensor x, int64_t cols, double c) {
  c10::cuda::CUDAGuard guard(x.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kInt && x.dim() == 2 && x.stride(1) == 1 && cols <= x.size(1));
  at::Tensor out = at::zeros({1}, x.options().dtype(at::kDouble));
  long long rows = x.size(0);
  if (rows * cols == 0) return 0.0;
  int grid = (int)std::min<long long>((rows * cols + 255) / 256, hb::sm_count() * 8);
  lgamma_sum_kernel<<<grid, 256, 0, stream>>>(x.data_ptr<int>(), rows, (int)cols, (int)x.stride(0), c,
                                              out.data_ptr<double>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out.item<double>();
}

}  // namespace hb
