// NMF mini-batch kernel (K7 + local-model K4 fused): SDDMM + both SpMMs over rank-wide vectors.
//
// Reference: dolphin/mlapps/nmf/NMFTrainer.java:328-367 -- per non-zero (i, j, v):
//   e = L_i . R_j - v ;  gL_i += 2 e R_j (+ 2 lambda L_i once per row) ;  gR_j += 2 e L_i
// and NMFETModelUpdateFunction.java:43-52 for the local model: L_i = clamp(L_i - step gL_i, 0, max).
//
// One warp owns one row i of the mini-batch: L_i stays in registers (rank/32 floats per lane),
// R_j comes from the batch-local staging buffer filled by ONE pull of the needed columns (the
// global model R is tiny -- rank x cols -- so each column crosses NVLink once per batch, not once
// per non-zero), gR_j is reduce-added into an L2-resident staging gradient which the PUSH kernel
// then applies at the owners with the SGD+clamp update function.
#include "hb_common.cuh"
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

namespace hb {

template <int VPL>   // float4 chunks per lane (rank <= 128 * VPL); rows are padded to 32 floats
__global__ void __launch_bounds__(256) nmf_batch_kernel(
    const int* __restrict__ csr_ptr, const int* __restrict__ col_slot, const float* __restrict__ val,
    int n_rows, int rank, float* __restrict__ L, int l_stride, const int* __restrict__ l_row,
    const float* __restrict__ Rst, float* __restrict__ Gst, int st_stride, float step, float lambda,
    float lo, float hi, float* __restrict__ loss_out) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = (rank + 3) >> 2;          // float4 chunks that carry data
  float loss = 0.f;
  for (int i = warp; i < n_rows; i += nwarps) {
    float* Li = L + (size_t)l_row[i] * l_stride;
    float4 l[VPL], g[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = lane + 32 * k;
      l[k] = (v < nvec) ? *reinterpret_cast<const float4*>(Li + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      g[k] = make_float4(2.f * lambda * l[k].x, 2.f * lambda * l[k].y, 2.f * lambda * l[k].z, 2.f * lambda * l[k].w);
    }
    const int beg = csr_ptr[i], end = csr_ptr[i + 1];
    for (int p = beg; p < end; ++p) {
      const int s = col_slot[p];
      const float* Rj = Rst + (size_t)s * st_stride;
      float4 r[VPL];
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int v = lane + 32 * k;
        r[k] = (v < nvec) ? *reinterpret_cast<const float4*>(Rj + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
        dot = fmaf(l[k].x, r[k].x, dot); dot = fmaf(l[k].y, r[k].y, dot);
        dot = fmaf(l[k].z, r[k].z, dot); dot = fmaf(l[k].w, r[k].w, dot);
      }
      dot = warp_sum(dot);
      const float e = dot - val[p];
      loss += e * e;
      const float e2 = 2.f * e;
      float* Gj = Gst + (size_t)s * st_stride;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int v = lane + 32 * k;
        g[k].x = fmaf(e2, r[k].x, g[k].x); g[k].y = fmaf(e2, r[k].y, g[k].y);
        g[k].z = fmaf(e2, r[k].z, g[k].z); g[k].w = fmaf(e2, r[k].w, g[k].w);
        // 128-bit RED into the L2-resident staging gradient (4x fewer atomic ops than scalar adds)
        if (v < nvec) red_add_f4_sys(Gj + 4 * v, make_float4(e2 * l[k].x, e2 * l[k].y, e2 * l[k].z, e2 * l[k].w));
      }
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = lane + 32 * k;
      if (v < nvec) {                          // local-model update (K4); padded lanes stay zero
        float4 o;
        o.x = fminf(fmaxf(l[k].x - step * g[k].x, lo), hi); o.y = fminf(fmaxf(l[k].y - step * g[k].y, lo), hi);
        o.z = fminf(fmaxf(l[k].z - step * g[k].z, lo), hi); o.w = fminf(fmaxf(l[k].w - step * g[k].w, lo), hi);
        const int c = 4 * v;
        if (c + 3 < rank) *reinterpret_cast<float4*>(Li + c) = o;
        else { float ov[4] = {o.x, o.y, o.z, o.w}; for (int q = 0; q < 4; ++q) if (c + q < rank) Li[c + q] = ov[q]; }
      }
    }
  }
  if (lane == 0 && loss_out != nullptr && loss != 0.f) atomicAdd(loss_out, loss);
}

// G[s, :] += 2 lambda R[s, :]  for the batch's distinct columns ("+2 lambda R_j once")
__global__ void nmf_reg_kernel(float* __restrict__ G, const float* __restrict__ R, long long n, float two_lambda) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) G[i] = fmaf(two_lambda, R[i], G[i]);
}

void nmf_batch(at::Tensor csr_ptr, at::Tensor col_slot, at::Tensor val, at::Tensor L, at::Tensor l_row,
               at::Tensor Rst, at::Tensor Gst, int64_t rank, double step, double lambda, double lo, double hi,
               at::Tensor loss) {
  c10::cuda::CUDAGuard guard(L.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int n_rows = (int)l_row.numel();
  if (n_rows == 0) return;
  TORCH_CHECK(csr_ptr.scalar_type() == at::kInt && col_slot.scalar_type() == at::kInt && l_row.scalar_type() == at::kInt);
  TORCH_CHECK(L.scalar_type() == at::kFloat && Rst.scalar_type() == at::kFloat && Gst.scalar_type() == at::kFloat);
  TORCH_CHECK(Rst.stride(0) == Gst.stride(0) && Rst.stride(1) == 1 && L.stride(1) == 1);
  TORCH_CHECK(rank <= 1024, "rank <= 1024 supported");
  TORCH_CHECK(L.stride(0) % 4 == 0 && Rst.stride(0) % 4 == 0, "row strides must be multiples of 4 floats");
  TORCH_CHECK(L.stride(0) >= ((rank + 3) / 4) * 4, "L rows must be padded to a multiple of 4 floats");
  const int rpl = (int)((rank + 127) / 128);
  const int grid = std::min((n_rows + 7) / 8, hb::sm_count() * 8);
#define LAUNCH_NMF(R)                                                                                   \
  nmf_batch_kernel<R><<<grid, 256, 0, stream>>>(                                                        \
      csr_ptr.data_ptr<int>(), col_slot.data_ptr<int>(), val.data_ptr<float>(), n_rows, (int)rank,      \
      L.data_ptr<float>(), (int)L.stride(0), l_row.data_ptr<int>(), Rst.data_ptr<float>(),              \
      Gst.data_ptr<float>(), (int)Rst.stride(0), (float)step, (float)lambda, (float)lo, (float)hi,      \
      loss.data_ptr<float>());
  if (rpl <= 1) { LAUNCH_NMF(1) } else if (rpl <= 2) { LAUNCH_NMF(2) } else if (rpl <= 4) { LAUNCH_NMF(4) }
  else { LAUNCH_NMF(8) }
#undef LAUNCH_NMF
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  if (lambda != 0.0) {
    long long n = Gst.numel();
    nmf_reg_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(Gst.data_ptr<float>(), Rst.data_ptr<float>(), n,
                                                              (float)(2.0 * lambda));
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
}

}  // namespace hb
