// LDA collapsed Gibbs sampling for one mini-batch of documents (K8), emitting the sparse
// (word, topic, +-1) deltas that the PUSH kernel (table_ops.cu: push_sparse_i32_kernel, K2+K5)
// reduce-adds into the owners' rows.
//
// Reference: dolphin/mlapps/lda/SparseLDASampler.java:141-274 (per-token resampling with the s/r/q
// bucket decomposition -- a CPU optimisation of the SAME conditional
//     p(z = k | .) ~ (n_dk + alpha) (n_wk + beta) / (n_k + V beta) ),
// LDATrainer.java:317-357 (delta table build), :113-194 (initial random assignment pushed as counts).
// GPU formulation: one warp per document; the document-topic histogram lives in shared memory,
// the word-topic rows come from the batch-local staging buffer filled by ONE pull (K1) of the
// batch vocabulary, every lane evaluates K/32 topics, a warp scan picks the new topic.
#include "hb_common.cuh"
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

namespace hb {

static constexpr int LDA_WARPS = 4;
static constexpr int LDA_UNROLL = 8;      // 128-bit row loads in flight per lane

__global__ void __launch_bounds__(LDA_WARPS * 32) lda_sample_kernel(
    const int* __restrict__ doc_ptr, const int* __restrict__ tok_slot, const long long* __restrict__ slot_key,
    int* __restrict__ z, int n_docs, const int* __restrict__ nwk, int nwk_stride, const int* __restrict__ nk,
    int K, float alpha, float beta, float vbeta, unsigned long long seed, long long tok_base,
    long long* __restrict__ out_keys, int* __restrict__ out_topics, int* __restrict__ out_deltas,
    int* __restrict__ sum_delta, int init_mode) {
  extern __shared__ unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int Kp = (K + 3) & ~3;             // per-warp arrays padded to whole 128-bit groups
  const int K4 = Kp >> 2;
  int* ndk = reinterpret_cast<int*>(smem_raw) + (size_t)wid * Kp;
  float* pk = reinterpret_cast<float*>(smem_raw + (size_t)LDA_WARPS * Kp * sizeof(int)) + (size_t)wid * Kp;
  for (int d = blockIdx.x * LDA_WARPS + wid; d < n_docs; d += gridDim.x * LDA_WARPS) {
    const int beg = doc_ptr[d], end = doc_ptr[d + 1];
    for (int k = lane; k < Kp; k += 32) ndk[k] = 0;
    __syncwarp();
    if (!init_mode) {
      for (int t = beg + lane; t < end; t += 32) atomicAdd(&ndk[z[t]], 1);
    }
    __syncwarp();
    // Tokens are processed in chunks of 32: lane l fetches the slot / old topic / key of token cb + l up front (one
    // round of global latency per 32 tokens instead of three dependent loads per token), the per-token loop gets them
    // by shuffle, prefetches the NEXT token's count row into L2 (one 128 B line per lane) and the chunk's outputs are
    // written back coalesced, one token per lane.
    for (int cb = beg; cb < end; cb += 32) {
     const int tl = cb + lane;
     int slot_l = 0, old_l = -1, nw_l = 0;
     long long key_l = 0;
     if (tl < end) {
       slot_l = tok_slot[tl];
       key_l = slot_key[slot_l];
       if (!init_mode) old_l = z[tl];
     }
     const int cnt = min(32, end - cb);
     for (int i = 0; i < cnt; ++i) {
      const int t = cb + i;
      const int slot = __shfl_sync(0xffffffffu, slot_l, i);
      const int* row = nwk + (size_t)slot * nwk_stride;
      if (!init_mode && i + 1 < cnt) {
        const int ns = __shfl_sync(0xffffffffu, slot_l, i + 1);
        if (lane * 32 < nwk_stride)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nwk + (size_t)ns * nwk_stride + lane * 32));
      }
      int old = -1, nw;
      if (init_mode) {
        // initial uniformly random assignment
        nw = (int)(rng_u32(seed, (unsigned long long)(tok_base + t)) % (unsigned)K);
      } else {
        old = __shfl_sync(0xffffffffu, old_l, i);
        if (lane == 0) ndk[old] -= 1;
        __syncwarp();
        // Every lane owns the topic groups g = lane, lane + 32, ... (4 consecutive topics per 128-bit load).  The loads
        // of LDA_UNROLL groups are issued before any of them is consumed: the row comes from L2 / HBM (the batch
        // staging buffer), so memory-level parallelism -- not arithmetic -- bounds the per-token latency.
        float lsum = 0.f;
        const int4* row4 = reinterpret_cast<const int4*>(row);
        const int4* nk4 = reinterpret_cast<const int4*>(nk);
        const int4* ndk4 = reinterpret_cast<const int4*>(ndk);
        float4* pk4 = reinterpret_cast<float4*>(pk);
        for (int g0 = lane; g0 < K4; g0 += 32 * LDA_UNROLL) {
          int4 r[LDA_UNROLL], n[LDA_UNROLL];
#pragma unroll
          for (int u = 0; u < LDA_UNROLL; ++u) {
            const int g = g0 + 32 * u;
            if (g < K4) { r[u] = row4[g]; n[u] = nk4[g]; }
          }
#pragma unroll
          for (int u = 0; u < LDA_UNROLL; ++u) {
            const int g = g0 + 32 * u;
            if (g < K4) {
              const int4 dk = ndk4[g];
              const int k0 = 4 * g;
              const int rr[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
              const int nn[4] = {n[u].x, n[u].y, n[u].z, n[u].w};
              const int dd[4] = {dk.x, dk.y, dk.z, dk.w};
              float pp[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int k = k0 + j;
                const int own = (k == old) ? 1 : 0;
                const float a = (float)dd[j] + alpha;
                const float b = (float)max(rr[j] - own, 0) + beta;
                const float c = (float)max(nn[j] - own, 0) + vbeta;
                pp[j] = (k < K) ? __fdividef(a * b, c) : 0.f;
                lsum += pp[j];
              }
              pk4[g] = make_float4(pp[0], pp[1], pp[2], pp[3]);
            }
          }
        }
        // inclusive scan of the lane sums (lane-major group order)
        float incl = lsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const float total = __shfl_sync(0xffffffffu, incl, 31);
        const float u = rng_uniform(seed, (unsigned long long)(tok_base + t)) * total;
        const unsigned hit = __ballot_sync(0xffffffffu, incl > u);
        const int src_lane = hit ? (__ffs(hit) - 1) : 31;
        int cand = K - 1;
        if (lane == src_lane) {
          float acc = incl - lsum;
          cand = -1;
          bool done = false;
          for (int g = lane; g < K4 && !done; g += 32) {
            const float4 q = pk4[g];
            const float qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (!done) {
                acc += qq[j];
                if (qq[j] > 0.f) cand = 4 * g + j;      // never land on a zero-probability (or padding) topic
                if (acc > u) done = true;
              }
            }
          }
          if (cand < 0) cand = old;
        }
        nw = __shfl_sync(0xffffffffu, cand, src_lane);
      }
      if (lane == 0) ndk[nw] += 1;
      if (lane == i) nw_l = nw;
      __syncwarp();
     }
     if (tl < end) {
       const int changed = (nw_l != old_l) ? 1 : 0;
       z[tl] = nw_l;
       out_keys[2 * (size_t)tl] = key_l;      out_topics[2 * (size_t)tl] = (old_l >= 0) ? old_l : 0;
       out_deltas[2 * (size_t)tl] = (old_l >= 0 && changed) ? -1 : 0;
       out_keys[2 * (size_t)tl + 1] = key_l;  out_topics[2 * (size_t)tl + 1] = nw_l;
       out_deltas[2 * (size_t)tl + 1] = changed;
       if (changed) {
         if (old_l >= 0) atomicAdd(sum_delta + old_l, -1);
         atomicAdd(sum_delta + nw_l, 1);
       }
     }
    }
  }
}

void lda_sample(at::Tensor doc_ptr, at::Tensor tok_slot, at::Tensor slot_key, at::Tensor z, at::Tensor nwk,
                at::Tensor nk, int64_t K, double alpha, double beta, int64_t V, int64_t seed, int64_t tok_base,
                at::Tensor out_keys, at::Tensor out_topics, at::Tensor out_deltas, at::Tensor sum_delta,
                bool init_mode) {
  c10::cuda::CUDAGuard guard(z.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int n_docs = (int)doc_ptr.numel() - 1;
  if (n_docs <= 0) return;
  TORCH_CHECK(doc_ptr.scalar_type() == at::kInt && tok_slot.scalar_type() == at::kInt && z.scalar_type() == at::kInt);
  TORCH_CHECK(slot_key.scalar_type() == at::kLong && out_keys.scalar_type() == at::kLong);
  TORCH_CHECK(nwk.scalar_type() == at::kInt && nk.scalar_type() == at::kInt && nwk.stride(1) == 1);
  const size_t smem = (size_t)LDA_WARPS * ((K + 3) & ~3) * (sizeof(int) + sizeof(float));
  TORCH_CHECK(nwk.stride(0) % 4 == 0 && reinterpret_cast<uintptr_t>(nwk.data_ptr<int>()) % 16 == 0 &&
              reinterpret_cast<uintptr_t>(nk.data_ptr<int>()) % 16 == 0 && nwk.stride(0) >= ((K + 3) & ~3),
              "lda_sample: count rows must be 16-byte aligned and padded to whole int4 groups");
  TORCH_CHECK(smem <= 200 * 1024, "num_topics too large for the shared-memory sampler");
  // per-device attribute: set on every launch (executors of one process may drive different GPUs)
  cudaFuncSetAttribute(lda_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int grid = std::min((n_docs + LDA_WARPS - 1) / LDA_WARPS, hb::sm_count() * 8);
  lda_sample_kernel<<<grid, LDA_WARPS * 32, smem, stream>>>(
      doc_ptr.data_ptr<int>(), tok_slot.data_ptr<int>(),
      reinterpret_cast<const long long*>(slot_key.data_ptr<int64_t>()), z.data_ptr<int>(), n_docs,
      nwk.data_ptr<int>(), (int)nwk.stride(0), nk.data_ptr<int>(), (int)K, (float)alpha, (float)beta,
      (float)(V * beta), (unsigned long long)seed, (long long)tok_base,
      reinterpret_cast<long long*>(out_keys.data_ptr<int64_t>()), out_topics.data_ptr<int>(),
      out_deltas.data_ptr<int>(), sum_delta.data_ptr<int>(), init_mode ? 1 : 0);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace hb
