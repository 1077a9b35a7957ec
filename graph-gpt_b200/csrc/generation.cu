// Un-masking generation (dLLM-style) — the per-step work that follows the forward pass:
//   gen_sample   : per (sample, position, feature) entry: temperature, top-p, top-k filtering, fp32 softmax, token choice
//                  (argmax or inverse-CDF sampling) and the confidence score (p[x0] | top1 - top2 | -entropy)
//   gen_unmask_* : which masked entries are revealed this step ("origin": Bernoulli(p_transfer); confidence algorithms:
//                  the k_b most confident masked entries of each sample, optional Gumbel perturbation)
// ref: src/utils/generation_utils.py:22-82 (top_p_logits, top_k_logits, sample_tokens), :138-228
// (_batch_unmask_without_for_loop).  HBM-bound: the logits [R, V] fp32 are read exactly once.
#include <float.h>

#include <algorithm>

#include "common.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

// order-preserving float -> uint32 key (larger float <=> larger key); -0.0 < +0.0 is harmless here
__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Cooperative group of G threads (a warp, or a whole CTA of 256) with sum / max / argmax reductions that return the
// result to every thread.  `scratch` is per-CTA shared memory (only used when G > 32).
template <int G>
struct Grp {
  float* scratch;
  __device__ __forceinline__ int tid() const { return G == 32 ? (threadIdx.x & 31) : threadIdx.x; }
  __device__ __forceinline__ void sync() const {
    if (G == 32) __syncwarp(); else __syncthreads();
  }
  __device__ __forceinline__ float sum(float v) const {
    v = warp_sum(v);
    if (G == 32) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < G / 32; ++i) t += scratch[i];
    return t;
  }
  __device__ __forceinline__ float max(float v) const {
    v = warp_max(v);
    if (G == 32) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = scratch[0];
#pragma unroll
    for (int i = 1; i < G / 32; ++i) t = fmaxf(t, scratch[i]);
    return t;
  }
  __device__ __forceinline__ int isum(int v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (G == 32) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) reinterpret_cast<int*>(scratch)[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int i = 0; i < G / 32; ++i) t += reinterpret_cast<int*>(scratch)[i];
    return t;
  }
  // smallest index attaining the maximum of (v, idx) pairs
  __device__ __forceinline__ void argmax(float& v, int& idx) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (G == 32) return;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      scratch[threadIdx.x >> 5] = v;
      reinterpret_cast<int*>(scratch)[8 + (threadIdx.x >> 5)] = idx;
    }
    __syncthreads();
    v = scratch[0];
    idx = reinterpret_cast<int*>(scratch)[8];
#pragma unroll
    for (int i = 1; i < G / 32; ++i) {
      const float ov = scratch[i];
      const int oi = reinterpret_cast<int*>(scratch)[8 + i];
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
  }
};

constexpr float kFiniteMin = -FLT_MAX;   // torch.finfo(torch.float32).min

// One group (warp or CTA) per logits row; the row lives in shared memory.
template <int G>
__global__ void __launch_bounds__(256) gen_sample_kernel(const float* __restrict__ logits, long long ldl, long long R, int V,
                                                         float temperature, int top_k, float top_p,
                                                         const float* __restrict__ u, int conf_mode,
                                                         long long* __restrict__ x0_out, float* __restrict__ conf_out,
                                                         float* __restrict__ probs_out, long long ldp) {
  extern __shared__ float smem[];
  constexpr int kGroups = 256 / G;
  Grp<G> g;
  g.scratch = smem;                                        // 16 floats of reduction scratch (CTA mode)
  float* row = smem + 16 + static_cast<size_t>(G == 32 ? (threadIdx.x >> 5) : 0) * V;
  const int t = g.tid();
  const long long group0 = static_cast<long long>(blockIdx.x) * kGroups + (G == 32 ? (threadIdx.x >> 5) : 0);
  const long long stride = static_cast<long long>(gridDim.x) * kGroups;
  for (long long r = group0; r < R; r += stride) {
    const float* src = logits + r * ldl;
    for (int j = t; j < V; j += G) row[j] = temperature > 0.f ? src[j] / temperature : src[j];
    g.sync();
    // ---- top-p: drop token j when the probability mass of the strictly larger logits exceeds top_p ------------
    if (top_p > 0.f && top_p < 1.f) {
      float m = kFiniteMin;
      for (int j = t; j < V; j += G) m = fmaxf(m, row[j]);
      m = g.max(m);
      float z = 0.f;
      for (int j = t; j < V; j += G) z += expf(row[j] - m);
      z = g.sum(z);
      const float budget = top_p * z;
      uint32_t lo = 0u, hi = 0xffffffffu;                   // smallest key K with mass(key > K) <= budget
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        float above = 0.f;
        for (int j = t; j < V; j += G) above += (fkey(row[j]) > mid) ? expf(row[j] - m) : 0.f;
        above = g.sum(above);
        if (above <= budget) hi = mid; else lo = mid + 1u;
      }
      g.sync();
      for (int j = t; j < V; j += G)
        if (fkey(row[j]) < lo) row[j] = kFiniteMin;
      g.sync();
    }
    // ---- top-k: drop logits below the k-th largest -------------------------------------------------------------
    if (top_k > 0 && top_k < V) {
      uint32_t lo = 0u, hi = 0xffffffffu;                   // largest key K with count(key >= K) >= k
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1) + 1u;
        int cnt = 0;
        for (int j = t; j < V; j += G) cnt += (fkey(row[j]) >= mid);
        cnt = g.isum(cnt);
        if (cnt >= top_k) lo = mid; else hi = mid - 1u;
      }
      g.sync();
      for (int j = t; j < V; j += G)
        if (fkey(row[j]) < lo) row[j] = kFiniteMin;
      g.sync();
    }
    // ---- softmax statistics + arg max ---------------------------------------------------------------------------
    float m = kFiniteMin;
    int am = 0x7fffffff;
    for (int j = t; j < V; j += G)
      if (row[j] > m || (row[j] == m && j < am)) { m = row[j]; am = j; }
    g.argmax(m, am);
    float z = 0.f;
    for (int j = t; j < V; j += G) z += expf(row[j] - m);
    z = g.sum(z);
    const float inv_z = 1.0f / z;
    int pick = am;
    if (temperature > 0.f && u != nullptr) {
      // inverse CDF in index order: first j with cumsum_{i<=j} e_i > u * z; per-thread contiguous chunks
      const float target = u[r] * z;
      const int chunk = (V + G - 1) / G;
      const int j0 = t * chunk, j1 = min(V, j0 + chunk);
      float part = 0.f;
      for (int j = j0; j < j1; ++j) part += expf(row[j] - m);
      // exclusive prefix of the per-thread partial sums (warp scan, then across warps in CTA mode)
      float incl = part;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += up;
      }
      float base = 0.f;
      if (G > 32) {
        __syncthreads();
        if ((threadIdx.x & 31) == 31) g.scratch[threadIdx.x >> 5] = incl;
        __syncthreads();
        for (int w = 0; w < (threadIdx.x >> 5); ++w) base += g.scratch[w];
      }
      const float before = base + incl - part;
      int found = 0x7fffffff;
      if (before <= target && target < before + part) {
        float c = before;
        for (int j = j0; j < j1; ++j) {
          c += expf(row[j] - m);
          if (c > target) { found = j; break; }
        }
      }
      // rounding can leave the target beyond the last partial sum: fall back to the last token with non-zero mass
      float fv = -static_cast<float>(found);
      int fi = found;
      g.argmax(fv, fi);
      if (fi != 0x7fffffff) pick = fi;
      else {
        int last = -1;
        for (int j = t; j < V; j += G)
          if (row[j] > kFiniteMin) last = max(last, j);
        float lv = static_cast<float>(last);
        int li = last;
        g.argmax(lv, li);
        pick = li >= 0 ? li : am;
      }
    }
    float conf = expf(row[pick] - m) * inv_z;
    if (conf_mode == 1) {                                   // margin: top1 - top2 probability
      float m2 = kFiniteMin;
      for (int j = t; j < V; j += G)
        if (j != am) m2 = fmaxf(m2, row[j]);
      m2 = g.max(m2);
      conf = inv_z - expf(m2 - m) * inv_z;
    } else if (conf_mode == 2) {                            // negative entropy: sum p log(p + 1e-10)
      float e = 0.f;
      for (int j = t; j < V; j += G) {
        const float p = expf(row[j] - m) * inv_z;
        e += p * logf(p + 1e-10f);
      }
      conf = g.sum(e);
    }
    if (probs_out != nullptr)
      for (int j = t; j < V; j += G) probs_out[r * ldp + j] = expf(row[j] - m) * inv_z;
    if (t == 0) {
      x0_out[r] = pick;
      if (conf_out != nullptr) conf_out[r] = conf;
    }
    g.sync();
  }
}

// "origin" algorithm: x = (x == mask && u < p_transfer) ? x0 : x     ref: generation_utils.py:156-170
__global__ void gen_unmask_origin_kernel(long long* __restrict__ x, const long long* __restrict__ x0,
                                         const float* __restrict__ u, float p_transfer, long long mask_token,
                                         long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    if (x[i] == mask_token && u[i] < p_transfer) x[i] = x0[i];
}

// Confidence algorithms: sample b reveals its k[b] most confident masked entries (ties: lowest position first).
// One CTA per sample.  gumbel_u != NULL: conf <- conf / alg_temp - log(-log(u + 1e-9) + 1e-9) first.
// ref: generation_utils.py:171-228 (topk over confidence with non-masked entries at -inf, then per-sample cut at k[b]).
__global__ void __launch_bounds__(256) gen_unmask_topk_kernel(long long* __restrict__ x, const long long* __restrict__ x0,
                                                              const float* __restrict__ conf,
                                                              const float* __restrict__ gumbel_u, float alg_temp,
                                                              const int* __restrict__ k_per_sample,
                                                              long long mask_token, int P) {
  __shared__ int red[8];
  __shared__ int pre[256];
  const int b = blockIdx.x;
  const int k = k_per_sample[b];
  if (k <= 0) return;
  long long* xb = x + static_cast<long long>(b) * P;
  const long long* x0b = x0 + static_cast<long long>(b) * P;
  const float* cb = conf + static_cast<long long>(b) * P;
  const float* ub = gumbel_u != nullptr ? gumbel_u + static_cast<long long>(b) * P : nullptr;
  auto key_at = [&](int j) -> uint32_t {                     // 0 for non-masked entries (below every real key)
    if (xb[j] != mask_token) return 0u;
    float c = cb[j];
    if (ub != nullptr) c = c / alg_temp - logf(-logf(ub[j] + 1e-9f) + 1e-9f);
    const uint32_t kk = fkey(c);
    return kk == 0u ? 1u : kk;
  };
  auto block_isum = [&](int v) -> int {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
  };
  uint32_t lo = 1u, hi = 0xffffffffu;                        // largest key K >= 1 with count(key >= K) >= k
  {
    int c = 0;
    for (int j = threadIdx.x; j < P; j += 256) c += (key_at(j) >= 1u);
    if (block_isum(c) < k) hi = lo = 1u;                     // fewer masked entries than k: reveal all of them
  }
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1) + 1u;
    int c = 0;
    for (int j = threadIdx.x; j < P; j += 256) c += (key_at(j) >= mid);
    if (block_isum(c) >= k) lo = mid; else hi = mid - 1u;
  }
  int gt = 0;
  for (int j = threadIdx.x; j < P; j += 256) gt += (key_at(j) > lo);
  gt = block_isum(gt);
  const int take_eq = k - gt;                                // ties at the threshold: first take_eq in position order
  const int chunk = (P + 255) / 256;
  const int j0 = threadIdx.x * chunk, j1 = min(P, j0 + chunk);
  int eq = 0;
  for (int j = j0; j < j1; ++j) eq += (key_at(j) == lo);
  pre[threadIdx.x] = eq;
  __syncthreads();
  int before = 0;
  for (int i = 0; i < threadIdx.x; ++i) before += pre[i];
  __syncthreads();
  for (int j = j0; j < j1; ++j) {
    const uint32_t kk = key_at(j);
    bool take = kk > lo;
    if (kk == lo) {
      take = before < take_eq;
      ++before;
    }
    if (take) xb[j] = x0b[j];
  }
}

}  // namespace ggpt

using namespace ggpt;

extern "C" {

int ggpt_gen_sample(const float* logits, long long ldl, long long R, int V, float temperature, int top_k, float top_p,
                    const float* u, int conf_mode, long long* x0, float* conf, float* probs, long long ldp,
                    void* stream) {
  GGPT_REQUIRE(logits && x0, "gen_sample: null pointer");
  GGPT_REQUIRE(R > 0 && V > 0 && ldl >= V, "gen_sample: bad sizes R=%lld V=%d ldl=%lld", R, V, ldl);
  GGPT_REQUIRE(conf_mode >= 0 && conf_mode <= 2, "gen_sample: conf_mode %d (0 = p[x0], 1 = margin, 2 = -entropy)", conf_mode);
  GGPT_REQUIRE(probs == nullptr || ldp >= V, "gen_sample: ldp < V");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (V <= 2048) {      // one warp per row, 8 rows per CTA
    const size_t smem = (16 + 8 * static_cast<size_t>(V)) * sizeof(float);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(gen_sample_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
      if (e != cudaSuccess) { set_error("gen_sample: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    }
    const long long blocks = std::min<long long>((R + 7) / 8, cap);
    gen_sample_kernel<32><<<static_cast<int>(blocks), 256, smem, s>>>(logits, ldl, R, V, temperature, top_k, top_p, u,
                                                                      conf_mode, x0, conf, probs, ldp);
  } else {              // one CTA per row
    const size_t smem = (16 + static_cast<size_t>(V)) * sizeof(float);
    GGPT_REQUIRE(smem <= 200 * 1024, "gen_sample: vocabulary %d does not fit in shared memory", V);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(gen_sample_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) { set_error("gen_sample: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
    }
    const long long blocks = std::min<long long>(R, cap);
    gen_sample_kernel<256><<<static_cast<int>(blocks), 256, smem, s>>>(logits, ldl, R, V, temperature, top_k, top_p, u,
                                                                       conf_mode, x0, conf, probs, ldp);
  }
  return check_launch("gen_sample_kernel");
}

int ggpt_gen_unmask_origin(long long* x, const long long* x0, const float* u, float p_transfer, long long mask_token,
                           long long n, void* stream) {
  GGPT_REQUIRE(x && x0 && u && n > 0, "gen_unmask_origin: null pointer / empty");
  const long long blocks = std::min<long long>((n + 255) / 256, static_cast<long long>(num_sms()) * 8);
  gen_unmask_origin_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x0, u, p_transfer,
                                                                                                   mask_token, n);
  return check_launch("gen_unmask_origin_kernel");
}

int ggpt_gen_unmask_topk(long long* x, const long long* x0, const float* conf, const float* gumbel_u, float alg_temp,
                         const int* k_per_sample, long long mask_token, int B, int P, void* stream) {
  GGPT_REQUIRE(x && x0 && conf && k_per_sample, "gen_unmask_topk: null pointer");
  GGPT_REQUIRE(B > 0 && P > 0, "gen_unmask_topk: bad sizes B=%d P=%d", B, P);
  GGPT_REQUIRE(gumbel_u == nullptr || alg_temp > 0.f, "gen_unmask_topk: Gumbel noise needs alg_temp > 0");
  gen_unmask_topk_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x0, conf, gumbel_u, alg_temp, k_per_sample,
                                                                            mask_token, P);
  return check_launch("gen_unmask_topk_kernel");
}

}  // extern "C"
