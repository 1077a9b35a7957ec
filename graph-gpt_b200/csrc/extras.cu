// Side branches of the GraphGPT hot path that the headline configs leave off but the reference supports:
//   * raw-embedding input branch (config.embed_dim > 0): mask-token swap + RMSNorm over the E raw features, feeding the
//     embed_proj GEMM (ref: modeling_pretrain.py:119-150, modeling_helpers.py:127-139)
//   * element dropout on activations (config.mlp_pdrop / embed_pdrop; ref: utils_graphgpt.py:69-83)
// All HBM-bound, one pass, warp-per-row or 16-byte grid-stride.
#include "common.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

static inline int grid_cap(long long blocks, int waves) {
  const long long cap = static_cast<long long>(num_sms()) * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// =============================================================================================
// Raw-embedding branch, forward.  Row t of the result is RMSNorm_E(src_t; w) in bf16 with
//   src_t = keep_t ? raw[t,:] : mask_tok[:],   keep_t = any_{f < fchk} labels[t*ldl + f] == -100   (labels NULL: keep)
// i.e. the reference's `embed_mask * raw + (~embed_mask) * emb_mask_token` followed by embed_layernorm.
// One warp per row; E % 4 == 0.
// =============================================================================================
__global__ void raw_embed_norm_fwd_kernel(const float* __restrict__ raw, const long long* __restrict__ labels,
                                          long long ldl, int fchk, const float* __restrict__ mask_tok,
                                          const float* __restrict__ w, __nv_bfloat16* __restrict__ h, long long ldh,
                                          float* __restrict__ rstd_out, unsigned char* __restrict__ keep_out, long long T,
                                          int E, float eps) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * wpb) {
    bool keep = true;
    if (labels != nullptr) {
      bool any = false;
      for (int f = lane; f < fchk; f += 32) any |= (labels[t * ldl + f] == -100);
      keep = __any_sync(0xffffffffu, any);
    }
    const float* src = keep ? raw + t * E : mask_tok;
    float ss = 0.f;
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(src + c);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / static_cast<float>(E) + eps);
    if (lane == 0) {
      if (rstd_out != nullptr) rstd_out[t] = rstd;
      if (keep_out != nullptr) keep_out[t] = keep ? 1 : 0;
    }
    __nv_bfloat16* hr = h + t * ldh;
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(src + c);
      const float4 ww = *reinterpret_cast<const float4*>(w + c);
      uint2 o;
      o.x = pack_bf16(ww.x * (v.x * rstd), ww.y * (v.y * rstd));
      o.y = pack_bf16(ww.z * (v.z * rstd), ww.w * (v.w * rstd));
      *reinterpret_cast<uint2*>(hr + c) = o;
    }
  }
}

// Backward of the above w.r.t. the parameters only (the raw features are data): dw += sum_t dh * xhat, and for the rows
// that took the mask token, dmask_tok += rstd * (g - xhat * mean(g * xhat)) with g = dh * w.  Block-level accumulation
// in shared memory (2 * E floats), one global atomic per column per block.
__global__ void raw_embed_norm_bwd_kernel(const __nv_bfloat16* __restrict__ dh, long long lddh,
                                          const float* __restrict__ raw, const unsigned char* __restrict__ keep_in,
                                          const float* __restrict__ mask_tok, const float* __restrict__ rstd_in,
                                          const float* __restrict__ w, float* __restrict__ dw,
                                          float* __restrict__ dmask_tok, long long T, int E) {
  extern __shared__ float acc[];   // [0,E): dw partial, [E,2E): dmask partial
  for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * wpb) {
    const bool keep = keep_in == nullptr || keep_in[t] != 0;
    const float* src = keep ? raw + t * E : mask_tok;
    const float rstd = rstd_in[t];
    const __nv_bfloat16* dr = dh + t * lddh;
    float dot = 0.f;
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(src + c);
      const float4 ww = *reinterpret_cast<const float4*>(w + c);
      const uint2 dv = *reinterpret_cast<const uint2*>(dr + c);
      const float2 d01 = unpack_bf16(dv.x), d23 = unpack_bf16(dv.y);
      const float x0 = v.x * rstd, x1 = v.y * rstd, x2 = v.z * rstd, x3 = v.w * rstd;
      atomicAdd(acc + c + 0, d01.x * x0); atomicAdd(acc + c + 1, d01.y * x1);
      atomicAdd(acc + c + 2, d23.x * x2); atomicAdd(acc + c + 3, d23.y * x3);
      dot += d01.x * ww.x * x0 + d01.y * ww.y * x1 + d23.x * ww.z * x2 + d23.y * ww.w * x3;
    }
    if (keep || dmask_tok == nullptr) continue;      // warp-uniform
    dot = warp_sum(dot) / static_cast<float>(E);
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(src + c);
      const float4 ww = *reinterpret_cast<const float4*>(w + c);
      const uint2 dv = *reinterpret_cast<const uint2*>(dr + c);
      const float2 d01 = unpack_bf16(dv.x), d23 = unpack_bf16(dv.y);
      atomicAdd(acc + E + c + 0, rstd * (d01.x * ww.x - v.x * rstd * dot));
      atomicAdd(acc + E + c + 1, rstd * (d01.y * ww.y - v.y * rstd * dot));
      atomicAdd(acc + E + c + 2, rstd * (d23.x * ww.z - v.z * rstd * dot));
      atomicAdd(acc + E + c + 3, rstd * (d23.y * ww.w - v.w * rstd * dot));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    atomicAdd(dw + i, acc[i]);
    if (dmask_tok != nullptr) atomicAdd(dmask_tok + i, acc[E + i]);
  }
}

// =============================================================================================
// Raw EDGE embeddings [N,S,S,E] (fine-tuning only; modeling_helpers.py:127-139 with a 4-D input): every (n, s, s') row is
// layer-normed, dropped out and projected, then summed over s'.  embed_proj has no bias, so the sum commutes with it:
//   h[t,:] = sum_j w * raw[t,j,:] * rstd[t,j] * drop(t,j,:)     (bf16, fed to ONE embed_proj GEMM of T rows)
// One warp per token t = (n, s); the partial sums stay in registers (E <= 1024).  Backward: only embed_layernorm.weight
// has a gradient (the features are data):  dw += sum_t dh[t,:] * sum_j raw[t,j,:] rstd[t,j] drop(t,j,:).
// =============================================================================================
__global__ void raw_embed_norm_sum_kernel(const float* __restrict__ raw, const float* __restrict__ w,
                                          __nv_bfloat16* __restrict__ h, long long ldh, const __nv_bfloat16* __restrict__ dh,
                                          long long lddh, float* __restrict__ dw, long long T, int S2, int E, float eps,
                                          DropParams dp) {
  extern __shared__ float acc_dw[];   // backward only: [E]
  const bool bwd = dh != nullptr;
  if (bwd) {
    for (int i = threadIdx.x; i < E; i += blockDim.x) acc_dw[i] = 0.f;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * wpb) {
    float4 sum[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sum[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < S2; ++j) {
      const float* src = raw + (t * S2 + j) * E;
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = lane * 4 + k * 128;
        if (c < E) {
          const float4 v = *reinterpret_cast<const float4*>(src + c);
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
      }
      ss = warp_sum(ss);
      if (ss == 0.f) continue;                       // an all-zero (absent edge) row contributes nothing
      const float rstd = rsqrtf(ss / static_cast<float>(E) + eps);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = lane * 4 + k * 128;
        if (c < E) {
          const float4 v = *reinterpret_cast<const float4*>(src + c);
          float2 s01 = make_float2(1.f, 1.f), s23 = make_float2(1.f, 1.f);
          if (dp.thresh != 0u) {
            const unsigned long long e0 = (static_cast<unsigned long long>(t) * S2 + j) * E + c;
            s01 = edrop_scale2(dp, e0 >> 1);
            s23 = edrop_scale2(dp, (e0 >> 1) + 1);
          }
          sum[k].x += v.x * rstd * s01.x; sum[k].y += v.y * rstd * s01.y;
          sum[k].z += v.z * rstd * s23.x; sum[k].w += v.w * rstd * s23.y;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = lane * 4 + k * 128;
      if (c < E) {
        if (!bwd) {
          const float4 ww = *reinterpret_cast<const float4*>(w + c);
          uint2 o;
          o.x = pack_bf16(ww.x * sum[k].x, ww.y * sum[k].y);
          o.y = pack_bf16(ww.z * sum[k].z, ww.w * sum[k].w);
          *reinterpret_cast<uint2*>(h + t * ldh + c) = o;
        } else {
          const uint2 dv = *reinterpret_cast<const uint2*>(dh + t * lddh + c);
          const float2 d01 = unpack_bf16(dv.x), d23 = unpack_bf16(dv.y);
          atomicAdd(acc_dw + c + 0, d01.x * sum[k].x); atomicAdd(acc_dw + c + 1, d01.y * sum[k].y);
          atomicAdd(acc_dw + c + 2, d23.x * sum[k].z); atomicAdd(acc_dw + c + 3, d23.y * sum[k].w);
        }
      }
    }
  }
  if (bwd) {
    __syncthreads();
    for (int i = threadIdx.x; i < E; i += blockDim.x) atomicAdd(dw + i, acc_dw[i]);
  }
}

// =============================================================================================
// LayerScale / DropPath backward of a residual branch  x_out = x_in + rowscale[t] * lam[c] * y  (utils_graphgpt.py:153-166):
//   dy[t,c]  = bf16(dx[t,c] * lam[c] * rowscale[t])                       gradient handed to the branch's dgrad / wgrad GEMMs
//   dlam[c] += sum_t dx[t,c] * rowscale[t] * y[t,c] = sum_t dx[t,c] * (x_out[t,c] - x_in[t,c]) / lam[c]
// (y itself is never stored: the fused add+norm pass consumed it).  One HBM pass; each lane owns columns
// lane*4 + k*128 for every row its warp visits, so the dlam partial sums stay in registers.
// =============================================================================================
template <int NV>
__global__ void __launch_bounds__(256) layerscale_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ x_out,
                                                             const float* __restrict__ x_in, const float* __restrict__ lam,
                                                             const float* __restrict__ rowscale,
                                                             __nv_bfloat16* __restrict__ dy, float* __restrict__ dlam,
                                                             long long T, int d) {
  extern __shared__ float s_acc[];   // [d]
  for (int c = threadIdx.x; c < d; c += blockDim.x) s_acc[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float4 acc[NV], lv[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane * 4 + k * 128;
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    lv[k] = make_float4(1.f, 1.f, 1.f, 1.f);
    if (c < d && lam != nullptr) lv[k] = *reinterpret_cast<const float4*>(lam + c);
  }
  for (long long t = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * wpb) {
    const float rs = rowscale != nullptr ? rowscale[t] : 1.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane * 4 + k * 128;
      if (c < d) {
        const float4 g = *reinterpret_cast<const float4*>(dx + t * d + c);
        if (dlam != nullptr) {
          const float4 a = *reinterpret_cast<const float4*>(x_out + t * d + c);
          const float4 b = *reinterpret_cast<const float4*>(x_in + t * d + c);
          acc[k].x += g.x * (a.x - b.x); acc[k].y += g.y * (a.y - b.y);
          acc[k].z += g.z * (a.z - b.z); acc[k].w += g.w * (a.w - b.w);
        }
        uint2 o;
        o.x = pack_bf16(g.x * lv[k].x * rs, g.y * lv[k].y * rs);
        o.y = pack_bf16(g.z * lv[k].z * rs, g.w * lv[k].w * rs);
        *reinterpret_cast<uint2*>(dy + t * d + c) = o;
      }
    }
  }
  if (dlam == nullptr) return;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane * 4 + k * 128;
    if (c < d) {
      atomicAdd(&s_acc[c + 0], acc[k].x / lv[k].x); atomicAdd(&s_acc[c + 1], acc[k].y / lv[k].y);
      atomicAdd(&s_acc[c + 2], acc[k].z / lv[k].z); atomicAdd(&s_acc[c + 3], acc[k].w / lv[k].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) atomicAdd(dlam + c, s_acc[c]);
}

// =============================================================================================
// Element dropout.  x[e] *= keep(e) / (1 - p) in place on a bf16 tensor viewed as a flat array of n elements;
// scale_f32 writes the keep/(1-p) factors themselves (tests and tooling read the mask through it).
// =============================================================================================
__global__ void dropout_bf16_kernel(__nv_bfloat16* __restrict__ x, long long n, DropParams dp) {
  const long long nvec = n >> 3;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 v = *reinterpret_cast<uint4*>(x + i * 8);
    uint32_t* q = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 s = edrop_scale2(dp, static_cast<unsigned long long>(i) * 4 + k);
      const float2 a = unpack_bf16(q[k]);
      q[k] = pack_bf16(a.x * s.x, a.y * s.y);
    }
    *reinterpret_cast<uint4*>(x + i * 8) = v;
  }
  if (blockIdx.x == 0) {
    for (long long e = nvec * 8 + threadIdx.x; e < n; e += blockDim.x)
      x[e] = __float2bfloat16_rn(__bfloat162float(x[e]) * edrop_scale1(dp, static_cast<unsigned long long>(e)));
  }
}

__global__ void dropout_scale_f32_kernel(float* __restrict__ out, long long n, DropParams dp) {
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * blockDim.x)
    out[e] = dp.thresh != 0u ? edrop_scale1(dp, static_cast<unsigned long long>(e)) : 1.0f;
}

// =============================================================================================
// Device-side sequence packer.  Packed sequence n is the concatenation of the graphs
// seq_graphs[cu_seq[n] .. cu_seq[n+1]), each followed by ONE separator row (the reference's <eos> row: pack_token_seq
// appends `seps` between graphs and prepare_inputs_for_pretrain_mlm appends the final one — tokenizer.py:359-415,
// tokenizer_utils.py:228-233,246), truncated to S rows (collator: val[:pad_to]) and padded with pad_id.
// Output: input_ids [N,S,F] and the segment id of every row (1-based, 0 = pad) — the [N,S] form of the block-diagonal
// attention mask of tokenizer_utils.py:351-355 that ggpt_attn_mask_build understands.  One CTA per packed sequence.
// =============================================================================================
constexpr int kPackMaxGraphs = 4096;
__global__ void __launch_bounds__(256) pack_sequences_kernel(const long long* __restrict__ rows, int F,
                                                             const int* __restrict__ cu_rows,
                                                             const int* __restrict__ seq_graphs,
                                                             const int* __restrict__ cu_seq, int S,
                                                             const long long* __restrict__ sep_row, long long pad_id,
                                                             long long* __restrict__ input_ids,
                                                             long long* __restrict__ seg_ids,
                                                             long long* __restrict__ position_ids,
                                                             int* __restrict__ n_valid) {
  __shared__ int seg_end[kPackMaxGraphs];
  __shared__ int n_seg;
  const int n = blockIdx.x;
  const int g0 = cu_seq[n], g1 = cu_seq[n + 1];
  if (threadIdx.x == 0) {
    int tot = 0, k = 0;
    for (int g = g0; g < g1 && k < kPackMaxGraphs && tot < S; ++g, ++k) {
      const int gi = seq_graphs[g];
      tot += cu_rows[gi + 1] - cu_rows[gi] + 1;           // the graph's rows + its separator row
      seg_end[k] = tot;
    }
    n_seg = k;
    if (n_valid != nullptr) n_valid[n] = min(tot, S);
  }
  __syncthreads();
  const int ns = n_seg;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    int lo = 0, hi = ns;                                  // first segment whose end is > s
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (seg_end[mid] > s) hi = mid; else lo = mid + 1;
    }
    long long* dst = input_ids + (static_cast<long long>(n) * S + s) * F;
    long long seg = 0;
    if (lo < ns) {
      const int start = lo == 0 ? 0 : seg_end[lo - 1];
      const int gi = seq_graphs[g0 + lo];
      const int r0 = cu_rows[gi], nr = cu_rows[gi + 1] - r0;
      const long long* src = (s - start < nr) ? rows + static_cast<long long>(r0 + s - start) * F : sep_row;
      for (int f = 0; f < F; ++f) dst[f] = src[f];
      seg = lo + 1;
    } else {
      for (int f = 0; f < F; ++f) dst[f] = pad_id;
    }
    seg_ids[static_cast<long long>(n) * S + s] = seg;
    if (position_ids != nullptr) position_ids[static_cast<long long>(n) * S + s] = s;
  }
}

}  // namespace ggpt

using namespace ggpt;

extern "C" {

int ggpt_raw_embed_norm_fwd(const float* raw, const long long* labels, long long ldl, int fchk, const float* mask_tok,
                            const float* w, void* h, long long ldh, float* rstd, unsigned char* keep, long long T, int E,
                            float eps, void* stream) {
  GGPT_REQUIRE(raw && w && h, "raw_embed_norm_fwd: null pointer");
  GGPT_REQUIRE(labels == nullptr || (mask_tok != nullptr && fchk > 0 && ldl >= fchk),
               "raw_embed_norm_fwd: labels need mask_tok and 0 < fchk <= ldl");
  GGPT_REQUIRE(T > 0 && E > 0 && E % 4 == 0 && ldh % 4 == 0, "raw_embed_norm_fwd: bad sizes T=%lld E=%d ldh=%lld", T, E, ldh);
  raw_embed_norm_fwd_kernel<<<grid_cap((T + 7) / 8, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      raw, labels, ldl, fchk, mask_tok, w, static_cast<__nv_bfloat16*>(h), ldh, rstd, keep, T, E, eps);
  return check_launch("raw_embed_norm_fwd_kernel");
}

int ggpt_raw_embed_norm_bwd(const void* dh, long long lddh, const float* raw, const unsigned char* keep,
                            const float* mask_tok, const float* rstd, const float* w, float* dw, float* dmask_tok,
                            long long T, int E, void* stream) {
  GGPT_REQUIRE(dh && raw && rstd && w && dw, "raw_embed_norm_bwd: null pointer");
  GGPT_REQUIRE(keep == nullptr || mask_tok != nullptr, "raw_embed_norm_bwd: keep flags need mask_tok");
  GGPT_REQUIRE(T > 0 && E > 0 && E % 4 == 0 && lddh % 4 == 0 && E <= 8192, "raw_embed_norm_bwd: bad sizes T=%lld E=%d", T, E);
  const size_t smem = 2 * static_cast<size_t>(E) * sizeof(float);
  if (smem > 48 * 1024) {
    static bool raised = false;
    if (!raised) {
      cudaError_t e = cudaFuncSetAttribute(raw_embed_norm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      if (e != cudaSuccess) {
        set_error("raw_embed_norm_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return -2;
      }
      raised = true;
    }
  }
  raw_embed_norm_bwd_kernel<<<grid_cap((T + 7) / 8, 2), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dh), lddh, raw, keep, mask_tok, rstd, w, dw, dmask_tok, T, E);
  return check_launch("raw_embed_norm_bwd_kernel");
}

int ggpt_raw_embed_norm_sum(const float* raw, const float* w, void* h, long long ldh, const void* dh, long long lddh,
                            float* dw, long long T, int S2, int E, float eps, float drop_p, unsigned long long drop_seed,
                            void* stream) {
  GGPT_REQUIRE(raw && T > 0 && S2 > 0 && E > 0 && E % 4 == 0 && E <= 1024, "raw_embed_norm_sum: bad sizes T=%lld S2=%d E=%d (E %% 4 == 0, E <= 1024)", T, S2, E);
  GGPT_REQUIRE((dh == nullptr) ? (w && h && ldh % 4 == 0) : (dw != nullptr && lddh % 4 == 0), "raw_embed_norm_sum: forward needs w / h, backward needs dh / dw");
  GGPT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "raw_embed_norm_sum: dropout p=%f outside [0,1)", drop_p);
  const size_t smem = dh != nullptr ? static_cast<size_t>(E) * sizeof(float) : 0;
  raw_embed_norm_sum_kernel<<<grid_cap((T + 7) / 8, 8), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      raw, w, static_cast<__nv_bfloat16*>(h), ldh, static_cast<const __nv_bfloat16*>(dh), lddh, dw, T, S2, E, eps,
      make_drop_params(drop_p, drop_seed));
  return check_launch("raw_embed_norm_sum_kernel");
}

int ggpt_layerscale_bwd(const float* dx, const float* x_out, const float* x_in, const float* lam, const float* rowscale,
                        void* dy, float* dlam, long long T, int d, void* stream) {
  GGPT_REQUIRE(dx && dy, "layerscale_bwd: null pointer");
  GGPT_REQUIRE(dlam == nullptr || (lam && x_out && x_in), "layerscale_bwd: dlam needs lam, x_out and x_in");
  GGPT_REQUIRE(T > 0 && d > 0 && d % 4 == 0 && d <= 2048, "layerscale_bwd: bad sizes T=%lld d=%d", T, d);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* dyb = static_cast<__nv_bfloat16*>(dy);
  const int grid = grid_cap((T + 7) / 8, 4);
  const size_t sm = d * sizeof(float);
  const int nv = (d + 127) / 128;
  if (nv <= 1) layerscale_bwd_kernel<1><<<grid, 256, sm, s>>>(dx, x_out, x_in, lam, rowscale, dyb, dlam, T, d);
  else if (nv <= 2) layerscale_bwd_kernel<2><<<grid, 256, sm, s>>>(dx, x_out, x_in, lam, rowscale, dyb, dlam, T, d);
  else if (nv <= 4) layerscale_bwd_kernel<4><<<grid, 256, sm, s>>>(dx, x_out, x_in, lam, rowscale, dyb, dlam, T, d);
  else if (nv <= 6) layerscale_bwd_kernel<6><<<grid, 256, sm, s>>>(dx, x_out, x_in, lam, rowscale, dyb, dlam, T, d);
  else if (nv <= 8) layerscale_bwd_kernel<8><<<grid, 256, sm, s>>>(dx, x_out, x_in, lam, rowscale, dyb, dlam, T, d);
  else layerscale_bwd_kernel<16><<<grid, 256, sm, s>>>(dx, x_out, x_in, lam, rowscale, dyb, dlam, T, d);
  return check_launch("layerscale_bwd_kernel");
}

int ggpt_pack_sequences(const long long* rows, int F, const int* cu_rows, const int* seq_graphs, const int* cu_seq, int N,
                        int S, const long long* sep_row, long long pad_id, long long* input_ids, long long* seg_ids,
                        long long* position_ids, int* n_valid, void* stream) {
  GGPT_REQUIRE(rows && cu_rows && seq_graphs && cu_seq && sep_row && input_ids && seg_ids, "pack_sequences: null pointer");
  GGPT_REQUIRE(N > 0 && S > 0 && F > 0, "pack_sequences: bad sizes N=%d S=%d F=%d", N, S, F);
  pack_sequences_kernel<<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, F, cu_rows, seq_graphs, cu_seq, S, sep_row,
                                                                           pad_id, input_ids, seg_ids, position_ids, n_valid);
  return check_launch("pack_sequences_kernel");
}

int ggpt_dropout_bf16(void* x, long long n, float p, unsigned long long seed, void* stream) {
  GGPT_REQUIRE(x != nullptr && n > 0, "dropout_bf16: null pointer / empty tensor");
  GGPT_REQUIRE(p >= 0.f && p < 1.f, "dropout_bf16: p=%f outside [0,1)", p);
  GGPT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "dropout_bf16: x must be 16-byte aligned");
  if (p == 0.f) return 0;
  dropout_bf16_kernel<<<grid_cap(((n >> 3) + 255) / 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(x), n, make_drop_params(p, seed));
  return check_launch("dropout_bf16_kernel");
}

int ggpt_dropout_scale_f32(float* out, long long n, float p, unsigned long long seed, void* stream) {
  GGPT_REQUIRE(out != nullptr && n > 0, "dropout_scale_f32: null pointer / empty tensor");
  GGPT_REQUIRE(p >= 0.f && p < 1.f, "dropout_scale_f32: p=%f outside [0,1)", p);
  dropout_scale_f32_kernel<<<grid_cap((n + 255) / 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out, n, make_drop_params(p, seed));
  return check_launch("dropout_scale_f32_kernel");
}

}  // extern "C"
