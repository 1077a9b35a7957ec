// HBM-bound kernels of the GraphGPT hot path: stacked-token embedding gather-sum, RMSNorm, GeGLU backward,
// loss-head compaction / gathers, fp32 cross-entropy, fused AdamW.  All are one-pass, vectorised (16-byte
// accesses), warp-shuffle reductions, grid sized from the SM count.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

static inline int grid_for_rows(long long rows, int rows_per_block, int max_waves = 8) {
  long long blocks = (rows + rows_per_block - 1) / rows_per_block;
  long long cap = static_cast<long long>(num_sms()) * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// =============================================================================================
// Stacked-token embedding: x[t,:] = sum_f gate[f,:] * E[ids[t,f],:]   (gate == NULL -> plain sum)
// ref: modeling_helpers.py:89-114 (_get_stacked_inputs_embeds), modeling_common.py:127-135
// =============================================================================================
__global__ void embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                 const float* __restrict__ gate, float* __restrict__ out, long long T, int F, int d,
                                 int V, int long_scale, int* __restrict__ err, DropParams dp) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * warps_per_block) {
    const long long* idrow = ids + t * F;
    int nnz = 0;
    for (int c = lane * 4; c < d; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int f = 0; f < F; ++f) {
        long long id = idrow[f];
        if (id < 0 || id >= V) {
          if (err) atomicExch(err, 1);
          id = 0;
        }
        if (c == lane * 4) nnz += (id != 0);
        float4 e = *reinterpret_cast<const float4*>(table + id * d + c);
        if (dp.thresh != 0u) {  // embed_dropout acts on the gathered [T,F,d] rows before aggregation (modeling_helpers.py:97-98)
          const unsigned long long pair = (static_cast<unsigned long long>(t * F + f) * d + c) >> 1;
          const float2 s01 = edrop_scale2(dp, pair), s23 = edrop_scale2(dp, pair + 1);
          e.x *= s01.x; e.y *= s01.y; e.z *= s23.x; e.w *= s23.y;
        }
        if (gate != nullptr) {
          const float4 g = *reinterpret_cast<const float4*>(gate + static_cast<long long>(f) * d + c);
          e.x *= g.x; e.y *= g.y; e.z *= g.z; e.w *= g.w;
        }
        acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
      }
      if (long_scale) {  // stack_method == "long": scale by 1/clamp(#non-pad features) (modeling_helpers.py:106-110)
        const float ratio = fminf(1.0f / (static_cast<float>(nnz) + 1e-7f), 1.0f);
        acc.x *= ratio; acc.y *= ratio; acc.z *= ratio; acc.w *= ratio;
      }
      *reinterpret_cast<float4*>(out + t * d + c) = acc;
    }
  }
}

// dE[id,:] += gate[f,:] * dx[t,:] for id != padding_idx;  dgate[f,:] += E[id,:] * dx[t,:]
__global__ void embed_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ dx,
                                 const float* __restrict__ table, const float* __restrict__ gate,
                                 float* __restrict__ dtable, float* __restrict__ dgate, long long T, int F, int d, int V,
                                 int padding_idx, int long_scale, DropParams dp) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * warps_per_block) {
    const long long* idrow = ids + t * F;
    float ratio = 1.f;
    if (long_scale) {
      int nnz = 0;
      for (int f = 0; f < F; ++f) nnz += (idrow[f] != 0);
      ratio = fminf(1.0f / (static_cast<float>(nnz) + 1e-7f), 1.0f);
    }
    for (int c = lane * 4; c < d; c += 128) {
      float4 g = *reinterpret_cast<const float4*>(dx + t * d + c);
      g.x *= ratio; g.y *= ratio; g.z *= ratio; g.w *= ratio;
      for (int f = 0; f < F; ++f) {
        long long id = idrow[f];
        if (id < 0 || id >= V) continue;
        float4 v = g;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
        if (dp.thresh != 0u) {
          const unsigned long long pair = (static_cast<unsigned long long>(t * F + f) * d + c) >> 1;
          const float2 s01 = edrop_scale2(dp, pair), s23 = edrop_scale2(dp, pair + 1);
          sc = make_float4(s01.x, s01.y, s23.x, s23.y);
        }
        if (gate != nullptr) {
          const float4 gt = *reinterpret_cast<const float4*>(gate + static_cast<long long>(f) * d + c);
          const float4 e = *reinterpret_cast<const float4*>(table + id * d + c);
          float* dg = dgate + static_cast<long long>(f) * d + c;
          atomicAdd(dg + 0, e.x * sc.x * g.x); atomicAdd(dg + 1, e.y * sc.y * g.y);
          atomicAdd(dg + 2, e.z * sc.z * g.z); atomicAdd(dg + 3, e.w * sc.w * g.w);
          v.x *= gt.x; v.y *= gt.y; v.z *= gt.z; v.w *= gt.w;
        }
        v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
        if (id == padding_idx) continue;  // nn.Embedding(padding_idx): row never receives gradient (HF:361)
        float* dst = dtable + id * d + c;
        atomicAdd(dst + 0, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
      }
    }
  }
}

// Token-count matrix for the embedding gradient as a GEMM:  C[t, v] = #{f : ids[t,f] == v} (bf16, exact small ints),
// C[t, padding_idx] = 0.  Then dE[V,d] = C^T dX runs on the tensor cores (wgrad GEMM) instead of T*F*d fp32 atomics
// that serialise on hot rows (<mask> carries ~half of all entries).  One warp per token; F <= 32.
__global__ void embed_count_kernel(const long long* __restrict__ ids, __nv_bfloat16* __restrict__ cnt, long long ldc,
                                   long long T, int F, int V, int padding_idx) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * warps_per_block) {
    __nv_bfloat16* row = cnt + t * ldc;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (long long c = lane * 8; c < ldc; c += 256) *reinterpret_cast<uint4*>(row + c) = z;
    __syncwarp();
    long long my = (lane < F) ? ids[t * F + lane] : -1;
    int count = 0;
    bool first = true;
    for (int f = 0; f < F; ++f) {
      const long long other = __shfl_sync(0xffffffffu, my, f);
      if (other == my) {
        ++count;
        if (f < lane) first = false;
      }
    }
    if (lane < F && first && my >= 0 && my < V && my != padding_idx) row[my] = __float2bfloat16_rn(static_cast<float>(count));
  }
}

// =============================================================================================
// In-model SMTP masking (config.smtp_inside): the collator ships raw ids [N, S, Ftot] whose column F+2 holds each
// row's node index; one uniform draw per (sample, node, feature) decides whether that feature of that NODE is masked,
// and every row of the Eulerian path that revisits the node looks the decision up — all rows of a node are masked
// together.  keep-rate threshold = mr[n]^power with mr ~ U(0,1) per sample.
//   out_ids[n,s,f] = masked ? mask_token : ids[n,s,f];   labels[n,s,f] = masked ? ids[n,s,f] : label_pad
//   masked = (u_node[n, node_idx[n,s], f] > mr[n]^power) && ids[n,s,f] > 0
// ref: modeling_helpers.py:399-452 (prepare_for_2d_smtp_inputs_labels with smtp_2d_rate = 1, replace_rate = 0,
// global_2d_mask = False, as called from modeling_pretrain.py:175-189).  One thread per (n, s, f) entry.
// =============================================================================================
__global__ void smtp_mask_2d_kernel(const long long* __restrict__ ids, int Ftot, int node_col,
                                    const float* __restrict__ mr, const float* __restrict__ u_node, float power,
                                    long long* __restrict__ out_ids, long long* __restrict__ labels, int N, int S, int F,
                                    long long mask_token, long long label_pad, int* __restrict__ err) {
  const long long total = static_cast<long long>(N) * S * F;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int f = static_cast<int>(i % F);
    const long long ns = i / F;
    const int n = static_cast<int>(ns / S);
    const long long v = ids[ns * Ftot + f];
    long long node = ids[ns * Ftot + node_col];
    if (node < 0) node += S;                       // torch advanced indexing wraps negative indices
    if (node < 0 || node >= S) {
      if (err) atomicExch(err, 3);
      node = 0;
    }
    const float thr = powf(mr[n], power);
    const bool masked = (u_node[(static_cast<long long>(n) * S + node) * F + f] > thr) && (v > 0);
    out_ids[i] = masked ? mask_token : v;
    labels[i] = masked ? v : label_pad;
  }
}

// =============================================================================================
// RMSNorm  y = bf16( w * x * rsqrt(mean(x^2) + eps) ), fp32 statistics.   ref: HF:59-64
// =============================================================================================
__global__ void rmsnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                   __nv_bfloat16* __restrict__ y, long long ldy, float* __restrict__ rstd_out,
                                   long long T, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * warps_per_block) {
    const float* xr = x + t * d;
    float ss = 0.f;
    for (int c = lane * 4; c < d; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / static_cast<float>(d) + eps);
    if (lane == 0 && rstd_out != nullptr) rstd_out[t] = rstd;
    __nv_bfloat16* yr = y + t * ldy;
    for (int c = lane * 4; c < d; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      const float4 ww = *reinterpret_cast<const float4*>(w + c);
      uint2 o;
      o.x = pack_bf16(ww.x * (v.x * rstd), ww.y * (v.y * rstd));
      o.y = pack_bf16(ww.z * (v.z * rstd), ww.w * (v.w * rstd));
      *reinterpret_cast<uint2*>(yr + c) = o;
    }
  }
}

// Fused residual add + RMSNorm:  x_out = x_in + rowscale[t] * colscale[:] * y   (y = bf16 branch output of o_proj /
// down_proj; colscale = LayerScale lambda, rowscale = DropPath scale, both optional), h = bf16(w * x_out * rstd).
// One HBM pass (600 B/token-dim... 4+2 in, 4+2 out) instead of a latency-bound fp32 read-modify-write in the GEMM
// epilogue followed by a separate norm pass.   ref: HF:325,331 (residual adds) + HF:59-64 (next RMSNorm)
__global__ void add_rmsnorm_fwd_kernel(const float* __restrict__ x_in, const __nv_bfloat16* __restrict__ y, long long ldy,
                                       const float* __restrict__ colscale, const float* __restrict__ rowscale,
                                       const float* __restrict__ w, float* __restrict__ x_out,
                                       __nv_bfloat16* __restrict__ h, float* __restrict__ rstd_out, long long T, int d,
                                       float eps) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * warps_per_block) {
    const float* xr = x_in + t * d;
    const __nv_bfloat16* yr = y + t * ldy;
    float* xo = x_out + t * d;
    const float rs = rowscale != nullptr ? rowscale[t] : 1.0f;
    float ss = 0.f;
    for (int c = lane * 4; c < d; c += 128) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + c);
      const uint2 yv = *reinterpret_cast<const uint2*>(yr + c);
      const float2 y01 = unpack_bf16(yv.x), y23 = unpack_bf16(yv.y);
      float4 sc = make_float4(rs, rs, rs, rs);
      if (colscale != nullptr) {
        const float4 cs = *reinterpret_cast<const float4*>(colscale + c);
        sc.x *= cs.x; sc.y *= cs.y; sc.z *= cs.z; sc.w *= cs.w;
      }
      float4 o;
      o.x = fmaf(y01.x, sc.x, xv.x); o.y = fmaf(y01.y, sc.y, xv.y);
      o.z = fmaf(y23.x, sc.z, xv.z); o.w = fmaf(y23.y, sc.w, xv.w);
      *reinterpret_cast<float4*>(xo + c) = o;
      ss += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / static_cast<float>(d) + eps);
    if (lane == 0 && rstd_out != nullptr) rstd_out[t] = rstd;
    if (h == nullptr) continue;
    __nv_bfloat16* hr = h + t * d;
    for (int c = lane * 4; c < d; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xo + c);     // just written by this lane: L1/L2 hit
      const float4 ww = *reinterpret_cast<const float4*>(w + c);
      uint2 o;
      o.x = pack_bf16(ww.x * (v.x * rstd), ww.y * (v.y * rstd));
      o.y = pack_bf16(ww.z * (v.z * rstd), ww.w * (v.w * rstd));
      *reinterpret_cast<uint2*>(hr + c) = o;
    }
  }
}

// Same, for d = NV * 128: the row stays in registers between the two phases (the generic kernel re-reads the x' it has just
// written, through L1) and so do this lane's slices of the norm weight and of colscale for all rows of the warp.
template <int NV>
__global__ void add_rmsnorm_fwd_reg_kernel(const float* __restrict__ x_in, const __nv_bfloat16* __restrict__ y, long long ldy,
                                           const float* __restrict__ colscale, const float* __restrict__ rowscale,
                                           const float* __restrict__ w, float* __restrict__ x_out,
                                           __nv_bfloat16* __restrict__ h, float* __restrict__ rstd_out, long long T,
                                           float eps) {
  constexpr int d = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  float4 wreg[NV], cs[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane * 4 + k * 128;
    wreg[k] = __ldg(reinterpret_cast<const float4*>(w + c));
    cs[k] = colscale != nullptr ? __ldg(reinterpret_cast<const float4*>(colscale + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  for (long long t = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * warps_per_block) {
    const float* xr = x_in + t * d;
    const __nv_bfloat16* yr = y + t * ldy;
    float4 xv[NV];
    uint2 yv[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {          // all loads of the row first
      const int c = lane * 4 + k * 128;
      xv[k] = *reinterpret_cast<const float4*>(xr + c);
      yv[k] = *reinterpret_cast<const uint2*>(yr + c);
    }
    const float rs = rowscale != nullptr ? rowscale[t] : 1.0f;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane * 4 + k * 128;
      const float2 y01 = unpack_bf16(yv[k].x), y23 = unpack_bf16(yv[k].y);
      float4 o;
      o.x = fmaf(y01.x, rs * cs[k].x, xv[k].x); o.y = fmaf(y01.y, rs * cs[k].y, xv[k].y);
      o.z = fmaf(y23.x, rs * cs[k].z, xv[k].z); o.w = fmaf(y23.y, rs * cs[k].w, xv[k].w);
      *reinterpret_cast<float4*>(x_out + t * d + c) = o;
      xv[k] = o;
      ss += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / static_cast<float>(d) + eps);
    if (lane == 0 && rstd_out != nullptr) rstd_out[t] = rstd;
    if (h == nullptr) continue;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane * 4 + k * 128;
      uint2 o;
      o.x = pack_bf16(wreg[k].x * (xv[k].x * rstd), wreg[k].y * (xv[k].y * rstd));
      o.y = pack_bf16(wreg[k].z * (xv[k].z * rstd), wreg[k].w * (xv[k].w * rstd));
      *reinterpret_cast<uint2*>(h + t * d + c) = o;
    }
  }
}

// dx_out = dresid + rstd * (g - xhat * mean(g * xhat)),  g = dy * w,  xhat = x * rstd;  dw += sum_t dy * xhat
// Also emits a bf16 copy of dx_out (the A operand of the next dgrad GEMMs).
// Each lane owns the same columns (lane*4 + k*128) for every row its warp visits, so the row is held in registers
// (one HBM pass) and the dw partial sums stay in registers until one smem + global reduction per block.
template <int NV>
__global__ void __launch_bounds__(128, 3) rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const float* __restrict__ x,
                                   const float* __restrict__ rstd_in, const float* __restrict__ w,
                                   const float* __restrict__ dresid, float* __restrict__ dx_out,
                                   __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dw, long long T, int d) {
  extern __shared__ float s_dw[];  // [d]
  for (int c = threadIdx.x; c < d; c += blockDim.x) s_dw[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_d = 1.0f / static_cast<float>(d);
  float4 dwacc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) dwacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long t = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * warps_per_block) {
    const float* xr = x + t * d;
    const __nv_bfloat16* dyr = dy + t * lddy;
    const float rstd = rstd_in[t];
    float4 xs[NV], gv[NV], rv[NV];   // xhat, dy (then g = dy*w), residual-path gradient
    float dot = 0.f;
    // issue every load of the row (x, dy, dresid) before the reduction so three streams are in flight per warp
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane * 4 + k * 128;
      rv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d && dresid != nullptr) rv[k] = *reinterpret_cast<const float4*>(dresid + t * d + c);
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane * 4 + k * 128;
      if (c < d) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + c);
        const uint2 dv = *reinterpret_cast<const uint2*>(dyr + c);
        const float2 d01 = unpack_bf16(dv.x), d23 = unpack_bf16(dv.y);
        xs[k] = make_float4(xv.x * rstd, xv.y * rstd, xv.z * rstd, xv.w * rstd);
        gv[k] = make_float4(d01.x, d01.y, d23.x, d23.y);
        dwacc[k].x += gv[k].x * xs[k].x; dwacc[k].y += gv[k].y * xs[k].y;
        dwacc[k].z += gv[k].z * xs[k].z; dwacc[k].w += gv[k].w * xs[k].w;
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c));   // 3 KB, L1-resident
        gv[k].x *= wv.x; gv[k].y *= wv.y; gv[k].z *= wv.z; gv[k].w *= wv.w;
        dot += gv[k].x * xs[k].x + gv[k].y * xs[k].y + gv[k].z * xs[k].z + gv[k].w * xs[k].w;
      }
    }
    dot = warp_sum(dot) * inv_d;  // mean(g * xhat)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane * 4 + k * 128;
      if (c < d) {
        const float4 r = rv[k];
        float4 o;
        o.x = r.x + rstd * (gv[k].x - xs[k].x * dot);
        o.y = r.y + rstd * (gv[k].y - xs[k].y * dot);
        o.z = r.z + rstd * (gv[k].z - xs[k].z * dot);
        o.w = r.w + rstd * (gv[k].w - xs[k].w * dot);
        *reinterpret_cast<float4*>(dx_out + t * d + c) = o;
        if (dx_bf16 != nullptr) {
          uint2 ob;
          ob.x = pack_bf16(o.x, o.y);
          ob.y = pack_bf16(o.z, o.w);
          *reinterpret_cast<uint2*>(dx_bf16 + t * d + c) = ob;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane * 4 + k * 128;
    if (c < d) {
      atomicAdd(&s_dw[c + 0], dwacc[k].x); atomicAdd(&s_dw[c + 1], dwacc[k].y);
      atomicAdd(&s_dw[c + 2], dwacc[k].z); atomicAdd(&s_dw[c + 3], dwacc[k].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) atomicAdd(dw + c, s_dw[c]);
}

// Same computation with the row inputs streamed through shared memory by 1-D bulk copies (cp.async.bulk, the TMA engine
// without a tensor map): every warp owns a private ring of kStages row slots (x | dresid | dy = 10*d bytes) guarded by
// its own mbarriers, lane 0 keeps kStages rows of loads in flight while the warp reduces / writes the current one.
// The register version above alternates "issue loads - wait - compute - store" per warp and reaches ~3.7 TB/s on a
// B200 (12 warps per SM at 154 registers); decoupling the loads from the registers keeps 24 rows per SM in flight.
template <int NV, int kStages>
__global__ void __launch_bounds__(256, 1)
rmsnorm_bwd_bulk_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const float* __restrict__ x,
                        const float* __restrict__ rstd_in, const float* __restrict__ w,
                        const float* __restrict__ dresid, float* __restrict__ dx_out,
                        __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dw, long long T, int d) {
  extern __shared__ __align__(128) uint8_t rb_smem[];
  constexpr int kWarps = 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row_bytes = 10 * d;                                   // x (4d) | dresid (4d) | dy (2d)
  float* s_dw = reinterpret_cast<float*>(rb_smem);                // [d]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rb_smem + ((4 * d + 127) / 128) * 128);   // [kWarps][kStages]
  uint8_t* ring = reinterpret_cast<uint8_t*>(bars) + 128 * ((kWarps * kStages * 8 + 127) / 128) +
                  static_cast<size_t>(warp) * kStages * row_bytes;
  uint64_t* my_bar = bars + warp * kStages;
  for (int c = threadIdx.x; c < d; c += blockDim.x) s_dw[c] = 0.f;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&my_bar[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = static_cast<long long>(blockIdx.x) * kWarps + warp;
  const long long tstep = static_cast<long long>(gridDim.x) * kWarps;
  const uint32_t tx_bytes = static_cast<uint32_t>((dresid != nullptr ? 10 : 6) * d);
  auto issue = [&](long long t, int s) {   // lane 0 only
    uint8_t* dst = ring + s * row_bytes;
    mbar_expect_tx(&my_bar[s], tx_bytes);
    bulk_load_1d(dst, x + t * d, 4 * d, &my_bar[s]);
    if (dresid != nullptr) bulk_load_1d(dst + 4 * d, dresid + t * d, 4 * d, &my_bar[s]);
    bulk_load_1d(dst + 8 * d, dy + t * lddy, 2 * d, &my_bar[s]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s)
      if (t0 + s * tstep < T) issue(t0 + s * tstep, s);
  }
  const float inv_d = 1.0f / static_cast<float>(d);
  float4 dwacc[NV], wreg[NV];               // this lane's slice of the norm weight stays in registers for all rows
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    dwacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = lane * 4 + k * 128;
    wreg[k] = (c < d) ? __ldg(reinterpret_cast<const float4*>(w + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int stage = 0;
  uint32_t phase = 0;
  float rstd_next = (t0 < T) ? rstd_in[t0] : 0.f;
  for (long long t = t0; t < T; t += tstep) {
    const float rstd = rstd_next;
    if (t + tstep < T) rstd_next = rstd_in[t + tstep];   // one row ahead: its latency overlaps this row's work
    mbar_wait(&my_bar[stage], phase);
    const uint8_t* src = ring + stage * row_bytes;
    float4 xs[NV], gv[NV];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane * 4 + k * 128;
      if (c < d) {
        const float4 xv = *reinterpret_cast<const float4*>(src + 4 * c);
        const uint2 dv = *reinterpret_cast<const uint2*>(src + 8 * d + 2 * c);
        const float2 d01 = unpack_bf16(dv.x), d23 = unpack_bf16(dv.y);
        xs[k] = make_float4(xv.x * rstd, xv.y * rstd, xv.z * rstd, xv.w * rstd);
        gv[k] = make_float4(d01.x, d01.y, d23.x, d23.y);
        dwacc[k].x += gv[k].x * xs[k].x; dwacc[k].y += gv[k].y * xs[k].y;
        dwacc[k].z += gv[k].z * xs[k].z; dwacc[k].w += gv[k].w * xs[k].w;
#ifdef GGPT_RMSBWD_LDG_W      // the previous form, for same-box A/B runs
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c));
#else
        const float4 wv = wreg[k];
#endif
        gv[k].x *= wv.x; gv[k].y *= wv.y; gv[k].z *= wv.z; gv[k].w *= wv.w;
        dot += gv[k].x * xs[k].x + gv[k].y * xs[k].y + gv[k].z * xs[k].z + gv[k].w * xs[k].w;
      }
    }
    dot = warp_sum(dot) * inv_d;  // mean(g * xhat)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane * 4 + k * 128;
      if (c < d) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dresid != nullptr) r = *reinterpret_cast<const float4*>(src + 4 * d + 4 * c);
        float4 o;
        o.x = r.x + rstd * (gv[k].x - xs[k].x * dot);
        o.y = r.y + rstd * (gv[k].y - xs[k].y * dot);
        o.z = r.z + rstd * (gv[k].z - xs[k].z * dot);
        o.w = r.w + rstd * (gv[k].w - xs[k].w * dot);
        *reinterpret_cast<float4*>(dx_out + t * d + c) = o;
        if (dx_bf16 != nullptr) {
          uint2 ob;
          ob.x = pack_bf16(o.x, o.y);
          ob.y = pack_bf16(o.z, o.w);
          *reinterpret_cast<uint2*>(dx_bf16 + t * d + c) = ob;
        }
      }
    }
    // the slot is free once every lane has read it: refill it with the row kStages iterations ahead
    __syncwarp();
    if (lane == 0 && t + kStages * tstep < T) {
      fence_proxy_async_smem();
      issue(t + kStages * tstep, stage);
    }
    if (++stage == kStages) {
      stage = 0;
      phase ^= 1;
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane * 4 + k * 128;
    if (c < d) {
      atomicAdd(&s_dw[c + 0], dwacc[k].x); atomicAdd(&s_dw[c + 1], dwacc[k].y);
      atomicAdd(&s_dw[c + 2], dwacc[k].z); atomicAdd(&s_dw[c + 3], dwacc[k].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) atomicAdd(dw + c, s_dw[c]);
}

template <int NV>
static int launch_rmsnorm_bwd_bulk(int grid, size_t smem, cudaStream_t s, const __nv_bfloat16* dy, long long lddy,
                                   const float* x, const float* rstd, const float* w, const float* dresid, float* dx_out,
                                   __nv_bfloat16* dx_bf16, float* dw, long long T, int d) {
  auto kern = rmsnorm_bwd_bulk_kernel<NV, 3>;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      set_error("rmsnorm_bwd: cudaFuncSetAttribute failed");
      return -2;
    }
    attr_set = true;
  }
  kern<<<grid, 256, smem, s>>>(dy, lddy, x, rstd, w, dresid, dx_out, dx_bf16, dw, T, d);
  return check_launch("rmsnorm_bwd_bulk_kernel");
}

// =============================================================================================
// GeGLU backward from the factors saved by the forward epilogue, gf = [u * gelu'(g) | gelu(g)] (bf16):
//   dgu = [dact * gf[:, 0:I] | dact * gf[:, I:2I]]
// (the stand-alone form of the fused down_proj-dgrad epilogue; used when dropout sits between the two)
// =============================================================================================
__global__ void geglu_bwd_kernel(const __nv_bfloat16* __restrict__ dact, const __nv_bfloat16* __restrict__ gf,
                                 __nv_bfloat16* __restrict__ dgu, long long T, int I) {
  const long long groups_per_row = I / 8;
  const long long total = T * groups_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = i / groups_per_row;
    const int c = static_cast<int>(i % groups_per_row) * 8;
    const uint4 da = *reinterpret_cast<const uint4*>(dact + t * I + c);
    const uint4 f1 = *reinterpret_cast<const uint4*>(gf + t * 2 * I + c);
    const uint4 f2 = *reinterpret_cast<const uint4*>(gf + t * 2 * I + I + c);
    const uint32_t dau[4] = {da.x, da.y, da.z, da.w}, f1u[4] = {f1.x, f1.y, f1.z, f1.w}, f2u[4] = {f2.x, f2.y, f2.z, f2.w};
    uint32_t og[4], ou[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = unpack_bf16(dau[j]), p = unpack_bf16(f1u[j]), q = unpack_bf16(f2u[j]);
      og[j] = pack_bf16(a.x * p.x, a.y * p.y);
      ou[j] = pack_bf16(a.x * q.x, a.y * q.y);
    }
    *reinterpret_cast<uint4*>(dgu + t * 2 * I + c) = make_uint4(og[0], og[1], og[2], og[3]);
    *reinterpret_cast<uint4*>(dgu + t * 2 * I + I + c) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
  }
}

// =============================================================================================
// Loss-head compaction (no host-side boolean indexing):
//   rows with >= 1 label, in (n,s) order -> sel_rows[M];  labelled (row, f) entries in row-major order ->
//   ent_src[L] = m*F + f (row of the [M*F, d] projected matrix), ent_label[L], ent_tok[L] = token index t.
// ref: modeling_helpers.py:263-301 (_prepare_for_stacked_feat_labels_per_mix_lvl)
// Three phases: per-block counts -> single-block scan of block totals -> write.
// =============================================================================================
constexpr int kScanBlock = 256;

__global__ void head_count_kernel(const long long* __restrict__ labels, long long T, int F, int* __restrict__ blk_rows,
                                  int* __restrict__ blk_ents) {
  const long long t = static_cast<long long>(blockIdx.x) * kScanBlock + threadIdx.x;
  int ents = 0;
  if (t < T) {
    for (int f = 0; f < F; ++f) ents += (labels[t * F + f] != -100);
  }
  int rows = ents > 0;
  __shared__ int s_r[kScanBlock / 32], s_e[kScanBlock / 32];
  int r = rows, e = ents;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r += __shfl_xor_sync(0xffffffffu, r, o);
    e += __shfl_xor_sync(0xffffffffu, e, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_r[threadIdx.x >> 5] = r;
    s_e[threadIdx.x >> 5] = e;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tr = 0, te = 0;
    for (int i = 0; i < kScanBlock / 32; ++i) {
      tr += s_r[i];
      te += s_e[i];
    }
    blk_rows[blockIdx.x] = tr;
    blk_ents[blockIdx.x] = te;
  }
}

__global__ void head_scan_blocks_kernel(int* __restrict__ blk_rows, int* __restrict__ blk_ents, int nblk,
                                        int* __restrict__ counts) {
  // single block; exclusive scan in place (nblk is small: T / 256)
  __shared__ int carry_r, carry_e;
  if (threadIdx.x == 0) {
    carry_r = 0;
    carry_e = 0;
  }
  __syncthreads();
  for (int base = 0; base < nblk; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int r = (i < nblk) ? blk_rows[i] : 0;
    int e = (i < nblk) ? blk_ents[i] : 0;
    // block-wide inclusive scan (Hillis-Steele in smem)
    __shared__ int sr[1024], se[1024];
    sr[threadIdx.x] = r;
    se[threadIdx.x] = e;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {
      int ar = 0, ae = 0;
      if (threadIdx.x >= o) {
        ar = sr[threadIdx.x - o];
        ae = se[threadIdx.x - o];
      }
      __syncthreads();
      sr[threadIdx.x] += ar;
      se[threadIdx.x] += ae;
      __syncthreads();
    }
    if (i < nblk) {
      blk_rows[i] = carry_r + sr[threadIdx.x] - r;
      blk_ents[i] = carry_e + se[threadIdx.x] - e;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) {
      carry_r += sr[threadIdx.x];
      carry_e += se[threadIdx.x];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[0] = carry_r;
    counts[1] = carry_e;
  }
}

__global__ void head_write_kernel(const long long* __restrict__ labels, long long T, int F,
                                  const int* __restrict__ blk_rows, const int* __restrict__ blk_ents,
                                  int* __restrict__ sel_rows, int* __restrict__ ent_src, int* __restrict__ ent_label,
                                  int* __restrict__ ent_tok) {
  const long long t = static_cast<long long>(blockIdx.x) * kScanBlock + threadIdx.x;
  int ents = 0;
  if (t < T) {
    for (int f = 0; f < F; ++f) ents += (labels[t * F + f] != -100);
  }
  const int rows = ents > 0;
  __shared__ int sr[kScanBlock], se[kScanBlock];
  sr[threadIdx.x] = rows;
  se[threadIdx.x] = ents;
  __syncthreads();
  for (int o = 1; o < kScanBlock; o <<= 1) {
    int ar = 0, ae = 0;
    if (threadIdx.x >= o) {
      ar = sr[threadIdx.x - o];
      ae = se[threadIdx.x - o];
    }
    __syncthreads();
    sr[threadIdx.x] += ar;
    se[threadIdx.x] += ae;
    __syncthreads();
  }
  if (t < T && rows) {
    const int m = blk_rows[blockIdx.x] + sr[threadIdx.x] - 1;
    int e = blk_ents[blockIdx.x] + se[threadIdx.x] - ents;
    sel_rows[m] = static_cast<int>(t);
    for (int f = 0; f < F; ++f) {
      const long long lab = labels[t * F + f];
      if (lab != -100) {
        ent_src[e] = m * F + f;
        ent_label[e] = static_cast<int>(lab);
        ent_tok[e] = static_cast<int>(t);
        ++e;
      }
    }
  }
}

// out[i,:] = src[idx[i],:]  (bf16 rows, 16-byte vectors); count read from device (*n_ptr) so no host sync is needed
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds, const int* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ out, long long ldo, const int* __restrict__ n_ptr,
                                   int n_max, int d) {
  const int n = n_ptr ? min(*n_ptr, n_max) : n_max;
  const int vec_per_row = d / 8;
  const long long total = static_cast<long long>(n) * vec_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    const long long s = idx ? idx[r] : r;
    *reinterpret_cast<uint4*>(out + r * ldo + c) = *reinterpret_cast<const uint4*>(src + s * lds + c);
  }
}

// out[idx[i],:] = src[i,:]   (distinct idx -> plain stores; `out` must be pre-zeroed by the caller)
__global__ void scatter_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds, const int* __restrict__ idx,
                                    __nv_bfloat16* __restrict__ out, long long ldo, const int* __restrict__ n_ptr,
                                    int n_max, int d) {
  const int n = n_ptr ? min(*n_ptr, n_max) : n_max;
  const int vec_per_row = d / 8;
  const long long total = static_cast<long long>(n) * vec_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    const long long s = idx[r];
    *reinterpret_cast<uint4*>(out + s * ldo + c) = *reinterpret_cast<const uint4*>(src + r * lds + c);
  }
}

// out[j,:] = src[i,:] if idx[i] == j for some i, else 0 — idx STRICTLY INCREASING (the compaction indices of
// head_compact are), so one kernel replaces memset + scatter: each CTA owns 256 consecutive output rows, finds the slice
// of idx that lands in them with one binary search, records it in a shared-memory map and then streams the rows.
__global__ void __launch_bounds__(256) expand_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds,
                                                          const int* __restrict__ idx, int n,
                                                          __nv_bfloat16* __restrict__ out, long long ldo,
                                                          long long n_out, int d) {
  __shared__ int map[256];
  __shared__ int i_first;
  const int vec_per_row = d / 8;
  const long long n_blocks = (n_out + 255) / 256;
  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const long long j0 = blk * 256;
    const long long left = n_out - j0;
    const int rows = left < 256 ? static_cast<int>(left) : 256;
    map[threadIdx.x] = -1;
    if (threadIdx.x == 0) {                 // first i with idx[i] >= j0
      int lo = 0, hi = n;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (idx[mid] < j0) lo = mid + 1; else hi = mid;
      }
      i_first = lo;
    }
    __syncthreads();
    {
      const int i = i_first + threadIdx.x;  // at most 256 indices can fall into 256 consecutive rows
      if (i < n) {
        const long long j = idx[i];
        if (j < j0 + rows) map[static_cast<int>(j - j0)] = i;
      }
    }
    __syncthreads();
    const int total = rows * vec_per_row;
    for (int e = threadIdx.x; e < total; e += 256) {
      const int r = e / vec_per_row, c = (e % vec_per_row) * 8;
      const int i = map[r];
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (i >= 0) v = *reinterpret_cast<const uint4*>(src + static_cast<long long>(i) * lds + c);
      *reinterpret_cast<uint4*>(out + (j0 + r) * ldo + c) = v;
    }
    __syncthreads();
  }
}

// =============================================================================================
// Cross-entropy over fp32 logits [L, ldl] (first V columns valid).  One warp per row.
//   row_lse[e] = logsumexp(logits[e,:V]);  loss_sum += wgt[e] * (row_lse[e] - logits[e, label[e]])
// focal_gamma > 0 (FocalLoss, utils_graphgpt.py:340-377): every entry is additionally weighted by (1 - p_t)^gamma with
// p_t = softmax(logits[e])[label[e]] treated as a constant (the reference detaches it: Variable(logpt.data.exp())).
// ref: modeling_helpers.py:145-198 (_get_ce_loss / _get_dlm_ce_loss on logits.float())
// =============================================================================================
__global__ void ce_fwd_kernel(const float* __restrict__ logits, long long ldl, const int* __restrict__ labels,
                              const float* __restrict__ wgt, float* __restrict__ row_lse, float* __restrict__ row_loss,
                              double* __restrict__ loss_sum, double* __restrict__ wgt_sum, int L, int V,
                              float focal_gamma, int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  double local = 0.0, local_w = 0.0;
  for (int e = blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < L; e += gridDim.x * warps_per_block) {
    const float* row = logits + static_cast<long long>(e) * ldl;
    float m = -INFINITY;
    for (int c = lane; c < V; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += __expf(row[c] - m);
    s = warp_sum(s);
    const float lse = m + logf(s);
    if (lane == 0) {
      int lab = labels[e];
      if (lab < 0 || lab >= V) {
        if (err) atomicExch(err, 2);
        lab = 0;
      }
      const float w = wgt ? wgt[e] : 1.0f;
      float l = lse - row[lab];
      if (focal_gamma > 0.f) l *= powf(fmaxf(1.0f - __expf(-l), 0.f), focal_gamma);
      row_lse[e] = lse;
      if (row_loss) row_loss[e] = l;
      local += static_cast<double>(l) * w;
      local_w += w;
    }
  }
  // block reduce of lane-0 partials
  __shared__ double s_l[32], s_w[32];
  if (lane == 0) {
    s_l[threadIdx.x >> 5] = local;
    s_w[threadIdx.x >> 5] = local_w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < warps_per_block; ++i) {
      a += s_l[i];
      b += s_w[i];
    }
    atomicAdd(loss_sum, a);
    if (wgt_sum) atomicAdd(wgt_sum, b);
  }
}

// Small vocabularies (V <= 1024, row pitch a multiple of 4 floats: the 756-entry vocabulary of the pre-training configs):
// one warp per row, the whole row in registers from up to eight 16-byte loads per lane that are all issued before the
// first use — one pass over the logits with 8 x 512 B in flight per warp, instead of two passes of scalar loads.
__global__ void ce_fwd_small_kernel(const float* __restrict__ logits, long long ldl, const int* __restrict__ labels,
                                    const float* __restrict__ wgt, float* __restrict__ row_lse, float* __restrict__ row_loss,
                                    double* __restrict__ loss_sum, double* __restrict__ wgt_sum, int L, int V,
                                    float focal_gamma, int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  double local = 0.0, local_w = 0.0;
  for (int e = blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < L; e += gridDim.x * warps_per_block) {
    const float* row = logits + static_cast<long long>(e) * ldl;
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = (k * 32 + lane) * 4;
      v[k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (c + 3 < V) {
        v[k] = *reinterpret_cast<const float4*>(row + c);
      } else if (c < V) {
        v[k].x = row[c];
        if (c + 1 < V) v[k].y = row[c + 1];
        if (c + 2 < V) v[k].z = row[c + 2];
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) m = fmaxf(m, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += (__expf(v[k].x - m) + __expf(v[k].y - m)) + (__expf(v[k].z - m) + __expf(v[k].w - m));
    s = warp_sum(s);
    const float lse = m + logf(s);
    if (lane == 0) {
      int lab = labels[e];
      if (lab < 0 || lab >= V) {
        if (err) atomicExch(err, 2);
        lab = 0;
      }
      const float w = wgt ? wgt[e] : 1.0f;
      float l = lse - row[lab];
      if (focal_gamma > 0.f) l *= powf(fmaxf(1.0f - __expf(-l), 0.f), focal_gamma);
      row_lse[e] = lse;
      if (row_loss) row_loss[e] = l;
      local += static_cast<double>(l) * w;
      local_w += w;
    }
  }
  __shared__ double s_l[32], s_w[32];
  if (lane == 0) {
    s_l[threadIdx.x >> 5] = local;
    s_w[threadIdx.x >> 5] = local_w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < warps_per_block; ++i) {
      a += s_l[i];
      b += s_w[i];
    }
    atomicAdd(loss_sum, a);
    if (wgt_sum) atomicAdd(wgt_sum, b);
  }
}

// Same idea for the gradient: all loads of the row first (16 bytes per lane each), then exp / subtract / pack, 8-byte stores.
__global__ void ce_bwd_small_kernel(const float* __restrict__ logits, long long ldl, const int* __restrict__ labels,
                                    const float* __restrict__ wgt, const float* __restrict__ row_lse,
                                    const float* __restrict__ scale, const float* __restrict__ gout,
                                    __nv_bfloat16* __restrict__ dlogits, long long ldd, int L, int V, int Vpad) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float g = scale[0] * (gout ? gout[0] : 1.0f);
  for (int e = blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < L; e += gridDim.x * warps_per_block) {
    const float* row = logits + static_cast<long long>(e) * ldl;
    __nv_bfloat16* drow = dlogits + static_cast<long long>(e) * ldd;
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = (k * 32 + lane) * 4;
      v[k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);      // exp(-inf - lse) = 0 for the pad columns
      if (c + 3 < V) {
        v[k] = *reinterpret_cast<const float4*>(row + c);
      } else if (c < V) {
        v[k].x = row[c];
        if (c + 1 < V) v[k].y = row[c + 1];
        if (c + 2 < V) v[k].z = row[c + 2];
      }
    }
    const float lse = row_lse[e];
    const int lab = labels[e];
    const float w = (wgt ? wgt[e] : 1.0f) * g;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < Vpad) {
        const float o0 = (__expf(v[k].x - lse) - (c == lab ? 1.f : 0.f)) * w;
        const float o1 = (__expf(v[k].y - lse) - (c + 1 == lab ? 1.f : 0.f)) * w;
        const float o2 = (__expf(v[k].z - lse) - (c + 2 == lab ? 1.f : 0.f)) * w;
        const float o3 = (__expf(v[k].w - lse) - (c + 3 == lab ? 1.f : 0.f)) * w;
        *reinterpret_cast<uint2*>(drow + c) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
      }
    }
  }
}

// dlogits[e,c] = (softmax(logits[e])[c] - [c == label[e]]) * wgt[e] * scale[0] * gout[0]   (bf16, pad columns = 0)
// (FOCAL is a template flag so that powf's slow path costs the plain cross-entropy kernel no registers)
template <bool FOCAL>
__global__ void ce_bwd_kernel(const float* __restrict__ logits, long long ldl, const int* __restrict__ labels,
                              const float* __restrict__ wgt, const float* __restrict__ row_lse,
                              const float* __restrict__ scale, const float* __restrict__ gout,
                              __nv_bfloat16* __restrict__ dlogits, long long ldd, int L, int V, int Vpad,
                              float focal_gamma) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float g = scale[0] * (gout ? gout[0] : 1.0f);
  for (int e = blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < L; e += gridDim.x * warps_per_block) {
    const float* row = logits + static_cast<long long>(e) * ldl;
    __nv_bfloat16* drow = dlogits + static_cast<long long>(e) * ldd;
    const float lse = row_lse[e];
    const int lab = labels[e];
    float w = (wgt ? wgt[e] : 1.0f) * g;
    if (FOCAL) w *= powf(fmaxf(1.0f - __expf(row[lab] - lse), 0.f), focal_gamma);   // detached (1 - p_t)^gamma
    for (int c = lane * 2; c < Vpad; c += 64) {
      float v0 = 0.f, v1 = 0.f;
      if (c < V) v0 = (__expf(row[c] - lse) - (c == lab ? 1.f : 0.f)) * w;
      if (c + 1 < V) v1 = (__expf(row[c + 1] - lse) - (c + 1 == lab ? 1.f : 0.f)) * w;
      *reinterpret_cast<uint32_t*>(drow + c) = pack_bf16(v0, v1);
    }
  }
}

// loss = loss_sum / denom  where denom = L (mean), wgt_sum + 1e-7, or a fixed divisor; also writes scale = 1/denom
__global__ void ce_finalize_kernel(const double* __restrict__ loss_sum, const double* __restrict__ wgt_sum,
                                   const int* __restrict__ count, int mode, float fixed_denom, float* __restrict__ loss,
                                   float* __restrict__ scale) {
  double denom;
  if (mode == 0) denom = static_cast<double>(*count);              // mean over L labelled entries
  else if (mode == 1) denom = *wgt_sum + 1e-7;                    // weighted mean
  else denom = static_cast<double>(fixed_denom);                  // dLM: N*S*F
  if (denom <= 0.0) {
    *loss = nanf("");  // CrossEntropyLoss over zero targets is NaN in torch
    *scale = 0.f;
  } else {
    *loss = static_cast<float>(*loss_sum / denom);
    *scale = static_cast<float>(1.0 / denom);
  }
}

// =============================================================================================
// Optimiser: squared-norm of the flat gradient, then fused AdamW with optional clipping, writing the fp32 master
// and its bf16 compute copy in one pass.  ref: training_utils.py:71-86 (clip_grad_norm_ + AdamW),
// configs/training/base.yaml:35-44 (betas 0.9/0.95, wd 0.1, clip 1.0)
// =============================================================================================
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  double acc = 0.0;
  const long long n4 = n / 4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc += static_cast<double>(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) acc += static_cast<double>(g[i]) * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += s[i];
    atomicAdd(out, t);
  }
}

__global__ void adamw_kernel(float* __restrict__ p, __nv_bfloat16* __restrict__ p_bf16, const float* __restrict__ g,
                             float* __restrict__ m, float* __restrict__ v, long long n, float lr, float beta1,
                             float beta2, float eps, float wd, float bc1, float bc2, const double* __restrict__ gnorm_sq,
                             float max_norm, float grad_scale) {
  float clip = grad_scale;
  if (gnorm_sq != nullptr && max_norm > 0.f) {
    const float norm = sqrtf(static_cast<float>(*gnorm_sq)) * grad_scale;
    clip = grad_scale * fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * clip;
    float pi = p[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= lr * wd * pi;                                   // decoupled weight decay (torch.optim.AdamW)
    pi -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    p[i] = pi;
    if (p_bf16 != nullptr) p_bf16[i] = __float2bfloat16_rn(pi);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long n4 = n / 4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n4 = n / 4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint2 v = reinterpret_cast<const uint2*>(src)[i];
    const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y);
    reinterpret_cast<float4*>(dst)[i] = make_float4(a.x, a.y, b.x, b.y);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) dst[i] = __bfloat162float(src[i]);
}

}  // namespace ggpt

using namespace ggpt;

extern "C" {

int ggpt_embed_fwd(const long long* ids, const float* table, const float* gate, float* out, long long T, int F, int d,
                   int V, int long_scale, int* err_flag, float drop_p, unsigned long long drop_seed, void* stream) {
  GGPT_REQUIRE(ids && table && out, "embed_fwd: null pointer");
  GGPT_REQUIRE(T > 0 && F > 0 && d > 0 && d % 4 == 0, "embed_fwd: bad sizes T=%lld F=%d d=%d", T, F, d);
  GGPT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "embed_fwd: dropout p=%f outside [0,1)", drop_p);
  embed_fwd_kernel<<<grid_for_rows(T, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, table, gate, out, T, F, d, V, long_scale, err_flag, make_drop_params(drop_p, drop_seed));
  return check_launch("embed_fwd_kernel");
}

int ggpt_embed_bwd(const long long* ids, const float* dx, const float* table, const float* gate, float* dtable,
                   float* dgate, long long T, int F, int d, int V, int padding_idx, int long_scale, float drop_p,
                   unsigned long long drop_seed, void* stream) {
  GGPT_REQUIRE(ids && dx && dtable, "embed_bwd: null pointer");
  GGPT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "embed_bwd: dropout p=%f outside [0,1)", drop_p);
  GGPT_REQUIRE(gate == nullptr || (table && dgate), "embed_bwd: gated aggregation needs table and dgate");
  embed_bwd_kernel<<<grid_for_rows(T, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, dx, table, gate, dtable, dgate, T, F, d, V, padding_idx, long_scale, make_drop_params(drop_p, drop_seed));
  return check_launch("embed_bwd_kernel");
}

int ggpt_embed_count(const long long* ids, void* cnt, long long ldc, long long T, int F, int V, int padding_idx,
                     void* stream) {
  GGPT_REQUIRE(ids && cnt, "embed_count: null pointer");
  GGPT_REQUIRE(T > 0 && F > 0 && F <= 32 && ldc % 8 == 0 && ldc >= V, "embed_count: bad sizes T=%lld F=%d V=%d ldc=%lld", T, F, V, ldc);
  embed_count_kernel<<<grid_for_rows(T, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, static_cast<__nv_bfloat16*>(cnt), ldc, T, F, V, padding_idx);
  return check_launch("embed_count_kernel");
}

int ggpt_smtp_mask_2d(const long long* ids, int Ftot, int node_col, const float* mr, const float* u_node, float power,
                      long long* out_ids, long long* labels, int N, int S, int F, long long mask_token,
                      long long label_pad, int* err_flag, void* stream) {
  GGPT_REQUIRE(ids && mr && u_node && out_ids && labels, "smtp_mask_2d: null pointer");
  GGPT_REQUIRE(N > 0 && S > 0 && F > 0 && Ftot >= F && node_col >= 0 && node_col < Ftot,
               "smtp_mask_2d: bad sizes N=%d S=%d F=%d Ftot=%d node_col=%d", N, S, F, Ftot, node_col);
  const long long total = static_cast<long long>(N) * S * F;
  smtp_mask_2d_kernel<<<grid_for_rows(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, Ftot, node_col, mr, u_node, power, out_ids, labels, N, S, F, mask_token, label_pad, err_flag);
  return check_launch("smtp_mask_2d_kernel");
}

int ggpt_rmsnorm_fwd(const float* x, const float* w, void* y, long long ldy, float* rstd, long long T, int d, float eps,
                     void* stream) {
  GGPT_REQUIRE(x && w && y, "rmsnorm_fwd: null pointer");
  GGPT_REQUIRE(T > 0 && d % 4 == 0 && ldy % 4 == 0, "rmsnorm_fwd: bad sizes");
  rmsnorm_fwd_kernel<<<grid_for_rows(T, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, w, static_cast<__nv_bfloat16*>(y), ldy, rstd, T, d, eps);
  return check_launch("rmsnorm_fwd_kernel");
}

int ggpt_add_rmsnorm_fwd(const float* x_in, const void* y, long long ldy, const float* colscale, const float* rowscale,
                         const float* w, float* x_out, void* h, float* rstd, long long T, int d, float eps, void* stream) {
  GGPT_REQUIRE(x_in && y && w && x_out, "add_rmsnorm_fwd: null pointer");
  GGPT_REQUIRE(T > 0 && d % 4 == 0 && ldy % 4 == 0, "add_rmsnorm_fwd: bad sizes");
  static const bool generic_only = getenv("GGPT_ADD_RMSNORM_GENERIC") != nullptr;     // A/B aid
  if (!generic_only && (d == 768 || d == 1024 || d == 512 || d == 256)) {
    const dim3 g(grid_for_rows(T, 8));
    const __nv_bfloat16* yb = static_cast<const __nv_bfloat16*>(y);
    __nv_bfloat16* hb = static_cast<__nv_bfloat16*>(h);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (d == 768) add_rmsnorm_fwd_reg_kernel<6><<<g, 256, 0, s>>>(x_in, yb, ldy, colscale, rowscale, w, x_out, hb, rstd, T, eps);
    else if (d == 1024) add_rmsnorm_fwd_reg_kernel<8><<<g, 256, 0, s>>>(x_in, yb, ldy, colscale, rowscale, w, x_out, hb, rstd, T, eps);
    else if (d == 512) add_rmsnorm_fwd_reg_kernel<4><<<g, 256, 0, s>>>(x_in, yb, ldy, colscale, rowscale, w, x_out, hb, rstd, T, eps);
    else add_rmsnorm_fwd_reg_kernel<2><<<g, 256, 0, s>>>(x_in, yb, ldy, colscale, rowscale, w, x_out, hb, rstd, T, eps);
    return check_launch("add_rmsnorm_fwd_reg_kernel");
  }
  add_rmsnorm_fwd_kernel<<<grid_for_rows(T, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_in, static_cast<const __nv_bfloat16*>(y), ldy, colscale, rowscale, w, x_out, static_cast<__nv_bfloat16*>(h), rstd, T,
      d, eps);
  return check_launch("add_rmsnorm_fwd_kernel");
}

int ggpt_rmsnorm_bwd(const void* dy, long long lddy, const float* x, const float* rstd, const float* w,
                     const float* dresid, float* dx_out, void* dx_bf16, float* dw, long long T, int d, void* stream) {
  GGPT_REQUIRE(dy && x && rstd && w && dx_out && dw, "rmsnorm_bwd: null pointer");
  GGPT_REQUIRE(T > 0 && d % 4 == 0 && lddy % 4 == 0, "rmsnorm_bwd: bad sizes");
  GGPT_REQUIRE(d <= 2048, "rmsnorm_bwd: hidden size %d > 2048 is not instantiated", d);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* dyb = static_cast<const __nv_bfloat16*>(dy);
  __nv_bfloat16* dxb = static_cast<__nv_bfloat16*>(dx_bf16);
  // bulk-copy pipelined version: 16-byte aligned rows, the per-warp rings must fit in shared memory
  {
    auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const size_t head = ((4 * static_cast<size_t>(d) + 127) / 128) * 128 + 128 * ((8 * 3 * 8 + 127) / 128);
    const size_t smem3 = head + static_cast<size_t>(8) * 3 * 10 * d;
    static const bool no_bulk = getenv("GGPT_RMSNORM_BWD_NO_BULK") != nullptr;
    if (!no_bulk && d % 8 == 0 && lddy % 8 == 0 && aligned16(x) && aligned16(dy) && (dresid == nullptr || aligned16(dresid)) &&
        smem3 <= 200 * 1024 && (d == 768 || d == 1024 || d == 512 || d == 256)) {
      const int grid_b = static_cast<int>(std::min<long long>((T + 7) / 8, num_sms()));
      if (d == 256) return launch_rmsnorm_bwd_bulk<2>(grid_b, smem3, s, dyb, lddy, x, rstd, w, dresid, dx_out, dxb, dw, T, d);
      if (d == 512) return launch_rmsnorm_bwd_bulk<4>(grid_b, smem3, s, dyb, lddy, x, rstd, w, dresid, dx_out, dxb, dw, T, d);
      if (d == 768) return launch_rmsnorm_bwd_bulk<6>(grid_b, smem3, s, dyb, lddy, x, rstd, w, dresid, dx_out, dxb, dw, T, d);
      return launch_rmsnorm_bwd_bulk<8>(grid_b, smem3, s, dyb, lddy, x, rstd, w, dresid, dx_out, dxb, dw, T, d);
    }
  }
  const int grid = grid_for_rows(T, 4, 12);
  const size_t sm = d * sizeof(float);
  const int nv = (d + 127) / 128;
  if (nv <= 1) rmsnorm_bwd_kernel<1><<<grid, 128, sm, s>>>(dyb, lddy, x, rstd, w, dresid, dx_out, dxb, dw, T, d);
  else if (nv <= 2) rmsnorm_bwd_kernel<2><<<grid, 128, sm, s>>>(dyb, lddy, x, rstd, w, dresid, dx_out, dxb, dw, T, d);
  else if (nv <= 4) rmsnorm_bwd_kernel<4><<<grid, 128, sm, s>>>(dyb, lddy, x, rstd, w, dresid, dx_out, dxb, dw, T, d);
  else if (nv <= 6) rmsnorm_bwd_kernel<6><<<grid, 128, sm, s>>>(dyb, lddy, x, rstd, w, dresid, dx_out, dxb, dw, T, d);
  else if (nv <= 8) rmsnorm_bwd_kernel<8><<<grid, 128, sm, s>>>(dyb, lddy, x, rstd, w, dresid, dx_out, dxb, dw, T, d);
  else rmsnorm_bwd_kernel<16><<<grid, 128, sm, s>>>(dyb, lddy, x, rstd, w, dresid, dx_out, dxb, dw, T, d);
  return check_launch("rmsnorm_bwd_kernel");
}

int ggpt_geglu_bwd(const void* dact, const void* gu, void* dgu, long long T, int I, void* stream) {
  GGPT_REQUIRE(dact && gu && dgu, "geglu_bwd: null pointer");
  GGPT_REQUIRE(T > 0 && I % 8 == 0, "geglu_bwd: bad sizes");
  const long long total = T * (I / 8);
  geglu_bwd_kernel<<<grid_for_rows(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dact), static_cast<const __nv_bfloat16*>(gu), static_cast<__nv_bfloat16*>(dgu), T,
      I);
  return check_launch("geglu_bwd_kernel");
}

long long ggpt_head_scratch_ints(long long T) { return 2 * ((T + kScanBlock - 1) / kScanBlock); }

int ggpt_head_compact(const long long* labels, long long T, int F, int* scratch, int* counts, int* sel_rows,
                      int* ent_src, int* ent_label, int* ent_tok, void* stream) {
  GGPT_REQUIRE(labels && scratch && counts && sel_rows && ent_src && ent_label && ent_tok, "head_compact: null pointer");
  GGPT_REQUIRE(T > 0 && F > 0 && T * F < (1ll << 31), "head_compact: bad sizes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nblk = static_cast<int>((T + kScanBlock - 1) / kScanBlock);
  int* blk_rows = scratch;
  int* blk_ents = scratch + nblk;
  head_count_kernel<<<nblk, kScanBlock, 0, s>>>(labels, T, F, blk_rows, blk_ents);
  if (int rc = check_launch("head_count_kernel")) return rc;
  head_scan_blocks_kernel<<<1, 1024, 0, s>>>(blk_rows, blk_ents, nblk, counts);
  if (int rc = check_launch("head_scan_blocks_kernel")) return rc;
  head_write_kernel<<<nblk, kScanBlock, 0, s>>>(labels, T, F, blk_rows, blk_ents, sel_rows, ent_src, ent_label, ent_tok);
  return check_launch("head_write_kernel");
}

int ggpt_gather_rows(const void* src, long long lds, const int* idx, void* out, long long ldo, const int* n_ptr,
                     int n_max, int d, void* stream) {
  GGPT_REQUIRE(src && out, "gather_rows: null pointer");
  GGPT_REQUIRE(d % 8 == 0 && lds % 8 == 0 && ldo % 8 == 0, "gather_rows: d/ld must be multiples of 8");
  if (n_max <= 0) return 0;
  const long long total = static_cast<long long>(n_max) * (d / 8);
  gather_rows_kernel<<<grid_for_rows(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), lds, idx, static_cast<__nv_bfloat16*>(out), ldo, n_ptr, n_max, d);
  return check_launch("gather_rows_kernel");
}

int ggpt_scatter_rows(const void* src, long long lds, const int* idx, void* out, long long ldo, const int* n_ptr,
                      int n_max, int d, void* stream) {
  GGPT_REQUIRE(src && out && idx, "scatter_rows: null pointer");
  GGPT_REQUIRE(d % 8 == 0 && lds % 8 == 0 && ldo % 8 == 0, "scatter_rows: d/ld must be multiples of 8");
  if (n_max <= 0) return 0;
  const long long total = static_cast<long long>(n_max) * (d / 8);
  scatter_rows_kernel<<<grid_for_rows(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), lds, idx, static_cast<__nv_bfloat16*>(out), ldo, n_ptr, n_max, d);
  return check_launch("scatter_rows_kernel");
}

int ggpt_expand_rows(const void* src, long long lds, const int* idx, int n, void* out, long long ldo, long long n_out, int d,
                     void* stream) {
  GGPT_REQUIRE(out && (n == 0 || (src && idx)), "expand_rows: null pointer");
  GGPT_REQUIRE(n >= 0 && n_out > 0 && d % 8 == 0 && lds % 8 == 0 && ldo % 8 == 0, "expand_rows: bad sizes n=%d n_out=%lld d=%d", n,
               n_out, d);
  const long long blocks = (n_out + 255) / 256;
  expand_rows_kernel<<<grid_for_rows(blocks, 1, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), lds, idx, n, static_cast<__nv_bfloat16*>(out), ldo, n_out, d);
  return check_launch("expand_rows_kernel");
}

int ggpt_ce_fwd(const float* logits, long long ldl, const int* labels, const float* wgt, float* row_lse, float* row_loss,
                double* loss_sum, double* wgt_sum, int L, int V, float focal_gamma, int* err_flag, void* stream) {
  GGPT_REQUIRE(logits && labels && row_lse && loss_sum, "ce_fwd: null pointer");
  if (L <= 0) return 0;
  if (V <= 1024 && ldl % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {
    ce_fwd_small_kernel<<<grid_for_rows(L, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, ldl, labels, wgt, row_lse, row_loss, loss_sum, wgt_sum, L, V, focal_gamma, err_flag);
    return check_launch("ce_fwd_small_kernel");
  }
  ce_fwd_kernel<<<grid_for_rows(L, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, ldl, labels, wgt, row_lse,
                                                                                    row_loss, loss_sum, wgt_sum, L, V,
                                                                                    focal_gamma, err_flag);
  return check_launch("ce_fwd_kernel");
}

int ggpt_ce_finalize(const double* loss_sum, const double* wgt_sum, const int* count, int mode, float fixed_denom,
                     float* loss, float* scale, void* stream) {
  GGPT_REQUIRE(loss_sum && loss && scale, "ce_finalize: null pointer");
  GGPT_REQUIRE(mode == 2 || (mode == 0 && count) || (mode == 1 && wgt_sum), "ce_finalize: bad mode %d", mode);
  ce_finalize_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(loss_sum, wgt_sum, count, mode, fixed_denom, loss,
                                                                     scale);
  return check_launch("ce_finalize_kernel");
}

int ggpt_ce_bwd(const float* logits, long long ldl, const int* labels, const float* wgt, const float* row_lse,
                const float* scale, const float* gout, void* dlogits, long long ldd, int L, int V, float focal_gamma,
                void* stream) {
  GGPT_REQUIRE(logits && labels && row_lse && scale && dlogits, "ce_bwd: null pointer");
  GGPT_REQUIRE(ldd % 8 == 0 && ldd >= ((V + 7) / 8) * 8, "ce_bwd: ldd must be a multiple of 8 covering V");
  if (L <= 0) return 0;
  const int Vpad8 = ((V + 7) / 8) * 8;
  if (focal_gamma <= 0.f && V <= 1024 && ldl % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dlogits) & 7) == 0) {
    ce_bwd_small_kernel<<<grid_for_rows(L, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, ldl, labels, wgt, row_lse, scale, gout, static_cast<__nv_bfloat16*>(dlogits), ldd, L, V, Vpad8);
    return check_launch("ce_bwd_small_kernel");
  }
  if (focal_gamma > 0.f)
    ce_bwd_kernel<true><<<grid_for_rows(L, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, ldl, labels, wgt, row_lse, scale, gout, static_cast<__nv_bfloat16*>(dlogits), ldd, L, V,
        ((V + 7) / 8) * 8, focal_gamma);
  else
    ce_bwd_kernel<false><<<grid_for_rows(L, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, ldl, labels, wgt, row_lse, scale, gout, static_cast<__nv_bfloat16*>(dlogits), ldd, L, V,
        ((V + 7) / 8) * 8, focal_gamma);
  return check_launch("ce_bwd_kernel");
}

int ggpt_sumsq(const float* g, long long n, double* out, void* stream) {
  GGPT_REQUIRE(g && out && n > 0, "sumsq: bad arguments");
  GGPT_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "sumsq: pointer must be 16-byte aligned");
  sumsq_kernel<<<grid_for_rows(n / 4 + 1, 256, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, n, out);
  return check_launch("sumsq_kernel");
}

int ggpt_adamw(float* p, void* p_bf16, const float* g, float* m, float* v, long long n, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, const double* gnorm_sq, float max_norm,
               float grad_scale, void* stream) {
  GGPT_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adamw: bad arguments");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adamw_kernel<<<grid_for_rows(n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, static_cast<__nv_bfloat16*>(p_bf16), g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, gnorm_sq,
      max_norm, grad_scale);
  return check_launch("adamw_kernel");
}

int ggpt_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  GGPT_REQUIRE(src && dst && n > 0, "cast: bad arguments");
  GGPT_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
               "cast: pointers must be 16/8-byte aligned");
  cast_f32_bf16_kernel<<<grid_for_rows(n / 4 + 1, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n);
  return check_launch("cast_f32_bf16_kernel");
}

int ggpt_cast_bf16_f32(const void* src, float* dst, long long n, void* stream) {
  GGPT_REQUIRE(src && dst && n > 0, "cast: bad arguments");
  GGPT_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0,
               "cast: pointers must be 16/8-byte aligned");
  cast_bf16_f32_kernel<<<grid_for_rows(n / 4 + 1, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), dst, n);
  return check_launch("cast_bf16_f32_kernel");
}

}  // extern "C"
