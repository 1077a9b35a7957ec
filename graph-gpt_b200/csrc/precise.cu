// Validation-only PRECISE mode (GGPT_PRECISE=1; never used by bench.py or a training step).
//
// Purpose: separate the bf16 storage rounding of the fast path from any arithmetic error of similar size.  In this mode
// every GEMM still runs through the SAME tcgen05 kernel (gemm_sm100.cu: same TMA maps, descriptors, MMA issue, TMEM
// epilogue), but on split-bf16 operands: an fp32 matrix X is written as X = hi + lo + O(2^-17 |X|) with hi = bf16(X),
// lo = bf16(X - hi), and the product A B^T ~= A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T is ONE GEMM over the concatenated
// reduction dimension K' = 3K:   A' = [A_hi | A_lo | A_hi],   B' = [B_hi | B_hi | B_lo]   (ggpt_vp_split3).
// Everything between the GEMMs (RMSNorm, RoPE, softmax attention, GeGLU, residual adds) runs in plain fp32 SIMT kernels
// below, straight from the reference formulas, and activations stay fp32.  The mode must reproduce the fp32 goldens of
// the reference to <= 1e-3 (tests/test_precise_mode_gpu.py measures ~1e-5); the fast path's 3..6e-3 on logits is then,
// by elimination, bf16 storage rounding.
//
// ref: HF:59-64 (RMSNorm), HF:138-168 (rotate-half RoPE), HF:199-221 (attention), HF:182-184 (GeGLU, erf GELU),
//      HF:325,331 + utils_graphgpt.py:153-166 (residual adds with LayerScale / DropPath scales),
//      modeling_pretrain.py:134-145 / modeling_helpers.py:127-139 (raw-embedding mask-token swap + norm).
#include "common.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

static inline unsigned vp_grid(long long n, int per_block) {
  long long b = (n + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

// out[r, 0:C | C:2C | 2C:3C] = role 0 (A operand): hi | lo | hi ;  role 1 (B operand): hi | hi | lo
__global__ void vp_split3_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out, long long ldo,
                                 long long R, int C, int role) {
  const long long total = R * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C;
    const int c = static_cast<int>(i % C);
    const float v = x[r * ldx + c];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + r * ldo + c;
    o[0] = hi;
    o[C] = role == 0 ? lo : hi;
    o[2 * C] = role == 0 ? hi : lo;
  }
}

// y = w * (x * rsqrt(mean(x^2) + eps)), all fp32 (one warp per row)
__global__ void vp_rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                  long long T, int d, float eps) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * wpb) {
    const float* xr = x + t * d;
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) ss += xr[c] * xr[c];
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / static_cast<float>(d) + eps);
    for (int c = lane; c < d; c += 32) y[t * d + c] = w[c] * (xr[c] * rstd);
  }
}

// raw-embedding branch: row keeps its raw features when any of its first fchk labels is -100 (or labels == NULL), else
// takes mask_tok; then RMSNorm over E
__global__ void vp_raw_embed_kernel(const float* __restrict__ raw, const long long* __restrict__ labels, long long ldl,
                                    int fchk, const float* __restrict__ mask_tok, const float* __restrict__ w,
                                    float* __restrict__ y, long long T, int E, float eps) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * wpb) {
    bool keep = true;
    if (labels != nullptr) {
      keep = false;
      for (int f = 0; f < fchk; ++f) keep |= (labels[t * ldl + f] == -100);
    }
    const float* xr = keep ? raw + t * E : mask_tok;
    float ss = 0.f;
    for (int c = lane; c < E; c += 32) ss += xr[c] * xr[c];
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / static_cast<float>(E) + eps);
    for (int c = lane; c < E; c += 32) y[t * E + c] = w[c] * (xr[c] * rstd);
  }
}

// in-place rotate-half RoPE on columns [0, rope_cols) of qkv (heads of 64): (x1, x2) -> (x1 c - x2 s, x2 c + x1 s)
__global__ void vp_rope_kernel(float* __restrict__ qkv, long long ld, const int* __restrict__ pos,
                               const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, long long T,
                               int rope_cols) {
  const int pairs = rope_cols / 2;
  const long long total = T * pairs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = i / pairs;
    const int pi = static_cast<int>(i % pairs);
    const int head = pi / 32, j = pi % 32;
    const long long p = pos[t];
    const float c = cos_tab[p * 32 + j], s = sin_tab[p * 32 + j];
    float* row = qkv + t * ld + head * 64;
    const float x1 = row[j], x2 = row[j + 32];
    row[j] = x1 * c - x2 * s;
    row[j + 32] = x2 * c + x1 * s;
  }
}

// fp32 attention: one warp per (n, h, q).  Scores are kept in shared memory (S floats per warp), softmax in fp32 with the
// reference's formula; keys whose mask bit is clear are excluded; a fully masked row yields 0 (as the fast kernels).
__global__ void vp_attn_kernel(const float* __restrict__ qkv, long long ld, int q_col0, int k_col0, int v_col0,
                               const uint32_t* __restrict__ mask_bits, int mask_words, float* __restrict__ out,
                               long long ldo, int N, int S, int H) {
  extern __shared__ float s_sc[];   // [warps][S]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* sc = s_sc + static_cast<size_t>(wib) * S;
  const long long total = static_cast<long long>(N) * H * S;
  for (long long item = static_cast<long long>(blockIdx.x) * wpb + wib; item < total;
       item += static_cast<long long>(gridDim.x) * wpb) {
    const int q = static_cast<int>(item % S);
    const int h = static_cast<int>((item / S) % H);
    const int n = static_cast<int>(item / (static_cast<long long>(S) * H));
    const float* qrow = qkv + (static_cast<long long>(n) * S + q) * ld + q_col0 + h * 64;
    const uint32_t* mrow = mask_bits + (static_cast<long long>(n) * S + q) * mask_words;
    float mx = -INFINITY;
    for (int k = lane; k < S; k += 32) {
      float s = -INFINITY;
      if ((mrow[k >> 5] >> (k & 31)) & 1u) {
        const float* krow = qkv + (static_cast<long long>(n) * S + k) * ld + k_col0 + h * 64;
        float acc = 0.f;
#pragma unroll 8
        for (int c = 0; c < 64; ++c) acc = fmaf(qrow[c], krow[c], acc);
        s = acc * 0.125f;
      }
      sc[k] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < S; k += 32) {
      const float e = (sc[k] == -INFINITY) ? 0.f : expf(sc[k] - mx);
      sc[k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
    float o0 = 0.f, o1 = 0.f;       // lane owns output columns lane and lane + 32
    for (int k = 0; k < S; ++k) {
      const float p = sc[k];
      if (p != 0.f) {
        const float* vrow = qkv + (static_cast<long long>(n) * S + k) * ld + v_col0 + h * 64;
        o0 = fmaf(p, vrow[lane], o0);
        o1 = fmaf(p, vrow[lane + 32], o1);
      }
    }
    float* orow = out + (static_cast<long long>(n) * S + q) * ldo + h * 64;
    orow[lane] = o0 * inv;
    orow[lane + 32] = o1 * inv;
    __syncwarp();
  }
}

// act[t, i] = gelu_erf(gu[t, i]) * gu[t, I + i]   (exact erf)
__global__ void vp_geglu_kernel(const float* __restrict__ gu, long long ldgu, float* __restrict__ act, long long T, int I) {
  const long long total = T * I;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = i / I;
    const int c = static_cast<int>(i % I);
    const float g = gu[t * ldgu + c], u = gu[t * ldgu + I + c];
    act[i] = 0.5f * g * (1.0f + erff(g * 0.70710678118654752f)) * u;
  }
}

// x_out = x_in + rowscale[t] * colscale[c] * y   (scales optional)
__global__ void vp_add_kernel(const float* __restrict__ x_in, const float* __restrict__ y, long long ldy,
                              const float* __restrict__ colscale, const float* __restrict__ rowscale,
                              float* __restrict__ x_out, long long T, int d) {
  const long long total = T * d;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = i / d;
    const int c = static_cast<int>(i % d);
    float v = y[t * ldy + c];
    if (colscale != nullptr) v = colscale[c] * v;
    if (rowscale != nullptr) v = v * rowscale[t];
    x_out[i] = x_in[i] + v;
  }
}

__global__ void vp_gather_rows_kernel(const float* __restrict__ src, long long lds, const int* __restrict__ idx,
                                      float* __restrict__ out, long long n, int d) {
  const long long total = n * d;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / d;
    out[i] = src[static_cast<long long>(idx[r]) * lds + (i % d)];
  }
}

}  // namespace ggpt

using namespace ggpt;

extern "C" {

int ggpt_vp_split3(const float* x, long long ldx, void* out, long long ldo, long long R, int C, int role, void* stream) {
  GGPT_REQUIRE(x && out && R > 0 && C > 0 && ldo >= 3LL * C && (role == 0 || role == 1), "vp_split3: bad arguments");
  vp_split3_kernel<<<vp_grid(R * C, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, static_cast<__nv_bfloat16*>(out), ldo, R, C, role);
  return check_launch("vp_split3_kernel");
}

int ggpt_vp_rmsnorm_f32(const float* x, const float* w, float* y, long long T, int d, float eps, void* stream) {
  GGPT_REQUIRE(x && w && y && T > 0 && d > 0, "vp_rmsnorm: bad arguments");
  vp_rmsnorm_kernel<<<vp_grid(T, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, y, T, d, eps);
  return check_launch("vp_rmsnorm_kernel");
}

int ggpt_vp_raw_embed_f32(const float* raw, const long long* labels, long long ldl, int fchk, const float* mask_tok,
                          const float* w, float* y, long long T, int E, float eps, void* stream) {
  GGPT_REQUIRE(raw && w && y && T > 0 && E > 0, "vp_raw_embed: bad arguments");
  GGPT_REQUIRE(labels == nullptr || (mask_tok != nullptr && fchk > 0), "vp_raw_embed: labels need mask_tok and fchk > 0");
  vp_raw_embed_kernel<<<vp_grid(T, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(raw, labels, ldl, fchk, mask_tok, w, y,
                                                                                    T, E, eps);
  return check_launch("vp_raw_embed_kernel");
}

int ggpt_vp_rope_f32(float* qkv, long long ld, const int* pos, const float* cos_tab, const float* sin_tab, long long T,
                     int rope_cols, void* stream) {
  GGPT_REQUIRE(qkv && pos && cos_tab && sin_tab && T > 0 && rope_cols % 64 == 0, "vp_rope: bad arguments");
  vp_rope_kernel<<<vp_grid(T * (rope_cols / 2), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(qkv, ld, pos, cos_tab,
                                                                                                    sin_tab, T, rope_cols);
  return check_launch("vp_rope_kernel");
}

int ggpt_vp_attn_f32(const float* qkv, long long ld, int q_col0, int k_col0, int v_col0, const uint32_t* mask_bits,
                     float* out, long long ldo, int N, int S, int H, void* stream) {
  GGPT_REQUIRE(qkv && mask_bits && out && N > 0 && S > 0 && H > 0, "vp_attn: bad arguments");
  const int warps = 4;
  const size_t smem = static_cast<size_t>(warps) * S * sizeof(float);
  GGPT_REQUIRE(smem <= 200 * 1024, "vp_attn: S=%d too long for the validation kernel", S);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(vp_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  vp_attn_kernel<<<vp_grid(static_cast<long long>(N) * S * H, warps), warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      qkv, ld, q_col0, k_col0, v_col0, mask_bits, ggpt_attn_mask_words(S), out, ldo, N, S, H);
  return check_launch("vp_attn_kernel");
}

int ggpt_vp_geglu_f32(const float* gu, long long ldgu, float* act, long long T, int I, void* stream) {
  GGPT_REQUIRE(gu && act && T > 0 && I > 0 && ldgu >= 2LL * I, "vp_geglu: bad arguments");
  vp_geglu_kernel<<<vp_grid(T * I, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(gu, ldgu, act, T, I);
  return check_launch("vp_geglu_kernel");
}

int ggpt_vp_add_f32(const float* x_in, const float* y, long long ldy, const float* colscale, const float* rowscale,
                    float* x_out, long long T, int d, void* stream) {
  GGPT_REQUIRE(x_in && y && x_out && T > 0 && d > 0, "vp_add: bad arguments");
  vp_add_kernel<<<vp_grid(T * d, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x_in, y, ldy, colscale, rowscale, x_out,
                                                                                    T, d);
  return check_launch("vp_add_kernel");
}

int ggpt_vp_gather_rows_f32(const float* src, long long lds, const int* idx, float* out, long long n, int d, void* stream) {
  GGPT_REQUIRE(src && idx && out && n > 0 && d > 0, "vp_gather_rows: bad arguments");
  vp_gather_rows_kernel<<<vp_grid(n * d, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, lds, idx, out, n, d);
  return check_launch("vp_gather_rows_kernel");
}

}  // extern "C"
