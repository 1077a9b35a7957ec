// Fine-tuning head in ONE kernel per direction: last-valid-row pooling, the `score` head
// (nn.Linear or the MLP of modules_utils.py:8-34) and the task loss with its backward.
//
// The reference scores EVERY position (`logits = self.score(hidden_states)`, modeling_finetune.py:281) and then indexes
// one row per sample (`pooled_logits = logits[arange(N), sequence_lengths]`, :289-296, sequence_lengths =
// ne(in_, pad).sum(-1) - 1, modeling_helpers.py:78-86); only that row reaches the loss (:167-234), so the head is
// [N, d] work, not [N*S, d]:  one CTA per sample
//   forward : count the non-pad ids of the sample -> row index; copy the pooled hidden row (returned as
//             task_hidden_states); x -> [gelu -> dropout ->] Linear ... (fp32, weights read from the fp32 masters);
//             per-sample loss term + denominator term; a one-CTA finalize adds them in a fixed order
//   backward: dlogits from the loss, back through the layers (weight / bias gradients by atomicAdd into the flat
//             gradient buffer), the gradient of the pooled row lands in the (pre-zeroed) [T, d] hidden gradient.
// Loss modes: 0 CE mean, 1 CE weighted by sample_wgt (sum l w / sum w), 2 MSE, 3 L1 (mean over all elements),
//             4 BCE-with-logits over the labelled (non-NaN) entries; -1 = no labels (logits only).
#include "common.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

constexpr int kFtMaxLayers = 4;
constexpr int kFtMaxDim = 2048;
constexpr int kFtThreads = 256;

struct FtHeadParams {
  const __nv_bfloat16* hidden;   // [N*S, ld]
  long long ld;
  const long long* ids;          // first-feature ids: element (n, s) at ids[n * ids_ld_n + s * ids_ld_s]
  long long ids_ld_n, ids_ld_s;
  long long pad_id;
  int N, S, d;
  int n_layers;
  int dims[kFtMaxLayers + 1];    // dims[0] = d, dims[n_layers] = num_labels
  const float* W[kFtMaxLayers];  // [dims[l+1], dims[l]]
  const float* b[kFtMaxLayers];  // [dims[l+1]] or NULL
  float* dW[kFtMaxLayers];
  float* db[kFtMaxLayers];
  int act;                       // 1: gelu (+ dropout) in front of EVERY Linear (the MLP head); 0: plain Linear
  DropParams drop;
  int mode;
  const long long* labels_i;     // CE: [N]
  const float* labels_f;         // MSE / L1 / BCE: [N, num_labels]
  const float* wgt;              // mode 1: [N]
  int* seq_idx;                  // [N] pooled row index (out fwd, in bwd)
  __nv_bfloat16* pooled;         // [N, d]
  float* logits;                 // [N, num_labels]
  float* pre;                    // stash [n_layers][N][kFtMaxDim-strided dims]: input of layer l BEFORE activation
  long long pre_stride_l;        // elements between layers in `pre` (= N * max_dim)
  int max_dim;
  float* row_loss;               // [N]
  float* row_den;                // [N]
  const float* fin;              // bwd: [loss, 1/denominator]
  const float* gout;             // bwd: upstream gradient of the loss (1 element)
  __nv_bfloat16* dhidden;        // bwd: [N*S, ldd], pre-zeroed
  long long ldd;
  int* err;
};

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_exact_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (kFtThreads >> 5); ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(kFtThreads) ft_head_fwd_kernel(const FtHeadParams p) {
  __shared__ float xa[kFtMaxDim], xb[kFtMaxDim];
  __shared__ float red[kFtThreads >> 5];
  __shared__ int s_row;
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- 1. last valid row: (#non-pad ids) - 1, negative wraps like a torch index
  float cnt = 0.f;
  for (int s = tid; s < p.S; s += kFtThreads) cnt += (p.ids[n * p.ids_ld_n + s * p.ids_ld_s] != p.pad_id) ? 1.f : 0.f;
  cnt = block_sum(cnt, red);
  if (tid == 0) {
    int sl = static_cast<int>(cnt) - 1;
    if (sl < 0) sl += p.S;
    s_row = sl;
    p.seq_idx[n] = sl;
  }
  __syncthreads();
  const long long row = static_cast<long long>(n) * p.S + s_row;
  // ---- 2. pooled hidden row
  for (int i = tid; i < p.d; i += kFtThreads) {
    const __nv_bfloat16 v = p.hidden[row * p.ld + i];
    p.pooled[static_cast<long long>(n) * p.d + i] = v;
    xa[i] = __bfloat162float(v);
  }
  __syncthreads();
  // ---- 3. score head
  float* x = xa;
  float* y = xb;
  for (int l = 0; l < p.n_layers; ++l) {
    const int in = p.dims[l], out = p.dims[l + 1];
    float* pre = p.pre + l * p.pre_stride_l + static_cast<long long>(n) * p.max_dim;
    for (int i = tid; i < in; i += kFtThreads) {
      float v = x[i];
      pre[i] = v;
      if (p.act) {
        v = gelu_exact(v);
        if (p.drop.thresh != 0u) v *= edrop_scale1(p.drop, (static_cast<unsigned long long>(l) * p.N + n) * p.max_dim + i);
      }
      x[i] = v;
    }
    __syncthreads();
    for (int j = warp; j < out; j += (kFtThreads >> 5)) {
      const float* w = p.W[l] + static_cast<long long>(j) * in;
      float acc = 0.f;
      for (int i = lane; i < in; i += 32) acc = fmaf(w[i], x[i], acc);
      acc = warp_sum(acc);
      if (lane == 0) y[j] = acc + (p.b[l] != nullptr ? p.b[l][j] : 0.f);
    }
    __syncthreads();
    float* t = x; x = y; y = t;
  }
  const int C = p.dims[p.n_layers];
  for (int j = tid; j < C; j += kFtThreads) p.logits[static_cast<long long>(n) * C + j] = x[j];
  if (p.mode < 0) return;
  // ---- 4. loss term and denominator term of this sample
  float lsum = 0.f, dsum = 0.f;
  if (p.mode <= 1) {
    float m = -INFINITY;
    for (int j = tid; j < C; j += kFtThreads) m = fmaxf(m, x[j]);
    m = warp_max(m);
    __syncthreads();
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < (kFtThreads >> 5); ++i) m = fmaxf(m, red[i]);
    float s = 0.f;
    for (int j = tid; j < C; j += kFtThreads) s += expf(x[j] - m);
    s = block_sum(s, red);
    if (tid == 0) {
      long long lab = p.labels_i[n];
      if (lab < 0 || lab >= C) {
        if (p.err) atomicExch(p.err, 2);
        lab = 0;
      }
      const float w = p.mode == 1 ? p.wgt[n] : 1.0f;
      lsum = (m + logf(s) - x[lab]) * w;
      dsum = w;
    }
  } else {
    float a = 0.f, c = 0.f;
    for (int j = tid; j < C; j += kFtThreads) {
      const float z = x[j], t = p.labels_f[static_cast<long long>(n) * C + j];
      if (p.mode == 2) { a += (z - t) * (z - t); c += 1.f; }
      else if (p.mode == 3) { a += fabsf(z - t); c += 1.f; }
      else if (t == t) { a += fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z))); c += 1.f; }   // BCE, labelled entries only
    }
    lsum = block_sum(a, red);
    dsum = block_sum(c, red);
  }
  if (tid == 0) {
    p.row_loss[n] = lsum;
    p.row_den[n] = dsum;
  }
}

// loss = sum(row_loss) / sum(row_den) in a fixed order;  out = [loss, 1 / denominator]
__global__ void ft_head_finalize_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_den, int N,
                                        float* __restrict__ out) {
  __shared__ double sl[32], sd[32];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    a += row_loss[i];
    b += row_den[i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sl[threadIdx.x >> 5] = a;
    sd[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = b = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) {
      a += sl[i];
      b += sd[i];
    }
    out[0] = static_cast<float>(a / b);          // 0 / 0 = NaN, as torch's mean over an empty selection
    out[1] = static_cast<float>(1.0 / b);
  }
}

__global__ void __launch_bounds__(kFtThreads) ft_head_bwd_kernel(const FtHeadParams p) {
  __shared__ float ga[kFtMaxDim], gb[kFtMaxDim], xin[kFtMaxDim];
  __shared__ float red[kFtThreads >> 5];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.dims[p.n_layers];
  const float scale = p.gout[0] * p.fin[1];
  const float* z = p.logits + static_cast<long long>(n) * C;
  // ---- dL/dlogits of this sample
  if (p.mode <= 1) {
    float m = -INFINITY;
    for (int j = tid; j < C; j += kFtThreads) m = fmaxf(m, z[j]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < (kFtThreads >> 5); ++i) m = fmaxf(m, red[i]);
    float s = 0.f;
    for (int j = tid; j < C; j += kFtThreads) s += expf(z[j] - m);
    s = block_sum(s, red);
    long long lab = p.labels_i[n];
    if (lab < 0 || lab >= C) lab = 0;
    const float w = p.mode == 1 ? p.wgt[n] : 1.0f;
    for (int j = tid; j < C; j += kFtThreads) ga[j] = (expf(z[j] - m) / s - (j == lab ? 1.f : 0.f)) * w * scale;
  } else {
    for (int j = tid; j < C; j += kFtThreads) {
      const float t = p.labels_f[static_cast<long long>(n) * C + j];
      float g;
      if (p.mode == 2) g = 2.f * (z[j] - t);
      else if (p.mode == 3) g = (z[j] > t) ? 1.f : ((z[j] < t) ? -1.f : 0.f);
      else g = (t == t) ? (1.f / (1.f + expf(-z[j])) - t) : 0.f;
      ga[j] = g * scale;
    }
  }
  __syncthreads();
  float* g = ga;
  float* gn = gb;
  for (int l = p.n_layers - 1; l >= 0; --l) {
    const int in = p.dims[l], out = p.dims[l + 1];
    const float* pre = p.pre + l * p.pre_stride_l + static_cast<long long>(n) * p.max_dim;
    for (int i = tid; i < in; i += kFtThreads) {
      float v = pre[i];
      if (p.act) {
        v = gelu_exact(v);
        if (p.drop.thresh != 0u) v *= edrop_scale1(p.drop, (static_cast<unsigned long long>(l) * p.N + n) * p.max_dim + i);
      }
      xin[i] = v;
    }
    __syncthreads();
    // dW[j, i] += g[j] * xin[i];  db[j] += g[j]
    if (p.dW[l] != nullptr) {
      for (long long e = tid; e < static_cast<long long>(out) * in; e += kFtThreads) {
        const int j = static_cast<int>(e / in), i = static_cast<int>(e % in);
        atomicAdd(p.dW[l] + e, g[j] * xin[i]);
      }
    }
    if (p.db[l] != nullptr)
      for (int j = tid; j < out; j += kFtThreads) atomicAdd(p.db[l] + j, g[j]);
    // gin[i] = sum_j W[j, i] g[j], then back through dropout and gelu
    for (int i = tid; i < in; i += kFtThreads) {
      float acc = 0.f;
      for (int j = 0; j < out; ++j) acc = fmaf(p.W[l][static_cast<long long>(j) * in + i], g[j], acc);
      if (p.act) {
        const float x0 = pre[i];
        float dact = gelu_exact_grad(x0);
        if (p.drop.thresh != 0u) dact *= edrop_scale1(p.drop, (static_cast<unsigned long long>(l) * p.N + n) * p.max_dim + i);
        acc *= dact;
      }
      gn[i] = acc;
    }
    __syncthreads();
    float* t = g; g = gn; gn = t;
  }
  if (p.dhidden != nullptr) {
    const long long row = static_cast<long long>(n) * p.S + p.seq_idx[n];
    for (int i = tid; i < p.d; i += kFtThreads) p.dhidden[row * p.ldd + i] = __float2bfloat16_rn(g[i]);
  }
}


// =============================================================================================
// loss_type "token_ce_intra" (modeling_finetune.py:137-165): the class "embeddings" of a sample are its OWN hidden states
// at positions cls_idx[n] .. cls_idx[n] + C - 1;  logits[n,s,c] = 20 * <a[n,s], a[n, cls_idx[n] + c]> with
// a = h / max(||h||, 1e-12) (F.normalize), CrossEntropy over the labelled positions (-100 ignored).
// Forward: one warp per position.  Backward in three deterministic steps (no atomics):
//   A  per position s : dl[s,:] = (softmax - onehot) * scale;  G[s,:] = 20 sum_c dl[s,c] a[e_c]          (query role)
//   B  per (n, c)     : G[e_c,:] += 20 sum_s dl[s,c] a[s]                                                (class-row role)
//   C  per position t : dh[t,:] = rn[t] * (G[t,:] - a[t] <a[t], G[t,:]>)                                 (normalize backward)
// =============================================================================================
__global__ void ft_intra_fwd_kernel(const __nv_bfloat16* __restrict__ h, long long ld, const long long* __restrict__ cls_idx,
                                    const long long* __restrict__ labels, float* __restrict__ rn, float* __restrict__ logits,
                                    float* __restrict__ row_loss, float* __restrict__ row_den, int N, int S, int C, int d,
                                    int phase, int* __restrict__ err) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long T = static_cast<long long>(N) * S;
  for (long long t = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * wpb) {
    const __nv_bfloat16* hs = h + t * ld;
    if (phase == 0) {                       // inverse norms
      float ss = 0.f;
      for (int i = lane; i < d; i += 32) {
        const float v = __bfloat162float(hs[i]);
        ss += v * v;
      }
      ss = warp_sum(ss);
      if (lane == 0) rn[t] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      continue;
    }
    const int n = static_cast<int>(t / S);
    long long e0 = cls_idx[n];
    if (e0 < 0 || e0 + C > S) {
      if (err && lane == 0) atomicExch(err, 2);
      e0 = 0;
    }
    const float rs = rn[t] * 20.0f;
    float m = -INFINITY, sum = 0.f, zlab = 0.f;
    long long lab = labels != nullptr ? labels[t] : -100;
    if (lab != -100 && (lab < 0 || lab >= C)) {
      if (err && lane == 0) atomicExch(err, 2);
      lab = 0;
    }
    for (int c = 0; c < C; ++c) {
      const long long te = static_cast<long long>(n) * S + e0 + c;
      const __nv_bfloat16* he = h + te * ld;
      float acc = 0.f;
      for (int i = lane; i < d; i += 32) acc = fmaf(__bfloat162float(hs[i]), __bfloat162float(he[i]), acc);
      const float z = warp_sum(acc) * rs * rn[te];
      if (lane == 0) logits[t * C + c] = z;
      const float mn = fmaxf(m, z);
      sum = sum * expf(m - mn) + expf(z - mn);
      m = mn;
      if (c == lab) zlab = z;
    }
    if (lane == 0 && row_loss != nullptr) {
      const bool on = lab != -100;
      row_loss[t] = on ? (m + logf(sum) - zlab) : 0.f;
      row_den[t] = on ? 1.f : 0.f;
    }
  }
}

__global__ void ft_intra_bwd_a_kernel(const __nv_bfloat16* __restrict__ h, long long ld, const long long* __restrict__ cls_idx,
                                      const long long* __restrict__ labels, const float* __restrict__ rn,
                                      const float* __restrict__ logits, const float* __restrict__ fin,
                                      const float* __restrict__ gout, float* __restrict__ dl, float* __restrict__ G, int N,
                                      int S, int C, int d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long T = static_cast<long long>(N) * S;
  const float scale = gout[0] * fin[1];
  for (long long t = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * wpb) {
    const int n = static_cast<int>(t / S);
    long long e0 = cls_idx[n];
    if (e0 < 0 || e0 + C > S) e0 = 0;
    long long lab = labels[t];
    const bool on = lab != -100;
    if (on && (lab < 0 || lab >= C)) lab = 0;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, logits[t * C + c]);
    m = warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += expf(logits[t * C + c] - m);
    sum = warp_sum(sum);
    for (int c = lane; c < C; c += 32)
      dl[t * C + c] = on ? (expf(logits[t * C + c] - m) / sum - (c == lab ? 1.f : 0.f)) * scale : 0.f;
    __syncwarp();
    for (int i = lane; i < d; i += 32) {
      float acc = 0.f;
      if (on) {
        for (int c = 0; c < C; ++c) {
          const long long te = static_cast<long long>(n) * S + e0 + c;
          acc = fmaf(dl[t * C + c] * rn[te], __bfloat162float(h[te * ld + i]), acc);
        }
      }
      G[t * d + i] = 20.0f * acc;
    }
  }
}

__global__ void ft_intra_bwd_b_kernel(const __nv_bfloat16* __restrict__ h, long long ld, const long long* __restrict__ cls_idx,
                                      const float* __restrict__ rn, const float* __restrict__ dl, float* __restrict__ G,
                                      int N, int S, int C, int d) {
  // one CTA per (n, c): thread i owns feature columns i, i + blockDim, ...
  const int n = blockIdx.x / C, c = blockIdx.x % C;
  long long e0 = cls_idx[n];
  if (e0 < 0 || e0 + C > S) e0 = 0;
  const long long te = static_cast<long long>(n) * S + e0 + c;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) {
      const long long t = static_cast<long long>(n) * S + s;
      const float w = dl[t * C + c];
      if (w != 0.f) acc = fmaf(w * rn[t], __bfloat162float(h[t * ld + i]), acc);
    }
    G[te * d + i] += 20.0f * acc;
  }
}

__global__ void ft_intra_bwd_c_kernel(const __nv_bfloat16* __restrict__ h, long long ld, const float* __restrict__ rn,
                                      const float* __restrict__ G, __nv_bfloat16* __restrict__ dh, long long ldd, long long T,
                                      int d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * wpb) {
    const float r = rn[t];
    float dot = 0.f;
    for (int i = lane; i < d; i += 32) dot = fmaf(__bfloat162float(h[t * ld + i]) * r, G[t * d + i], dot);
    dot = warp_sum(dot);
    for (int i = lane; i < d; i += 32) {
      const float a = __bfloat162float(h[t * ld + i]) * r;
      dh[t * ldd + i] = __float2bfloat16_rn(r * (G[t * d + i] - a * dot));
    }
  }
}

static inline unsigned intra_grid(long long rows) {
  long long b = (rows + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

static int fill_params(FtHeadParams& p, const void* hidden, long long ld, const long long* ids, long long ids_ld_n,
                       long long ids_ld_s, long long pad_id, int N, int S, int n_layers, const int* dims,
                       const float* const* W, const float* const* b, int act, float drop_p, unsigned long long seed, int mode,
                       const long long* labels_i, const float* labels_f, const float* wgt, int* seq_idx, float* logits,
                       float* pre, int max_dim) {
  GGPT_REQUIRE(hidden && ids && dims && W && seq_idx && logits && pre, "ft_head: null pointer");
  GGPT_REQUIRE(N > 0 && S > 0 && n_layers >= 1 && n_layers <= kFtMaxLayers, "ft_head: N=%d S=%d layers=%d (at most %d Linear layers)",
               N, S, n_layers, kFtMaxLayers);
  GGPT_REQUIRE(max_dim <= kFtMaxDim, "ft_head: layer width %d exceeds %d", max_dim, kFtMaxDim);
  GGPT_REQUIRE(mode >= -1 && mode <= 4, "ft_head: loss mode %d", mode);
  GGPT_REQUIRE(mode < 0 || (mode <= 1 ? labels_i != nullptr : labels_f != nullptr), "ft_head: labels missing for mode %d", mode);
  GGPT_REQUIRE(mode != 1 || wgt != nullptr, "ft_head: weighted CE needs sample weights");
  GGPT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "ft_head: dropout p=%f outside [0,1)", drop_p);
  p.hidden = static_cast<const __nv_bfloat16*>(hidden);
  p.ld = ld; p.ids = ids; p.ids_ld_n = ids_ld_n; p.ids_ld_s = ids_ld_s; p.pad_id = pad_id;
  p.N = N; p.S = S; p.d = dims[0]; p.n_layers = n_layers;
  for (int l = 0; l <= n_layers; ++l) {
    GGPT_REQUIRE(dims[l] > 0 && dims[l] <= max_dim, "ft_head: bad layer width %d", dims[l]);
    p.dims[l] = dims[l];
  }
  for (int l = 0; l < n_layers; ++l) {
    GGPT_REQUIRE(W[l] != nullptr, "ft_head: null weight of layer %d", l);
    p.W[l] = W[l];
    p.b[l] = b ? b[l] : nullptr;
    p.dW[l] = nullptr;
    p.db[l] = nullptr;
  }
  p.act = act; p.drop = make_drop_params(drop_p, seed); p.mode = mode;
  p.labels_i = labels_i; p.labels_f = labels_f; p.wgt = wgt;
  p.seq_idx = seq_idx; p.logits = logits; p.pre = pre; p.max_dim = max_dim;
  p.pre_stride_l = static_cast<long long>(N) * max_dim;
  return 0;
}

}  // namespace ggpt

using namespace ggpt;

extern "C" {

int ggpt_ft_head_fwd(const void* hidden, long long ld, const long long* ids, long long ids_ld_n, long long ids_ld_s,
                     long long pad_id, int N, int S, int n_layers, const int* dims, const float* const* W,
                     const float* const* b, int act, float drop_p, unsigned long long drop_seed, int mode,
                     const long long* labels_i, const float* labels_f, const float* sample_wgt, int* seq_idx, void* pooled,
                     float* logits, float* pre, int max_dim, float* row_loss, float* row_den, float* loss_out, int* err_flag,
                     void* stream) {
  FtHeadParams p{};
  if (int rc = fill_params(p, hidden, ld, ids, ids_ld_n, ids_ld_s, pad_id, N, S, n_layers, dims, W, b, act, drop_p, drop_seed,
                           mode, labels_i, labels_f, sample_wgt, seq_idx, logits, pre, max_dim))
    return rc;
  GGPT_REQUIRE(pooled != nullptr, "ft_head_fwd: null pooled output");
  GGPT_REQUIRE(mode < 0 || (row_loss && row_den && loss_out), "ft_head_fwd: loss outputs missing");
  p.pooled = static_cast<__nv_bfloat16*>(pooled);
  p.row_loss = row_loss; p.row_den = row_den; p.err = err_flag;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ft_head_fwd_kernel<<<N, kFtThreads, 0, s>>>(p);
  if (int rc = check_launch("ft_head_fwd_kernel")) return rc;
  if (mode >= 0) {
    ft_head_finalize_kernel<<<1, 256, 0, s>>>(row_loss, row_den, N, loss_out);
    return check_launch("ft_head_finalize_kernel");
  }
  return 0;
}

int ggpt_ft_head_bwd(int N, int S, int n_layers, const int* dims, const float* const* W,
                     float* const* dW, float* const* db, int act, float drop_p, unsigned long long drop_seed, int mode,
                     const long long* labels_i, const float* labels_f, const float* sample_wgt, const int* seq_idx,
                     const float* logits, const float* pre, int max_dim, const float* loss_out, const float* gout,
                     void* dhidden, long long ldd, void* stream) {
  GGPT_REQUIRE(dims && W && seq_idx && logits && pre && loss_out && gout, "ft_head_bwd: null pointer");
  GGPT_REQUIRE(N > 0 && S > 0 && n_layers >= 1 && n_layers <= kFtMaxLayers && max_dim <= kFtMaxDim, "ft_head_bwd: bad sizes");
  GGPT_REQUIRE(mode >= 0 && mode <= 4, "ft_head_bwd: loss mode %d", mode);
  FtHeadParams p{};
  p.N = N; p.S = S; p.d = dims[0]; p.n_layers = n_layers;
  for (int l = 0; l <= n_layers; ++l) p.dims[l] = dims[l];
  for (int l = 0; l < n_layers; ++l) {
    p.W[l] = W[l];
    p.dW[l] = dW ? dW[l] : nullptr;
    p.db[l] = db ? db[l] : nullptr;
  }
  p.act = act; p.drop = make_drop_params(drop_p, drop_seed); p.mode = mode;
  p.labels_i = labels_i; p.labels_f = labels_f; p.wgt = sample_wgt;
  p.seq_idx = const_cast<int*>(seq_idx); p.logits = const_cast<float*>(logits); p.pre = const_cast<float*>(pre);
  p.max_dim = max_dim; p.pre_stride_l = static_cast<long long>(N) * max_dim;
  p.fin = loss_out; p.gout = gout;
  p.dhidden = static_cast<__nv_bfloat16*>(dhidden); p.ldd = ldd;
  ft_head_bwd_kernel<<<N, kFtThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("ft_head_bwd_kernel");
}

int ggpt_ft_intra_fwd(const void* hidden, long long ld, const long long* cls_idx, const long long* labels, float* rn,
                      float* logits, float* row_loss, float* row_den, float* loss_out, int N, int S, int C, int d,
                      int* err_flag, void* stream) {
  GGPT_REQUIRE(hidden && cls_idx && rn && logits, "ft_intra_fwd: null pointer");
  GGPT_REQUIRE(N > 0 && S > 0 && C > 0 && C <= S && d > 0, "ft_intra_fwd: bad sizes N=%d S=%d C=%d d=%d", N, S, C, d);
  GGPT_REQUIRE(labels == nullptr || (row_loss && row_den && loss_out), "ft_intra_fwd: loss outputs missing");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long T = static_cast<long long>(N) * S;
  const __nv_bfloat16* h = static_cast<const __nv_bfloat16*>(hidden);
  ft_intra_fwd_kernel<<<intra_grid(T), 256, 0, s>>>(h, ld, cls_idx, labels, rn, logits, nullptr, nullptr, N, S, C, d, 0, err_flag);
  if (int rc = check_launch("ft_intra_fwd_kernel(norms)")) return rc;
  ft_intra_fwd_kernel<<<intra_grid(T), 256, 0, s>>>(h, ld, cls_idx, labels, rn, logits, labels ? row_loss : nullptr, row_den, N,
                                                    S, C, d, 1, err_flag);
  if (int rc = check_launch("ft_intra_fwd_kernel")) return rc;
  if (labels != nullptr) {
    ft_head_finalize_kernel<<<1, 256, 0, s>>>(row_loss, row_den, static_cast<int>(T), loss_out);
    return check_launch("ft_head_finalize_kernel");
  }
  return 0;
}

int ggpt_ft_intra_bwd(const void* hidden, long long ld, const long long* cls_idx, const long long* labels, const float* rn,
                      const float* logits, const float* loss_out, const float* gout, float* dl_scratch, float* g_scratch,
                      void* dhidden, long long ldd, int N, int S, int C, int d, void* stream) {
  GGPT_REQUIRE(hidden && cls_idx && labels && rn && logits && loss_out && gout && dl_scratch && g_scratch && dhidden,
               "ft_intra_bwd: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long T = static_cast<long long>(N) * S;
  const __nv_bfloat16* h = static_cast<const __nv_bfloat16*>(hidden);
  ft_intra_bwd_a_kernel<<<intra_grid(T), 256, 0, s>>>(h, ld, cls_idx, labels, rn, logits, loss_out, gout, dl_scratch, g_scratch,
                                                      N, S, C, d);
  if (int rc = check_launch("ft_intra_bwd_a_kernel")) return rc;
  ft_intra_bwd_b_kernel<<<N * C, 256, 0, s>>>(h, ld, cls_idx, rn, dl_scratch, g_scratch, N, S, C, d);
  if (int rc = check_launch("ft_intra_bwd_b_kernel")) return rc;
  ft_intra_bwd_c_kernel<<<intra_grid(T), 256, 0, s>>>(h, ld, rn, g_scratch, static_cast<__nv_bfloat16*>(dhidden), ldd, T, d);
  return check_launch("ft_intra_bwd_c_kernel");
}

}  // extern "C"
