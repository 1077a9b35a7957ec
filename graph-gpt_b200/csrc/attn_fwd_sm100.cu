// Flash-style softmax attention forward on tcgen05 (sm_100a) for GraphGPT's padded / packed / causal sequences.
//
//   O[n, q, h, :] = softmax_k( Q[n,q,h,:] . K[n,k,h,:] / sqrt(64) + mask[n,q,k] ) V[n,k,h,:]
//
// The reference materialises an additive [N,1,S,S] mask (0 / finfo.min) and runs matmul-softmax-matmul
// (HF:199-221 eager_attention_forward; mask from modeling_helpers.py:38-64).  Here the mask is a bit matrix
// (1 bit per (q,k), built once per step by ggpt_attn_mask_build) plus a per-tile-pair class
// (0 = all masked -> tile skipped, 1 = all visible -> no mask loads, 2 = mixed), which covers right-padding,
// block-diagonal packing, causal and any other 2-D/3-D mask the reference accepts without an O(S^2) fp tensor.
//
// Row tiles are VARIABLE: each sequence is cut into tiles of <= 128 rows (shared by queries and keys) and a cut is
// placed on a block boundary of the mask (no visible pair straddles it) whenever one exists 64..128 rows after the
// previous cut.  Packed Eulerian-path segments (~24 rows) therefore never straddle a tile: every query tile sees
// exactly one key tile, instead of the ~3 an aligned 128-grid gives.  Dense / padded / causal masks have no interior
// boundaries and fall back to the uniform grid.
//
// One CTA per (128-query tile, head, sequence); 6 warps: TMA producer, MMA issuer, 4 softmax warps (one query row
// per thread).  QK^T and PV run on tcgen05 with accumulators in TMEM; P goes through shared memory (bf16, 128B
// swizzle) as the A operand of the PV MMA; V is consumed in place as an MN-major B operand.  Two CTAs fit per SM
// (112 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
#include "common.cuh"
#include "attn_diag.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

constexpr int kAttThreads = 192;
constexpr int kTileQ = 128;
constexpr int kTileK = 128;
constexpr int kHeadDim = 64;
constexpr int kAttSmem = 16384 /*Q*/ + 2 * 16384 /*K*/ + 2 * 16384 /*V*/ + 32768 /*P*/ + 256 /*barriers*/;

struct AttnFwdParams {
  int N, S, H;
  int max_tiles;              // allocation stride of the tile arrays (= ceil(S/64))
  int mask_words;             // uint32 words per mask row (multiple of 4, padded by 4)
  const uint32_t* mask_bits;  // [N, S, mask_words]
  const int* tile_start;      // [N, max_tiles+1]  row offsets of the variable row tiles
  const int* n_tiles;         // [N]
  const uint8_t* tile_cls;    // [N, max_tiles, max_tiles]  (query tile, key tile)
  const uint8_t* iso_flags;   // [N, max_tiles] tiles handled by the persistent diagonal kernel (may be NULL)
  __nv_bfloat16* out;         // [N*S, H*64]
  long long ldo;
  float* lse;                 // [N, H, S]  natural-log logsumexp of the scaled scores (for backward)
  int q_col0, k_col0, v_col0; // column offsets of q/k/v inside the fused qkv row
  float scale_log2;           // (1/sqrt(64)) * log2(e)
  DropParams drop;            // attention-probability dropout (training)
};

// 128 mask bits of query row `mrow` for the key tile starting at key k0 (any alignment) with klen valid keys.
__device__ __forceinline__ void load_mask_words(const uint32_t* __restrict__ mrow, int k0, int klen, uint32_t (&mw)[4]) {
  const int w0 = k0 >> 5, sh = k0 & 31;
  uint32_t w[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) w[i] = mrow[w0 + i];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t v = __funnelshift_r(w[i], w[i + 1], sh);
    const int nvalid = klen - 32 * i;
    const uint32_t keep = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
    mw[i] = v & keep;
  }
}

// Online-softmax update of one thread's query row for one 128-key tile: reads the scores twice from TMEM (row max, then
// probabilities), writes bf16 P into the swizzled K-major A-operand buffer, returns the rescale factor of the running
// output.  MASKED = the tile pair is "mixed" (class 2) and every element consults its mask bit; full tiles (class 1) skip
// all mask work.  DROP = attention dropout is on (one 32-bit draw per key pair, see common.cuh).
template <bool MASKED, bool DROP>
__device__ __forceinline__ float fwd_softmax_tile(uint32_t tS, uint8_t* prow, int r, const uint32_t (&mw)[4],
                                                  float scale_log2, float& m_run, float& l_run, uint32_t rk1,
                                                  uint32_t rk2, int k0, uint32_t thresh) {
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains instead of a 128-deep max
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t s[32];
    tmem_ld32(tS + c * 32, s);
    tmem_ld_wait();
    const uint32_t w = (c == 0) ? mw[0] : (c == 1) ? mw[1] : (c == 2) ? mw[2] : mw[3];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float v = __uint_as_float(s[j]);
      if (MASKED) v = ((w >> j) & 1u) ? v : -INFINITY;
      m4[j & 3] = fmaxf(m4[j & 3], v);
    }
  }
  const float m_tile = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;   // scale > 0 commutes with max
  const float m_new = fmaxf(m_run, m_tile);
  const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
  const float alpha = (m_run == -INFINITY) ? 0.f : fast_exp2(m_run - m_use);
  float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t s[32];
    tmem_ld32(tS + c * 32, s);
    tmem_ld_wait();
    const uint32_t w = (c == 0) ? mw[0] : (c == 1) ? mw[1] : (c == 2) ? mw[2] : mw[3];
    // columns c*32 .. c*32+31 of P: half = c/2 (64-col K block), 16-byte chunks (c&1)*4 .. +3
    uint8_t* pbase = prow + (c >> 1) * 16384;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float pv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float e = fast_exp2(fmaf(__uint_as_float(s[g * 8 + j]), scale_log2, -m_use));
        if (MASKED) e = ((w >> (g * 8 + j)) & 1u) ? e : 0.f;
        pv[j] = e;
        l4[j & 3] += e;
      }
      if (DROP) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const uint32_t bits = drop_bits(rk1, rk2, k0 + c * 32 + g * 8 + j);
          if (!drop_keep_even(bits, thresh)) pv[j] = 0.f;
          if (!drop_keep_odd(bits, thresh)) pv[j + 1] = 0.f;
        }
      }
      uint4 o;
      o.x = pack_bf16(pv[0], pv[1]);
      o.y = pack_bf16(pv[2], pv[3]);
      o.z = pack_bf16(pv[4], pv[5]);
      o.w = pack_bf16(pv[6], pv[7]);
      const int chunk = ((c & 1) * 4 + g) ^ (r & 7);
      *reinterpret_cast<uint4*>(pbase + chunk * 16) = o;
    }
  }
  l_run = l_run * alpha + ((l4[0] + l4[1]) + (l4[2] + l4[3]));
  m_run = m_new;
  return alpha;
}

template <bool DROP>
__global__ void __launch_bounds__(kAttThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sV = smem + 16384 + 32768;
  uint8_t* sP = smem + 16384 + 65536;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 65536 + 32768);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int n_kt = p.n_tiles[n];
  if (qt >= n_kt) return;                       // uniform for the whole CTA, before any barrier / TMEM state
  if (p.iso_flags != nullptr && p.iso_flags[n * p.max_tiles + qt]) return;   // done by attn_diag_fwd_kernel
  const int* ts = p.tile_start + static_cast<size_t>(n) * (p.max_tiles + 1);
  const int q0 = ts[qt], qlen = ts[qt + 1] - q0;
  const uint8_t* cls_row = p.tile_cls + (static_cast<size_t>(n) * p.max_tiles + qt) * p.max_tiles;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("ggpt attn_fwd: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;         // 128 columns
  const uint32_t tmem_O = tmem_base + 128;   // 64 columns

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, 16384);
      tma_load_3d(sQ, &tmQKV, q_full, p.q_col0 + h * kHeadDim, q0, n);
      int it = 0;
      for (int kt = 0; kt < n_kt; ++kt) {
        if (cls_row[kt] == 0) continue;
        const int st = it & 1;
        mbar_wait(&kv_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 32768);
        tma_load_3d(sK + st * 16384, &tmQKV, &kv_full[st], p.k_col0 + h * kHeadDim, ts[kt], n);
        tma_load_3d(sV + st * 16384, &tmQKV, &kv_full[st], p.v_col0 + h * kHeadDim, ts[kt], n);
        ++it;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      auto issue_S = [&](int st) {
        const uint32_t aK = smem_u32(sK + st * 16384);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_bf16(tmem_S, umma_desc_sw128(aQ + kk * 32, 16, 1024), umma_desc_sw128(aK + kk * 32, 16, 1024),
                      idesc_s, kk != 0);
        tc_commit(s_full);
      };
      int n_active = 0;
      for (int kt = 0; kt < n_kt; ++kt) n_active += (cls_row[kt] != 0);
      if (n_active > 0) {
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        issue_S(0);
        for (int it = 0; it < n_active; ++it) {
          const int st = it & 1;
          mbar_wait(p_full, it & 1);   // P(it) in smem, S(it) and O(it-1) drained from TMEM
          tc_fence_after();
          const uint32_t aV = smem_u32(sV + st * 16384);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            tc_mma_bf16(tmem_O, umma_desc_sw128(aP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                        umma_desc_sw128(aV + kk * 2048, 8192, 1024), idesc_o, kk != 0);
          tc_commit(o_full);
          tc_commit(&kv_empty[st]);
          if (it + 1 < n_active) {
            const int st2 = (it + 1) & 1;
            mbar_wait(&kv_full[st2], ((it + 1) >> 1) & 1);
            tc_fence_after();
            issue_S(st2);
          }
        }
      }
    }
  } else {
    // ===================== softmax / epilogue warps: one query row per thread =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;              // row inside the tile
    const int q_row = q0 + r;
    const bool row_ok = r < qlen;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t* mrow = p.mask_bits + (static_cast<size_t>(n) * p.S + (row_ok ? q_row : 0)) * p.mask_words;

    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    const uint32_t rowkey = drop_rowkey(p.drop.seed_lo, p.drop.seed_hi, n, h, q_row);
    const uint32_t rowkey2 = drop_rowkey2(rowkey);
    float m_run = -INFINITY;   // running max of scaled (log2-domain) scores
    float l_run = 0.f;

    int it = 0;
    for (int kt = 0; kt < n_kt; ++kt) {
      const int cls = cls_row[kt];
      if (cls == 0) continue;
      uint32_t mw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (cls == 2) {
        if (row_ok) {
          load_mask_words(mrow, ts[kt], ts[kt + 1] - ts[kt], mw);
        } else {
          mw[0] = mw[1] = mw[2] = mw[3] = 0u;
        }
      }
      const int k0 = ts[kt];
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      const float alpha = (cls == 2)
          ? fwd_softmax_tile<true, DROP>(tmem_S + lane_addr, sP + r * 128, r, mw, p.scale_log2, m_run, l_run, rowkey, rowkey2,
                                         k0, p.drop.thresh)
          : fwd_softmax_tile<false, DROP>(tmem_S + lane_addr, sP + r * 128, r, mw, p.scale_log2, m_run, l_run, rowkey,
                                          rowkey2, k0, p.drop.thresh);
      fence_proxy_async_smem();   // make P visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // accumulate O = O * alpha + P V
      mbar_wait(o_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(tmem_O + lane_addr + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o_acc[c * 32 + j] = fmaf(o_acc[c * 32 + j], alpha, __uint_as_float(o[j]));
      }
      ++it;
    }

    {
      // every PV has retired (o_full of the last tile was waited for), so this warp's own 32 rows of the P buffer stage the
      // output for full-line stores (stage_flush_rows) instead of 32 scattered 16-byte pieces per instruction
      const float inv = (l_run > 0.f) ? p.drop.inv_keep / l_run : 0.f;   // dropout keeps are rescaled by 1/(1-p)
      uint8_t* stg = sP + quad * 4096;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 o;
        o.x = pack_bf16(o_acc[g * 8 + 0] * inv, o_acc[g * 8 + 1] * inv);
        o.y = pack_bf16(o_acc[g * 8 + 2] * inv, o_acc[g * 8 + 3] * inv);
        o.z = pack_bf16(o_acc[g * 8 + 4] * inv, o_acc[g * 8 + 5] * inv);
        o.w = pack_bf16(o_acc[g * 8 + 6] * inv, o_acc[g * 8 + 7] * inv);
        stage_put16(stg, lane, g, o);
      }
      stage_flush_rows(stg, lane, p.out + (static_cast<long long>(n) * p.S + q0 + quad * 32) * p.ldo + h * kHeadDim, p.ldo,
                       qlen - quad * 32);
    }
    if (row_ok) {
      if (p.lse != nullptr) {
        const float lse = (l_run > 0.f) ? (m_run * 0.6931471805599453f + logf(l_run)) : 0.f;
        p.lse[(static_cast<size_t>(n) * p.H + h) * p.S + q_row] = lse;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------------------------------------
// Mask preparation: int64 attention_mask ([N,S] key mask or [N,S,S]) (+ causal) -> bit matrix + tile classes
// ref: modeling_helpers.py:38-64 (_update_causal_mask / _expand_mask_from_3d_mask), HF:398-405 (causal)
// ---------------------------------------------------------------------------------------------
// kind[n] = 1 when the [N,S] mask of sequence n carries segment ids (some value >= 2), else 0 (plain 0/1 key mask)
__global__ void attn_mask_kind_kernel(const long long* __restrict__ am, int S, int* __restrict__ kind) {
  __shared__ int any;
  if (threadIdx.x == 0) any = 0;
  __syncthreads();
  const long long* row = am + static_cast<long long>(blockIdx.x) * S;
  bool mine = false;
  for (int k = threadIdx.x; k < S; k += blockDim.x) mine |= (row[k] >= 2);
  if (mine) atomicOr(&any, 1);
  __syncthreads();
  if (threadIdx.x == 0) kind[blockIdx.x] = any;
}

__global__ void attn_mask_bits_kernel(const long long* __restrict__ am, int am_dims, int N, int S, int causal,
                                      int mask_words, uint32_t* __restrict__ bits, const int* __restrict__ kind) {
  // one warp per (n, q, group of 4 words = 128 keys): four independent 8-byte loads per lane are in flight at once
  // (the [N,S,S] int64 mask of a packed batch is 537 MB per step — this pass is pure HBM streaming)
  const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int groups = mask_words >> 2;   // mask_words is a multiple of 4
  const long long total = static_cast<long long>(N) * S * groups;
  if (gw >= total) return;
  const int w0 = static_cast<int>(gw % groups) * 4;
  const long long nq = gw / groups;
  const int q = static_cast<int>(nq % S);
  const int n = static_cast<int>(nq / S);
  long long v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = (w0 + i) * 32 + lane;
    v[i] = 1;
    if (k < S && am != nullptr)
      v[i] = (am_dims == 2) ? am[static_cast<long long>(n) * S + k] : am[(static_cast<long long>(n) * S + q) * S + k];
  }
  // [N,S] masks: 0/1 = the reference's key-padding mask (every query, pad rows included, sees the non-zero keys).
  // A sequence whose mask holds values >= 2 carries SEGMENT IDS of a packed sequence (ggpt_pack_sequences): a query
  // sees exactly the keys of its own segment and pad rows (id 0) see nothing, i.e. the block-diagonal [N,S,S] mask of
  // tokenizer_utils.py:351-355 without materialising it.
  bool q_all = true;          // this query row sees every non-zero key
  long long aq = 0;
  if (am_dims == 2 && am != nullptr && kind != nullptr && kind[n] != 0) {
    aq = am[static_cast<long long>(n) * S + q];
    q_all = false;
  }
  uint32_t word[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = (w0 + i) * 32 + lane;
    const bool keep = (k < S) && (v[i] != 0) && (q_all || aq == v[i]) && !(causal && k > q);
    word[i] = __ballot_sync(0xffffffffu, keep);
  }
  if (lane == 0) *reinterpret_cast<uint4*>(bits + nq * mask_words + w0) = make_uint4(word[0], word[1], word[2], word[3]);
}

// Row-tile plan for one sequence (one CTA per sequence).  lo/hi = first/last visible key of each query row;
// r is a block boundary iff no visible pair straddles it: max_{q<r} hi[q] < r and min_{q>=r} lo[q] >= r.
__global__ void attn_plan_kernel(const uint32_t* __restrict__ bits, int S, int mask_words, int max_tiles,
                                 int* __restrict__ tile_start, int* __restrict__ n_tiles) {
  extern __shared__ int s_plan[];   // lo[S], hi[S]
  int* lo = s_plan;
  int* hi = s_plan + S;
  const int n = blockIdx.x;
  const int nw = (S + 31) / 32;
  for (int q = threadIdx.x; q < S; q += blockDim.x) {
    const uint32_t* row = bits + (static_cast<size_t>(n) * S + q) * mask_words;
    int l = 0x7fffffff, h = -1;
    for (int w = 0; w < nw; ++w) {
      const uint32_t v = row[w];
      if (v) {
        if (l == 0x7fffffff) l = w * 32 + (__ffs(v) - 1);
        h = w * 32 + (31 - __clz(v));
      }
    }
    lo[q] = l;
    hi[q] = h;
  }
  __syncthreads();
  if (threadIdx.x == 0) {          // inclusive prefix max of hi
    int m = -1;
    for (int q = 0; q < S; ++q) {
      m = max(m, hi[q]);
      hi[q] = m;
    }
  } else if (threadIdx.x == 32) {  // inclusive suffix min of lo
    int m = 0x7fffffff;
    for (int q = S - 1; q >= 0; --q) {
      m = min(m, lo[q]);
      lo[q] = m;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int* ts = tile_start + static_cast<size_t>(n) * (max_tiles + 1);
    int start = 0, t = 0;
    while (start < S) {
      int end = min(start + 128, S);
      if (end < S) {
        for (int b = end; b >= start + 64; --b) {
          if (hi[b - 1] < b && lo[b] >= b) {
            end = b;
            break;
          }
        }
      }
      ts[t++] = start;
      start = end;
    }
    ts[t] = S;
    n_tiles[n] = t;
  }
}

// class of every (query tile, key tile) pair: one warp per pair
__global__ void attn_tile_cls_kernel(const uint32_t* __restrict__ bits, int N, int S, int mask_words, int max_tiles,
                                     const int* __restrict__ tile_start, const int* __restrict__ n_tiles,
                                     uint8_t* __restrict__ cls) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= N * max_tiles * max_tiles) return;
  const int kt = gw % max_tiles;
  const int qt = (gw / max_tiles) % max_tiles;
  const int n = gw / (max_tiles * max_tiles);
  const int nt = n_tiles[n];
  if (qt >= nt || kt >= nt) {
    if (lane == 0) cls[gw] = 0;
    return;
  }
  const int* ts = tile_start + static_cast<size_t>(n) * (max_tiles + 1);
  const int q0 = ts[qt], qlen = ts[qt + 1] - q0, k0 = ts[kt], klen = ts[kt + 1] - k0;
  uint32_t any = 0, all = 0xffffffffu;
  for (int r = lane; r < qlen; r += 32) {
    uint32_t mw[4];
    load_mask_words(bits + (static_cast<size_t>(n) * S + q0 + r) * mask_words, k0, klen, mw);
    any |= mw[0] | mw[1] | mw[2] | mw[3];
    all &= mw[0] & mw[1] & mw[2] & mw[3];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    any |= __shfl_xor_sync(0xffffffffu, any, o);
    all &= __shfl_xor_sync(0xffffffffu, all, o);
  }
  const bool full = (qlen == 128) && (klen == 128) && (all == 0xffffffffu);
  if (lane == 0) cls[gw] = (any == 0) ? 0 : (full ? 1 : 2);
}

}  // namespace ggpt

using namespace ggpt;

extern "C" {

int ggpt_attn_mask_words(int S) { return ((S + 127) / 128) * 4 + 4; }

int ggpt_attn_max_tiles(int S) { return (S + 63) / 64; }

int ggpt_attn_mask_build(const long long* attention_mask, int mask_dims, int N, int S, int causal, uint32_t* mask_bits,
                         int* tile_start, int* n_tiles, uint8_t* tile_cls, uint8_t* iso_flags, int* iso_list,
                         int* iso_count, void* stream) {
  GGPT_REQUIRE(N > 0 && S > 0, "attn_mask_build: empty batch");
  GGPT_REQUIRE(attention_mask == nullptr || mask_dims == 2 || mask_dims == 3,
               "attention_mask of %d dims is not implemented (expected [N,S] or [N,S,S])", mask_dims);
  GGPT_REQUIRE(mask_bits && tile_start && n_tiles && tile_cls && iso_flags && iso_list && iso_count,
               "attn_mask_build: null output");
  GGPT_REQUIRE(S <= 16384, "attn_mask_build: S=%d exceeds the plan kernel's shared-memory budget", S);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int words = ggpt_attn_mask_words(S);
  const int mt = ggpt_attn_max_tiles(S);
  const long long warps = static_cast<long long>(N) * S * (words / 4);
  const long long blocks = (warps * 32 + 255) / 256;
  const int* kind = nullptr;
  if (mask_dims == 2 && attention_mask != nullptr) {   // n_tiles doubles as the per-sequence scratch until the plan kernel fills it
    attn_mask_kind_kernel<<<N, 256, 0, s>>>(attention_mask, S, n_tiles);
    if (int rc = check_launch("attn_mask_kind_kernel")) return rc;
    kind = n_tiles;
  }
  attn_mask_bits_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(attention_mask, mask_dims, N, S, causal, words,
                                                                       mask_bits, kind);
  if (int rc = check_launch("attn_mask_bits_kernel")) return rc;
  const size_t plan_smem = 2 * static_cast<size_t>(S) * sizeof(int);
  static bool attr_set = false;
  if (!attr_set && plan_smem > 48 * 1024) {
    cudaFuncSetAttribute(attn_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    attr_set = true;
  }
  attn_plan_kernel<<<N, 256, plan_smem, s>>>(mask_bits, S, words, mt, tile_start, n_tiles);
  if (int rc = check_launch("attn_plan_kernel")) return rc;
  const long long warps2 = static_cast<long long>(N) * mt * mt;
  attn_tile_cls_kernel<<<static_cast<unsigned>((warps2 * 32 + 255) / 256), 256, 0, s>>>(mask_bits, N, S, words, mt,
                                                                                         tile_start, n_tiles, tile_cls);
  if (int rc = check_launch("attn_tile_cls_kernel")) return rc;
  GGPT_REQUIRE((reinterpret_cast<uintptr_t>(iso_list) & 15) == 0, "attn_mask_build: iso_list must be 16-byte aligned");
  return attn_iso_build(tile_cls, n_tiles, tile_start, N, mt, iso_flags, iso_list, iso_count, s);
}

int ggpt_attn_fwd(const void* qkv, long long ld_qkv, int q_col0, int k_col0, int v_col0, const uint32_t* mask_bits,
                  const int* tile_start, const int* n_tiles, const uint8_t* tile_cls, const uint8_t* iso_flags,
                  const int* iso_list, const int* iso_count, int run_general, float dropout_p, unsigned long long seed,
                  void* out, long long ldo, float* lse, int N, int S, int H, void* stream) {
  GGPT_REQUIRE(qkv && mask_bits && tile_start && n_tiles && tile_cls && out, "attn_fwd: null pointer");
  GGPT_REQUIRE(N > 0 && S > 0 && H > 0, "attn_fwd: empty problem");
  GGPT_REQUIRE(ld_qkv % 8 == 0 && ldo % 8 == 0 && q_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0,
               "attn_fwd: leading dimensions / column offsets must be multiples of 8");
  CUtensorMap tm;
  if (int rc = make_tmap_3d_bf16(&tm, qkv, N, S, ld_qkv, static_cast<uint64_t>(S) * ld_qkv, ld_qkv, 128, 64)) return rc;
  AttnFwdParams p{};
  p.N = N; p.S = S; p.H = H;
  p.max_tiles = ggpt_attn_max_tiles(S);
  p.mask_words = ggpt_attn_mask_words(S);
  p.mask_bits = mask_bits; p.tile_start = tile_start; p.n_tiles = n_tiles; p.tile_cls = tile_cls;
  p.iso_flags = iso_flags;
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo; p.lse = lse;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  GGPT_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "attn_fwd: dropout_p must be in [0,1)");
  p.drop = make_drop_params(dropout_p, seed);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem);
    if (e != cudaSuccess) {
      set_error("attn_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return -2;
    }
    attr_set = true;
  }
  GGPT_REQUIRE(run_general || iso_flags, "attn_fwd: run_general == 0 needs the isolated-tile work list");
  if (run_general) {
    dim3 grid(p.max_tiles, H, N);
    if (p.drop.thresh != 0u) attn_fwd_kernel<true><<<grid, kAttThreads, kAttSmem, static_cast<cudaStream_t>(stream)>>>(tm, p);
    else attn_fwd_kernel<false><<<grid, kAttThreads, kAttSmem, static_cast<cudaStream_t>(stream)>>>(tm, p);
    if (int rc = check_launch("attn_fwd_kernel")) return rc;
  }
  if (iso_flags == nullptr) return 0;
  GGPT_REQUIRE(iso_list && iso_count, "attn_fwd: iso_flags given without iso_list / iso_count");
  DiagParams d{};
  d.N = N; d.S = S; d.H = H; d.max_tiles = p.max_tiles; d.mask_words = p.mask_words;
  d.mask_bits = mask_bits; d.tile_start = tile_start; d.tile_cls = tile_cls; d.iso_list = iso_list; d.iso_count = iso_count;
  d.q_col0 = q_col0; d.k_col0 = k_col0; d.v_col0 = v_col0; d.scale = 0.125f; d.scale_log2 = p.scale_log2;
  d.out = p.out; d.ldo = ldo; d.lse = lse; d.drop = p.drop;
  return attn_diag_fwd_launch(tm, d, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
