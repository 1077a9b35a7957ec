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
// One CTA per (128-query tile, group of heads, sequence); 10 warps: TMA producer, MMA issuer, 8 softmax warps (two
// threads per query row, 64 keys each — see the pipeline comment above attn_fwd_kernel).  QK^T and PV run on tcgen05
// with accumulators in TMEM; P goes through shared memory (bf16, 128B swizzle) as the A operand of the PV MMA; V is
// consumed in place as an MN-major B operand.  Two CTAs fit per SM (113 KB smem, 256 TMEM columns each) so one CTA's
// softmax overlaps the other's MMAs.
#include <stdlib.h>

#include "common.cuh"
#include "attn_diag.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

constexpr int kAttSoftmaxWarps = 8;
constexpr int kAttThreads = 64 + 32 * kAttSoftmaxWarps;
constexpr int kTileQ = 128;
constexpr int kTileK = 128;
constexpr int kHeadDim = 64;
// 2 CTAs per SM: 2 * (kAttSmem + 1 KB reserved) must stay within the SM's 228 KB — the 1 KB tail (barriers + tile plan)
// is all that is left
constexpr int kAttSmem = 16384 /*Q*/ + 2 * 16384 /*K*/ + 2 * 16384 /*V*/ + 32768 /*P*/ + 1024 /*barriers, tile plan*/;
static_assert(2 * (kAttSmem + 1024) <= 233472, "two attention-forward CTAs must fit one SM");

struct AttnFwdParams {
  int N, S, H;
  int heads_per_cta;          // consecutive heads served by one CTA of the general kernel
  int max_tiles;              // allocation stride of the tile arrays (= ceil(S/64))
  int mask_words;             // uint32 words per mask row (multiple of 4, padded by 4)
  const uint32_t* mask_bits;  // [N, S, mask_words]
  const int* tile_start;      // [N, max_tiles+1]  row offsets of the variable row tiles
  const int* n_tiles;         // [N]
  const uint8_t* tile_cls;    // [N, max_tiles, max_tiles]  (query tile, key tile)
  const uint8_t* iso_flags;   // [N, max_tiles] tiles handled by the persistent diagonal kernel (may be NULL)
  __nv_bfloat16* out;         // [N*S, H*64]
  __nv_bfloat16* out_lo;      // [N*S, H*64] bf16(O - bf16(O)), same ld (may be NULL): makes D = dO.(O + O_lo) exact in backward
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

// Softmax work of one thread: HALF a query row (64 of the tile's 128 keys; two warps share a TMEM lane quadrant and split
// the columns).  The 64 scores arrive in REGISTERS — the caller has already drained them from TMEM and handed the TMEM
// buffer back to the MMA thread, so the score MMA of the next tile overlaps these functions.
template <bool MASKED>
__device__ __forceinline__ float fwd_half_max(const uint32_t (&s)[2][32], const uint32_t (&mw)[2]) {
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains instead of a 64-deep max
#pragma unroll
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float v = __uint_as_float(s[c][j]);
      if (MASKED) v = ((mw[c] >> j) & 1u) ? v : -INFINITY;
      m4[j & 3] = fmaxf(m4[j & 3], v);
    }
  }
  return fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
}

// P = exp2(S * scale - m_used) (masked, then dropped) packed to bf16 IN PLACE: s[c][0..15] receive the 32 columns of
// chunk c; returns the (undropped) sum of the half row.
template <bool MASKED, bool DROP>
__device__ __forceinline__ float fwd_half_exp(uint32_t (&s)[2][32], const uint32_t (&mw)[2], float scale_log2, float m_used,
                                              uint32_t rk1, uint32_t rk2, int k0, uint32_t thresh) {
  float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float pv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float e = fast_exp2(fmaf(__uint_as_float(s[c][g * 8 + j]), scale_log2, -m_used));
        if (MASKED) e = ((mw[c] >> (g * 8 + j)) & 1u) ? e : 0.f;
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
      s[c][g * 4 + 0] = pack_bf16(pv[0], pv[1]);
      s[c][g * 4 + 1] = pack_bf16(pv[2], pv[3]);
      s[c][g * 4 + 2] = pack_bf16(pv[4], pv[5]);
      s[c][g * 4 + 3] = pack_bf16(pv[6], pv[7]);
    }
  }
  return (l4[0] + l4[1]) + (l4[2] + l4[3]);
}

// Pipeline of one CTA (the two resident CTAs of an SM interleave two such pipelines):
//   MMA thread     : S(0) | for g: [s_empty(g) & K(g+1)] -> S(g+1) ; per key half hc: [p_full[hc](g) & V(g)] ->
//                    O_hc (+)= P_hc(g) V_hc(g)
//   softmax warps  : s_full(g) -> scores TMEM -> registers -> arrive s_empty   (the S buffer is free again: the score
//                    MMA of tile g+1 runs while this tile's max / exp / pack are computed)
//                    -> maximum -> lazy rescale of O_hc in TMEM -> exp2 / mask / dropout / pack -> pv_done[hc](g-1)
//                    -> P_hc to smem -> arrive p_full[hc]
// EIGHT softmax warps, two threads per query row.  The two threads of a row split the KEYS of every tile (64 each) and
// run two INDEPENDENT online softmaxes: each has its own reference maximum, its own row sum and its own accumulator
// O_hc = sum over its keys of P V in TMEM (the PV product of a tile is issued as two K = 64 halves into two 64-column
// accumulators), and the halves are merged once per row in the epilogue,
//     O = (w_0 O_0 + w_1 O_1) / (w_0 l_0 + w_1 l_1),   w_hc = 2^(m_hc - max(m_0, m_1)).
// No value ever crosses a warp inside the key loop: a first version that kept one accumulator per row had the two
// threads agree on the maximum every tile (named barrier + TMEM exchange) and spent 17 % of its time there (clock64
// trace, tools/attn_trace.py).  With four warps per CTA the MUFU pipe, the binding resource at head_dim 64, was busy
// 36 % of the time (ncu): the phases of a warp that do not use it (barrier waits, the TMEM drain, the maximum, the P
// stores and their proxy fence, the output epilogue) are covered by the exponentials of the scheduler's other warps.
// The maximum each half's sums are expressed against (m_used) is only advanced — and O_hc / l rescaled, by the row's own
// thread through tcgen05.ld / st — when the new maximum exceeds it by more than 8 (log2 domain): stale maxima leave
// P <= 2^8, harmless in bf16 / fp32, and after the first tile or two the rescale almost never fires (the
// FlashAttention-4 "lazy rescale").
constexpr float kRescaleThreshold = 8.0f;

#ifdef GGPT_ATTN_TRACE      // profiling aid: per-phase clock64 stamps of one softmax warp of a few CTAs (tools/attn_trace.py)
__device__ long long g_attn_trace[64 * 16 * 8];
#define ATTN_TRACE(slot)                                                                                        \
  do {                                                                                                          \
    if (trace_on && lane == 0 && it < 16) g_attn_trace[(trace_cta * 16 + it) * 8 + (slot)] = clock64();          \
  } while (0)
#define ATTN_TRACE_H(slot)                                                                                      \
  do {                                                                                                          \
    if (trace_on && lane == 0) g_attn_trace[(trace_cta * 16 + 8 + (h - h_begin)) * 8 + (slot)] = clock64();      \
  } while (0)
#else
#define ATTN_TRACE(slot) do { } while (0)
#define ATTN_TRACE_H(slot) do { } while (0)
#endif

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
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]  score MMA that read the stage has retired
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]  both PV halves that read the stage have retired
  uint64_t* s_full = bars + 9;     // scores of tile g in TMEM
  uint64_t* s_empty = bars + 10;   // ... and drained into registers (8 warps)
  uint64_t* p_full = bars + 11;    // [2] P_hc(g) in smem, O_hc rescaled if needed (4 warps each)
  uint64_t* pv_done = bars + 13;   // [2] O_hc += P_hc(g) V_hc(g) retired: P_hc buffer reusable, O_hc readable
  uint64_t* q_empty = bars + 15;   // every score MMA of the current head has retired: Q buffer reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 1016);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  // One CTA serves ONE query tile for `heads_per_cta` consecutive heads: the tile plan, the TMEM allocation and the
  // barrier set-up are paid once, and the output epilogue of head h overlaps the Q / K / V loads and the first score MMA
  // of head h+1.
  const int qt = blockIdx.x, n = blockIdx.z;
  const int h_begin = blockIdx.y * p.heads_per_cta;
  const int h_end = min(p.H, h_begin + p.heads_per_cta);
  const int n_kt = p.n_tiles[n];
  if (qt >= n_kt) return;                       // uniform for the whole CTA, before any barrier / TMEM state
  const int* gts = p.tile_start + static_cast<size_t>(n) * (p.max_tiles + 1);
  const int q0 = gts[qt], qlen = gts[qt + 1] - q0;
  if (p.iso_flags != nullptr && p.iso_flags[n * p.max_tiles + qt]) {   // done by attn_diag_fwd_kernel
    if (p.out_lo != nullptr) {                  // ... which stores no rounding residual: D of these rows uses bf16 O only
      const int row_chunks = (h_end - h_begin) * 8;
      for (int i = threadIdx.x; i < qlen * row_chunks; i += blockDim.x)
        *reinterpret_cast<uint4*>(p.out_lo + (static_cast<long long>(n) * p.S + q0 + i / row_chunks) * p.ldo + h_begin * kHeadDim +
                                  (i % row_chunks) * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  // this query tile's row of the tile plan (row offsets as uint16: S <= 16384; classes) is copied to shared memory once —
  // every role reads it for every key tile, and a global load there is an L2 round trip on each tile's critical path
  uint16_t* ts = reinterpret_cast<uint16_t*>(bars + 16);               // [<= 257]
  uint8_t* cls_row = reinterpret_cast<uint8_t*>(bars + 16) + 520;      // [<= 256]
  {
    const uint8_t* g_cls = p.tile_cls + (static_cast<size_t>(n) * p.max_tiles + qt) * p.max_tiles;
    for (int t = threadIdx.x; t <= n_kt; t += blockDim.x) {
      ts[t] = static_cast<uint16_t>(gts[t]);
      if (t < n_kt) cls_row[t] = g_cls[t];
    }
  }

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("ggpt attn_fwd: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&p_full[i], kAttSoftmaxWarps / 2);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, kAttSoftmaxWarps);
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
  const uint32_t tmem_O = tmem_base + 128;   // 2 x 64 columns: O_0 (keys 0..63 of every tile) | O_1 (keys 64..127)

  int n_active = 0;                          // active key tiles of this query tile (the same for every head)
  for (int kt = 0; kt < n_kt; ++kt) n_active += (cls_row[kt] != 0);
  // g = running index of (head, active key tile) pairs: stage = g & 1 and phase = (g >> 1) & 1 of the K / V rings,
  // parity g & 1 of the per-tile barriers — all barrier phases simply continue from one head to the next

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && n_active > 0) {
      int g = 0;
      for (int h = h_begin; h < h_end; ++h) {
        if (h > h_begin) mbar_wait(q_empty, (h - h_begin - 1) & 1);
        mbar_expect_tx(q_full, 16384);
        tma_load_3d(sQ, &tmQKV, q_full, p.q_col0 + h * kHeadDim, q0, n);
        for (int kt = 0; kt < n_kt; ++kt) {
          if (cls_row[kt] == 0) continue;
          const int st = g & 1;
          const uint32_t ph = ((g >> 1) & 1) ^ 1;
          mbar_wait(&k_empty[st], ph);
          mbar_expect_tx(&k_full[st], 16384);
          tma_load_3d(sK + st * 16384, &tmQKV, &k_full[st], p.k_col0 + h * kHeadDim, ts[kt], n);
          mbar_wait(&v_empty[st], ph);
          mbar_expect_tx(&v_full[st], 16384);
          tma_load_3d(sV + st * 16384, &tmQKV, &v_full[st], p.v_col0 + h * kHeadDim, ts[kt], n);
          ++g;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (n_active > 0) {                      // whole warp, one elected lane issues (see tc_mma_bf16_e)
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      auto issue_S = [&](int g) {
        const int st = g & 1;
        mbar_wait(&k_full[st], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + st * 16384);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_bf16_e(tmem_S, umma_desc_sw128(aQ + kk * 32, 16, 1024), umma_desc_sw128(aK + kk * 32, 16, 1024),
                      idesc_s, kk != 0);
        tc_commit_e(s_full);
        tc_commit_e(&k_empty[st]);
      };
      int g0 = 0;
      for (int h = h_begin; h < h_end; ++h, g0 += n_active) {
        mbar_wait(q_full, (h - h_begin) & 1);
        if (g0 > 0) mbar_wait(s_empty, (g0 - 1) & 1);   // the last scores of the previous head are in registers
        tc_fence_after();
        issue_S(g0);
        for (int it = 0; it < n_active; ++it) {
          const int g = g0 + it, st = g & 1;
          if (it + 1 < n_active) {
            mbar_wait(s_empty, g & 1);      // S(g) sits in the softmax warps' registers
            tc_fence_after();
            issue_S(g + 1);
          } else {
            tc_commit_e(q_empty);             // every score MMA of this head has been issued
          }
          mbar_wait(&v_full[st], (g >> 1) & 1);
          const uint32_t aV = smem_u32(sV + st * 16384);
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {  // keys [64 hc, 64 hc + 64) of the tile -> accumulator O_hc
            mbar_wait(&p_full[hc], g & 1);  // P_hc(g) in smem, O_hc rescaled
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma_bf16_e(tmem_O + hc * 64, umma_desc_sw128(aP + hc * 16384 + kk * 32, 16, 1024),
                          umma_desc_sw128(aV + (hc * 4 + kk) * 2048, 8192, 1024), idesc_o, (it | kk) != 0);
            tc_commit_e(&pv_done[hc]);
          }
          tc_commit_e(&v_empty[st]);
        }
      }
    }
  } else {
    // ===================== softmax / epilogue warps: two threads per query row, 64 keys of every tile each ==============
    const int quad = warp & 3;
    const int hc = (warp - 2) >> 2;              // key half of every tile / column half of the output row
    const int r = quad * 32 + lane;              // row inside the tile
    const int q_row = q0 + r;
    const bool row_ok = r < qlen;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t* mrow = p.mask_bits + (static_cast<size_t>(n) * p.S + (row_ok ? q_row : 0)) * p.mask_words;
    uint8_t* prow = sP + hc * 16384 + r * 128;   // this thread's 64 columns are one 128-byte row of K block hc
    const uint32_t tmem_Omine = tmem_O + lane_addr + hc * 64;
    const uint32_t tmem_Oother = tmem_O + lane_addr + (hc ^ 1) * 64;
#ifdef GGPT_ATTN_TRACE
    const int trace_cta = (blockIdx.z % 8) * 8 + blockIdx.x % 8;     // 64 traced CTAs: head group 1, sequences 8..15
    const bool trace_on = (warp == 2) && blockIdx.y == 1 && blockIdx.z >= 8 && blockIdx.z < 16;
#endif

    int g0 = 0;
    for (int h = h_begin; h < h_end; ++h, g0 += n_active) {
      const uint32_t rowkey = drop_rowkey(p.drop.seed_lo, p.drop.seed_hi, n, h, q_row);
      const uint32_t rowkey2 = drop_rowkey2(rowkey);
      float m_used = 0.f;     // log2-domain maximum this half's P / l / O_hc are expressed against
      bool seen = false;      // some key of this half of the row has been visible so far
      float l_run = 0.f;      // this half's sum

      int it = 0;
      ATTN_TRACE_H(0);
      for (int kt = 0; kt < n_kt; ++kt) {
        const int cls = cls_row[kt];
        if (cls == 0) continue;
        const int g = g0 + it;
        ATTN_TRACE(0);
        uint32_t mw[2] = {0xffffffffu, 0xffffffffu};
        if (cls == 2) {
          if (row_ok) {
            uint32_t w4[4];
            load_mask_words(mrow, ts[kt], ts[kt + 1] - ts[kt], w4);
            mw[0] = hc ? w4[2] : w4[0];
            mw[1] = hc ? w4[3] : w4[1];
          } else {
            mw[0] = mw[1] = 0u;
          }
        }
        const int k0 = ts[kt] + hc * 64;
        uint32_t s[2][32];
        mbar_wait(s_full, g & 1);
        ATTN_TRACE(6);
        tc_fence_after();
        tmem_ld32(tmem_S + lane_addr + hc * 64, s[0]);
        tmem_ld32(tmem_S + lane_addr + hc * 64 + 32, s[1]);
        tmem_ld_wait();
        ATTN_TRACE(1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);

        const float m_half = ((cls == 2) ? fwd_half_max<true>(s, mw) : fwd_half_max<false>(s, mw)) * p.scale_log2;
        bool need = false;
        float alpha = 1.f;
        if (m_half != -INFINITY) {
          if (!seen) {                 // first visible keys of this half: O_hc and l are still exactly zero
            seen = true;
            m_used = m_half;
          } else if (m_half > m_used + kRescaleThreshold) {
            need = true;
            alpha = fast_exp2(m_used - m_half);
            m_used = m_half;
          }
        }
        ATTN_TRACE(2);
        if (it > 0 && __any_sync(0xffffffffu, need)) {   // warp-uniform: tcgen05.ld / st are warp-collective
          mbar_wait(&pv_done[hc], (g - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {                  // this half's 64 accumulator columns, 8 at a time
            uint32_t o[8];
            tmem_ld8(tmem_Omine + c * 8, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st8(tmem_Omine + c * 8, o);
          }
          tmem_st_wait();
          l_run *= alpha;
        }
        const float l_tile = (cls == 2)
            ? fwd_half_exp<true, DROP>(s, mw, p.scale_log2, m_used, rowkey, rowkey2, k0, p.drop.thresh)
            : fwd_half_exp<false, DROP>(s, mw, p.scale_log2, m_used, rowkey, rowkey2, k0, p.drop.thresh);
        l_run += l_tile;
        ATTN_TRACE(3);
        // the PV MMA of the previous tile has read this half's P buffer (tile 0 of a head: the epilogue of the previous
        // head has waited for both last PVs, and its final pair barrier ordered this store behind the partner's flush of
        // the staging rows that alias this buffer)
        if (it > 0) mbar_wait(&pv_done[hc], (g - 1) & 1);
        ATTN_TRACE(4);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const int chunk = (c * 4 + gq) ^ (r & 7);
            *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(s[c][gq * 4 + 0], s[c][gq * 4 + 1], s[c][gq * 4 + 2], s[c][gq * 4 + 3]);
          }
        }
        fence_proxy_async_smem();   // make P visible to the tensor-core (async) proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[hc]);
        ATTN_TRACE(5);
        ++it;
      }

      {
        // Merge of the two key halves and output.  Every PV of this head has retired, so the P buffer is free: first it
        // carries each thread's (m, l) to its partner — written into the 16-byte chunk of the PARTNER's part of the
        // staging row, which only the partner overwrites later, so one pair barrier suffices — then it stages the
        // output: the two warps of a quadrant fill one 32 x 128 B block (32 output columns each) and flush 16 rows each
        // as full 128-byte lines.
        ATTN_TRACE_H(1);
        if (n_active > 0) {
          const int gl = g0 + n_active - 1;
          mbar_wait(&pv_done[hc], gl & 1);
          mbar_wait(&pv_done[hc ^ 1], gl & 1);
          tc_fence_after();
        }
        ATTN_TRACE_H(2);
        uint8_t* stg = sP + quad * 4096;
        uint8_t* stg_lo = sP + 16384 + quad * 4096;
        const float m_mine = seen ? m_used : -INFINITY;
        *reinterpret_cast<float2*>(stg + lane * 128 + ((((hc ^ 1) * 4) ^ (lane & 7)) << 4)) = make_float2(m_mine, l_run);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        const float2 other = *reinterpret_cast<const float2*>(stg + lane * 128 + (((hc * 4) ^ (lane & 7)) << 4));
        const float m_row = fmaxf(m_mine, other.x);
        const float w_mine = (m_mine == -INFINITY) ? 0.f : fast_exp2(m_mine - m_row);
        const float w_other = (other.x == -INFINITY) ? 0.f : fast_exp2(other.x - m_row);
        const float l_row = l_run * w_mine + other.y * w_other;
        const float inv = (l_row > 0.f) ? p.drop.inv_keep / l_row : 0.f;   // dropout keeps are rescaled by 1/(1-p)
        const float f_mine = w_mine * inv, f_other = w_other * inv;
        ATTN_TRACE_H(3);
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {         // this thread's 32 output columns [32 hc, 32 hc + 32), 16 at a time
          uint32_t oa[16], ob[16];
          if (n_active > 0) {
            tmem_ld16(tmem_Omine + hc * 32 + cb * 16, oa);
            tmem_ld16(tmem_Oother + hc * 32 + cb * 16, ob);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) oa[j] = ob[j] = 0u;
          }
#pragma unroll
          for (int gq = 0; gq < 2; ++gq) {
            float f[8];
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              f[j] = __uint_as_float(oa[gq * 8 + j]) * f_mine + __uint_as_float(ob[gq * 8 + j]) * f_other;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              hi[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
              const float2 back = unpack_bf16(hi[j]);
              lo[j] = pack_bf16(f[2 * j] - back.x, f[2 * j + 1] - back.y);     // rounding residual of the bf16 output
            }
            stage_put16(stg, lane, hc * 4 + cb * 2 + gq, make_uint4(hi[0], hi[1], hi[2], hi[3]));
            if (p.out_lo != nullptr) stage_put16(stg_lo, lane, hc * 4 + cb * 2 + gq, make_uint4(lo[0], lo[1], lo[2], lo[3]));
          }
        }
        tc_fence_before();
        ATTN_TRACE_H(4);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        ATTN_TRACE_H(5);
        const long long row0 = static_cast<long long>(n) * p.S + q0 + quad * 32 + hc * 16;
        stage_flush_rows16(stg, lane, hc * 16, p.out + row0 * p.ldo + h * kHeadDim, p.ldo, qlen - quad * 32 - hc * 16);
        if (p.out_lo != nullptr)
          stage_flush_rows16(stg_lo, lane, hc * 16, p.out_lo + row0 * p.ldo + h * kHeadDim, p.ldo,
                             qlen - quad * 32 - hc * 16);
        if (row_ok && hc == 0 && p.lse != nullptr) {
          const float lse = (l_row > 0.f) ? (m_row * 0.6931471805599453f + logf(l_row)) : 0.f;
          p.lse[(static_cast<size_t>(n) * p.H + h) * p.S + q_row] = lse;
        }
        // the partner's flush reads staging rows that alias THIS thread's next P row: order the next head behind it
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        ATTN_TRACE_H(6);
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

#ifdef GGPT_ATTN_TRACE
int ggpt_debug_diag_trace(long long* out, int n) { return ggpt::diag_trace_read(out, n); }
int ggpt_debug_bwd_trace(long long* out, int n) { return ggpt::bwd_trace_read(out, n); }
int ggpt_debug_attn_trace(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, ggpt::g_attn_trace, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
#endif

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
                  void* out, void* out_lo, long long ldo, float* lse, int N, int S, int H, void* stream) {
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
  p.out = static_cast<__nv_bfloat16*>(out); p.out_lo = static_cast<__nv_bfloat16*>(out_lo); p.ldo = ldo; p.lse = lse;
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
    // heads per CTA: 3 when that divides H (12 heads -> 4 groups), else 4 / 2 / 1
    p.heads_per_cta = (H % 3 == 0) ? 3 : (H % 4 == 0) ? 4 : (H % 2 == 0) ? 2 : 1;
    static const char* hpc_env = getenv("GGPT_ATTN_HEADS_PER_CTA");
    if (hpc_env != nullptr && atoi(hpc_env) > 0) p.heads_per_cta = atoi(hpc_env);
    dim3 grid(p.max_tiles, (H + p.heads_per_cta - 1) / p.heads_per_cta, N);
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
