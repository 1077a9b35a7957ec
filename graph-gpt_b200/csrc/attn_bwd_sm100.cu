// Attention backward on tcgen05 (sm_100a): dQ, dK, dV of softmax(Q K^T / 8 + mask) V with recomputation of the
// probabilities from the saved log-sum-exp (no [N,H,S,S] tensor is ever stored).  autograd of HF:199-221.
//
// Two deterministic passes of one templated kernel (no atomics):
//   DKV : one CTA per (key tile, head, sequence), streams the query tiles:   dV += P^T dO,  dK += dS^T Q
//   DQ  : one CTA per (query tile, head, sequence), streams the key tiles:   dQ += dS K
// with  S = Q K^T,  P = exp(S/8 - lse),  dP = dO V^T,  dS = P * (dP - D) / 8,  D = rowsum(dO * O).
// All five products run on tcgen05 with TMEM accumulators; P / dS are written by the softmax warps to shared
// memory in the 128-byte-swizzled layout and consumed either as a K-major A operand (dS K) or, through the
// MN-major descriptor, as the transposed A operand (P^T dO, dS^T Q) — no explicit transposes.  The epilogue
// un-rotates dQ / dK (inverse RoPE) so the result is the gradient of the fused q|k|v projection output.
//
// Row-sum consistency: D is formed from the bf16 forward output, so sum_k dS[q,k] = -eps_q is not exactly zero and
// dQ would carry the common-mode error -eps_q * (sum_k P[q,k] K[k]) — visible when keys share a large common
// component (rows of identical <mask> tokens).  The DQ pass therefore also accumulates PK = P K (one extra
// 128x64x128 MMA per tile) and eps_q = sum_k P (dP - D) in registers, and emits dQ = dS K - eps_q/8 * PK, which is
// the gradient for the exact D = sum_k P dP.  dK and dV have no such common-mode term (eps varies per query row).
#include "common.cuh"
#include "attn_diag.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

// TMA producer, MMA issuer, 16 softmax-gradient warps: four per TMEM lane quadrant, i.e. four threads per query row with
// one 32-column chunk each.  The softmax-gradient phase is a chain of dependent fixed-latency instructions (FFMA -> MUFU ->
// FMUL -> F2FP -> STS); with two warps per scheduler it ran at ~0.15 instructions / clk / warp (ncu: issue slots 30 %
// busy, tensor pipe 20 %), so the remedy is more resident warps, not fewer instructions.
constexpr int kBwdSoftmaxWarps = 16;
constexpr int kBwdThreads = 64 + 32 * kBwdSoftmaxWarps;

struct AttnBwdParams {
  int N, S, H;
  int max_tiles;              // allocation stride of the tile arrays (= ceil(S/64))
  int mask_words;
  const uint32_t* mask_bits;  // [N,S,words]
  const int* tile_start;      // [N, max_tiles+1]  variable row tiles (see attn_fwd_sm100.cu)
  const int* n_tiles;         // [N]
  const uint8_t* tile_cls;    // [N,max_tiles,max_tiles]  (query tile major)
  const uint8_t* iso_flags;   // [N,max_tiles] tiles handled by attn_diag_bwd_kernel (may be NULL)
  const float* lse;           // [N,H,S]
  const float* dsum;          // [N,H,S]   D = rowsum(dO*O)
  __nv_bfloat16* dqkv;        // [N*S, ld]
  long long ld;
  int q_col0, k_col0, v_col0;
  const int* pos;             // [N*S]
  const float* cos_tab;
  const float* sin_tab;
  float scale;                // 1/8
  float scale_log2;           // scale * log2(e)
  DropParams drop;            // attention-probability dropout (same mask as the forward)
};

// smem (one CTA per SM): fixed pair 2x16K | streamed pair 2 stages x 2 x 16K | P 2 x 32K | dS 2 x 32K | eps | barriers.
// P / dS are DOUBLE-buffered so that the softmax-gradient warps can write tile i+1 while the tensor core still reads
// tile i (the gradient MMAs of tile i are issued AFTER the score MMAs of tile i+1, see the MMA issuer).  The epilogue's
// RoPE-table gather buffers alias the streamed stages and its eps exchange the P buffers (idle once the last MMA has
// retired).
template <bool DKV>
struct BwdSmem {
  static constexpr int kFixed = 0;
  static constexpr int kStream = 32768;
  static constexpr int kP = kStream + 65536;       // 2 x 32 KB
  static constexpr int kDS = kP + 65536;           // 2 x 32 KB
  static constexpr int kGather = kStream;          // 8 warps x 4 KB, epilogue only
  static constexpr int kEps = kP;                  // float [4 chunks][128]: per-row eps partial sums (DQ epilogue; aliases P)
  static constexpr int kPlan = kDS + 65536;       // int ts[<= 257] | uint8 cls[<= 256]: this sequence's tile plan
  static constexpr int kBars = kPlan + 1536;
  static constexpr int kTotal = kBars + 256;
};

// 128 mask bits of query row `mrow` for the key tile starting at key k0 (any alignment) with klen valid keys.
__device__ __forceinline__ void load_mask_words_bwd(const uint32_t* __restrict__ mrow, int k0, int klen, uint32_t (&mw)[4]) {
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

// One 32-column chunk of a thread's query row: P = exp(S/8 - lse), dS = P (dP - D) / 8 -> bf16 into the swizzled P / dS
// buffers.  MASKED / DROP are compile-time so that full tiles and dropout-free runs carry no mask or RNG instructions
// (predicated-off instructions still take issue slots: the kernel is issue-bound once the pipeline is full).
template <bool DKV, bool MASKED, bool DROP>
__device__ __forceinline__ void bwd_chunk(const uint32_t (&s)[32], const uint32_t (&dp)[32], int c, uint32_t w, int r,
                                          uint8_t* sP, uint8_t* sDS, float scale_log2, float scale, float lse2, float dsum,
                                          uint32_t rk1, uint32_t rk2, int kbase, uint32_t thresh, float inv_keep,
                                          float& eps_run) {
  uint8_t* pbase = sP + (c >> 1) * 16384 + r * 128;
  uint8_t* dbase = sDS + (c >> 1) * 16384 + r * 128;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float pv[8], dv[8];
    uint32_t bits = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float e = fast_exp2(fmaf(__uint_as_float(s[g * 8 + j]), scale_log2, -lse2));
      if (MASKED) e = ((w >> (g * 8 + j)) & 1u) ? e : 0.f;
      pv[j] = e;
      float dpe = __uint_as_float(dp[g * 8 + j]);
      bool keep = true;
      if (DROP) {
        if ((j & 1) == 0) bits = drop_bits(rk1, rk2, kbase + c * 32 + g * 8 + j);
        keep = (j & 1) ? drop_keep_odd(bits, thresh) : drop_keep_even(bits, thresh);
        dpe = keep ? dpe * inv_keep : 0.f;
      }
      const float t0 = pv[j] * (dpe - dsum);
      if (!DKV) eps_run += t0;
      dv[j] = t0 * scale;
      if (DROP && DKV && !keep) pv[j] = 0.f;   // DKV: sP feeds dV = (P o mask)^T dO / (1-p); DQ: sP feeds P K (undropped)
    }
    const int chunk = ((c & 1) * 4 + g) ^ (r & 7);
    uint4 o;
    o.x = pack_bf16(pv[0], pv[1]); o.y = pack_bf16(pv[2], pv[3]);
    o.z = pack_bf16(pv[4], pv[5]); o.w = pack_bf16(pv[6], pv[7]);
    *reinterpret_cast<uint4*>(pbase + chunk * 16) = o;
    uint4 o2;
    o2.x = pack_bf16(dv[0], dv[1]); o2.y = pack_bf16(dv[2], dv[3]);
    o2.z = pack_bf16(dv[4], dv[5]); o2.w = pack_bf16(dv[6], dv[7]);
    *reinterpret_cast<uint4*>(dbase + chunk * 16) = o2;
  }
}

// This thread's 32-column chunk c (columns [32 c, 32 c + 32)) of one tile.
template <bool DKV, bool MASKED, bool DROP>
__device__ __forceinline__ void bwd_tile(uint32_t tS, uint32_t tDP, int c, uint32_t mw_c, int r, uint8_t* sP,
                                         uint8_t* sDS, float scale_log2, float scale, float lse2, float dsum, uint32_t rk1,
                                         uint32_t rk2, int kbase, uint32_t thresh, float inv_keep, float& eps_run) {
  uint32_t s0[32], d0[32];
  tmem_ld32(tS + c * 32, s0);
  tmem_ld32(tDP + c * 32, d0);
  tmem_ld_wait();
  bwd_chunk<DKV, MASKED, DROP>(s0, d0, c, mw_c, r, sP, sDS, scale_log2, scale, lse2, dsum, rk1, rk2, kbase, thresh, inv_keep,
                               eps_run);
}

template <bool DKV, bool DROP>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const AttnBwdParams p) {
  using L = BwdSmem<DKV>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sFix0 = smem + L::kFixed;            // DKV: K_j      DQ: Q_i
  uint8_t* sFix1 = smem + L::kFixed + 16384;    // DKV: V_j      DQ: dO_i
  uint8_t* sStr = smem + L::kStream;            // stage s: [s*32768] = (DKV ? Q_i : K_j), [+16384] = (DKV ? dO_i : V_j)
  uint8_t* sP = smem + L::kP;                   // [2][32768]
  uint8_t* sDS = smem + L::kDS;                 // [2][32768]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint64_t* fix_full = bars + 0;
  uint64_t* str_full = bars + 1;    // [2]
  uint64_t* str_empty = bars + 3;   // [2]
  uint64_t* sdp_full = bars + 5;    // S and dP accumulators of the next tile ready
  uint64_t* pds_full = bars + 6;    // [2] P / dS of tile i written to buffer i & 1, S / dP drained     (count 8)
  uint64_t* pds_empty = bars + 8;   // [2] gradient MMAs that read buffer i & 1 have retired
  uint64_t* acc_full = bars + 10;   // final accumulators ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int n_t = p.n_tiles[n];
  if (tile >= n_t) return;                      // uniform for the whole CTA, before any barrier / TMEM state
  if (p.iso_flags != nullptr && p.iso_flags[n * p.max_tiles + tile]) return;   // done by attn_diag_bwd_kernel
  const int* ts = p.tile_start + static_cast<size_t>(n) * (p.max_tiles + 1);
  const int row_base = ts[tile], row_len = ts[tile + 1] - row_base;   // rows of the fixed tile
  const uint8_t* cls_n = p.tile_cls + static_cast<size_t>(n) * p.max_tiles * p.max_tiles;
  // the tile plan of this sequence (row offsets, class of (query tile, key tile) for every streamed index t) is copied
  // to shared memory once: it is consulted by every role for every tile, and a global load on that path costs an L2
  // round trip per tile
  int* s_ts = reinterpret_cast<int*>(smem + L::kPlan);
  uint8_t* s_cls = reinterpret_cast<uint8_t*>(smem + L::kPlan + 1040);     // S <= 16384 => at most 256 row tiles
  for (int t = threadIdx.x; t <= n_t; t += blockDim.x) {
    s_ts[t] = ts[t];
    if (t < n_t) s_cls[t] = DKV ? cls_n[t * p.max_tiles + tile] : cls_n[tile * p.max_tiles + t];
  }
  auto cls_of = [&](int t) -> int { return s_cls[t]; };

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("ggpt attn_bwd: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(fix_full, 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&str_full[i], 1);
      mbar_init(&str_empty[i], 1);
      mbar_init(&pds_full[i], kBwdSoftmaxWarps);
      mbar_init(&pds_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;           // [0,128)
  const uint32_t tmem_dP = tmem_base + 128;    // [128,256)
  const uint32_t tmem_A0 = tmem_base + 256;    // DKV: dV   DQ: dQ      (64 columns)
  const uint32_t tmem_A1 = tmem_base + 320;    // DKV: dK   DQ: PK = P K (row-sum correction)

  int n_active = 0;
  for (int t = 0; t < n_t; ++t) n_active += (cls_of(t) != 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && n_active > 0) {
      mbar_expect_tx(fix_full, 32768);
      if (DKV) {
        tma_load_3d(sFix0, &tmQKV, fix_full, p.k_col0 + h * 64, row_base, n);
        tma_load_3d(sFix1, &tmQKV, fix_full, p.v_col0 + h * 64, row_base, n);
      } else {
        tma_load_3d(sFix0, &tmQKV, fix_full, p.q_col0 + h * 64, row_base, n);
        tma_load_3d(sFix1, &tmDO, fix_full, h * 64, row_base, n);
      }
      int it = 0;
      for (int t = 0; t < n_t; ++t) {
        if (cls_of(t) == 0) continue;
        const int st = it & 1;
        mbar_wait(&str_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&str_full[st], 32768);
        if (DKV) {
          tma_load_3d(sStr + st * 32768, &tmQKV, &str_full[st], p.q_col0 + h * 64, s_ts[t], n);
          tma_load_3d(sStr + st * 32768 + 16384, &tmDO, &str_full[st], h * 64, s_ts[t], n);
        } else {
          tma_load_3d(sStr + st * 32768, &tmQKV, &str_full[st], p.k_col0 + h * 64, s_ts[t], n);
          tma_load_3d(sStr + st * 32768 + 16384, &tmQKV, &str_full[st], p.v_col0 + h * 64, s_ts[t], n);
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_active > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);     // S, dP : K-major x K-major
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, true, true);        // P^T dO, dS^T Q : MN x MN
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, false, true);       // dS K : K-major x MN-major
      const uint32_t aFix0 = smem_u32(sFix0), aFix1 = smem_u32(sFix1);
      auto issue_scores = [&](int it) {
        const int st = it & 1;
        mbar_wait(&str_full[st], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sStr + st * 32768), a1 = a0 + 16384;
        const uint32_t aQ = DKV ? a0 : aFix0, aK = DKV ? aFix0 : a0;
        const uint32_t aDO = DKV ? a1 : aFix1, aV = DKV ? aFix1 : a1;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_bf16(tmem_S, umma_desc_sw128(aQ + kk * 32, 16, 1024), umma_desc_sw128(aK + kk * 32, 16, 1024),
                      idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_bf16(tmem_dP, umma_desc_sw128(aDO + kk * 32, 16, 1024), umma_desc_sw128(aV + kk * 32, 16, 1024),
                      idesc_s, kk != 0);
        tc_commit(sdp_full);
      };
      mbar_wait(fix_full, 0);
      issue_scores(0);
      for (int it = 0; it < n_active; ++it) {
        const int st = it & 1;
        mbar_wait(&pds_full[st], (it >> 1) & 1);   // P / dS of tile it in buffer st; S / dP drained from TMEM
        tc_fence_after();
        // the scores of the NEXT tile go first: the softmax-gradient warps start on them while the gradient MMAs of this
        // tile run (P / dS are double-buffered, so they may write buffer st ^ 1 meanwhile)
        if (it + 1 < n_active) issue_scores(it + 1);
        const uint32_t a0 = smem_u32(sStr + st * 32768), a1 = a0 + 16384;
        const uint32_t aP = smem_u32(sP + st * 32768), aDS = smem_u32(sDS + st * 32768);
        if (DKV) {
          // dV += P^T dO_i ; dK += dS^T Q_i      (A = P / dS read MN-major: M = keys, K = query rows)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            tc_mma_bf16(tmem_A0, umma_desc_sw128(aP + kk * 2048, 16384, 1024),
                        umma_desc_sw128(a1 + kk * 2048, 8192, 1024), idesc_t, (it | kk) != 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            tc_mma_bf16(tmem_A1, umma_desc_sw128(aDS + kk * 2048, 16384, 1024),
                        umma_desc_sw128(a0 + kk * 2048, 8192, 1024), idesc_t, (it | kk) != 0);
        } else {
          // dQ += dS K_j ; PK += P K_j      (A = dS / P K-major: M = query rows, K = keys; B = K_j MN-major)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            tc_mma_bf16(tmem_A0, umma_desc_sw128(aDS + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                        umma_desc_sw128(a0 + kk * 2048, 8192, 1024), idesc_q, (it | kk) != 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            tc_mma_bf16(tmem_A1, umma_desc_sw128(aP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                        umma_desc_sw128(a0 + kk * 2048, 8192, 1024), idesc_q, (it | kk) != 0);
        }
        tc_commit(&pds_empty[st]);
        tc_commit(&str_empty[st]);
        if (it + 1 == n_active) tc_commit(acc_full);
      }
    }
  } else {
    // ===================== softmax-gradient warps: four threads per query row (32 key columns each) =====================
    const int quad = warp & 3;
    const int chunk_c = (warp - 2) >> 2;      // which 32-column chunk of the score tile this thread owns
    const int half = chunk_c & 1;             // epilogue role (warps with chunk_c < 2 write the outputs)
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float eps_run = 0.f;   // DQ: sum_k P (dP - D) of this thread's query row

    // per-(tile, row) inputs; fetched one tile ahead so that their global-memory latency is off the critical path
    // (only RAW loaded values are kept: an in-order warp stalls at the first instruction that consumes a load, so every
    // consumer — the log2(e) scaling, the funnel shifts of the mask words — sits at the point of use, one tile later)
    struct RowIn {
      float lse, dsum;
      uint32_t w[2];                          // the two mask words that cover this thread's 32 columns
      int q_row, kbase, klen, cls;
      bool row_ok;
    };
    auto fetch = [&](int t, RowIn& ri) {
      ri.cls = s_cls[t];
      const int qt = DKV ? t : tile;
      const int kt = DKV ? tile : t;
      ri.q_row = s_ts[qt] + r;
      ri.row_ok = r < s_ts[qt + 1] - s_ts[qt];
      ri.kbase = s_ts[kt];
      ri.klen = s_ts[kt + 1] - s_ts[kt];
      ri.lse = 0.f;
      ri.dsum = 0.f;
      ri.w[0] = ri.w[1] = 0xffffffffu;
      if (ri.row_ok) {
        const size_t li = (static_cast<size_t>(n) * p.H + h) * p.S + ri.q_row;
        ri.lse = p.lse[li];
        ri.dsum = p.dsum[li];
        if (ri.cls == 2) {
          const uint32_t* mrow = p.mask_bits + (static_cast<size_t>(n) * p.S + ri.q_row) * p.mask_words + (ri.kbase >> 5) + chunk_c;
          ri.w[0] = mrow[0];
          ri.w[1] = mrow[1];
        }
      }
    };
    auto next_active = [&](int t) -> int {
      ++t;
      while (t < n_t && s_cls[t] == 0) ++t;
      return t;
    };

    RowIn cur, nxt;
    int t = next_active(-1);
    if (t < n_t) fetch(t, cur);
    for (int it = 0; it < n_active; ++it) {
      const int st = it & 1;
      const int t_next = next_active(t);
      if (t_next < n_t) fetch(t_next, nxt);
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
      if (it >= 2) mbar_wait(&pds_empty[st], ((it - 2) >> 1) & 1);   // gradient MMAs of tile it-2 have consumed buffer st
      uint8_t* bP = sP + st * 32768;
      uint8_t* bDS = sDS + st * 32768;
      const float lse2 = cur.lse * 1.4426950408889634f;
      uint32_t rk1 = 0u, rk2 = 0u;
      if (DROP) {
        rk1 = drop_rowkey(p.drop.seed_lo, p.drop.seed_hi, n, h, cur.q_row);
        rk2 = drop_rowkey2(rk1);
      }
      // class 2 (mixed) tiles and partial query tiles (rows beyond the tile's length see nothing) take the masked variant
      if (cur.cls == 2 || __any_sync(0xffffffffu, !cur.row_ok)) {
        const uint32_t v = __funnelshift_r(cur.w[0], cur.w[1], cur.kbase & 31);
        const int nvalid = cur.klen - 32 * chunk_c;
        const uint32_t keep = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
        const uint32_t mw_c = (cur.row_ok && cur.cls == 2) ? (v & keep) : (cur.row_ok ? 0xffffffffu : 0u);
        bwd_tile<DKV, true, DROP>(tmem_S + lane_addr, tmem_dP + lane_addr, chunk_c, mw_c, r, bP, bDS, p.scale_log2, p.scale,
                                  lse2, cur.dsum, rk1, rk2, cur.kbase, p.drop.thresh, p.drop.inv_keep, eps_run);
      } else {
        bwd_tile<DKV, false, DROP>(tmem_S + lane_addr, tmem_dP + lane_addr, chunk_c, 0xffffffffu, r, bP, bDS, p.scale_log2,
                                   p.scale, lse2, cur.dsum, rk1, rk2, cur.kbase, p.drop.thresh, p.drop.inv_keep, eps_run);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[st]);
      cur = nxt;
      t = t_next;
    }

    // ---- epilogue: this thread owns output row `tile*128 + r`
    const int out_row = row_base + r;
    if (n_active > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    const bool ok = r < row_len;
    const long long grow = static_cast<long long>(n) * p.S + out_row;
    if (!DKV) {   // the four threads of a row exchange their eps partial sums
      float* epsx = reinterpret_cast<float*>(smem + L::kEps);
      epsx[chunk_c * 128 + r] = eps_run;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kBwdSoftmaxWarps) : "memory");
      eps_run = (epsx[r] + epsx[128 + r]) + (epsx[256 + r] + epsx[384 + r]);
    }
    // Work split between the first two warps of a quadrant (the other two are done): DKV — warp half 0 stores dV, half 1
    // stores dK; DQ — each stores two of the four 8-column groups of both rotation halves.
    if (chunk_c < 2) {
      const int a = DKV ? half : 0;
      // a == 0: dV (DKV, no rotation) or dQ (DQ, rotated);  a == 1: dK (rotated)
      const bool rot = DKV ? (a == 1) : true;
      const int col0 = (DKV ? (a == 0 ? p.v_col0 : p.k_col0) : p.q_col0) + h * 64;
      // cos / sin rows of this thread's token stay in shared memory (two swizzled 4 KB buffers per warp, aliasing the idle
      // streamed stages) and are read 8 columns at a time: holding them in registers next to the 64 accumulator values
      // does not fit the 112-register budget of a 576-thread CTA
      uint8_t* gcos = smem + L::kGather + (warp - 2) * 8192;
      uint8_t* gsin = gcos + 4096;
      if (rot) {   // warp-uniform
        const int pos = ok ? p.pos[grow] : 0;
        const int rd_row = lane >> 3, rd_j = lane & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + rd_row;
          const int src = __shfl_sync(0xffffffffu, pos, rr);
          const float4 vc = *reinterpret_cast<const float4*>(p.cos_tab + static_cast<long long>(src) * 32 + rd_j * 4);
          const float4 vs = *reinterpret_cast<const float4*>(p.sin_tab + static_cast<long long>(src) * 32 + rd_j * 4);
          *reinterpret_cast<float4*>(gcos + rr * 128 + ((rd_j ^ (rr & 7)) << 4)) = vc;
          *reinterpret_cast<float4*>(gsin + rr * 128 + ((rd_j ^ (rr & 7)) << 4)) = vs;
        }
        __syncwarp();
      }
      uint32_t x1[32], x2[32];
      if (n_active > 0) {
        tmem_ld32((a == 0 ? tmem_A0 : tmem_A1) + lane_addr, x1);
        tmem_ld32((a == 0 ? tmem_A0 : tmem_A1) + lane_addr + 32, x2);
        tmem_ld_wait();
        if (!DKV) {   // dQ = dS K - eps/8 * (P K)
          const float ce = eps_run * p.scale;
          uint32_t y[32];
          tmem_ld32(tmem_A1 + lane_addr, y);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) x1[j] = __float_as_uint(__uint_as_float(x1[j]) - ce * __uint_as_float(y[j]));
          tmem_ld32(tmem_A1 + lane_addr + 32, y);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) x2[j] = __float_as_uint(__uint_as_float(x2[j]) - ce * __uint_as_float(y[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) x1[j] = x2[j] = 0u;
      }
      if (DROP && DKV && a == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x1[j] = __float_as_uint(__uint_as_float(x1[j]) * p.drop.inv_keep);
          x2[j] = __float_as_uint(__uint_as_float(x2[j]) * p.drop.inv_keep);
        }
      }
      if (ok) {
        __nv_bfloat16* orow = p.dqkv + grow * p.ld + col0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (!DKV && (g >> 1) != half) continue;
          float o1[8], o2[8];
          if (rot) {
            float cs[8], sn[8];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int ch = ((2 * g + q) ^ (lane & 7)) << 4;
              const float4 vc = *reinterpret_cast<const float4*>(gcos + lane * 128 + ch);
              const float4 vs = *reinterpret_cast<const float4*>(gsin + lane * 128 + ch);
              cs[q * 4 + 0] = vc.x; cs[q * 4 + 1] = vc.y; cs[q * 4 + 2] = vc.z; cs[q * 4 + 3] = vc.w;
              sn[q * 4 + 0] = vs.x; sn[q * 4 + 1] = vs.y; sn[q * 4 + 2] = vs.z; sn[q * 4 + 3] = vs.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float d1 = __uint_as_float(x1[g * 8 + j]);
              const float d2 = __uint_as_float(x2[g * 8 + j]);
              o1[j] = d1 * cs[j] + d2 * sn[j];     // transpose of the forward rotation
              o2[j] = d2 * cs[j] - d1 * sn[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              o1[j] = __uint_as_float(x1[g * 8 + j]);
              o2[j] = __uint_as_float(x2[g * 8 + j]);
            }
          }
          uint4 v1, v2;
          v1.x = pack_bf16(o1[0], o1[1]); v1.y = pack_bf16(o1[2], o1[3]);
          v1.z = pack_bf16(o1[4], o1[5]); v1.w = pack_bf16(o1[6], o1[7]);
          v2.x = pack_bf16(o2[0], o2[1]); v2.y = pack_bf16(o2[2], o2[3]);
          v2.z = pack_bf16(o2[4], o2[5]); v2.w = pack_bf16(o2[6], o2[7]);
          *reinterpret_cast<uint4*>(orow + g * 8) = v1;
          *reinterpret_cast<uint4*>(orow + 32 + g * 8) = v2;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// D[n,h,s] = sum_d dO[n,s,h,d] * O[n,s,h,d]     (one warp per row; 8 lanes per head)
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO,
                                     long long ldo, long long lddo, float* __restrict__ dsum, int N, int S, int H) {
  const int lane = threadIdx.x & 31;
  const long long T = static_cast<long long>(N) * S;
  const int warps_per_block = blockDim.x >> 5;
  for (long long t = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); t < T;
       t += static_cast<long long>(gridDim.x) * warps_per_block) {
    const int n = static_cast<int>(t / S), s = static_cast<int>(t % S);
    for (int h0 = 0; h0 < H; h0 += 4) {
      const int hh = h0 + (lane >> 3);
      float acc = 0.f;
      if (hh < H) {
        const int c = hh * 64 + (lane & 7) * 8;
        const uint4 a = *reinterpret_cast<const uint4*>(o + t * ldo + c);
        const uint4 b = *reinterpret_cast<const uint4*>(dO + t * lddo + c);
        const uint32_t au[4] = {a.x, a.y, a.z, a.w}, bu[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16(au[j]), y = unpack_bf16(bu[j]);
          acc += x.x * y.x + x.y * y.y;
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (hh < H && (lane & 7) == 0) dsum[(static_cast<size_t>(n) * H + hh) * S + s] = acc;
    }
  }
}

template <bool DKV, bool DROP>
static int launch_bwd(const CUtensorMap& tmQKV, const CUtensorMap& tmDO, const AttnBwdParams& p, cudaStream_t s) {
  using L = BwdSmem<DKV>;
  auto kern = attn_bwd_kernel<DKV, DROP>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) {
      set_error("attn_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return -2;
    }
    attr_set = true;
  }
  dim3 grid(p.max_tiles, p.H, p.N);
  kern<<<grid, kBwdThreads, L::kTotal, s>>>(tmQKV, tmDO, p);
  return check_launch(DKV ? "attn_bwd_kernel<dkv>" : "attn_bwd_kernel<dq>");
}

}  // namespace ggpt

using namespace ggpt;

extern "C" {

int ggpt_attn_bwd(const void* qkv, long long ld_qkv, int q_col0, int k_col0, int v_col0, const void* out, long long ldo,
                  const void* dout, long long lddo, const float* lse, const uint32_t* mask_bits, const int* tile_start,
                  const int* n_tiles, const uint8_t* tile_cls, const uint8_t* iso_flags, const int* iso_list,
                  const int* iso_count, int run_general, float dropout_p, unsigned long long seed, const int* pos, const float* cos_tab, const float* sin_tab, float* dsum_scratch, void* dqkv,
                  long long ld_dqkv, int N, int S, int H, void* stream) {
  GGPT_REQUIRE(qkv && out && dout && lse && mask_bits && tile_start && n_tiles && tile_cls && pos && cos_tab && sin_tab && dsum_scratch && dqkv,
               "attn_bwd: null pointer");
  GGPT_REQUIRE(N > 0 && S > 0 && H > 0, "attn_bwd: empty problem");
  GGPT_REQUIRE(ld_qkv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && ld_dqkv % 8 == 0, "attn_bwd: ld must be multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long T = static_cast<long long>(N) * S;
  long long blocks = (T + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  attn_bwd_prep_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(out),
                                                                     static_cast<const __nv_bfloat16*>(dout), ldo, lddo,
                                                                     dsum_scratch, N, S, H);
  if (int rc = check_launch("attn_bwd_prep_kernel")) return rc;
  CUtensorMap tmQKV, tmDO;
  if (int rc = make_tmap_3d_bf16(&tmQKV, qkv, N, S, ld_qkv, static_cast<uint64_t>(S) * ld_qkv, ld_qkv, 128, 64)) return rc;
  if (int rc = make_tmap_3d_bf16(&tmDO, dout, N, S, static_cast<uint64_t>(H) * 64, static_cast<uint64_t>(S) * lddo, lddo, 128, 64))
    return rc;
  AttnBwdParams p{};
  p.N = N; p.S = S; p.H = H;
  p.max_tiles = ggpt_attn_max_tiles(S);
  p.mask_words = ggpt_attn_mask_words(S);
  p.mask_bits = mask_bits; p.tile_start = tile_start; p.n_tiles = n_tiles; p.tile_cls = tile_cls; p.lse = lse;
  p.dsum = dsum_scratch; p.iso_flags = iso_flags;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv); p.ld = ld_dqkv;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.pos = pos; p.cos_tab = cos_tab; p.sin_tab = sin_tab;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  GGPT_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "attn_bwd: dropout_p must be in [0,1)");
  p.drop = make_drop_params(dropout_p, seed);
  GGPT_REQUIRE(run_general || iso_flags, "attn_bwd: run_general == 0 needs the isolated-tile work list");
  if (run_general) {
    if (p.drop.thresh != 0u) {
      if (int rc = launch_bwd<true, true>(tmQKV, tmDO, p, s)) return rc;
      if (int rc = launch_bwd<false, true>(tmQKV, tmDO, p, s)) return rc;
    } else {
      if (int rc = launch_bwd<true, false>(tmQKV, tmDO, p, s)) return rc;
      if (int rc = launch_bwd<false, false>(tmQKV, tmDO, p, s)) return rc;
    }
  }
  if (iso_flags == nullptr) return 0;
  GGPT_REQUIRE(iso_list && iso_count, "attn_bwd: iso_flags given without iso_list / iso_count");
  DiagParams d{};
  d.N = N; d.S = S; d.H = H; d.max_tiles = p.max_tiles; d.mask_words = p.mask_words;
  d.mask_bits = mask_bits; d.tile_start = tile_start; d.tile_cls = tile_cls; d.iso_list = iso_list; d.iso_count = iso_count;
  d.q_col0 = q_col0; d.k_col0 = k_col0; d.v_col0 = v_col0; d.scale = p.scale; d.scale_log2 = p.scale_log2;
  d.lse_in = lse; d.dsum = dsum_scratch; d.dqkv = p.dqkv; d.ld_dqkv = ld_dqkv;
  d.pos = pos; d.cos_tab = cos_tab; d.sin_tab = sin_tab; d.drop = p.drop;
  return attn_diag_bwd_launch(tmQKV, tmDO, d, s);
}

}  // extern "C"
