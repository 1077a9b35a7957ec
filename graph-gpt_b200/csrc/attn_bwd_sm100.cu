// Attention backward on tcgen05 (sm_100a): dQ, dK, dV of softmax(Q K^T / 8 + mask) V with recomputation of the
// probabilities from the saved log-sum-exp (no [N,H,S,S] tensor is ever stored).  autograd of HF:199-221.
//
// ONE pass: one CTA per (key tile j, head, sequence) keeps K_j, V_j in shared memory and the dK_j / dV_j accumulators in
// TMEM, and streams the query tiles i that see it:
//     S = Q_i K_j^T,  dP = dO_i V_j^T                         (2 MMAs, 128 x 128 x 64, accumulators in TMEM)
//     P = exp(S/8 - lse_i),  dS = P * (dP - D_i) / 8          (16 warps, four threads per query row)
//     dV_j += P^T dO_i,  dK_j += dS^T Q_i,  dQ_i(j) = dS K_j  (3 MMAs; P / dS through swizzled shared memory, read K-major
//                                                              or, via the MN-major descriptor, transposed)
// Every score tile is computed once (5 MMAs and one exponential per element instead of the 8 + 2 of a dK/dV pass followed
// by a dQ pass).  The partial dQ_i(j) of the different key tiles are summed in an fp32 accumulator [N,S,H*64] in global
// memory with TMA reductions (cp.reduce.async.bulk.tensor ... add: the tile leaves TMEM through registers into a swizzled
// staging buffer and the TMA engine adds it at L2 — no per-thread atomics); attn_dq_finish_kernel then applies the
// inverse RoPE, writes bf16 dQ and re-zeroes the accumulator.  dK / dV leave through the epilogue (dK un-rotated).
// The fp32 summation order of dQ across key tiles is the only non-determinism.
//
// Pipeline: S / dP are drained from TMEM into registers at the start of the softmax-gradient phase and handed back at
// once, so the score MMAs of tile i+1 run while the warps compute tile i; P / dS stay in registers until the gradient
// MMAs of tile i-1 have retired and are then stored, so one 32 KB buffer each suffices; Q_i / dO_i arrive through a
// 3-stage TMA ring.
//
// D = rowsum(dO * O) must equal sum_k P dP to well below bf16 precision: its error eps_q shifts every dS of the row and
// dQ by -eps_q/8 * (sum_k P K) — visible when keys share a large common component (rows of identical <mask> tokens).
// The forward pass therefore also stores the bf16 rounding residual of its output (out_lo = bf16(O - bf16(O))), and
// D = dO . (O + O_lo) is exact to ~2^-17 (the isolated-tile kernels correct eps in-kernel instead, attn_diag_sm100.cu).
#include "common.cuh"
#include "attn_diag.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

// TMA producer, MMA issuer, 16 softmax-gradient warps: four per TMEM lane quadrant, i.e. four threads per query row with
// one 32-column chunk each.  The softmax-gradient phase is a chain of dependent fixed-latency instructions (FFMA -> MUFU ->
// FMUL -> F2FP -> STS); the remedy is resident warps, not fewer instructions.
constexpr int kBwdSoftmaxWarps = 16;
constexpr int kBwdThreads = 64 + 32 * kBwdSoftmaxWarps;
constexpr int kBwdStages = 3;

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

// smem (one CTA per SM): K_j | V_j | 3 stages x (Q_i | dO_i) | P | dS | dQ staging | tile plan | barriers.
// The epilogue's RoPE-table gather buffers alias the streamed stages (idle once the last MMA has retired).
struct BwdSmem {
  static constexpr int kFixed = 0;
  static constexpr int kStream = 32768;
  static constexpr int kP = kStream + kBwdStages * 32768;
  static constexpr int kDS = kP + 32768;
  static constexpr int kDQ = kDS + 32768;          // 16 warps x one box of [32 rows x 16 fp32] (SWIZZLE_64B)
  static constexpr int kGather = kStream;          // 8 warps x 8 KB, epilogue only
  static constexpr int kPlan = kDQ + 32768;        // int ts[<= 257] | uint8 cls[<= 256]: this sequence's tile plan
  static constexpr int kBars = kPlan + 1536;
  static constexpr int kTotal = kBars + 256;
};
static_assert(BwdSmem::kTotal <= 232448, "attention backward exceeds the 227 KB of dynamic shared memory");

// One 32-column chunk of a thread's query row: P = exp(S/8 - lse), dS = P (dP - D) / 8, packed to bf16 IN PLACE
// (s[0..15] <- P, dp[0..15] <- dS).  MASKED / DROP are compile-time so that full tiles and dropout-free runs carry no mask
// or RNG instructions (predicated-off instructions still take issue slots).
template <bool MASKED, bool DROP>
__device__ __forceinline__ void bwd_chunk(uint32_t (&s)[32], uint32_t (&dp)[32], int c, uint32_t w, float scale_log2,
                                          float scale, float lse2, float dsum, uint32_t rk1, uint32_t rk2, int kbase,
                                          uint32_t thresh, float inv_keep) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float pv[8], dv[8];
    uint32_t bits = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float e = fast_exp2(fmaf(__uint_as_float(s[g * 8 + j]), scale_log2, -lse2));
      if (MASKED) e = ((w >> (g * 8 + j)) & 1u) ? e : 0.f;
      float dpe = __uint_as_float(dp[g * 8 + j]);
      bool keep = true;
      if (DROP) {
        if ((j & 1) == 0) bits = drop_bits(rk1, rk2, kbase + c * 32 + g * 8 + j);
        keep = (j & 1) ? drop_keep_odd(bits, thresh) : drop_keep_even(bits, thresh);
        dpe = keep ? dpe * inv_keep : 0.f;
      }
      dv[j] = e * (dpe - dsum) * scale;
      pv[j] = (DROP && !keep) ? 0.f : e;       // the stored P feeds dV = (P o keep)^T dO / (1-p)
    }
    s[g * 4 + 0] = pack_bf16(pv[0], pv[1]); s[g * 4 + 1] = pack_bf16(pv[2], pv[3]);
    s[g * 4 + 2] = pack_bf16(pv[4], pv[5]); s[g * 4 + 3] = pack_bf16(pv[6], pv[7]);
    dp[g * 4 + 0] = pack_bf16(dv[0], dv[1]); dp[g * 4 + 1] = pack_bf16(dv[2], dv[3]);
    dp[g * 4 + 2] = pack_bf16(dv[4], dv[5]); dp[g * 4 + 3] = pack_bf16(dv[6], dv[7]);
  }
}

#ifdef GGPT_ATTN_TRACE      // profiling aid: clock64 stamps of softmax-gradient warp 2 of 64 CTAs (tools/attn_trace.py bwd)
__device__ long long g_bwd_trace[64 * 12 * 8];
#define BWD_TRACE(row, slot)                                                                                     \
  do {                                                                                                           \
    if (trace_on && lane == 0 && (row) < 12) g_bwd_trace[(trace_cta * 12 + (row)) * 8 + (slot)] = clock64();      \
  } while (0)
int bwd_trace_read(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_bwd_trace, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
#else
#define BWD_TRACE(row, slot) do { } while (0)
#endif

template <bool DROP>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const AttnBwdParams p) {
  using L = BwdSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + L::kFixed;
  uint8_t* sV = smem + L::kFixed + 16384;
  uint8_t* sStr = smem + L::kStream;            // stage s: [s*32768] = Q_i, [+16384] = dO_i
  uint8_t* sP = smem + L::kP;
  uint8_t* sDS = smem + L::kDS;
  uint8_t* sDQ = smem + L::kDQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint64_t* fix_full = bars + 0;
  uint64_t* str_full = bars + 1;    // [3]
  uint64_t* str_empty = bars + 4;   // [3]
  uint64_t* sdp_full = bars + 7;    // S and dP accumulators of the next tile ready
  uint64_t* sdp_empty = bars + 8;   // ... and drained into registers                                  (count 16)
  uint64_t* pds_full = bars + 9;    // P / dS of tile i in smem, dQ of tile i-1 drained from TMEM         (count 16)
  uint64_t* pds_empty = bars + 10;  // gradient MMAs of tile i retired: P / dS reusable, dQ_i(j) in TMEM
  uint64_t* acc_full = bars + 11;   // final accumulators ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
#ifdef GGPT_ATTN_TRACE
  const int trace_cta = (blockIdx.z % 8) * 8 + blockIdx.x % 8;     // 64 traced CTAs: head 5, sequences 8..15
  const bool trace_on = (warp == 2) && blockIdx.y == 5 && blockIdx.z >= 8 && blockIdx.z < 16;
  BWD_TRACE(8, 0);
#endif
  const int n_t = p.n_tiles[n];
  if (tile >= n_t) return;                      // uniform for the whole CTA, before any barrier / TMEM state
  if (p.iso_flags != nullptr && p.iso_flags[n * p.max_tiles + tile]) return;   // done by attn_diag_bwd_kernel
  const int* ts = p.tile_start + static_cast<size_t>(n) * (p.max_tiles + 1);
  const int row_base = ts[tile], row_len = ts[tile + 1] - row_base;   // rows of the key tile
  const uint8_t* cls_n = p.tile_cls + static_cast<size_t>(n) * p.max_tiles * p.max_tiles;
  // the tile plan of this sequence (row offsets, class of (query tile t, this key tile)) is copied to shared memory once:
  // every role consults it for every tile, and a global load on that path costs an L2 round trip per tile
  int* s_ts = reinterpret_cast<int*>(smem + L::kPlan);
  uint8_t* s_cls = reinterpret_cast<uint8_t*>(smem + L::kPlan + 1040);     // S <= 16384 => at most 256 row tiles
  for (int t = threadIdx.x; t <= n_t; t += blockDim.x) {
    s_ts[t] = ts[t];
    if (t < n_t) s_cls[t] = cls_n[t * p.max_tiles + tile];
  }

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("ggpt attn_bwd: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    mbar_init(fix_full, 1);
#pragma unroll
    for (int i = 0; i < kBwdStages; ++i) {
      mbar_init(&str_full[i], 1);
      mbar_init(&str_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, kBwdSoftmaxWarps);
    mbar_init(pds_full, kBwdSoftmaxWarps);
    mbar_init(pds_empty, 1);
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
  const uint32_t tmem_dV = tmem_base + 256;    // 64 columns
  const uint32_t tmem_dK = tmem_base + 320;    // 64 columns
  const uint32_t tmem_dQ = tmem_base + 384;    // 64 columns: dS K_j of the current query tile (not accumulated)

  int n_active = 0;
  for (int t = 0; t < n_t; ++t) n_active += (s_cls[t] != 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && n_active > 0) {
      mbar_expect_tx(fix_full, 32768);
      tma_load_3d(sK, &tmQKV, fix_full, p.k_col0 + h * 64, row_base, n);
      tma_load_3d(sV, &tmQKV, fix_full, p.v_col0 + h * 64, row_base, n);
      int st = 0;
      uint32_t ph = 1;
      for (int t = 0; t < n_t; ++t) {
        if (s_cls[t] == 0) continue;
        mbar_wait(&str_empty[st], ph);
        mbar_expect_tx(&str_full[st], 32768);
        tma_load_3d(sStr + st * 32768, &tmQKV, &str_full[st], p.q_col0 + h * 64, s_ts[t], n);
        tma_load_3d(sStr + st * 32768 + 16384, &tmDO, &str_full[st], h * 64, s_ts[t], n);
        if (++st == kBwdStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (n_active > 0) {                      // whole warp, one elected lane issues (see tc_mma_bf16_e)
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);     // S, dP : K-major x K-major
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, true, true);        // P^T dO, dS^T Q : MN x MN
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, false, true);       // dS K : K-major x MN-major
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      const uint32_t aP = smem_u32(sP), aDS = smem_u32(sDS);
      int st_s = 0;                 // stage / phase of the next score issue
      uint32_t ph_s = 0;
      auto issue_scores = [&]() {
        mbar_wait(&str_full[st_s], ph_s);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sStr + st_s * 32768), aDO = aQ + 16384;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_bf16_e(tmem_S, umma_desc_sw128(aQ + kk * 32, 16, 1024), umma_desc_sw128(aK + kk * 32, 16, 1024),
                      idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_bf16_e(tmem_dP, umma_desc_sw128(aDO + kk * 32, 16, 1024), umma_desc_sw128(aV + kk * 32, 16, 1024),
                      idesc_s, kk != 0);
        tc_commit_e(sdp_full);
        if (++st_s == kBwdStages) {
          st_s = 0;
          ph_s ^= 1;
        }
      };
      mbar_wait(fix_full, 0);
      issue_scores();
      int st = 0;
      for (int it = 0; it < n_active; ++it) {
        // the scores of the NEXT tile are issued as soon as this tile's S / dP sit in the softmax-gradient warps'
        // registers, i.e. they run WHILE those warps compute P / dS of this tile
        if (it + 1 < n_active) {
          mbar_wait(sdp_empty, it & 1);
          tc_fence_after();
          issue_scores();
        }
        mbar_wait(pds_full, it & 1);     // P / dS of tile it in smem; dQ of tile it-1 drained
        tc_fence_after();
        const uint32_t aQ = smem_u32(sStr + st * 32768), aDO = aQ + 16384;
        // dV += P^T dO_i ; dK += dS^T Q_i      (A = P / dS read MN-major: M = keys, K = query rows)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          tc_mma_bf16_e(tmem_dV, umma_desc_sw128(aP + kk * 2048, 16384, 1024), umma_desc_sw128(aDO + kk * 2048, 8192, 1024),
                      idesc_t, (it | kk) != 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          tc_mma_bf16_e(tmem_dK, umma_desc_sw128(aDS + kk * 2048, 16384, 1024), umma_desc_sw128(aQ + kk * 2048, 8192, 1024),
                      idesc_t, (it | kk) != 0);
        // dQ_i(j) = dS K_j                       (A = dS K-major: M = query rows, K = keys; B = K_j MN-major)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          tc_mma_bf16_e(tmem_dQ, umma_desc_sw128(aDS + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                      umma_desc_sw128(aK + kk * 2048, 8192, 1024), idesc_q, kk != 0);
        tc_commit_e(pds_empty);
        tc_commit_e(&str_empty[st]);
        if (it + 1 == n_active) tc_commit_e(acc_full);
        if (++st == kBwdStages) st = 0;
      }
    }
  } else {
    // ===================== softmax-gradient warps: four threads per query row (32 key columns each) =====================
    const int quad = warp & 3;
    const int chunk_c = (warp - 2) >> 2;      // which 32-column chunk of the score tile this thread owns
    const int half = chunk_c & 1;             // epilogue role (warps with chunk_c < 2 write dV / dK)
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    // dQ staging: every warp owns a [32 rows x 16 fp32] box (64-byte rows, SWIZZLE_64B) and issues its own TMA reduction,
    // so no warp ever waits for another one on this path
    uint8_t* dq_box = sDQ + (warp - 2) * 2048;
    uint8_t* dq_stage = dq_box + lane * 64;
    const int dq_sw = (lane >> 1) & 3;                       // 64-byte swizzle: 16-byte chunk index ^= (row >> 1) & 3

    // per-(tile, row) inputs; fetched one tile ahead so that their global-memory latency is off the critical path
    // (only RAW loaded values are kept: an in-order warp stalls at the first instruction that consumes a load, so every
    // consumer — the log2(e) scaling, the funnel shifts of the mask words — sits at the point of use, one tile later)
    struct RowIn {
      float lse, dsum;
      uint32_t w0, w1;                        // the two mask words that cover this thread's 32 columns
    };
    auto fetch = [&](int t, RowIn& ri) {      // tile plan entries come from shared memory at the point of use
      const int q0 = s_ts[t];
      const bool row_ok = r < s_ts[t + 1] - q0;
      ri.lse = 0.f;
      ri.dsum = 0.f;
      ri.w0 = ri.w1 = 0u;
      if (row_ok) {
        const size_t li = (static_cast<size_t>(n) * p.H + h) * p.S + q0 + r;
        ri.lse = p.lse[li];
        ri.dsum = p.dsum[li];
        if (s_cls[t] == 2) {
          const uint32_t* mrow = p.mask_bits + (static_cast<size_t>(n) * p.S + q0 + r) * p.mask_words + (row_base >> 5) + chunk_c;
          ri.w0 = mrow[0];
          ri.w1 = mrow[1];
        }
      }
    };
    auto next_active = [&](int t) -> int {
      ++t;
      while (t < n_t && s_cls[t] == 0) ++t;
      return t;
    };
    // dQ_i(j) (this thread's 16 columns, already in registers) -> swizzled staging -> TMA reduce-add into dq_acc
    auto dq_reduce = [&](const uint32_t (&dq)[16], int q0) {
      if (lane == 0) tma_store_wait_read<0>();          // this warp's previous reduction has read the staging box
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(dq_stage + ((i ^ dq_sw) << 4)) = make_uint4(dq[i * 4 + 0], dq[i * 4 + 1], dq[i * 4 + 2], dq[i * 4 + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_3d(&tmDQ, dq_box, h * 64 + chunk_c * 16, q0 + quad * 32, n);
        tma_store_commit();
      }
    };

    RowIn cur, nxt;
    int t = next_active(-1);
    if (t < n_t) fetch(t, cur);
    int prev_q0 = 0;
    BWD_TRACE(8, 1);
    for (int it = 0; it < n_active; ++it) {
      BWD_TRACE(it, 0);
      const int t_next = next_active(t);
      if (t_next < n_t) fetch(t_next, nxt);
      mbar_wait(sdp_full, it & 1);
      BWD_TRACE(it, 1);
      tc_fence_after();
      uint32_t s0[32], d0[32];
      tmem_ld32(tmem_S + lane_addr + chunk_c * 32, s0);
      tmem_ld32(tmem_dP + lane_addr + chunk_c * 32, d0);
      tmem_ld_wait();
      tc_fence_before();              // scores are in registers: hand S / dP back before the arithmetic, not after it
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_empty);
      BWD_TRACE(it, 2);
      const float lse2 = cur.lse * 1.4426950408889634f;
      const int cur_q0 = s_ts[t];
      uint32_t rk1 = 0u, rk2 = 0u;
      if (DROP) {
        rk1 = drop_rowkey(p.drop.seed_lo, p.drop.seed_hi, n, h, cur_q0 + r);
        rk2 = drop_rowkey2(rk1);
      }
      // class 2 (mixed) tiles — which include every partial tile — take the masked variant (rows beyond the query tile's
      // length were fetched with all-zero mask words)
      if (s_cls[t] == 2) {
        const uint32_t v = __funnelshift_r(cur.w0, cur.w1, row_base & 31);
        const int nvalid = row_len - 32 * chunk_c;
        const uint32_t keep = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
        bwd_chunk<true, DROP>(s0, d0, chunk_c, v & keep, p.scale_log2, p.scale, lse2, cur.dsum, rk1, rk2, row_base, p.drop.thresh,
                              p.drop.inv_keep);
      } else {
        bwd_chunk<false, DROP>(s0, d0, chunk_c, 0xffffffffu, p.scale_log2, p.scale, lse2, cur.dsum, rk1, rk2, row_base,
                               p.drop.thresh, p.drop.inv_keep);
      }
      BWD_TRACE(it, 3);
      uint32_t dq[16];
      if (it > 0) {
        mbar_wait(pds_empty, (it - 1) & 1);   // gradient MMAs of tile it-1 retired: P / dS free, dQ_{it-1}(j) complete
        tc_fence_after();
        tmem_ld16(tmem_dQ + lane_addr + chunk_c * 16, dq);
        tmem_ld_wait();
      }
      BWD_TRACE(it, 4);
      if (it > 0) {                              // this warp's previous dQ reduction has read the staging box
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
      {
        uint8_t* pbase = sP + (chunk_c >> 1) * 16384 + r * 128;
        uint8_t* dbase = sDS + (chunk_c >> 1) * 16384 + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = ((chunk_c & 1) * 4 + g) ^ (r & 7);
          *reinterpret_cast<uint4*>(pbase + chunk * 16) = make_uint4(s0[g * 4 + 0], s0[g * 4 + 1], s0[g * 4 + 2], s0[g * 4 + 3]);
          *reinterpret_cast<uint4*>(dbase + chunk * 16) = make_uint4(d0[g * 4 + 0], d0[g * 4 + 1], d0[g * 4 + 2], d0[g * 4 + 3]);
        }
        if (it > 0) {                            // dQ_{it-1}(j) rides on the same proxy fence as P / dS
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(dq_stage + ((i ^ dq_sw) << 4)) = make_uint4(dq[i * 4 + 0], dq[i * 4 + 1], dq[i * 4 + 2], dq[i * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(pds_full);
        if (it > 0) {
#ifndef GGPT_BWD_NO_REDUCE      // (timing experiment of tools/attn_trace.py: how much of the tile period is the TMA reduction?)
          tma_reduce_add_3d(&tmDQ, dq_box, h * 64 + chunk_c * 16, prev_q0 + quad * 32, n);
#endif
          tma_store_commit();
        }
      }
      BWD_TRACE(it, 5);
      BWD_TRACE(it, 6);
      prev_q0 = cur_q0;
      cur = nxt;
      t = t_next;
    }

    // ---- last dQ tile, then the epilogue: this thread owns output row `row_base + r`
    BWD_TRACE(8, 2);
    if (n_active > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      uint32_t dq[16];
      tmem_ld16(tmem_dQ + lane_addr + chunk_c * 16, dq);
      tmem_ld_wait();
      dq_reduce(dq, prev_q0);
    }
    const int out_row = row_base + r;
    const bool ok = r < row_len;
    const long long grow = static_cast<long long>(n) * p.S + out_row;
    // the first two warps of a quadrant write the outputs (the other two are done): half 0 stores dV, half 1 stores dK
    if (chunk_c < 2) {
      const int a = half;                       // a == 0: dV (no rotation);  a == 1: dK (un-rotated)
      const bool rot = (a == 1);
      const int col0 = (a == 0 ? p.v_col0 : p.k_col0) + h * 64;
      // cos / sin rows of this thread's token stay in shared memory (two swizzled 4 KB buffers per warp, aliasing the idle
      // streamed stages) and are read 8 columns at a time: holding them in registers next to the 64 accumulator values
      // does not fit the register budget of a 576-thread CTA
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
        tmem_ld32((a == 0 ? tmem_dV : tmem_dK) + lane_addr, x1);
        tmem_ld32((a == 0 ? tmem_dV : tmem_dK) + lane_addr + 32, x2);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) x1[j] = x2[j] = 0u;
      }
      if (DROP && a == 0) {
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
    if (lane == 0) tma_store_wait_all();        // the reductions have left shared memory and are globally performed
    BWD_TRACE(8, 3);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// D[n,h,s] = sum_d dO[n,s,h,d] * (O + O_lo)[n,s,h,d]     (one warp per row; 8 lanes per head; O_lo may be NULL)
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ olo,
                                     const __nv_bfloat16* __restrict__ dO, long long ldo, long long lddo,
                                     float* __restrict__ dsum, int N, int S, int H) {
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
        uint4 l = make_uint4(0u, 0u, 0u, 0u);
        if (olo != nullptr) l = *reinterpret_cast<const uint4*>(olo + t * ldo + c);
        const uint32_t au[4] = {a.x, a.y, a.z, a.w}, bu[4] = {b.x, b.y, b.z, b.w}, lu[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16(au[j]), y = unpack_bf16(bu[j]), z = unpack_bf16(lu[j]);
          acc += (x.x + z.x) * y.x + (x.y + z.y) * y.y;
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (hh < H && (lane & 7) == 0) dsum[(static_cast<size_t>(n) * H + hh) * S + s] = acc;
    }
  }
}

// dQ = inverse RoPE of the fp32 accumulator -> bf16 q columns of dqkv; the accumulator is re-zeroed for the next call.
// One CTA per (row tile, sequence) — tiles of the isolated-diagonal kernel (which writes its dQ itself) are skipped.
// A work element is (row, head, 4 rotation pairs): two 16-byte accumulator loads (columns j..j+3 and 32+j..35+j), two
// 8-byte bf16 stores and two 16-byte zero stores, all independent of the other elements (the loop is unrolled so that
// several elements' loads are in flight per thread: the kernel is a pure HBM stream, 12 B in + 6 B out per element).
__global__ void __launch_bounds__(256)
attn_dq_finish_kernel(float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, long long ld, int q_col0,
                      const int* __restrict__ tile_start, const int* __restrict__ n_tiles,
                      const uint8_t* __restrict__ iso_flags, int max_tiles, const int* __restrict__ pos,
                      const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, int S, int H) {
  const int tile = blockIdx.x, n = blockIdx.y;
  if (tile >= n_tiles[n]) return;
  if (iso_flags != nullptr && iso_flags[n * max_tiles + tile]) return;
  const int* ts = tile_start + static_cast<size_t>(n) * (max_tiles + 1);
  const int r0 = ts[tile], rows = ts[tile + 1] - r0;
  const int per_row = H * 8;
  const int total = rows * per_row;
  const long long row_base = static_cast<long long>(n) * S + r0;
  for (int e0 = threadIdx.x; e0 < total; e0 += 4 * 256) {
    float4 d1[4], d2[4];
    int psv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {                        // all loads of four elements first
      const int e = e0 + u * 256;
      if (e < total) {
        const int rr = e / per_row, rem = e - rr * per_row;
        const long long row = row_base + rr;
        const float4* a1 = reinterpret_cast<const float4*>(dq_acc + row * (static_cast<long long>(H) * 64) + (rem >> 3) * 64 + (rem & 7) * 4);
        d1[u] = a1[0];
        d2[u] = a1[8];
        psv[u] = pos[row];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      if (e < total) {
        const int rr = e / per_row, rem = e - rr * per_row;
        const int hh = rem >> 3, j = (rem & 7) * 4;
        const long long row = row_base + rr;
        const float4 c = *reinterpret_cast<const float4*>(cos_tab + static_cast<long long>(psv[u]) * 32 + j);
        const float4 sn = *reinterpret_cast<const float4*>(sin_tab + static_cast<long long>(psv[u]) * 32 + j);
        float4* a1 = reinterpret_cast<float4*>(dq_acc + row * (static_cast<long long>(H) * 64) + hh * 64 + j);
        a1[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        a1[8] = make_float4(0.f, 0.f, 0.f, 0.f);
        __nv_bfloat16* o = dqkv + row * ld + q_col0 + hh * 64 + j;
        uint2 o1, o2;                                    // transpose of the forward rotation
        o1.x = pack_bf16(d1[u].x * c.x + d2[u].x * sn.x, d1[u].y * c.y + d2[u].y * sn.y);
        o1.y = pack_bf16(d1[u].z * c.z + d2[u].z * sn.z, d1[u].w * c.w + d2[u].w * sn.w);
        o2.x = pack_bf16(d2[u].x * c.x - d1[u].x * sn.x, d2[u].y * c.y - d1[u].y * sn.y);
        o2.y = pack_bf16(d2[u].z * c.z - d1[u].z * sn.z, d2[u].w * c.w - d1[u].w * sn.w);
        *reinterpret_cast<uint2*>(o) = o1;
        *reinterpret_cast<uint2*>(o + 32) = o2;
      }
    }
  }
}

template <bool DROP>
static int launch_bwd(const CUtensorMap& tmQKV, const CUtensorMap& tmDO, const CUtensorMap& tmDQ, const AttnBwdParams& p,
                      cudaStream_t s) {
  auto kern = attn_bwd_kernel<DROP>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kTotal);
    if (e != cudaSuccess) {
      set_error("attn_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return -2;
    }
    attr_set = true;
  }
  dim3 grid(p.max_tiles, p.H, p.N);
  kern<<<grid, kBwdThreads, BwdSmem::kTotal, s>>>(tmQKV, tmDO, tmDQ, p);
  return check_launch("attn_bwd_kernel");
}

}  // namespace ggpt

using namespace ggpt;

extern "C" {

int ggpt_attn_bwd(const void* qkv, long long ld_qkv, int q_col0, int k_col0, int v_col0, const void* out, const void* out_lo,
                  long long ldo, const void* dout, long long lddo, const float* lse, const uint32_t* mask_bits,
                  const int* tile_start, const int* n_tiles, const uint8_t* tile_cls, const uint8_t* iso_flags,
                  const int* iso_list, const int* iso_count, int run_general, float dropout_p, unsigned long long seed,
                  const int* pos, const float* cos_tab, const float* sin_tab, float* dsum_scratch, float* dq_acc, void* dqkv,
                  long long ld_dqkv, int N, int S, int H, void* stream) {
  GGPT_REQUIRE(qkv && out && dout && lse && mask_bits && tile_start && n_tiles && tile_cls && pos && cos_tab && sin_tab && dsum_scratch && dqkv,
               "attn_bwd: null pointer");
  GGPT_REQUIRE(N > 0 && S > 0 && H > 0, "attn_bwd: empty problem");
  GGPT_REQUIRE(ld_qkv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && ld_dqkv % 8 == 0, "attn_bwd: ld must be multiples of 8");
  GGPT_REQUIRE(!run_general || dq_acc,
               "attn_bwd: the general (tile-loop) path needs the zero-filled fp32 dQ accumulator dq_acc [N*S, H*64]");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long T = static_cast<long long>(N) * S;
  long long blocks = (T + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  attn_bwd_prep_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(out),
                                                                     static_cast<const __nv_bfloat16*>(out_lo),
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
    CUtensorMap tmDQ;
    const uint64_t dq_ld = static_cast<uint64_t>(H) * 64;
    if (int rc = make_tmap_3d_f32(&tmDQ, dq_acc, N, S, dq_ld, static_cast<uint64_t>(S) * dq_ld, dq_ld, 32, 16)) return rc;
    if (p.drop.thresh != 0u) {
      if (int rc = launch_bwd<true>(tmQKV, tmDO, tmDQ, p, s)) return rc;
    } else {
      if (int rc = launch_bwd<false>(tmQKV, tmDO, tmDQ, p, s)) return rc;
    }
    attn_dq_finish_kernel<<<dim3(p.max_tiles, N), 256, 0, s>>>(dq_acc, p.dqkv, ld_dqkv, q_col0, tile_start, n_tiles, iso_flags,
                                                               p.max_tiles, pos, cos_tab, sin_tab, S, H);
    if (int rc = check_launch("attn_dq_finish_kernel")) return rc;
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
