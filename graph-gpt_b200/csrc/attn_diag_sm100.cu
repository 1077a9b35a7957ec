// Persistent attention kernels for ISOLATED DIAGONAL tiles — the packed-sequence fast path.
//
// With the variable row tiles of ggpt_attn_mask_build, a packed batch (block-diagonal mask over ~24-row
// Eulerian-path segments) decomposes into row tiles whose only visible key tile is the tile itself.  For such a
// tile there is no online-softmax loop at all: one QK^T, one softmax, one PV — and in backward one score pass
// produces dQ, dK and dV together.  What is left is latency (TMA -> MMA -> softmax -> MMA -> store), so instead of
// one short-lived CTA per tile these kernels are persistent (one CTA per SM) and software-pipelined over a device-side
// work list of (sequence, tile) x head items: the TMA producer runs stages ahead, the MMA warp issues the scores of
// item i+1 while the softmax warps work on item i, and the epilogue of item i-1 is interleaved after the softmax
// of item i.  Two TMEM slots hold the accumulators of the two items in flight; in backward the gradient accumulators
// alias the (by then consumed) S / dP columns of the same slot so that two slots fit in 512 columns.
//
// Arithmetic is identical to attn_fwd_sm100.cu / attn_bwd_sm100.cu (same bf16 rounding points), see there for the
// reference citations (HF:199-221, HF:146-168).
#include "common.cuh"
#include "attn_diag.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

struct DiagItem {
  int n, h, r0, len, cls;
};
// Work item idx = (isolated tile idx / H, head idx % H).  The tile's {sequence, first row, rows, class} come from ONE
// 16-byte descriptor load (written by attn_iso_kernel) — no dependent chain of list -> tile_start -> class loads.
__device__ __forceinline__ DiagItem diag_item(const DiagParams& p, int idx) {
  const int4 d = __ldg(reinterpret_cast<const int4*>(p.iso_list) + idx / p.H);
  DiagItem it;
  it.n = d.x; it.r0 = d.y; it.len = d.z; it.cls = d.w;
  it.h = idx % p.H;
  return it;
}
// raw 160 mask bits of one query row covering key columns [k0, k0 + 128) (any alignment); see diag_mask_finish
__device__ __forceinline__ void diag_mask_raw(const uint32_t* __restrict__ mrow, int k0, uint32_t (&w)[5]) {
  const int w0 = k0 >> 5;
#pragma unroll
  for (int i = 0; i < 5; ++i) w[i] = __ldg(mrow + w0 + i);
}
__device__ __forceinline__ void diag_mask_finish(const uint32_t (&w)[5], int k0, int klen, uint32_t (&mw)[4]) {
  const int sh = k0 & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t v = __funnelshift_r(w[i], w[i + 1], sh);
    const int nvalid = klen - 32 * i;
    const uint32_t keep = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
    mw[i] = v & keep;
  }
}

// =====================================================================================================
// forward
// =====================================================================================================
// warps: 0 = TMA producer, 1 = MMA issuer, 2..5 = softmax group 0, 6..9 = softmax group 1.  Item i lives in TMEM
// slot / P buffer / softmax group (i & 1): while one group runs the softmax of item i the other finishes item i-1
// (waits for its PV, normalises, stores), so eight warps keep the SM busy and no row statistic crosses a warp.
constexpr int kDFThreads = 320;
constexpr int kDFStages = 3;
constexpr int kDFStageBytes = 49152;                  // Q | K | V
constexpr int kDFSmemP = kDFStages * kDFStageBytes;   // 2 x 32 KB P buffers
constexpr int kDFSmemBars = kDFSmemP + 65536;
constexpr int kDFSmem = kDFSmemBars + 256;

__global__ void __launch_bounds__(kDFThreads, 1)
attn_diag_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const DiagParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDFSmemBars);
  uint64_t* full = bars;            // [3]
  uint64_t* empty = bars + 3;       // [3]
  uint64_t* s_full = bars + 6;      // [2]
  uint64_t* p_full = bars + 8;      // [2] count 4
  uint64_t* o_full = bars + 10;     // [2]
  uint64_t* slot_free = bars + 12;  // [2] count 4: group finished reading S / O of its slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp: provably uniform
  const int total = p.iso_count[0] * p.H;
  if (static_cast<int>(blockIdx.x) >= total) return;   // uniform; nothing allocated yet
  const int n_items = (total - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&slot_free[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;   // slot b: S at b*256, O at b*256 + 128

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_items; ++i) {
        const DiagItem it = diag_item(p, blockIdx.x + i * gridDim.x);
        const int st = i % kDFStages;
        mbar_wait(&empty[st], ((i / kDFStages) & 1) ^ 1);
        mbar_expect_tx(&full[st], kDFStageBytes);
        uint8_t* s = smem + st * kDFStageBytes;
        tma_load_3d(s, &tmQKV, &full[st], p.q_col0 + it.h * 64, it.r0, it.n);
        tma_load_3d(s + 16384, &tmQKV, &full[st], p.k_col0 + it.h * 64, it.r0, it.n);
        tma_load_3d(s + 32768, &tmQKV, &full[st], p.v_col0 + it.h * 64, it.r0, it.n);
      }
    }
  } else if (warp == 1) {
    {                                        // whole warp, one elected lane issues (see tc_mma_bf16_e)
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
      auto issue_S = [&](int i) {
        const int st = i % kDFStages, b = i & 1;
        mbar_wait(&full[st], (i / kDFStages) & 1);
        if (i >= 2) mbar_wait(&slot_free[b], ((i >> 1) - 1) & 1);   // group b is done with item i-2
        tc_fence_after();
        const uint32_t aQ = smem_u32(smem + st * kDFStageBytes), aK = aQ + 16384;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_bf16_e(tmem_base + b * 256, umma_desc_sw128(aQ + kk * 32, 16, 1024), umma_desc_sw128(aK + kk * 32, 16, 1024),
                      idesc_s, kk != 0);
        tc_commit_e(&s_full[b]);
      };
      auto issue_PV = [&](int i) {
        const int st = i % kDFStages, b = i & 1;
        mbar_wait(&p_full[b], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t aV = smem_u32(smem + st * kDFStageBytes + 32768);
        const uint32_t aP = smem_u32(smem + kDFSmemP + b * 32768);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          tc_mma_bf16_e(tmem_base + b * 256 + 128, umma_desc_sw128(aP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                      umma_desc_sw128(aV + kk * 2048, 8192, 1024), idesc_o, kk != 0);
        tc_commit_e(&o_full[b]);
        tc_commit_e(&empty[st]);
      };
      for (int i = 0; i < n_items; ++i) {
        issue_S(i);
        if (i > 0) issue_PV(i - 1);
      }
      issue_PV(n_items - 1);
    }
  } else {
    const int grp = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    uint8_t* sP = smem + kDFSmemP + grp * 32768;
    const uint32_t tS = tmem_base + grp * 256 + lane_addr;
    // Global loads are kept off the critical path: the descriptor of this group's item after next and the mask words of
    // its next item are requested while the current item is processed (they land one iteration before they are used).
    auto load_item = [&](int i) -> DiagItem {
      return (i < n_items) ? diag_item(p, blockIdx.x + i * gridDim.x) : DiagItem{0, 0, 0, 0, 0};
    };
    auto load_mask = [&](const DiagItem& it, uint32_t (&w)[5]) {
      const int row = it.r0 + min(r, max(it.len - 1, 0));
      diag_mask_raw(p.mask_bits + (static_cast<size_t>(it.n) * p.S + row) * p.mask_words, it.r0, w);
    };
    DiagItem it = load_item(grp), it_next = load_item(grp + 2);
    uint32_t mraw[5];
    load_mask(it, mraw);
    for (int i = grp; i < n_items; i += 2) {
      const DiagItem it_next2 = load_item(i + 4);
      const bool row_ok = r < it.len;
      const uint32_t ph = (i >> 1) & 1;
      uint32_t mw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (!row_ok) mw[0] = mw[1] = mw[2] = mw[3] = 0u;
      else if (it.cls == 2) diag_mask_finish(mraw, it.r0, it.len, mw);
      load_mask(it_next, mraw);            // for the next iteration
      const uint32_t rowkey = drop_rowkey(p.drop.seed_lo, p.drop.seed_hi, it.n, it.h, it.r0 + r);
      const uint32_t rowkey2 = drop_rowkey2(rowkey);
      mbar_wait(&s_full[grp], ph);
      tc_fence_after();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains: ILP instead of a 128-deep max
      // Packed segments leave most of a 128 x 128 tile masked (~24 visible keys per row): 32-column chunks / 8-column
      // groups that no row of this warp can see are skipped (warp-uniform votes), they only get their zeros stored.
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const uint32_t w = (c == 0) ? mw[0] : (c == 1) ? mw[1] : (c == 2) ? mw[2] : mw[3];
        if (!__any_sync(0xffffffffu, w != 0u)) continue;
        uint32_t s[32];
        tmem_ld32(tS + c * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (!__any_sync(0xffffffffu, ((w >> (g * 8)) & 0xffu) != 0u)) continue;
#pragma unroll
          for (int j = g * 8; j < g * 8 + 8; ++j)
            m4[j & 3] = fmaxf(m4[j & 3], ((w >> j) & 1u) ? __uint_as_float(s[j]) : -INFINITY);
        }
      }
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;   // scale > 0 commutes with max
      const float m_use = (m == -INFINITY) ? 0.f : m;
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const uint32_t w = (c == 0) ? mw[0] : (c == 1) ? mw[1] : (c == 2) ? mw[2] : mw[3];
        uint8_t* pbase = sP + (c >> 1) * 16384 + r * 128;
        if (!__any_sync(0xffffffffu, w != 0u)) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(pbase + ((((c & 1) * 4 + g) ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
          continue;
        }
        uint32_t s[32];
        tmem_ld32(tS + c * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (__any_sync(0xffffffffu, ((w >> (g * 8)) & 0xffu) != 0u)) {
            float pv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float e = fast_exp2(fmaf(__uint_as_float(s[g * 8 + j]), p.scale_log2, -m_use));
              pv[j] = ((w >> (g * 8 + j)) & 1u) ? e : 0.f;
              l4[j & 3] += pv[j];
            }
            if (p.drop.thresh != 0u) {
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                const uint32_t bits = drop_bits(rowkey, rowkey2, it.r0 + c * 32 + g * 8 + j);
                if (!drop_keep_even(bits, p.drop.thresh)) pv[j] = 0.f;
                if (!drop_keep_odd(bits, p.drop.thresh)) pv[j + 1] = 0.f;
              }
            }
            o.x = pack_bf16(pv[0], pv[1]); o.y = pack_bf16(pv[2], pv[3]);
            o.z = pack_bf16(pv[4], pv[5]); o.w = pack_bf16(pv[6], pv[7]);
          }
          *reinterpret_cast<uint4*>(pbase + ((((c & 1) * 4 + g) ^ (r & 7)) << 4)) = o;
        }
      }
      const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[grp]);
      // ---- epilogue of the same item (the other group runs its softmax meanwhile)
      mbar_wait(&o_full[grp], ph);
      tc_fence_after();
      const float inv = (l > 0.f) ? p.drop.inv_keep / l : 0.f;
      // PV has retired (o_full), so this warp's own 32 rows of the P buffer are free: they stage the coalesced store
      uint8_t* stg = sP + quad * 4096;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(tS + 128 + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
          v.y = pack_bf16(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
          v.z = pack_bf16(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
          v.w = pack_bf16(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
          stage_put16(stg, lane, c * 4 + g, v);
        }
      }
      stage_flush_rows(stg, lane, p.out + (static_cast<long long>(it.n) * p.S + it.r0 + quad * 32) * p.ldo + it.h * 64, p.ldo,
                       it.len - quad * 32);
      if (row_ok && p.lse != nullptr)
        p.lse[(static_cast<size_t>(it.n) * p.H + it.h) * p.S + it.r0 + r] =
            (l > 0.f) ? (m * 0.6931471805599453f + logf(l)) : 0.f;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&slot_free[grp]);
      it = it_next;
      it_next = it_next2;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================
// backward: dQ, dK, dV of one isolated tile in a single pass
// =====================================================================================================
#ifdef GGPT_ATTN_TRACE      // profiling aid: clock64 stamps of compute warp 2 of 32 CTAs, 16 items each (tools/attn_trace.py diag)
__device__ long long g_diag_trace[32 * 16 * 8];
#define DIAG_TRACE(slot)                                                                                           \
  do {                                                                                                             \
    if (warp == 2 && lane == 0 && blockIdx.x < 32 && i >= 8 && i < 24)                                              \
      g_diag_trace[(blockIdx.x * 16 + (i - 8)) * 8 + (slot)] = clock64();                                           \
  } while (0)
#define DIAG_TRACE_EPI(slot)                                                                                       \
  do {                                                                                                             \
    if (warp == 10 && lane == 0 && blockIdx.x < 32 && i >= 8 && i < 24)                                             \
      g_diag_trace[(blockIdx.x * 16 + (i - 8)) * 8 + (slot)] = clock64();                                           \
  } while (0)
#else
#define DIAG_TRACE(slot) do { } while (0)
#define DIAG_TRACE_EPI(slot) do { } while (0)
#endif
constexpr int kDBThreads = 576;                       // producer, MMA, 8 softmax-gradient warps, 8 epilogue warps
constexpr int kDBStageBytes = 65536;                  // Q | K | V | dO
constexpr int kDBSmemP = 2 * kDBStageBytes;
constexpr int kDBSmemDS = kDBSmemP + 32768;
constexpr int kDBSmemEps = kDBSmemDS + 32768;         // float [2 items][2 halves][128]
constexpr int kDBSmemGather = kDBSmemEps + 2048;      // 8 warps x 4 KB: coalesced RoPE-table gather (epilogue)
constexpr int kDBSmemBars = kDBSmemGather + 32768;
constexpr int kDBSmem = kDBSmemBars + 256;

__global__ void __launch_bounds__(kDBThreads, 1)
attn_diag_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const DiagParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sP = smem + kDBSmemP;
  uint8_t* sDS = smem + kDBSmemDS;
  float* epsx = reinterpret_cast<float*>(smem + kDBSmemEps);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDBSmemBars);
  uint64_t* full = bars;             // [2]
  uint64_t* empty = bars + 2;        // [2]
  uint64_t* sdp_full = bars + 4;     // [2]
  uint64_t* pds_full = bars + 6;     // count 8
  uint64_t* pds_empty = bars + 7;
  uint64_t* acc_full = bars + 8;     // [2]
  uint64_t* acc_empty = bars + 10;   // [2] count 8
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp: provably uniform
  const int total = p.iso_count[0] * p.H;
  if (static_cast<int>(blockIdx.x) >= total) return;
  const int n_items = (total - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&sdp_full[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
    }
    mbar_init(pds_full, 8);
    mbar_init(pds_empty, 1);
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
  // slot b (256 columns): S [0,128) dP [128,256); after the softmax pass the same columns are reused as
  // dV [0,64) dK [64,128) dQ [128,192) PK [192,256)

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_items; ++i) {
        const DiagItem it = diag_item(p, blockIdx.x + i * gridDim.x);
        const int st = i & 1;
        mbar_wait(&empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&full[st], kDBStageBytes);
        uint8_t* s = smem + st * kDBStageBytes;
        tma_load_3d(s, &tmQKV, &full[st], p.q_col0 + it.h * 64, it.r0, it.n);
        tma_load_3d(s + 16384, &tmQKV, &full[st], p.k_col0 + it.h * 64, it.r0, it.n);
        tma_load_3d(s + 32768, &tmQKV, &full[st], p.v_col0 + it.h * 64, it.r0, it.n);
        tma_load_3d(s + 49152, &tmDO, &full[st], it.h * 64, it.r0, it.n);
      }
    }
  } else if (warp == 1) {
    {                                        // whole warp, one elected lane issues (see tc_mma_bf16_e)
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, true, true);
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, false, true);
      const uint32_t aP = smem_u32(sP), aDS = smem_u32(sDS);
      auto scores_ready = [&](int i) -> bool {
        const int st = i & 1;
        if (!mbar_try_wait_warp(&full[st], (i >> 1) & 1)) return false;
        return i < 2 || mbar_try_wait_warp(&acc_empty[st], ((i >> 1) - 1) & 1);   // epilogue of item i-2 drained this slot
      };
      auto issue_scores = [&](int i) {
        const int st = i & 1;
        tc_fence_after();
        const uint32_t aQ = smem_u32(smem + st * kDBStageBytes), aK = aQ + 16384, aV = aQ + 32768, aDO = aQ + 49152;
        const uint32_t d = tmem_base + st * 256;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_bf16_e(d, umma_desc_sw128(aQ + kk * 32, 16, 1024), umma_desc_sw128(aK + kk * 32, 16, 1024), idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_bf16_e(d + 128, umma_desc_sw128(aDO + kk * 32, 16, 1024), umma_desc_sw128(aV + kk * 32, 16, 1024), idesc_s,
                      kk != 0);
        tc_commit_e(&sdp_full[st]);
      };
      auto issue_grads = [&](int i) {
#ifdef GGPT_ATTN_TRACE
        if (lane == 0 && blockIdx.x < 32 && i >= 8 && i < 24) g_diag_trace[(blockIdx.x * 16 + (i - 8)) * 8 + 5] = clock64();
#endif
        tc_fence_after();
        const int st = i & 1;
        const uint32_t aQ = smem_u32(smem + st * kDBStageBytes), aK = aQ + 16384, aDO = aQ + 49152;
        const uint32_t d = tmem_base + st * 256;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dV = P^T dO
          tc_mma_bf16_e(d, umma_desc_sw128(aP + kk * 2048, 16384, 1024), umma_desc_sw128(aDO + kk * 2048, 8192, 1024),
                      idesc_t, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dK = dS^T Q
          tc_mma_bf16_e(d + 64, umma_desc_sw128(aDS + kk * 2048, 16384, 1024), umma_desc_sw128(aQ + kk * 2048, 8192, 1024),
                      idesc_t, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dQ = dS K
          tc_mma_bf16_e(d + 128, umma_desc_sw128(aDS + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                      umma_desc_sw128(aK + kk * 2048, 8192, 1024), idesc_q, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // PK = P K  (row-sum correction, see attn_bwd_sm100.cu)
          tc_mma_bf16_e(d + 192, umma_desc_sw128(aP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                      umma_desc_sw128(aK + kk * 2048, 8192, 1024), idesc_q, kk != 0);
        tc_commit_e(&acc_full[st]);
        tc_commit_e(pds_empty);
        tc_commit_e(&empty[st]);
      };
      // Two independent streams of work for the tensor core: the scores of item a (need its stage loaded and the TMEM
      // slot of item a-2 drained by the epilogue warps) and the gradient products of item b (need P / dS of item b).
      // The thread polls both and issues whichever is ready: the gradient MMAs of item b must never sit behind a wait for
      // the epilogue of item b-1 (they release the P / dS buffer the softmax-gradient warps are waiting for), nor the
      // scores of item b+1 behind the softmax of item b.
      int a = 0, b = 0;
      uint32_t spins = 0;
      while (b < n_items) {
        bool progressed = false;
        if (a < n_items && a <= b + 1 && scores_ready(a)) {
          issue_scores(a);
          ++a;
          progressed = true;
        }
        if (b < a && mbar_try_wait_warp(pds_full, b & 1)) {
          issue_grads(b);
          ++b;
          progressed = true;
        }
        if (progressed) {
          spins = 0;
        } else if (++spins > (1u << 26)) {
          printf("ggpt attn_diag_bwd: MMA issuer stuck block=%d a=%d b=%d\n", blockIdx.x, a, b);
          __trap();
        }
      }
    }
  } else if (warp < 10) {
    // ===================== softmax-gradient warps (2..9): two per TMEM lane quadrant, 64 key columns each ==============
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    // Per-row inputs (the 3 mask words covering this warp's 64 key columns, lse, D) of item i+1 and the descriptor of item
    // i+1 are requested while item i is processed: no global-load latency on the critical path.
    struct RowPre {
      uint32_t w[5];
      float lse, dsum;
    };
    auto load_item = [&](int i) -> DiagItem {
      return (i < n_items) ? diag_item(p, blockIdx.x + i * gridDim.x) : DiagItem{0, 0, 0, 0, 0};
    };
    auto load_row = [&](const DiagItem& it, RowPre& o) {
      const int rr = min(r, max(it.len - 1, 0));
      const size_t row = static_cast<size_t>(it.n) * p.S + it.r0 + rr;
      const uint32_t* mrow = p.mask_bits + row * p.mask_words + (it.r0 >> 5);
#pragma unroll
      for (int k = 0; k < 5; ++k) o.w[k] = __ldg(mrow + k);
      const size_t li = (static_cast<size_t>(it.n) * p.H + it.h) * p.S + it.r0 + rr;
      o.lse = __ldg(p.lse_in + li);
      o.dsum = __ldg(p.dsum + li);
    };
    DiagItem it = load_item(0), it_next = load_item(1);
    RowPre rp;
    load_row(it, rp);
    for (int i = 0; i < n_items; ++i) {
      DIAG_TRACE(0);
      const DiagItem it_next2 = load_item(i + 2);
      RowPre rp_next;
      load_row(it_next, rp_next);
      const bool row_ok = r < it.len;
      const float lse2 = row_ok ? rp.lse * 1.4426950408889634f : 0.f;
      const float dsum = row_ok ? rp.dsum : 0.f;
      // The two warps of a quadrant take ALTERNATE 16-column pieces of the 128 keys (warp `half`: pieces half, half + 2,
      // half + 4, half + 6).  Splitting the keys into a left and a right half left all the work of a block-diagonal mask
      // with one of the two warps (clock64 trace: the busy warps needed 8k cycles per item, their partners 3.7k idling at
      // the barrier).  mw16[q] = the 16 mask bits of this row for piece 2 q + half.
      uint32_t mw16[4] = {0xffffu, 0xffffu, 0xffffu, 0xffffu};
      if (!row_ok) {
        mw16[0] = mw16[1] = mw16[2] = mw16[3] = 0u;
      } else if (it.cls == 2) {
        const int sh = it.r0 & 31;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t v = __funnelshift_r(rp.w[q], rp.w[q + 1], sh);
          const int nvalid = it.len - 32 * q;
          const uint32_t m32 = v & (nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u)));
          mw16[q] = (m32 >> (half * 16)) & 0xffffu;
        }
      } else if (it.len < 128) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int nvalid = it.len - 32 * q - 16 * half;
          mw16[q] = nvalid >= 16 ? 0xffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
        }
      }
      const int st = i & 1;
      const uint32_t rowkey = drop_rowkey(p.drop.seed_lo, p.drop.seed_hi, it.n, it.h, it.r0 + r);
      const uint32_t rowkey2 = drop_rowkey2(rowkey);
      DIAG_TRACE(1);
      mbar_wait(&sdp_full[st], (i >> 1) & 1);
      tc_fence_after();
      DIAG_TRACE(2);
      const uint32_t tS = tmem_base + st * 256 + lane_addr;
      float eps = 0.f;
      if (i > 0) mbar_wait(pds_empty, (i - 1) & 1);   // P / dS smem is single-buffered: the gradient MMAs of item i-1 have read it
      DIAG_TRACE(3);
      // 16 key columns at a time (two tcgen05.ld x16 per piece): with 32-column pieces the warp needed more than the 96
      // registers an 18-warp CTA allows, and the spill stores of the PREFETCHED per-row inputs then waited for their global
      // loads (4.5k cycles per item, clock64 trace).  Pieces / 8-column groups no row of this warp can see only get zeros
      // stored (see the forward kernel).
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int c16 = 2 * q + half;                                 // 16-column piece of the 128-key tile
        const uint32_t w = (q == 0) ? mw16[0] : (q == 1) ? mw16[1] : (q == 2) ? mw16[2] : mw16[3];
        uint8_t* pbase = sP + (c16 >> 2) * 16384 + r * 128;           // 64-key K block, then the row
        uint8_t* dbase = sDS + (c16 >> 2) * 16384 + r * 128;
        const int ch0 = (c16 & 3) * 2;                                // first 16-byte chunk of the piece inside the row
        if (!__any_sync(0xffffffffu, w != 0u)) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int chunk = ((ch0 + g) ^ (r & 7)) << 4;
            *reinterpret_cast<uint4*>(pbase + chunk) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(dbase + chunk) = make_uint4(0u, 0u, 0u, 0u);
          }
          continue;
        }
        uint32_t s[16], dp[16];
        tmem_ld16(tS + c16 * 16, s);
        tmem_ld16(tS + 128 + c16 * 16, dp);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 o = make_uint4(0u, 0u, 0u, 0u), o2 = make_uint4(0u, 0u, 0u, 0u);
          if (__any_sync(0xffffffffu, ((w >> (g * 8)) & 0xffu) != 0u)) {
            float pv[8], dv[8];
            uint32_t bits = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float e = fast_exp2(fmaf(__uint_as_float(s[g * 8 + j]), p.scale_log2, -lse2));
              pv[j] = ((w >> (g * 8 + j)) & 1u) ? e : 0.f;
              float dpe = __uint_as_float(dp[g * 8 + j]);
              if (p.drop.thresh != 0u) {
                if ((j & 1) == 0) bits = drop_bits(rowkey, rowkey2, it.r0 + c16 * 16 + g * 8 + j);
                const bool keep = (j & 1) ? drop_keep_odd(bits, p.drop.thresh) : drop_keep_even(bits, p.drop.thresh);
                dpe = keep ? dpe * p.drop.inv_keep : 0.f;
                const float t0 = pv[j] * (dpe - dsum);
                dv[j] = t0 * p.scale;
                if (!keep) pv[j] = 0.f;          // sP holds the dropped P (unscaled) for dV = (P o mask)^T dO / (1-p)
              } else {
                const float t0 = pv[j] * (dpe - dsum);
                eps += t0;
                dv[j] = t0 * p.scale;
              }
            }
            o.x = pack_bf16(pv[0], pv[1]); o.y = pack_bf16(pv[2], pv[3]);
            o.z = pack_bf16(pv[4], pv[5]); o.w = pack_bf16(pv[6], pv[7]);
            o2.x = pack_bf16(dv[0], dv[1]); o2.y = pack_bf16(dv[2], dv[3]);
            o2.z = pack_bf16(dv[4], dv[5]); o2.w = pack_bf16(dv[6], dv[7]);
          }
          const int chunk = ((ch0 + g) ^ (r & 7)) << 4;
          *reinterpret_cast<uint4*>(pbase + chunk) = o;
          *reinterpret_cast<uint4*>(dbase + chunk) = o2;
        }
      }
      epsx[(i & 1) * 256 + half * 128 + r] = eps;   // read by the epilogue warps behind acc_full (barrier chain)
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      DIAG_TRACE(4);
      it = it_next;
      it_next = it_next2;
      rp = rp_next;
    }
  } else {
    // ===================== epilogue warps (10..17): two per TMEM lane quadrant ==========================================
    // They drain item i's gradient accumulators, un-rotate dQ / dK and store, WHILE the softmax-gradient warps are on
    // item i+1 (before: the same eight warps did both, one after the other — the epilogue was 7.3k of the 12.5k cycles
    // per item, clock64 trace tools/attn_trace.py).  half 0: dV and dK, half 1: dQ with the eps * (P K) correction.
    // Accumulators are taken 16 columns at a time (two or four tcgen05.ld x16 per wait) so that the warp stays within the
    // 96 registers an 18-warp CTA allows, and every thread stores 32 contiguous bytes (one full sector) per instruction.
    const int quad = warp & 3;
    const int half = (warp - 10) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    // RoPE table rows of this quadrant's tokens: the two warps of a quadrant need the same 32 cos rows and 32 sin rows, so
    // warp `half` fetches table `half` for both with cp.async into a 4 KB swizzled buffer; 64-thread named barriers
    // separate the partner's reads of the previous tables from the refill and publish the new ones.
    uint8_t* gb_cos = smem + kDBSmemGather + (quad * 2) * 4096;
    uint8_t* gb_sin = gb_cos + 4096;
    auto pair_barrier = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory"); };
    auto load_item = [&](int i) -> DiagItem {
      return (i < n_items) ? diag_item(p, blockIdx.x + i * gridDim.x) : DiagItem{0, 0, 0, 0, 0};
    };
    auto load_pos = [&](const DiagItem& it) -> int {
      return __ldg(p.pos + static_cast<size_t>(it.n) * p.S + it.r0 + min(r, max(it.len - 1, 0)));
    };
    auto st256 = [](__nv_bfloat16* dst, const float (&o)[16]) {
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(pack_bf16(o[0], o[1])),
                   "r"(pack_bf16(o[2], o[3])), "r"(pack_bf16(o[4], o[5])), "r"(pack_bf16(o[6], o[7])),
                   "r"(pack_bf16(o[8], o[9])), "r"(pack_bf16(o[10], o[11])), "r"(pack_bf16(o[12], o[13])),
                   "r"(pack_bf16(o[14], o[15]))
                   : "memory");
    };
    DiagItem it = load_item(0), it_next = load_item(1);
    int pos = load_pos(it);
    for (int i = 0; i < n_items; ++i) {
      const DiagItem it_next2 = load_item(i + 2);
      const int pos_next = load_pos(it_next);
      const int st = i & 1;
      pair_barrier();                                 // the partner has finished reading the previous tables
      {
        const float* tab = half ? p.sin_tab : p.cos_tab;
        uint8_t* dst = half ? gb_sin : gb_cos;
        const int rd_row = lane >> 3, rd_j = lane & 7;
#pragma unroll
        for (int it8 = 0; it8 < 8; ++it8) {
          const int rr = it8 * 4 + rd_row;
          const int src = __shfl_sync(0xffffffffu, pos, rr);
          const float* g = tab + static_cast<long long>(src) * 32 + rd_j * 4;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + rr * 128 + ((rd_j ^ (rr & 7)) << 4))),
                       "l"(g)
                       : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      const uint32_t d = tmem_base + st * 256 + lane_addr;
      const bool ok = r < it.len;
      __nv_bfloat16* orow = p.dqkv + (static_cast<long long>(it.n) * p.S + it.r0 + r) * p.ld_dqkv + it.h * 64;
      mbar_wait(&acc_full[st], (i >> 1) & 1);
      tc_fence_after();
      DIAG_TRACE_EPI(6);
      if (half == 0) {
        // ---- dV: no rotation (dropout: keeps were rescaled by 1 / (1 - p))
        const float vs = (p.drop.thresh != 0u) ? p.drop.inv_keep : 1.0f;
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
          uint32_t x1[16], x2[16];
          tmem_ld16(d + cb * 32, x1);
          tmem_ld16(d + cb * 32 + 16, x2);
          tmem_ld_wait();
          float o1[16], o2[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            o1[j] = __uint_as_float(x1[j]) * vs;
            o2[j] = __uint_as_float(x2[j]) * vs;
          }
          if (ok) {
            st256(orow + p.v_col0 + cb * 32, o1);
            st256(orow + p.v_col0 + cb * 32 + 16, o2);
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      pair_barrier();                                 // both tables have landed
      // ---- rotated outputs: half 0 -> dK (accumulator columns [64,128)), half 1 -> dQ ([128,192), PK at [192,256))
      const uint32_t dacc = d + (half == 0 ? 64 : 128);
      const int col0 = half == 0 ? p.k_col0 : p.q_col0;
      float ce = 0.f;
      if (half == 1) {
        const float* e = epsx + (i & 1) * 256;
        ce = (e[r] + e[128 + r]) * p.scale;
      }
#pragma unroll 1
      for (int jb = 0; jb < 2; ++jb) {              // rotation pairs (j, j + 32), j in [16 jb, 16 jb + 16)
        uint32_t x1[16], x2[16];
        tmem_ld16(dacc + jb * 16, x1);
        tmem_ld16(dacc + 32 + jb * 16, x2);
        if (half == 1) {
          uint32_t y1[16], y2[16];
          tmem_ld16(d + 192 + jb * 16, y1);
          tmem_ld16(d + 192 + 32 + jb * 16, y2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            x1[j] = __float_as_uint(__uint_as_float(x1[j]) - ce * __uint_as_float(y1[j]));
            x2[j] = __float_as_uint(__uint_as_float(x2[j]) - ce * __uint_as_float(y2[j]));
          }
        } else {
          tmem_ld_wait();
        }
        float o1[16], o2[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = ((jb * 4 + q) ^ (lane & 7)) << 4;
          const float4 a = *reinterpret_cast<const float4*>(gb_cos + lane * 128 + ch);
          const float4 b4 = *reinterpret_cast<const float4*>(gb_sin + lane * 128 + ch);
          const float cs[4] = {a.x, a.y, a.z, a.w}, sn[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float d1 = __uint_as_float(x1[q * 4 + j]), d2 = __uint_as_float(x2[q * 4 + j]);
            o1[q * 4 + j] = d1 * cs[j] + d2 * sn[j];       // transpose of the forward rotation
            o2[q * 4 + j] = d2 * cs[j] - d1 * sn[j];
          }
        }
        if (ok) {
          st256(orow + col0 + jb * 16, o1);
          st256(orow + col0 + 32 + jb * 16, o2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[st]);
      DIAG_TRACE_EPI(7);
      it = it_next;
      it_next = it_next2;
      pos = pos_next;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// isolated tile: its diagonal pair is active and nothing else in its row or column of the class matrix is
// The work list holds one 16-byte descriptor {sequence, first row, rows, class} per isolated tile.
__global__ void attn_iso_kernel(const uint8_t* __restrict__ cls, const int* __restrict__ n_tiles,
                                const int* __restrict__ tile_start, int N, int max_tiles,
                                uint8_t* __restrict__ iso_flags, int* __restrict__ iso_list, int* __restrict__ iso_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * max_tiles) return;
  const int n = i / max_tiles, t = i % max_tiles;
  const int nt = n_tiles[n];
  bool iso = false;
  if (t < nt) {
    const uint8_t* c = cls + static_cast<size_t>(n) * max_tiles * max_tiles;
    iso = c[t * max_tiles + t] != 0;
    for (int j = 0; j < nt && iso; ++j)
      if (j != t && (c[t * max_tiles + j] != 0 || c[j * max_tiles + t] != 0)) iso = false;
  }
  iso_flags[i] = iso ? 1 : 0;
  if (iso) {
    const int* ts = tile_start + static_cast<size_t>(n) * (max_tiles + 1);
    const int slot = atomicAdd(iso_count, 1);
    reinterpret_cast<int4*>(iso_list)[slot] =
        make_int4(n, ts[t], ts[t + 1] - ts[t], cls[(static_cast<size_t>(n) * max_tiles + t) * max_tiles + t]);
  }
  if (t == 0) atomicAdd(iso_count + 1, nt);   // iso_count[1] = total number of row tiles in the batch
}

static int set_smem_attr(const void* fn, int bytes, bool* done) {
  if (*done) return 0;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("attn_diag: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return -2;
  }
  *done = true;
  return 0;
}

#ifdef GGPT_ATTN_TRACE
int diag_trace_read(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_diag_trace, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
#endif

int attn_iso_build(const uint8_t* cls, const int* n_tiles, const int* tile_start, int N, int max_tiles, uint8_t* iso_flags,
                   int* iso_list, int* iso_count, cudaStream_t s) {
  cudaMemsetAsync(iso_count, 0, 2 * sizeof(int), s);
  const int n = N * max_tiles;
  attn_iso_kernel<<<(n + 255) / 256, 256, 0, s>>>(cls, n_tiles, tile_start, N, max_tiles, iso_flags, iso_list, iso_count);
  return check_launch("attn_iso_kernel");
}

int attn_diag_fwd_launch(const CUtensorMap& tm, const DiagParams& p, cudaStream_t s) {
  static bool done = false;
  if (int rc = set_smem_attr(reinterpret_cast<const void*>(attn_diag_fwd_kernel), kDFSmem, &done)) return rc;
  attn_diag_fwd_kernel<<<num_sms(), kDFThreads, kDFSmem, s>>>(tm, p);
  return check_launch("attn_diag_fwd_kernel");
}

int attn_diag_bwd_launch(const CUtensorMap& tmQKV, const CUtensorMap& tmDO, const DiagParams& p, cudaStream_t s) {
  static bool done = false;
  if (int rc = set_smem_attr(reinterpret_cast<const void*>(attn_diag_bwd_kernel), kDBSmem, &done)) return rc;
  attn_diag_bwd_kernel<<<num_sms(), kDBThreads, kDBSmem, s>>>(tmQKV, tmDO, p);
  return check_launch("attn_diag_bwd_kernel");
}

}  // namespace ggpt
