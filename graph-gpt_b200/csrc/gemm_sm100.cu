// tcgen05 GEMM for the GraphGPT hot path (sm_100a).
//
//   C[M,N] = A[M,K] * B[N,K]^T      bf16 operands, fp32 accumulation in TMEM
//
// One persistent CTA per SM, 10 warps: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) and TMEM
// owner, warps 2..9 = epilogue (TMEM -> registers -> fused epilogue -> HBM; each TMEM lane quadrant is served by
// two warps that split the tile's columns, so the epilogue issue rate keeps up with the 128x256x64 mainloop).  Tiles are 128 x BN x 64 with a
// multi-stage smem ring (TMA SWIZZLE_128B) and two TMEM accumulator buffers so the epilogue of tile i overlaps
// the MMAs of tile i+1.
//
// Operand majors cover the three GEMMs of a linear layer y = x W^T (W is nn.Linear's [out,in]):
//   forward  y  = x  W^T : A = x  [T,in]  K-major,  B = W [out,in] K-major
//   dgrad    dx = dy W   : A = dy [T,out] K-major,  B = W [out,in] read MN-major (K = out rows)
//   wgrad    dW = dy^T x : A = dy [T,out] MN-major, B = x [T,in]   MN-major      (K = T rows)
// so no transposed copies of weights or activations are ever made.
//
// CTA pairs (PAIR = true, the default for problems with >= 2 m-blocks): the two CTAs of a cluster — the two SMs of a
// TPC — execute ONE tcgen05.mma.cta_group::2 of shape 256 x BN x 16 per step.  Each CTA stages its own 128 rows of A and
// only HALF of the B tile's rows (16 + 16 KB per stage instead of 16 + 32 KB), all TMA bytes complete on the leader's
// full barrier, the leader's elected thread issues every MMA, a multicast tcgen05.commit releases the smem slot in both
// CTAs and hands each CTA its own 128 x BN accumulator, and the peer's epilogue warps arrive remotely on the leader's
// accumulator-empty barrier.  Halving the B bytes per FLOP (L2 -> SM and smem -> tensor core) is worth +5..9 % on every
// shape of the training step.  Two older modes remain for single-m-block problems and A/B profiling (GGPT_GEMM_MODE):
// a single CTA per tile, and a cluster of two CTAs along M that each fetch half of the shared B tile with TMA
// .multicast::cluster (32 KB of L2 reads per k-block per CTA instead of 48) and release slots with a multicast commit.
//
// Replaces (reference, all via torch/cuBLAS): q/k/v/o_proj HF modeling_llama.py:262-264,288; gate/up/down_proj
// HF:182-184; n_token_proj modeling_pretrain.py:89-93; lm_head modeling_pretrain.py:218; and their autograd.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/ggpt_b200.h"

namespace ggpt {

enum : int { EPI_BF16 = 0, EPI_F32 = 1, EPI_RESID = 2, EPI_GEGLU = 3, EPI_QKV_ROPE = 4, EPI_DGEGLU = 5, EPI_DGEGLU_TMA = 6 };

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int num_splits;       // split-K factor (wgrad: K = tokens is huge while M x N has few tiles); >1 => atomic fp32 adds
  int kb_per_split;
  int b_half_rows;      // GEGLU: row offset of the `up` half inside B (= N/2)
  // epilogue
  void* C;              // bf16 (EPI_BF16/GEGLU/QKV_ROPE) or f32 (EPI_F32/RESID) output
  long long ldc;
  void* C2;             // GEGLU: act = gelu(gate)*up, bf16 [M, N/2]
  long long ldc2;
  const float* resid;   // RESID: fp32 [M,N]
  long long ldr;
  const float* colscale;  // RESID: optional per-column scale (LayerScale lambda), fp32 [N]
  const float* rowscale;  // RESID: optional per-row scale (DropPath keep/p), fp32 [M]
  int accumulate;         // F32: C += acc
  const void* gu;         // DGEGLU: saved [gate | up] of the forward pass, bf16 [M, 2N]
  long long ldgu;
  const int* pos;         // QKV_ROPE: position per row
  const float* cos_tab;   // QKV_ROPE: [max_pos, 32]
  const float* sin_tab;
  int rope_cols;          // QKV_ROPE: rotate columns [0, rope_cols) in heads of 64
  int direct_store;       // bf16 epilogues: every thread stores its row's columns straight from registers, 32 bytes (one
                          // full sector) per instruction, instead of transposing through the staging buffer
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)

template <int BN, bool PAIR, int EPI>
struct GemmSmem {
  static constexpr int kABytes = BM * BK * 2;  // 16 KB
  static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * BK * 2;   // a CTA of a pair holds half of the B tile's rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  // EPI_DGEGLU_TMA trades mainloop stages for a 3-deep ring of TMA-staged factor tiles (see its epilogue)
  static constexpr int kStages = (EPI == EPI_DGEGLU_TMA) ? (PAIR ? 4 : 2) : ((BN == 256 && !PAIR) ? 4 : 6);
  // one 32 x 128 B transposition buffer per epilogue warp;  DGEGLU_TMA: 3 x [gf1 chunk 16 KB | gf2 chunk 16 KB]
  static constexpr int kStagingBytes = (EPI == EPI_DGEGLU_TMA) ? 3 * 32768 : 8 * 4096;
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;  // +1024 alignment slack
};

#ifdef GGPT_ATTN_TRACE      // profiling aid (tools/attn_trace.py gemm): clock64 stamps of epilogue warp 2 / the MMA thread, 32 CTAs x 24 tiles
__device__ long long g_gemm_trace[32 * 24 * 8];
#define GEMM_TRACE(slot)                                                                                         \
  do {                                                                                                           \
    if (lane == 0 && blockIdx.x < 64 && (blockIdx.x & 1) == 0 && tcount < 24)                                     \
      g_gemm_trace[((blockIdx.x >> 1) * 24 + tcount) * 8 + (slot)] = clock64();                                   \
  } while (0)
#else
#define GEMM_TRACE(slot) do { } while (0)
#endif

template <int BN, bool A_MN, bool B_MN, int EPI, int CL, bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmF, const GemmParams p) {
  static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
  using S = GemmSmem<BN, PAIR, EPI>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + kStages * S::kStageBytes + S::kStagingBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  // work items = (m-block group of CL, n-block, k-split); the CTAs of a cluster take consecutive m-blocks
  // k-split is the SLOW index: the CTAs running concurrently work on the same K range of different output tiles, so the
  // A / B slabs of that range are fetched from HBM once and shared through L2 (split-fastest order re-read them ~2x).
  const int mn_tiles = ((p.num_m_blocks + CL - 1) / CL) * p.num_n_blocks;
  const int num_tiles = mn_tiles * p.num_splits;
  const int rank = (CL > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int item0 = blockIdx.x / CL;
  const int item_stride = gridDim.x / CL;
  // Item order.  Default: round-robin (item0, item0 + stride, ...).  QKV + RoPE: each cluster takes a CONTIGUOUS range of
  // tiles with the n-block fastest, i.e. it sweeps the 9 n-blocks of one m-block back to back — the cos / sin rows of the
  // tile's tokens are then gathered once per m-block instead of once per tile (clock64 trace: the gather was 4.3k of the
  // epilogue's 10.2k cycles per tile against a 7.5k-cycle mainloop), and the A tile is re-read from L2 while it is hot.
  constexpr bool kContiguous = (EPI == EPI_QKV_ROPE);
  const int items_per = (num_tiles + item_stride - 1) / item_stride;
  const int it_begin = kContiguous ? item0 * items_per : item0;
  const int it_end = kContiguous ? min(num_tiles, it_begin + items_per) : num_tiles;
  const int it_step = kContiguous ? 1 : item_stride;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      // multicast mode: released by the MMA warp of every CTA that receives the multicast B tile;
      // pair mode: released by the leader's (multicast) commit only
      mbar_init(&empty_bar[s], PAIR ? 1 : CL);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], PAIR ? 16 : 8);   // pair mode: the epilogue warps of BOTH CTAs arrive on the leader's barrier
    mbar_init(&tempty_bar[1], PAIR ? 16 : 8);
    if constexpr (EPI == EPI_DGEGLU_TMA) {
      uint64_t* fbar = reinterpret_cast<uint64_t*>(bar_base + 192);
#pragma unroll
      for (int i = 0; i < 3; ++i) mbar_init(&fbar[i], 1);
      tma_prefetch_desc(&tmE);
      tma_prefetch_desc(&tmF);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_pair(tmem_slot, 2 * BN);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, 2 * BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();   // peers' barriers must be initialised before any multicast / remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = it_begin; item < it_end; item += it_step) {
        const int tile = item % mn_tiles, split = item / mn_tiles;
        const int m0 = ((tile / p.num_n_blocks) * CL + rank) * BM;
        const int nb = tile % p.num_n_blocks;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          const int k0 = kb * BK;
          if constexpr (PAIR) {
            // Both CTAs load their own 128 rows of A and their own half of B's rows into their own smem; every byte
            // completes on the LEADER's full barrier, which therefore expects two stages' worth.
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
            if (!A_MN) {
              tma_load_2d_pair(sa, &tmA, &full_bar[stage], k0, m0);
            } else {
              tma_load_2d_pair(sa, &tmA, &full_bar[stage], m0, k0);
              tma_load_2d_pair(sa + 8192, &tmA, &full_bar[stage], m0 + 64, k0);
            }
            if (EPI == EPI_GEGLU) {          // leader: gate rows, peer: the matching up rows
              const int nh = nb * (BN / 2);
              tma_load_2d_pair(sb, &tmB, &full_bar[stage], k0, rank ? p.b_half_rows + nh : nh);
            } else if (!B_MN) {
              tma_load_2d_pair(sb, &tmB, &full_bar[stage], k0, nb * BN + rank * (BN / 2));
            } else {
#pragma unroll
              for (int i = 0; i < BN / 128; ++i)
                tma_load_2d_pair(sb + i * 8192, &tmB, &full_bar[stage], nb * BN + (rank * (BN / 128) + i) * 64, k0);
            }
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], S::kStageBytes);
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], m0, k0);
            tma_load_2d(sa + 8192, &tmA, &full_bar[stage], m0 + 64, k0);
          }
          if (CL == 1) {
            if (EPI == EPI_GEGLU) {
              // B rows: [gate tile | up tile], each BN/2 rows, from the two halves of the fused weight
              const int nh = nb * (BN / 2);
              tma_load_2d(sb, &tmB, &full_bar[stage], k0, nh);
              tma_load_2d(sb + (BN / 2) * 128, &tmB, &full_bar[stage], k0, p.b_half_rows + nh);
            } else if (!B_MN) {
              tma_load_2d(sb, &tmB, &full_bar[stage], k0, nb * BN);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_2d(sb + i * 8192, &tmB, &full_bar[stage], nb * BN + i * 64, k0);
            }
          } else {
            // this CTA fetches half `rank` of the B tile and multicasts it into both CTAs of the cluster
            if (EPI == EPI_GEGLU) {
              const int nh = nb * (BN / 2);
              tma_load_2d_mc(sb + rank * (BN / 2) * 128, &tmB, &full_bar[stage], k0, rank ? p.b_half_rows + nh : nh, 3);
            } else if (!B_MN) {
              tma_load_2d_mc(sb + rank * (BN / 2) * 128, &tmB, &full_bar[stage], k0, nb * BN + rank * (BN / 2), 3);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 128; ++i) {
                const int bi = rank * (BN / 128) + i;
                tma_load_2d_mc(sb + bi * 8192, &tmB, &full_bar[stage], nb * BN + bi * 64, k0, 3);
              }
            }
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop (warp-uniform control flow and values) and ONE elected lane executes the tcgen05
    // instructions: descriptors and barrier addresses then live in uniform registers.  Run by lane 0 alone, every MMA
    // cost ~20 instructions (per-lane descriptor arithmetic + an R2UR / ELECT / branch waterfall into the uniform operands
    // of UTCHMMA), ~100 per k-block from one thread that shares its scheduler with the epilogue warps.
#ifdef GGPT_MMA_LANE0      // the previous form, for same-box A/B runs (tools/gemm_ab.sh)
#define GGPT_MMA_ELECT() true
#define GGPT_MMA_SYNCWARP() do { } while (0)
    if (lane == 0 && (!PAIR || rank == 0)) {
#else
#define GGPT_MMA_ELECT() elect_one()
#define GGPT_MMA_SYNCWARP() __syncwarp()
    if (!PAIR || rank == 0) {                  // pair mode: the leader issues every MMA for both CTAs
#endif
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BM : BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tcount = 0;
      for (int item = it_begin; item < it_end; item += it_step, ++tcount) {
        const int split = item / mn_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        GEMM_TRACE(5);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        GEMM_TRACE(6);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adesc = A_MN ? umma_desc_sw128(sa + kk * 2048, 8192, 1024)
                                        : umma_desc_sw128(sa + kk * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_desc_sw128(sb + kk * 2048, 8192, 1024)
                                        : umma_desc_sw128(sb + kk * 32, 16, 1024);
            if (GGPT_MMA_ELECT()) {
              if (PAIR) tc_mma_bf16_pair(d_tmem, adesc, bdesc, idesc, (kb != kb0 || kk != 0) ? 1u : 0u);
              else tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (kb != kb0 || kk != 0) ? 1u : 0u);
            }
          }
          if (GGPT_MMA_ELECT()) {
            if (PAIR) tc_commit_pair_mc(&empty_bar[stage], 3);     // frees the slot in both CTAs of the pair
            else if (CL > 1) tc_commit_mc(&empty_bar[stage], 3);   // frees the slot in BOTH CTAs (multicast B lands in both)
            else tc_commit(&empty_bar[stage]);                     // frees the smem slot when these MMAs retire
          }
          GGPT_MMA_SYNCWARP();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        GEMM_TRACE(7);
        if (GGPT_MMA_ELECT()) {
          if (PAIR) tc_commit_pair_mc(&tfull_bar[acc], 3);   // each CTA's epilogue drains its own 128 rows
          else tc_commit(&tfull_bar[acc]);                   // accumulator complete
        }
        GGPT_MMA_SYNCWARP();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if constexpr (EPI == EPI_DGEGLU_TMA) {
    // ===================== Epilogue warps (2..9), down_proj dgrad + GeGLU backward with TMA-staged factors =============
    // acc = dact;  dgu = [ dact * gf[:, 0:N] | dact * gf[:, N:2N] ]  moves 8 bytes per accumulator element (4 in, 4 out):
    // HBM-bound, 0.25 ms per layer at 65 536 tokens against a 0.22 ms mainloop.  Per-lane __ldg factor loads could not
    // keep enough bytes in flight (0.55 ms).  Here the factor tiles travel by TMA: the accumulator tile is processed
    // in 64-column chunks; for chunk g a [128 x 64] box of each factor half is loaded into one of THREE 32 KB smem buffers
    // (SWIZZLE_128B, two chunks ahead of the compute), every thread multiplies its row's 32 columns IN PLACE (row-per-lane
    // 16-byte accesses are conflict-free under the 128-byte swizzle), and the same buffer leaves through a TMA store
    // (UTMASTG).  The ring spans tiles, so the loads of the next tile's first chunks overlap the current tile's tail.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* fring = smem + kStages * S::kStageBytes;
    uint64_t* fbar = reinterpret_cast<uint64_t*>(bar_base + 192);
    const bool leader = (warp == 2 && lane == 0);
    constexpr int kCh = BN / 64;
    const int my_tiles = (num_tiles > item0) ? (num_tiles - item0 + item_stride - 1) / item_stride : 0;
    const int n_chunks = my_tiles * kCh;
    auto chunk_coords = [&](int g, int& m0, int& col) {
      const int item = item0 + (g / kCh) * item_stride;
      const int tile = item % mn_tiles;
      m0 = ((tile / p.num_n_blocks) * CL + rank) * BM;
      col = (tile % p.num_n_blocks) * BN + (g % kCh) * 64;
    };
    auto issue_load = [&](int g) {
      int m0, col;
      chunk_coords(g, m0, col);
      uint8_t* buf = fring + (g % 3) * 32768;
      mbar_expect_tx(&fbar[g % 3], 32768);
      tma_load_2d(buf, &tmE, &fbar[g % 3], col, m0);
      tma_load_2d(buf + 16384, &tmE, &fbar[g % 3], p.N + col, m0);
    };
    if (leader) {
      if (n_chunks > 0) issue_load(0);
      if (n_chunks > 1) issue_load(1);
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    const int rr = quad * 32 + lane;             // this thread's row of the tile
    for (int g = 0; g < n_chunks; ++g) {
      const int c = g % kCh;
      if (c == 0) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
      }
      mbar_wait(&fbar[g % 3], (g / 3) & 1);
      uint8_t* buf = fring + (g % 3) * 32768;
      uint32_t r[32];
      tmem_ld32(tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16) + c * 64 + half * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t off = rr * 128u + (((half * 4 + q) ^ (rr & 7)) << 4);
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          uint4* ptr = reinterpret_cast<uint4*>(buf + f * 16384 + off);
          const uint4 v = *ptr;
          const float2 f01 = unpack_bf16(v.x), f23 = unpack_bf16(v.y), f45 = unpack_bf16(v.z), f67 = unpack_bf16(v.w);
          uint4 o;
          o.x = pack_bf16(__uint_as_float(r[q * 8 + 0]) * f01.x, __uint_as_float(r[q * 8 + 1]) * f01.y);
          o.y = pack_bf16(__uint_as_float(r[q * 8 + 2]) * f23.x, __uint_as_float(r[q * 8 + 3]) * f23.y);
          o.z = pack_bf16(__uint_as_float(r[q * 8 + 4]) * f45.x, __uint_as_float(r[q * 8 + 5]) * f45.y);
          o.w = pack_bf16(__uint_as_float(r[q * 8 + 6]) * f67.x, __uint_as_float(r[q * 8 + 7]) * f67.y);
          *ptr = o;
        }
      }
      fence_proxy_async_smem();                  // the in-place products must be visible to the TMA store
      if (c == kCh - 1) {                        // accumulator fully drained: hand the TMEM buffer back to the MMA thread
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(&tempty_bar[acc], 0);
          else mbar_arrive(&tempty_bar[acc]);
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");   // all eight epilogue warps have written this chunk
      if (leader) {
        int m0, col;
        chunk_coords(g, m0, col);
        tma_store_2d(&tmF, buf, col, m0);
        tma_store_2d(&tmF, buf + 16384, p.N + col, m0);
        tma_store_commit();
        if (g + 2 < n_chunks) {
          tma_store_wait_read<1>();              // the store of chunk g-1 has read buffer (g+2) % 3: refill it
          issue_load(g + 2);
        }
      }
    }
    if (leader) tma_store_wait_all();
  } else {
    // ===================== Epilogue warps (2..9) =====================
    // TMEM rows are thread-private (lane = row), so writing them straight to HBM would touch 32 different
    // 128-byte lines per instruction.  Each warp instead transposes 32 x 128 B chunks through its own swizzled
    // smem buffer: write phase lane = row, read phase 8 lanes cover one row's 128 B -> every global access is a
    // full-line, fully coalesced 16 B/lane transaction (4 rows per instruction).
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;     // which half of the tile's columns this warp drains
    uint8_t* stg = smem + kStages * S::kStageBytes + (warp - 2) * 4096;
    const int rd_row = lane >> 3;         // read phase: row (it*4 + rd_row), 16-byte chunk rd_j
    const int rd_j = lane & 7;
    auto stage_put = [&](int j, uint4 v) {
      *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = v;
    };
    auto stage_get = [&](int rr) -> uint4 {
      return *reinterpret_cast<const uint4*>(stg + rr * 128 + ((rd_j ^ (rr & 7)) << 4));
    };
    auto put_bf16_32 = [&](int jbase, const uint32_t (&r)[32]) {   // 32 fp32 accumulators -> 4 chunks of 8 bf16
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
        o.y = pack_bf16(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
        o.z = pack_bf16(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
        o.w = pack_bf16(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
        stage_put(jbase + g, o);
      }
    };
    // 16 consecutive bf16 columns of this thread's row, packed, as ONE 32-byte store (a full sector)
    auto st256 = [](__nv_bfloat16* dst, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                    uint32_t a6, uint32_t a7) {
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4),
                   "r"(a5), "r"(a6), "r"(a7)
                   : "memory");
    };
    auto st_bf16_32 = [&](__nv_bfloat16* dst, const uint32_t (&r)[32]) {   // 32 fp32 accumulators -> 32 bf16 columns
#pragma unroll
      for (int g = 0; g < 2; ++g)
        st256(dst + g * 16, pack_bf16(__uint_as_float(r[g * 16 + 0]), __uint_as_float(r[g * 16 + 1])),
              pack_bf16(__uint_as_float(r[g * 16 + 2]), __uint_as_float(r[g * 16 + 3])),
              pack_bf16(__uint_as_float(r[g * 16 + 4]), __uint_as_float(r[g * 16 + 5])),
              pack_bf16(__uint_as_float(r[g * 16 + 6]), __uint_as_float(r[g * 16 + 7])),
              pack_bf16(__uint_as_float(r[g * 16 + 8]), __uint_as_float(r[g * 16 + 9])),
              pack_bf16(__uint_as_float(r[g * 16 + 10]), __uint_as_float(r[g * 16 + 11])),
              pack_bf16(__uint_as_float(r[g * 16 + 12]), __uint_as_float(r[g * 16 + 13])),
              pack_bf16(__uint_as_float(r[g * 16 + 14]), __uint_as_float(r[g * 16 + 15])));
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    int tcount = 0;
    // QKV_ROPE: cos/sin rows of this thread's token — the same for every head and every n-block of the m-block — gathered
    // (coalesced, through the staging buffer) once per m-block, BEFORE waiting for the accumulator
    float cc[(EPI == EPI_QKV_ROPE) ? 32 : 1], ss[(EPI == EPI_QKV_ROPE) ? 32 : 1];
    int rope_m0 = -1;
    for (int item = it_begin; item < it_end; item += it_step, ++tcount) {
      if (warp == 2) GEMM_TRACE(0);
      const int tile = item % mn_tiles;
      const int m0 = ((tile / p.num_n_blocks) * CL + rank) * BM;
      const int nb = tile % p.num_n_blocks;
      const int row0 = m0 + quad * 32;                 // first global row of this warp's quadrant
      const int row = row0 + lane;                     // write-phase row of this thread
      const bool row_ok = row < p.M;
      if constexpr (EPI == EPI_QKV_ROPE) {
        if (nb * BN + half * (BN / 2) < p.rope_cols && m0 != rope_m0) {   // warp-uniform
          const int pos = row_ok ? p.pos[row] : 0;
          warp_gather_rows32(p.cos_tab, pos, stg, lane, cc);
          warp_gather_rows32(p.sin_tab, pos, stg, lane, ss);
          rope_m0 = m0;
        }
      }
      // DGEGLU: saved GeGLU factors of this thread's (row, 4-column) pieces of one 32-column chunk (read-phase layout)
      uint2 f1[(EPI == EPI_DGEGLU) ? 8 : 1], f2[(EPI == EPI_DGEGLU) ? 8 : 1];
      auto dgeglu_load = [&](const __nv_bfloat16* gf, int c, uint2 (&o1)[8], uint2 (&o2)[8]) {
        const int gcol = nb * BN + c + rd_j * 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow = row0 + it * 4 + rd_row;
          o1[it] = make_uint2(0u, 0u);
          o2[it] = make_uint2(0u, 0u);
          if (grow < p.M && gcol < p.N) {
            const __nv_bfloat16* src = gf + static_cast<long long>(grow) * p.ldgu + gcol;
            o1[it] = __ldg(reinterpret_cast<const uint2*>(src));
            o2[it] = __ldg(reinterpret_cast<const uint2*>(src + p.N));
          }
        }
      };
      if constexpr (EPI == EPI_DGEGLU)   // chunk 0's factors are fetched while the mainloop of this tile still runs
        dgeglu_load(reinterpret_cast<const __nv_bfloat16*>(p.gu), half * (BN / 2), f1, f2);
      if (warp == 2) GEMM_TRACE(1);
      mbar_wait(&tfull_bar[acc], acc_phase);
      if (warp == 2) GEMM_TRACE(2);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);

      // staged 32 x 64 bf16 chunk -> dst[row0.., col0..col0+64), columns clipped at col_limit
      auto copy_out_bf16 = [&](__nv_bfloat16* dst, long long ld, int col0, int col_limit) {
        __syncwarp();
        const int gcol = col0 + rd_j * 8;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + rd_row;
          if (row0 + rr < p.M && gcol < col_limit)
            *reinterpret_cast<uint4*>(dst + static_cast<long long>(row0 + rr) * ld + gcol) = stage_get(rr);
        }
        __syncwarp();
      };

      if constexpr (EPI == EPI_BF16) {
        const int n0 = nb * BN;
#pragma unroll 1
        for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 64) {
          uint32_t r0[32], r1[32];
          tmem_ld32(taddr + c, r0);
          tmem_ld32(taddr + c + 32, r1);
          tmem_ld_wait();
          if (p.direct_store) {
            if (row_ok && n0 + c < p.N) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<long long>(row) * p.ldc + n0 + c;
              st_bf16_32(dst, r0);
              st_bf16_32(dst + 32, r1);
            }
            continue;
          }
          put_bf16_32(0, r0);
          put_bf16_32(4, r1);
          copy_out_bf16(reinterpret_cast<__nv_bfloat16*>(p.C), p.ldc, n0 + c, p.N);
        }
      } else if constexpr (EPI == EPI_F32 || EPI == EPI_RESID) {
        const int n0 = nb * BN;
        float* C = reinterpret_cast<float*>(p.C);
#pragma unroll 1
        for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) stage_put(j, make_uint4(r[j * 4 + 0], r[j * 4 + 1], r[j * 4 + 2], r[j * 4 + 3]));
          __syncwarp();
          const int gcol = n0 + c + rd_j * 4;
          float4 cs = make_float4(1.f, 1.f, 1.f, 1.f);
          if (EPI == EPI_RESID && p.colscale != nullptr && gcol < p.N)
            cs = *reinterpret_cast<const float4*>(p.colscale + gcol);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rd_row;
            const int grow = row0 + rr;
            if (grow < p.M && gcol < p.N) {
              const uint4 u = stage_get(rr);
              float4 o = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
              float4* dst = reinterpret_cast<float4*>(C + static_cast<long long>(grow) * p.ldc + gcol);
              if constexpr (EPI == EPI_RESID) {
                const float4 res = *reinterpret_cast<const float4*>(p.resid + static_cast<long long>(grow) * p.ldr + gcol);
                const float rs = p.rowscale != nullptr ? p.rowscale[grow] : 1.0f;
                o.x = fmaf(o.x, cs.x * rs, res.x);
                o.y = fmaf(o.y, cs.y * rs, res.y);
                o.z = fmaf(o.z, cs.z * rs, res.z);
                o.w = fmaf(o.w, cs.w * rs, res.w);
                *dst = o;
              } else if (p.num_splits > 1) {   // split-K partial sums: fire-and-forget vector reduction into L2
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z),
                             "f"(o.w)
                             : "memory");
              } else {
                if (p.accumulate) {
                  const float4 old = *dst;
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *dst = o;
              }
            }
          }
          __syncwarp();
        }
      } else if constexpr (EPI == EPI_GEGLU) {
        // tile columns [0,BN/2) = gate for fused columns nh.., [BN/2,BN) = the matching up columns;
        // this warp owns gate/up columns [c, c+64)
        const int nh = nb * (BN / 2);
        const int half_n = p.N / 2;
        const int c = half * (BN / 4);
        __nv_bfloat16* gf = reinterpret_cast<__nv_bfloat16*>(p.C);
        if (gf != nullptr) {
          // training: besides act = gelu(g) * u, store the two factors the backward pass multiplies dact with,
          //   gf = [ u * gelu'(g) | gelu(g) ]   (gelu = g Phi(g), gelu' = Phi(g) + g phi(g), both from ONE exponential),
          // so that GeGLU backward is transcendental-free and fits the epilogue of the down_proj dgrad GEMM.
          uint32_t actp[2][16];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t rg[32], ru[32];
            tmem_ld32(taddr + c + hh * 32, rg);
            tmem_ld32(taddr + BN / 2 + c + hh * 32, ru);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float a1[2], a2[2], ac[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float g = __uint_as_float(rg[j + e]), u = __uint_as_float(ru[j + e]);
                float cdf, pdf;
                gelu_cdf_pdf(g, cdf, pdf);
                a2[e] = g * cdf;
                a1[e] = u * fmaf(g, pdf, cdf);
                ac[e] = a2[e] * u;
              }
              rg[j >> 1] = pack_bf16(a1[0], a1[1]);          // packed in place: entries [0,16) of rg / ru
              ru[j >> 1] = pack_bf16(a2[0], a2[1]);
              actp[hh][j >> 1] = pack_bf16(ac[0], ac[1]);
            }
            if (p.direct_store) {      // this row's 32 columns of each factor: two 32-byte stores per factor
              if (row_ok && nh + c + hh * 32 < half_n) {
                __nv_bfloat16* d1 = gf + static_cast<long long>(row) * p.ldc + nh + c + hh * 32;
                st256(d1, rg[0], rg[1], rg[2], rg[3], rg[4], rg[5], rg[6], rg[7]);
                st256(d1 + 16, rg[8], rg[9], rg[10], rg[11], rg[12], rg[13], rg[14], rg[15]);
                st256(d1 + half_n, ru[0], ru[1], ru[2], ru[3], ru[4], ru[5], ru[6], ru[7]);
                st256(d1 + half_n + 16, ru[8], ru[9], ru[10], ru[11], ru[12], ru[13], ru[14], ru[15]);
              }
              continue;
            }
            // staging row = [ 32 cols of u*gelu' (64 B) | 32 cols of gelu (64 B) ]; lanes 0-3 of a row group write the first
            // factor, lanes 4-7 the second — 64-byte contiguous pieces, two full sectors each
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              stage_put(q, make_uint4(rg[q * 4 + 0], rg[q * 4 + 1], rg[q * 4 + 2], rg[q * 4 + 3]));
              stage_put(4 + q, make_uint4(ru[q * 4 + 0], ru[q * 4 + 1], ru[q * 4 + 2], ru[q * 4 + 3]));
            }
            __syncwarp();
            {
              const int fcol = nh + c + hh * 32 + (rd_j & 3) * 8;            // column inside the factor's half
              __nv_bfloat16* dst = gf + (rd_j >> 2) * half_n + fcol;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + rd_row;
                if (row0 + rr < p.M && fcol < half_n)
                  *reinterpret_cast<uint4*>(dst + static_cast<long long>(row0 + rr) * p.ldc) = stage_get(rr);
              }
            }
            __syncwarp();
          }
          if (p.direct_store) {
            if (row_ok && nh + c < half_n) {
              __nv_bfloat16* da = reinterpret_cast<__nv_bfloat16*>(p.C2) + static_cast<long long>(row) * p.ldc2 + nh + c;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                st256(da + hh * 32, actp[hh][0], actp[hh][1], actp[hh][2], actp[hh][3], actp[hh][4], actp[hh][5], actp[hh][6],
                      actp[hh][7]);
                st256(da + hh * 32 + 16, actp[hh][8], actp[hh][9], actp[hh][10], actp[hh][11], actp[hh][12], actp[hh][13],
                      actp[hh][14], actp[hh][15]);
              }
            }
          } else {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
              for (int q = 0; q < 4; ++q)
                stage_put(hh * 4 + q, make_uint4(actp[hh][q * 4 + 0], actp[hh][q * 4 + 1], actp[hh][q * 4 + 2], actp[hh][q * 4 + 3]));
          }
        } else {
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {                  // inference: act only, 1-MUFU GELU
            uint32_t rg[32], ru[32];
            tmem_ld32(taddr + c + hh * 32, rg);
            tmem_ld32(taddr + BN / 2 + c + hh * 32, ru);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              rg[j] = __float_as_uint(gelu_erf_fwd(__uint_as_float(rg[j])) * __uint_as_float(ru[j]));
            put_bf16_32(hh * 4, rg);
          }
        }
        if (!(p.direct_store && gf != nullptr)) copy_out_bf16(reinterpret_cast<__nv_bfloat16*>(p.C2), p.ldc2, nh + c, half_n);
      } else if constexpr (EPI == EPI_DGEGLU) {
        // acc = dact = d(act); fused GeGLU backward with the factors saved by the forward epilogue:
        //   dgate = dact * gf[:, 0:N] (= u gelu'(g)),  dup = dact * gf[:, N:2N] (= gelu(g))  — two multiplies per element.
        // The epilogue moves 8 bytes per accumulator element (4 in, 4 out), so it is the factor loads that must be kept
        // in flight: the loads of chunk k+1 are issued before chunk k is processed (and those of chunk 0 before the
        // accumulator wait, see above) — they do not depend on the accumulator.
        const int n0 = nb * BN;
        const __nv_bfloat16* gf = reinterpret_cast<const __nv_bfloat16*>(p.gu);
        __nv_bfloat16* dgu = reinterpret_cast<__nv_bfloat16*>(p.C);
        constexpr int kChunks = BN / 64;                       // 32-column chunks in this warp's half of the tile
        const int cbase = half * (BN / 2);
#pragma unroll
        for (int k = 0; k < kChunks; ++k) {
          const int c = cbase + k * 32;
          const int gcol = n0 + c + rd_j * 4;
          uint2 n1[8], n2[8];
          if (k + 1 < kChunks) dgeglu_load(gf, c + 32, n1, n2);
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) stage_put(j, make_uint4(r[j * 4 + 0], r[j * 4 + 1], r[j * 4 + 2], r[j * 4 + 3]));
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rd_row;
            const int grow = row0 + rr;
            if (grow < p.M && gcol < p.N) {
              const uint4 a4 = stage_get(rr);
              const float a0 = __uint_as_float(a4.x), a1 = __uint_as_float(a4.y), a2 = __uint_as_float(a4.z),
                          a3 = __uint_as_float(a4.w);
              const float2 p01 = unpack_bf16(f1[it].x), p23 = unpack_bf16(f1[it].y);
              const float2 q01 = unpack_bf16(f2[it].x), q23 = unpack_bf16(f2[it].y);
              __nv_bfloat16* dst = dgu + static_cast<long long>(grow) * p.ldc + gcol;
              uint2 o;
              o.x = pack_bf16(a0 * p01.x, a1 * p01.y); o.y = pack_bf16(a2 * p23.x, a3 * p23.y);
              *reinterpret_cast<uint2*>(dst) = o;
              o.x = pack_bf16(a0 * q01.x, a1 * q01.y); o.y = pack_bf16(a2 * q23.x, a3 * q23.y);
              *reinterpret_cast<uint2*>(dst + p.N) = o;
            }
          }
          __syncwarp();
          if (k + 1 < kChunks) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              f1[it] = n1[it];
              f2[it] = n2[it];
            }
          }
        }
      } else if constexpr (EPI == EPI_QKV_ROPE) {
        // one head (64 columns) per iteration, the rotation pairs (j, j + 32) 16 at a time: two tcgen05.ld x16 per wait keep
        // the live accumulator values at 32 registers next to the 64 cos / sin values that persist across the m-block;
        // every thread stores its row's 16 columns as one 32-byte sector
        const int n0 = nb * BN;
#pragma unroll 1
        for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 64) {
          const bool rot = (n0 + c) < p.rope_cols;
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<long long>(row) * p.ldc + n0 + c;
#pragma unroll
          for (int jb = 0; jb < 2; ++jb) {
            uint32_t x1[16], x2[16];
            tmem_ld16(taddr + c + jb * 16, x1);
            tmem_ld16(taddr + c + 32 + jb * 16, x2);
            tmem_ld_wait();
            if (rot) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float a = __uint_as_float(x1[j]), b = __uint_as_float(x2[j]);
                x1[j] = __float_as_uint(a * cc[jb * 16 + j] - b * ss[jb * 16 + j]);   // q*cos + rotate_half(q)*sin, first half
                x2[j] = __float_as_uint(b * cc[jb * 16 + j] + a * ss[jb * 16 + j]);   // second half
              }
            }
            if (row_ok && n0 + c < p.N) {
              auto pk = [](uint32_t lo, uint32_t hi) { return pack_bf16(__uint_as_float(lo), __uint_as_float(hi)); };
              st256(dst + jb * 16, pk(x1[0], x1[1]), pk(x1[2], x1[3]), pk(x1[4], x1[5]), pk(x1[6], x1[7]), pk(x1[8], x1[9]),
                    pk(x1[10], x1[11]), pk(x1[12], x1[13]), pk(x1[14], x1[15]));
              st256(dst + 32 + jb * 16, pk(x2[0], x2[1]), pk(x2[2], x2[3]), pk(x2[4], x2[5]), pk(x2[6], x2[7]), pk(x2[8], x2[9]),
                    pk(x2[10], x2[11]), pk(x2[12], x2[13]), pk(x2[14], x2[15]));
            }
          }
        }
      }

      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(&tempty_bar[acc], 0);
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (warp == 2) GEMM_TRACE(3);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all();   // the peer may still multicast into this CTA's smem / arrive on its barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 2 * BN);
    else tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---------------------------------------------------------------------------------------------
// Host launcher
// ---------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, int EPI, int CL, bool PAIR>
static int launch_gemm_cl(const void* A, long long lda, const void* B, long long ldb, GemmParams p, cudaStream_t stream) {
  using S = GemmSmem<BN, PAIR, EPI>;
  CUtensorMap tmA, tmB, tmE, tmF;
  int rc;
  // A: K-major -> global [M rows, K cols], box 128 x 64.  MN-major -> global [K rows, M cols], box 64 x 64.
  if (!A_MN) rc = make_tmap_2d_bf16(&tmA, A, p.M, p.K, lda, BM, 64);
  else rc = make_tmap_2d_bf16(&tmA, A, p.K, p.M, lda, 64, 64);
  if (rc) return rc;
  const uint64_t b_rows_total = static_cast<uint64_t>(p.N);
  // B K-major: box = whole tile (CL 1) or the half this CTA multicasts (CL 2; GEGLU always loads gate/up halves)
  if (EPI == EPI_GEGLU || (!B_MN && CL == 2)) rc = make_tmap_2d_bf16(&tmB, B, b_rows_total, p.K, ldb, BN / 2, 64);
  else if (!B_MN) rc = make_tmap_2d_bf16(&tmB, B, b_rows_total, p.K, ldb, BN, 64);
  else rc = make_tmap_2d_bf16(&tmB, B, p.K, b_rows_total, ldb, 64, 64);
  if (rc) return rc;

  if (EPI == EPI_DGEGLU_TMA) {   // factor tiles in, products out: [M, 2N] bf16, boxes of 128 rows x 64 columns
    rc = make_tmap_2d_bf16(&tmE, p.gu, p.M, 2ull * p.N, p.ldgu, BM, 64);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tmF, p.C, p.M, 2ull * p.N, p.ldc, BM, 64);
    if (rc) return rc;
  } else {
    tmE = tmA;
    tmF = tmA;
  }
  const int items = ((p.num_m_blocks + CL - 1) / CL) * p.num_n_blocks * p.num_splits;
  int grid = num_sms() / CL * CL;
  if (grid > items * CL) grid = items * CL;

  auto kern = gemm_kernel<BN, A_MN, B_MN, EPI, CL, PAIR>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) {
      set_error("gemm: cudaFuncSetAttribute(%d B smem) failed: %s", S::kTotal, cudaGetErrorString(e));
      return -2;
    }
    attr_set = true;
  }
  if (CL == 1) {
    kern<<<grid, kGemmThreads, S::kTotal, stream>>>(tmA, tmB, tmE, tmF, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = S::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmE, tmF, p);
    if (e != cudaSuccess) {
      set_error("gemm: cluster launch failed: %s", cudaGetErrorString(e));
      return -2;
    }
  }
  return check_launch("gemm_kernel");
}

// Split-K plan of an accumulating fp32 GEMM (wgrad): equal-sized work items, so what matters is wave quantisation over
// the CTAs (or CTA pairs) resident at once.  First the best achievable occupancy with at most four waves (more, shorter
// splits only add reduction traffic; every split keeps >= 16 k-blocks), then the smallest split factor within 1 % of it
// that gives every CTA at least `min_waves` items (measured: with one or two items per CTA the fp32 epilogue of a tile is
// poorly overlapped by the next mainloop — gate|up wgrad 1475 TF/s at three waves, 1290 at one or two), else two, else any.
// Pure host arithmetic: ggpt_gemm_split_plan exposes it to the CPU tests.
static void plan_split_k(int num_m_blocks, int num_n_blocks, int num_k_blocks, int sms, bool clustered, int min_waves,
                         int* num_splits, int* kb_per_split) {
  const int cl = clustered ? 2 : 1;
  const int slots = sms / cl > 0 ? sms / cl : 1;                       // CTAs / CTA pairs resident at once
  const int tiles = ((num_m_blocks + cl - 1) / cl) * num_n_blocks;
  const int max_by_k = num_k_blocks / 16;
  auto eval = [&](int sp, int& kb, int& real, long long& waves) -> double {
    kb = (num_k_blocks + sp - 1) / sp;
    real = (num_k_blocks + kb - 1) / kb;                               // no empty splits
    const long long items = static_cast<long long>(tiles) * real;
    waves = (items + slots - 1) / slots;
    return static_cast<double>(items) / static_cast<double>(waves * slots);
  };
  const int sp_max = max_by_k < 64 ? max_by_k : 64;
  double best_eff = 0.0;
  int best_sp = 1, best_kb = num_k_blocks;
  for (int sp = 1; sp <= sp_max; ++sp) {
    int kb, real;
    long long waves;
    const double eff = eval(sp, kb, real, waves);
    if (eff > best_eff && (waves <= 4 || sp == 1)) best_eff = eff;
  }
  bool found = false;
  for (int pass = 0; pass < 3 && !found; ++pass) {
    for (int sp = 1; sp <= sp_max; ++sp) {
      int kb, real;
      long long waves;
      const double eff = eval(sp, kb, real, waves);
      if (eff >= best_eff - 0.01 && (waves <= 4 || sp == 1) && (pass == 2 || waves >= (pass == 0 ? min_waves : 2))) {
        best_sp = real;
        best_kb = kb;
        found = true;
        break;
      }
    }
  }
  *num_splits = best_sp;
  *kb_per_split = best_kb;
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const void* A, long long lda, const void* B, long long ldb, GemmParams p, cudaStream_t stream) {
  p.num_m_blocks = (p.M + BM - 1) / BM;
  if (EPI == EPI_GEGLU) p.num_n_blocks = (p.N / 2 + BN / 2 - 1) / (BN / 2);
  else p.num_n_blocks = (p.N + BN - 1) / BN;
  p.num_k_blocks = (p.K + BK - 1) / BK;
  p.b_half_rows = p.N / 2;
  // split-K only where it is safe (fp32 output that accumulates into a pre-initialised buffer) and useful (wgrad: K =
  // tokens is long while M x N has few tiles); plan_split_k picks the factor (configs[1]: o_proj 9 tiles x 24, qkv 27 x 8,
  // down 36 x 6, gate|up 72 x 3 — all 216 items = three waves of the 74 CTA pairs at 97 %).
  p.num_splits = 1;
  p.kb_per_split = p.num_k_blocks;
  if (EPI == EPI_F32 && p.accumulate) {
    static const char* mode_env0 = getenv("GGPT_GEMM_MODE");
    const bool clustered = p.num_m_blocks >= 2 && !(mode_env0 != nullptr && mode_env0[0] == 's');
    static const char* waves_env = getenv("GGPT_WGRAD_MIN_WAVES");
    const int max_by_k = p.num_k_blocks / 16;
    plan_split_k(p.num_m_blocks, p.num_n_blocks, p.num_k_blocks, num_sms(), clustered,
                 waves_env != nullptr ? atoi(waves_env) : 3, &p.num_splits, &p.kb_per_split);
    static const char* split_env = getenv("GGPT_WGRAD_SPLIT");      // "twowaves": the previous rule, for A/B runs
    if (split_env != nullptr && split_env[0] == 't') {
      const int mn_tiles = p.num_m_blocks * p.num_n_blocks;
      const int want = (2 * num_sms() + mn_tiles - 1) / mn_tiles;
      int sp = want < max_by_k ? want : max_by_k;
      if (sp < 1) sp = 1;
      p.kb_per_split = (p.num_k_blocks + sp - 1) / sp;
      p.num_splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    }
  }
  // clusters of two CTAs along M share (multicast) the B tile; single m-block problems stay un-clustered
  // (GGPT_GEMM_MODE = single | multicast | pair overrides the choice — a profiling aid)
  static const char* mode_env = getenv("GGPT_GEMM_MODE");
  static const int mode = mode_env == nullptr ? 2 : (mode_env[0] == 's' ? 0 : (mode_env[0] == 'm' ? 1 : 2));
  if (p.num_m_blocks >= 2 && mode == 2) return launch_gemm_cl<BN, A_MN, B_MN, EPI, 2, true>(A, lda, B, ldb, p, stream);
  if (p.num_m_blocks >= 2 && mode == 1) return launch_gemm_cl<BN, A_MN, B_MN, EPI, 2, false>(A, lda, B, ldb, p, stream);
  return launch_gemm_cl<BN, A_MN, B_MN, EPI, 1, false>(A, lda, B, ldb, p, stream);
}

template <int BN, int EPI>
static int dispatch_major(int a_mn, int b_mn, const void* A, long long lda, const void* B, long long ldb,
                          const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false, EPI>(A, lda, B, ldb, p, s);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true, EPI>(A, lda, B, ldb, p, s);
  if (a_mn && b_mn) return launch_gemm<BN, true, true, EPI>(A, lda, B, ldb, p, s);
  set_error("gemm: operand majors (A MN-major, B K-major) are not instantiated");
  return -1;
}

// direct 32-byte stores need 32-byte aligned rows and whole 64-column chunks (GGPT_GEMM_DIRECT_STORE=0/1 overrides: A/B runs)
static int want_direct_store(const void* C, long long ldc, int n_cols, int dflt) {
  static const char* env = getenv("GGPT_GEMM_DIRECT_STORE");
  const int on = env != nullptr ? (env[0] != '0') : dflt;   // default on: +1..4 % per GEMM, same-box A/B (profiles/r2aq_*)
  return on && C != nullptr && (reinterpret_cast<uintptr_t>(C) & 31) == 0 && (ldc % 16) == 0 && (n_cols % 64) == 0;
}

static int check_common(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K) {
  GGPT_REQUIRE(A && B, "gemm: null operand");
  GGPT_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  GGPT_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0, "gemm: lda/ldb must be multiples of 8 elements (TMA 16 B stride), got %lld %lld", lda, ldb);
  GGPT_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  return 0;
}

}  // namespace ggpt

using namespace ggpt;

extern "C" {

#ifdef GGPT_ATTN_TRACE
int ggpt_debug_gemm_trace(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, ggpt::g_gemm_trace, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
#endif

int ggpt_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                   void* C, long long ldc, int out_f32, int accumulate, int M, int N, int K, void* stream) {
  if (int rc = check_common(A, lda, B, ldb, M, N, K)) return rc;
  GGPT_REQUIRE(C != nullptr, "gemm: null output");
  GGPT_REQUIRE(ldc % 8 == 0 && ldc >= ((N + 7) / 8) * 8, "gemm: ldc=%lld must be a multiple of 8 and >= N rounded up to 8", ldc);
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.accumulate = accumulate;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool wide = N > 128;
  if (out_f32) {
    return wide ? dispatch_major<256, EPI_F32>(a_mn_major, b_mn_major, A, lda, B, ldb, p, s)
                : dispatch_major<128, EPI_F32>(a_mn_major, b_mn_major, A, lda, B, ldb, p, s);
  }
  GGPT_REQUIRE(!accumulate, "gemm: accumulate needs an fp32 output");
  p.direct_store = want_direct_store(C, ldc, N, 1);
  return wide ? dispatch_major<256, EPI_BF16>(a_mn_major, b_mn_major, A, lda, B, ldb, p, s)
              : dispatch_major<128, EPI_BF16>(a_mn_major, b_mn_major, A, lda, B, ldb, p, s);
}

int ggpt_gemm_split_plan(int M, int N, int K, int sm_count, int* num_splits, int* kb_per_split) {
  GGPT_REQUIRE(M > 0 && N > 0 && K > 0 && sm_count > 0, "gemm_split_plan: bad arguments M=%d N=%d K=%d sm_count=%d", M, N, K, sm_count);
  GGPT_REQUIRE(num_splits != nullptr && kb_per_split != nullptr, "gemm_split_plan: null output");
  const int bn = N > 128 ? 256 : 128;
  const int mb = (M + BM - 1) / BM;
  plan_split_k(mb, (N + bn - 1) / bn, (K + BK - 1) / BK, sm_count, mb >= 2, 3, num_splits, kb_per_split);
  return 0;
}

int ggpt_gemm_bf16_resid(const void* A, long long lda, const void* B, long long ldb, const float* resid,
                         long long ldr, const float* colscale, const float* rowscale, float* out, long long ldo,
                         int M, int N, int K, void* stream) {
  if (int rc = check_common(A, lda, B, ldb, M, N, K)) return rc;
  GGPT_REQUIRE(resid && out, "gemm_resid: null residual/output");
  GGPT_REQUIRE(N % 4 == 0 && ldr % 4 == 0 && ldo % 4 == 0, "gemm_resid: N, ldr, ldo must be multiples of 4");
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.C = out; p.ldc = ldo; p.resid = resid; p.ldr = ldr;
  p.colscale = colscale; p.rowscale = rowscale;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return N > 128 ? launch_gemm<256, false, false, EPI_RESID>(A, lda, B, ldb, p, s)
                 : launch_gemm<128, false, false, EPI_RESID>(A, lda, B, ldb, p, s);
}

int ggpt_gemm_bf16_geglu(const void* A, long long lda, const void* Wgu, long long ldb, void* gu, long long ldgu,
                         void* act, long long ldact, int M, int N2, int K, void* stream) {
  if (int rc = check_common(A, lda, Wgu, ldb, M, N2, K)) return rc;
  GGPT_REQUIRE(act != nullptr, "gemm_geglu: null act output");
  GGPT_REQUIRE(N2 % 16 == 0, "gemm_geglu: fused width must be a multiple of 16, got %d", N2);
  GGPT_REQUIRE(ldact % 8 == 0 && (gu == nullptr || ldgu % 8 == 0), "gemm_geglu: ld must be a multiple of 8");
  GemmParams p{};
  p.M = M; p.N = N2; p.K = K; p.C = gu; p.ldc = ldgu; p.C2 = act; p.ldc2 = ldact;
  p.direct_store = gu != nullptr && want_direct_store(gu, ldgu, N2 / 2, 1) && want_direct_store(act, ldact, N2 / 2, 1);
  return launch_gemm<256, false, false, EPI_GEGLU>(A, lda, Wgu, ldb, p, static_cast<cudaStream_t>(stream));
}

int ggpt_gemm_bf16_dgeglu(const void* dy, long long lda, const void* Wd, long long ldb, const void* gu, long long ldgu,
                          void* dgu, long long lddgu, int M, int I, int K, void* stream) {
  if (int rc = check_common(dy, lda, Wd, ldb, M, I, K)) return rc;
  GGPT_REQUIRE(gu && dgu, "gemm_dgeglu: null gu / dgu");
  GGPT_REQUIRE(I % 8 == 0 && ldgu % 8 == 0 && lddgu % 8 == 0 && ldgu >= 2 * I && lddgu >= 2 * I,
               "gemm_dgeglu: I and leading dimensions must be multiples of 8 and cover [gate | up]");
  GemmParams p{};
  p.M = M; p.N = I; p.K = K; p.C = dgu; p.ldc = lddgu; p.gu = gu; p.ldgu = ldgu;
  // TMA-staged factor tiles need whole 256-column accumulator tiles inside each half (I = 4 d, d a multiple of 64, always
  // satisfies this); GGPT_DGEGLU_TMA=0 selects the per-lane-load epilogue for A/B profiling
  static const char* tma_env = getenv("GGPT_DGEGLU_TMA");
  const bool use_tma = (I % 256 == 0) && !(tma_env != nullptr && tma_env[0] == '0') &&
                       (reinterpret_cast<uintptr_t>(gu) & 15) == 0 && (reinterpret_cast<uintptr_t>(dgu) & 15) == 0;
  if (use_tma) return launch_gemm<256, false, true, EPI_DGEGLU_TMA>(dy, lda, Wd, ldb, p, static_cast<cudaStream_t>(stream));
  return launch_gemm<256, false, true, EPI_DGEGLU>(dy, lda, Wd, ldb, p, static_cast<cudaStream_t>(stream));
}

int ggpt_gemm_bf16_qkv_rope(const void* A, long long lda, const void* Wqkv, long long ldb, void* qkv, long long ldc,
                            const int* pos, const float* cos_tab, const float* sin_tab, int rope_cols, int M, int N,
                            int K, void* stream) {
  if (int rc = check_common(A, lda, Wqkv, ldb, M, N, K)) return rc;
  GGPT_REQUIRE(qkv && pos && cos_tab && sin_tab, "gemm_qkv_rope: null pointer");
  GGPT_REQUIRE(N % 64 == 0 && rope_cols % 64 == 0 && ldc % 8 == 0, "gemm_qkv_rope: N and rope_cols must be multiples of 64 (head_dim)");
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.C = qkv; p.ldc = ldc; p.pos = pos; p.cos_tab = cos_tab; p.sin_tab = sin_tab;
  p.rope_cols = rope_cols;
  GGPT_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 31) == 0 && ldc % 16 == 0,
               "gemm_qkv_rope: the output must be 32-byte aligned with a row pitch that is a multiple of 16 elements");
  p.direct_store = 1;
  return launch_gemm<256, false, false, EPI_QKV_ROPE>(A, lda, Wqkv, ldb, p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
