// Shared device/host helpers for the sm_100a kernels of the GraphGPT hot path.
// Everything here is raw PTX for Blackwell (tcgen05 / TMEM / TMA / mbarrier); no CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace ggpt {

// ---------------------------------------------------------------------------------------------
// Host-side error plumbing (C-ABI returns 0 / negative code, text via ggpt_last_error)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  check_launch(const char* what);   // cudaGetLastError -> 0 / -2

#define GGPT_REQUIRE(cond, ...)                         \
  do {                                                  \
    if (!(cond)) {                                      \
      ::ggpt::set_error(__VA_ARGS__);                   \
      return -1;                                        \
    }                                                   \
  } while (0)

// Encode a 2-D bf16 row-major tensor [rows, cols] (row pitch `ld` elements) as a TMA map whose box is
// box_rows x box_cols with 128-byte swizzle (box_cols * 2 bytes must be 128).
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);
// Same for a 3-D view [d2, rows, cols] with element strides (ld2, ld, 1); box = 1 x box_rows x box_cols.
int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d2, uint64_t rows, uint64_t cols,
                      uint64_t ld2, uint64_t ld, uint32_t box_rows, uint32_t box_cols);

// fp32 [d2, rows, cols] view, box = 1 x box_rows x box_cols, swizzle = the box row size (128 B / 64 B, else none): TMA
// reduce-add target.
int make_tmap_3d_f32(CUtensorMap* map, const void* base, uint64_t d2, uint64_t rows, uint64_t cols,
                     uint64_t ld2, uint64_t ld, uint32_t box_rows, uint32_t box_cols);

int num_sms();

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a descriptor / protocol bug must not hang the GPU box. On timeout the kernel traps.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef GGPT_UNBOUNDED_WAIT
  while (!mbar_try_wait(bar, parity)) {
  }
#else
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) {
      printf("ggpt: mbarrier wait timeout block=(%d,%d,%d) thread=%d bar=%p parity=%u\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, (void*)bar, parity);
      __trap();
    }
  }
#endif
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMA ------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}

// TMA store shared -> global through a tensor map (SASS UTMASTG), bulk-group completion.  The smem tile must have been
// written with the map's swizzle and made visible to the async proxy (fence.proxy.async) before the store is issued.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// TMA reduction shared -> global (element-wise fp32 add into the tensor, performed at L2): same completion mechanism as
// the TMA store.  Out-of-range rows / columns of the box are clipped.
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's committed store groups have not finished READING their smem source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 1-D bulk copy global -> shared (TMA engine, no tensor map): 16-byte aligned addresses, size a multiple of 16.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- thread-block clusters -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load multicast to the same smem offset (and the same mbarrier offset) in every CTA of `mask`.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// tcgen05.commit that arrives on the mbarrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// ---- CTA pairs (cta_group::2): one MMA spans the two SMs of a TPC -----------------------------------------
// shared::cluster addresses of the two CTAs of a pair differ in bit 24; clearing it names the leader's (rank 0) copy.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
// TMA load issued by either CTA of a pair into ITS OWN smem, completing on the LEADER's mbarrier (same offset).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the mbarrier at this offset in CTA `cta` of the cluster.  Default semantics (release at CTA scope), as in
// CUTLASS's ClusterBarrier::arrive: the barrier only hands a drained TMEM accumulator back to the MMA thread, and the
// tcgen05 fences order the TMEM reads — a `.release.cluster` arrive would additionally wait until every global store of
// the epilogue warp is visible cluster-wide (ncu: "membar" was the second largest stall of the pair-mode GEMMs).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 (128 rows from each CTA's smem) and each CTA supplying half of B's N rows.
// Issued by ONE thread of the leader CTA; descriptors are smem offsets valid in both CTAs.
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of pair MMAs: arrives on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_pair_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar) & kPeerBitMask),
               "h"(mask)
               : "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: arrive on `bar` once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate. Issued by ONE thread.
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-uniform issue: the whole warp executes the surrounding (uniform) control flow and address arithmetic — which then
// lives in uniform registers — and ONE elected lane executes the tcgen05 instruction.  Issued from divergent single-lane
// code every MMA costs ~20 instructions (per-lane descriptor arithmetic + an R2UR / ELECT / branch waterfall).
__device__ __forceinline__ void tc_mma_bf16_e(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) tc_mma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
}
__device__ __forceinline__ void tc_commit_e(uint64_t* bar) {
  if (elect_one()) tc_commit(bar);
  __syncwarp();
}
// non-blocking barrier test by lane 0, result broadcast: keeps a polling loop warp-uniform
__device__ __forceinline__ bool mbar_try_wait_warp(uint64_t* bar, uint32_t parity) {
  int ok = 0;
  if ((threadIdx.x & 31) == 0) ok = mbar_try_wait(bar, parity) ? 1 : 0;
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (base_lane + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// registers -> TMEM, same 32 lanes x 32 columns shape (used to rescale an accumulator in place)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 8-column variants (32 lanes x 8 columns): for rarely taken paths that must not cost 32 live registers
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors -------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor, 128-byte swizzle (layout_type 2), descriptor version 1.
//   [0,14)  start address >> 4        [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4   [46,48) version = 1      [61,64) layout type
// K-major operand (rows = M/N index, 128 B = 64 bf16 along K): 8-row groups are 1024 B apart (SBO),
//   LBO is unused for swizzled K-major (encoded 1).  One MMA (K=16) advances the start by 32 B.
// MN-major operand (rows = K index, 128 B = 64 bf16 along M/N): 8 K-rows are 1024 B apart (SBO),
//   successive 64-element M/N blocks are `lbo_bytes` apart (LBO).  One MMA advances the start by 16 rows.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// 32-bit instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate.
//   [4,6) c_format=1 (f32)  [7,10) a_format=1 (bf16)  [10,13) b_format=1 (bf16)
//   [15] a_major  [16] b_major  (0 = K-major, 1 = MN-major)   [17,23) N>>3   [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// Byte offset of element (row, col) inside a [rows x 64] bf16 tile stored with the 128-byte swizzle that
// TMA SWIZZLE_128B and the UMMA SW128 descriptors use (tile base must be 1024-byte aligned).
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t col) {
  uint32_t chunk = (col >> 3) ^ (row & 7);  // 16-byte chunk index XOR row-in-group
  return row * 128u + chunk * 16u + (col & 7) * 2u;
}

// ---- small math helpers ---------------------------------------------------------------------
// 2^x on the MUFU pipe, no range fix-up code (arguments here are <= 0 up to rounding; tiny results flush to 0)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t v) {
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(b);
}
// Every lane receives the 32 floats of table row `row_idx` (its own row; e.g. the RoPE cos/sin row of its token).
// A per-lane read would touch 32 different 128-byte lines per instruction (32 L1 wavefronts for 16 useful bytes each),
// so rows are fetched cooperatively — 8 lanes per row, 4 rows per 16-byte/lane instruction — and transposed through a
// 4 KB per-warp smem buffer with the 128-byte XOR swizzle (conflict-free both ways).
__device__ __forceinline__ void warp_gather_rows32(const float* __restrict__ tab, int row_idx, uint8_t* stg, int lane,
                                                   float (&out)[32]) {
  const int rd_row = lane >> 3, rd_j = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + rd_row;
    const int src = __shfl_sync(0xffffffffu, row_idx, rr);
    const float4 v = *reinterpret_cast<const float4*>(tab + static_cast<long long>(src) * 32 + rd_j * 4);
    *reinterpret_cast<float4*>(stg + rr * 128 + ((rd_j ^ (rr & 7)) << 4)) = v;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4));
    out[j * 4 + 0] = v.x; out[j * 4 + 1] = v.y; out[j * 4 + 2] = v.z; out[j * 4 + 3] = v.w;
  }
  __syncwarp();
}

// Coalesced store of a warp's 32 rows x 128 bytes (64 bf16) that are held one row per lane: the rows are transposed
// through a 4 KB swizzled smem buffer so that each store instruction writes 4 complete 128-byte lines (8 lanes per row)
// instead of 32 scattered 16-byte pieces (32 LSU wavefronts and half-used sectors per instruction).
//   put: lane writes 16-byte chunk j (0..7) of its row;   flush: rows [0, rows_valid) go to dst_row0 + row * ld.
__device__ __forceinline__ void stage_put16(uint8_t* stg, int lane, int j, uint4 v) {
  *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = v;
}
__device__ __forceinline__ void stage_flush_rows(uint8_t* stg, int lane, __nv_bfloat16* dst_row0, long long ld, int rows_valid) {
  __syncwarp();
  const int rd_row = lane >> 3, rd_j = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + rd_row;
    if (rr < rows_valid)
      *reinterpret_cast<uint4*>(dst_row0 + static_cast<long long>(rr) * ld + rd_j * 8) =
          *reinterpret_cast<const uint4*>(stg + rr * 128 + ((rd_j ^ (rr & 7)) << 4));
  }
  __syncwarp();
}

// Same, for a 32-row block filled by TWO warps (each wrote half of every row's chunks): this warp flushes the 16 rows
// [row_first, row_first + 16) of the block; dst_row0 is the global address of row `row_first`.
__device__ __forceinline__ void stage_flush_rows16(uint8_t* stg_block, int lane, int row_first, __nv_bfloat16* dst_row0,
                                                   long long ld, int rows_valid) {
  const int rd_row = lane >> 3, rd_j = lane & 7;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int k = it * 4 + rd_row;
    const int rr = row_first + k;
    if (k < rows_valid)
      *reinterpret_cast<uint4*>(dst_row0 + static_cast<long long>(k) * ld + rd_j * 8) =
          *reinterpret_cast<const uint4*>(stg_block + rr * 128 + ((rd_j ^ (rr & 7)) << 4));
  }
}

// Counter-based dropout mask for attention probabilities: keep(n,h,q,k) is a pure function of (seed, n, h, q, k) and of
// the row-tile plan, so the forward and the backward kernels regenerate identical masks with no stored state
// (ref: HF:217 attention dropout).  One 32-bit draw serves a PAIR of adjacent keys of a key tile (16 bits each:
// the keep probability is quantised to 1/65536 and inv_keep is derived from the quantised value): two rounds of a
// wide multiply (Philox-style hi ^ lo folding) keyed by a well-mixed per-row key — 5 integer instructions per pair.
__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t drop_rowkey(uint32_t seed_lo, uint32_t seed_hi, int n, int h, int q) {
  uint32_t x = fmix32(seed_lo ^ (static_cast<uint32_t>(n) * 0x9E3779B1u));
  x = fmix32(x ^ seed_hi ^ (static_cast<uint32_t>(h) * 0x85EBCA77u));
  return fmix32(x ^ (static_cast<uint32_t>(q) * 0xC2B2AE3Du));
}
__device__ __forceinline__ uint32_t drop_rowkey2(uint32_t rowkey) { return fmix32(rowkey ^ 0x68BC21EBu); }
// 32 random bits for the key pair whose first (even tile-local) key has absolute index k_even
__device__ __forceinline__ uint32_t drop_bits(uint32_t rk1, uint32_t rk2, int k_even) {
  const uint64_t m1 = static_cast<uint64_t>(static_cast<uint32_t>(k_even) ^ rk1) * 0xD2511F53ull;
  const uint32_t y = static_cast<uint32_t>(m1 >> 32) ^ static_cast<uint32_t>(m1) ^ rk2;
  const uint64_t m2 = static_cast<uint64_t>(y) * 0xCD9E8D57ull;
  return static_cast<uint32_t>(m2 >> 32) ^ static_cast<uint32_t>(m2);
}
// thresh = (16-bit drop threshold) << 16;  even key of the pair uses the low half of the draw, odd key the high half
__device__ __forceinline__ bool drop_keep_even(uint32_t bits, uint32_t thresh) { return (bits << 16) >= thresh; }
__device__ __forceinline__ bool drop_keep_odd(uint32_t bits, uint32_t thresh) { return bits >= thresh; }
#endif  // __CUDACC__

struct DropParams {
  uint32_t thresh;      // (round(p * 65536) << 16); 0 = dropout off
  uint32_t seed_lo, seed_hi;
  float inv_keep;       // 1 / (1 - p_quantised)
};

inline DropParams make_drop_params(float p, unsigned long long seed) {
  DropParams d;
  d.thresh = 0; d.seed_lo = static_cast<uint32_t>(seed); d.seed_hi = static_cast<uint32_t>(seed >> 32); d.inv_keep = 1.0f;
  if (p > 0.f) {
    long t = static_cast<long>(static_cast<double>(p) * 65536.0 + 0.5);
    if (t > 65535) t = 65535;
    d.thresh = static_cast<uint32_t>(t) << 16;
    d.inv_keep = static_cast<float>(1.0 / (1.0 - static_cast<double>(t) / 65536.0));
  }
  return d;
}
#ifdef __CUDACC__

// Element dropout (nn.Dropout on activations: embed_pdrop / mlp_pdrop, ref: modeling_helpers.py:97-98,
// utils_graphgpt.py:69-83): keep(e) is a pure function of (seed, e) with e the linear element index of the tensor, so
// forward and backward regenerate the same mask.  One 32-bit draw per PAIR of adjacent elements (16 bits each).
__device__ __forceinline__ uint32_t edrop_bits(const DropParams& dp, unsigned long long pair) {
  const uint32_t x = fmix32(static_cast<uint32_t>(pair) ^ dp.seed_lo);
  return fmix32(x ^ dp.seed_hi ^ (static_cast<uint32_t>(pair >> 32) * 0x9E3779B1u));
}
// keep/(1-p) scales of elements 2*pair and 2*pair+1
__device__ __forceinline__ float2 edrop_scale2(const DropParams& dp, unsigned long long pair) {
  const uint32_t bits = edrop_bits(dp, pair);
  return make_float2(drop_keep_even(bits, dp.thresh) ? dp.inv_keep : 0.f, drop_keep_odd(bits, dp.thresh) ? dp.inv_keep : 0.f);
}
__device__ __forceinline__ float edrop_scale1(const DropParams& dp, unsigned long long e) {
  const float2 s = edrop_scale2(dp, e >> 1);
  return (e & 1ull) ? s.y : s.x;
}

// erf GELU (transformers ACT2FN["gelu"]) and its derivative from ONE exponential:
//   Phi(g) = 0.5 (1 + erf(g / sqrt2)),  erf(x) = sign(x) (1 - poly(t) e^{-x^2}),  t = 1 / (1 + 0.3275911 |x|)
//   (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7 — far below the bf16 rounding of the outputs), and with
//   x = g / sqrt2 the same e^{-x^2} = e^{-g^2/2} is the Gaussian pdf needed by gelu'(g) = Phi(g) + g phi(g).
__device__ __forceinline__ void gelu_cdf_pdf(float g, float& cdf, float& pdf) {
  // (t = 1 / (1 + 0.3275911 |g| / sqrt2): the constant is folded; rcp / ex2 are the bare MUFU forms — exp2f / __fdividef
  // wrap them in range fix-ups (FSETP + 2 FMUL each) that an argument <= 0 / >= 1 never needs)
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, fabsf(g), 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(g * g * -0.72134752044448170f));   // e^{-g^2/2}
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float erf_abs = fmaf(-poly, e, 1.0f);
  cdf = 0.5f + 0.5f * copysignf(erf_abs, g);
  pdf = 0.3989422804014327f * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, pdf;
  gelu_cdf_pdf(x, cdf, pdf);
  return x * cdf;
}
// Forward-only erf GELU with ONE MUFU op (the GEMM epilogue is MUFU-limited): Abramowitz-Stegun 7.1.28,
//   erf(x) = 1 - 1 / (1 + a1 x + ... + a6 x^6)^16  (x >= 0, |error| <= 3e-7);  |gelu error| <= 8.2e-7 over [-12, 12]
//   (checked against scipy in fp32 emulation), far below the bf16 rounding of the output.  Overflow of the 16th power
//   for |x| > ~17 gives 1/inf = 0, i.e. Phi = 0 or 1 exactly.
__device__ __forceinline__ float gelu_erf_fwd(float g) {
  const float ax = fabsf(g) * 0.70710678118654752f;
  float p = fmaf(0.0000430638f, ax, 0.0002765672f);
  p = fmaf(p, ax, 0.0001520143f);
  p = fmaf(p, ax, 0.0092705272f);
  p = fmaf(p, ax, 0.0422820123f);
  p = fmaf(p, ax, 0.0705230784f);
  p = fmaf(p, ax, 1.0f);
  p *= p; p *= p; p *= p; p *= p;
  const float r = __fdividef(0.5f, p);          // 0.5 * (1 - erf|x|)
  return g * (g >= 0.f ? 1.0f - r : r);
}
#endif  // __CUDACC__

}  // namespace ggpt
