// Parameters / launchers of the isolated-diagonal-tile attention kernels (attn_diag_sm100.cu).
#pragma once
#include "common.cuh"

namespace ggpt {

struct DiagParams {
  int N, S, H;
  int max_tiles;
  int mask_words;
  const uint32_t* mask_bits;   // [N,S,words]
  const int* tile_start;       // [N,max_tiles+1]
  const uint8_t* tile_cls;     // [N,max_tiles,max_tiles]
  const int* iso_list;         // [count] 16-byte descriptors {sequence, first row, rows, class} (16-byte aligned)
  const int* iso_count;        // [1]
  int q_col0, k_col0, v_col0;
  float scale, scale_log2;
  DropParams drop;             // attention-probability dropout (training)
  // forward
  __nv_bfloat16* out;          // [N*S, H*64]
  long long ldo;
  float* lse;                  // [N,H,S]
  // backward
  const float* lse_in;
  const float* dsum;           // [N,H,S]
  __nv_bfloat16* dqkv;
  long long ld_dqkv;
  const int* pos;
  const float* cos_tab;
  const float* sin_tab;
};

// iso_flags[N*max_tiles] = 1 for tiles whose only active pair (row and column of tile_cls) is the diagonal one;
// iso_list / iso_count = compact device-side work list of those tiles.
int attn_iso_build(const uint8_t* cls, const int* n_tiles, const int* tile_start, int N, int max_tiles, uint8_t* iso_flags,
                   int* iso_list, int* iso_count, cudaStream_t s);
#ifdef GGPT_ATTN_TRACE
int diag_trace_read(long long* out, int n);
int bwd_trace_read(long long* out, int n);
#endif
int attn_diag_fwd_launch(const CUtensorMap& tm, const DiagParams& p, cudaStream_t s);
int attn_diag_bwd_launch(const CUtensorMap& tmQKV, const CUtensorMap& tmDO, const DiagParams& p, cudaStream_t s);

}  // namespace ggpt
