/* ggpt_b200.h — C ABI of the B200-native GraphGPT hot path (libggpt_b200.so).
 *
 * The reference (alibaba/graph-gpt, pure Python over torch + HF transformers) has no native boundary; its hot
 * path is GraphGPTPretrainBase.forward / GraphGPTTaskModel.forward -> LlamaModel.forward -> torch kernels.
 * This library replaces those torch/cuBLAS/ATen calls one fused op at a time.  Each entry point cites the
 * reference lines whose arithmetic it reproduces ("ref:" = /root/reference, "HF:" = transformers
 * models/llama/modeling_llama.py, the un-vendored third-party dependency pinned at 4.53.3).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - the caller owns all buffers; no entry point allocates, frees or synchronises
 *   - `stream` is a cudaStream_t passed as void*
 *   - returns 0 on success, <0 on error; ggpt_last_error() returns the message (thread-local)
 *   - bf16 tensors are row-major with an explicit leading dimension in ELEMENTS; TMA needs ld % 8 == 0 and
 *     16-byte aligned bases
 *   - sm_100a only; there is no CPU or other-arch fallback
 */
#ifndef GGPT_B200_H_
#define GGPT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGPT_ABI_VERSION 2

const char* ggpt_last_error(void);
int ggpt_abi_version(void);
int ggpt_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * tcgen05 GEMM: C[M,N] = A[M,K] * B[N,K]^T, bf16 in, fp32 accumulate (TMEM).
 * a_mn_major / b_mn_major = 0: operand stored [rows = M or N, cols = K] (K contiguous);
 *                         = 1: operand stored [rows = K, cols = M or N] (M/N contiguous) — used so that
 *   dgrad (dx = dy W) and wgrad (dW = dy^T x) read weights/activations in place, without transposes.
 * out_f32 = 0: C is bf16;  1: C is fp32 and `accumulate` != 0 adds into C (gradient accumulation).
 * ref: every nn.Linear on the path — HF:262-264,288 (q/k/v/o_proj), HF:182-184 (gate/up/down_proj),
 *      modeling_pretrain.py:89-93 (n_token_proj), :218 (lm_head), and their autograd backward.
 * ------------------------------------------------------------------------------------------- */
int ggpt_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                   void* C, long long ldc, int out_f32, int accumulate, int M, int N, int K, void* stream);

/* Host-only query (no GPU needed): the split-K plan ggpt_gemm_bf16 uses for an accumulating fp32 output of this shape
 * (weight gradients, K = tokens) on a device with `sm_count` SMs.  The k range is cut into `num_splits` work items of
 * `kb_per_split` 64-wide k-blocks each (the last one may be shorter); every split keeps >= 16 k-blocks.  Chosen so that
 * tiles x splits fills whole waves of resident CTAs (CTA pairs when M spans >= 2 row tiles). */
int ggpt_gemm_split_plan(int M, int N, int K, int sm_count, int* num_splits, int* kb_per_split);

/* out[M,N] (fp32) = resid[M,N] (fp32) + rowscale[m] * colscale[n] * (A[M,K] * B[N,K]^T)
 * colscale (LayerScale lambda_1/lambda_2) and rowscale (DropPath keep-mask / keep-prob) may be NULL.
 * ref: HF:325,331 (residual adds after o_proj / down_proj); utils_graphgpt.py:153-166 (lambda, drop_path). */
int ggpt_gemm_bf16_resid(const void* A, long long lda, const void* B, long long ldb, const float* resid,
                         long long ldr, const float* colscale, const float* rowscale, float* out, long long ldo,
                         int M, int N, int K, void* stream);

/* Fused gate|up projection with GeGLU epilogue.  Wgu is [2I, K]: rows [0,I) = gate_proj.weight, rows [I,2I) =
 * up_proj.weight.  Writes act[M,I] = gelu_erf(gate) * up (bf16) and, when gf != NULL (training), the two factors GeGLU
 * backward multiplies d(act) with: gf[M,2I] = [ up * gelu'(gate) | gelu(gate) ] (bf16) — gate and up themselves are not
 * needed again.   ref: HF:182-184 with hidden_act="gelu" (exact erf GELU). */
int ggpt_gemm_bf16_geglu(const void* A, long long lda, const void* Wgu, long long ldb, void* gf, long long ldgf,
                         void* act, long long ldact, int M, int N2, int K, void* stream);

/* Fused down_proj dgrad + GeGLU backward: dact = dy[M,K] * Wd[K,I] (Wd = down_proj.weight [d, I] read in place) stays
 * in TMEM and the epilogue writes dgu[M,2I] = [ dact * gf[:,0:I] | dact * gf[:,I:2I] ] from the saved factors.
 * ref: autograd of HF:182-184 (down_proj, erf GELU, gate*up). */
int ggpt_gemm_bf16_dgeglu(const void* dy, long long lda, const void* Wd, long long ldb, const void* gf, long long ldgf,
                          void* dgu, long long lddgu, int M, int I, int K, void* stream);

/* Fused q|k|v projection with rotary embedding applied to the first rope_cols columns (q and k heads of 64).
 * pos[M] = position id of each row; cos_tab/sin_tab = fp32 [max_pos, 32] (HF inv_freq table).
 * ref: HF:262-267 (projections + apply_rotary_pos_emb), HF:124-135,138-168 (cos/sin, rotate_half). */
int ggpt_gemm_bf16_qkv_rope(const void* A, long long lda, const void* Wqkv, long long ldb, void* qkv, long long ldc,
                            const int* pos, const float* cos_tab, const float* sin_tab, int rope_cols, int M, int N,
                            int K, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Attention.  qkv is the fused projection output [N*S, ld_qkv] (bf16); q/k/v of head h live at columns
 * {q,k,v}_col0 + 64*h.  The mask is a bit matrix built once per step.
 * ------------------------------------------------------------------------------------------- */
/* uint32 words per mask row for sequence length S (one uint4 per 128 keys + 4 pad words so unaligned 128-key
 * windows can be read), and the allocation stride of the per-sequence tile arrays (= ceil(S/64)). */
int ggpt_attn_mask_words(int S);
int ggpt_attn_max_tiles(int S);

/* attention_mask: int64 [N,S] if mask_dims == 2, [N,S,S] if 3, or NULL (all visible); causal != 0 additionally hides
 * keys k > q.  An [N,S] mask of 0/1 values is the reference's key-padding mask (every query sees the non-zero keys); a
 * sequence whose [N,S] mask holds values >= 2 carries the SEGMENT IDS written by ggpt_pack_sequences (1-based, 0 = pad)
 * and expands to the block-diagonal mask — query q sees key k iff both lie in the same non-zero segment.  Outputs:
 *   mask_bits  [N,S,words]            1 bit per (q,k)
 *   tile_start [N,max_tiles+1], n_tiles [N]   variable row tiles (<= 128 rows, cut on block boundaries of the mask so
 *                                     packed segments never straddle a tile; uniform 128 grid otherwise)
 *   tile_cls   [N,max_tiles,max_tiles] class of every (query tile, key tile): 0 masked, 1 fully visible, 2 mixed
 *   iso_flags  [N,max_tiles], iso_list [4*N*max_tiles] (16-byte aligned), iso_count [2]   tiles whose only active pair
 *                                     is their own diagonal one (every tile of a packed batch): flags + compact work
 *                                     list (one {sequence, first row, rows, class} descriptor of 4 ints per tile) for
 *                                     the persistent single-pass kernels (iso_count[0] = #isolated, [1] = #tiles);
 *                                     pass iso_flags = NULL to ggpt_attn_fwd/bwd to force the general tile-loop
 *                                     kernels, or run_general = 0 when the caller knows every tile is isolated
 * ref: modeling_helpers.py:38-64 (_update_causal_mask, _expand_mask_from_3d_mask: additive 0/finfo.min mask),
 *      HF:398-405 (causal mask when config.causal_attention); packing: tokenizer_utils.py:351-355.  Fully masked
 *      query rows yield 0 here (the reference yields a uniform average; such rows are padding, never consumed). */
int ggpt_attn_mask_build(const long long* attention_mask, int mask_dims, int N, int S, int causal, uint32_t* mask_bits,
                         int* tile_start, int* n_tiles, uint8_t* tile_cls, uint8_t* iso_flags, int* iso_list,
                         int* iso_count, void* stream);

/* out[N*S, H*64] = dropout(softmax(q k^T / 8 + mask)) v per head; lse[N,H,S] (may be NULL) = log-sum-exp of the scaled
 * scores, kept for the backward pass.  out_lo (may be NULL; same shape / ld as out) receives the bf16 rounding residual
 * bf16(O - bf16(O)) of the rows computed by the general tile-loop kernel (zeros for isolated tiles): ggpt_attn_bwd uses
 * it to form D = dO.(O + O_lo) to fp32 accuracy.  dropout_p > 0 (training, config.attention_dropout) drops attention
 * probabilities with a counter-based mask that is a pure function of (seed, n, h, q, k); pass the same (dropout_p, seed)
 * to ggpt_attn_bwd.   ref: HF:199-221 (eager_attention_forward: fp32 softmax, dropout on the weights). */
int ggpt_attn_fwd(const void* qkv, long long ld_qkv, int q_col0, int k_col0, int v_col0, const uint32_t* mask_bits,
                  const int* tile_start, const int* n_tiles, const uint8_t* tile_cls, const uint8_t* iso_flags,
                  const int* iso_list, const int* iso_count, int run_general, float dropout_p, unsigned long long seed,
                  void* out, void* out_lo, long long ldo, float* lse, int N, int S, int H, void* stream);

/* dqkv[N*S, ld_dqkv] (bf16) = gradient of the fused q|k|v projection output given dout = dL/d(attention output).
 * Recomputes P from lse in ONE tcgen05 pass over the score tiles (per key tile: dK, dV in TMEM; the dQ contributions of
 * the key tiles are summed with TMA reductions into dq_acc and converted by a finishing kernel); dQ/dK are un-rotated
 * (inverse RoPE) with the same pos / cos / sin tables as ggpt_gemm_bf16_qkv_rope.  out_lo: see ggpt_attn_fwd (may be
 * NULL).  Scratch: dsum_scratch fp32 [N,H,S]; dq_acc fp32 [N*S, H*64], ZERO on entry and zero again on return (needed
 * only when run_general != 0).  ref: autograd of HF:199-221 and HF:146-168. */
int ggpt_attn_bwd(const void* qkv, long long ld_qkv, int q_col0, int k_col0, int v_col0, const void* out, const void* out_lo,
                  long long ldo, const void* dout, long long lddo, const float* lse, const uint32_t* mask_bits,
                  const int* tile_start, const int* n_tiles, const uint8_t* tile_cls, const uint8_t* iso_flags,
                  const int* iso_list, const int* iso_count, int run_general, float dropout_p, unsigned long long seed,
                  const int* pos, const float* cos_tab, const float* sin_tab, float* dsum_scratch, float* dq_acc, void* dqkv,
                  long long ld_dqkv, int N, int S, int H, void* stream);

/* ---------------------------------------------------------------------------------------------
 * HBM-bound kernels
 * ------------------------------------------------------------------------------------------- */
/* x[t,:] = sum_f gate[f,:] * drop(table[ids[t,f],:]) (gate NULL -> plain sum); long_scale applies the
 * stack_method=="long" 1/nnz rescale.  err_flag (may be NULL) is set to 1 on an out-of-range id.
 * drop_p > 0 (training, config.embed_pdrop): element dropout on the gathered [T,F,d] rows before aggregation, mask =
 * the counter-based one of ggpt_dropout_scale_f32 over the linear index (t*F+f)*d+j.
 * ref: modeling_helpers.py:89-114 (embed_dropout :97-98), modeling_common.py:127-135. */
int ggpt_embed_fwd(const long long* ids, const float* table, const float* gate, float* out, long long T, int F, int d,
                   int V, int long_scale, int* err_flag, float drop_p, unsigned long long drop_seed, void* stream);
/* dtable[id,:] += gate[f,:]*dx[t,:] for id != padding_idx (dtable pre-zeroed / accumulating), dgate likewise.
 * ref: autograd of the above; nn.Embedding(padding_idx=0) HF:361. */
int ggpt_embed_bwd(const long long* ids, const float* dx, const float* table, const float* gate, float* dtable,
                   float* dgate, long long T, int F, int d, int V, int padding_idx, int long_scale, float drop_p,
                   unsigned long long drop_seed, void* stream);

/* cnt[t, v] (bf16 [T, ldc], fully overwritten) = number of features f with ids[t,f] == v, 0 for v == padding_idx.
 * The embedding gradient is then the wgrad GEMM dE = cnt^T dX (ggpt_gemm_bf16 with both operands MN-major),
 * replacing T*F*d contended atomics.   ref: autograd of modeling_helpers.py:96 / modeling_common.py:134. */
int ggpt_embed_count(const long long* ids, void* cnt, long long ldc, long long T, int F, int V, int padding_idx,
                     void* stream);

/* In-model SMTP masking (config.smtp_inside).  ids int64 [N,S,Ftot]: columns [0,F) are the stacked token ids, column
 * node_col (= F+2 in the reference's layout) is the row's node index.  mr f32 [N] and u_node f32 [N,S,F] are uniform
 * draws supplied by the caller (torch's generator, so the stream of random numbers stays the reference's).
 * masked = u_node[n, node_idx[n,s], f] > mr[n]^power && ids[n,s,f] > 0;  out_ids = masked ? mask_token : id;
 * labels = masked ? id : label_pad.  A node index outside [-S, S) sets *err_flag = 3 (may be NULL).
 * ref: modeling_helpers.py:399-452 (prepare_for_2d_smtp_inputs_labels), modeling_pretrain.py:175-189. */
int ggpt_smtp_mask_2d(const long long* ids, int Ftot, int node_col, const float* mr, const float* u_node, float power,
                      long long* out_ids, long long* labels, int N, int S, int F, long long mask_token,
                      long long label_pad, int* err_flag, void* stream);

/* Raw-embedding input branch (config.embed_dim = E > 0).  h[t,:] (bf16 [T, ldh]) = RMSNorm_E(src_t; w) with
 * src_t = keep_t ? raw[t,:] : mask_tok[:], keep_t = any_{f<fchk}(labels[t*ldl+f] == -100); labels NULL (fine-tuning:
 * no mask-token swap) keeps every row.  rstd[T] / keep[T] (u8) are saved for backward (either may be NULL).  The
 * result feeds embed_proj (ggpt_gemm_bf16) and is added to the token-embedding sum by ggpt_add_rmsnorm_fwd.
 * ref: modeling_pretrain.py:119-150 (pre-training: fchk = F, or 1 when smtp_inside), modeling_helpers.py:127-139 (FT). */
int ggpt_raw_embed_norm_fwd(const float* raw, const long long* labels, long long ldl, int fchk, const float* mask_tok,
                            const float* w, void* h, long long ldh, float* rstd, unsigned char* keep, long long T, int E,
                            float eps, void* stream);
/* Parameter gradients of the above from dh (bf16 [T, lddh] = d loss / d h): dw[E] += sum_t dh*xhat; dmask_tok[E]
 * (may be NULL) += sum over swapped rows of dRMSNorm(dh).  The raw features are data and receive no gradient. */
int ggpt_raw_embed_norm_bwd(const void* dh, long long lddh, const float* raw, const unsigned char* keep,
                            const float* mask_tok, const float* rstd, const float* w, float* dw, float* dmask_tok,
                            long long T, int E, void* stream);

/* Backward of a LayerScale / DropPath residual branch x_out = x_in + rowscale[t]*lam[c]*y (ppa fine-tuning: lsi = 1,
 * path_dropout = 0.2): dy (bf16 [T,d]) = dx*lam*rowscale, the gradient handed to the branch's GEMMs, and
 * dlam[c] += sum_t dx*(x_out - x_in)/lam (y itself is not kept).  lam / rowscale may be NULL (factor 1); dlam NULL skips
 * the reduction (then x_out / x_in may be NULL).   ref: autograd of utils_graphgpt.py:153-166. */
int ggpt_layerscale_bwd(const float* dx, const float* x_out, const float* x_in, const float* lam, const float* rowscale,
                        void* dy, float* dlam, long long T, int d, void* stream);

/* Device-side sequence packer.  rows int64 [R,F] = the token rows of a pool of graphs, graph g owning rows
 * [cu_rows[g], cu_rows[g+1]); packed sequence n = graphs seq_graphs[cu_seq[n] .. cu_seq[n+1]) in order, each followed by
 * one separator row sep_row[F] (<eos>), truncated to S rows, padded with pad_id (at most 4096 graphs per sequence).
 * Writes input_ids [N,S,F], seg_ids int64 [N,S] (1-based segment per row, 0 = pad: pass it as the [N,S] attention_mask —
 * ggpt_attn_mask_build expands segment ids to the block-diagonal mask), position_ids [N,S] = 0..S-1 (may be NULL) and
 * n_valid[N] non-pad rows (may be NULL).
 * ref: tokenizer.py:359-415 (pack_token_seq), tokenizer_utils.py:228-233,351-355 (final <eos>, block_diag mask). */
int ggpt_pack_sequences(const long long* rows, int F, const int* cu_rows, const int* seq_graphs, const int* cu_seq, int N,
                        int S, const long long* sep_row, long long pad_id, long long* input_ids, long long* seg_ids,
                        long long* position_ids, int* n_valid, void* stream);

/* Element dropout on activations, training only (config.mlp_pdrop: GeGLU output and down_proj output; embed_pdrop:
 * normalised raw embeddings).  x (bf16, n contiguous elements, 16-byte aligned) is scaled in place by keep(e)/(1-p);
 * keep(e) is a pure function of (seed, e) — the backward pass applies the same call to the incoming gradient.
 * p is quantised to 1/65536.  ggpt_dropout_scale_f32 writes the factors keep(e)/(1-p) for e in [0, n).
 * ref: utils_graphgpt.py:69-83 (mlp_act_dropout, mlp_dropout), modeling_pretrain.py:146-147 (raw_embed_dropout). */
int ggpt_dropout_bf16(void* x, long long n, float p, unsigned long long seed, void* stream);
int ggpt_dropout_scale_f32(float* out, long long n, float p, unsigned long long seed, void* stream);

/* y = bf16(w * x * rsqrt(mean(x^2)+eps)), rstd[T] saved for backward (may be NULL).   ref: HF:59-64 */
int ggpt_rmsnorm_fwd(const float* x, const float* w, void* y, long long ldy, float* rstd, long long T, int d, float eps,
                     void* stream);
/* Fused residual add + RMSNorm: x_out = x_in + rowscale[t]*colscale[:]*y (y = bf16 output of o_proj / down_proj;
 * colscale = LayerScale lambda, rowscale = DropPath scale, either may be NULL), h = bf16(w * x_out * rstd) (h may be
 * NULL when only the sum is wanted), rstd[T] saved for backward (may be NULL).
 * ref: HF:325,331 (residual adds; utils_graphgpt.py:153-166 for lambda / drop_path) followed by HF:59-64. */
int ggpt_add_rmsnorm_fwd(const float* x_in, const void* y, long long ldy, const float* colscale, const float* rowscale,
                         const float* w, float* x_out, void* h, float* rstd, long long T, int d, float eps, void* stream);
/* dx_out = dresid (may be NULL) + dRMSNorm(dy); dx_bf16 (may be NULL) = bf16 copy; dw += sum_t dy * xhat. */
int ggpt_rmsnorm_bwd(const void* dy, long long lddy, const float* x, const float* rstd, const float* w,
                     const float* dresid, float* dx_out, void* dx_bf16, float* dw, long long T, int d, void* stream);

/* dgu = [dact * gf[:,0:I] | dact * gf[:,I:2I]] for the saved factors gf = [u*gelu'(g) | gelu(g)] — the stand-alone form of
 * the ggpt_gemm_bf16_dgeglu epilogue (used when dropout sits between the dgrad GEMM and the multiply).
 * ref: autograd of HF:182-184 with erf GELU */
int ggpt_geglu_bwd(const void* dact, const void* gf, void* dgu, long long T, int I, void* stream);

/* Loss-head compaction without host-side boolean indexing.  labels int64 [T,F] (-100 = ignore).
 * counts[0] = M rows with >= 1 label, counts[1] = L labelled entries; sel_rows[M] token index per selected row;
 * ent_src[L] = m*F+f, ent_label[L], ent_tok[L] in the reference's row-major order.
 * scratch needs ggpt_head_scratch_ints(T) ints.   ref: modeling_helpers.py:263-301. */
long long ggpt_head_scratch_ints(long long T);
int ggpt_head_compact(const long long* labels, long long T, int F, int* scratch, int* counts, int* sel_rows,
                      int* ent_src, int* ent_label, int* ent_tok, void* stream);
/* out[i,:] = src[idx[i],:] (idx NULL -> identity) / out[idx[i],:] = src[i,:]; bf16 rows of d elements;
 * the row count is min(*n_ptr, n_max) when n_ptr != NULL (device-side count, no host sync). */
int ggpt_gather_rows(const void* src, long long lds, const int* idx, void* out, long long ldo, const int* n_ptr,
                     int n_max, int d, void* stream);
int ggpt_scatter_rows(const void* src, long long lds, const int* idx, void* out, long long ldo, const int* n_ptr,
                      int n_max, int d, void* stream);
/* out[j,:] (bf16 [n_out, ldo], fully written) = src[i,:] where idx[i] == j, zero rows elsewhere; idx[n] must be strictly
 * increasing (sel_rows / ent_src of ggpt_head_compact are).  One pass instead of memset + ggpt_scatter_rows: the
 * backward of the boolean-mask indexing of modeling_helpers.py:288-300. */
int ggpt_expand_rows(const void* src, long long lds, const int* idx, int n, void* out, long long ldo, long long n_out, int d,
                     void* stream);

/* fp32 cross-entropy over logits [L, ldl] (V valid columns): row_lse, optional row_loss, and
 * loss_sum += wgt[e]*(lse - logit[label]) (wgt NULL -> 1), wgt_sum += wgt[e] (may be NULL).
 * focal_gamma > 0: each entry's loss is additionally scaled by the detached (1 - p_t)^gamma of FocalLoss.
 * ref: modeling_helpers.py:145-198 (CrossEntropyLoss on logits.float()); utils_graphgpt.py:340-377 (FocalLoss). */
int ggpt_ce_fwd(const float* logits, long long ldl, const int* labels, const float* wgt, float* row_lse, float* row_loss,
                double* loss_sum, double* wgt_sum, int L, int V, float focal_gamma, int* err_flag, void* stream);
/* loss = loss_sum / denom, scale = 1/denom; mode 0: denom = *count (mean), 1: *wgt_sum + 1e-7, 2: fixed_denom. */
int ggpt_ce_finalize(const double* loss_sum, const double* wgt_sum, const int* count, int mode, float fixed_denom,
                     float* loss, float* scale, void* stream);
/* dlogits (bf16 [L, ldd], pad columns zeroed) = (softmax - onehot) * wgt[e] * scale[0] * gout[0] (gout NULL -> 1),
 * times (1 - p_t)^focal_gamma when focal_gamma > 0 (p_t is a constant in the reference's FocalLoss). */
int ggpt_ce_bwd(const float* logits, long long ldl, const int* labels, const float* wgt, const float* row_lse,
                const float* scale, const float* gout, void* dlogits, long long ldd, int L, int V, float focal_gamma,
                void* stream);

/* ---------------------------------------------------------------------------------------------
 * Un-masking generation.  ref: src/utils/generation_utils.py
 * ------------------------------------------------------------------------------------------- */
/* sample_tokens (:44-82) over logits f32 [R, ldl] (V valid columns), one row per (sample, position, feature) entry:
 * logits / temperature (temperature > 0), top-p filter (0 < top_p < 1; :22-33: a token is dropped when the softmax mass
 * of the strictly larger logits exceeds top_p), top-k filter (0 < top_k < V; :36-41), fp32 softmax, then
 *   x0[r]  = temperature > 0 && u != NULL ? inverse-CDF sample with the uniform draw u[r] : arg max (lowest index)
 *   conf[r] (may be NULL) = conf_mode 0: p[x0];  1: top1 - top2 probability (margin_confidence);  2: sum p log(p + 1e-10)
 * probs (may be NULL; f32 [R, ldp]) receives the filtered softmax (tests / tooling). */
int ggpt_gen_sample(const float* logits, long long ldl, long long R, int V, float temperature, int top_k, float top_p,
                    const float* u, int conf_mode, long long* x0, float* conf, float* probs, long long ldp,
                    void* stream);
/* alg "origin" (:156-170): x[i] = x0[i] where x[i] == mask_token and u[i] < p_transfer (u = caller's uniform draws). */
int ggpt_gen_unmask_origin(long long* x, const long long* x0, const float* u, float p_transfer, long long mask_token,
                           long long n, void* stream);
/* algs maskgit_plus / topk_margin / entropy (:171-228): in each sample b of x int64 [B,P] the k_per_sample[b] most
 * confident masked entries take x0 (ties: lowest position first).  gumbel_u (may be NULL; f32 [B,P] uniform draws)
 * perturbs the scores as conf / alg_temp - log(-log(u + 1e-9) + 1e-9) (:197-205). */
int ggpt_gen_unmask_topk(long long* x, const long long* x0, const float* conf, const float* gumbel_u, float alg_temp,
                         const int* k_per_sample, long long mask_token, int B, int P, void* stream);

/* out += sum(g^2) (double accumulator, pre-zeroed by the caller). */
int ggpt_sumsq(const float* g, long long n, double* out, void* stream);
/* Fused AdamW (torch.optim.AdamW semantics) over a flat fp32 parameter buffer, with optional global-norm clipping
 * (gnorm_sq = device pointer to sum of squared grads, max_norm > 0) and grad_scale (1/loss-scale or 1/world).
 * Also writes the bf16 compute copy p_bf16 (may be NULL).   ref: training_utils.py:71-86, opt_utils.py:18-24. */
int ggpt_adamw(float* p, void* p_bf16, const float* g, float* m, float* v, long long n, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, const double* gnorm_sq, float max_norm,
               float grad_scale, void* stream);
int ggpt_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);
/* bf16 -> fp32 (gradient segments come back from a bf16 all-reduce; ref: ds_config2_pt_bf16.json:24-27 reduces in bf16) */
int ggpt_cast_bf16_f32(const void* src, float* dst, long long n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Validation-only PRECISE mode (GGPT_PRECISE=1 in the Python host; never on the bench / training path).
 * Every GEMM still goes through ggpt_gemm_bf16 (the same tcgen05 kernel) on split-bf16 operands — ggpt_vp_split3 writes
 * an fp32 matrix [R,C] as bf16 [R,3C]: role 0 (A operand) = hi | lo | hi, role 1 (B operand) = hi | hi | lo, with
 * hi = bf16(x), lo = bf16(x - hi), so that A' B'^T = A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T (error ~2^-17) — and the
 * element-wise steps between the GEMMs run as plain fp32 SIMT kernels straight from the reference formulas.  The mode
 * exists to show that the fast path's distance from the fp32 reference is bf16 storage rounding and nothing else.
 * ref: HF:59-64 (RMSNorm), HF:138-168 (RoPE), HF:199-221 (attention), HF:182-184 (GeGLU), HF:325,331 (residual adds),
 *      modeling_pretrain.py:134-145 / modeling_helpers.py:127-139 (raw-embedding swap + norm).
 * ------------------------------------------------------------------------------------------- */
int ggpt_vp_split3(const float* x, long long ldx, void* out, long long ldo, long long R, int C, int role, void* stream);
int ggpt_vp_rmsnorm_f32(const float* x, const float* w, float* y, long long T, int d, float eps, void* stream);
int ggpt_vp_raw_embed_f32(const float* raw, const long long* labels, long long ldl, int fchk, const float* mask_tok,
                          const float* w, float* y, long long T, int E, float eps, void* stream);
int ggpt_vp_rope_f32(float* qkv, long long ld, const int* pos, const float* cos_tab, const float* sin_tab, long long T,
                     int rope_cols, void* stream);
int ggpt_vp_attn_f32(const float* qkv, long long ld, int q_col0, int k_col0, int v_col0, const uint32_t* mask_bits,
                     float* out, long long ldo, int N, int S, int H, void* stream);
int ggpt_vp_geglu_f32(const float* gu, long long ldgu, float* act, long long T, int I, void* stream);
int ggpt_vp_add_f32(const float* x_in, const float* y, long long ldy, const float* colscale, const float* rowscale,
                    float* x_out, long long T, int d, void* stream);
int ggpt_vp_gather_rows_f32(const float* src, long long lds, const int* idx, float* out, long long n, int d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fine-tuning head: last-valid-row pooling + `score` head + task loss, one CTA per sample.
 *   seq_idx[n] = (#ids of sample n != pad_id) - 1 (negative wraps)          ref: modeling_helpers.py:78-86
 *   pooled[n]  = hidden[n, seq_idx[n]] (bf16 copy = task_hidden_states)     ref: modeling_finetune.py:293-296
 *   logits[n]  = score(pooled[n]) in fp32: n_layers Linear layers (weights W[l] fp32 [dims[l+1], dims[l]], optional
 *                biases), act = 1 puts gelu_erf (+ dropout drop_p) in front of EVERY Linear — the MLP head of
 *                modules_utils.py:8-34 — act = 0 is the plain nn.Linear head     ref: modeling_finetune.py:281,289-292
 *   loss_out   = [loss, 1 / denominator]   mode 0: CrossEntropy mean; 1: sum(l w) / sum(w) with sample_wgt; 2: MSE;
 *                3: L1 (both: mean over all N x num_labels elements); 4: BCE-with-logits over the non-NaN labels;
 *                -1: no labels, logits only                                   ref: modeling_finetune.py:167-234
 * `ids` = first-feature token ids, element (n, s) at ids[n * ids_ld_n + s * ids_ld_s].  W / b / dW / db are HOST arrays
 * of n_layers DEVICE pointers.  `pre` (fp32 [n_layers, N, max_dim]) keeps every layer's input for the backward pass.
 * Backward adds the weight / bias gradients into dW / db (entries may be NULL) and writes the gradient of the pooled
 * rows into the PRE-ZEROED dhidden [N*S, ldd] (bf16); gout = upstream gradient of the loss (1 float, device).
 * ------------------------------------------------------------------------------------------- */
int ggpt_ft_head_fwd(const void* hidden, long long ld, const long long* ids, long long ids_ld_n, long long ids_ld_s,
                     long long pad_id, int N, int S, int n_layers, const int* dims, const float* const* W,
                     const float* const* b, int act, float drop_p, unsigned long long drop_seed, int mode,
                     const long long* labels_i, const float* labels_f, const float* sample_wgt, int* seq_idx, void* pooled,
                     float* logits, float* pre, int max_dim, float* row_loss, float* row_den, float* loss_out, int* err_flag,
                     void* stream);
int ggpt_ft_head_bwd(int N, int S, int n_layers, const int* dims, const float* const* W, float* const* dW,
                     float* const* db, int act, float drop_p, unsigned long long drop_seed, int mode,
                     const long long* labels_i, const float* labels_f, const float* sample_wgt, const int* seq_idx,
                     const float* logits, const float* pre, int max_dim, const float* loss_out, const float* gout,
                     void* dhidden, long long ldd, void* stream);

/* Raw EDGE embeddings [N,S,S,E] of the fine-tuning models (ref: modeling_helpers.py:127-139, 4-D input): raw is fp32
 * [T, S2, E] (T = N*S tokens, S2 = S neighbours).  Forward (dh == NULL): h[t,:] = bf16(w * sum_j raw[t,j,:] rstd[t,j]
 * drop(t,j,:)) — the layer-normed, dropped-out rows summed over j; embed_proj (bias-free) then runs ONCE on the T sums.
 * Backward (dh != NULL, bf16 [T, lddh]): dw[E] += sum_t dh[t,:] * sum_j raw[t,j,:] rstd[t,j] drop(t,j,:). */
int ggpt_raw_embed_norm_sum(const float* raw, const float* w, void* h, long long ldh, const void* dh, long long lddh,
                            float* dw, long long T, int S2, int E, float eps, float drop_p, unsigned long long drop_seed,
                            void* stream);

/* loss_type "token_ce_intra" (ref: modeling_finetune.py:137-165): logits[n,s,c] = 20 * <a[n,s], a[n, cls_idx[n] + c]> with
 * a = hidden / max(||hidden||, 1e-12) (the sample's own hidden states at its C class-token positions serve as label
 * embeddings), CrossEntropy over the labelled positions (labels int64 [N*S], -100 ignored; NULL = logits only).
 * rn = fp32 [N*S] scratch (inverse norms, reused by backward), logits fp32 [N*S, C], loss_out = [loss, 1/#labelled].
 * Backward writes dhidden [N*S, ldd] (bf16, every row); dl_scratch fp32 [N*S, C], g_scratch fp32 [N*S, d]. */
int ggpt_ft_intra_fwd(const void* hidden, long long ld, const long long* cls_idx, const long long* labels, float* rn,
                      float* logits, float* row_loss, float* row_den, float* loss_out, int N, int S, int C, int d,
                      int* err_flag, void* stream);
int ggpt_ft_intra_bwd(const void* hidden, long long ld, const long long* cls_idx, const long long* labels, const float* rn,
                      const float* logits, const float* loss_out, const float* gout, float* dl_scratch, float* g_scratch,
                      void* dhidden, long long ldd, int N, int S, int C, int d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Eulerian-path serialisation of a batch of graphs.  ggpt_euler_paths is a HOST routine (all pointers are
 * host pointers; multi-threaded C++, n_threads <= 0 = all cores): for graph g with node_count[g] nodes and the undirected
 * edges edges[2*edge_off[g] .. 2*edge_off[g+1]) (local node ids) it walks every connected component along an Euler
 * circuit of the eulerised component (odd-degree nodes joined along shortest paths), cut at the step that covers the
 * last edge, visits components in random order joined by jump edges, and re-indexes the nodes cyclically in order of
 * first appearance: node_map[node_off[g] + v] = (start_g + k_v) % scope.  steps = (src, tgt, original edge index within
 * the graph's de-duplicated edge list | -1 for a jump edge) triples, graph g at steps[3*step_off[g] ..).  Returns the
 * total number of steps; when steps == NULL or steps_cap is too small nothing is copied (sizing call).
 * ref: src/utils/nx_utils.py:388-422 (graph2path_v2, connected_graph2path), :331-348 (shorten_path), :234-260 (cyclic map).
 * The walk itself is a random object (networkx matching + Python's RNG in the reference): parity is property-level
 * (oracle/euler_oracle.py), the re-index given the walk is exact.
 * ggpt_stack_path_rows (DEVICE): one row of stacked tokens per path position — node-id token, the node's attribute tokens,
 * the traversed edge's attribute tokens (default tokens for the first row and for jump edges).
 * ref: src/data/tokenizer.py:1196-1266 (stack_node_edge_graph_attr_to_node).
 * ------------------------------------------------------------------------------------------- */
long long ggpt_euler_paths(int n_graphs, const int* node_count, const long long* edge_off, const int* edges,
                           unsigned long long seed, int scope, int n_threads, long long* step_off, int* steps,
                           long long steps_cap, long long* node_off, int* node_map);
int ggpt_stack_path_rows(const int* row_node, const int* row_edge, const int* node_map, const long long* node_attr,
                         int n_node_attr, const long long* edge_attr, int n_edge_attr, const long long* default_edge,
                         long long node_base, long long* rows, long long R, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GGPT_B200_H_ */
