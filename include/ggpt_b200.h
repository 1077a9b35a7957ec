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

#define GGPT_ABI_VERSION 1

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

/* out[M,N] (fp32) = resid[M,N] (fp32) + rowscale[m] * colscale[n] * (A[M,K] * B[N,K]^T)
 * colscale (LayerScale lambda_1/lambda_2) and rowscale (DropPath keep-mask / keep-prob) may be NULL.
 * ref: HF:325,331 (residual adds after o_proj / down_proj); utils_graphgpt.py:153-166 (lambda, drop_path). */
int ggpt_gemm_bf16_resid(const void* A, long long lda, const void* B, long long ldb, const float* resid,
                         long long ldr, const float* colscale, const float* rowscale, float* out, long long ldo,
                         int M, int N, int K, void* stream);

/* Fused gate|up projection with GeGLU epilogue.  Wgu is [2I, K]: rows [0,I) = gate_proj.weight, rows [I,2I) =
 * up_proj.weight.  Writes gu[M,2I] = [gate | up] (bf16, may be NULL for inference) and
 * act[M,I] = gelu_erf(gate) * up (bf16).   ref: HF:182-184 with hidden_act="gelu" (exact erf GELU). */
int ggpt_gemm_bf16_geglu(const void* A, long long lda, const void* Wgu, long long ldb, void* gu, long long ldgu,
                         void* act, long long ldact, int M, int N2, int K, void* stream);

/* Fused q|k|v projection with rotary embedding applied to the first rope_cols columns (q and k heads of 64).
 * pos[M] = position id of each row; cos_tab/sin_tab = fp32 [max_pos, 32] (HF inv_freq table).
 * ref: HF:262-267 (projections + apply_rotary_pos_emb), HF:124-135,138-168 (cos/sin, rotate_half). */
int ggpt_gemm_bf16_qkv_rope(const void* A, long long lda, const void* Wqkv, long long ldb, void* qkv, long long ldc,
                            const int* pos, const float* cos_tab, const float* sin_tab, int rope_cols, int M, int N,
                            int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GGPT_B200_H_ */
