/* eec_experiments.h -- entry points of kernels that are NOT on the product path.  They are built only by
 * `make -C early-exit-transformer_b200 experiments` into eec/libeec_exp.so (libeec.so does not contain them) and exist for A/B
 * timing with tools/kbench.py. */
#ifndef EEC_EXPERIMENTS_H_
#define EEC_EXPERIMENTS_H_
#include "eec.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- fused feed-forward module (TA:91-119 + half-step residual TA:185-187, 207-209 + the LayerNorm that follows:
 *      TA:151 self_attn_layer_norm after ffn1, TA:211 final_layer_norm after ffn2).  bf16 operands only (tcgen05):
 *        x_out  = residual + alpha * ( SiLU(u W1^T + b1) W2^T + b2 )          fp32 [rows, 256]
 *        ln_out = LayerNorm(x_out; ln_gamma, ln_beta)                          bf16|fp32 [rows, 256], eps 1e-5
 *      u [rows,256] bf16 = LayerNorm output feeding the module; w1 [f,256], w2 [256,f] bf16 (nn.Linear layout);
 *      hpre (optional, bf16 [rows,f]) receives u W1^T + b1 for the backward pass; the [rows,f] activation itself
 *      never leaves the SM.  ln_mean / ln_rstd optional.  d must be 256, f a multiple of 128. */
int eec_ffn_fwd(const void* u, const void* w1, const float* b1, const void* w2, const float* b2, const float* residual,
                float alpha, const float* ln_gamma, const float* ln_beta, float* x_out, void* ln_out, int ln_dtype,
                float* ln_mean, float* ln_rstd, void* hpre, int rows, int d, int f, eec_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EEC_EXPERIMENTS_H_ */
