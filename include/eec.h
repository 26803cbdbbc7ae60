/* eec.h -- C ABI of libeec.so: hand-written sm_100a kernels for the early-exit conformer
 * hot path (Early_conformer encoder fwd/bwd + per-exit CTC heads).
 *
 * The reference (augustgw/early-exit-transformer) has no FFI of its own: its "operator
 * interface" for this path is the Python class models/model/early_exit.py:565-634 plus the
 * torch / torchaudio library calls it issues.  Each entry point below replaces one of those
 * call sites (cited per function; TA = torchaudio/models/conformer.py 2.11.0).  The host-side
 * mirror of the reference class that binds these symbols with ctypes lives in
 * early-exit-transformer_b200/eec/ ; INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer unless named h_*; nothing is allocated
 *     inside except small cached TMA descriptors; no implicit synchronisation;
 *   - all work is enqueued on `stream` (pass torch's current stream);
 *   - return 0 on success, non-zero on error; eec_last_error() gives a thread-local message;
 *   - activations are frame-major: row r = b*T + t, channel contiguous;
 *   - dtype enum selects the storage type of GEMM/attention operands (EEC_F32 = FFMA parity
 *     path, EEC_BF16 = tcgen05/TMEM/TMA path).  The residual stream, statistics, log-probs,
 *     CTC and all gradients of parameters are fp32 in both modes.
 */
#ifndef EEC_H_
#define EEC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* eec_stream_t; /* == cudaStream_t */

typedef enum { EEC_F32 = 0, EEC_BF16 = 1 } eec_dtype;
typedef enum {
  EEC_ACT_NONE = 0,
  EEC_ACT_SILU = 1,  /* out = silu(acc+bias); optional preact store                        */
  EEC_ACT_GLU = 2,   /* N even; out[:, n] = z[n] * sigmoid(z[n+N/2]), out has N/2 columns   */
  EEC_ACT_DSILU = 3, /* out = (acc) * silu'(preact[m,n])  (backward of SILU)                */
  EEC_ACT_RELU = 4,  /* out = max(acc+bias, 0)   (nn.TransformerDecoderLayer's feed-forward, early_exit.py:703-711) */
  EEC_ACT_DRELU = 5, /* out = (acc) * [preact[m,n] > 0]; preact = the forward's ReLU OUTPUT (backward of RELU) */
} eec_act;

const char* eec_last_error(void);
int eec_version(void);
/* 1 when the current device is compute capability 10.x (the only supported target) */
int eec_device_ok(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long eec_launch_count(void);

/* ---- GEMM with fused epilogue -------------------------------------------------------
 * C[M,N] = epi( sum_k A(m,k) * B(n,k) ),  replaces torch F.linear / 1x1 Conv1d / cuBLAS calls
 * at TA:104-108 (FFN), torch nn/functional.py in-proj/out-proj (TA:194-200), TA:44-51,66-73
 * (pointwise convs), early_exit.py:28-43 (front-end convs after im2col), :629 (exit heads),
 * and their autograd dgrad/wgrad.
 *   a_kmajor=1: A(m,k) = A[m*lda + k]   a_kmajor=0: A(m,k) = A[k*lda + m]
 *   b_kmajor=1: B(n,k) = B[n*ldb + k]   b_kmajor=0: B(n,k) = B[k*ldb + n]
 * epi(v) for element (m,n):  v += bias[n];  v = act(v);  v = alpha*v;
 *                            v += residual[(res_row_mod ? m % res_row_mod : m)*ldr + n];
 *                            accumulate ? C[m,n] += v (fp32 atomics, split-K allowed) : C[m,n] = v
 * Optional row-wise tails (need N == 256 so that one CTA owns whole rows):
 *   ln_out != NULL : ln_out[m,:] = LayerNorm(C[m,:]; ln_gamma, ln_beta, eps 1e-5) stored as ln_dtype,
 *                    optional ln_mean/ln_rstd [M] (TA:103,151,42,211)
 *   ln2_*          : a second LayerNorm applied to ln_out's fp32 value (final_layer_norm followed by
 *                    the next module's LayerNorm); when set, C receives the FIRST LN's output.
 */
typedef struct {
  int M, N, K;
  const void* A; int lda; int a_kmajor;
  const void* B; int ldb; int b_kmajor;
  int in_dtype;             /* eec_dtype of A and B                                         */
  const float* bias;        /* [N] or NULL                                                  */
  int act;                  /* eec_act                                                      */
  void* preact; int ldp;    /* SILU: optional store of acc+bias (in_dtype... see preact_dtype) */
  int preact_dtype;         /* eec_dtype of preact (read for DSILU, written for SILU/GLU)   */
  float alpha;              /* scale applied after act (1.0f if unused)                     */
  const float* residual; int ldr; int res_row_mod;
  void* C; int ldc; int out_dtype;
  int accumulate;           /* 1: fp32 C += result                                          */
  const float* ln_gamma; const float* ln_beta; void* ln_out; int ln_dtype; int ld_ln;
  float* ln_mean; float* ln_rstd;
  const float* ln2_gamma; const float* ln2_beta; float* ln2_mean; float* ln2_rstd;
  float* x_pre;             /* with ln2: optional fp32 store of the pre-LN1 value (needed for bwd) */
  float* a_colsum;          /* weight-gradient form only (a_kmajor = 0, b_kmajor = 0, accumulate, no bias, bf16):          */
  float a_colsum_scale;     /*   a_colsum[m] += a_colsum_scale * sum_k A(m,k)  -- the bias gradient of the same layer, summed */
                            /*   from the A tiles while they sit in shared memory (no second pass over the [rows, M] tensor) */
  /* dropout (see "dropout" below) applied right after the activation: v = act(v + bias); v = dropout(v); v = alpha * v; ...
   * element index of (m, n) is m*N + n.  SILU / DSILU / NONE epilogues and the LayerNorm tail; drop_state NULL = off.     */
  const uint64_t* drop_state; float drop_p; uint32_t drop_site;
  const void* drop_bits;    /* bf16 (tensor-core) path: keep-mask words of eec_dropout_bits for this site -- R = M, C = Cs = N, W = 16
                             * (SILU / DSILU / NONE epilogues) or W = 32 (LayerNorm tail); the fp32 path evaluates Philox in place */
} eec_gemm_desc;
int eec_gemm(const eec_gemm_desc* d, eec_stream_t stream);

/* ---- dropout (train mode, drop_prob > 0): nn.Dropout at positional_encoding.py:72, TA:106 and TA:108 (feed-forward),
 *      TA:73 (convolution module), TA:201 (after the attention out-projection) and the dropout nn.MultiheadAttention applies
 *      to the attention probabilities (TA:152).  Masks are COUNTER-BASED, never stored: element i of the logical tensor at
 *      dropout site `site` is kept iff   u16_{i%8}( Philox4x32-10( key = seed, counter = {i/8 lo, i/8 hi, site, offset} ) ) >= thr,
 *      thr = round(p * 65536), and kept values are scaled by 65536 / (65536 - thr)  (p is quantised to 1/65536).
 *      `state` is a DEVICE array {seed, offset}: the backward kernels regenerate the forward masks from the same
 *      (state, site) and a CUDA-graph replay sees a new offset without being re-captured (eec_dropout_advance).
 *      PyTorch's own RNG streams cannot be reproduced (SURVEY App. A "Dropout sites"); rate, scale and determinism can. */
/* y[i] = x[i] * mask_i * scale  (dtype enums of x / y; x == y allowed when the dtypes agree) */
int eec_dropout(const void* x, int in_dtype, void* y, int out_dtype, int64_t n, const uint64_t* state, float p,
                uint32_t site, eec_stream_t stream);
/* Keep-mask words for the tensor-core kernels.  Their epilogues are instruction-issue bound, so they do not evaluate Philox
 * themselves: this kernel evaluates the same draws once per forward into 1 bit per element, and the forward AND backward kernels of
 * the site read them (one coalesced word per thread and tile).  Logical tensor [R, C] with element index r*Cs + c (Cs % 8 == 0),
 * word width W = 16 | 32:  word (c/W, r) -> bits[(c/W)*R + r] (uint16 / uint32), bit j <-> element (c/W)*W + j is kept. */
int eec_dropout_bits(const uint64_t* state, float p, uint32_t site, int64_t R, int C, int64_t Cs, int W, void* bits,
                     eec_stream_t stream);
/* state[1] += 1 (one launch; lives inside the captured training step) */
int eec_dropout_advance(uint64_t* state, eec_stream_t stream);

/* ---- LayerNorm (TA:103,151,42,211) ------------------------------------------------- */
int eec_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* out, int out_dtype,
                      float* mean, float* rstd, int rows, int d, eec_stream_t stream);
/* dx (+)= LN'(dy); dgamma += sum dy*xhat; dbeta += sum dy  (dgamma/dbeta accumulate);
 * dx_copy (optional, dtype dx_copy_dtype): copy of the final dx as the operand of the next dgrad/wgrad GEMM;
 * dx_colsum (optional): dx_colsum[c] += colsum_scale * sum_rows dx_copy[:,c]  (the bias gradient of the
 * projection that precedes this LayerNorm's residual branch in the forward pass);
 * drop_state != NULL: that projection's output went through dropout (site drop_site, element index r*256 + c) before
 * joining the residual stream, so dx_copy and dx_colsum receive dropout'(dx) = dx * mask * scale (dx itself does not). */
int eec_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                      const float* gamma, float* dx, int dx_accumulate, float* dgamma, float* dbeta,
                      void* dx_copy, int dx_copy_dtype, float* dx_colsum, float colsum_scale,
                      const uint64_t* drop_state, float drop_p, uint32_t drop_site, int rows, int d,
                      eec_stream_t stream);
/* the same with the upstream gradient dy in dy_dtype (EEC_F32 | EEC_BF16): on the bf16 path the data-gradient GEMM that produces dy
 * (torch autograd of the nn.Linear / 1x1 Conv1d behind a LayerNorm, TA:104,194,44) writes it in bf16 */
int eec_layernorm_bwd_dy(const void* dy, int dy_dtype, const float* x, const float* mean, const float* rstd,
                         const float* gamma, float* dx, int dx_accumulate, float* dgamma, float* dbeta,
                         void* dx_copy, int dx_copy_dtype, float* dx_colsum, float colsum_scale,
                         const uint64_t* drop_state, float drop_p, uint32_t drop_site, int rows, int d,
                         eec_stream_t stream);

/* ---- multi-head self-attention core (nn.MultiheadAttention SDPA branch, TA:194-200) -
 * qkv [B*T, 3*H*dh] rows = [q | k | v], head h = columns [h*dh, (h+1)*dh) of each third.
 * keys t' >= key_len[b] are masked; fully-masked rows produce 0.  lse [B,H,T] (natural log).
 * drop_state != NULL: dropout on the attention probabilities (after the softmax, before the product with V);
 * element index of probability (b, h, t, t') is ((b*H + h)*T + t) * (8*ceil(T/8)) + t'. */
int eec_attn_fwd(const void* qkv, int dtype, const int32_t* key_len, void* ctx, float* lse,
                 int B, int T, int H, int dh, const uint64_t* drop_state, float drop_p, uint32_t drop_site,
                 const void* drop_bits /* EEC_BF16: eec_dropout_bits(R = B*H*T, C = T, Cs = 8*ceil(T/8), W = 32) */,
                 eec_stream_t stream);
/* dvec: fp32 workspace [B*H*T] (row dots dO.O); dq32: fp32 workspace [B*T, H*dh] (bf16 path: dQ
 * partials of the key blocks are summed there with vector atomics; may be NULL for EEC_F32) */
int eec_attn_bwd(const void* qkv, const void* ctx, const void* dctx, int dtype, const float* lse,
                 const int32_t* key_len, void* dqkv, float* dvec, float* dq32, int B, int T, int H, int dh,
                 const uint64_t* drop_state, float drop_p, uint32_t drop_site, const void* drop_bits, eec_stream_t stream);

/* ---- general attention core: the AED decoder stacks (nn.TransformerDecoderLayer's self_attn with the causal
 *      tgt_mask + tgt_key_padding_mask and multihead_attn over the encoder states, early_exit.py:701-717, :742-800) -----------------
 * Queries and keys / values may come from different tensors with different lengths:
 *   Q(b, t, h, :)  = q[(b*Tq + t)*ldq + h*dh ...]   t < Tq        K(b, t', h, :) = k[(b*Tk + t')*ldk + h*dh ...]   t' < Tk   (V likewise)
 * (point q / k / v at the first column of the respective block of a packed projection output).  Key t' is visible to query t iff
 * t' < key_len[b] (key_len != NULL), bit (t' & 31) of key_valid_bits[b*ceil(Tk/32) + t'/32] is set (key_valid_bits != NULL) and, with
 * causal != 0, t' <= t.  Rows without a visible key produce 0.  ctx(b, t, h, :) = ctx[(b*Tq + t)*ldo + h*dh ...]; lse [B, H, Tq]. */
typedef struct {
  int B, H, dh, Tq, Tk;
  const void* q; int ldq;
  const void* k; int ldk;
  const void* v; int ldv;
  int dtype;                        /* eec_dtype of q / k / v / ctx and of the gradients */
  const int32_t* key_len;           /* [B] or NULL */
  const uint32_t* key_valid_bits;   /* [B, ceil(Tk/32)] or NULL */
  int causal;
  /* dropout on the attention probabilities (nn.MultiheadAttention(dropout = p) inside the decoder layers), same conventions as
   * eec_attn_fwd: element index of probability (b, h, t, t') is ((b*H + h)*Tq + t) * (8*ceil(Tk/8)) + t'; EEC_BF16 needs the keep-mask
   * words of eec_dropout_bits(R = B*H*Tq, C = Tk, Cs = 8*ceil(Tk/8), W = 32) in drop_bits; drop_state NULL = off */
  const uint64_t* drop_state; float drop_p; uint32_t drop_site; const void* drop_bits;
} eec_attn_desc;
int eec_attn_general_fwd(const eec_attn_desc* d, void* ctx, int ldo, float* lse, eec_stream_t stream);
/* dq [B*Tq rows, lddq], dk / dv [B*Tk rows, lddk / lddv] are WRITTEN (head h at column h*dh of the given pointers); dvec: fp32
 * workspace [B*H*Tq]; dq32: fp32 workspace [B*Tq, H*dh] (EEC_BF16 tensor-core path; may be NULL for EEC_F32). */
int eec_attn_general_bwd(const eec_attn_desc* d, const void* ctx, const void* dctx, int ldo, const float* lse, void* dq, int lddq,
                         void* dk, int lddk, void* dv, int lddv, float* dvec, float* dq32, eec_stream_t stream);
/* key_valid_bits for a padded token matrix: bit set iff tokens[b, t] != pad  (tgt_key_padding_mask = (trg == pad), early_exit.py:802-805) */
int eec_key_bits_from_tokens(const int64_t* tokens, int B, int L, int64_t pad, uint32_t* bits, eec_stream_t stream);

/* ---- AED decoder glue ---------------------------------------------------------------------------------------
 * x[b*L + t, :] = emb[tokens[b, t], :] + pe[t, :]      (nn.Embedding + PositionalEncoding, early_exit.py:776-777; no sqrt(d) scaling) */
int eec_embed_pe(const int64_t* tokens, const float* emb, const float* pe, float* x, int B, int L, int D, int V, eec_stream_t stream);
/* demb[tokens[b, t], :] += dx[b*L + t, :]   (fp32 atomics; demb accumulates) */
int eec_embed_bwd(const int64_t* tokens, const float* dx, float* demb, int B, int L, int D, int V, eec_stream_t stream);
/* nn.CrossEntropyLoss() with its defaults (mean over ALL rows, no ignore_index: the pad id is scored, train.py:47, :258):
 * loss_out[0] += mean_r( logsumexp(logits[r, :]) - logits[r, target[r]] );  dlogits (may be NULL) = (softmax - onehot) / rows. */
int eec_cross_entropy(const float* logits, const int64_t* targets, int rows, int V, float* loss_out, float* dlogits, eec_stream_t stream);

/* ---- conformer convolution module interior (TA:52-65) ------------------------------
 * g [B,T,C] (dtype) -> depthwise conv k (SAME, zero pad per utterance) + bias.
 * eval : out = silu(bn_eval(c))                     (one kernel)
 * train: pass A writes c and per-channel sum / sumsq (double[2*C], caller zeroes);
 *        pass B normalises with batch stats, SiLU, and updates running stats. */
int eec_dwconv_bn_silu_eval(const void* g, int dtype, const float* w, const float* bias,
                            const float* bn_w, const float* bn_b, const float* run_mean,
                            const float* run_var, void* out, int B, int T, int C, int K,
                            eec_stream_t stream);
int eec_dwconv_stats(const void* g, int dtype, const float* w, const float* bias, float* c,
                     double* sums, int B, int T, int C, int K, eec_stream_t stream);
int eec_bn_silu_train(const float* c, const double* sums, const float* bn_w, const float* bn_b,
                      float* run_mean, float* run_var, int64_t* num_batches_tracked, float momentum,
                      float* save_mean, float* save_rstd, void* out, int dtype, int rows, int C,
                      int64_t stat_rows /* 0 = rows; data parallel with synchronised statistics: `sums` was all-reduced (SUM)
                                           over the ranks and stat_rows is the row count of ALL ranks */,
                      eec_stream_t stream);
/* backward of train-mode BN+SiLU+dwconv.  Step 1: dn = ds*silu'(n); sums2 = {sum dn, sum dn*nhat}.
 * Step 2: dc = gamma*rstd*(dn - mean(dn) - nhat*mean(dn*nhat)); dgamma/dbeta accumulate.
 * Step 3: dg = dwconv^T(dc); dW/dbias accumulate. */
int eec_bn_silu_bwd_stats(const void* ds, int dtype, const float* c, const float* save_mean,
                          const float* save_rstd, const float* bn_w, const float* bn_b, double* sums2,
                          int rows, int C, eec_stream_t stream);
int eec_bn_silu_bwd_apply(const void* ds, int dtype, const float* c, const float* save_mean,
                          const float* save_rstd, const float* bn_w, const float* bn_b,
                          const double* sums2, float* dc, float* dgamma, float* dbeta, int rows, int C,
                          int64_t stat_rows /* 0 = rows */,
                          const double* sums2_local /* NULL, or (synchronised statistics) this rank's own sums for dgamma / dbeta
                                                       while `sums2` holds the all-reduced ones */,
                          eec_stream_t stream);
/* workspace: eec_dwconv_bwd_workspace_bytes(B, T, C) bytes of device memory (per-block partial weight gradients: the kernel
 * writes them without atomics and a second launch reduces them into dw / dbias, both accumulated) */
int64_t eec_dwconv_bwd_workspace_bytes(int B, int T, int C);
int eec_dwconv_bwd(const float* dc, const void* g, int dtype, const float* w, void* dg, float* dw,
                   float* dbias, int B, int T, int C, int K, void* workspace, eec_stream_t stream);
/* GLU backward: z [rows, 2C] (dtype), dg [rows,C] (dtype) -> dz [rows,2C] (dtype) */
int eec_glu_bwd(const void* z, const void* dg, void* dz, int dtype, int rows, int C, eec_stream_t stream);

/* ---- exit head tail (early_exit.py:630): log_softmax over V, + by-products ---------
 * argmax (int32 [rows]) and entropy (fp32 [rows], -sum p log p) may be NULL. */
int eec_logsoftmax_fwd(const float* logits, float* out, int32_t* argmax, float* entropy, int rows, int V,
                       eec_stream_t stream);

/* generic log_softmax backward: dlogits = g - exp(lp) * rowsum(g) */
int eec_logsoftmax_bwd(const float* g, const float* lp, float* dlogits, int rows, int V, eec_stream_t stream);

/* fused exit head: Linear(d->V) + log_softmax (+argmax, +frame entropy), early_exit.py:629-630.
 * bf16: one tcgen05 GEMM with the log-softmax in its epilogue; fp32: FFMA GEMM into logits_ws
 * [rows,V] followed by the row kernel. */
int eec_head_logsoftmax(const void* x, int dtype, const void* w, const float* bias, float* out,
                        int32_t* argmax, float* entropy, float* logits_ws, int rows, int d, int V,
                        eec_stream_t stream);

/* ---- multi-exit CTC (train.py:57-65, nn.CTCLoss(blank=0, 'mean', zero_infinity=True)) --
 * All E exits in one launch.  lp [E,B,T,V] fp32 log-probs (batch-major), targets [B,Lmax] int64,
 * target_len [B] int64, input length = T for every utterance.  nll [E,B] (0 where infeasible).
 * grad [E,B,T,V] (may be NULL) = gscale/(B*max(U_b,1)) * (exp(lp) - occupancy), i.e.
 * d(loss)/d(logits).  loss_out[e] += sum_b nll_b/(B*max(U_b,1))  (caller zeroes).
 * workspace: alpha scratch of eec_ctc_workspace_bytes(E,B,T,Lmax). */
int64_t eec_ctc_workspace_bytes(int E, int B, int T, int Lmax);
int eec_ctc_fwd_bwd(const float* lp, const int64_t* targets, const int64_t* target_len, int E, int B, int T,
                    int V, int Lmax, int blank, float gscale, float* nll, float* loss_out, float* grad,
                    void* workspace, eec_stream_t stream);

/* ---- greedy CTC decode (util/beam_infer.py:21-23): collapse repeats, drop blank ---- */
int eec_greedy_collapse(const int32_t* argmax, int32_t* tokens, int32_t* n_tokens, int B, int T, int blank,
                        eec_stream_t stream);

/* ---- CTC prefix beam search: torchaudio's cuda_ctc_decoder(tokens, nbest=1, beam_size, blank_skip_threshold=0.95)
 *      as called at util/beam_infer.py:100-110 once per exit (inference.py:66-79) -- here for all exits and utterances of a forward in
 *      ONE launch (one CTA per emission matrix walks its frames on chip).
 * lp [n_utt, T, V] fp32 log-probabilities (the (E, B, T', V) output of Early_conformer.forward is n_utt = E*B), enc_len [n_utt] int32 or
 * NULL (= T for every utterance, what the reference passes).  Frames with lp[t, blank] > log_blank_skip are not expanded but
 * counted as a pure blank emission, p_b' = (p_b + p_nb) p(blank), p_nb' = 0 -- the library's "blank skipping" (pass log(0.95); +inf disables it).  Prefix beam search with merged prefixes, `beam` <= 16 hypotheses, V <= 1024.
 * Outputs, best hypothesis first: tokens [n_utt, nbest, T] int32 padded with -1, n_tokens [n_utt, nbest], scores [n_utt, nbest]
 * (log(p_blank + p_non_blank); -inf for missing hypotheses).  workspace: eec_ctc_beam_workspace_bytes(n_utt, T, beam) bytes
 * (back-pointer trie of the prefixes). */
int64_t eec_ctc_beam_workspace_bytes(int n_utt, int T, int beam);
int eec_ctc_beam_search(const float* lp, const int32_t* enc_len, int n_utt, int T, int V, int beam, int nbest, int blank,
                        float log_blank_skip, int32_t* tokens, int32_t* n_tokens, float* scores, void* workspace,
                        eec_stream_t stream);

/* ---- front end (early_exit.py:24-48, positional_encoding.py:70-72) ----------------
 * im2col for Conv1d(k=3,s=2): out[(b,t), c*3+j] = in[b*sb + c*sc + (2t+j)*st]; out row stride ldo */
int eec_im2col_k3s2(const void* in, int in_dtype, int64_t sb, int64_t sc, int64_t st, void* out,
                    int out_dtype, int ldo, int B, int C, int T_out, eec_stream_t stream);
/* transpose of the above for the SECOND conv's input gradient (frame-major x1 [B,T_in,C]):
 * dx[b,f,c] = sum_{j} dcols[(b,(f-j)/2), c*3+j] over valid (f-j) even */
int eec_col2im_k3s2(const float* dcols, int ldc, float* dx, int B, int C, int T_in, int T_out,
                    eec_stream_t stream);
/* key_len[b] = (int) min(lengths[b]/4.0, T)   (early_exit.py:623) */
int eec_encoder_lengths(const int64_t* lengths, int32_t* key_len, int B, int T, int div, int add,
                        eec_stream_t stream);

/* ---- small utilities ----------------------------------------------------------------- */
int eec_cast(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, eec_stream_t stream);
/* out[n] += scale * sum_m in[m, n]   (bias gradients; out accumulates) */
int eec_colsum(const void* in, int dtype, int ld, float* out, float scale, int rows, int cols,
               eec_stream_t stream);
/* out_bf16[r, c] = bf16(in[r, c]) and colsum[c] += scale * sum_r in[r, c] in one pass (cols % 256 == 0): the bf16 operand copy of
 * d(logits) and the exit head's bias gradient (autograd of early_exit.py:629) */
int eec_cast_colsum(const float* in, int ld, void* out_bf16, int ldo, float* colsum, float scale, int rows, int cols,
                    eec_stream_t stream);
/* y = (*s_dev) * x  (device scalar; used to apply an upstream loss gradient without a host sync) */
int eec_scale_dev(const float* x, const float* s_dev, float* y, int64_t n, eec_stream_t stream);
/* y[r, 0:slab] = s_dev[r] * x[r, 0:slab], r < rows (x, y contiguous; slab % 4 == 0): every exit's CTC gradient scaled by its
 * upstream scalar (train.py:63 sums the per-exit losses) in one launch */
int eec_scale_rows_dev(const float* x, const float* s_dev, float* y, int rows, int64_t slab, eec_stream_t stream);

/* ---- optimiser step over flat buffers: torch.nn.utils.clip_grad_norm_ (train.py:69) + NoamOpt.step
 *      (util/noam_opt.py:26-40) + torch.optim.AdamW.step (train.py:261-262) in three launches, no host sync.
 *      params / grads / exp_avg / exp_avg_sq: fp32 [n], 16-byte aligned.  bf16_shadow (optional, bf16 [n]) receives the
 *      updated parameters as GEMM operands.  state: double[4] on the device = {step t, sum g^2, last lr, last clip coefficient};
 *      the call increments t.  clip <= 0 disables clipping; lr_fixed >= 0 replaces the Noam rate
 *      lr = model_size^-0.5 * min(t^-0.5, t * warmup^-1.5). */
int eec_noam_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* bf16_shadow, int64_t n,
                        double* state, float model_size, float warmup, float beta1, float beta2, float eps, float weight_decay,
                        float clip, float lr_fixed, eec_stream_t stream);
/* y = a*x + y over n floats */
int eec_axpy(const float* x, float a, float* y, int64_t n, eec_stream_t stream);

/* ---- early exit (north star; NOT in the reference) -------------------------------------
 * mean frame entropy over t < key_len[b]; finalise rows with H < threshold (or last exit):
 * writes exit_index / tokens for finalised ORIGINAL rows, compacts survivors (stable) and
 * updates *n_alive on device.  No host sync. */
int eec_exit_select(const float* entropy, const int32_t* argmax, const int32_t* key_len_alive,
                    const int32_t* row_map, int32_t* n_alive, int exit_idx, int is_last, float threshold,
                    int32_t* exit_index, int32_t* tokens, int32_t* n_tokens, int32_t* new_row_map,
                    int32_t* new_key_len, int32_t* gather_idx, float* mean_entropy_out, int B, int T,
                    int blank, eec_stream_t stream);
int eec_gather_rows(const float* x, float* y, const int32_t* gather_idx, const int32_t* n_alive,
                    int B, int64_t row_elems, eec_stream_t stream);
/* Active-item limit of ONE STREAM (the handle is the stream the caller already passes to every entry point, so two models or two
 * streams in one process never see each other's limit): while n_items_dev != NULL, the forward (inference) forms of eec_gemm (K-major A),
 * eec_layernorm_fwd, eec_attn_fwd and eec_dwconv_bn_silu_eval launched on that stream process only the first *n_items_dev utterances, i.e. the leading
 * *n_items_dev * rows_per_item rows of every frame-major tensor; tiles / rows / utterance blocks past the limit return at once,
 * so compaction after an exit actually removes the finished utterances' work from the later layers.  The count is read ON THE
 * DEVICE at kernel start (eec_exit_select keeps it current): no host sync, and a captured CUDA graph stays valid when it changes.
 * pad_items further utterances behind the count are processed as well: the tensor-core attention loads 128-row K/V tiles that
 * overhang into the following utterance (masked in the softmax, but 0 x NaN = NaN in P.V), so the rows right behind the last
 * survivor must be finite; the caller keeps one utterance of stale-but-finite data there (pad_items = 1 when T' >= 128).
 * The limit is looked up on the host when a kernel is LAUNCHED on `stream`: it applies to the launches (or graph-captured launches)
 * issued on that stream between the set and the clear.
 * Rows past the limit are left untouched (their contents are unspecified).  Pass NULL to clear. */
int eec_set_active_items(const int32_t* n_items_dev, int rows_per_item, int pad_items, eec_stream_t stream);
/* dst[i] = src[idx[i]], i < n (idx clamped to [0, n)): the raw lengths of the surviving utterances in compacted order (Splitformer's
 * parallel branch masks with the RAW fbank lengths, early_exit.py:332-338) */
int eec_gather_i64(const int64_t* src, const int32_t* idx, int64_t* dst, int n, eec_stream_t stream);

/* ---- feature front end: util/data_loader.py:7-18 -- torchaudio Spectrogram(n_fft = 1024, win_length = 320,
 *      hop_length = 160; hann, center, reflect, power 2) followed by MelScale(16 kHz, 80 mels, n_stft = 513), computed by the
 *      reference per utterance on the CPU in the collate function (data_loader.py:124-125).  Batched here:
 *        eec_fbank_frames -> eec_gemm ([re|im] = frames x [cos|sin]^T) -> eec_fbank_power -> eec_gemm (mel = power x fb) -> eec_fbank_finish
 *      Both GEMMs run on the bf16 tensor cores at fp32 accuracy: operands are split v = hi + lo (bf16 each) and the partial
 *      products hi*hi + hi*lo + lo*hi are contracted in one GEMM over K' = 3K: activations are laid out [hi | hi | lo] (written
 *      by eec_fbank_frames / eec_fbank_power), constant operands [hi | lo | hi] (eec_fbank_split_operand).
 * frames_bf16x3 [B*T, 3*win] bf16: row (b, t) = window[i] * wave[b][reflect(t*hop - win/2 + i)], i < win, zero for t >= 1 + wave_len[b]/hop
 * (only the win_length samples under the centred, zero-padded window are non-zero: the n_fft-point DFT is a K = win contraction). */
int eec_fbank_frames(const float* wave, const int64_t* wave_len, int64_t ldw, const float* window, void* frames_bf16x3,
                     int B, int T, int win, int hop, eec_stream_t stream);
/* spec [rows, lds] fp32 = [re (n_freqs) | im (n_freqs) | pad] -> power_bf16x3 [rows, 3*kp] bf16 ([hi | hi | lo] of re^2 + im^2, zero padded) */
int eec_fbank_power(const float* spec, int lds, void* power_bf16x3, int64_t rows, int n_freqs, int kp, eec_stream_t stream);
/* mel [B*T, ldm] fp32 (frame-major) -> out [B, n_mels, T] fp32, the layout Early_conformer.forward takes (early_exit.py:617-620) */
int eec_fbank_finish(const float* mel, int ldm, float* out, int B, int T, int n_mels, eec_stream_t stream);
/* constant operand w [rows, k] fp32 -> out_bf16x3 [rows, 3*kp] bf16 = [hi | lo | hi], zero padded to kp columns per part */
int eec_fbank_split_operand(const float* w, int k, void* out_bf16x3, int64_t rows, int kp, eec_stream_t stream);

/* ---- Splitformer branch glue (early_exit.py:318-356) --------------------------------- */
int eec_stride2_gather(const float* x, float* y, int B, int T, int D, eec_stream_t stream);
/* y[b,t,:] += up[b,t/2,:] */
int eec_repeat2_add(const float* up, float* y, int B, int T, int D, eec_stream_t stream);
/* dy_half[b,t2,:] = dy[b,2*t2,:] + dy[b,2*t2+1,:] (t < T) */
int eec_repeat2_bwd(const float* dy, float* dhalf, int B, int T, int D, eec_stream_t stream);
/* dx[b,2*t2,:] += dhalf[b,t2,:] */
int eec_stride2_scatter_add(const float* dhalf, float* dx, int B, int T, int D, eec_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EEC_H_ */
