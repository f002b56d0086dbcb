/*
 * neuspeech_b200.h -- C-ABI of the B200-native NeuSpeech hot path (libneuspeech_b200.so).
 *
 * The reference (NeuSpeech/NeuSpeech1) has no FFI of its own: its hot path is PyTorch library calls made from
 * utils/load_model.py / utils/model_utils.py / utils/augment_eeg.py.  Each entry point below
 * therefore cites the reference call site whose arithmetic it replaces.  The Python host (neuspeech1_b200/) binds these
 * with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ns_status otherwise; ns_last_error_string() explains (thread-local)
 *   - all data pointers are DEVICE pointers (HBM), row-major; `stream` is a cudaStream_t passed as void*
 *   - functions never allocate device memory, never synchronise, never touch the host copy of the data
 *   - ns_dtype says how activations are stored (NS_F32 / NS_BF16); arithmetic accumulates in fp32 in both cases
 *   - "fast" variants (tcgen05/TMEM/TMA, sm_100a) are selected automatically when dtype == NS_BF16 and the shape
 *     qualifies; ns_set_path() can force the plain SIMT kernels (parity debugging), ns_get_counters() reports which ran
 */
#ifndef NEUSPEECH_B200_H
#define NEUSPEECH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { NS_OK = 0, NS_ERR_ARG = -1, NS_ERR_CUDA = -2, NS_ERR_UNSUPPORTED = -3, NS_ERR_WORKSPACE = -4 } ns_status;
typedef enum { NS_F32 = 0, NS_BF16 = 1 } ns_dtype;
typedef enum { NS_ACT_NONE = 0, NS_ACT_GELU = 1, NS_ACT_DGELU = 2 } ns_act;
typedef enum { NS_PATH_AUTO = 0, NS_PATH_SIMT = 1, NS_PATH_FAST = 2 } ns_path;

/* ---- library ---- */
int         ns_version(void);
const char* ns_last_error_string(void);
int         ns_set_path(int path);                 /* ns_path; returns previous */
/* Programmatic dependent launch for the calling thread's subsequent launches (returns the previous setting).  The kernels of
 * the one-token decoder step (embedding, LayerNorm forward, tcgen05 GEMM, single-query attention, greedy pick) are then
 * launched with cudaLaunchAttributeProgrammaticStreamSerialization: their prologues overlap the tail of the kernel before them
 * and each waits (griddepcontrol.wait) before its first global-memory access.  Streams captured into CUDA graphs keep the
 * programmatic edges.  Off by default. */
int         ns_set_pdl(int on);
int         ns_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* counters[0]=tcgen05 GEMM launches, [1]=SIMT GEMM launches, [2]=tensor-core attention launches, [3]=SIMT attention
 * launches, [4]=other kernel launches, [5]=tcgen05 wgrad launches.  Reset with ns_reset_counters(). */
int         ns_get_counters(long long* counters, int n);
int         ns_reset_counters(void);

/* ---- fused GEMM:  D = epilogue( A[M,K] * W[N,K]^T  (+ A2[M,K2] * W2[N,K2]^T) )
 * Replaces nn.Linear (+ PEFT lora.Linear, finetune.py:194-212) at HF modeling_whisper.py:310,331-332,355,404,406 and
 * the tied proj_out (utils/load_model.py:1047).  The second product is the LoRA branch: A2 = t = x*A^T (M,K2), W2 = s*B.
 * Epilogue order: +bias[n]; *alpha for columns < alpha_cols (q pre-scale, HF:310); act; +residual.
 *   act NS_ACT_GELU : y = gelu_erf(z); if aux_out != NULL the pre-activation z is stored there (ld = ldaux)
 *   act NS_ACT_DGELU: y = z * gelu'(aux_in[m,n])                (backward through GELU; aux_in = saved pre-activation)
 *   residual: D += R[row % res_mod, n] when res_mod > 0 (position table, utils/load_model.py:413-416) else R[row, n]
 * Output tiles are stored with 16-byte granularity: when N is not a multiple of 8 (bf16) and ldd leaves room, the pad
 * columns [N, round_up(N, 8)) of D (and of aux_out) may be written with zeros; nothing beyond that is touched. */
typedef struct {
  const float* bias;
  float        alpha;
  int          alpha_cols;
  int          act;
  const void*  aux_in;
  void*        aux_out;
  long long    ldaux;
  const void*  residual;
  long long    ldr;
  int          res_mod;
  int          out_dtype;     /* ns_dtype of D (NS_F32 allowed with NS_BF16 inputs) */
  int          a2_group_cols; /* > 0: stacked adapters -- output columns [g*a2_group_cols, (g+1)*a2_group_cols) use
                                 A2[:, g*K2 : (g+1)*K2] (q/k/v share one t = x*[Aq;Ak;Av]^T buffer, lda2 >= G*K2) */
  const unsigned int* drop_bits; /* != NULL: LoRA-branch dropout in the INPUT-gradient product (PEFT: the branch input is
                                 dropout(x), so d/dx of the branch is keep . (dt A)):  D = epilogue-activation( A*W^T +
                                 keep . (A2*W2^T) ), keep(m, n) = !bit (n % 32) of drop_bits[m * drop_ld + n / 32] (one adapter's
                                 plane of ns_dropout_bits).  The second product gets its own TMEM accumulator and is masked in
                                 the epilogue before the sum.  bf16 tcgen05 path only (N % 64 == 0); else NS_ERR_UNSUPPORTED. */
  long long    drop_ld;       /* words per row of drop_bits */
  int          drop_mode;     /* 0: mask the second product (above).  1: mask the A OPERAND of the (single) product -- the LoRA
                                 down product t = alpha * (x . keep_g) * A_g^T under branch dropout (finetune.py:210): adapter g
                                 owns output columns [32 g, 32 g + 32) (rank 32) and the plane at drop_bits + g * drop_gstride;
                                 the elements are zeroed in shared memory between the TMA and the MMA (bf16 tcgen05 path only,
                                 N % 32 == 0, K % 64 == 0) */
  long long    drop_gstride;  /* words between the planes of stacked adapters (drop_mode 1) */
  int          a_group_cols;  /* > 0: BLOCK-DIAGONAL main product -- output columns [g*a_group_cols, (g+1)*a_group_cols) contract
                                 A[:, g*K : (g+1)*K] (lda >= G*K) with their own rows of W: dt_g = dy_g * B_g for the q/k/v adapters
                                 in one launch (dy = [dq|dk|dv], W = [B_q^T; B_k^T; B_v^T]).  Any multiple of 32 that divides N
                                 (32-wide tiles).  No second product with it. */
  int          aux_deriv;     /* != 0: the aux tensor holds gelu'(z), not z.  NS_ACT_GELU stores gelu'(z) into aux_out (same tanh as
                                 the activation: a handful of FMAs more), NS_ACT_DGELU multiplies by aux_in as it is -- the
                                 backward epilogue loses its transcendental and a dozen instructions per element.  The forward
                                 and the backward call of one layer must agree on it. */
  const unsigned int* drop_seed;  /* drop_mode 2: like 1, but the mask stage DRAWS the planes itself (the counter hash of
                                 ns_dropout_bits: same words for the same seed / salt / p) and stores them to drop_bits for the
                                 backward consumers -- no separate ns_dropout_bits launch, the integer hash runs under the
                                 memory-bound stream of x.  drop_seed: device word holding the step seed. */
  const unsigned int* drop_salts; /* drop_mode 2: HOST array, one salt per stacked adapter (N / 32 <= 4 of them) */
  float        drop_p;            /* drop_mode 2: dropout probability */
} ns_epilogue;

int ns_gemm_nt(int dtype, long long M, int N, int K, const void* A, long long lda, const void* W, long long ldw,
               void* D, long long ldd, const ns_epilogue* ep,
               const void* A2, long long lda2, const void* W2, long long ldw2, int K2, void* stream);

/* ---- LayerNorm + GEMM for few rows (the one-token decoder step, M <= 128): D = epilogue( LN(X)[M,K] * W[N,K]^T ), LN over the K
 * columns with gamma / beta (both NULL: no LayerNorm, then K is unrestricted), eps as ns_layernorm_fwd.  One launch in place of
 * ns_layernorm_fwd + ns_gemm_nt (HF modeling_whisper.py:393-414: every decoder sub-block starts with a LayerNorm); bias,
 * alpha / alpha_cols, NS_ACT_GELU and residual of ns_epilogue.  bf16 storage, N <= 8192, K <= 512 with LayerNorm; returns
 * NS_ERR_UNSUPPORTED otherwise.  (Measured at 128 rows: no faster than the two calls, see csrc/ns_skinny.cu; ns_decode_step
 * uses it only on request.) */
int ns_ln_gemm_nt(int dtype, long long M, int N, int K, const void* X, long long ldx, const float* gamma, const float* beta,
                  float eps, const void* W, long long ldw, void* D, long long ldd, const ns_epilogue* ep, void* stream);

/* ---- weight-gradient GEMM:  G[i*si + j*sj] += alpha * sum_m X[m,i] * Y[m,j]   (G fp32, caller zeroes it)
 * Replaces autograd's wgrad of the LoRA A/B linears (PEFT) : dB = s*dy^T t, dA = dt^T x. */
int ns_gemm_tn(int dtype, long long M, int I, int J, const void* X, long long ldx, const void* Y, long long ldy,
               float* G, long long si, long long sj, float alpha, void* stream);

/* Block-diagonal form: `groups` independent gradients in one launch, X (M, groups*I), Y (M, groups*J),
 *   G[(g*I + i)*si + j*sj] += alphas[g] * sum_m X[m, g*I + i] * Y[m, g*J + j]      (dB_q, dB_k, dB_v from dy = [dq|dk|dv], t = [t_q|t_k|t_v]).
 * alphas: host array of `groups` (<= 4) scales. */
int ns_gemm_tn_grouped(int dtype, long long M, int I, int J, int groups, const void* X, long long ldx, const void* Y, long long ldy,
                       float* G, long long si, long long sj, const float* alphas, void* stream);
/* B side of a LoRA branch's backward in ONE pass over dy (PEFT lora.Linear under autograd, finetune.py:194-212: y += t B^T):
 *   dt[m, g*r + j]        = alpha_dt[g] * sum_n dy[m, g*N + n] * Bt[g*r + j, n]      (dt = alpha' dy B; replaces ns_gemm_nt on a rank-r tile)
 *   dB[(g*N + n)*r + j]  += alpha_db[g] * sum_m dy[m, g*N + n] * t[m, g*r + j]       (fp32, caller zeroes; replaces ns_gemm_tn[_grouped])
 * for `groups` (<= 4) adapters stacked along the columns of dy (M, groups*N), t and dt (M, groups*r) and the rows of Bt
 * (groups*r, N) = B^T.  bf16 tcgen05 path only: r == 32, N % 128 == 0, 16-byte aligned operands; NS_ERR_UNSUPPORTED otherwise
 * (the caller then issues the two separate products).  alpha_dt / alpha_db: host arrays of `groups` scales.
 * N > 1792 (fc1's 2048 output columns) runs as column parts on different CTAs that combine their partial dt through a
 * caller-owned workspace: ns_lora_bwd_b_workspace_bytes() bytes (0: none needed, -1: shape not supported), 16-byte aligned,
 * its first (groups * ceil(M/128) * 16, rounded up to 256) bytes zeroed ONCE by the caller -- the kernel leaves them zero. */
int ns_lora_bwd_b(int dtype, long long M, int N, int r, int groups, const void* dy, long long lddy, const void* Bt, long long ldbt,
                  const void* t, long long ldt, void* dt, long long lddt, float* dB, const float* alpha_dt, const float* alpha_db,
                  void* workspace, long long workspace_bytes, void* stream);
long long ns_lora_bwd_b_workspace_bytes(long long M, int N, int r, int groups);
/* Same with X masked by a dropout plane (ns_dropout_bits, one adapter) on its way to the tensor cores:
 *   G += alpha * (X . keep)^T Y  --  dA = dt'^T (x . keep) of a LoRA branch under dropout, x read once.  bf16 tcgen05 path only
 *   (I % 64 == 0, xbits_ld even); NS_ERR_UNSUPPORTED otherwise. */
int ns_gemm_tn_masked(int dtype, long long M, int I, int J, const void* X, long long ldx, const void* Y, long long ldy,
                      float* G, long long si, long long sj, float alpha, const unsigned int* xbits, long long xbits_ld,
                      void* stream);

/* ---- stem convolution (kernel 3, pad 1, stride 1|2) on channels-last activations, as an implicit GEMM.
 * Replaces nn.Conv1d + GELU at utils/model_utils.py:12-16 and utils/load_model.py:410-411 (+ the permute and
 * embed_positions add of :413-416 through ep->residual/res_mod).
 *   x  (B, Tin, Cp)  channels-last, Cp = padded channel count (multiple of 16; pad channels are zero)
 *   w  (3, N, Cp)    tap-major re-layout of Conv1d.weight (N, C, 3)
 *   y  (B, Tout, N)  Tout = Tin/stride.   Epilogue as ns_gemm_nt (row index for res_mod is the output time t). */
int ns_conv3_fwd(int dtype, int B, int Tin, int Cp, int N, int stride, const void* x, const void* w, void* y,
                 const ns_epilogue* ep, void* stream);
/* input gradient: dx (B,Tin,Cp) = conv_transpose(dz (B,Tout,N), w);  wt = (3, Cp, N) tap-major transposed weight.
 * ep may carry NS_ACT_DGELU (dx *= gelu'(aux_in)) so the previous conv's GELU backward is fused. */
int ns_conv3_dgrad(int dtype, int B, int Tin, int Cp, int N, int stride, const void* dz, const void* wt, void* dx,
                   const ns_epilogue* ep, void* stream);
/* weight gradient: dw (3, N, Cp) fp32 += sum_{b,t} dz[b,t,n] * x[b, stride*t + k - 1, c];  db (N) fp32 += sum dz */
int ns_conv3_wgrad(int dtype, int B, int Tin, int Cp, int N, int stride, const void* dz, const void* x, float* dw,
                   float* db, void* stream);

/* ---- LayerNorm over the last dim (eps 1e-5, affine).  HF modeling_whisper.py:393,403; utils/load_model.py:468.
 * mean/rstd (fp32, per row) are saved for the backward when non-NULL. */
int ns_layernorm_fwd(int dtype, long long rows, int d, const void* x, const float* gamma, const float* beta, void* y,
                     float* mean, float* rstd, float eps, void* stream);
/* dx = LN'(dy) (gamma/beta are frozen in the reference: no parameter grads).  If dres != NULL: dx += dres (residual). */
int ns_layernorm_bwd(int dtype, long long rows, int d, const void* dy, const void* x, const float* gamma,
                     const float* mean, const float* rstd, const void* dres, void* dx, void* stream);

/* ---- multi-head attention, head_dim 64|32, q pre-scaled (HF:215-238, :310).  Tensors are addressed with element strides
 *   q[b, i, h, :] = q + b*q_bs + i*q_rs + h*Dh   (same for k, v, o) so packed qkv / cross-KV buffers are read in place.
 *   causal != 0: key j visible to query i iff j <= i + (Lk - Lq).   lse (B,H,Lq) fp32 saved for the backward. */
typedef struct {
  int B, H, Lq, Lk, Dh, causal;
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;
} ns_attn_shape;
int ns_attention_fwd(int dtype, const ns_attn_shape* s, const void* q, const void* k, const void* v, void* o, float* lse,
                     void* stream);
/* do has o's strides; dq/dk/dv have q/k/v's strides.  delta (B,H,Lq) fp32 is scratch. */
int ns_attention_bwd(int dtype, const ns_attn_shape* s, const void* q, const void* k, const void* v, const void* o,
                     const void* d_o, const float* lse, float* delta, void* dq, void* dk, void* dv, void* stream);
/* Same gradients through the fused single-pass kernel (exp evaluated once, dQ accumulated in fp32 with TMA reduce-add) when
 * the shape qualifies (bf16, head_dim 64, non-causal) and `workspace` (1024-byte aligned device memory) holds at least
 * ns_attention_bwd_workspace_bytes(s) bytes; otherwise it behaves exactly like ns_attention_bwd.  The library never
 * allocates: the caller owns the workspace (contents are scratch). */
long long ns_attention_bwd_workspace_bytes(const ns_attn_shape* s);
/* Developer aid: when a device buffer of 4*512*2 int64 is registered, CTA (0,0,0) of the fused backward records a
 * (tag, clock64) timeline of its warp roles into it; NULL (the default) disables it. */
int ns_debug_attn_trace(long long* device_buffer);
int ns_attention_bwd_ws(int dtype, const ns_attn_shape* s, const void* q, const void* k, const void* v, const void* o,
                        const void* d_o, const float* lse, float* delta, void* dq, void* dk, void* dv, void* workspace,
                        long long workspace_bytes, void* stream);

/* ---- decoder embedding: h[b,l,:] = E[ids[b,l]] + P[pos0 + l]   (utils/load_model.py:646-660) */
int ns_embed(int dtype, int B, int L, int d, const long long* ids, const void* E, const void* P, int pos0, void* h,
             void* stream);
/* ---- cross-entropy over logits (rows, ld>=V): per-row loss (0 where label==-100), optional in-place dlogits
 *  (softmax - onehot) * (grad_scale / n_valid); pad columns [V, ld) are zeroed.  utils/load_model.py:1051-1054.
 *  n_valid is counted on device into *n_valid_out (int); loss_sum_out = sum of row losses (fp32). */
int ns_cross_entropy(int dtype, long long rows, int V, long long ld, void* logits, const long long* labels,
                     float* row_loss, float* loss_sum_out, int* n_valid_out, int write_grad, float grad_scale,
                     void* stream);
/* ---- greedy next token: argmax over logits[b, :V] with `suppress` ids (at most 64) at -inf, finished rows emit pad
 *  (GenerationMixin greedy + SuppressTokensAtBegin, HF generation_whisper.py:1774-1813).  out != NULL: the token is also
 *  appended to the output sequence, out[b * out_ld] (the caller passes &sequences[0][step]). */
int ns_greedy_pick(int dtype, int B, int V, long long ld, const void* logits, const int* suppress, int n_suppress,
                   int eos, int pad, unsigned char* finished, long long* next_ids, long long* out, long long out_ld,
                   void* stream);

/* ---- the decode loop in two calls (SURVEY.md 8b decode_prefill / decode_step; utils/load_model.py:1072-1351, :1332-1351
 *  prepare_inputs_for_generation feeds the last token only; HF modeling_whisper.py:314-336 KV cache).
 *  ns_decode_prefill: cross-attention K|V of every decoder layer from the encoder output, enc (B*S, d) -> layers[i].cross_kv.
 *  ns_decode_step   : ONE decoder pass for the token ids[b] at cache position pos (embedding + positions, per layer
 *                     LN -> q|k|v written into self_cache[:, pos] -> single-query self-attention -> out_proj + residual -> LN ->
 *                     cross-attention over cross_kv -> out_proj + residual -> LN -> fc1 + GELU -> fc2 + residual; final LN, tied
 *                     vocabulary projection) followed by the greedy pick (ns_greedy_pick semantics).  About 70 kernel launches
 *                     issued back to back from C++; nothing is allocated, every buffer is the caller's. */
typedef struct {
  const float* ln1_g; const float* ln1_b; const void* wqkv; const float* bqkv; const void* wo; const float* bo;
  const float* ln2_g; const float* ln2_b; const void* wqc; const float* bqc; const void* woc; const float* boc;
  const float* ln3_g; const float* ln3_b; const void* w1; const float* b1; const void* w2; const float* b2;
  const void* wkv; const float* bkv;   /* (2d, d) / (2d): cross K|V projection of this layer (prefill) */
  void* self_cache;                    /* (B, Tmax, 3d): q|k|v rows of this layer */
  void* cross_kv;                      /* (B*S, cross_ld >= 2d): K|V of the encoder output (unused in the absorbed form) */
  const void* wq_abs;                  /* absorbed form (below): (heads * d, d), the query AND key projections of this layer in one
                                          weight, wq_abs[h * d + n, m] = Dh^-0.5 * sum_c Wk[h * Dh + c, n] * Wq[h * Dh + c, m] */
  const float* bq_abs;                 /* (heads * d): bq_abs[h * d + n] = Dh^-0.5 * sum_c bq[h * Dh + c] * Wk[h * Dh + c, n] */
} ns_decoder_layer;
typedef struct {
  int dtype, n_layers, d, heads, ffn, vocab, S, Tmax, B, logits_dtype;
  long long cross_ld, logits_ld;
  const void* E; const void* pos_table; const float* lnf_g; const float* lnf_b;
  const ns_decoder_layer* layers;      /* host array of n_layers entries */
  void* h0; void* u; void* o; void* h1; void* qc; void* h2; void* mm; void* h3a; void* h3b; void* y;   /* (B, d) scratch; mm (B, ffn) */
  void* logits;                        /* (B, logits_ld) */
  /* Absorbed cross-attention (csrc/ns_attention_absorbed.cu): when enc, qp, cp and every layers[i].wq_abs / bq_abs are set
   * (bf16, d == 512, heads <= 8) the step never reads cross_kv: the key projection moves to the query side and folds into the
   * query projection (Q'_h = LN(x) wq_abs_h^T + bq_abs_h, ONE GEMM instead of the q projection), all heads attend over the
   * encoder rows themselves (C'_h = softmax(Q'_h enc^T) enc) and the value projection follows (out_h = C'_h Wv_h^T + bv_h,
   * block-diagonal GEMM on rows [d, 2d) of wkv) -- S*d instead of 2*S*d elements read per (sample, layer) and position, and
   * ns_decode_prefill has nothing to do. */
  const void* enc;                     /* (B*S, d) encoder output */
  void* qp; void* cp;                  /* (B, heads * d) scratch: Q' and C' */
} ns_decoder;
/* The middle part of the absorbed cross-attention on its own (tests, other callers): ctx[b, h, :] = softmax_j(qp[b, h, :] .
 * enc[b, j, :]) enc[b], all H <= 8 heads of a sample in one CTA.  qp / ctx rows of a sample are ldq / ldo elements apart
 * (>= H * d), enc samples enc_bs elements apart (rows d apart).  bf16, d == 512 only; NS_ERR_UNSUPPORTED otherwise. */
int ns_cross_attention_absorbed(int dtype, int B, int S, int H, int d, const void* qp, long long ldq, const void* enc, long long enc_bs,
                                void* ctx, long long ldo, void* stream);
int ns_decode_prefill(const ns_decoder* dec, const void* enc, void* stream);
int ns_decode_step(const ns_decoder* dec, const long long* ids, int pos, const int* suppress, int n_suppress, int eos, int pad,
                   unsigned char* finished, long long* next_ids, long long* out, long long out_ld, void* stream);

/* ---- EEG augmentation + pad + cast + layout pass (utils/reader.py:552-594, :496-506; utils/augment_eeg.py:15-26,54-56;
 *  utils/utils.py:33-60).  One read of x, one write of y.  x is either the collator's dense (B,C,Tin) batch (src_off NULL) or a
 *  RAGGED SAMPLE STORE resident in HBM (utils/reader.py:253-303 keeps recordings as unpadded .npy files): channel row c of
 *  batch slot b starts at x + src_off[b] + c * src_ld[b] (elements), n[b] valid samples; in_dtype says how x is stored.
 *   per sample b: n[b] valid samples, shift[b], taylor edges e0[b], e1[b]; keep-grid bits grid + b*grid_stride (row-major
 *   gc x gl bytes, NULL pointer or flags bit0 clear = no mask) expanded with rep_c[b], rep_t[b];
 *   noise (flags bit1): y = 2x + sigma[b,c]*N(0,1) (Philox, seed) -- the reference's 2x quirk kept.
 *   layout 0: y (B,C,T) like the reference collator;  layout 1: y (B,T,Cp) channels-last, pad channels zero. */
typedef struct {
  int B, C, Tin, T, Cp, layout, out_dtype;
  const int* n; const int* shift; const int* e0; const int* e1; const int* flags;
  const unsigned char* grid; long long grid_stride; const int* gl; const int* rep_c; const int* rep_t;
  const float* sigma; unsigned long long seed;
  int in_dtype; const long long* src_off; const int* src_ld;
} ns_aug_args;
int ns_aug_pass(const ns_aug_args* a, const void* x, void* y, void* stream);
/* per-(b,c) mean square over the first n[b] samples (for the noise sigma): ms (B,C) fp32; same source addressing as ns_aug_pass */
int ns_channel_meansq(int dtype, int B, int C, int Tin, const int* n, const void* x, float* ms, const long long* src_off,
                      const int* src_ld, void* stream);

/* ---- small utilities */
int ns_cast(int src_dtype, int dst_dtype, long long n, const void* src, void* dst, void* stream);
/* dst (cols, rows_pad>=rows) = scale * src (rows, cols)^T, dst leading dim ldd; pad rows zero-filled up to ldd */
int ns_transpose(int src_dtype, int dst_dtype, int rows, int cols, const void* src, long long lds, void* dst,
                 long long ldd, float scale, void* stream);
/* n_jobs independent transposes (same dtypes) in ONE launch; `jobs` is a DEVICE array.  max_rows_pad / max_cols bound the
 * largest job's ldd / cols.  Used for the per-step refresh of the LoRA operand layouts (60 small matrices). */
typedef struct {
  const void* src;
  void* dst;
  int rows, cols;
  long long lds, ldd;
  float scale;
  int pad_;
} ns_transpose_job;
int ns_transpose_batched(int src_dtype, int dst_dtype, int n_jobs, int max_rows_pad, int max_cols,
                         const ns_transpose_job* jobs, void* stream);
/* conv weight re-layouts: w (N,C,3) fp32 -> (3,N,Cp) and (3,Cp,N) in dtype; and the fp32 gradient back (3,N,Cp)->(N,C,3) */
int ns_conv_weight_pack(int dtype, int N, int C, int Cp, const float* w, void* w_tap, void* w_tap_t, void* stream);
int ns_conv_weight_unpack_grad(int N, int C, int Cp, const float* dw_tap, float* dw, void* stream);
/* y = a + b elementwise (dtype) */
int ns_add(int dtype, long long n, const void* a, const void* b, void* y, void* stream);
/* dz = dy * gelu_erf'(z) elementwise: backward through F.gelu(conv2(.)) at utils/load_model.py:411 (the only GELU on the
 * path whose backward cannot ride in a GEMM epilogue) */
int ns_dgelu_mul(int dtype, long long n, const void* dy, const void* z, void* dz, void* stream);

/* ---- beam search (evaluation.py:370-385: generate(num_beams=5, repetition_penalty=5.0, no_repeat_ngram_size=2)).
 * ns_beam_row_topk: for every beam row the C (= 2 * num_beams <= 16) best continuations of
 *   run_score[row] + processors(log_softmax(logits[row]))   with the processors of the reference's generate call in HF order:
 * repetition penalty over the tokens in seqs[row, :t] (lp < 0: lp * penalty, else lp / penalty), no-repeat-n-gram ban, and
 * `suppress` (begin_suppress_tokens, first generated position only) -- one pass over the logits (history as shared-memory bitmaps
 * over the vocabulary, online softmax, per-thread partial top lists).  Outputs sorted by score (descending).  The per-sample merge
 * of num_beams rows is a 2K x K problem left to the host loop.
 * ns_attention_decode_rows: single-query attention whose key/value j of batch row b is read from cache row kv_row[b * kv_ld + j]:
 * the beam reorder (utils/load_model.py:1353-1360 _reorder_cache) permutes this table instead of copying every layer's cache. */
int ns_beam_row_topk(int dtype, int rows, int V, long long ld, const void* logits, const long long* seqs, long long lds, int t,
                     const float* run_score, float penalty, int ngram, const int* suppress, int n_suppress, int C,
                     float* out_score, int* out_tok, void* stream);
int ns_attention_decode_rows(int dtype, const ns_attn_shape* s, const void* q, const void* k, const void* v, void* o,
                             const int* kv_row, long long kv_ld, void* stream);

/* ---- LoRA branch: dropout and the rank-r products around it.
 * Replaces PEFT lora.Linear's  result += lora_B(lora_A(lora_dropout(x))) * scaling  (finetune.py:206-212: lora_dropout 0.05,
 * 0.1 for AdaLoRA) and its autograd backward.  The keep mask of a module is a counter-based bit plane, drawn once per step:
 *   bits[g][rows][(cols+31)/32] (32-bit words, row-major like the activation); bit (col % 32) of word (g, row, w = col / 32) set
 *   <=> element (row, col) of adapter g is dropped.  The 32 flags of a word are drawn together from 16 hashed words
 *   R_i = mix1(km + (i + 1) * 0xC2B2AE35), km = lowbias32((row * 0x9E3779B1) ^ (w * 0x85EBCA77) ^ *seed ^ salts[g]), mix1 = the
 *   first multiply-xorshift round of lowbias32, combined along the binary expansion of thr = round(p * 65536), least
 *   significant bit first: D = bit_i(thr) ? (D | R_i) : (D & R_i)  (P(flag) = thr / 65536).
 * `seed` is a DEVICE word (one per training step, advanced on the device so a replayed CUDA graph draws a new mask); `salts[g]`
 * identify the modules (crc32 of their names).  The kernels use the UNSCALED masked input: the caller folds 1/(1-p) into alpha
 * (t) and into dt.  G = adapters stacked on the same input (1, or 3 for q/k/v), r = rank. */
int ns_seed_advance(unsigned int* seed, void* stream);                          /* *seed = lowbias32(*seed + 0x9E3779B9) */
long long ns_dropout_bits_words(long long rows, int cols);                      /* words per adapter of the bit plane */
/* G <= 8 planes of the same shape in one launch (e.g. q, k, v, out_proj, fc1 of one encoder layer: all d_model columns wide) */
int ns_dropout_bits(long long rows, int cols, int G, const unsigned int* seed, const unsigned int* salts, float p,
                    unsigned int* bits, void* stream);
/* y = x with the dropped elements of ONE adapter's plane zeroed (materialised masked input: fp32 parity mode and tests) */
int ns_dropout_apply(int dtype, long long rows, int cols, const void* x, long long ldx, void* y, long long ldy,
                     const unsigned int* bits, void* stream);
/* t[M, G*r] = alpha * (x . keep_g) * A_g^T for the G stacked adapters A = [A_0; ..; A_{G-1}] (G*r, K), bf16 storage.
 * bits == NULL (evaluation, parity runs): plain t = alpha * x * A^T.  HBM-bound: x is read once. */
int ns_lora_down(long long M, int K, int G, int r, const void* x, long long ldx, const void* A, long long lda, void* t,
                 long long ldt, float alpha, const unsigned int* bits, void* stream);
/* dA[G*r, K] (fp32, row stride ldg) += dt_g^T * (x . keep_g): gradient of the stacked A_g, bf16 activations, x read once.
 * With dx != NULL the same pass also removes the dropped terms from the input gradient,
 *   dx[m,k] -= sum_g dropped_g(m,k) * (dt_g[m,:] . At[k, g*r:(g+1)*r]) (* gelu'(z[m,k]) when z != NULL),
 * which the input-gradient GEMM added as a K-segment as if nothing had been dropped (At = A^T, (K, ldat >= G*r), bf16). */
int ns_lora_da(long long M, int K, int G, int r, const void* x, long long ldx, const void* dt, long long lddt, float* dA,
               long long ldg, const unsigned int* bits, void* dx, long long lddx, const void* At, long long ldat, const void* z,
               long long ldz, void* stream);
/* the dx correction alone, any storage dtype (fp32 parity mode) */
int ns_lora_dx_fix(int dtype, long long rows, int K, int G, int r, void* dx, long long lddx, const void* dt, long long lddt,
                   const void* At, long long ldat, const unsigned int* bits, const void* z, long long ldz, void* stream);

/* ---- fused clip + AdamW over one flat fp32 parameter/gradient buffer (HF trainer.py:2493,1760; finetune.py:236-247).
 *  Step 1: ns_sumsq accumulates sum(g^2) into *out (caller zeroes).  Step 2: ns_adamw_clip reads *sumsq on device,
 *  scales g by min(1, max_norm/(sqrt(sumsq*gscale^2)+1e-6))*gscale and applies torch.optim.AdamW semantics. */
int ns_sumsq(long long n, const float* g, float* out, void* stream);
int ns_adamw_clip(long long n, float* p, const float* g, float* m, float* v, const float* sumsq, float gscale,
                  float max_norm, float lr, float beta1, float beta2, float eps, float wd, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEUSPEECH_B200_H */
