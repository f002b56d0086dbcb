// Native decode loop body (decode_prefill / decode_step).  utils/load_model.py:1072-1351 -> GenerationMixin greedy:
// the reference spends one Python-dispatched ATen call per op and position; here a position is ONE C call that launches the
// ~70 kernels of the decoder pass + the greedy pick back to back from C++ (about 2 us of host time per launch: the host stays
// ahead of the ~0.9 ms of device time per position, so no CUDA graph per position is needed to keep the GPU busy).
#include "ns_common.cuh"
#include "ns_gemm.cuh"

#include <stdlib.h>

namespace ns {
int cross_attention_absorbed(int B, int S, int H, int d, const void* qp, long long ldq, const void* enc, long long enc_bs, void* out,
                             long long ldo, cudaStream_t st);   // ns_attention_absorbed.cu
}
using namespace ns;

extern "C" {

static ns_epilogue plain_epi(int dtype) {
  ns_epilogue e;
  memset(&e, 0, sizeof(e));
  e.alpha = 1.0f;
  e.out_dtype = dtype;
  return e;
}

// D = epi(LN(x) W^T) (gamma == NULL: no LayerNorm): one skinny-GEMM launch when the shape qualifies (bf16, M <= 128), else
// ns_layernorm_fwd into `scratch` followed by ns_gemm_nt
static int ln_gemm(int dt, int M, int N, int K, const void* x, const float* gamma, const float* beta, void* scratch, const void* W,
                   void* D, long long ldd, const ns_epilogue* e, void* stream) {
  static const bool skinny = getenv("NS_SKINNY") != nullptr;      // measured slower at B = 128 (see ns_skinny.cu): opt-in
  if (skinny && dt == NS_BF16) {
    const int r = skinny_gemm(M, N, K, x, K, gamma, beta, 1e-5f, W, K, D, ldd, e, reinterpret_cast<cudaStream_t>(stream));
    if (r != NS_ERR_UNSUPPORTED) return r;
  }
  const void* a = x;
  if (gamma) {
    const int r = ns_layernorm_fwd(dt, M, K, x, gamma, beta, scratch, nullptr, nullptr, 1e-5f, stream);
    if (r != NS_OK) return r;
    a = scratch;
  }
  return ns_gemm_nt(dt, M, N, K, a, K, W, K, D, ldd, e, nullptr, 0, nullptr, 0, 0, stream);
}

// The absorbed cross-attention (include/neuspeech_b200.h ns_decoder.enc) applies: every buffer and weight layout is there
static bool absorbed(const ns_decoder* dec) {
  if (dec->dtype != NS_BF16 || dec->d != 512 || dec->heads > 8 || !dec->enc || !dec->qp || !dec->cp) return false;
  for (int i = 0; i < dec->n_layers; ++i)
    if (!dec->layers[i].wq_abs || !dec->layers[i].bq_abs) return false;
  return true;
}

#define NS_TRY(expr)            \
  do {                          \
    const int _r = (expr);      \
    if (_r != NS_OK) return _r; \
  } while (0)

int ns_cross_attention_absorbed(int dtype, int B, int S, int H, int d, const void* qp, long long ldq, const void* enc, long long enc_bs,
                                void* ctx, long long ldo, void* stream) {
  NS_CHECK_ARG(B >= 0 && S > 0 && H > 0 && d > 0 && qp && enc && ctx, "ns_cross_attention_absorbed: bad shape/pointers");
  NS_CHECK_ARG(ldq >= static_cast<long long>(H) * d && ldo >= static_cast<long long>(H) * d && enc_bs >= static_cast<long long>(S) * d,
               "ns_cross_attention_absorbed: leading dimension too small");
  if (B == 0) return NS_OK;
  set_error("ns_cross_attention_absorbed: bf16, d_model 512, at most 8 heads, 16-byte aligned operands only");
  if (dtype != NS_BF16) return NS_ERR_UNSUPPORTED;
  return cross_attention_absorbed(B, S, H, d, qp, ldq, enc, enc_bs, ctx, ldo, reinterpret_cast<cudaStream_t>(stream));
}

int ns_decode_prefill(const ns_decoder* dec, const void* enc, void* stream) {
  NS_CHECK_ARG(dec && enc && dec->layers && dec->n_layers > 0, "ns_decode_prefill: null argument");
  const int d = dec->d;
  const long long M = static_cast<long long>(dec->B) * dec->S;
  if (absorbed(dec)) {
    NS_CHECK_ARG(enc == dec->enc, "ns_decode_prefill: the absorbed form attends over dec->enc itself; enc must be that buffer");
    return NS_OK;                                                  // keys and values ARE the encoder output: nothing to project
  }
  for (int i = 0; i < dec->n_layers; ++i) {
    const ns_decoder_layer& L = dec->layers[i];
    NS_CHECK_ARG(L.wkv && L.cross_kv, "ns_decode_prefill: layer %d has no cross K/V weights / buffer", i);
    ns_epilogue e = plain_epi(dec->dtype);
    e.bias = L.bkv;
    NS_TRY(ns_gemm_nt(dec->dtype, M, 2 * d, d, enc, d, L.wkv, d, L.cross_kv, dec->cross_ld, &e, nullptr, 0, nullptr, 0, 0, stream));
  }
  return NS_OK;
}

int ns_decode_step(const ns_decoder* dec, const long long* ids, int pos, const int* suppress, int n_suppress, int eos, int pad,
                   unsigned char* finished, long long* next_ids, long long* out, long long out_ld, void* stream) {
  NS_CHECK_ARG(dec && ids && next_ids && dec->layers && dec->n_layers > 0, "ns_decode_step: null argument");
  NS_CHECK_ARG(pos >= 0 && pos < dec->Tmax, "ns_decode_step: position %d outside the cache (Tmax %d)", pos, dec->Tmax);
  NS_CHECK_ARG(dec->d % dec->heads == 0, "ns_decode_step: d_model %d not divisible by %d heads", dec->d, dec->heads);
  const int dt = dec->dtype, B = dec->B, d = dec->d, H = dec->heads, F = dec->ffn, S = dec->S, Tmax = dec->Tmax;
  const int Dh = d / H;
  const float qscale = 1.0f / sqrtf(static_cast<float>(Dh));
  const size_t es = dsize(dt);
  const bool absorb = absorbed(dec);
  NS_TRY(ns_embed(dt, B, 1, d, ids, dec->E, dec->pos_table, pos, dec->h0, stream));
  const void* hd = dec->h0;
  ns_attn_shape ss{B, H, 1, pos + 1, Dh, 1, (long long)Tmax * 3 * d, 3LL * d, (long long)Tmax * 3 * d, 3LL * d, (long long)Tmax * 3 * d, 3LL * d, d, d};
  ns_attn_shape sc{B, H, 1, S, Dh, 0, d, d, (long long)S * dec->cross_ld, dec->cross_ld, (long long)S * dec->cross_ld, dec->cross_ld, d, d};
  for (int i = 0; i < dec->n_layers; ++i) {
    const ns_decoder_layer& L = dec->layers[i];
    char* cache = static_cast<char*>(L.self_cache);
    char* row = cache + static_cast<size_t>(pos) * 3 * d * es;          // (b, pos, :) = row + b * Tmax * 3d
    // self-attention: q|k|v of the new token go straight into the cache row (HF modeling_whisper.py:314-336)
    ns_epilogue e = plain_epi(dt);
    e.bias = L.bqkv; e.alpha = qscale; e.alpha_cols = d;
    NS_TRY(ln_gemm(dt, B, 3 * d, d, hd, L.ln1_g, L.ln1_b, dec->u, L.wqkv, row, (long long)Tmax * 3 * d, &e, stream));
    NS_TRY(ns_attention_fwd(dt, &ss, row, cache + static_cast<size_t>(d) * es, cache + static_cast<size_t>(2 * d) * es, dec->o, nullptr, stream));
    e = plain_epi(dt);
    e.bias = L.bo; e.residual = hd; e.ldr = d;
    NS_TRY(ln_gemm(dt, B, d, d, dec->o, nullptr, nullptr, nullptr, L.wo, dec->h1, d, &e, stream));
    // cross-attention: over the precomputed K|V of this layer, or (absorbed form) over the encoder rows themselves
    if (absorb) {
      // Q'_h = q_h Wk_h folded into the query projection (wq_abs), all heads over the encoder rows, then out_h = C'_h Wv_h^T +
      // bv_h (block-diagonal: output columns [h Dh, (h+1) Dh) contract C'[:, h d : (h+1) d])
      e = plain_epi(dt);
      e.bias = L.bq_abs;
      NS_TRY(ln_gemm(dt, B, H * d, d, dec->h1, L.ln2_g, L.ln2_b, dec->u, L.wq_abs, dec->qp, static_cast<long long>(H) * d, &e, stream));
      NS_TRY(cross_attention_absorbed(B, S, H, d, dec->qp, static_cast<long long>(H) * d, dec->enc, static_cast<long long>(S) * d, dec->cp,
                                      static_cast<long long>(H) * d, reinterpret_cast<cudaStream_t>(stream)));
      e = plain_epi(dt);
      e.a_group_cols = Dh;
      e.bias = L.bkv + d;
      NS_TRY(ns_gemm_nt(dt, B, d, d, dec->cp, static_cast<long long>(H) * d, static_cast<const char*>(L.wkv) + static_cast<size_t>(d) * d * es, d,
                        dec->o, d, &e, nullptr, 0, nullptr, 0, 0, stream));
    } else {
      e = plain_epi(dt);
      e.bias = L.bqc; e.alpha = qscale; e.alpha_cols = d;
      NS_TRY(ln_gemm(dt, B, d, d, dec->h1, L.ln2_g, L.ln2_b, dec->u, L.wqc, dec->qc, d, &e, stream));
      const char* ckv = static_cast<const char*>(L.cross_kv);
      NS_TRY(ns_attention_fwd(dt, &sc, dec->qc, ckv, ckv + static_cast<size_t>(d) * es, dec->o, nullptr, stream));
    }
    e = plain_epi(dt);
    e.bias = L.boc; e.residual = dec->h1; e.ldr = d;
    NS_TRY(ln_gemm(dt, B, d, d, dec->o, nullptr, nullptr, nullptr, L.woc, dec->h2, d, &e, stream));
    // MLP
    e = plain_epi(dt);
    e.bias = L.b1; e.act = NS_ACT_GELU;
    NS_TRY(ln_gemm(dt, B, F, d, dec->h2, L.ln3_g, L.ln3_b, dec->u, L.w1, dec->mm, F, &e, stream));
    void* hn = (i & 1) ? dec->h3b : dec->h3a;
    e = plain_epi(dt);
    e.bias = L.b2; e.residual = dec->h2; e.ldr = d;
    NS_TRY(ln_gemm(dt, B, d, F, dec->mm, nullptr, nullptr, nullptr, L.w2, hn, d, &e, stream));
    hd = hn;
  }
  NS_TRY(ns_layernorm_fwd(dt, B, d, hd, dec->lnf_g, dec->lnf_b, dec->y, nullptr, nullptr, 1e-5f, stream));
  ns_epilogue e = plain_epi(dec->logits_dtype);
  NS_TRY(ns_gemm_nt(dt, B, dec->vocab, d, dec->y, d, dec->E, d, dec->logits, dec->logits_ld, &e, nullptr, 0, nullptr, 0, 0, stream));
  return ns_greedy_pick(dec->logits_dtype, B, dec->vocab, dec->logits_ld, dec->logits, suppress, n_suppress, eos, pad, finished, next_ids, out, out_ld,
                        stream);
}

}  // extern "C"
