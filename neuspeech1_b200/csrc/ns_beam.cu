// Beam-search scoring step (evaluation.py:370-385: generate(num_beams=5, repetition_penalty=5.0, no_repeat_ngram_size=2)).
//
// Per beam row the reference's generation loop (transformers GenerationMixin beam search) computes, over the whole vocabulary,
//   lp = log_softmax(logits);  repetition penalty on every token already in the row (lp < 0: lp * penalty, else lp / penalty);
//   no-repeat-n-gram ban (-inf);  begin-suppress (-inf, first generated position);  + running beam score;  top 2K over K * V.
// Done with ATen that is ~12 passes over the (B*K, V) matrix plus a radix select.  Here ONE pass per row:
//   * the row's token history is turned into two shared-memory bitmaps over the vocabulary (seen / banned),
//   * a single sweep over the logits keeps the online softmax statistics and, per thread, the best 2K RAW logits among the
//     tokens that are neither seen nor banned (log-softmax and "+ running score" are monotone: order by raw logit),
//   * the few seen tokens are scored exactly (they need the row's log-sum-exp, known after the sweep) and merged,
//   * the block selects the row's 2K best; the per-sample merge of K rows x 2K candidates is a 50-element problem left to the host
//     loop (neuspeech1_b200/generation.py).
// HBM-bound: the logits are read once (B*K * V * 2 bytes).
#include "ns_common.cuh"

namespace ns {

constexpr int BEAM_MAXC = 16;          // candidates per row (2K <= 16)
constexpr int BEAM_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(BEAM_THREADS) beam_row_topk_kernel(int V, long long ld, const T* __restrict__ logits,
                                                                    const long long* __restrict__ seqs, long long lds, int t,
                                                                    const float* __restrict__ run_score, float penalty, int ngram,
                                                                    const int* __restrict__ suppress, int n_sup, int C,
                                                                    float* __restrict__ out_score, int* __restrict__ out_tok) {
  extern __shared__ uint32_t sm[];
  const int words = (V + 31) >> 5;
  uint32_t* seen = sm;                                   // token occurs in the row (repetition penalty)
  uint32_t* banned = sm + words;                         // -inf
  float* c_val = reinterpret_cast<float*>(sm + 2 * words);          // [BEAM_THREADS * C + t] candidate scores
  int* c_tok = reinterpret_cast<int*>(c_val + BEAM_THREADS * BEAM_MAXC + t);
  __shared__ float red_m[BEAM_THREADS / 32], red_l[BEAM_THREADS / 32];
  __shared__ float sel_v[BEAM_THREADS / 32];
  __shared__ int sel_i[BEAM_THREADS / 32];
  __shared__ int n_seen_cand;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* x = logits + static_cast<long long>(row) * ld;
  const long long* sq = seqs + static_cast<long long>(row) * lds;

  for (int i = tid; i < 2 * words; i += BEAM_THREADS) sm[i] = 0u;
  if (tid == 0) n_seen_cand = 0;
  __syncthreads();
  for (int j = tid; j < t; j += BEAM_THREADS) {
    const int v = static_cast<int>(sq[j]);
    if (v >= 0 && v < V) {
      atomicOr(&seen[v >> 5], 1u << (v & 31));
      if (ngram == 1) atomicOr(&banned[v >> 5], 1u << (v & 31));
    }
  }
  if (ngram >= 2 && t + 1 >= ngram) {
    // windows of n-1 tokens starting at i whose continuation seqs[i + n - 1] exists, compared with the last n-1 tokens
    const int nwin = t - (ngram - 1);
    for (int i = tid; i < nwin; i += BEAM_THREADS) {
      bool match = true;
      for (int j = 0; j < ngram - 1; ++j) match = match && sq[i + j] == sq[t - (ngram - 1) + j];
      if (match) {
        const int v = static_cast<int>(sq[i + ngram - 1]);
        if (v >= 0 && v < V) atomicOr(&banned[v >> 5], 1u << (v & 31));
      }
    }
  }
  for (int j = tid; j < n_sup; j += BEAM_THREADS) {
    const int v = suppress[j];
    if (v >= 0 && v < V) atomicOr(&banned[v >> 5], 1u << (v & 31));
  }
  __syncthreads();

  // ---- one sweep: online softmax statistics over ALL tokens, top-C raw logits over the plain (not seen, not banned) ones
  float m = -INFINITY, l = 0.f;
  float tv[BEAM_MAXC];
  int ti[BEAM_MAXC];
#pragma unroll
  for (int i = 0; i < BEAM_MAXC; ++i) { tv[i] = -INFINITY; ti[i] = -1; }
  for (int v = tid; v < V; v += BEAM_THREADS) {           // lanes of a warp share one word of the bitmaps
    const float f = to_f<T>(x[v]);
    if (f > m) { l = l * __expf(m - f) + 1.f; m = f; } else { l += __expf(f - m); }
    const uint32_t skip = (seen[v >> 5] | banned[v >> 5]) >> (v & 31);
    if (!(skip & 1u) && f > tv[BEAM_MAXC - 1]) {
      float cv = f; int ci = v;                          // insert into the sorted (descending) list
#pragma unroll
      for (int i = 0; i < BEAM_MAXC; ++i) {
        if (cv > tv[i]) { const float a = tv[i]; const int bi = ti[i]; tv[i] = cv; ti[i] = ci; cv = a; ci = bi; }
      }
    }
  }
  // block reduction of (m, l)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
    const float mn = fmaxf(m, m2);
    l = (m == -INFINITY ? 0.f : l * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : l2 * __expf(m2 - mn));
    m = mn;
  }
  if (lane == 0) { red_m[warp] = m; red_l[warp] = l; }
  __syncthreads();
  float M = -INFINITY, L = 0.f;
  for (int i = 0; i < BEAM_THREADS / 32; ++i) {
    const float mn = fmaxf(M, red_m[i]);
    L = (M == -INFINITY ? 0.f : L * __expf(M - mn)) + (red_m[i] == -INFINITY ? 0.f : red_l[i] * __expf(red_m[i] - mn));
    M = mn;
  }
  const float lse = M + logf(L);
  const float base = run_score[row];
  // ---- candidates -> shared: the per-thread lists, then the seen tokens (each once: the thread that clears the bit owns it)
#pragma unroll
  for (int i = 0; i < BEAM_MAXC; ++i) {                  // (a thread's entries beyond the C best can never be selected)
    c_val[tid * BEAM_MAXC + i] = (ti[i] >= 0 && i < C) ? base + (tv[i] - lse) : -INFINITY;
    c_tok[tid * BEAM_MAXC + i] = ti[i];
  }
  __syncthreads();
  for (int j = tid; j < t; j += BEAM_THREADS) {
    const int v = static_cast<int>(sq[j]);
    if (v >= 0 && v < V) {
      const uint32_t bit = 1u << (v & 31);
      const uint32_t old = atomicAnd(&seen[v >> 5], ~bit);
      if ((old & bit) && !(banned[v >> 5] & bit)) {
        float lp = to_f<T>(x[v]) - lse;
        lp = lp < 0.f ? lp * penalty : lp / penalty;
        const int slot = atomicAdd(&n_seen_cand, 1);
        c_val[BEAM_THREADS * BEAM_MAXC + slot] = base + lp;
        c_tok[BEAM_THREADS * BEAM_MAXC + slot] = v;
      }
    }
  }
  __syncthreads();
  const int n_cand = BEAM_THREADS * BEAM_MAXC + n_seen_cand;
  // ---- C rounds of block-wide argmax (ties: lower token id first, like a stable sort by score)
  for (int r = 0; r < C; ++r) {
    float bv = -INFINITY; int bi = -1;
    for (int i = tid; i < n_cand; i += BEAM_THREADS) {
      const float v = c_val[i];
      if (v > bv || (v == bv && bi >= 0 && c_tok[i] >= 0 && c_tok[i] < c_tok[bi])) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v2 = __shfl_xor_sync(0xffffffffu, bv, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      if (i2 >= 0 && (bi < 0 || v2 > bv || (v2 == bv && c_tok[i2] < c_tok[bi]))) { bv = v2; bi = i2; }
    }
    if (lane == 0) { sel_v[warp] = bv; sel_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float fv = -INFINITY; int fi = -1;
      for (int i = 0; i < BEAM_THREADS / 32; ++i)
        if (sel_i[i] >= 0 && (fi < 0 || sel_v[i] > fv || (sel_v[i] == fv && c_tok[sel_i[i]] < c_tok[fi]))) { fv = sel_v[i]; fi = sel_i[i]; }
      out_score[static_cast<long long>(row) * C + r] = fi >= 0 ? fv : -INFINITY;
      out_tok[static_cast<long long>(row) * C + r] = fi >= 0 ? c_tok[fi] : 0;
      if (fi >= 0) c_val[fi] = -INFINITY;               // taken
    }
    __syncthreads();
  }
}

}  // namespace ns

using namespace ns;

extern "C" int ns_beam_row_topk(int dtype, int rows, int V, long long ld, const void* logits, const long long* seqs, long long lds, int t,
                                const float* run_score, float penalty, int ngram, const int* suppress, int n_suppress, int C,
                                float* out_score, int* out_tok, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_beam_row_topk: bad dtype %d", dtype);
  NS_CHECK_ARG(rows >= 0 && V > 0 && ld >= V && logits && seqs && t >= 0 && lds >= t && run_score && out_score && out_tok,
               "ns_beam_row_topk: bad shape/pointers");
  NS_CHECK_ARG(C >= 1 && C <= BEAM_MAXC, "ns_beam_row_topk: candidates per row must be in [1, %d] (2 x num_beams)", BEAM_MAXC);
  NS_CHECK_ARG(penalty > 0.f && ngram >= 0 && (n_suppress == 0 || suppress), "ns_beam_row_topk: bad penalty / n-gram / suppress arguments");
  if (rows == 0) return NS_OK;
  const int words = (V + 31) / 32;
  const size_t smem = static_cast<size_t>(2 * words) * 4 + static_cast<size_t>(BEAM_THREADS * BEAM_MAXC + t) * 8;
  NS_CHECK_ARG(smem <= 200 * 1024, "ns_beam_row_topk: vocabulary / history too large for the shared-memory bitmaps");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == NS_BF16) {
    static size_t attr = 0;
    if (smem > attr) { NS_CUDA(cudaFuncSetAttribute(beam_row_topk_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
    beam_row_topk_kernel<__nv_bfloat16><<<rows, BEAM_THREADS, smem, st>>>(V, ld, static_cast<const __nv_bfloat16*>(logits), seqs, lds, t, run_score,
                                                                         penalty, ngram, suppress, n_suppress, C, out_score, out_tok);
  } else {
    static size_t attr = 0;
    if (smem > attr) { NS_CUDA(cudaFuncSetAttribute(beam_row_topk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
    beam_row_topk_kernel<float><<<rows, BEAM_THREADS, smem, st>>>(V, ld, static_cast<const float*>(logits), seqs, lds, t, run_score, penalty, ngram,
                                                                 suppress, n_suppress, C, out_score, out_tok);
  }
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}
