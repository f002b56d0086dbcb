// C-ABI entry points: library state + GEMM / convolution dispatch between the tcgen05 and the SIMT kernels.
#include "ns_common.cuh"
#include "ns_gemm.cuh"

#include <atomic>
#include <stdarg.h>

namespace ns {

static thread_local char g_err[512] = "";
thread_local int g_path = NS_PATH_AUTO;
thread_local int g_pdl = 0;
static std::atomic<long long> g_counters[C_NUM];

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count(int which, long long n) { g_counters[which].fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int n = []() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    return v;
  }();
  return n;
}

int attention_fwd_simt(int dtype, const ns_attn_shape& s, const void* q, const void* k, const void* v, void* o, float* lse, cudaStream_t st);
int attention_decode_rows(int dtype, const ns_attn_shape& s, const void* q, const void* k, const void* v, void* o, const int* kv_row,
                          long long kv_ld, cudaStream_t st);
int attention_bwd_simt(int dtype, const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                       const float* lse, float* delta, void* dq, void* dk, void* dv, cudaStream_t st);
int attention_fwd_tc(const ns_attn_shape& s, const void* q, const void* k, const void* v, void* o, float* lse, cudaStream_t st);
int attention_bwd_tc(const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                     const float* lse, float* delta, void* dq, void* dk, void* dv, cudaStream_t st);

size_t attention_bwd_fused_ws(const ns_attn_shape& s);
int attention_bwd_fused(const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                        const float* lse, float* delta, void* dq, void* dk, void* dv, void* ws, size_t ws_bytes, cudaStream_t st);

int attention_bwd_smallq(const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                         const float* lse, float* delta, void* dq, void* dk, void* dv, cudaStream_t st);
void set_attn_trace(long long* p);
static bool want_fast(int dtype) { return dtype == NS_BF16 && g_path != NS_PATH_SIMT; }
static int fast_required_failed(const char* what) {
  set_error("%s: NS_PATH_FAST requested but the shape/dtype does not qualify for the tcgen05 path", what);
  return NS_ERR_UNSUPPORTED;
}

}  // namespace ns

using namespace ns;

extern "C" {

int ns_version(void) { return 100; }
const char* ns_last_error_string(void) { return g_err; }
int ns_set_path(int path) {
  const int prev = g_path;
  if (path >= NS_PATH_AUTO && path <= NS_PATH_FAST) g_path = path;
  return prev;
}
int ns_device_info(int* sms, int* cc_major, int* cc_minor) {
  int dev = 0;
  NS_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  NS_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sms) *sms = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return NS_OK;
}
int ns_get_counters(long long* counters, int n) {
  NS_CHECK_ARG(counters && n > 0, "ns_get_counters: bad arguments");
  for (int i = 0; i < n && i < C_NUM; ++i) counters[i] = g_counters[i].load();
  return NS_OK;
}
int ns_reset_counters(void) {
  for (int i = 0; i < C_NUM; ++i) g_counters[i].store(0);
  return NS_OK;
}

int ns_set_pdl(int on) {
  const int prev = g_pdl;
  g_pdl = on ? 1 : 0;
  return prev;
}

int ns_gemm_nt(int dtype, long long M, int N, int K, const void* A, long long lda, const void* W, long long ldw, void* D,
               long long ldd, const ns_epilogue* ep, const void* A2, long long lda2, const void* W2, long long ldw2,
               int K2, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_gemm_nt: bad dtype %d", dtype);
  NS_CHECK_ARG(M >= 0 && N >= 0 && K > 0 && A && W && D, "ns_gemm_nt: bad shape/pointers (M=%lld N=%d K=%d)", M, N, K);
  NS_CHECK_ARG(lda >= K && ldw >= K && ldd >= N, "ns_gemm_nt: leading dimension too small");
  NS_CHECK_ARG(!ep || ep->a_group_cols <= 0 || (N % ep->a_group_cols == 0 && lda >= static_cast<long long>(N / ep->a_group_cols) * K),
               "ns_gemm_nt: a_group_cols needs N a multiple of it and lda >= groups * K");
  NS_CHECK_ARG((A2 == nullptr) == (W2 == nullptr), "ns_gemm_nt: A2 and W2 must be given together");
  NS_CHECK_ARG(!A2 || K2 > 0, "ns_gemm_nt: K2 must be positive with A2");
  if (ep) NS_CHECK_ARG(ep->act != NS_ACT_DGELU || ep->aux_in, "ns_gemm_nt: NS_ACT_DGELU needs aux_in");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const EpiDev e = make_epi(ep, dtype);
  const int ngrp = (ep && A2) ? ep->a2_group_cols : 0;
  if (e.drop_bits) {
    // the masked second product exists on the tcgen05 path only: never fall through to a kernel that would ignore the mask
    if (!want_fast(dtype)) {
      set_error("ns_gemm_nt: drop_bits needs bf16 storage and the tcgen05 path");
      return NS_ERR_UNSUPPORTED;
    }
    return gemm_nt_fast(M, N, K, A, lda, W, ldw, D, ldd, e, A2, lda2, W2, ldw2, K2, ngrp, st);
  }
  if (want_fast(dtype)) {
    const int r = gemm_nt_fast(M, N, K, A, lda, W, ldw, D, ldd, e, A2, lda2, W2, ldw2, K2, ngrp, st);
    if (r != NS_ERR_UNSUPPORTED) return r;
    if (g_path == NS_PATH_FAST) return fast_required_failed("ns_gemm_nt");
  }
  if (e.a_group_cols > 0) {
    // block-diagonal main product off the tcgen05 path: one plain launch per group
    NS_CHECK_ARG(!A2 && N % e.a_group_cols == 0, "ns_gemm_nt: a_group_cols needs a single product and N a multiple of it");
    const size_t es = dsize(dtype), eo = ep->out_dtype == NS_F32 ? 4 : 2;
    ns_epilogue eg = *ep;
    eg.a_group_cols = 0;
    for (int g = 0; g < N / e.a_group_cols; ++g) {
      const int c0 = g * e.a_group_cols;
      eg.bias = ep->bias ? ep->bias + c0 : nullptr;
      eg.alpha_cols = ep->alpha_cols - c0 < 0 ? 0 : (ep->alpha_cols - c0 > e.a_group_cols ? e.a_group_cols : ep->alpha_cols - c0);
      NS_CHECK_ARG(!ep->residual && !ep->aux_in && !ep->aux_out, "ns_gemm_nt: a_group_cols with residual / aux is not supported off the fast path");
      const int r = ns_gemm_nt(dtype, M, e.a_group_cols, K, static_cast<const char*>(A) + static_cast<size_t>(g) * K * es, lda,
                               static_cast<const char*>(W) + static_cast<size_t>(c0) * ldw * es, ldw, static_cast<char*>(D) + static_cast<size_t>(c0) * eo,
                               ldd, &eg, nullptr, 0, nullptr, 0, 0, stream);
      if (r != NS_OK) return r;
    }
    return NS_OK;
  }
  SimtProg p;
  memset(&p, 0, sizeof(p));
  p.nseg = A2 ? 2 : 1;
  p.seg[0] = SimtSeg{A, W, lda, ldw, 0, 1, 0, 0, (int)M, K, 0, 0};
  if (A2) p.seg[1] = SimtSeg{A2, W2, lda2, ldw2, 0, 1, 0, 0, (int)M, K2, ngrp, K2};
  p.batches = 1; p.tout = (int)M; p.N = N;
  p.out_bs = 0; p.out_rs = 1; p.out_off = 0; p.ldd = ldd; p.D = D; p.epi = e;
  return launch_nt_simt(dtype, p, st);
}

int ns_gemm_tn(int dtype, long long M, int I, int J, const void* X, long long ldx, const void* Y, long long ldy, float* G,
               long long si, long long sj, float alpha, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_gemm_tn: bad dtype %d", dtype);
  NS_CHECK_ARG(M >= 0 && I > 0 && J > 0 && X && Y && G, "ns_gemm_tn: bad shape/pointers");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M == 0) return NS_OK;
  if (want_fast(dtype)) {
    const int r = gemm_tn_fast(M, I, J, X, ldx, Y, ldy, G, si, sj, alpha, st);
    if (r != NS_ERR_UNSUPPORTED) return r;
    if (g_path == NS_PATH_FAST) return fast_required_failed("ns_gemm_tn");
  }
  SimtTnProg p;
  memset(&p, 0, sizeof(p));
  p.X = X; p.Y = Y; p.ldx = ldx; p.ldy = ldy; p.batches = 1; p.tout = (int)M; p.x_bs = 0;
  p.y_bs = 0; p.y_rs = 1; p.y_rows = (int)M; p.ntaps = 1; p.I = I; p.J = J; p.si = si; p.sj = sj; p.stap = 0;
  p.G = G; p.alpha = alpha;
  return launch_tn_simt(dtype, p, st);
}

int ns_gemm_tn_grouped(int dtype, long long M, int I, int J, int groups, const void* X, long long ldx, const void* Y, long long ldy,
                       float* G, long long si, long long sj, const float* alphas, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_gemm_tn_grouped: bad dtype %d", dtype);
  NS_CHECK_ARG(M >= 0 && I > 0 && J > 0 && groups >= 1 && groups <= 4 && X && Y && G && alphas, "ns_gemm_tn_grouped: bad shape/pointers");
  NS_CHECK_ARG(ldx >= static_cast<long long>(groups) * I && ldy >= static_cast<long long>(groups) * J, "ns_gemm_tn_grouped: leading dimension too small");
  if (M == 0) return NS_OK;
  if (want_fast(dtype)) {
    const int r = gemm_tn_grouped_fast(M, I, J, groups, X, ldx, Y, ldy, G, si, sj, alphas, reinterpret_cast<cudaStream_t>(stream));
    if (r != NS_ERR_UNSUPPORTED) return r;
    if (g_path == NS_PATH_FAST) return fast_required_failed("ns_gemm_tn_grouped");
  }
  const size_t es = dsize(dtype);
  for (int g = 0; g < groups; ++g) {                      // one plain launch per group (SIMT or tcgen05, whichever takes the shape)
    const int r = ns_gemm_tn(dtype, M, I, J, static_cast<const char*>(X) + static_cast<size_t>(g) * I * es, ldx,
                             static_cast<const char*>(Y) + static_cast<size_t>(g) * J * es, ldy, G + static_cast<long long>(g) * I * si, si, sj,
                             alphas[g], stream);
    if (r != NS_OK) return r;
  }
  return NS_OK;
}

int ns_lora_bwd_b(int dtype, long long M, int N, int r, int groups, const void* dy, long long lddy, const void* Bt, long long ldbt,
                  const void* t, long long ldt, void* dt, long long lddt, float* dB, const float* alpha_dt, const float* alpha_db,
                  void* workspace, long long workspace_bytes, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_lora_bwd_b: bad dtype %d", dtype);
  NS_CHECK_ARG(M >= 0 && N > 0 && r > 0 && groups >= 1 && groups <= 4 && dy && Bt && t && dt && dB && alpha_dt && alpha_db,
               "ns_lora_bwd_b: bad shape/pointers");
  NS_CHECK_ARG(lddy >= static_cast<long long>(groups) * N && ldbt >= N && ldt >= static_cast<long long>(groups) * r &&
                   lddt >= static_cast<long long>(groups) * r,
               "ns_lora_bwd_b: leading dimension too small");
  if (M == 0) return NS_OK;
  if (!want_fast(dtype)) {
    set_error("ns_lora_bwd_b: bf16 storage and the tcgen05 path only");
    return NS_ERR_UNSUPPORTED;
  }
  set_error("ns_lora_bwd_b: shape does not qualify (r == 32, N %% 128 == 0, at most 4 column parts of 14 chunks, 16-byte aligned operands)");
  return lora_bwd_b_fast(M, N, r, groups, dy, lddy, Bt, ldbt, t, ldt, dt, lddt, dB, alpha_dt, alpha_db, workspace, workspace_bytes,
                         reinterpret_cast<cudaStream_t>(stream));
}

long long ns_lora_bwd_b_workspace_bytes(long long M, int N, int r, int groups) { return lora_bwd_b_workspace_bytes(M, N, r, groups); }

int ns_gemm_tn_masked(int dtype, long long M, int I, int J, const void* X, long long ldx, const void* Y, long long ldy, float* G,
                      long long si, long long sj, float alpha, const unsigned int* xbits, long long xbits_ld, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_gemm_tn_masked: bad dtype %d", dtype);
  NS_CHECK_ARG(M >= 0 && I > 0 && J > 0 && X && Y && G && xbits, "ns_gemm_tn_masked: bad shape/pointers");
  if (M == 0) return NS_OK;
  if (!want_fast(dtype)) {
    set_error("ns_gemm_tn_masked: bf16 storage and the tcgen05 path only");
    return NS_ERR_UNSUPPORTED;
  }
  set_error("ns_gemm_tn_masked: operands do not qualify for the tcgen05 path (alignment / sizes)");
  return gemm_tn_fast(M, I, J, X, ldx, Y, ldy, G, si, sj, alpha, reinterpret_cast<cudaStream_t>(stream), xbits, xbits_ld);
}

int ns_conv3_fwd(int dtype, int B, int Tin, int Cp, int N, int stride, const void* x, const void* w, void* y,
                 const ns_epilogue* ep, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_conv3_fwd: bad dtype");
  NS_CHECK_ARG(B > 0 && Tin > 0 && Cp > 0 && N > 0 && (stride == 1 || stride == 2) && Tin % stride == 0 && x && w && y,
               "ns_conv3_fwd: bad arguments (B=%d Tin=%d Cp=%d N=%d stride=%d)", B, Tin, Cp, N, stride);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const EpiDev e = make_epi(ep, dtype);
  if (want_fast(dtype)) {
    const int r = conv3_fwd_fast(B, Tin, Cp, N, stride, x, w, y, e, st);
    if (r != NS_ERR_UNSUPPORTED) return r;
    if (g_path == NS_PATH_FAST) return fast_required_failed("ns_conv3_fwd");
  }
  const int Tout = Tin / stride;
  const size_t es = dsize(dtype);
  SimtProg p;
  memset(&p, 0, sizeof(p));
  p.nseg = 3;
  for (int k = 0; k < 3; ++k) {
    // input row stride*t + k - 1 = stride*(t + off) + add
    int off, add;
    if (stride == 1) { off = k - 1; add = 0; }
    else { off = (k == 0) ? -1 : 0; add = (k == 1) ? 0 : 1; }
    const char* wk = reinterpret_cast<const char*>(w) + static_cast<size_t>(k) * N * Cp * es;
    p.seg[k] = SimtSeg{x, wk, Cp, Cp, Tin, stride, add, off, Tout, Cp, 0, 0};
  }
  p.batches = B; p.tout = Tout; p.N = N;
  p.out_bs = Tout; p.out_rs = 1; p.out_off = 0; p.ldd = N; p.D = y; p.epi = e;
  return launch_nt_simt(dtype, p, st);
}

int ns_conv3_dgrad(int dtype, int B, int Tin, int Cp, int N, int stride, const void* dz, const void* wt, void* dx,
                   const ns_epilogue* ep, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_conv3_dgrad: bad dtype");
  NS_CHECK_ARG(B > 0 && Tin > 0 && Cp > 0 && N > 0 && stride == 2 && Tin % 2 == 0 && dz && wt && dx,
               "ns_conv3_dgrad: only stride 2 is on the path (conv A's input needs no gradient)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const EpiDev e = make_epi(ep, dtype);
  if (want_fast(dtype)) {
    const int r = conv3_dgrad_fast(B, Tin, Cp, N, stride, dz, wt, dx, e, st);
    if (r != NS_ERR_UNSUPPORTED) return r;
    if (g_path == NS_PATH_FAST) return fast_required_failed("ns_conv3_dgrad");
  }
  const int Tout = Tin / 2;
  const size_t es = dsize(dtype);
  auto tap = [&](int k) { return reinterpret_cast<const char*>(wt) + static_cast<size_t>(k) * Cp * N * es; };
  for (int par = 0; par < 2; ++par) {
    SimtProg p;
    memset(&p, 0, sizeof(p));
    if (par == 0) {
      p.nseg = 1;
      p.seg[0] = SimtSeg{dz, tap(1), N, N, Tout, 1, 0, 0, Tout, N, 0, 0};
    } else {
      p.nseg = 2;
      p.seg[0] = SimtSeg{dz, tap(2), N, N, Tout, 1, 0, 0, Tout, N, 0, 0};
      p.seg[1] = SimtSeg{dz, tap(0), N, N, Tout, 1, 0, 1, Tout, N, 0, 0};
    }
    p.batches = B; p.tout = Tout; p.N = Cp;
    p.out_bs = Tin; p.out_rs = 2; p.out_off = par; p.ldd = Cp; p.D = dx; p.epi = e;
    const int r = launch_nt_simt(dtype, p, st);
    if (r) return r;
  }
  return NS_OK;
}

int ns_conv3_wgrad(int dtype, int B, int Tin, int Cp, int N, int stride, const void* dz, const void* x, float* dw,
                   float* db, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_conv3_wgrad: bad dtype");
  NS_CHECK_ARG(B > 0 && Tin > 0 && Cp > 0 && N > 0 && (stride == 1 || stride == 2) && Tin % stride == 0 && dz && x && dw,
               "ns_conv3_wgrad: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int Tout = Tin / stride;
  if (db) {
    const int r = launch_colsum(dtype, static_cast<long long>(B) * Tout, N, dz, N, db, st);
    if (r) return r;
  }
  if (want_fast(dtype)) {
    const int r = conv3_wgrad_fast(B, Tin, Cp, N, stride, dz, x, dw, st);
    if (r != NS_ERR_UNSUPPORTED) return r;
    if (g_path == NS_PATH_FAST) return fast_required_failed("ns_conv3_wgrad");
  }
  SimtTnProg p;
  memset(&p, 0, sizeof(p));
  p.X = dz; p.Y = x; p.ldx = N; p.ldy = Cp; p.batches = B; p.tout = Tout; p.x_bs = Tout;
  p.y_bs = Tin; p.y_rs = stride; p.y_rows = Tout;
  for (int k = 0; k < 3; ++k) {
    if (stride == 1) { p.y_off[k] = k - 1; p.y_add[k] = 0; }
    else { p.y_off[k] = (k == 0) ? -1 : 0; p.y_add[k] = (k == 1) ? 0 : 1; }
  }
  p.ntaps = 3; p.I = N; p.J = Cp; p.si = Cp; p.sj = 1; p.stap = static_cast<long long>(N) * Cp; p.G = dw; p.alpha = 1.0f;
  return launch_tn_simt(dtype, p, st);
}

static int check_attn(const ns_attn_shape* s) {
  NS_CHECK_ARG(s, "attention: null shape");
  NS_CHECK_ARG(s->B > 0 && s->H > 0 && s->Lq > 0 && s->Lk > 0 && (s->Dh == 64 || s->Dh == 32),
               "attention: bad shape B=%d H=%d Lq=%d Lk=%d Dh=%d (head_dim must be 32 or 64)", s->B, s->H, s->Lq, s->Lk, s->Dh);
  return NS_OK;
}

int ns_attention_fwd(int dtype, const ns_attn_shape* s, const void* q, const void* k, const void* v, void* o, float* lse,
                     void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && q && k && v && o, "ns_attention_fwd: bad arguments");
  if (int r = check_attn(s)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (want_fast(dtype) && s->Lq > 1) {       // single-query (decode) attention is a streaming kernel on the SIMT side
    const int r = attention_fwd_tc(*s, q, k, v, o, lse, st);
    if (r != NS_ERR_UNSUPPORTED) return r;
    if (g_path == NS_PATH_FAST) return fast_required_failed("ns_attention_fwd");
  }
  return attention_fwd_simt(dtype, *s, q, k, v, o, lse, st);
}

int ns_attention_decode_rows(int dtype, const ns_attn_shape* s, const void* q, const void* k, const void* v, void* o, const int* kv_row,
                             long long kv_ld, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && q && k && v && o && kv_row, "ns_attention_decode_rows: bad arguments");
  if (int r = check_attn(s)) return r;
  NS_CHECK_ARG(s->Lq == 1 && kv_ld >= s->Lk, "ns_attention_decode_rows: one query per row and a table of at least Lk entries per row");
  return attention_decode_rows(dtype, *s, q, k, v, o, kv_row, kv_ld, reinterpret_cast<cudaStream_t>(stream));
}

int ns_attention_bwd(int dtype, const ns_attn_shape* s, const void* q, const void* k, const void* v, const void* o,
                     const void* d_o, const float* lse, float* delta, void* dq, void* dk, void* dv, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && q && k && v && o && d_o && lse && delta && dq && dk && dv, "ns_attention_bwd: bad arguments");
  if (int r = check_attn(s)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (want_fast(dtype) && s->Lq <= 64) {      // short query axis (decoder cross-attention): one CTA per (batch, head)
    const int r = attention_bwd_smallq(*s, q, k, v, o, d_o, lse, delta, dq, dk, dv, st);
    if (r != NS_ERR_UNSUPPORTED) return r;
  }
  if (want_fast(dtype)) {
    const int r = attention_bwd_tc(*s, q, k, v, o, d_o, lse, delta, dq, dk, dv, st);
    if (r != NS_ERR_UNSUPPORTED) return r;
    if (g_path == NS_PATH_FAST) return fast_required_failed("ns_attention_bwd");
  }
  return attention_bwd_simt(dtype, *s, q, k, v, o, d_o, lse, delta, dq, dk, dv, st);
}

int ns_debug_attn_trace(long long* device_buffer) {
  set_attn_trace(device_buffer);
  return NS_OK;
}

long long ns_attention_bwd_workspace_bytes(const ns_attn_shape* s) {
  if (!s || check_attn(s)) return 0;
  return static_cast<long long>(attention_bwd_fused_ws(*s));
}

int ns_attention_bwd_ws(int dtype, const ns_attn_shape* s, const void* q, const void* k, const void* v, const void* o,
                        const void* d_o, const float* lse, float* delta, void* dq, void* dk, void* dv, void* workspace,
                        long long workspace_bytes, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && q && k && v && o && d_o && lse && delta && dq && dk && dv, "ns_attention_bwd_ws: bad arguments");
  if (int r = check_attn(s)) return r;
  if (want_fast(dtype) && workspace && workspace_bytes > 0) {
    const int r = attention_bwd_fused(*s, q, k, v, o, d_o, lse, delta, dq, dk, dv, workspace, static_cast<size_t>(workspace_bytes),
                                      reinterpret_cast<cudaStream_t>(stream));
    if (r != NS_ERR_UNSUPPORTED) return r;
  }
  return ns_attention_bwd(dtype, s, q, k, v, o, d_o, lse, delta, dq, dk, dv, stream);
}

}  // extern "C"
