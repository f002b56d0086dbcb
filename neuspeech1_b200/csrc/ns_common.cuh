// Shared host/device helpers for libneuspeech_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/neuspeech_b200.h"

namespace ns {

// ---------------------------------------------------------------- error handling (no exceptions across the ABI)
void set_error(const char* fmt, ...);
extern thread_local int g_path;           // ns_path
enum Counter { C_GEMM_TC = 0, C_GEMM_SIMT = 1, C_ATTN_TC = 2, C_ATTN_SIMT = 3, C_OTHER = 4, C_WGRAD_TC = 5, C_NUM = 8 };
void count(int which, long long n = 1);
int sm_count();

// ---------------------------------------------------------------- programmatic dependent launch (ns_set_pdl)
// The decoder's one-token steps are ~70 tiny dependent kernels (M = batch rows): with the launch attribute below the next
// kernel's CTAs are scheduled and run their prologue (barrier init, TMEM allocation, descriptor prefetch) while the previous
// kernel is still running; every kernel that may be launched this way executes pdl_wait() before its first global-memory
// access (griddepcontrol.wait: all prerequisite grids complete and flushed; a no-op without the attribute).
extern thread_local int g_pdl;
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// <<<grid, block, smem, stream>>> with the programmatic-stream-serialization attribute when ns_set_pdl(1) is in effect
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define NS_CHECK_ARG(cond, ...)                       \
  do {                                                \
    if (!(cond)) {                                    \
      ns::set_error(__VA_ARGS__);                     \
      return NS_ERR_ARG;                              \
    }                                                 \
  } while (0)

#define NS_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ns::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NS_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define NS_LAUNCH_CHECK() NS_CUDA(cudaPeekAtLastError())

inline size_t dsize(int dt) { return dt == NS_BF16 ? 2 : 4; }
inline bool valid_dtype(int dt) { return dt == NS_F32 || dt == NS_BF16; }

// ---------------------------------------------------------------- device numeric helpers
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// d/dx [x * Phi(x)] = Phi(x) + x * phi(x)
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// GELU for the tcgen05 epilogues (bf16 outputs).  erf(z/sqrt(2)) ~= tanh(z * (a + b z^2 + c z^4)) with the three coefficients
// least-squares fitted to erf itself (NOT the classic two-term "tanh GELU"): max abs error 4.9e-5 on gelu and 1.3e-4 on
// its derivative over |z| <= 10, i.e. two orders of magnitude below a bf16 ulp of the result, plus the 2^-11 relative error
// of tanh.approx.  One MUFU + 7 FMA-pipe instructions per element instead of ~14 for a polynomial erf, which keeps the
// fc1 / conv epilogues shorter than their main loops.  z^2 is clamped at 36 (the quartic coefficient is negative); tanh
// has saturated to +-1 long before.  The exact erff stays in the SIMT path (fp32 parity mode).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kGeluA = 7.97704294e-01f, kGeluB = 3.68194288e-02f, kGeluC = -3.20606757e-04f;
__device__ __forceinline__ float gelu_fast(float z) {
  const float t = fminf(z * z, 36.0f);
  const float T = tanh_approx(z * fmaf(t, fmaf(t, kGeluC, kGeluB), kGeluA));
  const float hz = 0.5f * z;
  return fmaf(hz, T, hz);
}
// exact derivative of gelu_fast: 0.5 (1 + T) + 0.5 z (1 - T^2) u'(z),  u' = a + 3 b z^2 + 5 c z^4
__device__ __forceinline__ float dgelu_fast(float z) {
  const float t = fminf(z * z, 36.0f);
  const float T = tanh_approx(z * fmaf(t, fmaf(t, kGeluC, kGeluB), kGeluA));
  const float up = fmaf(t, fmaf(t, 5.0f * kGeluC, 3.0f * kGeluB), kGeluA);
  const float s = fmaf(-T, T, 1.0f);
  return fmaf(0.5f * z * s, up, fmaf(0.5f, T, 0.5f));
}

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): the same fp32 arithmetic in half the issue slots -- the GEMM epilogues
// are bound by instruction issue, not by the FMA pipe.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t f2_splat(float v) { return f2_pack(v, v); }
// tanh(u(z)) of a pair and the clamped z^2 it was built from: the part gelu_fast and dgelu_fast share
__device__ __forceinline__ uint64_t gelu_tanh2(uint64_t z, uint64_t& t) {
  float t0, t1, u0, u1;
  f2_unpack(f2_mul(z, z), t0, t1);
  t = f2_pack(fminf(t0, 36.0f), fminf(t1, 36.0f));
  const uint64_t pz = f2_fma(t, f2_fma(t, f2_splat(kGeluC), f2_splat(kGeluB)), f2_splat(kGeluA));
  f2_unpack(f2_mul(z, pz), u0, u1);
  return f2_pack(tanh_approx(u0), tanh_approx(u1));
}
// gelu_fast / dgelu_fast on a pair: bit-identical to the scalar forms (same operations, same order, same rounding)
__device__ __forceinline__ uint64_t gelu_fast2(uint64_t z) {
  uint64_t t;
  const uint64_t T = gelu_tanh2(z, t);
  const uint64_t hz = f2_mul(z, f2_splat(0.5f));
  return f2_fma(hz, T, hz);
}
// gelu_fast and its derivative of a pair from ONE tanh (the forward epilogue that saves gelu'(z) for the backward)
__device__ __forceinline__ uint64_t gelu_both2(uint64_t z, uint64_t& dg) {
  uint64_t t;
  const uint64_t T = gelu_tanh2(z, t);
  const uint64_t hz = f2_mul(z, f2_splat(0.5f));
  const uint64_t up = f2_fma(t, f2_fma(t, f2_splat(5.0f * kGeluC), f2_splat(3.0f * kGeluB)), f2_splat(kGeluA));
  const uint64_t s = f2_fma(f2_mul(T, f2_splat(-1.0f)), T, f2_splat(1.0f));
  dg = f2_fma(f2_mul(hz, s), up, f2_fma(f2_splat(0.5f), T, f2_splat(0.5f)));
  return f2_fma(hz, T, hz);
}
__device__ __forceinline__ uint64_t dgelu_fast2(uint64_t z) {
  uint64_t t;
  const uint64_t T = gelu_tanh2(z, t);
  const uint64_t up = f2_fma(t, f2_fma(t, f2_splat(5.0f * kGeluC), f2_splat(3.0f * kGeluB)), f2_splat(kGeluA));
  const uint64_t s = f2_fma(f2_mul(T, f2_splat(-1.0f)), T, f2_splat(1.0f));
  return f2_fma(f2_mul(f2_mul(z, f2_splat(0.5f)), s), up, f2_fma(f2_splat(0.5f), T, f2_splat(0.5f)));
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

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// Epilogue description in device form (shared by the SIMT and the tcgen05 GEMMs).
struct EpiDev {
  const float* bias;
  float alpha;
  int alpha_cols;
  int act;
  const void* aux_in;
  void* aux_out;
  long long ldaux;
  const void* residual;
  long long ldr;
  int res_mod;
  int out_f32;   // 1: D is fp32
  const uint32_t* drop_bits;   // masked second product / masked A operand (see ns_epilogue::drop_bits, drop_mode)
  long long drop_ld;
  int drop_mode;
  long long drop_gstride;
  int a_group_cols;            // block-diagonal main product (see ns_epilogue::a_group_cols)
  int aux_deriv;               // aux holds gelu'(z) instead of z (see ns_epilogue::aux_deriv)
  const uint32_t* drop_seed;   // drop_mode 2: the mask stage draws the planes (see ns_epilogue::drop_seed)
  uint32_t drop_salts[4];
  float drop_p;
};

inline EpiDev make_epi(const ns_epilogue* ep, int dtype) {
  EpiDev e;
  memset(&e, 0, sizeof(e));
  e.alpha = 1.0f;
  e.out_f32 = (dtype == NS_F32);
  if (ep) {
    e.bias = ep->bias; e.alpha = ep->alpha; e.alpha_cols = ep->alpha_cols; e.act = ep->act;
    e.aux_in = ep->aux_in; e.aux_out = ep->aux_out; e.ldaux = ep->ldaux;
    e.residual = ep->residual; e.ldr = ep->ldr; e.res_mod = ep->res_mod;
    e.out_f32 = (ep->out_dtype == NS_F32);
    e.drop_bits = ep->drop_bits; e.drop_ld = ep->drop_ld; e.drop_mode = ep->drop_mode; e.drop_gstride = ep->drop_gstride;
    e.a_group_cols = ep->a_group_cols; e.aux_deriv = ep->aux_deriv;
    if (ep->drop_mode == 2 && ep->drop_salts) {
      e.drop_seed = ep->drop_seed; e.drop_p = ep->drop_p;
      for (int i = 0; i < 4; ++i) e.drop_salts[i] = ep->drop_salts[i];     // the caller sizes the array for 4 adapters
    }
  }
  return e;
}

// Apply the element epilogue.  `T` = storage type of aux/residual (the activation dtype).
template <typename T>
__device__ __forceinline__ float epi_apply(const EpiDev& e, float acc, long long row, long long res_row, int col) {
  float x = acc;
  if (e.bias) x += __ldg(e.bias + col);
  if (col < e.alpha_cols) x *= e.alpha;
  if (e.act == NS_ACT_GELU) {
    if (e.aux_out) reinterpret_cast<T*>(e.aux_out)[row * e.ldaux + col] = from_f<T>(e.aux_deriv ? dgelu_erf(x) : x);
    x = gelu_erf(x);
  } else if (e.act == NS_ACT_DGELU) {
    const float a = to_f<T>(reinterpret_cast<const T*>(e.aux_in)[row * e.ldaux + col]);
    x *= e.aux_deriv ? a : dgelu_erf(a);
  }
  if (e.residual) x += to_f<T>(reinterpret_cast<const T*>(e.residual)[res_row * e.ldr + col]);
  return x;
}

}  // namespace ns
