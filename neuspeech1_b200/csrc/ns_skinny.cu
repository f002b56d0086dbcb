// Skinny GEMM for the one-token decoder step (utils/load_model.py:1332-1351: M = batch rows <= 128, N, K in {d_model, ffn}):
//
//     D[M, N] = epilogue( LN(X)[M, K] * W[N, K]^T )          LN optional (pre-LN decoder blocks: HF modeling_whisper.py:393-414)
//
// At M <= 128 the persistent tcgen05 kernel is all fixed cost: one row tile, 8..64 CTAs, ~10 us per call (barrier / TMEM /
// descriptor set-up, pipeline fill, epilogue through shared memory) plus ~5 us for the LayerNorm launch in front of it -- 37 + 19
// such launches per decoded position.  Here a CTA owns TN output columns: it copies the whole activation block (128 x K bf16,
// L2 resident) and its TN weight rows (the only HBM traffic) into shared memory with cp.async, normalises the rows in place
// (warp = 16 rows), and runs mma.sync m16n8k16 from ldmatrix fragments; bias / scale / GELU / residual on the fragments.  One
// launch per sub-block instead of two.  MEASURED: slower than the two calls it replaces at B = 128 (1.23 ms against 0.87 ms per
// position): with 16-column tiles only 32..128 CTAs run, each repeating the 128 KB activation copy and the LayerNorm, and the
// fixed cost per launch stays where it was.  ns_decode_step therefore uses it only with NS_SKINNY=1; the entry point stays for
// callers with a handful of rows, where one launch does beat two.
#include "ns_common.cuh"

#include <stdlib.h>

namespace ns {

struct SkinnyProg {
  int M, N, K;
  const __nv_bfloat16* x; long long ldx;
  const __nv_bfloat16* w; long long ldw;
  void* d; long long ldd; int out_f32;
  const float* gamma; const float* beta; float eps;       // LN over the K columns of x (gamma == nullptr: none)
  const float* bias; float alpha; int alpha_cols; int act; // act: NS_ACT_NONE | NS_ACT_GELU
  const __nv_bfloat16* residual; long long ldr;
};

constexpr int kSkM = 128, kSkKC = 512, kSkPitch = (kSkKC + 8) * 2;   // bytes; +16 B per row: conflict-free ldmatrix

__device__ __forceinline__ void sk_cp16(uint32_t saddr, const void* g, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void sk_ldsm4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void sk_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int TN>
__global__ void __launch_bounds__(256) skinny_gemm_kernel(const SkinnyProg p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = TN / 8;
  unsigned char* xs = smem_raw;                              // [128][pitch]
  unsigned char* wsm = smem_raw + kSkM * kSkPitch;           // [TN][pitch]
  const uint32_t xs_u = static_cast<uint32_t>(__cvta_generic_to_shared(xs)), ws_u = static_cast<uint32_t>(__cvta_generic_to_shared(wsm));
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TN;
  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool rows_live = warp * 16 < p.M;                    // this warp's 16 rows hold at least one real row

  for (int kc = 0; kc < p.K; kc += kSkKC) {
    const int kw = min(kSkKC, p.K - kc);                     // columns of this chunk (multiple of 16)
    const int vec = kw / 8;                                  // 16-byte vectors per row
    if (kc) __syncthreads();                                 // everyone is done with the previous chunk
    for (int i = threadIdx.x; i < kSkM * vec; i += 256) {
      const int r = i / vec, v = i - r * vec;
      const bool ok = r < p.M;
      sk_cp16(xs_u + r * kSkPitch + v * 16, p.x + (ok ? static_cast<long long>(r) * p.ldx + kc + v * 8 : 0), ok ? 16 : 0);
    }
    for (int i = threadIdx.x; i < TN * vec; i += 256) {
      const int r = i / vec, v = i - r * vec;
      const bool ok = n0 + r < p.N;
      sk_cp16(ws_u + r * kSkPitch + v * 16, p.w + (ok ? static_cast<long long>(n0 + r) * p.ldw + kc + v * 8 : 0), ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();
    if (p.gamma != nullptr && rows_live) {
      // LayerNorm of this warp's 16 rows in place (K <= 512: one chunk holds the whole row); same formulas as ln_fwd_bf16_kernel
      const int per = kw / 32;                               // elements per lane: lane + 32 j
      for (int rr = 0; rr < 16; ++rr) {
        __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(xs + (warp * 16 + rr) * kSkPitch);
        float v[kSkKC / 32];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kSkKC / 32; ++j) {
          v[j] = j < per ? __bfloat162float(row[lane + 32 * j]) : 0.f;
          s += v[j];
        }
        s = warp_sum(s);
        const float mean = s / static_cast<float>(kw);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < kSkKC / 32; ++j) {
          const float c = j < per ? v[j] - mean : 0.f;
          q = fmaf(c, c, q);
        }
        q = warp_sum(q);
        const float rstd = rsqrtf(q / static_cast<float>(kw) + p.eps);
#pragma unroll
        for (int j = 0; j < kSkKC / 32; ++j)
          if (j < per) {
            const int c = lane + 32 * j;
            row[c] = __float2bfloat16_rn(fmaf((v[j] - mean) * rstd, __ldg(p.gamma + c), __ldg(p.beta + c)));
          }
      }
      __syncwarp();
    }
    if (rows_live) {
      // A fragment rows: 16 * warp + (lane % 16), columns k0 + 8 * (lane / 16);  B: [n][k] rows, four 8x8 matrices per ldmatrix
      const uint32_t a_base = xs_u + (warp * 16 + (lane & 15)) * kSkPitch + (lane >> 4) * 16;
      const uint32_t b_base = ws_u + ((lane >> 4) * 8 + (lane & 7)) * kSkPitch + ((lane >> 3) & 1) * 16;
      for (int k0 = 0; k0 < kw; k0 += 16) {
        uint32_t a[4];
        sk_ldsm4(a, a_base + k0 * 2);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t b[4];
          sk_ldsm4(b, b_base + np * 16 * kSkPitch + k0 * 2);
          sk_mma(acc[2 * np], a, b[0], b[1]);
          sk_mma(acc[2 * np + 1], a, b[2], b[3]);
        }
      }
    }
  }
  if (!rows_live) return;
  // epilogue on the fragments: c0,c1 = row g, columns 2t, 2t+1; c2,c3 = row g + 8
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = n0 + nt * 8 + 2 * t;
    if (col >= p.N) continue;
    const bool two = col + 1 < p.N;
    const float b0 = p.bias ? __ldg(p.bias + col) : 0.f, b1 = (p.bias && two) ? __ldg(p.bias + col + 1) : 0.f;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int row = warp * 16 + g + 8 * hh;
      if (row >= p.M) continue;
      float y0 = acc[nt][2 * hh] + b0, y1 = acc[nt][2 * hh + 1] + b1;
      if (col < p.alpha_cols) y0 *= p.alpha;
      if (col + 1 < p.alpha_cols) y1 *= p.alpha;
      if (p.act == NS_ACT_GELU) { y0 = gelu_fast(y0); y1 = gelu_fast(y1); }
      if (p.residual) {
        const __nv_bfloat16* rp = p.residual + static_cast<long long>(row) * p.ldr + col;
        y0 += __bfloat162float(rp[0]);
        if (two) y1 += __bfloat162float(rp[1]);
      }
      if (p.out_f32) {
        float* dp = static_cast<float*>(p.d) + static_cast<long long>(row) * p.ldd + col;
        dp[0] = y0;
        if (two) dp[1] = y1;
      } else {
        __nv_bfloat16* dp = static_cast<__nv_bfloat16*>(p.d) + static_cast<long long>(row) * p.ldd + col;
        if (two && (reinterpret_cast<uintptr_t>(dp) & 3) == 0) *reinterpret_cast<uint32_t*>(dp) = pack_bf16x2(y0, y1);
        else { dp[0] = __float2bfloat16_rn(y0); if (two) dp[1] = __float2bfloat16_rn(y1); }
      }
    }
  }
}

// Returns NS_ERR_UNSUPPORTED when the shape does not qualify (the caller takes LayerNorm + ns_gemm_nt instead).
int skinny_gemm(long long M, int N, int K, const void* x, long long ldx, const float* gamma, const float* beta, float eps,
                const void* w, long long ldw, void* d, long long ldd, const ns_epilogue* ep, cudaStream_t st) {
  if (M <= 0 || M > kSkM || N <= 0 || N > 8192 || K <= 0 || K % 16 != 0 || ldx % 8 != 0 || ldw % 8 != 0) return NS_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(w) & 15) != 0) return NS_ERR_UNSUPPORTED;
  if (gamma && (K > kSkKC || K % 32 != 0 || !beta)) return NS_ERR_UNSUPPORTED;
  if (ep && (ep->aux_in || ep->aux_out || ep->res_mod || ep->drop_bits || ep->a_group_cols || ep->a2_group_cols ||
             (ep->act != NS_ACT_NONE && ep->act != NS_ACT_GELU)))
    return NS_ERR_UNSUPPORTED;
  SkinnyProg p;
  memset(&p, 0, sizeof(p));
  p.M = static_cast<int>(M); p.N = N; p.K = K;
  p.x = static_cast<const __nv_bfloat16*>(x); p.ldx = ldx;
  p.w = static_cast<const __nv_bfloat16*>(w); p.ldw = ldw;
  p.d = d; p.ldd = ldd; p.out_f32 = ep && ep->out_dtype == NS_F32;
  p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.alpha = 1.0f;
  if (ep) {
    p.bias = ep->bias; p.alpha = ep->alpha; p.alpha_cols = ep->alpha_cols; p.act = ep->act;
    p.residual = static_cast<const __nv_bfloat16*>(ep->residual); p.ldr = ep->ldr;
  }
  // narrow column tiles while the grid is small: every CTA re-reads the activation block from L2, the weights come once from HBM
  const bool narrow = N <= 1024;
  const int tn = narrow ? 16 : 32;
  const size_t smem = static_cast<size_t>(kSkM + tn) * kSkPitch;
  const unsigned grid = static_cast<unsigned>((N + tn - 1) / tn);
  static bool attr16 = false, attr32 = false;
  if (narrow) {
    if (!attr16) { NS_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kSkM + 16) * kSkPitch))); attr16 = true; }
    NS_CUDA(launch_pdl(skinny_gemm_kernel<16>, dim3(grid), dim3(256), smem, st, p));
  } else {
    if (!attr32) { NS_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kSkM + 32) * kSkPitch))); attr32 = true; }
    NS_CUDA(launch_pdl(skinny_gemm_kernel<32>, dim3(grid), dim3(256), smem, st, p));
  }
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

}  // namespace ns

using namespace ns;

extern "C" int ns_ln_gemm_nt(int dtype, long long M, int N, int K, const void* X, long long ldx, const float* gamma, const float* beta,
                             float eps, const void* W, long long ldw, void* D, long long ldd, const ns_epilogue* ep, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && M >= 0 && N > 0 && K > 0 && X && W && D, "ns_ln_gemm_nt: bad shape/pointers");
  NS_CHECK_ARG(ldx >= K && ldw >= K && ldd >= N, "ns_ln_gemm_nt: leading dimension too small");
  NS_CHECK_ARG((gamma == nullptr) == (beta == nullptr), "ns_ln_gemm_nt: gamma and beta go together");
  if (M == 0) return NS_OK;
  if (dtype != NS_BF16) {
    set_error("ns_ln_gemm_nt: bf16 storage only (fp32 takes ns_layernorm_fwd + ns_gemm_nt)");
    return NS_ERR_UNSUPPORTED;
  }
  const int r = skinny_gemm(M, N, K, X, ldx, gamma, beta, eps, W, ldw, D, ldd, ep, reinterpret_cast<cudaStream_t>(stream));
  if (r == NS_ERR_UNSUPPORTED) set_error("ns_ln_gemm_nt: shape not supported (M <= 128, N <= 8192, K %% 16 == 0, K <= 512 with LayerNorm, plain / GELU / bias / residual epilogue)");
  return r;
}
