// Decode-step cross-attention WITHOUT materialised keys and values (utils/load_model.py:534-767 -> HF modeling_whisper.py
// WhisperAttention with is_cross_attention: K = enc Wk^T (no bias), V = enc Wv^T + bv, recomputed never, read every position).
// The greedy loop reads the cross K|V of every sample and decoder layer once per generated token: 2 * S * d elements per (sample,
// layer), 2.36 GB per position at B = 128 -- the whole cost of a position.  The projections are linear, so they move to the
// query side (the "weight absorption" of latent-attention decoders):
//     scores_h[j] = q_h . K_h[j] = (q_h Wk_h) . enc[j]                      Q'_h = q_h Wk_h             (H x d per sample)
//     out_h       = sum_j P_h[j] V_h[j] = (sum_j P_h[j] enc[j]) Wv_h^T + bv_h    C'_h = P_h enc         (H x d per sample)
// (softmax rows sum to one, so the value bias passes through).  The keys AND values of all heads and of every decoder layer are
// then the encoder output itself: S * d elements per (sample, layer) -- half the bytes -- and no cross K|V buffer at all.
// The two small projections run on the block-diagonal tcgen05 GEMM (ns_epilogue.a_group_cols); this file is the middle part:
//     C'[b, h, :] = softmax_j( Q'[b, h, :] . enc[b, j, :] ) enc[b]          one CTA per sample, all H <= 8 heads at once
// as a flash-decoding loop on mma.sync m16n8k16 (the 8 heads are the 8 valid rows of the 16-row A tile): the encoder rows stream
// through a 3-stage cp.async ring of 64-key tiles (64 KB each); phase 1: warp w scores keys [8w, 8w+8) of the tile over all 512
// dimensions; row maxima / sums meet in shared memory; phase 2: warp w accumulates dimensions [64w, 64w+64) of C' over the 64
// keys (P through shared memory as bf16, the tile read a second time with ldmatrix.trans).  d_model = 512 only (Whisper-base).
#include "ns_common.cuh"

namespace ns {

namespace {

constexpr int AB_D = 512;                 // model width = "head dimension" of the absorbed form
constexpr int AB_KT = 64;                 // keys per tile
constexpr int AB_STAGES = 3;
constexpr int AB_WARPS = 8;
constexpr int AB_ROW = AB_D * 2 + 16;     // bytes per shared-memory row: 16 bytes of padding -> conflict-free ldmatrix
constexpr int AB_PROW = AB_KT * 2 + 16;   // P rows
constexpr int AB_SMEM = AB_STAGES * AB_KT * AB_ROW + 8 * AB_ROW + 8 * AB_PROW + 2 * AB_WARPS * 8 * 4;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ab_cp16(uint32_t saddr, const void* g, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void ab_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void ab_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
__device__ __forceinline__ void ab_ldsm4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ab_ldsm4t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ab_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// qp (B, H, 512) bf16 (row stride ldq per sample), enc (B, S, 512) bf16 (sample stride enc_bs, row stride 512), out like qp
__global__ void __launch_bounds__(AB_WARPS * 32, 1) cross_absorbed_kernel(int S, int H, const __nv_bfloat16* __restrict__ qp, long long ldq,
                                                                          const __nv_bfloat16* __restrict__ enc, long long enc_bs,
                                                                          __nv_bfloat16* __restrict__ out, long long ldo) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* sE = smem;                                          // [STAGES][KT][AB_ROW]
  unsigned char* sQ = sE + AB_STAGES * AB_KT * AB_ROW;               // [8][AB_ROW]
  unsigned char* sP = sQ + 8 * AB_ROW;                               // [8][AB_PROW]
  float* sMax = reinterpret_cast<float*>(sP + 8 * AB_PROW);          // [WARPS][8]
  float* sSum = sMax + AB_WARPS * 8;                                 // [WARPS][8]
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const __nv_bfloat16* eb = enc + static_cast<long long>(b) * enc_bs;
  const int n_tiles = (S + AB_KT - 1) / AB_KT;

  auto issue_tile = [&](int t) {
    if (t < n_tiles) {
      const uint32_t st = smem_addr(sE) + (t % AB_STAGES) * (AB_KT * AB_ROW);
#pragma unroll
      for (int j = 0; j < (AB_KT * 64) / (AB_WARPS * 32); ++j) {
        const int i = threadIdx.x + j * (AB_WARPS * 32);
        const int row = i >> 6, ch = i & 63;
        const int key = t * AB_KT + row;
        const bool ok = key < S;
        ab_cp16(st + row * AB_ROW + ch * 16, eb + static_cast<long long>(ok ? key : 0) * AB_D + ch * 8, ok ? 16 : 0);
      }
    }
    ab_commit();
  };
  // Q' rows of the valid heads (rows >= H stay zero: they only feed accumulator rows nobody reads)
  for (int i = threadIdx.x; i < 8 * 64; i += AB_WARPS * 32) {
    const int row = i >> 6, ch = i & 63;
    ab_cp16(smem_addr(sQ) + row * AB_ROW + ch * 16, qp + static_cast<long long>(b) * ldq + static_cast<long long>(row < H ? row : 0) * AB_D + ch * 8,
            row < H ? 16 : 0);
  }
  issue_tile(0);                                                     // group 0 = Q' + tile 0
  issue_tile(1);

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;
  // ldmatrix lane addressing shared by the Q' / P (A operand) and phase-1 key (B operand) loads: row = lane % 8, 8 columns per matrix
  const uint32_t a_lane = static_cast<uint32_t>((lane & 7) * AB_ROW + (lane >> 3) * 16);
  const uint32_t p_lane = static_cast<uint32_t>((lane & 7) * AB_PROW + (lane >> 3) * 16);
  // phase-2 transposed loads: row = key (lane % 16), matrices 2, 3 = the next 8 dimensions
  const uint32_t t_lane = static_cast<uint32_t>((lane & 15) * AB_ROW + (lane >> 4) * 16);

  for (int t = 0; t < n_tiles; ++t) {
    ab_wait<AB_STAGES - 2>();
    __syncthreads();                                                 // tile t landed for everyone; the stage of tile t-1 is free
    issue_tile(t + AB_STAGES - 1);
    const uint32_t st = smem_addr(sE) + (t % AB_STAGES) * (AB_KT * AB_ROW);
    // ---- phase 1: scores of keys [8 warp, 8 warp + 8) x 8 heads over all 512 dimensions
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t qa = smem_addr(sQ) + a_lane;
    const uint32_t ka = st + warp * 8 * AB_ROW + a_lane;
#pragma unroll 4
    for (int k2 = 0; k2 < AB_D / 32; ++k2) {                         // two k steps (32 dimensions) per pair of ldmatrix.x4
      uint32_t a[4], bb[4];
      ab_ldsm4(a, qa + k2 * 64);
      ab_ldsm4(bb, ka + k2 * 64);
      ab_mma(s, a[0], 0u, a[1], 0u, bb[0], bb[1]);
      ab_mma(s, a[2], 0u, a[3], 0u, bb[2], bb[3]);
    }
    const int key0 = t * AB_KT + warp * 8 + 2 * tq;
    float s0 = key0 < S ? s[0] * kLog2e : -INFINITY;
    float s1 = key0 + 1 < S ? s[1] * kLog2e : -INFINITY;
    float mx = fmaxf(s0, s1);
    mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 2));
    if (tq == 0) sMax[warp * 8 + g] = mx;
    __syncthreads();
    float m_new = m_run;
#pragma unroll
    for (int w = 0; w < AB_WARPS; ++w) m_new = fmaxf(m_new, sMax[w * 8 + g]);
    const float alpha = exp2f(m_run - m_new);                        // first tile: exp2(-inf) = 0
    const float p0 = exp2f(s0 - m_new), p1 = exp2f(s1 - m_new);
    float rs = p0 + p1;
    rs += __shfl_xor_sync(0xFFFFFFFFu, rs, 1);
    rs += __shfl_xor_sync(0xFFFFFFFFu, rs, 2);
    if (tq == 0) sSum[warp * 8 + g] = rs;
    {
      const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
      *reinterpret_cast<__nv_bfloat162*>(sP + g * AB_PROW + (warp * 8 + 2 * tq) * 2) = pb;
    }
    __syncthreads();
    float ls = 0.f;
#pragma unroll
    for (int w = 0; w < AB_WARPS; ++w) ls += sSum[w * 8 + g];
    l_run = l_run * alpha + ls;
    m_run = m_new;
    // ---- phase 2: C'[heads][64 warp .. 64 warp + 64) += P (8 x 64 keys) * tile (64 keys x 64 dimensions)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= alpha; o[nt][1] *= alpha; }
    const uint32_t pa = smem_addr(sP) + p_lane;
    const uint32_t va = st + t_lane + warp * 128;
#pragma unroll
    for (int k2 = 0; k2 < AB_KT / 32; ++k2) {                        // 32 keys per ldmatrix.x4 of P
      uint32_t a[4];
      ab_ldsm4(a, pa + k2 * 64);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint32_t vrow = va + (k2 * 32 + kk * 16) * AB_ROW;
#pragma unroll
        for (int np = 0; np < 4; ++np) {                             // 16 dimensions (two n tiles) per transposed load
          uint32_t bb[4];
          ab_ldsm4t(bb, vrow + np * 32);
          ab_mma(o[2 * np], a[2 * kk], 0u, a[2 * kk + 1], 0u, bb[0], bb[1]);
          ab_mma(o[2 * np + 1], a[2 * kk], 0u, a[2 * kk + 1], 0u, bb[2], bb[3]);
        }
      }
    }
  }
  ab_wait<0>();
  if (g < H) {
    const float inv = 1.0f / l_run;
    __nv_bfloat16* orow = out + static_cast<long long>(b) * ldo + static_cast<long long>(g) * AB_D + warp * 64 + 2 * tq;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<__nv_bfloat162*>(orow + nt * 8) = __floats2bfloat162_rn(o[nt][0] * inv, o[nt][1] * inv);
  }
}

}  // namespace

// C'[b, h, :] = softmax_j(Q'[b, h, :] . enc[b, j, :]) enc[b]; Q' already carries the Dh^-0.5 of the query projection
int cross_attention_absorbed(int B, int S, int H, int d, const void* qp, long long ldq, const void* enc, long long enc_bs, void* out,
                             long long ldo, cudaStream_t st) {
  if (d != AB_D || H < 1 || H > 8 || S < 1 || ldq % 8 != 0 || ldo % 8 != 0 || enc_bs % 8 != 0) return NS_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(qp) | reinterpret_cast<uintptr_t>(enc) | reinterpret_cast<uintptr_t>(out)) & 15) return NS_ERR_UNSUPPORTED;
  static bool attr_done = false;
  if (!attr_done) {
    NS_CUDA(cudaFuncSetAttribute(cross_absorbed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    attr_done = true;
  }
  NS_CUDA(launch_pdl(cross_absorbed_kernel, dim3(B), dim3(AB_WARPS * 32), static_cast<size_t>(AB_SMEM), st, S, H, static_cast<const __nv_bfloat16*>(qp), ldq,
                     static_cast<const __nv_bfloat16*>(enc), enc_bs, static_cast<__nv_bfloat16*>(out), ldo));
  NS_LAUNCH_CHECK();
  count(C_ATTN_TC);
  return NS_OK;
}

}  // namespace ns
