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
// as a flash-decoding loop on mma.sync m16n8k16 (the 8 heads are the 8 valid rows of the 16-row A tile).  Two groups of four
// warps walk the even / odd 32-key tiles (32 KB each) with their own 3-stage cp.async ring, softmax state and named barrier, and
// merge at the end; per tile, phase 1: warp w scores keys [8w, 8w+8) over all 512 dimensions; row maxima / sums meet in shared
// memory; phase 2: warp w accumulates dimensions [128w, 128w+128) of C' over the tile's keys (P through shared memory as bf16,
// the tile read a second time with ldmatrix.trans).  d_model = 512 only (Whisper-base).
#include "ns_common.cuh"

namespace ns {

namespace {

constexpr int AB_D = 512;                 // model width = "head dimension" of the absorbed form
constexpr int AB_KT = 32;                 // keys per tile
constexpr int AB_GROUPS = 2;              // independent pipelines per CTA (4 warps each): one runs its MMAs while the other sits at a barrier
constexpr int AB_GW = 4;                  // warps per group
constexpr int AB_STAGES = 3;              // ring stages per group
constexpr int AB_ROW = AB_D * 2 + 16;     // bytes per shared-memory row: 16 bytes of padding -> conflict-free ldmatrix
constexpr int AB_PROW = AB_KT * 2 + 16;   // P rows
constexpr int AB_GROUP_BYTES = AB_STAGES * AB_KT * AB_ROW + 8 * AB_PROW + 2 * AB_GW * 8 * 4;   // ring + P + maxima + sums
constexpr int AB_SMEM = AB_GROUPS * AB_GROUP_BYTES + 8 * AB_ROW;                               // + Q'
static_assert(8 * AB_D * 4 + 2 * 8 * 4 * 32 <= AB_STAGES * AB_KT * AB_ROW, "the merge buffer reuses group 1's ring");

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ab_cp16(uint32_t saddr, const void* g, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void ab_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void ab_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
__device__ __forceinline__ void ab_group_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(AB_GW * 32) : "memory"); }
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

// qp (B, H, 512) bf16 (row stride ldq per sample), enc (B, S, 512) bf16 (sample stride enc_bs, row stride 512), out like qp.
// The two warp groups walk the even / odd 32-key tiles with their own cp.async ring, softmax state and named barrier -- a group
// waiting at one of its three barriers per tile leaves the tensor and load pipes to the other -- and merge at the end.
__global__ void __launch_bounds__(AB_GROUPS * AB_GW * 32, 1) cross_absorbed_kernel(int S, int H, const __nv_bfloat16* __restrict__ qp, long long ldq,
                                                                                    const __nv_bfloat16* __restrict__ enc, long long enc_bs,
                                                                                    __nv_bfloat16* __restrict__ out, long long ldo) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int grp = threadIdx.x >> 7;                                  // warp group
  const int gtid = threadIdx.x & 127;
  const int warp = gtid >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  unsigned char* sQ = smem + AB_GROUPS * AB_GROUP_BYTES;             // [8][AB_ROW], shared by both groups
  unsigned char* sE = smem + grp * AB_GROUP_BYTES;                   // [STAGES][KT][AB_ROW]
  unsigned char* sP = sE + AB_STAGES * AB_KT * AB_ROW;               // [8][AB_PROW]
  float* sMax = reinterpret_cast<float*>(sP + 8 * AB_PROW);          // [GW][8]
  float* sSum = sMax + AB_GW * 8;                                    // [GW][8]
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const __nv_bfloat16* eb = enc + static_cast<long long>(b) * enc_bs;
  const int n_tiles = (S + AB_KT - 1) / AB_KT;
  const int my_tiles = (n_tiles - grp + AB_GROUPS - 1) / AB_GROUPS;  // tiles grp, grp + 2, ...

  auto issue_tile = [&](int i) {                                     // i-th tile of this group
    if (i < my_tiles) {
      const int t = grp + i * AB_GROUPS;
      const uint32_t st = smem_addr(sE) + (i % AB_STAGES) * (AB_KT * AB_ROW);
#pragma unroll
      for (int j = 0; j < (AB_KT * 64) / (AB_GW * 32); ++j) {
        const int c = gtid + j * (AB_GW * 32);
        const int row = c >> 6, ch = c & 63;
        const int key = t * AB_KT + row;
        const bool ok = key < S;
        ab_cp16(st + row * AB_ROW + ch * 16, eb + static_cast<long long>(ok ? key : 0) * AB_D + ch * 8, ok ? 16 : 0);
      }
    }
    ab_commit();
  };
  // Q' rows of the valid heads (rows >= H stay zero: they only feed accumulator rows nobody reads); every thread's first group
  for (int i = threadIdx.x; i < 8 * 64; i += AB_GROUPS * AB_GW * 32) {
    const int row = i >> 6, ch = i & 63;
    ab_cp16(smem_addr(sQ) + row * AB_ROW + ch * 16, qp + static_cast<long long>(b) * ldq + static_cast<long long>(row < H ? row : 0) * AB_D + ch * 8,
            row < H ? 16 : 0);
  }
  issue_tile(0);                                                     // cp.async group 0 of every thread = its share of Q' + the first tile
  issue_tile(1);

  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;
  // ldmatrix lane addressing shared by the Q' / P (A operand) and phase-1 key (B operand) loads: row = lane % 8, 8 columns per matrix
  const uint32_t a_lane = static_cast<uint32_t>((lane & 7) * AB_ROW + (lane >> 3) * 16);
  const uint32_t p_lane = static_cast<uint32_t>((lane & 7) * AB_PROW + (lane >> 3) * 16);
  // phase-2 transposed loads: row = key (lane % 16), matrices 2, 3 = the next 8 dimensions
  const uint32_t t_lane = static_cast<uint32_t>((lane & 15) * AB_ROW + (lane >> 4) * 16);

  // Q' has to be visible to BOTH groups before anyone scores: every thread waits for its own copies, then one block barrier
  ab_wait<1>();
  __syncthreads();
  // Q' as A fragments in registers for the whole kernel (rows 0-7 only: a1 = a3 = 0): 64 registers instead of 8 KB of
  // shared-memory reads per warp and tile -- a third of the kernel's shared-memory traffic
  uint32_t qf[AB_D / 32][4];
#pragma unroll
  for (int k2 = 0; k2 < AB_D / 32; ++k2) ab_ldsm4(qf[k2], smem_addr(sQ) + a_lane + k2 * 64);
  for (int i = 0; i < my_tiles; ++i) {
    ab_wait<1>();                                                    // this thread's copies of tile i (tile i + 1 may be in flight)
    ab_group_sync(grp);                                              // tile i landed for the group; everyone left iteration i - 1
    issue_tile(i + 2);                                               // into the stage tile i - 1 was read from
    const int t = grp + i * AB_GROUPS;
    const uint32_t st = smem_addr(sE) + (i % AB_STAGES) * (AB_KT * AB_ROW);
    // ---- phase 1: scores of keys [8 warp, 8 warp + 8) x 8 heads over all 512 dimensions
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t ka = st + warp * 8 * AB_ROW + a_lane;
#pragma unroll
    for (int k2 = 0; k2 < AB_D / 32; ++k2) {                         // two k steps (32 dimensions) per ldmatrix.x4 of the keys
      uint32_t bb[4];
      ab_ldsm4(bb, ka + k2 * 64);
      ab_mma(s, qf[k2][0], 0u, qf[k2][1], 0u, bb[0], bb[1]);
      ab_mma(s, qf[k2][2], 0u, qf[k2][3], 0u, bb[2], bb[3]);
    }
    const int key0 = t * AB_KT + warp * 8 + 2 * tq;
    float s0 = key0 < S ? s[0] * kLog2e : -INFINITY;
    float s1 = key0 + 1 < S ? s[1] * kLog2e : -INFINITY;
    float mx = fmaxf(s0, s1);
    mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 2));
    if (tq == 0) sMax[warp * 8 + g] = mx;
    ab_group_sync(grp);
    float m_new = m_run;
#pragma unroll
    for (int w = 0; w < AB_GW; ++w) m_new = fmaxf(m_new, sMax[w * 8 + g]);
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
    ab_group_sync(grp);
    float ls = 0.f;
#pragma unroll
    for (int w = 0; w < AB_GW; ++w) ls += sSum[w * 8 + g];
    l_run = l_run * alpha + ls;
    m_run = m_new;
    // ---- phase 2: C'[heads][128 warp .. 128 warp + 128) += P (8 x 32 keys) * tile (32 keys x 128 dimensions)
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) { o[nt][0] *= alpha; o[nt][1] *= alpha; }
    uint32_t a[4];
    ab_ldsm4(a, smem_addr(sP) + p_lane);                             // all 32 keys of the tile: two k steps
    const uint32_t va = st + t_lane + warp * 256;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const uint32_t vrow = va + kk * 16 * AB_ROW;
#pragma unroll
      for (int np = 0; np < 8; ++np) {                               // 16 dimensions (two n tiles) per transposed load
        uint32_t bb[4];
        ab_ldsm4t(bb, vrow + np * 32);
        ab_mma(o[2 * np], a[2 * kk], 0u, a[2 * kk + 1], 0u, bb[0], bb[1]);
        ab_mma(o[2 * np + 1], a[2 * kk], 0u, a[2 * kk + 1], 0u, bb[2], bb[3]);
      }
    }
    // (the barrier at the top of the next iteration separates these reads of the stage and of P / maxima / sums from the
    // writes of iteration i + 1; the stage itself is refilled only after that barrier)
  }
  ab_wait<0>();
  __syncthreads();                                                   // both groups done with their rings
  // ---- merge: group 1 leaves (m, l, C') in its ring, group 0 combines and writes
  float* mo = reinterpret_cast<float*>(smem + AB_GROUP_BYTES);       // [16 n tiles][2][128 threads] accumulators, then [2][128] m, l
  if (grp == 1) {
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) { mo[(nt * 2) * 128 + gtid] = o[nt][0]; mo[(nt * 2 + 1) * 128 + gtid] = o[nt][1]; }
    mo[32 * 128 + gtid] = m_run;
    mo[33 * 128 + gtid] = l_run;
  }
  __syncthreads();
  if (grp == 0 && g < H) {
    const float m1 = mo[32 * 128 + gtid], l1 = mo[33 * 128 + gtid];
    const float m = fmaxf(m_run, m1);
    const float f0 = exp2f(m_run - m), f1 = exp2f(m1 - m);           // a group without tiles has m = -inf, l = 0: factor 0
    const float inv = 1.0f / (l_run * f0 + l1 * f1);
    __nv_bfloat16* orow = out + static_cast<long long>(b) * ldo + static_cast<long long>(g) * AB_D + warp * 128 + 2 * tq;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const float v0 = (o[nt][0] * f0 + mo[(nt * 2) * 128 + gtid] * f1) * inv;
      const float v1 = (o[nt][1] * f0 + mo[(nt * 2 + 1) * 128 + gtid] * f1) * inv;
      *reinterpret_cast<__nv_bfloat162*>(orow + nt * 8) = __floats2bfloat162_rn(v0, v1);
    }
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
  NS_CUDA(launch_pdl(cross_absorbed_kernel, dim3(B), dim3(AB_GROUPS * AB_GW * 32), static_cast<size_t>(AB_SMEM), st, S, H, static_cast<const __nv_bfloat16*>(qp), ldq,
                     static_cast<const __nv_bfloat16*>(enc), enc_bs, static_cast<__nv_bfloat16*>(out), ldo));
  NS_LAUNCH_CHECK();
  count(C_ATTN_TC);
  return NS_OK;
}

}  // namespace ns
