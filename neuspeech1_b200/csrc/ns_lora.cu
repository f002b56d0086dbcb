// LoRA side kernels (finetune.py:194-212): the branch dropout of PEFT's lora.Linear and the rank-r products around it.
//
//   y = base(x) + (alpha/r) * B(A(dropout_p(x)))          finetune.py:210  lora_dropout = 0.05 (0.1 for AdaLoRA, :206-207)
//
// The keep mask is a counter-based bit plane, row-major like the activation it masks: word (module, row, w) holds the DROPPED
// flags of columns [32 w, 32 w + 32) of that row, bit (col % 32).  One thread that owns a row of a tile (a tcgen05 epilogue
// lane, a mask-stage lane) reads its flags as consecutive words.  The 32 Bernoulli(p) bits of a word are drawn together from 16
// hashed words
//     km = lowbias32( (row * 0x9E3779B1) ^ (w * 0x85EBCA77) ^ seed ^ salt )
//     R_i = mix1( km + (i + 1) * 0xC2B2AE35 ),   i = 0..15        mix1(x): x ^= x >> 16; x *= 0x7FEB352D; x ^= x >> 15
// combined along the binary expansion of thr16 = round(p * 65536), least significant bit first:
//     D = 0;   D = bit_i(thr16) ? (D | R_i) : (D & R_i)         =>  every bit of D is set with probability thr16 / 65536
// (bit position b of (R_15 .. R_0) read as a 16-bit uniform number U_b: D_b = [U_b < thr16]) -- half a hash per element and no
// per-element compare.  The oracle restates exactly this (oracle/whisper_eeg.py::lora_dropout_keep).  A module's mask is needed
// three times per step (t forward, dA and the dx correction backward); hashing in every consumer made all three instruction-
// bound (ncu: 65.8 M warp instructions for the q/k/v down product, 58 % issue active, DRAM 11 %), so the plane is hashed ONCE
// per step (ns_dropout_bits, 1/16 of the activation's bytes) and every consumer reads bits.
// Convention: kernels work with the UNSCALED masked input x (.) keep; the 1/(1-p) lives in alpha' = (alpha/r)/(1-p), which
// scales t = alpha' (x.keep) A^T forward and dt' = alpha' g B backward, so dA = dt'^T (x.keep) and dx += (dt' A).keep.
//
//   ns_dropout_bits   bits[g][row][col / 32]: bit (col % 32) set <=> dropped
//   ns_lora_down      t[M, G*r] = alpha' * (x . keep_g) A_g^T      HBM-bound: x streams through registers once, A_g in shared
//                     memory, mma.sync m16n8k16 on register fragments (a rank-32 product cannot feed tcgen05's 128-row tiles
//                     from registers; the kernel is bound by the read of x, not by the tensor pipe)
//   ns_lora_da        dA_g[r, K] += dt'_g^T (x . keep_g)           split over row slabs, cp.async ring (x, dt', bits), mask applied
//                     on the ldmatrix fragments, fp32 vector reductions into the flat gradient buffer; the SAME pass removes the
//                     dropped terms from dx (below) when the caller hands it dx
//   ns_lora_dx_fix    dx[m,k] -= dropped_g(m,k) * (dt'_g[m,:] . A_g[:,k]) [* gelu'(z[m,k])]   sparse correction after the input-
//                     gradient GEMM, which carries the LoRA product as a K-segment as if nothing had been dropped (generic
//                     storage; the bf16 path does it inside ns_lora_da)
//   ns_dropout_apply  y = x . keep          (fp32 parity mode / reference path of the tests: materialises the masked input)
//   ns_seed_advance   seed <- lowbias32(seed + 0x9E3779B9)          (inside the captured training step: a new mask per replay)
#include "ns_common.cuh"
#include "ns_dropout.cuh"

#include <stdlib.h>

namespace ns {

// AND-mask for a packed bf16 pair from two drop bits: 0 where dropped
__device__ __forceinline__ uint32_t keep_mask2(uint32_t drop_lo, uint32_t drop_hi) {
  return ~(((drop_lo & 1u) | ((drop_hi & 1u) << 16)) * 0xFFFFu);
}

constexpr int kMaxPlanes = 8;
struct Salts { uint32_t s[kMaxPlanes]; };

// ------------------------------------------------------------------------------------------------ seed / mask planes
__global__ void seed_advance_kernel(uint32_t* seed) { *seed = lowbias32(*seed + 0x9E3779B9u); }

// one thread per output word = 32 columns of one row; consecutive threads write consecutive words of the plane
__global__ void __launch_bounds__(256) dropout_bits_kernel(long long rows, int cols, int words, const uint32_t* __restrict__ seed,
                                                           Salts salts, uint32_t thr, uint32_t* __restrict__ bits) {
  const int g = blockIdx.y;
  const uint32_t ms = *seed ^ salts.s[g];
  const long long total = rows * words;
  uint32_t* out = bits + static_cast<long long>(g) * total;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / words;
    const int w = static_cast<int>(i - row * words);
    const int valid = cols - w * 32;                              // columns of this block that exist
    const uint32_t vmask = valid < 32 ? (1u << valid) - 1u : 0xFFFFFFFFu;
    out[i] = drop_plane_word(static_cast<uint32_t>(row), static_cast<uint32_t>(w), ms, thr) & vmask;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_apply_kernel(long long rows, int cols, const T* __restrict__ x, long long ldx,
                                                            T* __restrict__ y, long long ldy, const uint32_t* __restrict__ bits) {
  const int words = (cols + 31) >> 5;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows * cols;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const uint32_t d = (bits[r * words + (c >> 5)] >> (c & 31)) & 1u;
    y[r * ldy + c] = d ? from_f<T>(0.f) : x[r * ldx + c];
  }
}

// ------------------------------------------------------------------------------------------------ dx correction, generic storage
template <typename T> __device__ __forceinline__ float dgelu_of(float z);
template <> __device__ __forceinline__ float dgelu_of<float>(float z) { return dgelu_erf(z); }
template <> __device__ __forceinline__ float dgelu_of<__nv_bfloat16>(float z) { return dgelu_fast(z); }

// One warp per row pair; lane = column (stride 32).  dt' rows of the pair sit in shared memory as fp32; At is A transposed,
// (K, ldat >= G*r): the r coefficients a dropped element needs are contiguous.  (fp32 parity mode and odd shapes.)
template <typename T>
__global__ void __launch_bounds__(256) lora_dx_fix_kernel(long long rows, int K, int G, int r, T* __restrict__ dx, long long lddx,
                                                          const T* __restrict__ dt, long long lddt, const T* __restrict__ At,
                                                          long long ldat, const uint32_t* __restrict__ bits,
                                                          const T* __restrict__ z, long long ldz) {
  extern __shared__ float s_dt[];                       // [8 warps][2 rows][G*r]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rp = static_cast<long long>(blockIdx.x) * 8 + warp;
  const long long r0 = rp * 2;
  if (r0 >= rows) return;
  const bool two = r0 + 1 < rows;
  const int gr = G * r;
  float* sd = s_dt + warp * 2 * gr;
  for (int j = lane; j < gr; j += 32) {
    sd[j] = to_f<T>(dt[r0 * lddt + j]);
    sd[gr + j] = two ? to_f<T>(dt[(r0 + 1) * lddt + j]) : 0.f;
  }
  __syncwarp();
  const int W = (K + 31) >> 5;
  for (int c = lane; c < K; c += 32) {
    float fix0 = 0.f, fix1 = 0.f;
    bool any0 = false, any1 = false;
    for (int g = 0; g < G; ++g) {
      const uint32_t* bw = bits + (static_cast<long long>(g) * rows + r0) * W + (c >> 5);
      const bool d0 = ((bw[0] >> (c & 31)) & 1u) != 0, d1 = two && ((bw[W] >> (c & 31)) & 1u) != 0;
      if (d0 | d1) {
        const T* a = At + static_cast<long long>(c) * ldat + g * r;
        float acc0 = 0.f, acc1 = 0.f;
        for (int j = 0; j < r; ++j) {
          const float av = to_f<T>(a[j]);
          acc0 = fmaf(sd[g * r + j], av, acc0);
          acc1 = fmaf(sd[gr + g * r + j], av, acc1);
        }
        if (d0) { fix0 += acc0; any0 = true; }
        if (d1) { fix1 += acc1; any1 = true; }
      }
    }
    if (any0) {
      if (z) fix0 *= dgelu_of<T>(to_f<T>(z[r0 * ldz + c]));
      dx[r0 * lddx + c] = from_f<T>(to_f<T>(dx[r0 * lddx + c]) - fix0);
    }
    if (any1) {
      if (z) fix1 *= dgelu_of<T>(to_f<T>(z[(r0 + 1) * ldz + c]));
      dx[(r0 + 1) * lddx + c] = from_f<T>(to_f<T>(dx[(r0 + 1) * lddx + c]) - fix1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ mma.sync helpers (bf16)
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// dx[0..1] += (a, b) as one packed bf16 reduction on global memory (no return value, no generic-address dispatch)
__device__ __forceinline__ void red_bf16x2(__nv_bfloat16* p, float a, float b) {
  const uint32_t v = pack_bf16x2(a, b);
  asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------ t = alpha' (x . keep) A^T
// Warp = MT 16-row tiles (fragment rows (g, g+8) <-> actual rows (2g, 2g+1) of the tile, so a thread owns both rows of a row
// pair).  Per 32-column chunk a thread loads 16 B of each of its rows; the 8 bf16 are consumed as the k-slots
// {2t,2t+1,2t+8,2t+9} of two MMAs (the k order inside a dot product is free as long as the A_g fragment uses the same order:
// the A_g fragment is the same 16 B of A_g's row).  A_g rows live in shared memory, pitch K*2+64 B (conflict-free LDS.128).
// D chunks of x are in flight per thread (register ring; ncu showed the two-deep version waiting on the long scoreboard).
template <int G, int NT, int WARPS, int MT, int D>
__global__ void __launch_bounds__(WARPS * 32) lora_down_kernel(long long M, int K, const __nv_bfloat16* __restrict__ x, long long ldx,
                                                               const __nv_bfloat16* __restrict__ A, long long lda,
                                                               __nv_bfloat16* __restrict__ t, long long ldt, float alpha,
                                                               const uint32_t* __restrict__ bits) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int R = NT * 8;
  const int pitch = K * 2 + 64;                                         // bytes
  for (int i = threadIdx.x; i < G * R * (K / 8); i += WARPS * 32) {     // A_g rows -> shared
    const int row = i / (K / 8), v = i - row * (K / 8);
    *reinterpret_cast<uint4*>(smem_raw + row * pitch + v * 16) = __ldg(reinterpret_cast<const uint4*>(A + row * lda) + v);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const long long n_warps = static_cast<long long>(gridDim.x) * WARPS;
  const long long rpw = (((M + n_warps - 1) / n_warps) + 1) & ~1LL;      // equal, contiguous, even-aligned row range per warp
  const long long r_begin = (static_cast<long long>(blockIdx.x) * WARPS + warp) * rpw;
  const long long r_end = r_begin + rpw < M ? r_begin + rpw : M;
  const int W = (K + 31) >> 5;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (long long R0 = r_begin; R0 < r_end; R0 += 16 * MT) {
    const bool two = MT == 2 && R0 + 16 < r_end;
    float acc[MT][G][NT][4];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
      for (int b = 0; b < G; ++b)
#pragma unroll
        for (int c = 0; c < NT; ++c)
#pragma unroll
          for (int d = 0; d < 4; ++d) acc[a][b][c][d] = 0.f;
    const __nv_bfloat16* xr[MT][2];
    bool ok[MT][2];
    const uint32_t* brow[MT][G];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const long long re = R0 + mt * 16 + 2 * g;
#pragma unroll
      for (int o = 0; o < 2; ++o) { ok[mt][o] = re + o < r_end; xr[mt][o] = x + (ok[mt][o] ? re + o : 0) * ldx + tq * 8; }
#pragma unroll
      for (int gi = 0; gi < G; ++gi)
        brow[mt][gi] = bits ? bits + (static_cast<long long>(gi) * M + (ok[mt][0] ? re : 0)) * W : nullptr;
    }
    uint4 buf[D][MT][2];
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int o = 0; o < 2; ++o)
          buf[d][mt][o] = (ok[mt][o] && d * 32 < K) ? __ldg(reinterpret_cast<const uint4*>(xr[mt][o] + d * 32)) : zero4;
    for (int cb = 0; cb < K; cb += 32 * D) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int c0 = cb + d * 32;
        if (c0 < K) {
          uint4 cur[MT][2];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              cur[mt][o] = buf[d][mt][o];
              buf[d][mt][o] = (ok[mt][o] && c0 + 32 * D < K) ? __ldg(reinterpret_cast<const uint4*>(xr[mt][o] + c0 + 32 * D)) : zero4;
            }
#pragma unroll
          for (int gi = 0; gi < G; ++gi) {
            uint32_t xe[MT][4], xo[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              if (mt == 1 && !two) continue;
              const uint32_t e[4] = {cur[mt][0].x, cur[mt][0].y, cur[mt][0].z, cur[mt][0].w};
              const uint32_t o[4] = {cur[mt][1].x, cur[mt][1].y, cur[mt][1].z, cur[mt][1].w};
              if (bits) {
                // the flags of this thread's 8 columns in its even and odd row (one word per row and 32-column chunk)
                const uint32_t be = __ldg(brow[mt][gi] + (c0 >> 5)) >> (8 * tq);
                const uint32_t bo = ok[mt][1] ? __ldg(brow[mt][gi] + W + (c0 >> 5)) >> (8 * tq) : 0u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  xe[mt][i] = e[i] & keep_mask2(be >> (2 * i), be >> (2 * i + 1));
                  xo[mt][i] = o[i] & keep_mask2(bo >> (2 * i), bo >> (2 * i + 1));
                }
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) { xe[mt][i] = e[i]; xo[mt][i] = o[i]; }
              }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const uint4 w = *reinterpret_cast<const uint4*>(smem_raw + (gi * R + nt * 8 + g) * pitch + (c0 + tq * 8) * 2);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                if (mt == 1 && !two) continue;
                mma_bf16_16816(acc[mt][gi][nt], xe[mt][0], xo[mt][0], xe[mt][1], xo[mt][1], w.x, w.y);
                mma_bf16_16816(acc[mt][gi][nt], xe[mt][2], xo[mt][2], xe[mt][3], xo[mt][3], w.z, w.w);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const long long re = R0 + mt * 16 + 2 * g;
#pragma unroll
      for (int gi = 0; gi < G; ++gi)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int col = gi * R + nt * 8 + tq * 2;
          if (re < r_end) *reinterpret_cast<uint32_t*>(t + re * ldt + col) = pack_bf16x2(alpha * acc[mt][gi][nt][0], alpha * acc[mt][gi][nt][1]);
          if (re + 1 < r_end) *reinterpret_cast<uint32_t*>(t + (re + 1) * ldt + col) = pack_bf16x2(alpha * acc[mt][gi][nt][2], alpha * acc[mt][gi][nt][3]);
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------ dA_g += dt'_g^T (x . keep_g)  (+ dx fix)
// CTA = (256-column slab of K, row slab), 16 warps x 16 columns.  D'[col, r] = sum_m x[m, col] dt'[m, r]: both operands are
// read "transposed" from their row-major shared tiles with ldmatrix.trans; a fragment register then holds the two rows of a row
// pair at one column.  cp.async ring of 32-row tiles of x, dt' and the mask words of the slab (32 rows x 8 words per adapter).
// FIX: the same pass removes the dropped terms from dx (the input-gradient GEMM added dt' A for every element).  Forming the
// whole rank-r product P = dt' A for the warp's 16 rows x 16 columns costs 4 MMAs per adapter (its A^T fragments stay in
// registers for the CTA's lifetime) -- cheaper than gathering r coefficients per dropped element with divergent lanes -- and P
// is kept only where a mask bit is set.  dx is updated with packed bf16 reductions (red.global.add.noftz.bf16x2: no load, no
// read-modify-write chain; a load-before-store per column tile chained DRAM round trips: ncu long-scoreboard 12.8 of 13 stall
// cycles, 193 us).  With z (backward through GELU: the correction carries gelu'(z)) the z tile of the slab rides in the ring next
// to x: a gather of z at the dropped positions (per-thread lists drained with six loads in flight) ran at 1.1 ms for K = 2048.
constexpr int DA_COLS = 256, DA_ROWS = 32, DA_WARPS = 16;
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
template <int G, int NT, bool FIX, bool ZG, int STAGES>
__global__ void __launch_bounds__(DA_WARPS * 32) lora_da_kernel(long long M, int K, long long rows_per_slab, const __nv_bfloat16* __restrict__ x,
                                                               long long ldx, const __nv_bfloat16* __restrict__ dt, long long lddt,
                                                               float* __restrict__ dA, long long ldg, const uint32_t* __restrict__ bits,
                                                               __nv_bfloat16* __restrict__ dx, long long lddx,
                                                               const __nv_bfloat16* __restrict__ At, long long ldat,
                                                               const __nv_bfloat16* __restrict__ z, long long ldz) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int R = NT * 8, GR = G * R, KS = (R + 15) / 16;
  constexpr int XP = (DA_COLS + 8) * 2;                 // x tile pitch (bytes): +16 B -> conflict-free ldmatrix
  constexpr int DP = (GR + 8) * 2;                      // dt' tile pitch
  constexpr int BITS_OFF = DA_ROWS * XP + DA_ROWS * DP; // mask words: [G][32 rows][8 words]
  constexpr int Z_OFF = BITS_OFF + G * 32 * 32;         // ZG: z tile, same pitch as x
  constexpr int STAGE = Z_OFF + (ZG ? DA_ROWS * XP : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int k0 = blockIdx.x * DA_COLS;
  const long long m_begin = static_cast<long long>(blockIdx.y) * rows_per_slab;
  const long long m_end = m_begin + rows_per_slab < M ? m_begin + rows_per_slab : M;
  const int n_tiles = m_end > m_begin ? static_cast<int>((m_end - m_begin + DA_ROWS - 1) / DA_ROWS) : 0;
  const uint32_t sbase = smem_u32(smem_raw);
  const int W = (K + 31) >> 5;

  auto issue = [&](int tile) {
    if (tile < n_tiles) {
      const uint32_t st = sbase + (tile % STAGES) * STAGE;
      const long long mrow = m_begin + static_cast<long long>(tile) * DA_ROWS;
      for (int i = threadIdx.x; i < DA_ROWS * (DA_COLS / 8); i += DA_WARPS * 32) {
        const int rr = i / (DA_COLS / 8), v = i - rr * (DA_COLS / 8);
        const bool ok = mrow + rr < m_end && k0 + v * 8 < K;
        cp_async16(st + rr * XP + v * 16, x + (ok ? (mrow + rr) * ldx + k0 + v * 8 : 0), ok ? 16 : 0);
        if constexpr (ZG) cp_async16(st + Z_OFF + rr * XP + v * 16, z + (ok ? (mrow + rr) * ldz + k0 + v * 8 : 0), ok ? 16 : 0);
      }
      for (int i = threadIdx.x; i < DA_ROWS * (GR / 8); i += DA_WARPS * 32) {
        const int rr = i / (GR / 8), v = i - rr * (GR / 8);
        const bool ok = mrow + rr < m_end;
        cp_async16(st + DA_ROWS * XP + rr * DP + v * 16, dt + (ok ? (mrow + rr) * lddt + v * 8 : 0), ok ? 16 : 0);
      }
      if (bits) {
        for (int i = threadIdx.x; i < G * 32 * 8; i += DA_WARPS * 32) {       // 32 rows x 8 words per adapter
          const int gi = i >> 8, rr = (i >> 3) & 31, v = i & 7;
          const bool ok = mrow + rr < M && (k0 >> 5) + v < W;
          cp_async4(st + BITS_OFF + (gi * 32 + rr) * 32 + v * 4, bits + (ok ? (static_cast<long long>(gi) * M + mrow + rr) * W + (k0 >> 5) + v : 0),
                    ok ? 4 : 0);
        }
      }
    }
    cp_async_commit();
  };

  float acc[G][NT][4];
#pragma unroll
  for (int a = 0; a < G; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

  const int wcol = warp * 16;                           // this warp's 16 columns inside the slab = one mask word
  // FIX: B fragments of P = dt' A for the warp's two 8-column tiles: B[k = rank][n = col] = At[col][rank]
  uint32_t bf[FIX ? G : 1][2][KS][2];
  if constexpr (FIX) {
#pragma unroll
    for (int gi = 0; gi < G; ++gi)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const int c = k0 + wcol + 8 * j + g;
          const __nv_bfloat16* ar = At + static_cast<long long>(c < K ? c : 0) * ldat + gi * R + ks * 16 + 2 * tq;
          bf[gi][j][ks][0] = c < K ? __ldg(reinterpret_cast<const uint32_t*>(ar)) : 0u;
          bf[gi][j][ks][1] = (c < K && ks * 16 + 8 < R) ? __ldg(reinterpret_cast<const uint32_t*>(ar + 8)) : 0u;
        }
  }
  for (int s = 0; s < STAGES - 1; ++s) issue(s);
  const int q = lane >> 3, rr8 = lane & 7;
  for (int tile = 0; tile < n_tiles; ++tile) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();                                    // tile `tile` has landed for everyone; stage (tile-1)%S is free again
    issue(tile + STAGES - 1);
    const uint32_t st = sbase + (tile % STAGES) * STAGE;
    const unsigned char* stp = smem_raw + (tile % STAGES) * STAGE;
    const long long mrow = m_begin + static_cast<long long>(tile) * DA_ROWS;
#pragma unroll
    for (int ks = 0; ks < DA_ROWS / 16; ++ks) {
      uint32_t a[4];
      ldsm_x4_trans(a, st + (ks * 16 + (q >> 1) * 8 + rr8) * XP + (wcol + (q & 1) * 8) * 2);
      float f[2][4];
      uint32_t fl[2] = {0u, 0u};
      if constexpr (FIX) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) f[j][i] = 0.f;
      }
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        uint32_t am[4];
        if (bits) {
          // dA side: fragment registers hold rows (2t, 2t+1) [a0, a1] and (2t+8, 2t+9) [a2, a3] at columns g [a0, a2], g+8 [a1, a3]
          const uint32_t* bw = reinterpret_cast<const uint32_t*>(stp + BITS_OFF + (gi * 32 + ks * 16 + 2 * tq) * 32) + (warp >> 1);
          const int sh = (warp & 1) * 16 + g;
          const uint32_t w0 = bw[0] >> sh, w1 = bw[8] >> sh, w2 = bw[64] >> sh, w3 = bw[72] >> sh;   // rows 2t, 2t+1, 2t+8, 2t+9
          am[0] = a[0] & keep_mask2(w0, w1);
          am[1] = a[1] & keep_mask2(w0 >> 8, w1 >> 8);
          am[2] = a[2] & keep_mask2(w2, w3);
          am[3] = a[3] & keep_mask2(w2 >> 8, w3 >> 8);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) am[i] = a[i];
        }
        const uint32_t dbase = st + DA_ROWS * XP + (ks * 16) * DP + (gi * R) * 2;
        if constexpr (NT >= 2) {
#pragma unroll
          for (int np = 0; np < NT / 2; ++np) {
            uint32_t b[4];
            ldsm_x4_trans(b, dbase + ((q & 1) * 8 + rr8) * DP + ((np * 2 + (q >> 1)) * 8) * 2);
            mma_bf16_16816(acc[gi][np * 2], am[0], am[1], am[2], am[3], b[0], b[1]);
            mma_bf16_16816(acc[gi][np * 2 + 1], am[0], am[1], am[2], am[3], b[2], b[3]);
          }
        } else {
          uint32_t b[2];
          ldsm_x2_trans(b, dbase + (((lane >> 3) & 1) * 8 + rr8) * DP);
          mma_bf16_16816(acc[gi][0], am[0], am[1], am[2], am[3], b[0], b[1]);
        }
        if constexpr (FIX) {
          // P = dt' A_g for 16 rows x 16 columns; fragment rows (g, g+8) <-> tile rows (2g, 2g+1): one row pair per thread
          float P[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) {
            uint32_t da[4];
            if (kk * 16 + 8 < R) {
              ldsm_x4(da, st + DA_ROWS * XP + (ks * 16 + 2 * rr8 + (q & 1)) * DP + (gi * R + kk * 16 + (q >> 1) * 8) * 2);
            } else {                                     // rank 8: only the low k half exists
              ldsm_x4(da, st + DA_ROWS * XP + (ks * 16 + 2 * rr8 + (q & 1)) * DP + (gi * R + kk * 16) * 2);
              da[2] = 0u; da[3] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) mma_bf16_16816(P[j], da[0], da[1], da[2], da[3], bf[gi][j][kk][0], bf[gi][j][kk][1]);
          }
          const uint32_t* bd = reinterpret_cast<const uint32_t*>(stp + BITS_OFF + (gi * 32 + ks * 16 + 2 * g) * 32) + (warp >> 1);
          const uint32_t we = bd[0] >> ((warp & 1) * 16), wo = bd[8] >> ((warp & 1) * 16);   // this thread's row pair
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int cj = 8 * j + 2 * tq;
            // bit0 (even row, col) bit1 (odd, col) bit2 (even, col+1) bit3 (odd, col+1)
            const uint32_t d = ((we >> cj) & 1u) | (((wo >> cj) & 1u) << 1) | (((we >> (cj + 1)) & 1u) << 2) | (((wo >> (cj + 1)) & 1u) << 3);
            f[j][0] += (d & 1u) ? P[j][0] : 0.f; f[j][1] += (d & 4u) ? P[j][1] : 0.f;
            f[j][2] += (d & 2u) ? P[j][2] : 0.f; f[j][3] += (d & 8u) ? P[j][3] : 0.f;
            fl[j] |= d;
          }
        }
      }
      if constexpr (FIX) {
        const long long re = mrow + ks * 16 + 2 * g;    // even row of this thread's pair
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col = k0 + wcol + 8 * j + 2 * tq;
          if (col < K) {
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              const bool need = (fl[j] & (o == 0 ? 5u : 10u)) != 0 && re + o < m_end;
              if (need) {
                const long long off = (re + o) * lddx + col;
                float f0 = f[j][2 * o], f1 = f[j][2 * o + 1];
                if constexpr (ZG) {
                  const float2 zf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(stp + Z_OFF + (ks * 16 + 2 * g + o) * XP + (wcol + 8 * j + 2 * tq) * 2));
                  f0 *= dgelu_fast(zf.x); f1 *= dgelu_fast(zf.y);
                }
                red_bf16x2(dx + off, -f0, -f1);
              }
            }
          }
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // accumulators -> shared [GR][DA_COLS] fp32 -> 16-byte vector reductions along the contiguous column axis of dA
  float* so = reinterpret_cast<float*>(smem_raw);
#pragma unroll
  for (int gi = 0; gi < G; ++gi)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int rr = gi * R + nt * 8 + tq * 2;
      so[rr * DA_COLS + wcol + g] = acc[gi][nt][0];
      so[(rr + 1) * DA_COLS + wcol + g] = acc[gi][nt][1];
      so[rr * DA_COLS + wcol + g + 8] = acc[gi][nt][2];
      so[(rr + 1) * DA_COLS + wcol + g + 8] = acc[gi][nt][3];
    }
  __syncthreads();
  if (n_tiles == 0) return;
  for (int i = threadIdx.x; i < GR * (DA_COLS / 4); i += DA_WARPS * 32) {
    const int rr = i / (DA_COLS / 4), c4 = (i - rr * (DA_COLS / 4)) * 4;
    if (k0 + c4 < K) {
      const float4 v = *reinterpret_cast<const float4*>(so + rr * DA_COLS + c4);
      float* dst = dA + static_cast<long long>(rr) * ldg + k0 + c4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  }
}

template <int G, int NT, bool FIX, bool ZG>
static int launch_da_variant(long long M, int K, const void* x, long long ldx, const void* dt, long long lddt, float* dA, long long ldg,
                             const uint32_t* bits, void* dx, long long lddx, const void* At, long long ldat, const void* z, long long ldz,
                             cudaStream_t st) {
  constexpr int R = NT * 8, GR = G * R;
  constexpr int STAGE = DA_ROWS * (DA_COLS + 8) * 2 * (ZG ? 2 : 1) + DA_ROWS * (GR + 8) * 2 + G * 32 * 32;
  constexpr int STAGES = (216 * 1024) / STAGE >= 8 ? 8 : (216 * 1024) / STAGE;
  static_assert(STAGES >= 3, "ring too shallow");
  constexpr size_t ring = static_cast<size_t>(STAGE) * STAGES;
  const size_t smem = ring > static_cast<size_t>(GR) * DA_COLS * 4 ? ring : static_cast<size_t>(GR) * DA_COLS * 4;
  const int col_slabs = (K + DA_COLS - 1) / DA_COLS;
  int row_slabs = sm_count() / col_slabs;
  if (row_slabs < 1) row_slabs = 1;
  long long rps = (M + row_slabs - 1) / row_slabs;
  rps = (rps + DA_ROWS - 1) / DA_ROWS * DA_ROWS;        // slabs start on multiples of 32 rows (row pairs never straddle two CTAs)
  row_slabs = static_cast<int>((M + rps - 1) / rps);
  static bool attr_done = false;
  auto kern = lora_da_kernel<G, NT, FIX, ZG, STAGES>;
  if (!attr_done) { NS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done = true; }
  // The column slabs of one row slab are launched as a thread-block CLUSTER: co-scheduled CTAs walk the same rows at the same
  // pace, so the 512-byte pieces of a row are requested together and DRAM sees whole rows (pages) instead of scattered halves.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(col_slabs, row_slabs);
  cfg.blockDim = dim3(DA_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (col_slabs <= 8 && getenv("NS_LORA_CLUSTER") != nullptr) ? col_slabs : 1;   // measured: no gain at 2 slabs, 2x slower at 8
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NS_CUDA(cudaLaunchKernelEx(&cfg, kern, M, K, rps, static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(dt), lddt, dA, ldg,
                             bits, static_cast<__nv_bfloat16*>(dx), lddx, static_cast<const __nv_bfloat16*>(At), ldat,
                             static_cast<const __nv_bfloat16*>(z), ldz));
  return NS_OK;
}

template <int G, int NT>
static int launch_da(long long M, int K, const void* x, long long ldx, const void* dt, long long lddt, float* dA, long long ldg,
                     const uint32_t* bits, void* dx, long long lddx, const void* At, long long ldat, const void* z, long long ldz,
                     cudaStream_t st) {
  if (dx && z) return launch_da_variant<G, NT, true, true>(M, K, x, ldx, dt, lddt, dA, ldg, bits, dx, lddx, At, ldat, z, ldz, st);
  if (dx) return launch_da_variant<G, NT, true, false>(M, K, x, ldx, dt, lddt, dA, ldg, bits, dx, lddx, At, ldat, z, ldz, st);
  return launch_da_variant<G, NT, false, false>(M, K, x, ldx, dt, lddt, dA, ldg, bits, dx, lddx, At, ldat, z, ldz, st);
}

template <typename Kern>
static int launch_down_kernel(Kern kern, long long ctas, int threads, size_t smem, long long M, int K, const void* x, long long ldx,
                              const void* A, long long lda, void* t, long long ldt, float alpha, const uint32_t* bits, cudaStream_t st) {
  NS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<static_cast<unsigned>(ctas), threads, smem, st>>>(M, K, static_cast<const __nv_bfloat16*>(x), ldx,
                                                           static_cast<const __nv_bfloat16*>(A), lda, static_cast<__nv_bfloat16*>(t), ldt,
                                                           alpha, bits);
  NS_LAUNCH_CHECK();
  return NS_OK;
}

template <int G, int NT, int WARPS, int MT>
static int launch_down(long long M, int K, const void* x, long long ldx, const void* A, long long lda, void* t, long long ldt,
                       float alpha, const uint32_t* bits, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(G) * NT * 8 * (K * 2 + 64);
  const long long tiles = (M + 16 * MT - 1) / (16 * MT);
  long long ctas = (tiles + WARPS - 1) / WARPS;
  if (ctas > sm_count()) ctas = sm_count();            // one CTA per SM (registers), equal row range per warp
  if (K % 64 == 0)
    return launch_down_kernel(lora_down_kernel<G, NT, WARPS, MT, 2>, ctas, WARPS * 32, smem, M, K, x, ldx, A, lda, t, ldt, alpha, bits, st);
  return launch_down_kernel(lora_down_kernel<G, NT, WARPS, MT, 1>, ctas, WARPS * 32, smem, M, K, x, ldx, A, lda, t, ldt, alpha, bits, st);
}


}  // namespace ns

using namespace ns;

extern "C" {

// (groups, rank) instantiations: r = 32 (finetune.py:210), 16 (AdaLoRA's 12 padded), 8 (tests)
#define NS_LORA_DISPATCH(FN, ...)                                                        \
  do {                                                                                   \
    if (G == 1 && r == 32) return FN<1, 4 NS_W1>(__VA_ARGS__);                           \
    if (G == 3 && r == 32) return FN<3, 4 NS_W3>(__VA_ARGS__);                           \
    if (G == 1 && r == 16) return FN<1, 2 NS_W1>(__VA_ARGS__);                           \
    if (G == 3 && r == 16) return FN<3, 2 NS_W3>(__VA_ARGS__);                           \
    if (G == 1 && r == 8) return FN<1, 1 NS_W1>(__VA_ARGS__);                            \
    if (G == 3 && r == 8) return FN<3, 1 NS_W1>(__VA_ARGS__);                            \
  } while (0)

int ns_seed_advance(unsigned int* seed, void* stream) {
  NS_CHECK_ARG(seed, "ns_seed_advance: null seed");
  seed_advance_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(seed);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

long long ns_dropout_bits_words(long long rows, int cols) { return rows * ((cols + 31) / 32); }

int ns_dropout_bits(long long rows, int cols, int G, const unsigned int* seed, const unsigned int* salts, float p, unsigned int* bits,
                    void* stream) {
  NS_CHECK_ARG(rows >= 0 && rows < (1LL << 31) && cols > 0 && G >= 1 && G <= kMaxPlanes && seed && salts && bits, "ns_dropout_bits: bad arguments");
  NS_CHECK_ARG(p >= 0.f && p < 1.f, "ns_dropout_bits: p = %f out of [0, 1)", p);
  if (rows == 0) return NS_OK;
  Salts s;
  for (int g = 0; g < kMaxPlanes; ++g) s.s[g] = g < G ? salts[g] : 0u;
  const int words = (cols + 31) / 32;
  long long blocks = (rows * words + 255) / 256;
  const long long cap = (static_cast<long long>(sm_count()) * 8 + G - 1) / G;    // about 8 CTAs per SM in total, grid-stride loops
  if (blocks > cap) blocks = cap;
  dropout_bits_kernel<<<dim3(static_cast<unsigned>(blocks), G), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rows, cols, words, seed, s, drop_thr16(p), bits);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_dropout_apply(int dtype, long long rows, int cols, const void* x, long long ldx, void* y, long long ldy,
                     const unsigned int* bits, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_dropout_apply: bad dtype %d", dtype);
  NS_CHECK_ARG(rows >= 0 && cols > 0 && x && y && bits && ldx >= cols && ldy >= cols, "ns_dropout_apply: bad shape/pointers");
  if (rows == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = rows * cols;
  long long blocks = (n + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (dtype == NS_BF16)
    dropout_apply_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(rows, cols, static_cast<const __nv_bfloat16*>(x), ldx,
                                                                          static_cast<__nv_bfloat16*>(y), ldy, bits);
  else
    dropout_apply_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(rows, cols, static_cast<const float*>(x), ldx, static_cast<float*>(y),
                                                                  ldy, bits);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_lora_dx_fix(int dtype, long long rows, int K, int G, int r, void* dx, long long lddx, const void* dt, long long lddt,
                   const void* At, long long ldat, const unsigned int* bits, const void* z, long long ldz, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_lora_dx_fix: bad dtype %d", dtype);
  NS_CHECK_ARG(rows >= 0 && K > 0 && G >= 1 && G <= 3 && r > 0 && dx && dt && At && bits, "ns_lora_dx_fix: bad arguments");
  NS_CHECK_ARG(lddx >= K && lddt >= G * r && ldat >= G * r && (!z || ldz >= K), "ns_lora_dx_fix: leading dimension too small");
  if (rows == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long pairs = (rows + 1) / 2;
  const unsigned blocks = static_cast<unsigned>((pairs + 7) / 8);
  const size_t smem = 8 * 2 * G * r * sizeof(float);
  if (dtype == NS_BF16)
    lora_dx_fix_kernel<__nv_bfloat16><<<blocks, 256, smem, st>>>(rows, K, G, r, static_cast<__nv_bfloat16*>(dx), lddx,
                                                                 static_cast<const __nv_bfloat16*>(dt), lddt, static_cast<const __nv_bfloat16*>(At),
                                                                 ldat, bits, static_cast<const __nv_bfloat16*>(z), ldz);
  else
    lora_dx_fix_kernel<float><<<blocks, 256, smem, st>>>(rows, K, G, r, static_cast<float*>(dx), lddx, static_cast<const float*>(dt), lddt,
                                                         static_cast<const float*>(At), ldat, bits, static_cast<const float*>(z), ldz);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_lora_down(long long M, int K, int G, int r, const void* x, long long ldx, const void* A, long long lda, void* t, long long ldt,
                 float alpha, const unsigned int* bits, void* stream) {
  NS_CHECK_ARG(M >= 0 && K > 0 && K % 32 == 0 && x && A && t, "ns_lora_down: bad shape/pointers (K must be a multiple of 32)");
  NS_CHECK_ARG(ldx >= K && lda >= K && ldt >= G * r && ldx % 8 == 0 && lda % 8 == 0 && ldt % 2 == 0, "ns_lora_down: bad leading dimensions");
  NS_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(t) & 3) == 0,
               "ns_lora_down: operands must be 16-byte aligned");
  if (M == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (static_cast<size_t>(G) * r * (K * 2 + 64) > 220 * 1024) {
    set_error("ns_lora_down: A (%d x %d) does not fit in shared memory", G * r, K);
    return NS_ERR_UNSUPPORTED;
  }
  count(C_OTHER);
#define NS_W1 , 16, 2
#define NS_W3 , 8, 2
  NS_LORA_DISPATCH(launch_down, M, K, x, ldx, A, lda, t, ldt, alpha, bits, st);
#undef NS_W1
#undef NS_W3
  set_error("ns_lora_down: unsupported (groups, rank) = (%d, %d)", G, r);
  return NS_ERR_UNSUPPORTED;
}

int ns_lora_da(long long M, int K, int G, int r, const void* x, long long ldx, const void* dt, long long lddt, float* dA, long long ldg,
               const unsigned int* bits, void* dx, long long lddx, const void* At, long long ldat, const void* z, long long ldz,
               void* stream) {
  NS_CHECK_ARG(M >= 0 && K > 0 && K % 64 == 0 && x && dt && dA, "ns_lora_da: bad shape/pointers (K must be a multiple of 64)");
  NS_CHECK_ARG(ldx >= K && lddt >= G * r && ldg >= K && ldx % 8 == 0 && lddt % 8 == 0 && ldg % 4 == 0, "ns_lora_da: bad leading dimensions");
  NS_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dt) & 15) == 0 && (reinterpret_cast<uintptr_t>(dA) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(bits) & 3) == 0,
               "ns_lora_da: operands must be 16-byte aligned");
  NS_CHECK_ARG(!dx || (bits && At && lddx >= K && lddx % 2 == 0 && ldat >= G * r && ldat % 2 == 0 && M * lddx < (1LL << 32) &&
                       (reinterpret_cast<uintptr_t>(dx) & 3) == 0 && (reinterpret_cast<uintptr_t>(At) & 3) == 0),
               "ns_lora_da: the dx correction needs bits, A^T and even leading dimensions");
  NS_CHECK_ARG(!z || (dx && ldz >= K && ldz % 2 == 0 && (reinterpret_cast<uintptr_t>(z) & 3) == 0), "ns_lora_da: z goes with dx");
  if (M == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  count(C_OTHER);
#define NS_W1
#define NS_W3
  NS_LORA_DISPATCH(launch_da, M, K, x, ldx, dt, lddt, dA, ldg, bits, dx, lddx, At, ldat, z, ldz, st);
#undef NS_W1
#undef NS_W3
  set_error("ns_lora_da: unsupported (groups, rank) = (%d, %d)", G, r);
  return NS_ERR_UNSUPPORTED;
}

}  // extern "C"
