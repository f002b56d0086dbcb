// LoRA side kernels (finetune.py:194-212): the branch dropout of PEFT's lora.Linear and the rank-r products around it.
//
//   y = base(x) + (alpha/r) * B(A(dropout_p(x)))          finetune.py:210  lora_dropout = 0.05 (0.1 for AdaLoRA, :206-207)
//
// The keep mask is a counter hash, never stored: element (row, col) of module `salt` at step seed `seed` is DROPPED iff
//     w = lowbias32( ((row >> 1) * 0x9E3779B1) ^ (col * 0x85EBCA77) ^ seed ^ salt );   half = (row & 1) ? w >> 16 : w & 0xFFFF
//     half < thr16,   thr16 = round(p * 65536)
// (one 32-bit hash serves the two rows of a row pair: every kernel below holds both rows of a pair in one thread, so the mask
// costs half a hash per element).  The oracle restates exactly this (oracle/whisper_eeg.py::lora_dropout_keep).
// Convention: kernels work with the UNSCALED masked input x (.) keep; the 1/(1-p) lives in alpha' = (alpha/r)/(1-p), which
// scales t = alpha' (x.keep) A^T forward and dt' = alpha' g B backward, so dA = dt'^T (x.keep) and dx += (dt' A).keep.
//
//   ns_lora_down    t[M, G*r] = alpha' * (x . keep_g) A_g^T          HBM-bound: x streams through registers once, A_g in shared
//                   memory, mma.sync m16n8k16 on register fragments (a rank-32 product cannot feed tcgen05's 128-row tiles from
//                   registers; the kernel is bound by the read of x, not by the tensor pipe)
//   ns_lora_da      dA_g[r, K] += dt'_g^T (x . keep_g)               split over row slabs, cp.async ring, mask applied on the
//                   ldmatrix fragments, fp32 vector reductions into the flat gradient buffer
//   ns_lora_dx_fix  dx[m,k] -= dropped_g(m,k) * (dt'_g[m,:] . A_g[:,k]) [* gelu'(z[m,k])]     sparse correction after the input-
//                   gradient GEMM, which carries the LoRA product as a K-segment as if nothing had been dropped
//   ns_dropout_apply  y = x . keep          (fp32 parity mode / reference path of the tests: materialises the masked input)
//   ns_seed_advance   seed <- lowbias32(seed + 0x9E3779B9)            (inside the captured training step: a new mask per replay)
#include "ns_common.cuh"

namespace ns {

__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
constexpr uint32_t kRowMul = 0x9E3779B1u, kColMul = 0x85EBCA77u;
__device__ __forceinline__ uint32_t pair_key(uint32_t row_pair, uint32_t module_seed) { return (row_pair * kRowMul) ^ module_seed; }
__device__ __forceinline__ uint32_t drop_word(uint32_t pkey, uint32_t col) { return lowbias32(pkey ^ (col * kColMul)); }
// AND-mask for a packed pair (lo element uses half `lo16`, hi element uses half `hi16`): 0 where dropped
__device__ __forceinline__ uint32_t keep_bits(uint32_t lo16, uint32_t hi16, uint32_t thr) {
  return (lo16 < thr ? 0u : 0x0000FFFFu) | (hi16 < thr ? 0u : 0xFFFF0000u);
}

struct Salts { uint32_t s[3]; };

// ------------------------------------------------------------------------------------------------ seed / materialised mask
__global__ void seed_advance_kernel(uint32_t* seed) { *seed = lowbias32(*seed + 0x9E3779B9u); }

template <typename T>
__global__ void __launch_bounds__(256) dropout_apply_kernel(long long rows, int cols, const T* __restrict__ x, long long ldx,
                                                            T* __restrict__ y, long long ldy, const uint32_t* __restrict__ seed,
                                                            uint32_t salt, uint32_t thr) {
  const long long pairs = (rows + 1) >> 1;
  const uint32_t ms = *seed ^ salt;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < pairs * cols;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long rp = i / cols;
    const int c = static_cast<int>(i - rp * cols);
    const uint32_t w = drop_word(pair_key(static_cast<uint32_t>(rp), ms), static_cast<uint32_t>(c));
    const long long r0 = rp * 2;
    y[r0 * ldy + c] = (w & 0xFFFFu) < thr ? from_f<T>(0.f) : x[r0 * ldx + c];
    if (r0 + 1 < rows) y[(r0 + 1) * ldy + c] = (w >> 16) < thr ? from_f<T>(0.f) : x[(r0 + 1) * ldx + c];
  }
}

// ------------------------------------------------------------------------------------------------ dx correction
template <typename T> __device__ __forceinline__ float dgelu_of(float z);
template <> __device__ __forceinline__ float dgelu_of<float>(float z) { return dgelu_erf(z); }
template <> __device__ __forceinline__ float dgelu_of<__nv_bfloat16>(float z) { return dgelu_fast(z); }

// One warp per row pair; lane = column (stride 32).  dt' rows of the pair sit in shared memory as fp32; At is A transposed,
// (K, ldat >= G*r): the r coefficients a dropped element needs are contiguous.
template <typename T>
__global__ void __launch_bounds__(256) lora_dx_fix_kernel(long long rows, int K, int G, int r, T* __restrict__ dx, long long lddx,
                                                          const T* __restrict__ dt, long long lddt, const T* __restrict__ At,
                                                          long long ldat, const uint32_t* __restrict__ seed, Salts salts,
                                                          uint32_t thr, const T* __restrict__ z, long long ldz) {
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
  const uint32_t sd0 = *seed;
  uint32_t pk[3];
  for (int g = 0; g < G; ++g) pk[g] = pair_key(static_cast<uint32_t>(rp), sd0 ^ salts.s[g]);
  for (int c = lane; c < K; c += 32) {
    float fix0 = 0.f, fix1 = 0.f;
    bool any0 = false, any1 = false;
    for (int g = 0; g < G; ++g) {
      const uint32_t w = drop_word(pk[g], static_cast<uint32_t>(c));
      const bool d0 = (w & 0xFFFFu) < thr, d1 = two && (w >> 16) < thr;
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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ------------------------------------------------------------------------------------------------ t = alpha' (x . keep) A^T
// Warp = 32 rows (two m16 tiles; fragment rows (g, g+8) <-> actual rows (2g, 2g+1) of the tile, so a thread owns both rows
// of a row pair).  Per 32-column chunk a thread loads 16 B of each of its 4 rows; the 8 bf16 are consumed as the k-slots
// {2t,2t+1,2t+8,2t+9} of two MMAs (k order inside a dot product is free as long as the A_g fragment uses the same order:
// the A_g fragment is the same 16 B of A_g's row).  A_g rows live in shared memory, pitch K*2+64 B (conflict-free LDS.128).
template <int G, int NT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) lora_down_kernel(long long M, int K, const __nv_bfloat16* __restrict__ x, long long ldx,
                                                               const __nv_bfloat16* __restrict__ A, long long lda,
                                                               __nv_bfloat16* __restrict__ t, long long ldt, float alpha,
                                                               const uint32_t* __restrict__ seed, Salts salts, uint32_t thr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int R = NT * 8;
  const int pitch = K * 2 + 64;                                         // bytes
  for (int i = threadIdx.x; i < G * R * (K / 8); i += WARPS * 32) {     // A_g rows -> shared
    const int row = i / (K / 8), v = i - row * (K / 8);
    *reinterpret_cast<uint4*>(smem_raw + row * pitch + v * 16) = __ldg(reinterpret_cast<const uint4*>(A + row * lda) + v);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const uint32_t sd0 = thr ? *seed : 0u;
  const long long tiles = (M + 31) / 32;
  for (long long tile = static_cast<long long>(blockIdx.x) * WARPS + warp; tile < tiles; tile += static_cast<long long>(gridDim.x) * WARPS) {
    const long long R0 = tile * 32;
    float acc[2][G][NT][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < G; ++b)
#pragma unroll
        for (int c = 0; c < NT; ++c)
#pragma unroll
          for (int d = 0; d < 4; ++d) acc[a][b][c][d] = 0.f;
    const __nv_bfloat16* xr[2][2];
    bool ok[2][2];
    uint32_t pk[2][G];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const long long re = R0 + mt * 16 + 2 * g;
#pragma unroll
      for (int o = 0; o < 2; ++o) { ok[mt][o] = re + o < M; xr[mt][o] = x + (ok[mt][o] ? re + o : 0) * ldx + tq * 8; }
#pragma unroll
      for (int gi = 0; gi < G; ++gi) pk[mt][gi] = pair_key(static_cast<uint32_t>(re >> 1), sd0 ^ salts.s[gi]);
    }
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    uint4 cur[2][2], nx1[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        cur[mt][o] = ok[mt][o] ? __ldg(reinterpret_cast<const uint4*>(xr[mt][o])) : zero4;
        nx1[mt][o] = (ok[mt][o] && K > 32) ? __ldg(reinterpret_cast<const uint4*>(xr[mt][o] + 32)) : zero4;
      }
    for (int c0 = 0; c0 < K; c0 += 32) {
      uint4 nx2[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int o = 0; o < 2; ++o)
          nx2[mt][o] = (ok[mt][o] && c0 + 64 < K) ? __ldg(reinterpret_cast<const uint4*>(xr[mt][o] + c0 + 64)) : zero4;
      const uint32_t col0 = static_cast<uint32_t>(c0 + tq * 8);
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        uint32_t xe[2][4], xo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t e[4] = {cur[mt][0].x, cur[mt][0].y, cur[mt][0].z, cur[mt][0].w};
          const uint32_t o[4] = {cur[mt][1].x, cur[mt][1].y, cur[mt][1].z, cur[mt][1].w};
          if (thr) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t w0 = drop_word(pk[mt][gi], col0 + 2 * i), w1 = drop_word(pk[mt][gi], col0 + 2 * i + 1);
              xe[mt][i] = e[i] & keep_bits(w0 & 0xFFFFu, w1 & 0xFFFFu, thr);
              xo[mt][i] = o[i] & keep_bits(w0 >> 16, w1 >> 16, thr);
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
          for (int mt = 0; mt < 2; ++mt) {
            mma_bf16_16816(acc[mt][gi][nt], xe[mt][0], xo[mt][0], xe[mt][1], xo[mt][1], w.x, w.y);
            mma_bf16_16816(acc[mt][gi][nt], xe[mt][2], xo[mt][2], xe[mt][3], xo[mt][3], w.z, w.w);
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int o = 0; o < 2; ++o) { cur[mt][o] = nx1[mt][o]; nx1[mt][o] = nx2[mt][o]; }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const long long re = R0 + mt * 16 + 2 * g;
#pragma unroll
      for (int gi = 0; gi < G; ++gi)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int col = gi * R + nt * 8 + tq * 2;
          if (re < M) *reinterpret_cast<uint32_t*>(t + re * ldt + col) = pack_bf16x2(alpha * acc[mt][gi][nt][0], alpha * acc[mt][gi][nt][1]);
          if (re + 1 < M) *reinterpret_cast<uint32_t*>(t + (re + 1) * ldt + col) = pack_bf16x2(alpha * acc[mt][gi][nt][2], alpha * acc[mt][gi][nt][3]);
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------ dA_g += dt'_g^T (x . keep_g)
// CTA = (256-column slab of K, row slab), 16 warps x 16 columns.  D'[col, r] = sum_m x[m, col] dt'[m, r]: both operands are
// read "transposed" from their row-major shared tiles with ldmatrix.trans; a fragment register then holds the two rows of a row
// pair at one column, which is exactly what one mask word covers.  4-stage cp.async ring of 32-row tiles.
constexpr int DA_COLS = 256, DA_ROWS = 32, DA_STAGES = 4, DA_WARPS = 16;
template <int G, int NT>
__global__ void __launch_bounds__(DA_WARPS * 32) lora_da_kernel(long long M, int K, long long rows_per_slab, const __nv_bfloat16* __restrict__ x,
                                                               long long ldx, const __nv_bfloat16* __restrict__ dt, long long lddt,
                                                               float* __restrict__ dA, long long ldg, const uint32_t* __restrict__ seed,
                                                               Salts salts, uint32_t thr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int R = NT * 8, GR = G * R;
  constexpr int XP = (DA_COLS + 8) * 2;                 // x tile pitch (bytes): +16 B -> conflict-free ldmatrix
  constexpr int DP = (GR + 8) * 2;                      // dt' tile pitch
  constexpr int STAGE = DA_ROWS * XP + DA_ROWS * DP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int k0 = blockIdx.x * DA_COLS;
  const long long m_begin = static_cast<long long>(blockIdx.y) * rows_per_slab;
  const long long m_end = m_begin + rows_per_slab < M ? m_begin + rows_per_slab : M;
  const int n_tiles = m_end > m_begin ? static_cast<int>((m_end - m_begin + DA_ROWS - 1) / DA_ROWS) : 0;
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sd0 = thr ? *seed : 0u;

  auto issue = [&](int tile) {
    if (tile < n_tiles) {
      const uint32_t st = sbase + (tile % DA_STAGES) * STAGE;
      const long long mrow = m_begin + static_cast<long long>(tile) * DA_ROWS;
      for (int i = threadIdx.x; i < DA_ROWS * (DA_COLS / 8); i += DA_WARPS * 32) {
        const int rr = i / (DA_COLS / 8), v = i - rr * (DA_COLS / 8);
        const bool ok = mrow + rr < m_end && k0 + v * 8 < K;
        cp_async16(st + rr * XP + v * 16, x + (ok ? (mrow + rr) * ldx + k0 + v * 8 : 0), ok ? 16 : 0);
      }
      for (int i = threadIdx.x; i < DA_ROWS * (GR / 8); i += DA_WARPS * 32) {
        const int rr = i / (GR / 8), v = i - rr * (GR / 8);
        const bool ok = mrow + rr < m_end;
        cp_async16(st + DA_ROWS * XP + rr * DP + v * 16, dt + (ok ? (mrow + rr) * lddt + v * 8 : 0), ok ? 16 : 0);
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

  for (int s = 0; s < DA_STAGES - 1; ++s) issue(s);
  const int wcol = warp * 16;                           // this warp's 16 columns inside the slab
  const int q = lane >> 3, rr8 = lane & 7;
  for (int tile = 0; tile < n_tiles; ++tile) {
    cp_async_wait<DA_STAGES - 2>();
    __syncthreads();                                    // tile `tile` has landed for everyone; stage (tile-1)%S is free again
    issue(tile + DA_STAGES - 1);
    const uint32_t st = sbase + (tile % DA_STAGES) * STAGE;
    const long long mrow = m_begin + static_cast<long long>(tile) * DA_ROWS;
#pragma unroll
    for (int ks = 0; ks < DA_ROWS / 16; ++ks) {
      uint32_t a[4];
      ldsm_x4_trans(a, st + (ks * 16 + (q >> 1) * 8 + rr8) * XP + (wcol + (q & 1) * 8) * 2);
      const uint32_t rp = static_cast<uint32_t>((mrow + ks * 16) >> 1) + tq;     // row pair of (m = 2t, 2t+1); +4 for the m+8 half
      const uint32_t col = static_cast<uint32_t>(k0 + wcol + g);
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        uint32_t am[4];
        if (thr) {
          const uint32_t ms = sd0 ^ salts.s[gi];
          const uint32_t w0 = drop_word(pair_key(rp, ms), col), w1 = drop_word(pair_key(rp, ms), col + 8);
          const uint32_t w2 = drop_word(pair_key(rp + 4, ms), col), w3 = drop_word(pair_key(rp + 4, ms), col + 8);
          am[0] = a[0] & keep_bits(w0 & 0xFFFFu, w0 >> 16, thr);
          am[1] = a[1] & keep_bits(w1 & 0xFFFFu, w1 >> 16, thr);
          am[2] = a[2] & keep_bits(w2 & 0xFFFFu, w2 >> 16, thr);
          am[3] = a[3] & keep_bits(w3 & 0xFFFFu, w3 >> 16, thr);
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

template <int G, int NT>
static int launch_da(long long M, int K, const void* x, long long ldx, const void* dt, long long lddt, float* dA, long long ldg,
                     const uint32_t* seed, Salts salts, uint32_t thr, cudaStream_t st) {
  constexpr int R = NT * 8, GR = G * R;
  constexpr int STAGE = DA_ROWS * (DA_COLS + 8) * 2 + DA_ROWS * (GR + 8) * 2;
  const size_t smem = STAGE * DA_STAGES > GR * DA_COLS * 4 ? STAGE * DA_STAGES : GR * DA_COLS * 4;
  const int col_slabs = (K + DA_COLS - 1) / DA_COLS;
  int row_slabs = sm_count() / col_slabs;
  if (row_slabs < 1) row_slabs = 1;
  long long rps = (M + row_slabs - 1) / row_slabs;
  rps = (rps + DA_ROWS - 1) / DA_ROWS * DA_ROWS;        // slabs start on even rows (row pairs never straddle two CTAs)
  row_slabs = static_cast<int>((M + rps - 1) / rps);
  static bool attr_done = false;
  auto kern = lora_da_kernel<G, NT>;
  if (!attr_done) { NS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done = true; }
  kern<<<dim3(col_slabs, row_slabs), DA_WARPS * 32, smem, st>>>(M, K, rps, static_cast<const __nv_bfloat16*>(x), ldx,
                                                               static_cast<const __nv_bfloat16*>(dt), lddt, dA, ldg, seed, salts, thr);
  NS_LAUNCH_CHECK();
  return NS_OK;
}

template <int G, int NT, int WARPS>
static int launch_down(long long M, int K, const void* x, long long ldx, const void* A, long long lda, void* t, long long ldt,
                       float alpha, const uint32_t* seed, Salts salts, uint32_t thr, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(G) * NT * 8 * (K * 2 + 64);
  static size_t attr_smem = 0;
  auto kern = lora_down_kernel<G, NT, WARPS>;
  if (smem > attr_smem) { NS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; }
  const long long tiles = (M + 31) / 32;
  long long ctas = (tiles + WARPS - 1) / WARPS;
  if (ctas > sm_count()) ctas = sm_count();            // one CTA per SM (registers: WARPS*32 threads x up to 128/255), warps stride the tiles
  kern<<<static_cast<unsigned>(ctas), WARPS * 32, smem, st>>>(M, K, static_cast<const __nv_bfloat16*>(x), ldx,
                                                            static_cast<const __nv_bfloat16*>(A), lda, static_cast<__nv_bfloat16*>(t),
                                                            ldt, alpha, seed, salts, thr);
  NS_LAUNCH_CHECK();
  return NS_OK;
}

static uint32_t thr16(float p) {
  const double v = static_cast<double>(p) * 65536.0 + 0.5;
  return v <= 0 ? 0u : (v >= 65535.0 ? 65535u : static_cast<uint32_t>(v));
}

}  // namespace ns

using namespace ns;

extern "C" {

int ns_seed_advance(unsigned int* seed, void* stream) {
  NS_CHECK_ARG(seed, "ns_seed_advance: null seed");
  seed_advance_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(seed);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_dropout_apply(int dtype, long long rows, int cols, const void* x, long long ldx, void* y, long long ldy,
                     const unsigned int* seed, unsigned int salt, float p, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_dropout_apply: bad dtype %d", dtype);
  NS_CHECK_ARG(rows >= 0 && cols > 0 && x && y && seed && ldx >= cols && ldy >= cols, "ns_dropout_apply: bad shape/pointers");
  NS_CHECK_ARG(p >= 0.f && p < 1.f, "ns_dropout_apply: p = %f out of [0, 1)", p);
  if (rows == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = ((rows + 1) / 2) * cols;
  long long blocks = (n + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (dtype == NS_BF16)
    dropout_apply_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(rows, cols, static_cast<const __nv_bfloat16*>(x), ldx,
                                                                          static_cast<__nv_bfloat16*>(y), ldy, seed, salt, thr16(p));
  else
    dropout_apply_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(rows, cols, static_cast<const float*>(x), ldx, static_cast<float*>(y),
                                                                  ldy, seed, salt, thr16(p));
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_lora_dx_fix(int dtype, long long rows, int K, int G, int r, void* dx, long long lddx, const void* dt, long long lddt,
                   const void* At, long long ldat, const unsigned int* seed, const unsigned int* salts, float p, const void* z,
                   long long ldz, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype), "ns_lora_dx_fix: bad dtype %d", dtype);
  NS_CHECK_ARG(rows >= 0 && K > 0 && G >= 1 && G <= 3 && r > 0 && dx && dt && At && seed && salts, "ns_lora_dx_fix: bad arguments");
  NS_CHECK_ARG(lddx >= K && lddt >= G * r && ldat >= G * r && (!z || ldz >= K), "ns_lora_dx_fix: leading dimension too small");
  NS_CHECK_ARG(p >= 0.f && p < 1.f, "ns_lora_dx_fix: p = %f out of [0, 1)", p);
  const uint32_t thr = thr16(p);
  if (rows == 0 || thr == 0) return NS_OK;
  Salts s{{salts[0], G > 1 ? salts[1] : 0u, G > 2 ? salts[2] : 0u}};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long pairs = (rows + 1) / 2;
  const unsigned blocks = static_cast<unsigned>((pairs + 7) / 8);
  const size_t smem = 8 * 2 * G * r * sizeof(float);
  if (dtype == NS_BF16)
    lora_dx_fix_kernel<__nv_bfloat16><<<blocks, 256, smem, st>>>(rows, K, G, r, static_cast<__nv_bfloat16*>(dx), lddx,
                                                                 static_cast<const __nv_bfloat16*>(dt), lddt, static_cast<const __nv_bfloat16*>(At),
                                                                 ldat, seed, s, thr, static_cast<const __nv_bfloat16*>(z), ldz);
  else
    lora_dx_fix_kernel<float><<<blocks, 256, smem, st>>>(rows, K, G, r, static_cast<float*>(dx), lddx, static_cast<const float*>(dt), lddt,
                                                         static_cast<const float*>(At), ldat, seed, s, thr, static_cast<const float*>(z), ldz);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

#define NS_LORA_DISPATCH(FN, ...)                                                        \
  do {                                                                                   \
    if (G == 1 && r == 32) return FN<1, 4 NS_W1>(__VA_ARGS__);                           \
    if (G == 3 && r == 32) return FN<3, 4 NS_W3>(__VA_ARGS__);                           \
    if (G == 1 && r == 16) return FN<1, 2 NS_W1>(__VA_ARGS__);                           \
    if (G == 3 && r == 16) return FN<3, 2 NS_W3>(__VA_ARGS__);                           \
    if (G == 1 && r == 8) return FN<1, 1 NS_W1>(__VA_ARGS__);                            \
    if (G == 3 && r == 8) return FN<3, 1 NS_W1>(__VA_ARGS__);                            \
  } while (0)

int ns_lora_down(long long M, int K, int G, int r, const void* x, long long ldx, const void* A, long long lda, void* t, long long ldt,
                 float alpha, const unsigned int* seed, const unsigned int* salts, float p, void* stream) {
  NS_CHECK_ARG(M >= 0 && K > 0 && K % 32 == 0 && x && A && t, "ns_lora_down: bad shape/pointers (K must be a multiple of 32)");
  NS_CHECK_ARG(ldx >= K && lda >= K && ldt >= G * r && ldx % 8 == 0 && lda % 8 == 0 && ldt % 2 == 0, "ns_lora_down: bad leading dimensions");
  NS_CHECK_ARG(p >= 0.f && p < 1.f && (p == 0.f || (seed && salts)), "ns_lora_down: dropout needs seed and salts");
  NS_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(t) & 3) == 0,
               "ns_lora_down: operands must be 16-byte aligned");
  if (M == 0) return NS_OK;
  const uint32_t thr = thr16(p);
  Salts s{{salts ? salts[0] : 0u, (salts && G > 1) ? salts[1] : 0u, (salts && G > 2) ? salts[2] : 0u}};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (static_cast<size_t>(G) * r * (K * 2 + 64) > 220 * 1024) {
    set_error("ns_lora_down: A (%d x %d) does not fit in shared memory", G * r, K);
    return NS_ERR_UNSUPPORTED;
  }
  count(C_OTHER);
#define NS_W1 , 16
#define NS_W3 , 8
  NS_LORA_DISPATCH(launch_down, M, K, x, ldx, A, lda, t, ldt, alpha, seed, s, thr, st);
#undef NS_W1
#undef NS_W3
  set_error("ns_lora_down: unsupported (groups, rank) = (%d, %d)", G, r);
  return NS_ERR_UNSUPPORTED;
}

int ns_lora_da(long long M, int K, int G, int r, const void* x, long long ldx, const void* dt, long long lddt, float* dA, long long ldg,
               const unsigned int* seed, const unsigned int* salts, float p, void* stream) {
  NS_CHECK_ARG(M >= 0 && K > 0 && K % 8 == 0 && x && dt && dA, "ns_lora_da: bad shape/pointers (K must be a multiple of 8)");
  NS_CHECK_ARG(ldx >= K && lddt >= G * r && ldg >= K && ldx % 8 == 0 && lddt % 8 == 0 && ldg % 4 == 0, "ns_lora_da: bad leading dimensions");
  NS_CHECK_ARG(p >= 0.f && p < 1.f && (p == 0.f || (seed && salts)), "ns_lora_da: dropout needs seed and salts");
  NS_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dt) & 15) == 0 && (reinterpret_cast<uintptr_t>(dA) & 15) == 0,
               "ns_lora_da: operands must be 16-byte aligned");
  if (M == 0) return NS_OK;
  const uint32_t thr = thr16(p);
  Salts s{{salts ? salts[0] : 0u, (salts && G > 1) ? salts[1] : 0u, (salts && G > 2) ? salts[2] : 0u}};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  count(C_OTHER);
#define NS_W1
#define NS_W3
  NS_LORA_DISPATCH(launch_da, M, K, x, ldx, dt, lddt, dA, ldg, seed, s, thr, st);
#undef NS_W1
#undef NS_W3
  set_error("ns_lora_da: unsupported (groups, rank) = (%d, %d)", G, r);
  return NS_ERR_UNSUPPORTED;
}

}  // extern "C"
