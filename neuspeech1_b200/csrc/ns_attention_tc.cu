// Flash-style attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), head_dim 64, bf16, non-causal.
//
// forward, one CTA per (batch, head, 128-query tile), two CTAs resident per SM so one CTA's softmax overlaps the other's MMAs:
//   warp 0      TMA producer : Q tile once, then K / V tiles of 128 keys through a 2-stage ring
//   warp 1      MMA issuer   : S = Q K^T (SS form, 128x128 fp32 in TMEM), then O += P V (TS form: P read from TMEM, V MN-major)
//   warps 2..5  softmax      : one query row per thread (TMEM lane = row): two passes over S with tcgen05.ld -- row max, then
//                              p = exp2((s - m) * log2e), row sum, bf16 P written back to TMEM with tcgen05.st.
//                              O is rescaled lazily (only when the running max grew by > 8 in log2 units), so most tiles skip
//                              the TMEM round trip of the accumulator.
// TMEM columns: [0,128) S, [128,192) P (bf16 pairs), [192,256) O.   q is pre-scaled by the caller (HF modeling_whisper.py:310).
#include "ns_common.cuh"
#include "ns_sm100.cuh"

#include <stdlib.h>

namespace ns {
using namespace sm100;

struct AttnMaps {
  CUtensorMap q, k, v;
};
struct AttnFwdProg {
  int B, H, Lq, Lk;
  long long o_bs, o_rs;
  __nv_bfloat16* o;
  float* lse;
  long long* trace;     // developer aid (ns_debug_attn_trace): CTA (0,0,0) of the ping-pong kernel records (tag, clock64)
  int take_turns;
};
long long* get_attn_trace();

constexpr int kAtThreads = 192;
constexpr int kTile = 128 * 64 * 2;                          // one [128][64] bf16 tile, 128B-swizzled
constexpr int kAtSmem = kTile * 5 + 1024 + 256;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA pipe for a fraction of the softmax (the MUFU pipe does 16 ex2 per clock and SM and was 65 % busy: the
// forward's first limit at head_dim 64).  Round-to-nearest split x = n + f by the 1.5 * 2^23 magic add, 2^f on [-0.5, 0.5] by a
// degree-3 polynomial (Lawson-weighted fit, max relative error 7.5e-5: fifty times below a bf16 ulp of P), 2^n by adding n to
// the exponent field.  Packed fp32 pairs: 7 FMA-pipe instructions + 2 clamps + 2 integer adds per pair.
#ifndef NS_ATTN_POLY_EVERY
#define NS_ATTN_POLY_EVERY 4        // every 4th pair of the 32 pairs of a half tile takes the polynomial (0 = never)
#endif
__device__ __forceinline__ void poly_exp2_pair(uint64_t x, float& p0, float& p1) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  x = f2_pack(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));                      // 2^-126: nothing below it survives the bf16 pack
  const uint64_t t = f2_add(x, f2_splat(12582912.0f));                       // low mantissa bits of t = round(x)
  const uint64_t f = f2_add(x, f2_fma(t, f2_splat(-1.0f), f2_splat(12582912.0f)));   // x - round(x)
  const uint64_t p = f2_fma(f, f2_fma(f, f2_fma(f, f2_splat(0.05517146f), f2_splat(0.24261086f)), f2_splat(0.69326097f)), f2_splat(0.9999281f));
  float t0, t1, q0, q1;
  f2_unpack(t, t0, t1);
  f2_unpack(p, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

__global__ void __launch_bounds__(kAtThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnFwdProg p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  auto sK = [&](int s) { return smem_base + kTile * (1 + s); };
  auto sV = [&](int s) { return smem_base + kTile * (3 + s); };
  const uint32_t bar = smem_base + kTile * 5;
  const uint32_t q_full = bar;
  auto k_full = [&](int s) { return bar + 8u * (1 + s); };
  auto v_full = [&](int s) { return bar + 8u * (3 + s); };
  auto k_empty = [&](int s) { return bar + 8u * (5 + s); };
  auto v_empty = [&](int s) { return bar + 8u * (7 + s); };
  const uint32_t s_full = bar + 8u * 9, p_full = bar + 8u * 10, o_full = bar + 8u * 11;
  const uint32_t tmem_slot = bar + 8u * 12;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Lk + 127) / 128;

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(k_full(s), 1); mbar_init(v_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_empty(s), 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tP = tmem + 128, tO = tmem + 192;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, kTile);
      tma_load_3d(&maps.q, q_full, sQ, h * 64, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1u;
        mbar_wait(k_empty(s), ph ^ 1u);
        mbar_expect_tx(k_full(s), kTile);
        tma_load_3d(&maps.k, k_full(s), sK(s), h * 64, j * 128, b);
        mbar_wait(v_empty(s), ph ^ 1u);
        mbar_expect_tx(v_full(s), kTile);
        tma_load_3d(&maps.v, v_full(s), sV(s), h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idescS = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idescO = umma_idesc_bf16(128, 64, 0, 1);          // B = V is MN-major (keys x head_dim rows)
    mbar_wait(q_full, 0);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1u;
      mbar_wait(k_full(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qd = umma_smem_desc(sQ, 16, 1024), kd = umma_smem_desc(sK(s), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tS, qd + 2u * k, kd + 2u * k, idescS, k > 0);
        umma_commit(k_empty(s));
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      mbar_wait(v_full(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t vd = umma_smem_desc(sV(s), 8192, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_f16_ts(tO, tP + 8u * k, vd + 128u * k, idescO, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(v_empty(s));
        if (j == n_kv - 1) umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int nvalid = min(128, p.Lk - j * 128);
      const bool full = nvalid == 128;
      uint32_t v[32];
      if (j == 0) {                                   // first tile: the row max has to be known before the exponentials
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(tS + lane_addr + 32u * c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (full || c * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        m_used = mx;
      }
      // One pass: p = exp2((s - m_used) * log2e) with the max of the PREVIOUS tiles, tracking this tile's max on the side.
      // Only if some row's max grew by more than 8 (log2 units) the accumulator is rescaled and the pass is redone.
      float l_tile = 0.f, mx = -INFINITY;
      for (int attempt = 0; attempt < 2; ++attempt) {
        const float mb = m_used * kLog2e;
        l_tile = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(tS + lane_addr + 32u * c, v);
          tmem_ld_wait();
          uint32_t pk[16];
          if (full) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
              mx = fmaxf(mx, fmaxf(s0, s1));
              const float p0 = fast_exp2(fmaf(s0, kLog2e, -mb)), p1 = fast_exp2(fmaf(s1, kLog2e, -mb));
              l_tile += p0 + p1;
              pk[i] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const bool ok0 = c * 32 + 2 * i < nvalid, ok1 = c * 32 + 2 * i + 1 < nvalid;
              const float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
              if (ok0) mx = fmaxf(mx, s0);
              if (ok1) mx = fmaxf(mx, s1);
              const float p0 = ok0 ? fast_exp2(fmaf(s0, kLog2e, -mb)) : 0.f;
              const float p1 = ok1 ? fast_exp2(fmaf(s1, kLog2e, -mb)) : 0.f;
              l_tile += p0 + p1;
              pk[i] = pack_bf16x2(p0, p1);
            }
          }
          tmem_st16(tP + lane_addr + 16u * c, pk);
        }
        const bool need = mx > m_used + 5.545177f;     // 8 / log2(e)
        if (attempt == 1 || !__any_sync(0xffffffffu, need)) break;
        const float m_new = fmaxf(m_used, mx);
        const float sc = fast_exp2((m_used - m_new) * kLog2e);
        tmem_st_wait();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          tmem_ld32(tO + lane_addr + 32u * c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * sc);
          tmem_st32(tO + lane_addr + 32u * c, v);
        }
        l *= sc;
        m_used = m_new;
      }
      l += l_tile;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int qi = q0 + row;
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = p.o + b * p.o_bs + static_cast<long long>(qi) * p.o_rs + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tO + lane_addr + 32u * c, v);
      tmem_ld_wait();
      if (qi < p.Lq) {
        uint4* dst = reinterpret_cast<uint4*>(orow + 32 * c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
          dst[i] = u;
        }
      }
    }
    if (p.lse && qi < p.Lq) p.lse[(static_cast<long long>(b) * p.H + h) * p.Lq + qi] = m_used + logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Forward with a DOUBLE-BUFFERED score tile (the default).  Same CTA shape as attn_fwd_tc_kernel (one 128-query tile, two CTAs
// per SM, 256 TMEM columns) but the key axis advances in HALF tiles of 64 keys:
//   TMEM columns: [0,64) S buffer 0, [64,128) S buffer 1, [128,192) O.  bf16 P overwrites the first 32 columns of its own S
//   buffer (every score of the half tile is in registers by then).
//   The MMA thread keeps S two half tiles ahead: S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ...  so the softmax warps find their
//   next scores ready when they finish a half tile -- in the single-buffer kernel they sat out PV(j) + S(j+1) + two barrier
//   round trips per tile and the MUFU pipe was busy 52 % of the time.
//   Softmax per half tile: both 32-column TMEM loads behind one wait, exact row max, then the exponentials (packed FFMA2 for the
//   argument, FADD2 for the row sum).  The running max moves only when a row's new max exceeds it by > 8 (log2 units, P stays
//   below 2^8); O is then rescaled after PV(h-1) has finished (pv_done), before P(h) is published.
constexpr int kDbThreads = 192;
__global__ void __launch_bounds__(kDbThreads, 2)
attn_fwd_db_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnFwdProg p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  auto sK = [&](int s) { return smem_base + kTile * (1 + s); };
  auto sV = [&](int s) { return smem_base + kTile * (3 + s); };
  const uint32_t bar = smem_base + kTile * 5;
  const uint32_t q_full = bar;
  auto k_full = [&](int s) { return bar + 8u * (1 + s); };
  auto v_full = [&](int s) { return bar + 8u * (3 + s); };
  auto k_empty = [&](int s) { return bar + 8u * (5 + s); };
  auto v_empty = [&](int s) { return bar + 8u * (7 + s); };
  auto s_full = [&](int b_) { return bar + 8u * (9 + b_); };
  auto p_full = [&](int b_) { return bar + 8u * (11 + b_); };
  auto pv_done = [&](int b_) { return bar + 8u * (13 + b_); };
  const uint32_t o_full = bar + 8u * 15;
  const uint32_t tmem_slot = bar + 8u * 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Lk + 127) / 128;
  const int nh = (p.Lk + 63) / 64;               // half tiles of 64 keys

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1); mbar_init(v_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_empty(s), 1);
      mbar_init(s_full(s), 1); mbar_init(p_full(s), 4); mbar_init(pv_done(s), 1);
    }
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  auto tS = [&](int b_) { return tmem + 64u * static_cast<uint32_t>(b_); };
  const uint32_t tO = tmem + 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, kTile);
      tma_load_3d(&maps.q, q_full, sQ, h * 64, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1u;
        mbar_wait(k_empty(s), ph ^ 1u);
        mbar_expect_tx(k_full(s), kTile);
        tma_load_3d(&maps.k, k_full(s), sK(s), h * 64, j * 128, b);
        mbar_wait(v_empty(s), ph ^ 1u);
        mbar_expect_tx(v_full(s), kTile);
        tma_load_3d(&maps.v, v_full(s), sV(s), h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idescS = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idescO = umma_idesc_bf16(128, 64, 0, 1);          // B = V is MN-major (keys x head_dim rows)
    mbar_wait(q_full, 0);
    // S(hf) = Q K_half^T into score buffer hf & 1 (keys 64 * half .. of key tile hf >> 1: 64 rows of 128 B = +8 KB)
    auto issue_qk = [&](int hf) {
      const int j = hf >> 1, s = j & 1, half = hf & 1;
      if (half == 0) {
        mbar_wait(k_full(s), (j >> 1) & 1u);
        tc_fence_after();
      }
      if (elect_one()) {
        const uint64_t qd = umma_smem_desc(sQ, 16, 1024), kd = umma_smem_desc(sK(s) + 8192u * half, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tS(hf & 1), qd + 2u * k, kd + 2u * k, idescS, k > 0);
        if (half == 1 || hf == nh - 1) umma_commit(k_empty(s));
        umma_commit(s_full(hf & 1));
      }
      __syncwarp();
    };
    issue_qk(0);
    if (nh > 1) issue_qk(1);
    for (int hf = 0; hf < nh; ++hf) {
      const int j = hf >> 1, s = j & 1, half = hf & 1, buf = hf & 1;
      mbar_wait(p_full(buf), (hf >> 1) & 1u);
      tc_fence_after();
      if (half == 0) {
        mbar_wait(v_full(s), (j >> 1) & 1u);
        tc_fence_after();
      }
      if (elect_one()) {
        const uint64_t vd = umma_smem_desc(sV(s), 8192, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)       // 16 keys per MMA: 8 TMEM columns of bf16 pairs, 2 KB of V rows
          umma_f16_ts(tO, tS(buf) + 8u * k, vd + 128u * (4 * half + k), idescO, (hf > 0 || k > 0) ? 1u : 0u);
        if (half == 1 || hf == nh - 1) umma_commit(v_empty(s));
        umma_commit(pv_done(buf));
        if (hf == nh - 1) umma_commit(o_full);
      }
      __syncwarp();
      if (hf + 2 < nh) issue_qk(hf + 2);    // in order behind PV(hf): it may overwrite the buffer PV(hf) reads P from
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float m_used = -INFINITY, l = 0.f;
    for (int hf = 0; hf < nh; ++hf) {
      const int buf = hf & 1;
      mbar_wait(s_full(buf), (hf >> 1) & 1u);
      tc_fence_after();
      const int nvalid = min(64, p.Lk - hf * 64);
      uint32_t v[2][32];
      tmem_ld32(tS(buf) + lane_addr, v[0]);
      tmem_ld32(tS(buf) + lane_addr + 32u, v[1]);
      tmem_ld_wait();
      if (nvalid < 64) {                              // keys past Lk: exp2(-inf) = 0 and they never win the max
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= nvalid) v[c][i] = 0xff800000u;
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};     // four chains instead of one 32-deep one
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if (hf == 0) {
        m_used = mx;
      } else if (__any_sync(0xffffffffu, mx > m_used + 5.545177f)) {     // 8 / log2(e)
        const float m_new = fmaxf(m_used, mx);
        const float sc = fast_exp2((m_used - m_new) * kLog2e);
        mbar_wait(pv_done((hf - 1) & 1), ((hf - 1) >> 1) & 1u);          // nobody accumulates into O right now
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + lane_addr + 32u * c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
          tmem_st32(tO + lane_addr + 32u * c, o);
        }
        l *= sc;
        m_used = m_new;
      }
      // (fetching the next half tile's scores here, behind the exponentials, was tried: 574 us vs 444 us -- 168 registers
      // with spills under the two-CTAs-per-SM cap)
      const uint64_t nmb = f2_splat(-m_used * kLog2e), l2e = f2_splat(kLog2e);
      uint64_t acc0 = f2_splat(0.f), acc1 = f2_splat(0.f);
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t xx = f2_fma(f2_pack(__uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1])), l2e, nmb);
          float p0, p1;
          if (NS_ATTN_POLY_EVERY > 0 && (16 * c + i) % (NS_ATTN_POLY_EVERY > 0 ? NS_ATTN_POLY_EVERY : 1) == NS_ATTN_POLY_EVERY - 1) {
            poly_exp2_pair(xx, p0, p1);                 // FMA pipe
          } else {
            float x0, x1;
            f2_unpack(xx, x0, x1);
            p0 = fast_exp2(x0); p1 = fast_exp2(x1);      // MUFU pipe
          }
          if (i & 1) acc1 = f2_add(acc1, f2_pack(p0, p1)); else acc0 = f2_add(acc0, f2_pack(p0, p1));
          pk[16 * c + i] = pack_bf16x2(p0, p1);
        }
      }
      tmem_st32(tS(buf) + lane_addr, pk);             // P over the first 32 columns of its own score buffer
      float a0, a1, a2, a3;
      f2_unpack(acc0, a0, a1); f2_unpack(acc1, a2, a3);
      l += (a0 + a1) + (a2 + a3);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(buf));
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int qi = q0 + row;
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = p.o + b * p.o_bs + static_cast<long long>(qi) * p.o_rs + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tO + lane_addr + 32u * c, v);
      tmem_ld_wait();
      if (qi < p.Lq) {
        uint4* dst = reinterpret_cast<uint4*>(orow + 32 * c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
          dst[i] = u;
        }
      }
    }
    if (p.lse && qi < p.Lq) p.lse[(static_cast<long long>(b) * p.H + h) * p.Lq + qi] = m_used + logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward, ping-pong variant for long query axes (the encoder: Lq = 1500).  One CTA per SM owns TWO 128-query tiles of one
// (batch, head) and streams the key/value tiles once for both:
//   warp 0        TMA producer : Q0, Q1 once; K / V tiles of 128 keys through 3-stage rings
//   warp 1        MMA issuer   : one elected thread; S_w = Q_w K^T (SS, N = 128), O_w += P_w V (TS, N = 64) for w = 0, 1,
//                                issued S0 S1 | PV0 S0' PV1 S1' | ... so that while softmax group 0 works on its tile the
//                                tensor pipe serves group 1 and vice versa
//   warps 4..7    softmax group 0 (query tile 0), warps 8..11 softmax group 1: one query row per thread
// The exponentials (MUFU, 16 per clock per SM) bound this kernel at head_dim 64; with the two groups half a period apart
// the MUFU pipe always has one group feeding it, instead of idling during every S / PV round trip.
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512).
constexpr int kF2Threads = 384;
constexpr int kF2Stages = 3;
constexpr int kF2Smem = kTile * (2 + 2 * kF2Stages) + 1024 + 256;

__global__ void __launch_bounds__(kF2Threads, 1)
attn_fwd2_tc_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnFwdProg p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sQ = [&](int w) { return smem_base + kTile * w; };
  auto sK = [&](int s) { return smem_base + kTile * (2 + s); };
  auto sV = [&](int s) { return smem_base + kTile * (2 + kF2Stages + s); };
  const uint32_t bar = smem_base + kTile * (2 + 2 * kF2Stages);
  const uint32_t q_full = bar;
  auto k_full = [&](int s) { return bar + 8u * (1 + s); };
  auto v_full = [&](int s) { return bar + 8u * (4 + s); };
  auto k_empty = [&](int s) { return bar + 8u * (7 + s); };
  auto v_empty = [&](int s) { return bar + 8u * (10 + s); };
  auto s_full = [&](int w) { return bar + 8u * (13 + w); };
  auto p_ready = [&](int w) { return bar + 8u * (15 + w); };
  const uint32_t o_done = bar + 8u * 17;
  const uint32_t tmem_slot = bar + 8u * 18;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Lk + 127) / 128;
  const bool tr_on = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  int tr_n = 0;
  auto trace = [&](int region, long long tag) {
    if (tr_on && tr_n < 512) {
      p.trace[(region * 512 + tr_n) * 2] = tag;
      p.trace[(region * 512 + tr_n) * 2 + 1] = clock64();
      ++tr_n;
    }
  };

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kF2Stages; ++s) { mbar_init(k_full(s), 1); mbar_init(v_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_empty(s), 1); }
    for (int w = 0; w < 2; ++w) { mbar_init(s_full(w), 1); mbar_init(p_ready(w), 4); }
    mbar_init(o_done, 1);
    mbar_fence_init();
    tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * kTile);
      tma_load_3d(&maps.q, q_full, sQ(0), h * 64, q0, b);
      tma_load_3d(&maps.q, q_full, sQ(1), h * 64, q0 + 128, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kF2Stages;
        const uint32_t ph = ((j / kF2Stages) & 1u) ^ 1u;
        mbar_wait(k_empty(s), ph);
        mbar_expect_tx(k_full(s), kTile);
        tma_load_3d(&maps.k, k_full(s), sK(s), h * 64, j * 128, b);
        mbar_wait(v_empty(s), ph);
        mbar_expect_tx(v_full(s), kTile);
        tma_load_3d(&maps.v, v_full(s), sV(s), h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idescS = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idescO = umma_idesc_bf16(128, 64, 0, 1);          // B = V is MN-major (keys x head_dim rows)
      const uint64_t qd0 = umma_smem_desc(sQ(0), 16, 1024), qd1 = umma_smem_desc(sQ(1), 16, 1024);
      const uint64_t kd0 = umma_smem_desc(sK(0), 16, 1024), vd0 = umma_smem_desc(sV(0), 8192, 1024);
      auto issue_S = [&](int w, int j) {
        const uint64_t kd = kd0 + 1024u * (j % kF2Stages);
        const uint64_t qd = w ? qd1 : qd0;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + 128u * w, qd + 2u * k, kd + 2u * k, idescS, k > 0);
        umma_commit(s_full(w));
      };
      auto issue_PV = [&](int w, int j) {
        const uint64_t vd = vd0 + 1024u * (j % kF2Stages);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16_ts(tmem + 256u + 64u * w, tmem + 384u + 64u * w + 8u * k, vd + 128u * k, idescO, (j > 0 || k > 0) ? 1u : 0u);
      };
      mbar_wait(q_full, 0);
      mbar_wait(k_full(0), 0);
      tc_fence_after();
      issue_S(0, 0);
      issue_S(1, 0);
      umma_commit(k_empty(0));
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kF2Stages;
        const uint32_t ph = (j / kF2Stages) & 1u;
        const bool more = j + 1 < n_kv;
        const int st1 = (j + 1) % kF2Stages;
        const uint32_t ph1 = ((j + 1) / kF2Stages) & 1u;
        mbar_wait(p_ready(0), j & 1);
        mbar_wait(v_full(st), ph);
        trace(0, 100 + j);
        tc_fence_after();
        issue_PV(0, j);
        if (more) {
          mbar_wait(k_full(st1), ph1);
          tc_fence_after();
          issue_S(0, j + 1);
        }
        mbar_wait(p_ready(1), j & 1);
        trace(0, 200 + j);
        tc_fence_after();
        issue_PV(1, j);
        umma_commit(v_empty(st));
        if (more) {
          issue_S(1, j + 1);
          umma_commit(k_empty(st1));
        }
      }
      umma_commit(o_done);
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int w = (warp - 4) >> 2;                      // softmax group = query tile
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem + 128u * w, tO = tmem + 256u + 64u * w, tP = tmem + 384u + 64u * w;
    float m_used = -INFINITY, l = 0.f;
    // Optional strict alternation of the two groups' exponential phases (token through named barriers 4 / 5, the warpgroup
    // ping-pong of FlashAttention-3).  Measured on B200 it LOSES here (627 us vs 543 us per encoder layer): one warp per
    // scheduler cannot overlap its own MUFU issue windows (~14 issue cycles per element), two concurrent warps can; so the
    // default lets both groups run together and only NS_ATTN_TAKE_TURNS=1 enables the token.
    const bool take_turns = p.take_turns != 0;
    if (take_turns && w == 1) named_bar_arrive(4, 256);
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full(w), j & 1);
      if (take_turns) named_bar_sync(4 + w, 256);
      if (quarter == 0 && lane == 0) trace(2 + w, 100 + j);
      tc_fence_after();
      const int nvalid = min(128, p.Lk - j * 128);
      const bool full = nvalid == 128;
      uint32_t v[32];
      if (j == 0) {                                   // first tile: the row max has to be known before the exponentials
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(tS + lane_addr + 32u * c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (full || c * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        m_used = mx;
      }
      // One pass: p = exp2((s - m_used) * log2e) with the max of the PREVIOUS tiles, tracking this tile's max on the side.
      // Only if some row's max grew by more than 8 (log2 units) the accumulator is rescaled and the pass is redone.
      float l_tile = 0.f, mx = -INFINITY;
      for (int attempt = 0; attempt < 2; ++attempt) {
        const float mb = m_used * kLog2e;
        l_tile = 0.f;
        // software pipeline over the four 32-column chunks: the tcgen05.ld of chunk c+1 is in flight while chunk c is
        // exponentiated, so this warp (the only one of its group on this scheduler) keeps the MUFU pipe fed
        uint32_t vb[2][32];
        tmem_ld32(tS + lane_addr, vb[0]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < 3) tmem_ld32(tS + lane_addr + 32u * (c + 1), vb[(c + 1) & 1]);
          const uint32_t (&vc)[32] = vb[c & 1];
          uint32_t pk[16];
          if (full) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float s0 = __uint_as_float(vc[2 * i]), s1 = __uint_as_float(vc[2 * i + 1]);
              mx = fmaxf(mx, fmaxf(s0, s1));
              const float p0 = fast_exp2(fmaf(s0, kLog2e, -mb)), p1 = fast_exp2(fmaf(s1, kLog2e, -mb));
              l_tile += p0 + p1;
              pk[i] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const bool ok0 = c * 32 + 2 * i < nvalid, ok1 = c * 32 + 2 * i + 1 < nvalid;
              const float s0 = __uint_as_float(vc[2 * i]), s1 = __uint_as_float(vc[2 * i + 1]);
              if (ok0) mx = fmaxf(mx, s0);
              if (ok1) mx = fmaxf(mx, s1);
              const float p0 = ok0 ? fast_exp2(fmaf(s0, kLog2e, -mb)) : 0.f;
              const float p1 = ok1 ? fast_exp2(fmaf(s1, kLog2e, -mb)) : 0.f;
              l_tile += p0 + p1;
              pk[i] = pack_bf16x2(p0, p1);
            }
          }
          tmem_st16(tP + lane_addr + 16u * c, pk);
          if (c < 3) tmem_ld_wait();
        }
        const bool need = mx > m_used + 5.545177f;     // 8 / log2(e)
        if (attempt == 1 || !__any_sync(0xffffffffu, need)) break;
        // s_full(j) was committed behind PV(j-1) of this tile, so the accumulator is quiescent here
        const float m_new = fmaxf(m_used, mx);
        const float sc = fast_exp2((m_used - m_new) * kLog2e);
        tmem_st_wait();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          tmem_ld32(tO + lane_addr + 32u * c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * sc);
          tmem_st32(tO + lane_addr + 32u * c, v);
        }
        l *= sc;
        m_used = m_new;
      }
      l += l_tile;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(w));
      if (take_turns && (w == 0 || j + 1 < n_kv)) named_bar_arrive(5 - w, 256);
      if (quarter == 0 && lane == 0) trace(2 + w, 200 + j);
    }
    mbar_wait(o_done, 0);
    tc_fence_after();
    const int qi = q0 + 128 * w + row;
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = p.o + b * p.o_bs + static_cast<long long>(qi) * p.o_rs + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tO + lane_addr + 32u * c, v);
      tmem_ld_wait();
      if (qi < p.Lq) {
        uint4* dst = reinterpret_cast<uint4*>(orow + 32 * c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
          dst[i] = u;
        }
      }
    }
    if (p.lse && qi < p.Lq) p.lse[(static_cast<long long>(b) * p.H + h) * p.Lq + qi] = m_used + logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int head_map(CUtensorMap* m, const void* base, int H, int L, int B, long long bs, long long rs) {
  uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)L, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)rs * 2, (uint64_t)bs * 2};
  uint32_t box[3] = {64, 128, 1};
  return make_map(m, base, 3, dims, str, box);
}

static bool tc_eligible(const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o) {
  return s.Dh == 64 && !s.causal && al16(q) && al16(k) && al16(v) && al16(o) && s.q_rs % 8 == 0 && s.k_rs % 8 == 0 &&
         s.v_rs % 8 == 0 && s.o_rs % 8 == 0 && s.q_bs % 8 == 0 && s.k_bs % 8 == 0 && s.v_bs % 8 == 0 && s.o_bs % 8 == 0 &&
         s.Lk >= 1 && s.H <= 65535 && s.B <= 65535;
}

int attention_fwd_tc(const ns_attn_shape& s, const void* q, const void* k, const void* v, void* o, float* lse, cudaStream_t st) {
  if (!tc_eligible(s, q, k, v, o)) return NS_ERR_UNSUPPORTED;
  static bool attr_done = false;
  if (!attr_done) {
    NS_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    attr_done = true;
  }
  AttnMaps maps;
  int r;
  if ((r = head_map(&maps.q, q, s.H, s.Lq, s.B, s.q_bs, s.q_rs))) return r;
  if ((r = head_map(&maps.k, k, s.H, s.Lk, s.B, s.k_bs, s.k_rs))) return r;
  if ((r = head_map(&maps.v, v, s.H, s.Lk, s.B, s.v_bs, s.v_rs))) return r;
  AttnFwdProg prog{s.B, s.H, s.Lq, s.Lk, s.o_bs, s.o_rs, reinterpret_cast<__nv_bfloat16*>(o), lse, get_attn_trace(),
                   getenv("NS_ATTN_TAKE_TURNS") != nullptr};
  // two-query-tile kernel: same speed as the 2-CTAs-per-SM kernel below on B200 (543 us per encoder layer) with half the
  // K/V traffic; opt-in until it wins
  static const bool pingpong = getenv("NS_ATTN_PINGPONG") != nullptr;
  if (s.Lq > 128 && pingpong) {
    static bool attr2_done = false;
    if (!attr2_done) {
      NS_CUDA(cudaFuncSetAttribute(attn_fwd2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2Smem));
      attr2_done = true;
    }
    dim3 grid2((s.Lq + 255) / 256, s.H, s.B);
    attn_fwd2_tc_kernel<<<grid2, kF2Threads, kF2Smem, st>>>(maps, prog);
    NS_LAUNCH_CHECK();
    count(C_ATTN_TC);
    return NS_OK;
  }
  dim3 grid((s.Lq + 127) / 128, s.H, s.B);
  static const bool single_buffer = getenv("NS_ATTN_FWD_SINGLE") != nullptr;     // the first kernel, kept for A/B runs
  if (single_buffer) {
    attn_fwd_tc_kernel<<<grid, kAtThreads, kAtSmem, st>>>(maps, prog);
  } else {
    static bool attr3_done = false;
    if (!attr3_done) {
      NS_CUDA(cudaFuncSetAttribute(attn_fwd_db_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
      attr3_done = true;
    }
    attn_fwd_db_kernel<<<grid, kDbThreads, kAtSmem, st>>>(maps, prog);
  }
  NS_LAUNCH_CHECK();
  count(C_ATTN_TC);
  return NS_OK;
}

// =================================================================================================== backward
// Two kernels, no atomics (deterministic):
//   attn_bwd_dkdv_tc_kernel : CTA = (b, h, 128 keys).  TMEM lanes = keys.  Per 64-query step:
//        S^T = K Q^T, dP^T = V dO^T (SS) -> P^T = exp2(S^T*log2e - lse[q]), dS^T = P^T (dP^T - delta[q]) (8 compute warps,
//        lse/delta per COLUMN via broadcast shared loads) -> bf16 P^T / dS^T written over the consumed S^T / dP^T columns
//        -> dV += P^T dO, dK += dS^T Q  (TS: A from TMEM, B = dO / Q tile MN-major).
//   attn_bwd_dq_tc_kernel   : CTA = (b, h, 128 queries).  TMEM lanes = queries (lse/delta are per-thread scalars).  Per 64-key
//        step: S = Q K^T, dP = dO V^T -> dS = P (dP - delta) -> dQ += dS K  (TS, B = K tile MN-major).
// TMEM budget 256 columns per CTA so that two CTAs share an SM (one computes while the other's MMAs run).
struct AttnBwdMaps {
  CUtensorMap q, k, v, d_o;     // box {64, 128} for the resident operand, {64, 64} for the streamed one (see host code)
};
struct AttnBwdProg {
  int B, H, Lq, Lk;
  const float* lse;
  const float* delta;
  long long g_bs, g_rs, g2_bs, g2_rs;   // output strides: dkdv kernel -> (dk, dv); dq kernel -> (dq, unused)
  __nv_bfloat16* out0;
  __nv_bfloat16* out1;
};

constexpr int kBwThreads = 320;                       // warp0 TMA, warp1 MMA, warps 2..9 compute
constexpr int kHalf = 64 * 64 * 2;                    // [64][64] bf16 tile
constexpr int kBwSmem = 2 * kTile + 4 * kHalf + 1024 + 2048 + 256;

// column offset (in 32-bit TMEM columns) of the bf16 pair block for step-local index k16 (16 contraction elements)
__device__ __forceinline__ uint32_t pk_col(int k16) { return static_cast<uint32_t>((k16 >> 1) * 32 + (k16 & 1) * 8); }

template <bool kDQ>
__global__ void __launch_bounds__(kBwThreads, 2)
attn_bwd_tc_kernel(const __grid_constant__ AttnBwdMaps maps, const __grid_constant__ AttnBwdProg p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // resident 128-row tiles: dkdv -> K, V ; dq -> Q, dO.   streamed 64-row tiles (2 stages): dkdv -> Q, dO ; dq -> K, V
  const uint32_t sR0 = smem_base, sR1 = smem_base + kTile;
  auto sS0 = [&](int s) { return smem_base + 2 * kTile + kHalf * (2 * s); };
  auto sS1 = [&](int s) { return smem_base + 2 * kTile + kHalf * (2 * s + 1); };
  const uint32_t stat_off = 2 * kTile + 4 * kHalf;                     // float lse_s[2][64], delta_s[2][64]
  float* stat = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + stat_off);
  const uint32_t bar = smem_base + stat_off + 2048;
  const uint32_t r_full = bar;
  auto s_full = [&](int s) { return bar + 8u * (1 + s); };
  auto s_empty = [&](int s) { return bar + 8u * (3 + s); };
  const uint32_t sdp_full = bar + 8u * 5, pds_full = bar + 8u * 6, acc_done = bar + 8u * 7;
  const uint32_t tmem_slot = bar + 8u * 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int Lstream = kDQ ? p.Lk : p.Lq;
  const int n_steps = (Lstream + 63) / 64;

  if (warp == 0 && lane == 0) {
    mbar_init(r_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(s_full(s), 1); mbar_init(s_empty(s), 1); }
    mbar_init(sdp_full, 1); mbar_init(pds_full, 8); mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tdP = tmem + 64, tA0 = tmem + 128, tA1 = tmem + 192;   // accumulators: dkdv -> dV, dK ; dq -> dQ

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(r_full, 2 * kTile);
      if (kDQ) {
        tma_load_3d(&maps.q, r_full, sR0, h * 64, r0, b);
        tma_load_3d(&maps.d_o, r_full, sR1, h * 64, r0, b);
      } else {
        tma_load_3d(&maps.k, r_full, sR0, h * 64, r0, b);
        tma_load_3d(&maps.v, r_full, sR1, h * 64, r0, b);
      }
      for (int i = 0; i < n_steps; ++i) {
        const int s = i & 1;
        mbar_wait(s_empty(s), ((i >> 1) & 1u) ^ 1u);
        mbar_expect_tx(s_full(s), 2 * kHalf);
        if (kDQ) {
          tma_load_3d(&maps.k, s_full(s), sS0(s), h * 64, i * 64, b);
          tma_load_3d(&maps.v, s_full(s), sS1(s), h * 64, i * 64, b);
        } else {
          tma_load_3d(&maps.q, s_full(s), sS0(s), h * 64, i * 64, b);
          tma_load_3d(&maps.d_o, s_full(s), sS1(s), h * 64, i * 64, b);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idescS = umma_idesc_bf16(128, 64, 0, 0);        // [128 x 64] = resident(128 rows) x streamed(64 rows)^T
    constexpr uint32_t idescA = umma_idesc_bf16(128, 64, 0, 1);        // accumulators: A from TMEM, B MN-major
    mbar_wait(r_full, 0);
    for (int i = 0; i < n_steps; ++i) {
      const int s = i & 1;
      mbar_wait(s_full(s), (i >> 1) & 1u);
      if (i > 0) mbar_wait(acc_done, (i - 1) & 1u);                    // P/dS columns of step i-1 fully consumed (WAR)
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a0 = umma_smem_desc(sR0, 16, 1024), a1 = umma_smem_desc(sR1, 16, 1024);
        const uint64_t b0 = umma_smem_desc(sS0(s), 16, 1024), b1 = umma_smem_desc(sS1(s), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tS, a0 + 2u * k, b0 + 2u * k, idescS, k > 0);     // dkdv: K Q^T ; dq: Q K^T
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tdP, a1 + 2u * k, b1 + 2u * k, idescS, k > 0);    // dkdv: V dO^T; dq: dO V^T
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(pds_full, i & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t acc = (i > 0) ? 1u : 0u;
        if (kDQ) {
          const uint64_t kd = umma_smem_desc(sS0(s), 4096, 1024);      // K tile [64 keys][64 dh] as MN-major B
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(tA0, tS + pk_col(k), kd + 128u * k, idescA, (acc | (k > 0)) ? 1u : 0u);
        } else {
          const uint64_t od = umma_smem_desc(sS1(s), 4096, 1024);      // dO tile [64 q][64 dh] as MN-major B
          const uint64_t qd = umma_smem_desc(sS0(s), 4096, 1024);      // Q  tile [64 q][64 dh] as MN-major B
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(tA0, tS + pk_col(k), od + 128u * k, idescA, (acc | (k > 0)) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(tA1, tdP + pk_col(k), qd + 128u * k, idescA, (acc | (k > 0)) ? 1u : 0u);
        }
        umma_commit(s_empty(s));
        umma_commit(acc_done);
      }
      __syncwarp();
    }
  } else {
    const int cw = warp - 2;                          // 0..7
    const int quarter = warp & 3;
    const int chalf = cw >> 2;                        // which 32 of the step's 64 columns
    const int row = quarter * 32 + lane;
    const int tid_c = cw * 32 + lane;                 // 0..255
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const long long stat_base = (static_cast<long long>(b) * p.H + h) * p.Lq;
    float my_lse = 0.f, my_delta = 0.f;               // dq kernel: per-row scalars
    if (kDQ) {
      const int qi = r0 + row;
      my_lse = qi < p.Lq ? p.lse[stat_base + qi] * kLog2e : INFINITY;
      my_delta = qi < p.Lq ? p.delta[stat_base + qi] : 0.f;
    }
    // dkdv kernel: stage the per-column lse/delta of step 0
    float pre = 0.f;
    auto fetch_stat = [&](int step) -> float {
      if (tid_c < 64) { const int qi = step * 64 + tid_c; return qi < p.Lq ? p.lse[stat_base + qi] * kLog2e : INFINITY; }
      if (tid_c < 128) { const int qi = step * 64 + tid_c - 64; return qi < p.Lq ? p.delta[stat_base + qi] : 0.f; }
      return 0.f;
    };
    if (!kDQ) {
      pre = fetch_stat(0);
      if (tid_c < 128) stat[tid_c] = pre;             // stage 0: [0,64) lse, [64,128) delta
    }
    for (int i = 0; i < n_steps; ++i) {
      const int s = i & 1;
      if (!kDQ) {
        if (i + 1 < n_steps) pre = fetch_stat(i + 1);
        named_bar_sync(1, 256);                       // stage s visible to all compute warps
      }
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
      const float* lse_s = stat + s * 128;
      const float* del_s = lse_s + 64;
      const int nvalid = min(64, Lstream - i * 64);   // valid streamed rows in this step (columns of S)
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {                   // 16 columns at a time
        const int col0 = chalf * 32 + c * 16;
        uint32_t sv[16], dv[16];
        tmem_ld16(tS + lane_addr + static_cast<uint32_t>(col0), sv);
        tmem_ld16(tdP + lane_addr + static_cast<uint32_t>(col0), dv);
        tmem_ld_wait();
        uint32_t pp[8], ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float pr[2], dd[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = col0 + 2 * e + u;
            const float ls = kDQ ? my_lse : lse_s[col];
            const float dl = kDQ ? my_delta : del_s[col];
            float pv = fast_exp2(fmaf(__uint_as_float(sv[2 * e + u]), kLog2e, -ls));
            if (kDQ && col >= nvalid) pv = 0.f;       // zero-filled keys would otherwise contribute exp(-lse)
            pr[u] = pv;
            dd[u] = pv * (__uint_as_float(dv[2 * e + u]) - dl);
          }
          pp[e] = pack_bf16x2(pr[0], pr[1]);
          ds[e] = pack_bf16x2(dd[0], dd[1]);
        }
        // bf16 pairs for columns [col0, col0+16) -> 8 TMEM columns at pk_col(col0/16); these lie inside the fp32 columns this
        // thread's warp has already consumed ([chalf*32, col0+16)), never in the other warp's half.
        const uint32_t pc = pk_col(col0 >> 4);
        if (kDQ) {
          tmem_st8(tS + lane_addr + pc, ds);          // dq only needs dS
        } else {
          tmem_st8(tS + lane_addr + pc, pp);
          tmem_st8(tdP + lane_addr + pc, ds);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      if (!kDQ && i + 1 < n_steps && tid_c < 128) stat[((i + 1) & 1) * 128 + tid_c] = pre;
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    // ---- write the accumulators
    mbar_wait(acc_done, (n_steps - 1) & 1);
    tc_fence_after();
    const int ri = r0 + row;
    const int Lres = kDQ ? p.Lq : p.Lk;
    if (kDQ) {
      // dQ (64 columns): warp half 0 writes columns 0..31, half 1 columns 32..63
      uint32_t v[32];
      tmem_ld32(tA0 + lane_addr + 32u * chalf, v);
      tmem_ld_wait();
      if (ri < Lres) {
        uint4* dst = reinterpret_cast<uint4*>(p.out0 + b * p.g_bs + static_cast<long long>(ri) * p.g_rs + h * 64 + 32 * chalf);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[8 * e + 0]), __uint_as_float(v[8 * e + 1]));
          u.y = pack_bf16x2(__uint_as_float(v[8 * e + 2]), __uint_as_float(v[8 * e + 3]));
          u.z = pack_bf16x2(__uint_as_float(v[8 * e + 4]), __uint_as_float(v[8 * e + 5]));
          u.w = pack_bf16x2(__uint_as_float(v[8 * e + 6]), __uint_as_float(v[8 * e + 7]));
          dst[e] = u;
        }
      }
    } else {
      // half 0 writes dV (tA0), half 1 writes dK (tA1): 64 columns each
      const uint32_t tacc = chalf == 0 ? tA0 : tA1;
      __nv_bfloat16* base = chalf == 0 ? p.out1 + b * p.g2_bs + static_cast<long long>(ri) * p.g2_rs
                                       : p.out0 + b * p.g_bs + static_cast<long long>(ri) * p.g_rs;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tacc + lane_addr + 32u * c, v);
        tmem_ld_wait();
        if (ri < Lres) {
          uint4* dst = reinterpret_cast<uint4*>(base + h * 64 + 32 * c);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v[8 * e + 0]), __uint_as_float(v[8 * e + 1]));
            u.y = pack_bf16x2(__uint_as_float(v[8 * e + 2]), __uint_as_float(v[8 * e + 3]));
            u.z = pack_bf16x2(__uint_as_float(v[8 * e + 4]), __uint_as_float(v[8 * e + 5]));
            u.w = pack_bf16x2(__uint_as_float(v[8 * e + 6]), __uint_as_float(v[8 * e + 7]));
            dst[e] = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

int attention_delta(int dtype, const ns_attn_shape& s, const void* o, const void* d_o, float* delta, cudaStream_t st);

static int head_map_rows(CUtensorMap* m, const void* base, int H, int L, int B, long long bs, long long rs, int rows) {
  uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)L, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)rs * 2, (uint64_t)bs * 2};
  uint32_t box[3] = {64, (uint32_t)rows, 1};
  return make_map(m, base, 3, dims, str, box);
}

int attention_bwd_tc(const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                     const float* lse, float* delta, void* dq, void* dk, void* dv, cudaStream_t st) {
  if (!tc_eligible(s, q, k, v, o) || !al16(d_o) || !al16(dq) || !al16(dk) || !al16(dv)) return NS_ERR_UNSUPPORTED;
  static bool attr_done = false;
  if (!attr_done) {
    NS_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwSmem));
    NS_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwSmem));
    attr_done = true;
  }
  int r = attention_delta(NS_BF16, s, o, d_o, delta, st);
  if (r) return r;
  {   // dK, dV: resident K, V (128-row boxes); streamed Q, dO (64-row boxes)
    AttnBwdMaps maps;
    if ((r = head_map_rows(&maps.k, k, s.H, s.Lk, s.B, s.k_bs, s.k_rs, 128))) return r;
    if ((r = head_map_rows(&maps.v, v, s.H, s.Lk, s.B, s.v_bs, s.v_rs, 128))) return r;
    if ((r = head_map_rows(&maps.q, q, s.H, s.Lq, s.B, s.q_bs, s.q_rs, 64))) return r;
    if ((r = head_map_rows(&maps.d_o, d_o, s.H, s.Lq, s.B, s.o_bs, s.o_rs, 64))) return r;
    AttnBwdProg prog{s.B, s.H, s.Lq, s.Lk, lse, delta, s.k_bs, s.k_rs, s.v_bs, s.v_rs,
                     reinterpret_cast<__nv_bfloat16*>(dk), reinterpret_cast<__nv_bfloat16*>(dv)};
    dim3 grid((s.Lk + 127) / 128, s.H, s.B);
    attn_bwd_tc_kernel<false><<<grid, kBwThreads, kBwSmem, st>>>(maps, prog);
    NS_LAUNCH_CHECK();
  }
  {   // dQ: resident Q, dO; streamed K, V
    AttnBwdMaps maps;
    if ((r = head_map_rows(&maps.q, q, s.H, s.Lq, s.B, s.q_bs, s.q_rs, 128))) return r;
    if ((r = head_map_rows(&maps.d_o, d_o, s.H, s.Lq, s.B, s.o_bs, s.o_rs, 128))) return r;
    if ((r = head_map_rows(&maps.k, k, s.H, s.Lk, s.B, s.k_bs, s.k_rs, 64))) return r;
    if ((r = head_map_rows(&maps.v, v, s.H, s.Lk, s.B, s.v_bs, s.v_rs, 64))) return r;
    AttnBwdProg prog{s.B, s.H, s.Lq, s.Lk, lse, delta, s.q_bs, s.q_rs, 0, 0, reinterpret_cast<__nv_bfloat16*>(dq), nullptr};
    dim3 grid((s.Lq + 127) / 128, s.H, s.B);
    attn_bwd_tc_kernel<true><<<grid, kBwThreads, kBwSmem, st>>>(maps, prog);
    NS_LAUNCH_CHECK();
  }
  count(C_ATTN_TC, 2);
  return NS_OK;
}
}  // namespace ns
