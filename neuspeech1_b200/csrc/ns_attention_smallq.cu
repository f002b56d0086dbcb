// Attention backward for SHORT query axes (Lq <= 64, head_dim 64, bf16, non-causal) on tcgen05 / TMEM / TMA: the decoder's
// cross-attention in the training step (32 label positions against the 1500 encoder positions, HF modeling_whisper.py:314-336).
//
// One CTA per (batch, head) walks the 128-key tiles once.  Because ALL queries fit in one block, dK and dV of a key tile are
// complete after that tile (no accumulation across tiles, no atomics); only dQ accumulates over the walk, in TMEM.
//   S^T  = K Q^T, dP^T = V dO^T            (SS MMAs, 128 keys x NQ queries, fp32 in TMEM)
//   P^T  = exp2(S^T log2e - lse log2e), dS^T = P^T (dP^T - delta)      4 compute warps, TMEM lane = key, one row per thread
//   dV   = P^T dO, dK = dS^T Q            (TS MMAs, A = the bf16 values written back over the consumed fp32 columns)
//   dQ  += dS K                            (SS MMA, both operands MN-major; dS^T staged in shared memory, query axis
//                                           zero-padded to M = 128)
// delta = rowsum(dO * O) and lse * log2e are computed by the CTA itself (<= 64 rows), so there is no prep kernel and no
// workspace.  The general kernels (ns_attention_tc.cu / ns_attention_bwd_fused.cu) pad the query axis to 128 and pay their
// per-CTA set-up once per key tile: 238 us / 324 us for this shape; the floor is reading K, V and writing dK, dV (~60 us).
#include "ns_common.cuh"
#include "ns_sm100.cuh"

namespace ns {
using namespace sm100;

struct SqMaps {
  CUtensorMap q, k, v, d_o;
};
struct SqProg {
  int B, H, Lq, Lk;
  const __nv_bfloat16* o;
  const __nv_bfloat16* d_o;
  long long o_bs, o_rs;
  const float* lse;
  float* delta;
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs;
  __nv_bfloat16* dq;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
};

constexpr int kSqThreads = 192;             // warp 0 TMA, warp 1 MMA, warps 2..5 compute
constexpr int kSqTile = 128 * 64 * 2;
constexpr float kSqL2e = 1.4426950408889634f;

template <int NQ> struct SqCfg {
  static constexpr int kQBytes = NQ * 128;                                    // [NQ rows][64] bf16
  static constexpr uint32_t kOffQ = 0, kOffdO = kQBytes, kOffK = 2 * kQBytes, kOffV = kOffK + 2 * kSqTile, kOffdS = kOffV + 2 * kSqTile,
                            kOffStat = kOffdS + 2 * kSqTile, kOffBar = kOffStat + 1024;
  static constexpr int kSmem = kOffBar + 128 + 1024;
  static constexpr int kTmemCols = (NQ == 32) ? 256 : 512;
  // TMEM columns
  static constexpr uint32_t tS = 0, tdP = NQ, tdV = 2 * NQ, tdK = 2 * NQ + 64, tdQ = 2 * NQ + 128;
};

__device__ __forceinline__ float sq_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NQ>
__global__ void __launch_bounds__(kSqThreads, 1)
attn_bwd_smallq_kernel(const __grid_constant__ SqMaps maps, const __grid_constant__ SqProg p) {
  using C = SqCfg<NQ>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + C::kOffQ, sdO = base + C::kOffdO, sdS = base + C::kOffdS;
  auto sK = [&](int s) { return base + C::kOffK + kSqTile * s; };
  auto sV = [&](int s) { return base + C::kOffV + kSqTile * s; };
  float* stat = reinterpret_cast<float*>(base_ptr + C::kOffStat);            // [0,NQ) lse*log2e, [64,64+NQ) delta
  const uint32_t bar = base + C::kOffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (3 + s); };
  const uint32_t sdp_full = bar + 8u * 5, pds_ready = bar + 8u * 6, acc_full = bar + 8u * 7, acc_free = bar + 8u * 8;
  const uint32_t tmem_slot = bar + 8u * 9;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(base_ptr + C::kOffBar + 8 * 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int n_kt = (p.Lk + 127) / 128;

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(sdp_full, 1); mbar_init(pds_ready, 4); mbar_init(acc_full, 1); mbar_init(acc_free, 4);
    mbar_fence_init();
    tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v); tma_prefetch_desc(&maps.d_o);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  // zero the dS^T staging tile once: the query columns beyond NQ (M is padded to 128 for the dQ product) must read as zero
  for (uint32_t off = threadIdx.x * 16u; off < 2u * kSqTile; off += kSqThreads * 16u) st_shared_v4(sdS + off, 0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * C::kQBytes);
      tma_load_3d(&maps.q, q_full, sQ, h * 64, 0, b);
      tma_load_3d(&maps.d_o, q_full, sdO, h * 64, 0, b);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j & 1;
        mbar_wait(kv_empty(s), ((j >> 1) & 1u) ^ 1u);
        mbar_expect_tx(kv_full(s), 2 * kSqTile);
        tma_load_3d(&maps.k, kv_full(s), sK(s), h * 64, j * 128, b);
        tma_load_3d(&maps.v, kv_full(s), sV(s), h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (one elected thread)
    if (elect_one()) {
      constexpr uint32_t idS = umma_idesc_bf16(128, NQ, 0, 0);   // S^T / dP^T: resident(128 keys) x (NQ queries)^T, both K-major
      constexpr uint32_t idA = umma_idesc_bf16(128, 64, 0, 1);   // dV / dK   : A from TMEM, B = dO / Q MN-major
      constexpr uint32_t idQ = umma_idesc_bf16(128, 64, 1, 1);   // dQ        : A = dS^T (smem, MN-major), B = K MN-major
      const uint64_t qd_k = umma_smem_desc(sQ, 16, 1024), od_k = umma_smem_desc(sdO, 16, 1024);
      const uint64_t qd_mn = umma_smem_desc(sQ, 16384, 1024), od_mn = umma_smem_desc(sdO, 16384, 1024);
      const uint64_t ds_mn = umma_smem_desc(sdS, 16384, 1024);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j & 1;
        const uint64_t kd = umma_smem_desc(sK(s), 16, 1024), vd = umma_smem_desc(sV(s), 16, 1024);
        const uint64_t kd_mn = umma_smem_desc(sK(s), 16384, 1024);
        mbar_wait(kv_full(s), (j >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + C::tS, kd + 2u * k, qd_k + 2u * k, idS, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + C::tdP, vd + 2u * k, od_k + 2u * k, idS, k > 0);
        umma_commit(sdp_full);
        mbar_wait(pds_ready, j & 1);
        if (j > 0) mbar_wait(acc_free, (j - 1) & 1);            // dV / dK of the previous tile have been read out
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < NQ / 16; ++k) umma_f16_ts(tmem + C::tdV, tmem + C::tS + 8u * k, od_mn + 128u * k, idA, k > 0);
#pragma unroll
        for (int k = 0; k < NQ / 16; ++k) umma_f16_ts(tmem + C::tdK, tmem + C::tdP + 8u * k, qd_mn + 128u * k, idA, k > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_f16(tmem + C::tdQ, ds_mn + 128u * k, kd_mn + 128u * k, idQ, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(kv_empty(s));
        umma_commit(acc_full);
      }
    }
    __syncwarp();
  } else {
    // ================================================================ compute warps (TMEM lane = key row of the tile)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    // softmax statistics of the (<= NQ) query rows: lse * log2e and delta = dO . O
    {
      const int t = threadIdx.x - 64;                            // 0..127
      if (t < NQ) {
        float l2 = INFINITY, dl = 0.f;
        if (t < p.Lq) {
          const long long li = (static_cast<long long>(b) * p.H + h) * p.Lq + t;
          l2 = p.lse[li] * kSqL2e;
          const uint4* po = reinterpret_cast<const uint4*>(p.o + b * p.o_bs + static_cast<long long>(t) * p.o_rs + h * 64);
          const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + b * p.o_bs + static_cast<long long>(t) * p.o_rs + h * 64);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint4 x = __ldg(po + e), y = __ldg(pd + e);
            float2 f, g;
            f = unpack_bf16x2(x.x); g = unpack_bf16x2(y.x); dl = fmaf(f.x, g.x, dl); dl = fmaf(f.y, g.y, dl);
            f = unpack_bf16x2(x.y); g = unpack_bf16x2(y.y); dl = fmaf(f.x, g.x, dl); dl = fmaf(f.y, g.y, dl);
            f = unpack_bf16x2(x.z); g = unpack_bf16x2(y.z); dl = fmaf(f.x, g.x, dl); dl = fmaf(f.y, g.y, dl);
            f = unpack_bf16x2(x.w); g = unpack_bf16x2(y.w); dl = fmaf(f.x, g.x, dl); dl = fmaf(f.y, g.y, dl);
          }
          if (p.delta) p.delta[li] = dl;
        }
        stat[t] = l2;
        stat[64 + t] = dl;
      }
      named_bar_sync(1, 128);
    }
    for (int j = 0; j < n_kt; ++j) {
      mbar_wait(sdp_full, j & 1);
      tc_fence_after();
      uint32_t pk[NQ / 2], dk[NQ / 2];
#pragma unroll
      for (int c = 0; c < NQ / 32; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld32(tmem + C::tS + lane_addr + 32u * c, sv);
        tmem_ld32(tmem + C::tdP + lane_addr + 32u * c, dv);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 l4 = *reinterpret_cast<const float4*>(stat + 32 * c + 4 * q4);
          const float4 d4 = *reinterpret_cast<const float4*>(stat + 64 + 32 * c + 4 * q4);
          const float p0 = sq_ex2(fmaf(__uint_as_float(sv[4 * q4 + 0]), kSqL2e, -l4.x));
          const float p1 = sq_ex2(fmaf(__uint_as_float(sv[4 * q4 + 1]), kSqL2e, -l4.y));
          const float p2 = sq_ex2(fmaf(__uint_as_float(sv[4 * q4 + 2]), kSqL2e, -l4.z));
          const float p3 = sq_ex2(fmaf(__uint_as_float(sv[4 * q4 + 3]), kSqL2e, -l4.w));
          pk[16 * c + 2 * q4] = pack_bf16x2(p0, p1);
          pk[16 * c + 2 * q4 + 1] = pack_bf16x2(p2, p3);
          dk[16 * c + 2 * q4] = pack_bf16x2(p0 * (__uint_as_float(dv[4 * q4 + 0]) - d4.x), p1 * (__uint_as_float(dv[4 * q4 + 1]) - d4.y));
          dk[16 * c + 2 * q4 + 1] = pack_bf16x2(p2 * (__uint_as_float(dv[4 * q4 + 2]) - d4.z), p3 * (__uint_as_float(dv[4 * q4 + 3]) - d4.w));
        }
      }
      // bf16 P^T / dS^T over the consumed fp32 columns of this thread's row (A operands of dV / dK)
      if constexpr (NQ == 32) {
        tmem_st16(tmem + C::tS + lane_addr, pk);
        tmem_st16(tmem + C::tdP + lane_addr, dk);
      } else {
        tmem_st32(tmem + C::tS + lane_addr, pk);
        tmem_st32(tmem + C::tdP + lane_addr, dk);
      }
      // dS^T row (this key) -> shared staging, query block 0, 8 queries per 16-byte chunk, 128B swizzle
#pragma unroll
      for (int e = 0; e < NQ / 8; ++e)
        st_shared_v4(sdS + static_cast<uint32_t>(row) * 128u + ((static_cast<uint32_t>(e) ^ sw) << 4), dk[4 * e], dk[4 * e + 1], dk[4 * e + 2],
                     dk[4 * e + 3]);
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);
      // ---- dV / dK rows of this key tile are final: TMEM -> global
      mbar_wait(acc_full, j & 1);
      tc_fence_after();
      const int ki = j * 128 + row;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* out = which == 0 ? p.dv + b * p.v_bs + static_cast<long long>(ki) * p.v_rs + h * 64
                                        : p.dk + b * p.k_bs + static_cast<long long>(ki) * p.k_rs + h * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem + (which == 0 ? C::tdV : C::tdK) + lane_addr + 32u * c, v);
          tmem_ld_wait();
          if (ki < p.Lk) {
            uint4* dst = reinterpret_cast<uint4*>(out + 32 * c);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
    }
    // ---- dQ (lanes = query rows; acc_full of the last tile covers the last dQ product)
    if (quarter * 32 < p.Lq) {                          // warp-uniform: tcgen05.ld is .sync.aligned (whole warp or nobody)
      __nv_bfloat16* out = p.dq + b * p.q_bs + static_cast<long long>(row) * p.q_rs + h * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + C::tdQ + lane_addr + 32u * c, v);
        tmem_ld_wait();
        if (row < p.Lq) {
          uint4* dst = reinterpret_cast<uint4*>(out + 32 * c);
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
    tmem_dealloc(tmem, C::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------- host side
static bool sq_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int sq_map(CUtensorMap* m, const void* base, int H, int L, int B, long long bs, long long rs, int rows) {
  uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)L, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)rs * 2, (uint64_t)bs * 2};
  uint32_t box[3] = {64, (uint32_t)rows, 1};
  return make_map(m, base, 3, dims, str, box);
}

template <int NQ>
static int launch_smallq(const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                         const float* lse, float* delta, void* dq, void* dk, void* dv, cudaStream_t st) {
  using C = SqCfg<NQ>;
  static bool attr_done = false;
  if (!attr_done) {
    NS_CUDA(cudaFuncSetAttribute(attn_bwd_smallq_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    attr_done = true;
  }
  SqMaps maps;
  int r;
  if ((r = sq_map(&maps.q, q, s.H, s.Lq, s.B, s.q_bs, s.q_rs, NQ))) return r;
  if ((r = sq_map(&maps.d_o, d_o, s.H, s.Lq, s.B, s.o_bs, s.o_rs, NQ))) return r;
  if ((r = sq_map(&maps.k, k, s.H, s.Lk, s.B, s.k_bs, s.k_rs, 128))) return r;
  if ((r = sq_map(&maps.v, v, s.H, s.Lk, s.B, s.v_bs, s.v_rs, 128))) return r;
  SqProg prog{s.B, s.H, s.Lq, s.Lk, reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(d_o), s.o_bs, s.o_rs,
              lse, delta, s.q_bs, s.q_rs, s.k_bs, s.k_rs, s.v_bs, s.v_rs, reinterpret_cast<__nv_bfloat16*>(dq),
              reinterpret_cast<__nv_bfloat16*>(dk), reinterpret_cast<__nv_bfloat16*>(dv)};
  attn_bwd_smallq_kernel<NQ><<<dim3(s.H, s.B), kSqThreads, C::kSmem, st>>>(maps, prog);
  NS_LAUNCH_CHECK();
  count(C_ATTN_TC);
  return NS_OK;
}

int attention_bwd_smallq(const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                         const float* lse, float* delta, void* dq, void* dk, void* dv, cudaStream_t st) {
  if (s.Dh != 64 || s.causal || s.Lq > 64 || s.Lk < 1 || s.H > 65535 || s.B > 65535 || !sq_al16(q) || !sq_al16(k) || !sq_al16(v) ||
      !sq_al16(o) || !sq_al16(d_o) || !sq_al16(dq) || !sq_al16(dk) || !sq_al16(dv) || s.q_rs % 8 || s.k_rs % 8 || s.v_rs % 8 || s.o_rs % 8 ||
      s.q_bs % 8 || s.k_bs % 8 || s.v_bs % 8 || s.o_bs % 8)
    return NS_ERR_UNSUPPORTED;
  if (s.Lq <= 32) return launch_smallq<32>(s, q, k, v, o, d_o, lse, delta, dq, dk, dv, st);
  return launch_smallq<64>(s, q, k, v, o, d_o, lse, delta, dq, dk, dv, st);
}
}  // namespace ns
