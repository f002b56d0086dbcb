// Fused flash-attention backward on tcgen05 / TMEM / TMA (head_dim 64, bf16, non-causal): ONE pass over the score tiles.
//
// CTA = (batch, head, 128-key tile); it walks the 128-query tiles of that (batch, head):
//   S^T  = K Q^T            (SS MMA, 128 keys x 128 queries, fp32 in TMEM)
//   dP^T = V dO^T           (SS MMA)
//   P^T  = exp2(S^T log2e - lse[q] log2e)                    8 compute warps, TMEM lane = key, one row per thread
//   dS^T = P^T * (dP^T - delta[q])
//   dV  += P^T  dO          (TS MMA: A = bf16 P^T written back over the consumed S^T columns)
//   dK  += dS^T Q           (TS MMA: A = bf16 dS^T over the consumed dP^T columns)
//   dQ_i = dS K             (SS MMA, both operands MN-major: A = dS^T staged in shared memory, B = the resident K tile)
//          -> TMEM -> shared (fp32, 128B swizzle) -> TMA reduce-add into an fp32 accumulation buffer (L2 atomics),
//          because the 12 key-tile CTAs of one (batch, head) all contribute to the same dQ rows.
// Compared with the two-kernel scheme (ns_attention_tc.cu) the exponentials are evaluated once instead of twice and five
// matrix products are issued per tile instead of seven; the softmax statistics (lse * log2e, delta = rowsum(dO * O)) are
// produced by a small prep kernel in tile-major order so the TMA warp streams them with the Q tile.
//
// Warp roles (704 threads, 1 CTA / SM): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 dQ epilogue,
// warps 6..13 and 14..21 two compute groups (exp + dS on alternating half tiles).  TMEM columns: [0,128) two S^T / P^T half-tile buffers, [128,256) two dP^T / dS^T
// buffers, [256,320) dV, [320,384) dK, [384,448) dQ, [448,512) bf16 K and V (A operands of S^T / dP^T).
// MMAs of one thread execute in issue order, which is what makes the in-place bf16 overwrites safe: S^T(j+2) is issued
// after dV(j), dP^T(j+2) after dK(j).
#include "ns_common.cuh"
#include "ns_sm100.cuh"

#include <stdlib.h>

#include <type_traits>

namespace ns {
using namespace sm100;

struct BwfMaps {
  CUtensorMap q, k, v, d_o, dqacc;
};
struct BwfProg {
  int B, H, Lq, Lk, nqt;
  const float* stats;                   // [B][H][nqt][2][128]: lse*log2e (+inf beyond Lq), delta (0 beyond Lq)
  long long dk_bs, dk_rs, dv_bs, dv_rs;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  int debug;                            // NS_BWF_DEBUG bit 0: skip the dQ reduce-add (timing experiments only)
  long long* trace;                     // optional timeline of CTA (0,0,0): 4 regions x 512 x {tag, clock} (ns_debug_attn_trace)
};

constexpr int kBfThreads = 704;        // warp 0 TMA, 1 MMA, 2..5 dQ epilogue, 6..13 / 14..21 compute groups
constexpr int kT16 = 128 * 64 * 2;      // [128][64] bf16 tile
// shared memory map: K | V (later Q stage 2) | Q stages 0,1 | dO stages 0..2 | dS^T staging x2 | dQ staging | statistics x3 | barriers
constexpr uint32_t kOffK = 0, kOffV = kT16, kOffQ = 2 * kT16, kOffdO = 4 * kT16, kOffdS = 7 * kT16 /* 2 x 32 KB */,
                   kOffdQ = 11 * kT16, kOffStat = 13 * kT16, kOffBar = 13 * kT16 + 3072;
constexpr int kBfSmem = kOffBar + 256 + 1024;
constexpr float kL2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t bf16x2_mul(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#define NS_TRACE(region, tag)                                                                         \
  do {                                                                                              \
    if (tr_on && lane == 0 && tr_n < 512) {                                                         \
      p.trace[((region) * 512 + tr_n) * 2] = (tag);                                                 \
      p.trace[((region) * 512 + tr_n) * 2 + 1] = clock64();                                         \
      ++tr_n;                                                                                       \
    }                                                                                               \
  } while (0)

#define NS_TRACE1(region, tag)   /* single-thread variant (inside an elected region) */                \
  do {                                                                                              \
    if (tr_on && tr_n < 512) {                                                                      \
      p.trace[((region) * 512 + tr_n) * 2] = (tag);                                                 \
      p.trace[((region) * 512 + tr_n) * 2 + 1] = clock64();                                         \
      ++tr_n;                                                                                       \
    }                                                                                               \
  } while (0)

// The query axis is walked in HALF tiles of 64 (index j, tile t = j/2): S^T / dP^T are 128 keys x 64 queries and
// double-buffered in TMEM (buffer j&1), so the exp warps work on half tile j+1 while the dS warps work on j and the tensor
// pipe runs the dV/dK/dQ products of j-1 -- three stages in flight.  dQ is issued once per full tile (M = 128 queries).
// Q / dO tiles (128 rows) stream through a 3-stage ring (stage t%3; the third Q stage reuses the V tile's shared memory
// once V has been copied to TMEM), which leaves ~2.5 half-tile steps for each TMA load to land.
__global__ void __launch_bounds__(kBfThreads, 1)
attn_bwd_fused_kernel(const __grid_constant__ BwfMaps maps, const __grid_constant__ BwfProg p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sK = base + kOffK, sV = base + kOffV, sdQ = base + kOffdQ;
  auto sdS = [&](int s) { return base + kOffdS + 2 * kT16 * s; };       // dS^T staging, double-buffered per query tile
  auto sQ = [&](int s) { return s == 2 ? sV : base + kOffQ + kT16 * s; };
  auto sdO = [&](int s) { return base + kOffdO + kT16 * s; };
  const float* stat_s = reinterpret_cast<const float*>(base_ptr + kOffStat);
  const uint32_t bar = base + kOffBar;
  auto q_full = [&](int s) { return bar + 8u * (0 + s); };
  auto q_empty = [&](int s) { return bar + 8u * (3 + s); };
  auto do_full = [&](int s) { return bar + 8u * (6 + s); };
  auto do_empty = [&](int s) { return bar + 8u * (9 + s); };
  auto s_full = [&](int s) { return bar + 8u * (12 + s); };
  auto p_ready = [&](int s) { return bar + 8u * (14 + s); };
  auto dp_full = [&](int s) { return bar + 8u * (18 + s); };
  auto ds_ready = [&](int s) { return bar + 8u * (20 + s); };
  auto ds_free = [&](int s) { return bar + 8u * (22 + s); };
  const uint32_t kv_full = bar + 8u * 24, dq_full = bar + 8u * 25, dq_empty = bar + 8u * 26, acc_done = bar + 8u * 27,
                 kvt_ready = bar + 8u * 28;
  const uint32_t dbg_bar = bar + 8u * 29;
  const uint32_t tmem_slot = bar + 8u * 30;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(base_ptr + kOffBar + 8 * 30);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nqt = p.nqt;
  const int nh = 2 * nqt;
  const bool tr_on = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  int tr_n = 0;
  if (warp == 1) NS_TRACE(0, 0);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 3; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); mbar_init(do_full(s), 1); mbar_init(do_empty(s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full(s), 1); mbar_init(p_ready(s), 8); mbar_init(dp_full(s), 1); mbar_init(ds_ready(s), 8);
      mbar_init(ds_free(s), 1);
    }
    mbar_init(dbg_bar, 1);
    mbar_init(kv_full, 1); mbar_init(dq_full, 1); mbar_init(dq_empty, 4); mbar_init(acc_done, 1); mbar_init(kvt_ready, 8);
    mbar_fence_init();
    tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v); tma_prefetch_desc(&maps.d_o);
    tma_prefetch_desc(&maps.dqacc);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tdV = tmem + 256, tdK = tmem + 320, tdQ = tmem + 384;
  const uint32_t tK = tmem + 448, tV = tmem + 480;         // bf16 copies of the resident K / V tiles: A operands of S^T / dP^T

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * kT16);
      tma_load_3d(&maps.k, kv_full, sK, h * 64, k0, b);
      tma_load_3d(&maps.v, kv_full, sV, h * 64, k0, b);
      const float* stats = p.stats + (static_cast<long long>(b) * p.H + h) * nqt * 256;
      for (int i = 0; i < nqt; ++i) {
        const int s = i % 3;
        const uint32_t ph = ((i / 3) & 1u) ^ 1u;
        if (i == 2) mbar_wait(kvt_ready, 0);                 // V has been copied to TMEM: its tile becomes Q stage 2
        mbar_wait(q_empty(s), ph);
        mbar_expect_tx(q_full(s), kT16 + 1024);
        tma_load_3d(&maps.q, q_full(s), sQ(s), h * 64, i * 128, b);
        bulk_load_1d(base + kOffStat + 1024u * s, stats + i * 256, 1024, q_full(s));
        mbar_wait(do_empty(s), ph);
        mbar_expect_tx(do_full(s), kT16);
        tma_load_3d(&maps.d_o, do_full(s), sdO(s), h * 64, i * 128, b);
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer: ONE elected thread runs the whole loop and
    // every cycle of it is on the critical path.  The loop body is unrolled over (half tile, Q/dO ring stage) so that every
    // descriptor is a uniform base plus an immediate, and each barrier is probed right before the first product that needs
    // it, so the ~90-cycle probe latencies overlap with the execution of the products issued just before.
    if (elect_one()) {
      constexpr uint32_t idS = umma_idesc_bf16(128, 64, 0, 0);  // S^T / dP^T : A = K / V in TMEM, B = 64 streamed query rows
      constexpr uint32_t idA = umma_idesc_bf16(128, 64, 0, 1);  // dV / dK    : A from TMEM, B = dO / Q half tile MN-major
      constexpr uint32_t idQ = umma_idesc_bf16(128, 64, 1, 1);  // dQ         : A = dS^T (smem, MN-major), B = K MN-major
      const uint64_t kd_mn = umma_smem_desc(sK, 16384, 1024);
      // stage s of a descriptor = stage 0 + offset (tile bytes >> 4); the start-address field never carries into the next one
      const uint64_t q_k = umma_smem_desc(sQ(0), 16, 1024), o_k = umma_smem_desc(sdO(0), 16, 1024);          // K-major views
      const uint64_t q_mn = umma_smem_desc(sQ(0), 16384, 1024), o_mn = umma_smem_desc(sdO(0), 16384, 1024);  // MN-major views
      const uint64_t ds_mn = umma_smem_desc(sdS(0), 16384, 1024);
      // HALF = j & 1 (also the TMEM buffer), QS = (j / 2) % 3: compile-time per unrolled step
      auto issue_S = [&](auto half_c, auto qs_c) {
        constexpr int HALF = decltype(half_c)::value, QS = decltype(qs_c)::value;
        constexpr long long off = (QS == 2 ? -1024 : 1024 * QS) + 512 * HALF;   // +8192 bytes: second half of the 128-row tile
        const uint64_t qd = q_k + static_cast<uint64_t>(off);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(tmem + 64u * HALF, tK + 8u * k, qd + 2u * k, idS, k > 0);
        umma_commit(s_full(HALF));
      };
      auto issue_dP = [&](auto half_c, auto qs_c) {
        constexpr int HALF = decltype(half_c)::value, QS = decltype(qs_c)::value;
        const uint64_t od = o_k + static_cast<uint64_t>(1024 * QS + 512 * HALF);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(tmem + 128u + 64u * HALF, tV + 8u * k, od + 2u * k, idS, k > 0);
        umma_commit(dp_full(HALF));
      };
      using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
      mbar_wait(kv_full, 0);
      mbar_wait(kvt_ready, 0);
      mbar_wait(q_full(0), 0);
      mbar_wait(do_full(0), 0);
      tc_fence_after();
      issue_S(I0{}, I0{}); issue_dP(I0{}, I0{}); issue_S(I1{}, I0{}); issue_dP(I1{}, I0{});
      auto step = [&](auto half_c, auto qs_c, int j) {
        constexpr int HALF = decltype(half_c)::value, QS = decltype(qs_c)::value;
        constexpr int QSN = (QS + 1) % 3;                                   // ring stage of the next query tile
        using QSN_c = std::integral_constant<int, QSN>;
        const int t = j >> 1, dss = t & 1;
        const uint32_t ph = static_cast<uint32_t>(t) & 1u;
        const uint32_t acc = j > 0 ? 1u : 0u;
        const bool more = j + 2 < nh;
        const uint32_t ph_next = ((t + 1) / 3) & 1u;
        const uint32_t tSb = tmem + 64u * HALF, tdPb = tmem + 128u + 64u * HALF;
        // ---- dV += P^T dO (K = 64 queries)
        mbar_wait(p_ready(HALF), ph);
        NS_TRACE1(0, 100 + j);
        tc_fence_after();
        {
          const uint64_t od = o_mn + static_cast<uint64_t>(1024 * QS + 512 * HALF);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(tdV, tSb + 32u * (k >> 1) + 8u * (k & 1), od + 128u * k, idA, (acc | (k > 0)) ? 1u : 0u);
          if (HALF) umma_commit(do_empty(QS));
        }
        // ---- S^T of half tile j+2 into the same TMEM buffer (behind dV by in-order execution; nobody else reads P^T)
        if (more) {
          if (!HALF) { mbar_wait(q_full(QSN), ph_next); tc_fence_after(); }
          issue_S(half_c, QSN_c{});                           // half tile j+2 belongs to query tile t+1 either way
        }
        // ---- dK += dS^T Q
        mbar_wait(ds_ready(HALF), ph);
        NS_TRACE1(0, 300 + j);
        tc_fence_after();
        {
          constexpr long long off = (QS == 2 ? -1024 : 1024 * QS) + 512 * HALF;
          const uint64_t qd = q_mn + static_cast<uint64_t>(off);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(tdK, tdPb + 32u * (k >> 1) + 8u * (k & 1), qd + 128u * k, idA, (acc | (k > 0)) ? 1u : 0u);
        }
        // ---- once per full tile: dQ = dS K
        if (HALF) {
          umma_commit(q_empty(QS));
          if (t > 0) { mbar_wait(dq_empty, (t - 1) & 1); tc_fence_after(); }
          const uint64_t dsd = ds_mn + 2048u * dss;
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_f16(tdQ, dsd + 128u * k, kd_mn + 128u * k, idQ, k > 0);
          umma_commit(dq_full);
          umma_commit(ds_free(dss));
        }
        // ---- dP^T of half tile j+2 (behind dK)
        if (more) {
          if (!HALF) { mbar_wait(do_full(QSN), ph_next); tc_fence_after(); }
          issue_dP(half_c, QSN_c{});
        }
      };
      for (int j0 = 0; j0 < nh; j0 += 6) {
        step(I0{}, I0{}, j0);
        step(I1{}, I0{}, j0 + 1);
        if (j0 + 2 >= nh) break;
        step(I0{}, I1{}, j0 + 2);
        step(I1{}, I1{}, j0 + 3);
        if (j0 + 4 >= nh) break;
        step(I0{}, I2{}, j0 + 4);
        step(I1{}, I2{}, j0 + 5);
      }
      umma_commit(acc_done);
      NS_TRACE1(0, 999);
    }
    __syncwarp();
  } else if (warp < 6) {
    // ================================================================ dQ epilogue: TMEM -> smem (fp32, swizzled) -> TMA reduce-add
    const int qd = warp & 3;
    const int row = qd * 32 + lane;                       // query row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    for (int i = 0; i < nqt; ++i) {
      mbar_wait(dq_full, i & 1);
      if (warp == 2) NS_TRACE(1, 100 + i);
      tc_fence_after();
      if (i > 0) {
        if (warp == 2) {                                  // the previous reduce has finished reading the staging tile
          if (elect_one()) bulk_wait_read0();             // (elect.sync is deterministic: the lane that committed the group)
        }
        named_bar_sync(2, 128);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tdQ + lane_addr + 32u * c, v);
        tmem_ld_wait();
        const uint32_t box = sdQ + 16384u * c + row_off;
#pragma unroll
        for (int e = 0; e < 8; ++e) st_shared_v4(box + ((static_cast<uint32_t>(e) ^ sw) << 4), v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      if (warp == 2) NS_TRACE(1, 200 + i);
      fence_proxy_async();
      named_bar_sync(2, 128);
      if (warp == 2 && !(p.debug & 1)) {
        if (elect_one()) {                                  // same lane every time: the bulk group belongs to one thread
          tma_reduce_add_3d(&maps.dqacc, sdQ, h * 64, i * 128, b);
          tma_reduce_add_3d(&maps.dqacc, sdQ + 16384u, h * 64 + 32, i * 128, b);
          bulk_commit();
        }
      }
    }
    if (warp == 2) {
      if (elect_one()) bulk_wait0();
    }
  } else {
    // ================================================================ compute warps, two groups of 8 (g = 0: warps 6..13, g = 1: 14..21).
    // Group g owns the half tiles j = g, g+2, ... (always TMEM buffer g) and runs both phases on them with P^T kept in
    // registers between the two, so nobody else ever reads the S^T / P^T buffer: S^T(j+2) only has to follow dV(j) in the MMA
    // queue.  The groups run half a period apart, so one group's exponentials (MUFU) overlap the other's dS arithmetic.
    //   phase 1: P^T = exp2(S^T log2e - lse log2e) -> bf16 over the consumed S^T columns (A operand of dV += P^T dO)
    //   phase 2: dS^T = P^T (dP^T - delta)         -> bf16 over the consumed dP^T columns (A of dK) + shared memory (A of dQ)
    const int g = (warp - 6) >> 3;
    const int qd = warp & 3;
    const int ch = ((warp - 6) >> 2) & 1;                 // which 32 of the half tile's 64 query columns
    const int row = qd * 32 + lane;                       // key row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t tSg = tmem + 64u * g + lane_addr + 32u * ch, tdPg = tmem + 128u + 64u * g + lane_addr + 32u * ch;
    if (g == 0) {
      // resident K (ch 0) / V (ch 1) row of this thread: shared (128B swizzle) -> registers -> TMEM, once per CTA.
      // With A in tensor memory the S^T / dP^T products read only their 2 KB B operand from shared memory per MMA.
      mbar_wait(kv_full, 0);
      const uint32_t src = (ch == 0 ? sK : sV) + static_cast<uint32_t>(row) * 128u;
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) ld_shared_v4(src + ((static_cast<uint32_t>(c) ^ sw) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      tmem_st32((ch == 0 ? tK : tV) + lane_addr, v);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(kvt_ready);
    }
    const bool tr_w = (warp == 6 || warp == 14);
    for (int j = g; j < nh; j += 2) {
      const int t = j >> 1, qs = t % 3, dss = t & 1;
      const uint32_t ph = static_cast<uint32_t>(t) & 1u;
      const float* st_l = stat_s + 256 * qs + 64 * g + 32 * ch;   // lse * log2e of this thread's 32 query columns
      const float* st_d = st_l + 128;                               // delta
      mbar_wait2(s_full(g), ph, q_full(qs), (t / 3) & 1u);          // q_full: this tile's statistics have landed
      if (tr_w) NS_TRACE(2 + g, 100 + j);
      tc_fence_after();
      uint32_t v[32], pk[16];
      const bool dbg_skip = (p.debug & 2) != 0;           // NS_BWF_DEBUG bit 1: no TMEM traffic from the compute warps (timing experiment)
      if (!dbg_skip) {
        tmem_ld32(tSg, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        // the table holds -lse * log2e: one packed FFMA per pair of exponents (the compute warps are issue-bound)
        const float4 l4 = *reinterpret_cast<const float4*>(st_l + 4 * q4);
        float x0, x1, x2, x3;
        f2_unpack(f2_fma(f2_pack(__uint_as_float(v[4 * q4 + 0]), __uint_as_float(v[4 * q4 + 1])), f2_splat(kL2e), f2_pack(l4.x, l4.y)), x0, x1);
        f2_unpack(f2_fma(f2_pack(__uint_as_float(v[4 * q4 + 2]), __uint_as_float(v[4 * q4 + 3])), f2_splat(kL2e), f2_pack(l4.z, l4.w)), x2, x3);
        pk[2 * q4] = pack_bf16x2(ex2f(x0), ex2f(x1));
        pk[2 * q4 + 1] = pack_bf16x2(ex2f(x2), ex2f(x3));
      }
      if (!dbg_skip) {
        tmem_st16(tSg, pk);                                 // bf16 P^T over the fp32 columns this thread has consumed
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(g));
      if (tr_w) NS_TRACE(2 + g, 200 + j);

      if (t > 1) mbar_wait2(dp_full(g), ph, ds_free(dss), ((t >> 1) - 1) & 1u);   // ds_free: dQ_{t-2} has read this staging tile
      else mbar_wait(dp_full(g), ph);
      if (tr_w) NS_TRACE(2 + g, 300 + j);
      tc_fence_after();
      if (!dbg_skip) {
        tmem_ld32(tdPg, v);
        tmem_ld_wait();
      }
      uint32_t dk[16];
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        // dS = P (dP - delta) on bf16 pairs: the table holds -delta, (dP - delta) is one packed add, rounded to bf16 and
        // multiplied with the bf16 P the dV product uses (3 instructions per pair instead of 7 with fp32 unpacking)
        const float4 d4 = *reinterpret_cast<const float4*>(st_d + 4 * q4);
        float e0, e1, e2, e3;
        f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * q4 + 0]), __uint_as_float(v[4 * q4 + 1])), f2_pack(d4.x, d4.y)), e0, e1);
        f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * q4 + 2]), __uint_as_float(v[4 * q4 + 3])), f2_pack(d4.z, d4.w)), e2, e3);
        dk[2 * q4] = bf16x2_mul(pk[2 * q4], pack_bf16x2(e0, e1));
        dk[2 * q4 + 1] = bf16x2_mul(pk[2 * q4 + 1], pack_bf16x2(e2, e3));
      }
      if (!dbg_skip) tmem_st16(tdPg, dk);                   // A operand of dK += dS^T Q
      const uint32_t ds_row = sdS(dss) + 16384u * g + static_cast<uint32_t>(row) * 128u;
#pragma unroll
      for (int e = 0; e < 4; ++e)                           // dS^T row (this key) x 8 queries per 16-byte chunk, 128B swizzle
        st_shared_v4(ds_row + ((static_cast<uint32_t>(4 * ch + e) ^ sw) << 4), dk[4 * e], dk[4 * e + 1], dk[4 * e + 2], dk[4 * e + 3]);
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready(g));
      if (tr_w) NS_TRACE(2 + g, 400 + j);
    }
    if (g == 0) {
      // ---- dV (ch 0) / dK (ch 1) of this key tile
      mbar_wait(acc_done, 0);
      tc_fence_after();
      const int ki = k0 + row;
      const uint32_t tacc = ch == 0 ? tdV : tdK;
      __nv_bfloat16* out = ch == 0 ? p.dv + b * p.dv_bs + static_cast<long long>(ki) * p.dv_rs + h * 64
                                   : p.dk + b * p.dk_bs + static_cast<long long>(ki) * p.dk_rs + h * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tacc + lane_addr + 32u * c, v);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------------- prep / finish kernels
// One warp per (b, q) row: delta[b,h,q] = sum_d dO * O for every head (16-byte loads, 4 lanes per head), written together
// with lse * log2e in tile-major order [b][h][q / 128][{-lse2, -delta}][q % 128]; rows beyond Lq get (-inf, 0).
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(int B, int H, int Lq, int nqt, long long o_bs, long long o_rs, const __nv_bfloat16* __restrict__ o,
                     const __nv_bfloat16* __restrict__ d_o, const float* __restrict__ lse, float* __restrict__ delta,
                     float* __restrict__ stats) {
  const long long rowid = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int Lp = nqt * 128;
  if (rowid >= static_cast<long long>(B) * Lp) return;
  const int b = static_cast<int>(rowid / Lp), q = static_cast<int>(rowid % Lp);
  const bool valid = q < Lq;
  for (int h0 = 0; h0 < H; h0 += 8) {
    const int hh = h0 + (lane >> 2);
    float a = 0.f;
    if (valid && hh < H) {
      const long long off = b * o_bs + static_cast<long long>(q) * o_rs + h0 * 64 + lane * 16;
      const uint4* po = reinterpret_cast<const uint4*>(o + off);
      const uint4* pd = reinterpret_cast<const uint4*>(d_o + off);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const uint4 x = __ldg(po + e), y = __ldg(pd + e);
        float2 f, g;
        f = unpack_bf16x2(x.x); g = unpack_bf16x2(y.x); a = fmaf(f.x, g.x, a); a = fmaf(f.y, g.y, a);
        f = unpack_bf16x2(x.y); g = unpack_bf16x2(y.y); a = fmaf(f.x, g.x, a); a = fmaf(f.y, g.y, a);
        f = unpack_bf16x2(x.z); g = unpack_bf16x2(y.z); a = fmaf(f.x, g.x, a); a = fmaf(f.y, g.y, a);
        f = unpack_bf16x2(x.w); g = unpack_bf16x2(y.w); a = fmaf(f.x, g.x, a); a = fmaf(f.y, g.y, a);
      }
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    if ((lane & 3) == 0 && hh < H) {
      const long long sidx = ((static_cast<long long>(b) * H + hh) * nqt + (q >> 7)) * 256 + (q & 127);
      if (valid) {
        const long long li = (static_cast<long long>(b) * H + hh) * Lq + q;
        stats[sidx] = -lse[li] * kL2e;          // both negated: the main kernel adds them (packed FFMA / FADD)
        stats[sidx + 128] = -a;
        delta[li] = a;
      } else {
        stats[sidx] = -INFINITY;                // exp2(-inf) = 0 for the padding rows
        stats[sidx + 128] = 0.f;
      }
    }
  }
}

// dq (bf16, caller strides) = fp32 accumulation buffer (B, Lq, H*64)
__global__ void __launch_bounds__(256)
attn_bwd_dq_convert_kernel(long long rows, int Lq, int W, long long q_bs, long long q_rs, const float* __restrict__ acc,
                           __nv_bfloat16* __restrict__ dq) {
  const int per_row = W / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= rows * per_row) return;
  const long long r = idx / per_row;
  const int c = static_cast<int>(idx % per_row) * 8;
  const float4* src = reinterpret_cast<const float4*>(acc + r * W + c);
  const float4 a = __ldg(src), bq = __ldg(src + 1);
  uint4 u;
  u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(bq.x, bq.y); u.w = pack_bf16x2(bq.z, bq.w);
  const long long b = r / Lq, q = r % Lq;
  *reinterpret_cast<uint4*>(dq + b * q_bs + q * q_rs + c) = u;
}

// ---------------------------------------------------------------------------------------------------- host side
static long long* g_attn_trace = nullptr;
void set_attn_trace(long long* p) { g_attn_trace = p; }
long long* get_attn_trace() { return g_attn_trace; }
static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int head_map128(CUtensorMap* m, const void* base, int H, int L, int B, long long bs, long long rs) {
  uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)L, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)rs * 2, (uint64_t)bs * 2};
  uint32_t box[3] = {64, 128, 1};
  return make_map(m, base, 3, dims, str, box);
}

static bool fused_eligible(const ns_attn_shape& s) {
  return s.Dh == 64 && !s.causal && s.Lq >= 1 && s.Lk >= 1 && s.H <= 65535 && s.B <= 65535 && s.q_rs % 8 == 0 && s.k_rs % 8 == 0 &&
         s.v_rs % 8 == 0 && s.o_rs % 8 == 0 && s.q_bs % 8 == 0 && s.k_bs % 8 == 0 && s.v_bs % 8 == 0 && s.o_bs % 8 == 0;
}

// workspace: fp32 dQ accumulator (B, Lq, H*64) followed by the tile-major statistics (B, H, nqt, 256)
size_t attention_bwd_fused_ws(const ns_attn_shape& s) {
  if (!fused_eligible(s)) return 0;
  const size_t nqt = (s.Lq + 127) / 128;
  const size_t acc = ((static_cast<size_t>(s.B) * s.Lq * s.H * 64 * 4) + 1023) / 1024 * 1024;
  return acc + static_cast<size_t>(s.B) * s.H * nqt * 256 * 4;
}

int attention_bwd_fused(const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                        const float* lse, float* delta, void* dq, void* dk, void* dv, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!fused_eligible(s) || !al16(q) || !al16(k) || !al16(v) || !al16(o) || !al16(d_o) || !al16(dq) || !al16(dk) || !al16(dv) ||
      !ws || (reinterpret_cast<uintptr_t>(ws) & 1023) != 0 || ws_bytes < attention_bwd_fused_ws(s))
    return NS_ERR_UNSUPPORTED;
  static bool attr_done = false;
  if (!attr_done) {
    NS_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBfSmem));
    attr_done = true;
  }
  const int nqt = (s.Lq + 127) / 128;
  const int W = s.H * 64;
  const size_t acc_bytes = ((static_cast<size_t>(s.B) * s.Lq * W * 4) + 1023) / 1024 * 1024;
  float* acc = reinterpret_cast<float*>(ws);
  float* stats = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + acc_bytes);
  NS_CUDA(cudaMemsetAsync(acc, 0, static_cast<size_t>(s.B) * s.Lq * W * 4, st));
  {
    const long long rows = static_cast<long long>(s.B) * nqt * 128;
    attn_bwd_prep_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(
        s.B, s.H, s.Lq, nqt, s.o_bs, s.o_rs, reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(d_o),
        lse, delta, stats);
    NS_LAUNCH_CHECK();
  }
  BwfMaps maps;
  int r;
  if ((r = head_map128(&maps.q, q, s.H, s.Lq, s.B, s.q_bs, s.q_rs))) return r;
  if ((r = head_map128(&maps.k, k, s.H, s.Lk, s.B, s.k_bs, s.k_rs))) return r;
  if ((r = head_map128(&maps.v, v, s.H, s.Lk, s.B, s.v_bs, s.v_rs))) return r;
  if ((r = head_map128(&maps.d_o, d_o, s.H, s.Lq, s.B, s.o_bs, s.o_rs))) return r;
  {
    uint64_t dims[3] = {(uint64_t)W, (uint64_t)s.Lq, (uint64_t)s.B};
    uint64_t str[2] = {(uint64_t)W * 4, (uint64_t)W * 4 * (uint64_t)s.Lq};
    uint32_t box[3] = {32, 128, 1};
    if ((r = make_map_f32(&maps.dqacc, acc, 3, dims, str, box))) return r;
  }
  BwfProg prog{s.B, s.H, s.Lq, s.Lk, nqt, stats, s.k_bs, s.k_rs, s.v_bs, s.v_rs,
               reinterpret_cast<__nv_bfloat16*>(dk), reinterpret_cast<__nv_bfloat16*>(dv), 0, g_attn_trace};
  if (const char* e = getenv("NS_BWF_DEBUG")) prog.debug = atoi(e);
  dim3 grid((s.Lk + 127) / 128, s.H, s.B);
  attn_bwd_fused_kernel<<<grid, kBfThreads, kBfSmem, st>>>(maps, prog);
  NS_LAUNCH_CHECK();
  {
    const long long rows = static_cast<long long>(s.B) * s.Lq;
    const long long n = rows * (W / 8);
    attn_bwd_dq_convert_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(rows, s.Lq, W, s.q_bs, s.q_rs, acc,
                                                                                      reinterpret_cast<__nv_bfloat16*>(dq));
    NS_LAUNCH_CHECK();
  }
  count(C_ATTN_TC, 1);
  count(C_OTHER, 2);
  return NS_OK;
}
}  // namespace ns
