// B-side backward of a LoRA branch in ONE pass over the output gradient (PEFT lora.Linear under autograd, finetune.py:194-212:
// y += t B^T with t = alpha' (x . keep) A^T).  Both products contract the same matrix dy:
//     dt = alpha' dy B        (M, r)   -- contraction over the N output columns: dy is the K-major A operand
//     dB += dy^T t            (N, r)   -- contraction over the M rows:          dy is the MN-major A operand
// ns_gemm_nt (rank-r tile) + ns_gemm_tn streamed dy from HBM twice (31 + 27 us for M = 96000, N = 512; 62 + 74 us for the
// stacked q/k/v adapters).  Here a CTA owns 128-row slabs of dy and streams each slab once through a TMA ring of
// [128 rows][128 columns] chunks; the SAME shared-memory bytes feed two tcgen05 products per chunk:
//     dB_c (128 x r) += chunk^T (MN-major view, K = 128 rows)  * t slab (MN-major, N = r)        accumulates over the CTA's slabs
//     dt   (128 x r) += chunk   (K-major view,  K = 128 cols)  * B^T chunk (K-major, resident)   accumulates over the slab's chunks
// TMEM: N/128 dB accumulators of r columns + two dt accumulators (slab parity) -> 14 chunks (N <= 1792) per CTA.  Warp roles: warp 0 TMA
// producer, warp 1 MMA issuer (one elected thread), warp 2 TMEM allocator, warps 4-7 epilogue (dt slab -> bf16 -> global
// while the next slab streams; the dB accumulators leave once, as 16-byte fp32 reductions into the flat gradient buffer).
// Groups (q, k, v stacked along the columns of dy / t / dt and the rows of B^T / dB) are independent work items of one launch.
// Wider N (fc1's 2048 columns = 16 dB accumulators = all 512 TMEM columns) is cut into `parts` column ranges on different CTAs:
// each part holds its own dB rows and a PARTIAL dt; the parts of a (slab, 32-row quarter) meet through a caller-owned workspace --
// every part stores its fp32 partial, fences, and takes a ticket; whoever draws the last ticket adds the other partials (L2
// hits) to its registers and writes the bf16 rows.  Nobody waits for anybody, so the order in which CTAs run does not matter;
// the tickets return to zero (the caller zeroes them once).
#include "ns_common.cuh"
#include "ns_sm100.cuh"
#include "ns_gemm.cuh"

#include <stdlib.h>

namespace ns {
using namespace sm100;

struct LbMaps {
  CUtensorMap x, t, bt;
};
struct LbProg {
  long long M;
  int N, r, groups, nchunk;      // nchunk: 128-column chunks per part
  int parts;                     // column ranges of one group on different CTAs (1: no workspace)
  int slabs, nsplit, stages;
  int dbg;                       // NS_LB_DEBUG (timing experiments, wrong results): 1 = no dB products, 2 = no dt products
  int* tickets;                  // [groups][slabs][4]           (parts > 1)
  float* partial;                // [groups][parts][slabs * 128][r]
  uint32_t bt_bytes, tmem_cols;
  __nv_bfloat16* dt;
  long long lddt;
  float* dB;
  float alpha_dt[4], alpha_db[4];
};
constexpr int kLbThreads = 256;
constexpr int kLbMaxStages = 5;
constexpr uint32_t kLbChunkBytes = 2 * 128 * 64 * 2;   // two [128 rows][64 columns] boxes
constexpr uint32_t kLbTBytes = 128 * 64 * 2;           // t slab box: 64 columns wide (r = 32 of them used by the MMA)
constexpr int kLbSmemMax = 232448;

__global__ void __launch_bounds__(kLbThreads, 1) lora_bwd_b_kernel(const __grid_constant__ LbMaps maps, const __grid_constant__ LbProg p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBt = smem_base;
  auto sT = [&](int b) { return smem_base + p.bt_bytes + static_cast<uint32_t>(b) * kLbTBytes; };
  auto sX = [&](int s) { return smem_base + p.bt_bytes + 2u * kLbTBytes + static_cast<uint32_t>(s) * kLbChunkBytes; };
  const uint32_t bar_base = sX(p.stages);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kLbMaxStages + s); };
  auto t_full = [&](int b) { return bar_base + 8u * (2 * kLbMaxStages + b); };
  auto t_empty = [&](int b) { return bar_base + 8u * (2 * kLbMaxStages + 2 + b); };
  auto dt_full = [&](int b) { return bar_base + 8u * (2 * kLbMaxStages + 4 + b); };
  auto dt_empty = [&](int b) { return bar_base + 8u * (2 * kLbMaxStages + 6 + b); };
  const uint32_t bt_full = bar_base + 8u * (2 * kLbMaxStages + 8);
  const uint32_t db_full = bar_base + 8u * (2 * kLbMaxStages + 9);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kLbMaxStages + 10);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x % p.nsplit, part = (blockIdx.x / p.nsplit) % p.parts, grp = blockIdx.x / (p.nsplit * p.parts);
  const int pcol0 = part * p.nchunk * 128;                         // first column of this part inside the group
  // balanced partition of the slabs: every CTA gets floor or ceil of slabs / nsplit (750 slabs over 148 CTAs: 5 or 6 each on
  // all SMs, instead of 6 each on 125)
  const int sl0 = static_cast<int>(static_cast<long long>(split) * p.slabs / p.nsplit);
  const int sl1 = static_cast<int>(static_cast<long long>(split + 1) * p.slabs / p.nsplit);
  const int S = p.stages;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(t_full(b), 1);
      mbar_init(t_empty(b), 1);
      mbar_init(dt_full(b), 1);
      mbar_init(dt_empty(b), 4);
    }
    mbar_init(bt_full, 1);
    mbar_init(db_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.t);
    tma_prefetch_desc(&maps.bt);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_dt = tmem_base + static_cast<uint32_t>(p.nchunk * p.r);

  if (warp == 0) {
    if (elect_one() && sl0 < sl1) {
      mbar_expect_tx(bt_full, p.bt_bytes);
      for (int b = 0; b < 2 * p.nchunk; ++b) tma_load_4d(&maps.bt, bt_full, sBt + 4096u * b, pcol0 + 64 * b, 0, grp * p.r, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int sl = sl0; sl < sl1; ++sl) {
        const int b = (sl - sl0) & 1;
        const uint32_t u = static_cast<uint32_t>((sl - sl0) >> 1);
        mbar_wait(t_empty(b), (u & 1u) ^ 1u);
        mbar_expect_tx(t_full(b), kLbTBytes);
        tma_load_4d(&maps.t, t_full(b), sT(b), grp * p.r, 0, sl * 128, 0);
        for (int c = 0; c < p.nchunk; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), kLbChunkBytes);
          const int col = grp * p.N + pcol0 + c * 128;
          tma_load_4d(&maps.x, full_bar(stage), sX(stage), col, 0, sl * 128, 0);
          tma_load_4d(&maps.x, full_bar(stage), sX(stage) + 16384u, col + 64, 0, sl * 128, 0);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (sl0 < sl1) {
      constexpr uint32_t idB = umma_idesc_bf16(128, 32, 1, 1);   // dB chunk: A = dy chunk^T, B = t slab, both MN-major
      constexpr uint32_t idT = umma_idesc_bf16(128, 32, 0, 0);   // dt slab : A = dy chunk,   B = B^T chunk, both K-major
      mbar_wait(bt_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int sl = sl0; sl < sl1; ++sl) {
        const int b = (sl - sl0) & 1;
        const uint32_t u = static_cast<uint32_t>((sl - sl0) >> 1);
        mbar_wait(t_full(b), u & 1u);
        mbar_wait(dt_empty(b), (u & 1u) ^ 1u);
        const uint32_t acc_db = sl > sl0 ? 1u : 0u;
        for (int c = 0; c < p.nchunk; ++c) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = sX(stage);
            // MN-major views: 64-column groups 16 KB apart (LBO), 8-row groups 1 KB apart (SBO); 16 rows per MMA = +2048 B
            const uint64_t a_mn = umma_smem_desc(sa, 16384, 1024), t_mn = umma_smem_desc(sT(b), 16384, 1024);
            if (!(p.dbg & 1)) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_f16(tmem_base + static_cast<uint32_t>(c * p.r), a_mn + 128u * k, t_mn + 128u * k, idB, (acc_db | (k > 0)) ? 1u : 0u);
            }
            // K-major views: one box = 64 contraction columns, 16 per MMA = +32 B inside the swizzled 128-byte row
#pragma unroll
            for (int box = 0; box < ((p.dbg & 2) ? 0 : 2); ++box) {
              const uint64_t a_k = umma_smem_desc(sa + 16384u * box, 16, 1024);
              const uint64_t b_k = umma_smem_desc(sBt + 4096u * static_cast<uint32_t>(2 * c + box), 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_dt + static_cast<uint32_t>(b * p.r), a_k + 2u * k, b_k + 2u * k, idT, (c > 0 || box > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));
            if (c == p.nchunk - 1) {
              umma_commit(dt_full(b));
              umma_commit(t_empty(b));
              if (sl == sl1 - 1) umma_commit(db_full);
            }
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const float a_dt = p.alpha_dt[grp], a_db = p.alpha_db[grp];
    for (int sl = sl0; sl < sl1; ++sl) {
      const int b = (sl - sl0) & 1;
      const uint32_t u = static_cast<uint32_t>((sl - sl0) >> 1);
      mbar_wait(dt_full(b), u & 1u);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_dt + lane_base + static_cast<uint32_t>(b * p.r), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dt_empty(b));
      const long long row = static_cast<long long>(sl) * 128 + q * 32 + lane;
      if (p.parts > 1) {
        // meet the other column parts of these 32 rows (see the file header)
        const long long prow = static_cast<long long>(p.slabs) * 128;
        float* mine = p.partial + ((static_cast<long long>(grp) * p.parts + part) * prow + row) * p.r;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          __stcg(reinterpret_cast<float4*>(mine + j), make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
        __threadfence();
        __syncwarp();
        int* tk = p.tickets + (static_cast<long long>(grp) * p.slabs + sl) * 4 + q;
        int old = 0;
        if (lane == 0) old = atomicAdd(tk, 1);
        old = __shfl_sync(0xFFFFFFFFu, old, 0);
        if (old != p.parts - 1) continue;                          // not the last part of these rows: done
        __threadfence();
        if (lane == 0) *tk = 0;
        for (int o = 0; o < p.parts; ++o) {
          if (o == part) continue;
          const float* other = p.partial + ((static_cast<long long>(grp) * p.parts + o) * prow + row) * p.r;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 w = __ldcg(reinterpret_cast<const float4*>(other + j));
            v[j] = __float_as_uint(__uint_as_float(v[j]) + w.x); v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + w.y);
            v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + w.z); v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + w.w);
          }
        }
      }
      if (row < p.M) {
        uint4* dst = reinterpret_cast<uint4*>(p.dt + row * p.lddt + grp * p.r);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(a_dt * __uint_as_float(v[j + 2 * e]), a_dt * __uint_as_float(v[j + 2 * e + 1]));
            w[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          dst[j >> 3] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    if (sl0 < sl1) {
      mbar_wait(db_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < p.nchunk; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(c * p.r), v);
        tmem_ld_wait();
        float* g0 = p.dB + (static_cast<long long>(grp) * p.N + pcol0 + c * 128 + q * 32 + lane) * p.r;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g0 + j), "f"(a_db * __uint_as_float(v[j])),
                       "f"(a_db * __uint_as_float(v[j + 1])), "f"(a_db * __uint_as_float(v[j + 2])), "f"(a_db * __uint_as_float(v[j + 3]))
                       : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

// column parts of one group: the fewest whose N/128/parts dB accumulators + two dt accumulators fit 512 TMEM columns
static int lb_parts(int N, int r) {
  const int chunks = N / 128, max_chunks = (512 - 2 * r) / r;
  int parts = 1;
  while (chunks % parts != 0 || chunks / parts > max_chunks) ++parts;
  return parts;
}
static bool lb_shape_ok(long long M, int N, int r, int groups) {
  return r == 32 && N >= 128 && N % 128 == 0 && groups >= 1 && groups <= 4 && M > 0 && M <= 0x7fffffffLL && lb_parts(N, r) <= 4;
}
long long lora_bwd_b_workspace_bytes(long long M, int N, int r, int groups) {
  if (!lb_shape_ok(M, N, r, groups)) return -1;
  const int parts = lb_parts(N, r);
  if (parts == 1) return 0;
  const long long slabs = (M + 127) / 128;
  const long long tickets = (groups * slabs * 4 * 4 + 255) / 256 * 256;
  return tickets + static_cast<long long>(groups) * parts * slabs * 128 * r * 4;
}

int lora_bwd_b_fast(long long M, int N, int r, int groups, const void* dy, long long lddy, const void* Bt, long long ldbt, const void* t,
                    long long ldt, void* dt, long long lddt, float* dB, const float* alpha_dt, const float* alpha_db, void* workspace,
                    long long workspace_bytes, cudaStream_t st) {
  if (!lb_shape_ok(M, N, r, groups)) return NS_ERR_UNSUPPORTED;
  if (!al16(dy) || !al16(Bt) || !al16(t) || !al16(dt) || !al16(dB) || lddy % 8 || ldbt % 8 || ldt % 8 || lddt % 8) return NS_ERR_UNSUPPORTED;
  LbProg p;
  memset(&p, 0, sizeof(p));
  static const int dbg = getenv("NS_LB_DEBUG") ? atoi(getenv("NS_LB_DEBUG")) : 0;
  p.dbg = dbg;
  p.M = M; p.N = N; p.r = r; p.groups = groups;
  p.parts = lb_parts(N, r);
  p.nchunk = N / 128 / p.parts;
  p.slabs = static_cast<int>((M + 127) / 128);
  if (p.parts > 1) {
    const long long need = lora_bwd_b_workspace_bytes(M, N, r, groups);
    if (need < 0) return NS_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < need || !al16(workspace)) {
      set_error("ns_lora_bwd_b: N = %d needs a %lld-byte workspace (ns_lora_bwd_b_workspace_bytes), 16-byte aligned, tickets zeroed once", N, need);
      return NS_ERR_ARG;
    }
    const long long tickets = (static_cast<long long>(groups) * p.slabs * 4 * 4 + 255) / 256 * 256;
    p.tickets = static_cast<int*>(workspace);
    p.partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + tickets);
  }
  const int cols = p.nchunk * r + 2 * r;
  p.tmem_cols = 32;
  while (static_cast<int>(p.tmem_cols) < cols) p.tmem_cols *= 2;
  p.bt_bytes = static_cast<uint32_t>(2 * p.nchunk) * 4096u;
  const int room = kLbSmemMax - 1024 - 256 - static_cast<int>(p.bt_bytes) - 2 * static_cast<int>(kLbTBytes);
  p.stages = room / static_cast<int>(kLbChunkBytes);
  if (p.stages > kLbMaxStages) p.stages = kLbMaxStages;
  if (p.stages < 2) return NS_ERR_UNSUPPORTED;
  int nsplit = sm_count() / (groups * p.parts);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > p.slabs) nsplit = p.slabs;
  p.nsplit = nsplit;
  p.dt = static_cast<__nv_bfloat16*>(dt); p.lddt = lddt; p.dB = dB;
  for (int g = 0; g < groups; ++g) { p.alpha_dt[g] = alpha_dt[g]; p.alpha_db[g] = alpha_db[g]; }
  LbMaps maps;
  const uint64_t dx[4] = {(uint64_t)N * groups, 1, (uint64_t)M, 1};
  const uint64_t sx[3] = {(uint64_t)lddy * 2, (uint64_t)lddy * 2, (uint64_t)lddy * 2 * (uint64_t)M};
  const uint32_t box[4] = {64, 1, 128, 1};
  int rc = make_map(&maps.x, dy, 4, dx, sx, box);
  if (rc) return rc;
  const uint64_t dtm[4] = {(uint64_t)r * groups, 1, (uint64_t)M, 1};
  const uint64_t stm[3] = {(uint64_t)ldt * 2, (uint64_t)ldt * 2, (uint64_t)ldt * 2 * (uint64_t)M};
  rc = make_map(&maps.t, t, 4, dtm, stm, box);
  if (rc) return rc;
  const uint64_t db[4] = {(uint64_t)N, 1, (uint64_t)r * groups, 1};
  const uint64_t sb[3] = {(uint64_t)ldbt * 2, (uint64_t)ldbt * 2, (uint64_t)ldbt * 2 * (uint64_t)r * groups};
  const uint32_t boxb[4] = {64, 1, 32, 1};
  rc = make_map(&maps.bt, Bt, 4, db, sb, boxb);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    NS_CUDA(cudaFuncSetAttribute(lora_bwd_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLbSmemMax));
    attr_done = true;
  }
  const int smem = 1024 + static_cast<int>(p.bt_bytes) + 2 * static_cast<int>(kLbTBytes) + p.stages * static_cast<int>(kLbChunkBytes) + 256;
  lora_bwd_b_kernel<<<groups * p.parts * p.nsplit, kLbThreads, smem, st>>>(maps, p);
  NS_LAUNCH_CHECK();
  count(C_WGRAD_TC);
  return NS_OK;
}

}  // namespace ns
