// tcgen05 / TMEM / TMA GEMM family for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
//   gemm_nt_kernel<BN> : persistent, warp-specialised  D = epi( sum_seg A_seg * W_seg^T )   (both operands K-major)
//        segments = main product, optional LoRA product (K2 = rank), or the 3 taps of a k=3 convolution (implicit GEMM:
//        the A tile of tap k is the same activation tensor at a shifted / parity-selected row coordinate; TMA zero-fills
//        the out-of-range rows, which is exactly the conv's zero padding).
//   gemm_tn_kernel     : split-K weight-gradient GEMM  G += alpha * X^T Y  (both operands MN-major, fp32 atomics).
//
// Warp roles (gemm_nt): warp0 = TMA producer, warp1 = MMA issuer (single thread), warp2 = TMEM allocator,
// warps 4..11 = epilogue (TMEM -> registers -> fused bias/scale/GELU/residual -> global).  Two TMEM accumulator stages
// let the epilogue of tile i overlap the main loop of tile i+1.
#include "ns_common.cuh"
#include "ns_sm100.cuh"
#include "ns_gemm.cuh"
#include "ns_dropout.cuh"

#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>
#include <utility>
#include <type_traits>

namespace ns {
using namespace sm100;

long long* get_attn_trace();   // ns_attention_bwd_fused.cu (developer aid)

// ------------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = []() -> PFN_encodeTiled {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<PFN_encodeTiled>(f);
  }();
  return fn;
}

// bf16 tensor map, 128B swizzle, zero OOB fill.  dims[0] is the contiguous dim; strides in BYTES for dims 1..rank-1.
static int make_map_t(CUtensorMap* out, CUtensorMapDataType dt, const void* ptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_b, const uint32_t* box);
int make_map(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_b,
             const uint32_t* box) {
  return make_map_t(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ptr, rank, dims, strides_b, box);
}
int make_map_f32(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_b,
                 const uint32_t* box) {
  return make_map_t(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ptr, rank, dims, strides_b, box);
}
static int make_map_t(CUtensorMap* out, CUtensorMapDataType dt, const void* ptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_b, const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return NS_ERR_CUDA;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gs[i - 1] = strides_b[i - 1];
      if (gs[i - 1] % 16 != 0) {
        set_error("tensor map stride %d = %llu bytes is not a multiple of 16", i, (unsigned long long)gs[i - 1]);
        return NS_ERR_ARG;
      }
    }
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) {
    set_error("tensor map base pointer not 16-byte aligned");
    return NS_ERR_ARG;
  }
  CUresult r = enc(out, dt, rank, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r, rank,
              (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
              (unsigned long long)(rank > 2 ? gd[2] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0);
    return NS_ERR_CUDA;
  }
  return NS_OK;
}

// ------------------------------------------------------------------------------------------------ kernel parameters
struct Seg {
  int kblocks;      // 64-wide K blocks in this segment
  int last_ksteps;  // UMMA K=16 steps issued in the last block (1..4)
  int a_map, b_map; // which tensor map
  int a_off;        // row offset added to the tile's first row (conv taps)
  int a_par;        // parity coordinate (stride-2 convs view the input as (.., T/2, 2, C))
  int b_tap;        // 3rd coordinate of the weight map
  int a_ngrp;       // if > 0: A's k coordinate is advanced by (n0 / a_ngrp) * a_kstep (stacked LoRA adapters)
  int a_kstep;
};

// n / d for 0 <= n < 2^31 as one multiply-high and a shift (the tile decode ran five integer divisions per tile and role)
struct FastDiv {
  uint32_t mul, shr, d;
  __device__ __forceinline__ int div(int n) const { return d == 1u ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul) >> shr); }
};
static FastDiv make_fastdiv(long long dd) {
  FastDiv f;
  const uint32_t d = dd < 1 ? 1u : static_cast<uint32_t>(dd);
  f.d = d; f.mul = 0; f.shr = 0;
  if (d > 1) {
    uint32_t lg = 0;
    while ((1ull << lg) < d) ++lg;
    const uint32_t pw = 31 + lg;
    f.mul = static_cast<uint32_t>(((1ull << pw) + d - 1) / d);
    f.shr = pw - 32;
  }
  return f;
}

struct TileProg {
  int nseg;
  Seg seg[4];
  int batches, tout, tiles_per_batch, n_tiles, N;
  long long out_bs, out_rs, out_off, ldd;
  void* D;
  EpiDev epi;
  int vec_out, vec_aux, vec_res, vec_bias;
  int tma_out;      // 1: bf16 output (and aux) tiles leave through shared memory + TMA store (BN >= 128 only)
  int tma_in;       // with tma_out: 1 = aux_in (NS_ACT_DGELU), 2 = residual tiles arrive through TMA + shared memory
  int in_batched;   // tma_in: the input tensor has a batch coordinate (0: one [rows][N] table shared by all batches)
  int m_fast;       // tile raster: 1 = consecutive tiles walk M first (small M, large N: the weight tile is the one to reuse)
  long long* trace; // developer aid (ns_debug_attn_trace buffer): CTAs 0/1 record (tag, clock64) of producer / MMA events
  int stages;       // operand ring depth (what fits beside the staging tiles)
  int staging_tiles;  // 0, 2 (one output tile per column half) or 4 (+ one aux-output or input tile per half)
  int epi_mode;     // which compiled copy of the epilogue loop runs (see gemm_nt_kernel)
  FastDiv fd_units, fd_ntiles, fd_tpb;   // m_units (row tiles, or row-tile pairs with CG = 2), n_tiles, tiles_per_batch
  // AM (A-operand mask, the LoRA down product under branch dropout): plane of adapter g = n_tile starts at am_bits + g * am_gstride
  const uint32_t* am_bits;
  long long am_ld, am_gstride;
  // am_seed != NULL: the mask stage draws the plane words (drop_plane_word) and stores them to am_bits instead of loading them
  const uint32_t* am_seed;
  uint32_t am_salts[4], am_thr;
};

struct Maps {
  CUtensorMap a[2];
  CUtensorMap b[2];
  CUtensorMap d;    // output, box {64 columns, 128 rows, 1}: used when tma_out
  CUtensorMap aux;  // pre-activation output (NS_ACT_GELU with aux_out), same box
  CUtensorMap in;   // epilogue input tile (aux_in or residual), same box
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kABytes = kBM * kBK * 2;
constexpr int kNtThreads = 384;

// CG = 2: the kernel runs as CTA pairs (clusters of 2, tcgen05 cta_group::2).  A pair owns a 256-row x BN tile: each CTA
// stages its own 128 A rows and HALF of the weight tile (BN/2 rows), the leader issues M = 256 MMAs that read both halves,
// and each CTA's TMEM receives its 128 accumulator rows.  Per MMA every SM then moves 8 KB through shared memory instead of
// 12 KB and fetches a third less from L2 -- the two limits of the single-CTA kernel on the encoder shapes.
// PM = 1 ("product mask"): the LAST segment (the LoRA product of an input-gradient GEMM) accumulates into its own TMEM tile
// and the epilogue adds it only where the branch's dropout kept the element (ns_epilogue::drop_bits):
//   D = act( A W^T + keep . (A2 W2^T) ).  Two stages x (main + masked product) x BN columns = all 512 TMEM columns at BN = 128.
template <int BN, int CG = 1, int PM = 0> struct NtCfg {
  static constexpr int kBBytes = (BN / CG) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kMaxSmem = 232448;             // the 227 KB opt-in limit (the runtime's 1 KB per CTA is outside it)
  static constexpr int kTailBytes = 256 + 2048;       // barriers + the epilogue's bias rows (2 column halves x 2 tile parities)
  static constexpr int kTmemCols = PM ? 4 * BN : ((2 * BN < 32) ? 32 : 2 * BN);
  static_assert(!PM || BN == 128, "the masked second product is built for 128-wide tiles");
  // staging for the TMA epilogue: [128 rows][64 columns] bf16 tiles (128B swizzle), see TileProg::staging_tiles
  static constexpr int kStagingTile = kBM * 128;
  static int stages_for(int staging_tiles) {
    const int s = (kMaxSmem - kTailBytes - staging_tiles * kStagingTile) / kStageBytes;
    return s > 8 ? 8 : s;
  }
  static int smem_bytes(int stages, int staging_tiles) { return stages * kStageBytes + staging_tiles * kStagingTile + kTailBytes; }
};

// ------------------------------------------------------------------------------------------------ epilogue helpers
__device__ __forceinline__ void load32_bf16(const __nv_bfloat16* p, bool vec, int ncols, float (&o)[32]) {
  if (vec) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u = __ldg(q + i);
      float2 f;
      f = unpack_bf16x2(u.x); o[8 * i + 0] = f.x; o[8 * i + 1] = f.y;
      f = unpack_bf16x2(u.y); o[8 * i + 2] = f.x; o[8 * i + 3] = f.y;
      f = unpack_bf16x2(u.z); o[8 * i + 4] = f.x; o[8 * i + 5] = f.y;
      f = unpack_bf16x2(u.w); o[8 * i + 6] = f.x; o[8 * i + 7] = f.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = (j < ncols) ? __bfloat162float(p[j]) : 0.0f;
  }
}
__device__ __forceinline__ void store32_bf16(__nv_bfloat16* p, bool vec, int ncols, const float (&x)[32]) {
  if (vec) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16x2(x[8 * i + 0], x[8 * i + 1]);
      u.y = pack_bf16x2(x[8 * i + 2], x[8 * i + 3]);
      u.z = pack_bf16x2(x[8 * i + 4], x[8 * i + 5]);
      u.w = pack_bf16x2(x[8 * i + 6], x[8 * i + 7]);
      q[i] = u;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) p[j] = __float2bfloat16_rn(x[j]);
  }
}
__device__ __forceinline__ void store32_f32(float* p, bool vec, int ncols, const float (&x)[32]) {
  if (vec) {
    float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) p[j] = x[j];
  }
}

// ------------------------------------------------------------------------------------------------ NT kernel
// Epilogue timeline of warp 4 of CTA 0 (tools/gemm_epi_trace.py): compiled in only with -DNS_GEMM_EPI_TRACE, the probes
// cost ~5 % on the short-K shapes even when no trace buffer is set.
#ifdef NS_GEMM_EPI_TRACE
#define NS_EPI_TRACE(tag) do { if (ew == 0 && lane == 0 && blockIdx.x == 0) trace(3, tag); } while (0)
#else
#define NS_EPI_TRACE(tag) do { } while (0)
#endif
// Mask stage helper: zero the dropped elements of one 128-byte row (64 bf16, eight 16-byte chunks stored 128B-swizzled:
// logical chunk c sits at position c ^ sw) of an operand tile in shared memory.  w = the row's 64 drop flags, bit = column.
// The loop runs over LOGICAL chunks: the 8 consecutive rows of a quarter warp then touch 8 different positions, i.e. all 32
// banks (walking positions instead puts every lane of the warp on the same four banks: 8-way conflicts on each of the 16
// accesses, measured 48 -> 80 us at K = 512).  All eight loads are issued before the first use.  Four flags become four byte
// sign bits with one multiply (bit j of a nibble lands on bit 8 j + 7: 0x10204080 = 2^7 + 2^14 + 2^21 + 2^28, no two partial
// products collide), and PRMT's sign-replicate mode (selector nibble bit 3; __byte_perm documents only 3 bits, hence the PTX)
// turns a pair of them into the 32-bit mask of a packed bf16 pair: 2 instructions per word after that.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ void mask_row128(uint32_t row_addr, uint32_t sw, uint2 w) {
  uint32_t v[8][4];
#pragma unroll
  for (int c = 0; c < 8; ++c) ld_shared_v4(row_addr + ((static_cast<uint32_t>(c) ^ sw) << 4), v[c][0], v[c][1], v[c][2], v[c][3]);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t m8 = (c < 4 ? w.x : w.y) >> (8 * (c & 3));
    const uint32_t slo = (m8 & 0xFu) * 0x10204080u, shi = ((m8 >> 4) & 0xFu) * 0x10204080u;
    v[c][0] &= ~prmt(slo, 0u, 0x9988u);
    v[c][1] &= ~prmt(slo, 0u, 0xBBAAu);
    v[c][2] &= ~prmt(shi, 0u, 0x9988u);
    v[c][3] &= ~prmt(shi, 0u, 0xBBAAu);
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) st_shared_v4(row_addr + ((static_cast<uint32_t>(c) ^ sw) << 4), v[c][0], v[c][1], v[c][2], v[c][3]);
}
// AM = 1 ("A-operand mask", BN = 32 only): t = alpha (x . keep_g) A_g^T, the LoRA down product under branch dropout.  The
// 32-wide kernel's second epilogue warpgroup (warps 8..11, idle at this width) becomes a MASK STAGE between the TMA and the
// MMA: thread = one row of the 128 x 64 A tile; it waits for the tile, zeroes the dropped elements of its 128-byte swizzled row
// in place (flags from the row-major bit plane: one 8-byte load per row and k block, fetched two blocks ahead;
// see mask_row128), fences the generic writes for the async proxy
// and arrives on the barrier the MMA issuer waits on.  Stacked adapters (q/k/v) are consecutive column tiles: tile g masks the
// same x tile (an L2 hit) with plane g.
template <int BN, int CG, int PM, int AM>
__global__ void __launch_bounds__(kNtThreads, 1)
gemm_nt_kernel(const __grid_constant__ Maps maps, const __grid_constant__ TileProg p) {
  using Cfg = NtCfg<BN, CG, PM>;
  static_assert(!AM || (BN == 32 && CG == 1 && !PM), "the mask stage lives in the 32-wide single-CTA kernel");
  const int S = p.stages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);      // 128B-swizzled tiles need 1 KB alignment: no static shared memory here,
  if (smem_base & 1023u) __trap();                    // so the dynamic window starts aligned (and every byte of it is budgeted)
  const uint32_t staging_base = smem_base + S * Cfg::kStageBytes;
  const uint32_t bar_base = staging_base + static_cast<uint32_t>(p.staging_tiles) * Cfg::kStagingTile;
  // barrier addresses
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  auto in_full = [&](int w) { return bar_base + 8u * (2 * S + 5 + w); };    // one per epilogue warp (8)
  auto mfull_bar = [&](int s) { return in_full(s); };                       // AM: "tile masked" per ring stage (the TMA-staged
                                                                            // epilogue inputs do not exist at BN = 32)
  const uint32_t bias_base = bar_base + 256u;         // [tile parity][column half][128] floats
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();     // ns_set_pdl: the next kernel may start its own prologue now
  // Work units.  CG = 1: one 128-row tile per CTA per step.  CG = 2: the pair walks "super tiles" of two consecutive row
  // tiles (same column tile); this CTA takes row tile 2 * pair_index + rank.  An odd row-tile count leaves one phantom tile
  // whose loads are all out of range (zero fill) and whose stores are clipped.
  const int m_tiles_real = p.batches * p.tiles_per_batch;
  const int m_units = (CG == 2) ? (m_tiles_real + 1) / 2 : m_tiles_real;
  const int total_tiles = m_units * p.n_tiles;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int tile_first = (CG == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = (CG == 2) ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  int tr_n = 0;
  auto trace = [&](int region, long long tag) {
    if (p.trace != nullptr && blockIdx.x < 2 && tr_n < 512) {
      p.trace[(region * 512 + tr_n) * 2] = tag;
      p.trace[(region * 512 + tr_n) * 2 + 1] = clock64();
      ++tr_n;
    }
  };
  auto decode = [&](int tile, int& m_tile, int& n_tile) {
    int mu;
    if (p.m_fast) { n_tile = p.fd_units.div(tile); mu = tile - n_tile * m_units; }
    else { mu = p.fd_ntiles.div(tile); n_tile = tile - mu * p.n_tiles; }
    m_tile = (CG == 2) ? 2 * mu + static_cast<int>(cta_rank) : mu;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b[0]);
    if (p.nseg > 1) {
      tma_prefetch_desc(&maps.a[1]);
      tma_prefetch_desc(&maps.b[1]);
    }
    if (p.tma_out) {
      tma_prefetch_desc(&maps.d);
      if (p.epi.aux_out) tma_prefetch_desc(&maps.aux);
      if (p.tma_in) tma_prefetch_desc(&maps.in);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);     // CG = 2: only the leader's copy is used; it counts the bytes of both CTAs
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), AM ? 4 : 8 * CG);   // one arrive per epilogue warp (of both CTAs)
    }
    for (int w = 0; w < 8; ++w) mbar_init(in_full(w), AM ? 4 : 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    if (CG == 2) { tmem_alloc2(tmem_slot, Cfg::kTmemCols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  pdl_wait();                  // everything above touched shared / tensor memory only; global memory from here on
  const uint32_t tmem_base = *tmem_slot_ptr;

  // warps 0..3 (TMA, MMA, TMEM owner, idle) give registers to the two epilogue warpgroups; the setmaxnreg sits inside
  // each role branch so that ptxas allocates every role under its own limit
  if (warp == 0) {
    // ================================================================ TMA producer
    setmaxnreg_dec<40>();
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        int m_tile, n_tile;
        decode(tile, m_tile, n_tile);
        const int b = p.fd_tpb.div(m_tile);
        const int t0 = (m_tile - b * p.tiles_per_batch) * kBM;
        const int n0 = n_tile * BN;
        for (int s = 0; s < p.nseg; ++s) {
          const Seg& sg = p.seg[s];
          const int kbase = sg.a_ngrp > 0 ? (n0 / sg.a_ngrp) * sg.a_kstep : 0;
          for (int kb = 0; kb < sg.kblocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            trace(blockIdx.x == 0 ? 0 : 2, 100 + stage);
            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
            const uint32_t sb = sa + kABytes;
            if (CG == 2) {
              // both CTAs' bytes are counted on the LEADER's barrier, which the leader arms with the pair's total.  The peer
              // needs no arrive of its own: it cannot touch phase n+1 of a stage before the leader's MMAs of phase n have
              // committed, and the phase cannot complete without the peer's bytes.
              const uint32_t lead_full = mapa_cluster(full_bar(stage), 0);
              if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
              tma_load_4d_2sm(&maps.a[sg.a_map], lead_full, sa, kbase + kb * kBK, sg.a_par, t0 + sg.a_off, b);
              tma_load_3d_2sm(&maps.b[sg.b_map], lead_full, sb, kb * kBK, n0 + static_cast<int>(cta_rank) * (BN / 2), sg.b_tap);
            } else {
              mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
              tma_load_4d(&maps.a[sg.a_map], full_bar(stage), sa, kbase + kb * kBK, sg.a_par, t0 + sg.a_off, b);
              tma_load_3d(&maps.b[sg.b_map], full_bar(stage), sb, kb * kBK, n0, sg.b_tap);
            }
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    setmaxnreg_dec<40>();
    constexpr uint32_t idesc = umma_idesc_bf16(kBM * CG, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = tile_first; tile < total_tiles && cta_rank == 0; tile += tile_step, ++it) {   // CG = 2: the leader issues
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_main = tmem_base + static_cast<uint32_t>(acc * BN);
      uint32_t acc_main = 0, acc_p = 0;                   // accumulate flags: main tile, masked-product tile (PM)
      for (int s = 0; s < p.nseg; ++s) {
        const Seg& sg = p.seg[s];
        const bool to_p = PM && s == p.nseg - 1;
        const uint32_t d_tmem = d_main + (to_p ? static_cast<uint32_t>(2 * BN) : 0u);
        uint32_t accumulate = to_p ? acc_p : acc_main;
        for (int kb = 0; kb < sg.kblocks; ++kb) {
          mbar_wait(AM ? mfull_bar(stage) : full_bar(stage), phase);
          if (lane == 0) trace(1, 200 + stage);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
            const uint32_t sb = sa + kABytes;
            const int ksteps = (kb == sg.kblocks - 1) ? sg.last_ksteps : 4;
            const uint64_t adesc = umma_smem_desc(sa, 16, 1024);
            const uint64_t bdesc = umma_smem_desc(sb, 16, 1024);
            for (int k = 0; k < ksteps; ++k) {
              // advance 16 elements (32 B) along K inside the 128B swizzle row: +2 in the (addr >> 4) field
              if (CG == 2) umma_f16_cg2(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, accumulate);
              else umma_f16(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, accumulate);
              accumulate = 1;
            }
            // frees the smem stage (in both CTAs of a pair) when these MMAs have read it
            if (CG == 2) umma_commit_mc2(empty_bar(stage), 3); else umma_commit(empty_bar(stage));
          }
          __syncwarp();
          accumulate = 1;                                // (the elected lane set it; keep the warp's copies equal)
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (to_p) acc_p = accumulate; else acc_main = accumulate;
      }
      if (elect_one()) {
        if (CG == 2) umma_commit_mc2(tfull_bar(acc), 3); else umma_commit(tfull_bar(acc));
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    setmaxnreg_dec<40>();
  } else if (AM && warp >= 8) {
    // ================================================================ mask stage (AM)
    setmaxnreg_inc<232>();
    const int r = (warp - 8) * 32 + lane;                // row of the A tile
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const int nkb = p.seg[0].kblocks;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      int m_tile, n_tile;
      decode(tile, m_tile, n_tile);
      const int b = p.fd_tpb.div(m_tile);
      const int t = (m_tile - b * p.tiles_per_batch) * kBM + r;
      const bool valid = t < p.tout;                     // rows past the end arrive zero-filled
      const long long grow = static_cast<long long>(b) * p.tout + (valid ? t : 0);      // row of the plane = row of x
      uint2* brow = const_cast<uint2*>(reinterpret_cast<const uint2*>(p.am_bits + static_cast<long long>(n_tile) * p.am_gstride + grow * p.am_ld));
      const uint2 none = make_uint2(0u, 0u);
      if (p.am_seed) {
        // draw: the integer hash (about 240 instructions per row and k block) hides under the stream of x -- the stage has
        // ~800 cycles per k block before the TMA ring runs dry, and the separate generator launch (and its read here) goes away
        const uint32_t ms = __ldg(p.am_seed) ^ p.am_salts[n_tile & 3];
        const uint32_t hrow = static_cast<uint32_t>(grow);
        for (int kb = 0; kb < nkb; ++kb) {
          uint2 wc = none;
          if (valid) {
            wc.x = drop_plane_word(hrow, static_cast<uint32_t>(2 * kb), ms, p.am_thr);
            wc.y = drop_plane_word(hrow, static_cast<uint32_t>(2 * kb + 1), ms, p.am_thr);
            brow[kb] = wc;
          }
          mbar_wait(full_bar(stage), phase);
          mask_row128(smem_base + stage * Cfg::kStageBytes + static_cast<uint32_t>(r) * 128u, sw, wc);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(mfull_bar(stage));
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        continue;
      }
      uint2 w0 = (valid && nkb > 0) ? __ldg(brow) : none;
      uint2 w1 = (valid && nkb > 1) ? __ldg(brow + 1) : none;
      for (int kb = 0; kb < nkb; ++kb) {
        const uint2 wc = w0;
        w0 = w1;
        w1 = (valid && kb + 2 < nkb) ? __ldg(brow + kb + 2) : none;
        mbar_wait(full_bar(stage), phase);
        mask_row128(smem_base + stage * Cfg::kStageBytes + static_cast<uint32_t>(r) * 128u, sw, wc);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(mfull_bar(stage));
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ================================================================ epilogue
    setmaxnreg_inc<232>();
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which half of the tile's columns
    constexpr int kHalfCols = (BN >= 64) ? BN / 2 : BN;
    const EpiDev& e = p.epi;
    const bool use_tma = BN >= 128 && p.tma_out;
    // Every epilogue warp moves its OWN 32-row slice of the staging tiles (TMA boxes of 64 columns x 32 rows, a per-warp
    // input barrier, per-lane bulk groups): no named barrier anywhere in the epilogue.  With one issuer per column half the
    // four warps met at 4 barriers per 64-column chunk and ncu showed 2 barrier-stall cycles per issued instruction.
    const int ew = warp - 4;
    const uint32_t slice_off = static_cast<uint32_t>(q) * 32u * 128u;
    const uint32_t out_stage = staging_base + static_cast<uint32_t>(half) * (kBM * 128);
    const uint32_t in_stage = staging_base + static_cast<uint32_t>(2 + half) * (kBM * 128);   // input tile, or the aux output tile
    uint32_t in_phase = 0;
    // this warp's slice of the epilogue input tile (aux_in / residual) of (tile, column chunk c) -> shared memory
    auto issue_in = [&](int tile_, int c_) {
      int m_tile_, n_tile_;
      decode(tile_, m_tile_, n_tile_);
      mbar_expect_tx(in_full(ew), 32 * 128);
      const int b_ = p.fd_tpb.div(m_tile_);
      tma_load_3d(&maps.in, in_full(ew), in_stage + slice_off, n_tile_ * BN + half * kHalfCols + c_,
                  (m_tile_ - b_ * p.tiles_per_batch) * kBM + q * 32, p.in_batched ? b_ : 0);
    };
    if (use_tma && p.tma_in && tile_first < total_tiles) {
      if (elect_one()) issue_in(tile_first, 0);
    }
    const uint32_t lead_tempty0 = (CG == 2) ? mapa_cluster(tempty_bar(0), 0) : 0u;   // the leader's accumulator-free barriers
    // Bias: global loads in the column loop cost an L2 round trip per 32-column slice (~500 cycles, measured: the
    // 227 KB carve-out leaves next to no L1).  Each lane fetches 4 of its half's columns one tile AHEAD into registers, the
    // warp parks them in shared memory after the accumulator-full wait and the slices read them back as broadcasts.  The
    // four warps of a half write identical rows; two tile parities keep a warp that runs ahead off the row a slower one
    // still reads (warps drift by at most one tile: the accumulator stage is recycled only after all of them arrived).
    constexpr int kBiasLanes = kHalfCols / 4;
    float4 bias_next = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch_bias = [&](int tile_) {
      if (e.bias == nullptr || lane >= kBiasLanes) return;
      int m_tile_, n_tile_;
      decode(tile_, m_tile_, n_tile_);
      const int col = n_tile_ * BN + half * kHalfCols + 4 * lane;
      if (p.vec_bias && col + 3 < p.N) {
        bias_next = __ldg(reinterpret_cast<const float4*>(e.bias + col));
      } else {
        bias_next.x = col < p.N ? __ldg(e.bias + col) : 0.f;
        bias_next.y = col + 1 < p.N ? __ldg(e.bias + col + 1) : 0.f;
        bias_next.z = col + 2 < p.N ? __ldg(e.bias + col + 2) : 0.f;
        bias_next.w = col + 3 < p.N ? __ldg(e.bias + col + 3) : 0.f;
      }
    };
    if (tile_first < total_tiles) fetch_bias(tile_first);
    // PM: this thread's dropout flags of the tile's columns in its half (kHalfCols / 32 words of its row), fetched one tile
    // ahead like the bias (an L2 round trip that must not sit between the accumulator wait and the arithmetic)
    uint2 drop_next = make_uint2(0u, 0u);
    auto fetch_drop = [&](int tile_) {
      if constexpr (PM) {
        int m_tile_, n_tile_;
        decode(tile_, m_tile_, n_tile_);
        const int b_ = p.fd_tpb.div(m_tile_);
        const int t_ = (m_tile_ - b_ * p.tiles_per_batch) * kBM + q * 32 + lane;
        const int col = n_tile_ * BN + half * kHalfCols;
        drop_next = make_uint2(0u, 0u);
        if (t_ < p.tout && m_tile_ < m_tiles_real && col < p.N) {
          const long long gm = static_cast<long long>(b_) * p.out_bs + static_cast<long long>(t_) * p.out_rs + p.out_off;
          drop_next = __ldg(reinterpret_cast<const uint2*>(e.drop_bits + gm * e.drop_ld + (col >> 5)));
        }
      }
    };
    if (tile_first < total_tiles) fetch_drop(tile_first);
    // The tile loop is compiled once per epilogue SHAPE (p.epi_mode, chosen on the host): with everything decided at run
    // time a plain bias epilogue executed ~1400 warp instructions per tile and warp for ~250 useful ones (branches over
    // the unused variants, register moves between them), spread over 90 KB of code.
    //   0 = generic (any epilogue, direct or TMA stores)   1 = TMA store, no activation / residual
    //   2 = GELU   3 = GELU + pre-activation output   4 = dGELU with the TMA-staged pre-activation   5 = TMA-staged residual
    auto run_tiles = [&](auto mode_c) {
    constexpr int MODE = decltype(mode_c)::value;
    const int act = MODE == 0 ? e.act : ((MODE == 2 || MODE == 3) ? NS_ACT_GELU : (MODE == 4 ? NS_ACT_DGELU : NS_ACT_NONE));
    const bool use_tma_m = MODE == 0 ? use_tma : true;
    const int tma_in = MODE == 0 ? p.tma_in : (MODE == 4 ? 1 : (MODE == 5 ? 2 : 0));
    const bool has_aux = MODE == 0 ? (e.act == NS_ACT_GELU && e.aux_out != nullptr) : (MODE == 3);
    const bool has_res = MODE == 0 ? (e.residual != nullptr) : (MODE == 5);
    int it = 0;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      int m_tile, n_tile;
      decode(tile, m_tile, n_tile);
      const int b = p.fd_tpb.div(m_tile);
      const int t_tile = (m_tile - b * p.tiles_per_batch) * kBM;
      const int t = t_tile + q * 32 + lane;
      const int n0 = n_tile * BN;
      const bool valid = t < p.tout && m_tile < m_tiles_real;
      const long long row = static_cast<long long>(b) * p.out_bs + static_cast<long long>(t) * p.out_rs + p.out_off;
      const long long res_row = e.res_mod > 0 ? ((static_cast<long long>(t) * p.out_rs + p.out_off) % e.res_mod) : row;
      NS_EPI_TRACE(380);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      NS_EPI_TRACE(300);
      const uint32_t bias_row = bias_base + static_cast<uint32_t>(((it & 1) * 2 + half) * 512);
      if (e.bias) {
        if (lane < kBiasLanes)
          st_shared_v4(bias_row + 16u * lane, __float_as_uint(bias_next.x), __float_as_uint(bias_next.y), __float_as_uint(bias_next.z),
                       __float_as_uint(bias_next.w));
        __syncwarp();
        if (tile + tile_step < total_tiles) fetch_bias(tile + tile_step);
      }
      const uint2 drop_cur = drop_next;
      (void)drop_cur;
      if (PM && tile + tile_step < total_tiles) fetch_drop(tile + tile_step);
      // One 32-column slice of this thread's row: TMEM -> registers -> bias / scale / activation / residual.
      // `z_out` receives the pre-activation when NS_ACT_GELU has an aux output.
      // `zin` (16 packed bf16 pairs) carries this slice of the TMA-staged input tile when has_in.
      const uint32_t acc_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + half * kHalfCols);
      // `v`: the 32 accumulator columns [c, c + 32) of this half, already read from TMEM (the caller keeps several loads in
      // flight behind one wait).
      auto slice = [&](const uint32_t (&v)[32], int c, int col0, int ncols, float (&x)[32], float (&z_out)[32], const uint32_t (&zin)[16],
                       bool has_in) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (e.bias) {                                     // columns past N hold zeros
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t b0, b1, b2, b3;
            ld_shared_v4(bias_row + 4u * static_cast<uint32_t>(c) + 16u * j, b0, b1, b2, b3);
            f2_unpack(f2_add(f2_pack(x[4 * j], x[4 * j + 1]), f2_pack(__uint_as_float(b0), __uint_as_float(b1))), x[4 * j], x[4 * j + 1]);
            f2_unpack(f2_add(f2_pack(x[4 * j + 2], x[4 * j + 3]), f2_pack(__uint_as_float(b2), __uint_as_float(b3))), x[4 * j + 2],
                      x[4 * j + 3]);
          }
        }
        if (col0 < e.alpha_cols) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < e.alpha_cols) x[j] *= e.alpha;
        }
        const bool full = (ncols == 32);
        if (act == NS_ACT_GELU) {
          if (e.aux_deriv) {                              // save gelu'(z) for the backward (one tanh serves both)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              uint64_t dg;
              f2_unpack(gelu_both2(f2_pack(x[2 * j], x[2 * j + 1]), dg), x[2 * j], x[2 * j + 1]);
              f2_unpack(dg, z_out[2 * j], z_out[2 * j + 1]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              z_out[2 * j] = x[2 * j]; z_out[2 * j + 1] = x[2 * j + 1];
              f2_unpack(gelu_fast2(f2_pack(x[2 * j], x[2 * j + 1])), x[2 * j], x[2 * j + 1]);
            }
          }
        } else if (act == NS_ACT_DGELU) {
          if (has_in && tma_in == 1) {
            if (e.aux_deriv) {                            // the saved tensor IS the derivative
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 zz = unpack_bf16x2(zin[j]);
                f2_unpack(f2_mul(f2_pack(x[2 * j], x[2 * j + 1]), f2_pack(zz.x, zz.y)), x[2 * j], x[2 * j + 1]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 zz = unpack_bf16x2(zin[j]);
                f2_unpack(f2_mul(f2_pack(x[2 * j], x[2 * j + 1]), dgelu_fast2(f2_pack(zz.x, zz.y))), x[2 * j], x[2 * j + 1]);
              }
            }
          } else if (valid) {
            float z[32];
            load32_bf16(reinterpret_cast<const __nv_bfloat16*>(e.aux_in) + row * e.ldaux + col0, p.vec_aux && full, ncols, z);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] *= e.aux_deriv ? z[j] : dgelu_fast(z[j]);
          }
        }
        if (has_res) {
          if (has_in && tma_in == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 rr = unpack_bf16x2(zin[j]);
              f2_unpack(f2_add(f2_pack(x[2 * j], x[2 * j + 1]), f2_pack(rr.x, rr.y)), x[2 * j], x[2 * j + 1]);
            }
          } else if (valid) {
            float r[32];
            load32_bf16(reinterpret_cast<const __nv_bfloat16*>(e.residual) + res_row * e.ldr + col0, p.vec_res && full, ncols, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] += r[j];
          }
        }
      };
      if (use_tma_m) {
        // ---- bf16 tiles travel through shared memory: each thread owns the 128-byte row slice (64 columns) of a
        // 128B-swizzled [128 rows][64 columns] staging tile; one lane per column half issues the TMA loads / stores.  The
        // direct path below moves 16 bytes per lane from / to 32 different rows per instruction (LSU bound).
        const uint32_t row_off = static_cast<uint32_t>(q * 32 + lane) * 128u;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        auto stage_wait = [&]() {                                             // the previous store of this slice has read it
          if (elect_one()) bulk_wait_read0();
          __syncwarp();
        };
        auto stage_write = [&](uint32_t tile_addr, int sidx, const float (&y)[32]) {   // this thread's 32 columns -> 4 swizzled chunks
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            st_shared_v4(tile_addr + row_off + ((static_cast<uint32_t>(4 * sidx + cc) ^ sw) << 4), pack_bf16x2(y[8 * cc], y[8 * cc + 1]),
                         pack_bf16x2(y[8 * cc + 2], y[8 * cc + 3]), pack_bf16x2(y[8 * cc + 4], y[8 * cc + 5]),
                         pack_bf16x2(y[8 * cc + 6], y[8 * cc + 7]));
        };
        auto stage_write_packed = [&](uint32_t tile_addr, int sidx, const uint32_t (&y)[16]) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            st_shared_v4(tile_addr + row_off + ((static_cast<uint32_t>(4 * sidx + cc) ^ sw) << 4), y[4 * cc], y[4 * cc + 1], y[4 * cc + 2],
                         y[4 * cc + 3]);
        };
        auto stage_commit = [&](const CUtensorMap* map, uint32_t tile_addr, int col0) {
          fence_proxy_async();
          __syncwarp();
          if (elect_one()) {
            tma_store_3d(map, tile_addr + slice_off, col0, t_tile + q * 32, b);
            bulk_commit();
          }
        };
#pragma unroll 1
        for (int c = 0; c < kHalfCols; c += 64) {
          const int col0 = n0 + half * kHalfCols + c;
          if (c > 0 && col0 >= p.N) break;                                    // uniform over the 4 warps of this half
          uint32_t zin[2][16];
          if (tma_in) {
            mbar_wait(in_full(ew), in_phase);
            in_phase ^= 1u;
#pragma unroll
            for (int sidx = 0; sidx < 2; ++sidx)
#pragma unroll
              for (int cc = 0; cc < 4; ++cc)
                ld_shared_v4(in_stage + row_off + ((static_cast<uint32_t>(4 * sidx + cc) ^ sw) << 4), zin[sidx][4 * cc], zin[sidx][4 * cc + 1],
                             zin[sidx][4 * cc + 2], zin[sidx][4 * cc + 3]);
            __syncwarp();                                                     // every lane has its row: the slice may be refilled
            if (elect_one()) {                                                // prefetch the next chunk (or the next tile's first)
              const int c2 = c + 64;
              if (c2 < kHalfCols && col0 + 64 < p.N) issue_in(tile, c2);
              else if (tile + tile_step < total_tiles) issue_in(tile + tile_step, 0);
            }
          }
          NS_EPI_TRACE(310);
          if (has_aux) {
            // pre-activation tile -> its own staging tile and out; then the activated tile.  Loads and arithmetic of both
            // slices first (the previous chunk's two stores drain meanwhile), then wait / write / commit per tile.
            uint32_t pk_z[2][16], pk_out[2][16];
            {
              uint32_t v[2][32];
              tmem_ld32(acc_addr + static_cast<uint32_t>(c), v[0]);
              tmem_ld32(acc_addr + static_cast<uint32_t>(c + 32), v[1]);
              tmem_ld_wait();
#pragma unroll
              for (int sidx = 0; sidx < 2; ++sidx) {
                const int cs = col0 + 32 * sidx;
                float x[32], z[32];
                if (cs < p.N) {
                  slice(v[sidx], c + 32 * sidx, cs, min(32, p.N - cs), x, z, zin[sidx], false);
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) { x[j] = 0.f; z[j] = 0.f; }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  pk_z[sidx][j] = pack_bf16x2(z[2 * j], z[2 * j + 1]);
                  pk_out[sidx][j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
                }
              }
            }
            // ONE bulk group per chunk for the two tiles: one wait, one proxy fence, one commit instead of two each (the store
            // plumbing -- commit, fence, the warp syncs around the elected lane -- was a quarter of this epilogue's samples)
            stage_wait();
#pragma unroll
            for (int sidx = 0; sidx < 2; ++sidx) stage_write_packed(in_stage, sidx, pk_z[sidx]);
#pragma unroll
            for (int sidx = 0; sidx < 2; ++sidx) stage_write_packed(out_stage, sidx, pk_out[sidx]);
            NS_EPI_TRACE(340);
            fence_proxy_async();
            __syncwarp();
            if (elect_one()) {
              tma_store_3d(&maps.aux, in_stage + slice_off, col0, t_tile + q * 32, b);
              tma_store_3d(&maps.d, out_stage + slice_off, col0, t_tile + q * 32, b);
              bulk_commit();
            }
          } else {
            // both 32-column slices of the chunk: two TMEM loads in flight behind one wait, then the arithmetic -- all of it
            // while the previous chunk's TMA store still reads this warp's staging slice; the wait for that store sits right
            // before the slice is rewritten.
            uint32_t v[2][32], pk[2][16];
            tmem_ld32(acc_addr + static_cast<uint32_t>(c), v[0]);
            tmem_ld32(acc_addr + static_cast<uint32_t>(c + 32), v[1]);
            if constexpr (PM) {
              // masked second product: v += keep ? P : 0 (kHalfCols == 64: the chunk is the whole half, drop_cur its two words)
              uint32_t pv[2][32];
              tmem_ld32(acc_addr + static_cast<uint32_t>(2 * BN + c), pv[0]);
              tmem_ld32(acc_addr + static_cast<uint32_t>(2 * BN + c + 32), pv[1]);
              tmem_ld_wait();
#pragma unroll
              for (int sidx = 0; sidx < 2; ++sidx) {
                const uint32_t dw = sidx == 0 ? drop_cur.x : drop_cur.y;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (!((dw >> j) & 1u)) v[sidx][j] = __float_as_uint(__uint_as_float(v[sidx][j]) + __uint_as_float(pv[sidx][j]));
              }
            } else {
              tmem_ld_wait();
            }
            NS_EPI_TRACE(322);
#pragma unroll
            for (int sidx = 0; sidx < 2; ++sidx) {
              const int cs = col0 + 32 * sidx;
              float x[32], z[32];
              if (cs < p.N) {
                slice(v[sidx], c + 32 * sidx, cs, min(32, p.N - cs), x, z, zin[sidx], tma_in != 0);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = 0.f;
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[sidx][j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
            }
            NS_EPI_TRACE(323);
            stage_wait();
            NS_EPI_TRACE(320);
#pragma unroll
            for (int sidx = 0; sidx < 2; ++sidx) stage_write_packed(out_stage, sidx, pk[sidx]);
          }
          NS_EPI_TRACE(350);
          if (!has_aux) stage_commit(&maps.d, out_stage, col0);
          NS_EPI_TRACE(360);
        }
      } else if (BN >= 64 || half == 0) {
#pragma unroll 1
        for (int c = 0; c < kHalfCols; c += 32) {
          const int col0 = n0 + half * kHalfCols + c;
          if (col0 >= p.N) break;   // warp-uniform
          const int ncols = min(32, p.N - col0);
          const bool full = (ncols == 32);
          float x[32], z[32];
          const uint32_t no_in[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
          uint32_t v[32];
          tmem_ld32(acc_addr + static_cast<uint32_t>(c), v);
          tmem_ld_wait();
          slice(v, c, col0, ncols, x, z, no_in, false);
          if (act == NS_ACT_GELU && e.aux_out && valid)
            store32_bf16(reinterpret_cast<__nv_bfloat16*>(e.aux_out) + row * e.ldaux + col0, p.vec_aux && full, ncols, z);
          if (valid) {
            if (e.out_f32)
              store32_f32(reinterpret_cast<float*>(p.D) + row * p.ldd + col0, p.vec_out && full, ncols, x);
            else
              store32_bf16(reinterpret_cast<__nv_bfloat16*>(p.D) + row * p.ldd + col0, p.vec_out && full, ncols, x);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(lead_tempty0 + 8u * acc); else mbar_arrive(tempty_bar(acc));
      }
      NS_EPI_TRACE(370);
    }
    };
    if constexpr (PM) {                                   // the host only sends plain and dGELU epilogues here
      if (p.epi_mode == 4) run_tiles(std::integral_constant<int, 4>{});
      else run_tiles(std::integral_constant<int, 1>{});
    } else if constexpr (BN >= 128) {
      switch (p.epi_mode) {
        case 1: run_tiles(std::integral_constant<int, 1>{}); break;
        case 2: run_tiles(std::integral_constant<int, 2>{}); break;
        case 3: run_tiles(std::integral_constant<int, 3>{}); break;
        case 4: run_tiles(std::integral_constant<int, 4>{}); break;
        case 5: run_tiles(std::integral_constant<int, 5>{}); break;
        default: run_tiles(std::integral_constant<int, 0>{}); break;
      }
    } else {
      run_tiles(std::integral_constant<int, 0>{});
    }
    if (use_tma) {
      if (elect_one()) bulk_wait0();                                          // outstanding tile stores of this thread
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // nobody leaves while the peer may still signal / read this CTA
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ TN (wgrad) kernel
struct TnProg {
  int batches, tout;          // contraction = batches x tout rows (tout padded per batch by TMA zero fill)
  int blocks_per_batch;       // ceil(tout / 64)
  int total_blocks, blocks_per_split, nsplit;
  int i_tiles, j_tiles, ntaps;
  int I, J, bj;               // logical sizes; bj = UMMA N (multiple of 16, <= 256)
  int y_off[3], y_par[3];     // per tap: row offset / parity coordinate of Y
  long long si, sj, stap;     // output strides
  float* G;
  float alpha;
  int stage_bytes;            // icta x 16 KB (X) + 8 KB per 64 columns of Y
  int icta;                   // 128-row output tiles per CTA (1..4): with a narrow Y (J <= 64) one CTA streams up to 512
                              // contiguous X columns per contraction row and keeps icta accumulators (icta * bj TMEM columns)
  int stages;                 // operand ring depth
  int grp_i, grp_j;           // > 0: block-diagonal -- i tile i0 belongs to group g = i0 / grp_i and uses Y columns [g*grp_j, g*grp_j + bj)
  float alpha_grp[4];
  const uint32_t* xbits;      // != NULL: X is masked with this row-major dropout plane on its way to the MMA (dA = dt'^T (x . keep))
  long long xbits_ld;
};
struct TnMaps {
  CUtensorMap x;
  CUtensorMap y;
};
constexpr int kTnThreads = 256;
constexpr int kTnMaxStages = 4;
constexpr int kTnABytes = 128 * 64 * 2;        // two [64 m][64 i] boxes per output row tile
constexpr int kTnSmemMax = 232448;

__global__ void __launch_bounds__(kTnThreads, 1)
gemm_tn_kernel(const __grid_constant__ TnMaps maps, const __grid_constant__ TnProg p) {
  const int S = p.stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S * p.stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kTnMaxStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kTnMaxStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kTnMaxStages + 1);
  auto mfull_bar = [&](int s) { return bar_base + 8u * (2 * kTnMaxStages + 2 + s); };   // X tile masked (xbits)
  const bool masked = p.xbits != nullptr;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // decode work item: (tile, split)
  const int split = blockIdx.x % p.nsplit;
  int tile = blockIdx.x / p.nsplit;
  const int j_tile = tile % p.j_tiles; tile /= p.j_tiles;
  const int i_tile = tile % p.i_tiles; tile /= p.i_tiles;
  const int tap = tile;
  const int blk0 = split * p.blocks_per_split;
  const int blk1 = min(p.total_blocks, blk0 + p.blocks_per_split);
  const int nblk = max(0, blk1 - blk0);
  const int i0 = i_tile * 128 * p.icta;         // i_tile counts groups of icta row tiles
  const int j0 = j_tile * 256;
  const int grp = p.grp_i > 0 ? i0 / p.grp_i : 0;
  const int yj0 = p.grp_i > 0 ? grp * p.grp_j : j0;      // column of Y this tile starts at (block-diagonal: the group's block)
  const float alpha = p.grp_i > 0 ? p.alpha_grp[grp] : p.alpha;
  const int jboxes = (p.bj + 63) / 64;
  const uint32_t a_bytes = static_cast<uint32_t>(p.icta) * kTnABytes;
  const uint32_t stage_tx = a_bytes + static_cast<uint32_t>(jboxes) * 8192u;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(mfull_bar(s), 4);
    }
    mbar_init(tfull_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int blk = blk0; blk < blk1; ++blk) {
        const int b = blk / p.blocks_per_batch;
        const int t0 = (blk % p.blocks_per_batch) * 64;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * p.stage_bytes;
        const uint32_t sb = sa + a_bytes;
        mbar_expect_tx(full_bar(stage), stage_tx);
        for (int g = 0; g < 2 * p.icta; ++g)                  // boxes past I arrive zero-filled
          tma_load_4d(&maps.x, full_bar(stage), sa + 8192u * g, i0 + 64 * g, 0, t0, b);
        for (int g = 0; g < jboxes; ++g)
          tma_load_4d(&maps.y, full_bar(stage), sb + 8192u * g, yj0 + 64 * g, p.y_par[tap], t0 + p.y_off[tap], b);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, p.bj, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (int blk = blk0; blk < blk1; ++blk) {
      mbar_wait(masked ? mfull_bar(stage) : full_bar(stage), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + stage * p.stage_bytes;
        const uint32_t sb = sa + a_bytes;
        // MN-major: LBO = distance between 64-element MN groups (8192 B), SBO = distance between 8-row K groups (1024 B)
        const uint64_t bdesc = umma_smem_desc(sb, 8192, 1024);
        for (int it = 0; it < p.icta; ++it) {
          const uint64_t adesc = umma_smem_desc(sa + static_cast<uint32_t>(it) * kTnABytes, 8192, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 16 contraction rows = 2048 B  -> +128 in the (addr >> 4) field
            umma_f16(tmem_base + static_cast<uint32_t>(it * p.bj), adesc + static_cast<uint64_t>(128 * k),
                     bdesc + static_cast<uint64_t>(128 * k), idesc, (accumulate | (k > 0)) ? 1u : 0u);
        }
        accumulate = 1;
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) umma_commit(tfull_bar);
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    if (masked) {
      // Mask stage: the epilogue warps have nothing to do until the last block.  X boxes are [64 contraction rows][64 columns]
      // (128-byte swizzled rows); thread = (row, box parity), icta boxes each, one 8-byte flag load per box, fetched a block ahead.
      const int tid = (warp - 4) * 32 + lane;
      const int row = tid & 63, bpar = tid >> 6;
      const uint32_t sw = static_cast<uint32_t>(row & 7);
      const uint2 none = make_uint2(0u, 0u);
      auto fetch = [&](int blk, uint2 (&w)[4]) {
        const int b = blk / p.blocks_per_batch;
        const int t = (blk % p.blocks_per_batch) * 64 + row;
        const bool ok = blk < blk1 && t < p.tout;
        const uint32_t* br = p.xbits + (static_cast<long long>(b) * p.tout + (ok ? t : 0)) * p.xbits_ld;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = i0 + 64 * (2 * g + bpar);
          w[g] = (ok && g < p.icta && col < p.I) ? __ldg(reinterpret_cast<const uint2*>(br + (col >> 5))) : none;
        }
      };
      uint2 wn[4];
      fetch(blk0, wn);
      int stage = 0;
      uint32_t phase = 0;
      for (int blk = blk0; blk < blk1; ++blk) {
        uint2 wc[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) wc[g] = wn[g];
        fetch(blk + 1, wn);
        mbar_wait(full_bar(stage), phase);
        const uint32_t sa = smem_base + stage * p.stage_bytes;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g < p.icta && (wc[g].x | wc[g].y))
            mask_row128(sa + 8192u * (2 * g + bpar) + static_cast<uint32_t>(row) * 128u, sw, wc[g]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(mfull_bar(stage));
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
    if (nblk > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int it = 0; it < p.icta; ++it) {
        const int i = i0 + it * 128 + q * 32 + lane;
        float* g_row = p.G + static_cast<long long>(tap) * p.stap + static_cast<long long>(i) * p.si;
#pragma unroll 1
        for (int c = 0; c < p.bj; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(it * p.bj + c), v);
          tmem_ld_wait();
          if (i < p.I) {
            float* g0 = g_row + static_cast<long long>(j0 + c) * p.sj;
            if (p.sj == 1 && j0 + c + 32 <= p.J && (reinterpret_cast<uintptr_t>(g0) & 15) == 0) {
              // this thread's 32 columns are contiguous: eight 16-byte reductions instead of 32 scalar ones to 32 different
              // lines per warp instruction (the split-K tail was a third of the kernel at the LoRA shapes)
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g0 + j), "f"(alpha * __uint_as_float(v[j])),
                             "f"(alpha * __uint_as_float(v[j + 1])), "f"(alpha * __uint_as_float(v[j + 2])),
                             "f"(alpha * __uint_as_float(v[j + 3]))
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int jj = j0 + c + j;
                if (jj < p.J) atomicAdd(g_row + static_cast<long long>(jj) * p.sj, alpha * __uint_as_float(v[j]));
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------ host launchers
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BN, int CG, int PM = 0, int AM = 0>
static int launch_nt(const Maps& maps, TileProg& prog, cudaStream_t st) {
  using Cfg = NtCfg<BN, CG, PM>;
  static bool attr_done = false;
  if (!attr_done) {
    NS_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<BN, CG, PM, AM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kMaxSmem));
    attr_done = true;
  }
  prog.n_tiles = (prog.N + BN - 1) / BN;
  const long long m_tiles = static_cast<long long>(prog.batches) * prog.tiles_per_batch;
  // consecutive tiles should share the LARGER operand tile stream's counterpart: few M tiles and many N tiles (the tied
  // vocabulary projection) -> walk M first so each weight tile is fetched once and reused from L2
  prog.m_fast = (m_tiles * 4 < prog.n_tiles) ? 1 : 0;
  const long long units = ((CG == 2) ? (m_tiles + 1) / 2 : m_tiles) * prog.n_tiles;     // work units per CTA (pair)
  int grid = static_cast<int>(units * CG < sm_count() ? units * CG : sm_count());
  if (CG == 2) grid &= ~1;
  if (grid <= 0) return NS_OK;
  const bool has_aux = (prog.epi.act == NS_ACT_GELU && prog.epi.aux_out);
  prog.fd_units = make_fastdiv((CG == 2) ? (m_tiles + 1) / 2 : m_tiles);
  prog.fd_ntiles = make_fastdiv(prog.n_tiles);
  prog.fd_tpb = make_fastdiv(prog.tiles_per_batch);
  prog.staging_tiles = (BN >= 128 && prog.tma_out) ? ((has_aux || prog.tma_in) ? 4 : 2) : 0;
  prog.stages = Cfg::stages_for(prog.staging_tiles);
  prog.epi_mode = 0;
  if (BN >= 128 && prog.tma_out) {
    const EpiDev& e = prog.epi;
    if (e.act == NS_ACT_NONE && !e.residual && !prog.tma_in) prog.epi_mode = 1;
    else if (e.act == NS_ACT_GELU && !e.residual && !prog.tma_in) prog.epi_mode = e.aux_out ? 3 : 2;
    else if (e.act == NS_ACT_DGELU && !e.residual && prog.tma_in == 1) prog.epi_mode = 4;
    else if (e.act == NS_ACT_NONE && e.residual && prog.tma_in == 2) prog.epi_mode = 5;
  }
  static const bool generic_only = getenv("NS_GEMM_GENERIC_EPILOGUE") != nullptr;
  if (generic_only) prog.epi_mode = 0;
  prog.trace = get_attn_trace();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kNtThreads);
  cfg.dynamicSmemBytes = Cfg::smem_bytes(prog.stages, prog.staging_tiles);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (PM && !(prog.epi_mode == 1 || prog.epi_mode == 4)) {
    set_error("ns_gemm_nt: the dropout-masked second product supports plain and NS_ACT_DGELU epilogues with bf16 TMA tiles only");
    return NS_ERR_UNSUPPORTED;
  }
  NS_CUDA(cudaLaunchKernelEx(&cfg, gemm_nt_kernel<BN, CG, PM, AM>, maps, prog));
  NS_LAUNCH_CHECK();
  count(C_GEMM_TC);
  return NS_OK;
}

// Tile width: the widest UMMA N that covers the output, narrowed (down to 64) while the grid would leave a quarter of the
// SMs idle -- the M = B*L = 2048 products of the decoder and the K = 51872 product dlogits * E have only 16 row tiles.
static int choose_bn(long long m_tiles, int N) {
  int bn = (N > 128) ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
  const long long want = static_cast<long long>(sm_count()) * 3 / 4;
  static const int min_bn = getenv("NS_GEMM_MIN_BN") ? atoi(getenv("NS_GEMM_MIN_BN")) : 32;   // (decode position: 0.954 -> 0.933 ms at 32)
  const int floor_bn = (m_tiles == 1 && min_bn >= 32) ? min_bn : 64;      // one row tile (decoder steps): narrower tiles, more CTAs
  while (bn > floor_bn && m_tiles * ((N + bn - 1) / bn) < want) bn >>= 1;
  return bn;
}

// CTA pairs (cta_group::2) for the wide tiles when there is enough work to fill the pairs
static int choose_cg(long long m_tiles, int N, int bn) {
  static const bool disabled = getenv("NS_GEMM_NO_2CTA") != nullptr;
  if (disabled || bn != 256) return 1;
  const long long units = ((m_tiles + 1) / 2) * ((N + bn - 1) / bn);
  return units * 2 >= static_cast<long long>(sm_count()) * 3 / 4 ? 2 : 1;
}

static int dispatch_nt(const Maps& maps, TileProg& prog, cudaStream_t st, int bn, int cg) {
  if (prog.am_bits) return launch_nt<32, 1, 0, 1>(maps, prog, st);
  if (prog.epi.drop_bits) return cg == 2 ? launch_nt<128, 2, 1>(maps, prog, st) : launch_nt<128, 1, 1>(maps, prog, st);
  if (bn == 256) return cg == 2 ? launch_nt<256, 2>(maps, prog, st) : launch_nt<256, 1>(maps, prog, st);
  if (bn == 128) return launch_nt<128, 1>(maps, prog, st);
  if (bn == 64) return launch_nt<64, 1>(maps, prog, st);
  return launch_nt<32, 1>(maps, prog, st);
}

static void fill_epi(TileProg& prog, const EpiDev& e) {
  prog.epi = e;
  prog.vec_bias = e.bias && aligned16(e.bias);
  const bool f32 = e.out_f32;
  prog.vec_out = aligned16(prog.D) && (prog.ldd % (f32 ? 4 : 8) == 0);
  const void* aux = e.act == NS_ACT_GELU ? e.aux_out : e.aux_in;
  prog.vec_aux = aux && aligned16(aux) && (e.ldaux % 8 == 0);
  prog.vec_res = e.residual && aligned16(e.residual) && (e.ldr % 8 == 0);
}

// Output tensor maps for the TMA-store epilogue (bf16 output, 16-byte aligned rows, N > 64 so that BN >= 128 runs).
static int setup_out_maps(Maps& maps, TileProg& prog, int bn) {
  const EpiDev& e = prog.epi;
  maps.d = maps.a[0];
  maps.aux = maps.a[0];
  maps.in = maps.a[0];
  prog.tma_out = 0;
  prog.tma_in = 0;
  static const bool disabled = getenv("NS_GEMM_NO_TMA_STORE") != nullptr;
  const bool has_aux = (e.act == NS_ACT_GELU && e.aux_out);
  if (disabled || e.out_f32 || bn < 128 || !prog.vec_out || (has_aux && !prog.vec_aux)) return NS_OK;
  auto mk = [&](CUtensorMap* m, void* base, long long ld) -> int {
    uint64_t dims[3] = {(uint64_t)prog.N, (uint64_t)prog.tout, (uint64_t)prog.batches};
    const long long bs = prog.batches > 1 ? prog.out_bs : static_cast<long long>(prog.tout) * prog.out_rs;
    uint64_t str[2] = {(uint64_t)(prog.out_rs * ld * 2), (uint64_t)(bs * ld * 2)};
    uint32_t box[3] = {64, 32, 1};           // one epilogue warp's slice
    return make_map(m, reinterpret_cast<char*>(base) + prog.out_off * ld * 2, 3, dims, str, box);
  };
  int r = mk(&maps.d, prog.D, prog.ldd);
  if (r) return r;
  if (has_aux && (r = mk(&maps.aux, e.aux_out, e.ldaux))) return r;
  prog.tma_out = 1;
  // one epilogue INPUT tile kind can ride the same way: the saved pre-activation of NS_ACT_DGELU, else the residual
  maps.in = maps.a[0];
  prog.tma_in = 0;
  prog.in_batched = 1;
  if (has_aux) {
    // the second staging tile of each half carries the pre-activation output; a residual (conv C's position table) stays on
    // the direct-load path
  } else if (e.act == NS_ACT_DGELU && prog.vec_aux) {
    if ((r = mk(&maps.in, const_cast<void*>(e.aux_in), e.ldaux))) return r;
    prog.tma_in = 1;
  } else if (e.residual && prog.vec_res) {
    if (e.res_mod == 0) {
      if ((r = mk(&maps.in, const_cast<void*>(e.residual), e.ldr))) return r;
      prog.tma_in = 2;
    } else if (prog.out_rs == 1 && prog.out_off == 0 && prog.tout <= e.res_mod) {
      // position table (rows t of every batch read table row t): a tile never wraps, no batch coordinate
      uint64_t dims[3] = {(uint64_t)prog.N, (uint64_t)e.res_mod, 1};
      uint64_t str[2] = {(uint64_t)(e.ldr * 2), (uint64_t)(e.ldr * 2) * (uint64_t)e.res_mod};
      uint32_t box[3] = {64, 32, 1};
      if ((r = make_map(&maps.in, e.residual, 3, dims, str, box))) return r;
      prog.tma_in = 2;
      prog.in_batched = 0;
    }
  }
  return NS_OK;
}

static void seg_from_k(Seg& s, int K) {
  s.kblocks = (K + kBK - 1) / kBK;
  const int rem = K - (s.kblocks - 1) * kBK;
  s.last_ksteps = (rem + 15) / 16;
}

// Fast path of ns_gemm_nt.  Returns NS_ERR_UNSUPPORTED when the shape does not qualify.
int gemm_nt_fast(long long M, int N, int K, const void* A, long long lda, const void* W, long long ldw, void* D,
                 long long ldd, const EpiDev& epi, const void* A2, long long lda2, const void* W2, long long ldw2,
                 int K2, int a2_ngrp, cudaStream_t st) {
  if (M <= 0 || N <= 0) return NS_OK;
  if (epi.drop_bits) set_error("ns_gemm_nt: operands of the dropout-masked product do not qualify for the tcgen05 path (alignment / K %% 16)");
  if (K % 16 != 0 || lda % 8 != 0 || ldw % 8 != 0 || !aligned16(A) || !aligned16(W)) return NS_ERR_UNSUPPORTED;
  if (A2 && (K2 % 16 != 0 || lda2 % 8 != 0 || ldw2 % 8 != 0 || !aligned16(A2) || !aligned16(W2))) return NS_ERR_UNSUPPORTED;
  if (M > 0x7fffffffLL) return NS_ERR_UNSUPPORTED;
  Maps maps;
  TileProg prog;
  memset(&prog, 0, sizeof(prog));
  int bn = choose_bn((M + kBM - 1) / kBM, N);
  int cg = choose_cg((M + kBM - 1) / kBM, N, bn);
  const int a1g = epi.a_group_cols;
  if (a1g > 0) {
    // block-diagonal main product on 32-wide tiles: a group is one tile (the rank-r products dt_g = dy_g B_g) or several (the
    // per-head projections of the absorbed cross-attention)
    if (A2 || epi.drop_bits || a1g % 32 != 0 || N % a1g != 0) {
      set_error("ns_gemm_nt: a_group_cols supports groups of a multiple of 32 columns of a single product, N a multiple of the group");
      return NS_ERR_UNSUPPORTED;
    }
    bn = 32; cg = 1;
  }
  const bool am = epi.drop_bits && (epi.drop_mode == 1 || epi.drop_mode == 2);
  if (am && epi.drop_mode == 2 && (!epi.drop_seed || N > 128 || epi.drop_p < 0.f || epi.drop_p >= 1.f)) {
    set_error("ns_gemm_nt: drop_mode 2 needs drop_seed, drop_salts, 0 <= drop_p < 1 and at most 4 stacked adapters");
    return NS_ERR_ARG;
  }
  if (am) {
    // A-operand mask (LoRA down product): 32-wide tiles, one adapter per column tile
    if (A2 || N % 32 != 0 || K % 64 != 0 || epi.drop_ld % 2 != 0 || (reinterpret_cast<uintptr_t>(epi.drop_bits) & 7) != 0 ||
        (N > 32 && epi.drop_gstride % 2 != 0)) {
      set_error("ns_gemm_nt: the dropout-masked A operand needs a single product, N %% 32 == 0 (rank-32 adapters), K %% 64 == 0, even drop_ld");
      return NS_ERR_UNSUPPORTED;
    }
    bn = 32; cg = 1;
  } else if (epi.drop_bits) {
    // masked second product: 128-wide tiles (TMEM holds main + product tiles twice), CTA pairs whenever there is work for them
    if (!A2 || a2_ngrp > 0 || N % 64 != 0 || epi.drop_ld % 2 != 0 || epi.out_f32 || epi.residual || epi.act == NS_ACT_GELU ||
        (reinterpret_cast<uintptr_t>(epi.drop_bits) & 7) != 0) {
      set_error("ns_gemm_nt: drop_bits needs a second product (one adapter), N %% 64 == 0, even drop_ld, bf16 output, no residual / GELU");
      return NS_ERR_UNSUPPORTED;
    }
    bn = 128;
    const long long units = (((M + kBM - 1) / kBM + 1) / 2) * ((N + bn - 1) / bn);
    static const bool no2 = getenv("NS_GEMM_NO_2CTA") != nullptr;
    cg = (!no2 && units * 2 >= static_cast<long long>(sm_count()) * 3 / 4) ? 2 : 1;
  }
  {
    uint64_t dims[4] = {(uint64_t)(a1g > 0 ? static_cast<long long>(N / a1g) * K : K), 1, (uint64_t)M, 1};
    uint64_t str[3] = {(uint64_t)lda * 2, (uint64_t)lda * 2, (uint64_t)lda * 2 * (uint64_t)M};
    uint32_t box[4] = {kBK, 1, kBM, 1};
    int r = make_map(&maps.a[0], A, 4, dims, str, box);
    if (r) return r;
    uint64_t dimb[3] = {(uint64_t)K, (uint64_t)N, 1};
    uint64_t strb[2] = {(uint64_t)ldw * 2, (uint64_t)ldw * 2 * (uint64_t)N};
    uint32_t boxb[3] = {kBK, (uint32_t)(bn / cg), 1};
    r = make_map(&maps.b[0], W, 3, dimb, strb, boxb);
    if (r) return r;
  }
  prog.nseg = 1;
  seg_from_k(prog.seg[0], K);
  if (a1g > 0) { prog.seg[0].a_ngrp = a1g; prog.seg[0].a_kstep = K; }
  if (A2) {
    // the LoRA operand may be a column window of a wider stacked buffer: expose the whole row (lda2 columns)
    uint64_t dims[4] = {(uint64_t)(a2_ngrp > 0 ? lda2 : K2), 1, (uint64_t)M, 1};
    uint64_t str[3] = {(uint64_t)lda2 * 2, (uint64_t)lda2 * 2, (uint64_t)lda2 * 2 * (uint64_t)M};
    uint32_t box[4] = {kBK, 1, kBM, 1};
    int r = make_map(&maps.a[1], A2, 4, dims, str, box);
    if (r) return r;
    uint64_t dimb[3] = {(uint64_t)K2, (uint64_t)N, 1};
    uint64_t strb[2] = {(uint64_t)ldw2 * 2, (uint64_t)ldw2 * 2 * (uint64_t)N};
    uint32_t boxb[3] = {kBK, (uint32_t)(bn / cg), 1};
    r = make_map(&maps.b[1], W2, 3, dimb, strb, boxb);
    if (r) return r;
    prog.nseg = 2;
    seg_from_k(prog.seg[1], K2);
    prog.seg[1].a_map = 1;
    prog.seg[1].b_map = 1;
    prog.seg[1].a_ngrp = a2_ngrp;
    prog.seg[1].a_kstep = K2;
  } else {
    maps.a[1] = maps.a[0];
    maps.b[1] = maps.b[0];
  }
  prog.batches = 1;
  prog.tout = static_cast<int>(M);
  prog.tiles_per_batch = static_cast<int>((M + kBM - 1) / kBM);
  prog.N = N;
  prog.out_bs = 0; prog.out_rs = 1; prog.out_off = 0;
  prog.ldd = ldd;
  prog.D = D;
  fill_epi(prog, epi);
  if (am) {
    prog.am_bits = epi.drop_bits; prog.am_ld = epi.drop_ld; prog.am_gstride = epi.drop_gstride;
    if (epi.drop_mode == 2) {
      prog.am_seed = epi.drop_seed; prog.am_thr = drop_thr16(epi.drop_p);
      for (int i = 0; i < 4; ++i) prog.am_salts[i] = epi.drop_salts[i];
    }
    prog.epi.drop_bits = nullptr;
  }
  if (int r = setup_out_maps(maps, prog, bn)) return r;
  return dispatch_nt(maps, prog, st, bn, cg);
}

// Implicit-GEMM k=3 convolution forward on channels-last bf16 (see ns_conv3_fwd).
int conv3_fwd_fast(int B, int Tin, int Cp, int N, int stride, const void* x, const void* w, void* y, const EpiDev& epi,
                   cudaStream_t st) {
  if (Cp % 16 != 0 || !aligned16(x) || !aligned16(w) || (stride != 1 && stride != 2) || Tin % stride != 0)
    return NS_ERR_UNSUPPORTED;
  Maps maps;
  TileProg prog;
  memset(&prog, 0, sizeof(prog));
  const int Tout = Tin / stride;
  const int bn = choose_bn(static_cast<long long>(B) * ((Tout + kBM - 1) / kBM), N);
  const int cg = choose_cg(static_cast<long long>(B) * ((Tout + kBM - 1) / kBM), N, bn);
  uint64_t dims[4] = {(uint64_t)Cp, (uint64_t)stride, (uint64_t)Tout, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)Cp * 2, (uint64_t)Cp * 2 * stride, (uint64_t)Cp * 2 * (uint64_t)Tin};
  uint32_t box[4] = {kBK, 1, kBM, 1};
  int r = make_map(&maps.a[0], x, 4, dims, str, box);
  if (r) return r;
  uint64_t dimb[3] = {(uint64_t)Cp, (uint64_t)N, 3};
  uint64_t strb[2] = {(uint64_t)Cp * 2, (uint64_t)Cp * 2 * (uint64_t)N};
  uint32_t boxb[3] = {kBK, (uint32_t)(bn / cg), 1};
  r = make_map(&maps.b[0], w, 3, dimb, strb, boxb);
  if (r) return r;
  maps.a[1] = maps.a[0];
  maps.b[1] = maps.b[0];
  prog.nseg = 3;
  for (int k = 0; k < 3; ++k) {
    seg_from_k(prog.seg[k], Cp);
    prog.seg[k].b_tap = k;
    if (stride == 1) {
      prog.seg[k].a_off = k - 1;
      prog.seg[k].a_par = 0;
    } else {  // input row 2t+k-1:  k=0 -> (t-1, parity 1), k=1 -> (t, 0), k=2 -> (t, 1)
      prog.seg[k].a_off = (k == 0) ? -1 : 0;
      prog.seg[k].a_par = (k == 1) ? 0 : 1;
    }
  }
  prog.batches = B;
  prog.tout = Tout;
  prog.tiles_per_batch = (Tout + kBM - 1) / kBM;
  prog.N = N;
  prog.out_bs = Tout; prog.out_rs = 1; prog.out_off = 0;
  prog.ldd = N;
  prog.D = y;
  fill_epi(prog, epi);
  prog.epi.drop_bits = nullptr;                           // convolutions carry no LoRA branch
  if (int r = setup_out_maps(maps, prog, bn)) return r;
  return dispatch_nt(maps, prog, st, bn, cg);
}

// Input gradient of the stride-2 conv: one launch per output-row parity (see ns_conv3_dgrad).
int conv3_dgrad_fast(int B, int Tin, int Cp, int N, int stride, const void* dz, const void* wt, void* dx,
                     const EpiDev& epi, cudaStream_t st) {
  if (stride != 2 || N % 16 != 0 || Cp % 8 != 0 || !aligned16(dz) || !aligned16(wt) || Tin % 2 != 0)
    return NS_ERR_UNSUPPORTED;
  const int Tout = Tin / 2;
  const int bn = choose_bn(static_cast<long long>(B) * ((Tout + kBM - 1) / kBM), Cp);
  const int cg = choose_cg(static_cast<long long>(B) * ((Tout + kBM - 1) / kBM), Cp, bn);
  Maps maps;
  uint64_t dims[4] = {(uint64_t)N, 1, (uint64_t)Tout, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)N * 2, (uint64_t)N * 2, (uint64_t)N * 2 * (uint64_t)Tout};
  uint32_t box[4] = {kBK, 1, kBM, 1};
  int r = make_map(&maps.a[0], dz, 4, dims, str, box);
  if (r) return r;
  uint64_t dimb[3] = {(uint64_t)N, (uint64_t)Cp, 3};
  uint64_t strb[2] = {(uint64_t)N * 2, (uint64_t)N * 2 * (uint64_t)Cp};
  uint32_t boxb[3] = {kBK, (uint32_t)(bn / cg), 1};
  r = make_map(&maps.b[0], wt, 3, dimb, strb, boxb);
  if (r) return r;
  maps.a[1] = maps.a[0];
  maps.b[1] = maps.b[0];
  for (int par = 0; par < 2; ++par) {
    TileProg prog;
    memset(&prog, 0, sizeof(prog));
    if (par == 0) {          // dx[2j]   = dz[j] W1
      prog.nseg = 1;
      seg_from_k(prog.seg[0], N);
      prog.seg[0].b_tap = 1;
    } else {                 // dx[2j+1] = dz[j] W2 + dz[j+1] W0
      prog.nseg = 2;
      seg_from_k(prog.seg[0], N); prog.seg[0].b_tap = 2; prog.seg[0].a_off = 0;
      seg_from_k(prog.seg[1], N); prog.seg[1].b_tap = 0; prog.seg[1].a_off = 1;
    }
    prog.batches = B;
    prog.tout = Tout;
    prog.tiles_per_batch = (Tout + kBM - 1) / kBM;
    prog.N = Cp;
    prog.out_bs = Tin; prog.out_rs = 2; prog.out_off = par;
    prog.ldd = Cp;
    prog.D = dx;
    fill_epi(prog, epi);
    prog.epi.drop_bits = nullptr;
    if ((r = setup_out_maps(maps, prog, bn))) return r;
    r = dispatch_nt(maps, prog, st, bn, cg);
    if (r) return r;
  }
  return NS_OK;
}

static int launch_tn(const TnMaps& maps, TnProg& p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    NS_CUDA(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTnSmemMax));
    attr_done = true;
  }
  p.blocks_per_batch = (p.tout + 63) / 64;
  p.total_blocks = p.batches * p.blocks_per_batch;
  const int jt = p.J < 256 ? p.J : 256;   // all j tiles use the same UMMA N (tail columns masked in the epilogue)
  p.bj = (jt + 15) / 16 * 16;
  // narrow Y (the LoRA rank): one CTA takes up to four row tiles, i.e. up to 1 KB of every X row it streams
  static const bool no_wide = getenv("NS_TN_NO_WIDE") != nullptr;
  p.icta = 1;
  if (!no_wide && p.ntaps == 1 && p.bj <= 64) {
    const int it = (p.I + 127) / 128;
    p.icta = it >= 4 ? 4 : it;
  }
  while (p.grp_i > 0 && p.icta > 1 && p.grp_i % (128 * p.icta) != 0) --p.icta;   // an i tile never straddles two groups
  p.i_tiles = (p.I + 128 * p.icta - 1) / (128 * p.icta);
  p.j_tiles = (p.J + 255) / 256;
  p.stage_bytes = p.icta * kTnABytes + ((p.bj + 63) / 64) * 8192;
  const int tiles = p.i_tiles * p.j_tiles * p.ntaps;
  int nsplit;
  if (p.icta == 1) {                                         // 4 stages of <= 48 KB; two CTAs per SM when they fit
    p.stages = kTnMaxStages;
    nsplit = (2 * sm_count() + tiles - 1) / tiles;
  } else {                                                   // wide stages (up to 72 KB): one CTA per SM, 3 stages
    p.stages = (kTnSmemMax - 1024 - 256) / p.stage_bytes;
    if (p.stages > kTnMaxStages) p.stages = kTnMaxStages;
    nsplit = (sm_count() + tiles - 1) / tiles;
  }
  const int max_split = (p.total_blocks + 3) / 4;           // at least 4 contraction blocks per CTA
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  p.blocks_per_split = (p.total_blocks + nsplit - 1) / nsplit;
  p.nsplit = (p.total_blocks + p.blocks_per_split - 1) / p.blocks_per_split;
  const int smem = p.stages * p.stage_bytes + 1024 + 256;
  gemm_tn_kernel<<<tiles * p.nsplit, kTnThreads, smem, st>>>(maps, p);
  NS_LAUNCH_CHECK();
  count(C_WGRAD_TC);
  return NS_OK;
}

int gemm_tn_fast(long long M, int I, int J, const void* X, long long ldx, const void* Y, long long ldy, float* G,
                 long long si, long long sj, float alpha, cudaStream_t st, const uint32_t* xbits, long long xbits_ld) {
  if (ldx % 8 != 0 || ldy % 8 != 0 || !aligned16(X) || !aligned16(Y) || M > 0x7fffffffLL) return NS_ERR_UNSUPPORTED;
  if (I % 8 != 0 || J % 8 != 0) return NS_ERR_UNSUPPORTED;
  if (xbits && (I % 64 != 0 || xbits_ld % 2 != 0 || (reinterpret_cast<uintptr_t>(xbits) & 7) != 0)) {
    set_error("ns_gemm_tn_masked: I must be a multiple of 64, xbits_ld even, xbits 8-byte aligned");
    return NS_ERR_UNSUPPORTED;
  }
  // G^T = Y^T X is the same sum: put the WIDE operand on the X side (rows of the MMA, several row tiles per CTA) and the
  // LoRA-rank-wide one on the Y side
  if (!xbits && I <= 64 && J >= 128) {
    std::swap(I, J); std::swap(X, Y); std::swap(ldx, ldy); std::swap(si, sj);
  }
  TnMaps maps;
  uint64_t dx[4] = {(uint64_t)I, 1, (uint64_t)M, 1};
  uint64_t sx[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2, (uint64_t)ldx * 2 * (uint64_t)M};
  uint32_t box[4] = {64, 1, 64, 1};
  int r = make_map(&maps.x, X, 4, dx, sx, box);
  if (r) return r;
  uint64_t dy[4] = {(uint64_t)J, 1, (uint64_t)M, 1};
  uint64_t sy[3] = {(uint64_t)ldy * 2, (uint64_t)ldy * 2, (uint64_t)ldy * 2 * (uint64_t)M};
  r = make_map(&maps.y, Y, 4, dy, sy, box);
  if (r) return r;
  TnProg p;
  memset(&p, 0, sizeof(p));
  p.batches = 1; p.tout = static_cast<int>(M); p.ntaps = 1;
  p.I = I; p.J = J; p.si = si; p.sj = sj; p.stap = 0; p.G = G; p.alpha = alpha;
  p.xbits = xbits; p.xbits_ld = xbits_ld;
  return launch_tn(maps, p, st);
}

// Block-diagonal weight gradients (see ns_gemm_tn_grouped): X (M, groups*I), Y (M, groups*J), group g -> G rows [g*I, (g+1)*I)
int gemm_tn_grouped_fast(long long M, int I, int J, int groups, const void* X, long long ldx, const void* Y, long long ldy, float* G,
                         long long si, long long sj, const float* alphas, cudaStream_t st) {
  if (ldx % 8 != 0 || ldy % 8 != 0 || !aligned16(X) || !aligned16(Y) || M > 0x7fffffffLL) return NS_ERR_UNSUPPORTED;
  if (I % 128 != 0 || J % 16 != 0 || J > 64 || groups < 1 || groups > 4) return NS_ERR_UNSUPPORTED;
  TnMaps maps;
  uint64_t dx[4] = {(uint64_t)I * groups, 1, (uint64_t)M, 1};
  uint64_t sx[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2, (uint64_t)ldx * 2 * (uint64_t)M};
  uint32_t box[4] = {64, 1, 64, 1};
  int r = make_map(&maps.x, X, 4, dx, sx, box);
  if (r) return r;
  uint64_t dy[4] = {(uint64_t)J * groups, 1, (uint64_t)M, 1};
  uint64_t sy[3] = {(uint64_t)ldy * 2, (uint64_t)ldy * 2, (uint64_t)ldy * 2 * (uint64_t)M};
  r = make_map(&maps.y, Y, 4, dy, sy, box);
  if (r) return r;
  TnProg p;
  memset(&p, 0, sizeof(p));
  p.batches = 1; p.tout = static_cast<int>(M); p.ntaps = 1;
  p.I = I * groups; p.J = J; p.si = si; p.sj = sj; p.stap = 0; p.G = G; p.alpha = 1.0f;
  p.grp_i = I; p.grp_j = J;
  for (int g = 0; g < groups; ++g) p.alpha_grp[g] = alphas[g];
  return launch_tn(maps, p, st);
}

// dw (3, N, Cp) += sum_{b,t} dz[b,t,n] x[b, stride*t + k - 1, c]
int conv3_wgrad_fast(int B, int Tin, int Cp, int N, int stride, const void* dz, const void* x, float* dw,
                     cudaStream_t st) {
  if (Cp % 8 != 0 || N % 8 != 0 || !aligned16(dz) || !aligned16(x) || (stride != 1 && stride != 2) || Tin % stride != 0)
    return NS_ERR_UNSUPPORTED;
  const int Tout = Tin / stride;
  TnMaps maps;
  uint64_t dxm[4] = {(uint64_t)N, 1, (uint64_t)Tout, (uint64_t)B};
  uint64_t sxm[3] = {(uint64_t)N * 2, (uint64_t)N * 2, (uint64_t)N * 2 * (uint64_t)Tout};
  uint32_t box[4] = {64, 1, 64, 1};
  int r = make_map(&maps.x, dz, 4, dxm, sxm, box);
  if (r) return r;
  uint64_t dym[4] = {(uint64_t)Cp, (uint64_t)stride, (uint64_t)Tout, (uint64_t)B};
  uint64_t sym[3] = {(uint64_t)Cp * 2, (uint64_t)Cp * 2 * stride, (uint64_t)Cp * 2 * (uint64_t)Tin};
  r = make_map(&maps.y, x, 4, dym, sym, box);
  if (r) return r;
  TnProg p;
  memset(&p, 0, sizeof(p));
  p.batches = B; p.tout = Tout; p.ntaps = 3;
  p.I = N; p.J = Cp; p.si = Cp; p.sj = 1; p.stap = static_cast<long long>(N) * Cp; p.G = dw; p.alpha = 1.0f;
  for (int k = 0; k < 3; ++k) {
    if (stride == 1) { p.y_off[k] = k - 1; p.y_par[k] = 0; }
    else { p.y_off[k] = (k == 0) ? -1 : 0; p.y_par[k] = (k == 1) ? 0 : 1; }
  }
  return launch_tn(maps, p, st);
}

}  // namespace ns
