// Micro-benchmark: how fast can 148 persistent CTAs stream a row-major bf16 matrix (M, N) from HBM through a TMA ring
// (128B-swizzled boxes, 64 columns wide), depending on the SHAPE and ORDER of the boxes of a ring stage?  No compute: one thread
// issues the loads, one waits for a full stage and hands it back.  The rank-r products of the LoRA branches (ns_gemm_nt 32-wide
// tiles, gemm_tn_kernel, lora_bwd_b_kernel) are exactly this stream plus a few MMAs, and they all sit at 3.3-4.1 TB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neuspeech1_b200/csrc -o gpurun_out/tma_stream_bench tools/ubench/tma_stream_bench.cu
// stage tile = SR rows x SC columns, loaded as boxes of BR rows x 64 columns, row groups outer, column boxes inner; tiles run
// column-tile fastest inside a slab of SR rows; every CTA takes a contiguous, balanced range of tiles.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
namespace ns { void set_error(const char*, ...) {} }
#include "ns_sm100.cuh"
using namespace ns::sm100;

struct Prog {
  int SR, SC, BR, stages, col_tiles;
  long long tiles;
  uint32_t stage_bytes;
};

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ Prog p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + p.stages * p.stage_bytes;
  auto full = [&](int s) { return bar + 8u * s; };
  auto empty = [&](int s) { return bar + 8u * (16 + s); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_fence_init();
  }
  __syncthreads();
  const long long t0 = p.tiles * blockIdx.x / gridDim.x, t1 = p.tiles * (blockIdx.x + 1) / gridDim.x;
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    for (long long t = t0; t < t1; ++t) {
      const int row0 = static_cast<int>(t / p.col_tiles) * p.SR, col0 = static_cast<int>(t % p.col_tiles) * p.SC;
      mbar_wait(empty(stage), phase ^ 1u);
      mbar_expect_tx(full(stage), p.stage_bytes);
      const uint32_t dst = base + stage * p.stage_bytes;
      // smem layout: [column box][SR rows][128 B]; a box of BR rows lands at its row offset inside its column box
      for (int rg = 0; rg < p.SR / p.BR; ++rg)
        for (int cb = 0; cb < p.SC / 64; ++cb)
          tma_load_3d(&map, full(stage), dst + static_cast<uint32_t>(cb * p.SR * 128 + rg * p.BR * 128), col0 + 64 * cb, row0 + rg * p.BR, 0);
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (long long t = t0; t < t1; ++t) {
      mbar_wait(full(stage), phase);
      mbar_arrive(empty(stage));
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(f);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  void* flush = nullptr;
  cudaMalloc(&flush, 256u << 20);
  struct Shape { long long M; int N; } shapes[] = {{96000, 512}, {384000, 512}, {96000, 2048}};
  struct Mode { int SR, SC, BR; const char* what; } modes[] = {
      {128, 128, 128, "128x128 tile, 128-row boxes (lora_bwd_b today)"},
      {128, 128, 32, "128x128 tile, 32-row boxes"},
      {128, 64, 128, "128x64 tile (gemm_nt k-block)"},
      {64, 256, 64, "64x256 tile (gemm_tn icta=2)"},
      {64, 512, 64, "64x512 tile (gemm_tn icta=4)"},
      {32, 512, 32, "32 full rows of 512 columns"},
      {16, 512, 16, "16 full rows of 512 columns"},
      {64, 512, 16, "64x512 tile, 16-row boxes"},
      {128, 256, 32, "128x256 tile, 32-row boxes"},
      {128, 256, 128, "128x256 tile, 128-row boxes"},
  };
  const int promos[] = {CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE};
  const char* pname[] = {"L2_256B", "L2_128B", "none"};
  for (auto sh : shapes) {
    __nv_bfloat16* x = nullptr;
    const size_t bytes = static_cast<size_t>(sh.M) * sh.N * 2;
    cudaMalloc(&x, bytes);
    cudaMemset(x, 1, bytes);
    for (int pi = 0; pi < 3; ++pi) {
      for (auto m : modes) {
        if (pi > 0 && !(m.SR == 128 && m.SC == 128 && m.BR == 128) && !(m.SR == 32 && m.SC == 512)) continue;
        if (sh.N % m.SC != 0) continue;
        Prog p;
        p.SR = m.SR; p.SC = m.SC; p.BR = m.BR;
        p.stage_bytes = static_cast<uint32_t>(m.SR) * m.SC * 2;
        p.stages = (200 * 1024) / p.stage_bytes;
        if (p.stages > 16) p.stages = 16;
        if (p.stages < 2) continue;
        p.col_tiles = sh.N / m.SC;
        p.tiles = (sh.M / m.SR) * p.col_tiles;
        CUtensorMap map;
        cuuint64_t gd[3] = {(cuuint64_t)sh.N, (cuuint64_t)sh.M, 1};
        cuuint64_t gs[2] = {(cuuint64_t)sh.N * 2, (cuuint64_t)sh.N * 2 * sh.M};
        cuuint32_t bx[3] = {64, (cuuint32_t)m.BR, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         (CUtensorMapL2promotion)promos[pi], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int smem = p.stages * p.stage_bytes + 1024 + 512;
        float best = 1e9f, sum = 0.f;
        const int reps = 6;
        for (int rep = 0; rep < reps; ++rep) {
          cudaMemsetAsync(flush, rep, 256u << 20);
          cudaEvent_t a, b;
          cudaEventCreate(&a); cudaEventCreate(&b);
          cudaEventRecord(a);
          stream_kernel<<<sms, 64, smem>>>(map, p);
          cudaEventRecord(b);
          cudaEventSynchronize(b);
          float ms = 0.f;
          cudaEventElapsedTime(&ms, a, b);
          if (rep > 0) { sum += ms; if (ms < best) best = ms; }
          cudaEventDestroy(a); cudaEventDestroy(b);
        }
        cudaError_t e = cudaGetLastError();
        const float avg = sum / (reps - 1);
        printf("M=%lld N=%d promo=%-7s %-48s stages=%2d  avg %.1f us  %.2f TB/s  (best %.1f us %.2f TB/s) %s\n", sh.M, sh.N, pname[pi], m.what, p.stages,
               avg * 1e3, bytes / (avg * 1e-3) / 1e12, best * 1e3, bytes / (best * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
    cudaFree(x);
  }
  return 0;
}
