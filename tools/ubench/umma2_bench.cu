// Micro-benchmark: tcgen05.mma.cta_group::2 (M=256 over a CTA pair) issue/execution rate, shared-memory operands.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
namespace ns { void set_error(const char*, ...) {} }
#include "ns_sm100.cuh"
using namespace ns::sm100;

template <int N>
__global__ void __launch_bounds__(128, 1) bench2(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 160 * 1024;
  const uint32_t slot = bar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc2(slot, 512); tmem_relinquish2(); }
  tc_fence_before(); cluster_sync_all(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (warp == 1 && rank == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(256, N, 0, 0);
    const uint64_t ad = umma_smem_desc(base, 16, 1024);
    const uint64_t bd = umma_smem_desc(base + 65536, 16, 1024);
    long long t0 = 0, t1 = 0;
    uint32_t ph = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_f16_cg2(tmem, ad + 2u * (k & 3), bd + 2u * (k & 3), idesc, 1);
        }
        umma_commit_mc2(bar, 1);
      }
      __syncwarp();
      mbar_wait(bar, ph); ph ^= 1;
      t1 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(tmem, 512); }
}

template <int N> void run(long long* d) {
  auto k = bench2<N>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 170 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const int iters = 512;
  cudaLaunchKernelEx(&cfg, k, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("cta_group::2 SS M=256 N=%3d : %7.1f cycles/MMA (ideal %d)  [%s]\n", N, (double)c / (iters * 8), N / 2, cudaGetErrorString(e));
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  run<256>(d); run<128>(d); run<64>(d);
  return 0;
}
