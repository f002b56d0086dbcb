// Micro-benchmark: cycles per tcgen05.mma for several shapes / dependency patterns (one CTA per SM, 1 issuing thread).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neuspeech1_b200/csrc -o gpurun_out/umma_bench tools/ubench/umma_bench.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
namespace ns { void set_error(const char*, ...) {} }
#include "ns_sm100.cuh"
using namespace ns::sm100;

template <int N, int NACC, int TS, int MN>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 160 * 1024;
  const uint32_t slot = bar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, MN, MN);
    const uint64_t ad = MN ? umma_smem_desc(base, 16384, 1024) : umma_smem_desc(base, 16, 1024);
    const uint64_t bd = MN ? umma_smem_desc(base + 65536, 16384, 1024) : umma_smem_desc(base + 65536, 16, 1024);
    long long t0 = 0, t1 = 0, ti = 0;
    uint32_t ph = 0;
    for (int rep = 0; rep < 2; ++rep) {          // rep 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t d = tmem + 256u + static_cast<uint32_t>((k % NACC) * N);
            const uint32_t adv = MN ? 128u * (k & 3) : 2u * (k & 3);
            if (TS) umma_f16_ts(d, tmem + 8u * (k & 3), bd + adv, idesc, 1);
            else umma_f16(d, ad + adv, bd + adv, idesc, 1);
          }
        }
        umma_commit(bar);
      }
      __syncwarp();
      ti = clock64();
      mbar_wait(bar, ph); ph ^= 1;
      t1 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int NACC, int TS, int MN>
void run(long long* d) {
  auto k = bench<N, NACC, TS, MN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  const int iters = 512;
  k<<<148, 128, 170 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc[2] = {0, 0}; cudaMemcpy(cyc, d, 16, cudaMemcpyDeviceToHost);
  printf("%s %s N=%3d nacc=%d : %7.1f cycles/MMA total, %7.1f issue-side (ideal %d)  [%s]\n", TS ? "TS" : "SS", MN ? "MN-major" : "K-major ", N, NACC,
         (double)cyc[0] / (iters * 8), (double)cyc[1] / (iters * 8), N / 2, cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<64, 1, 0, 0>(d); run<64, 2, 0, 0>(d); run<128, 1, 0, 0>(d); run<128, 2, 0, 0>(d); run<256, 1, 0, 0>(d);
  run<64, 1, 0, 1>(d); run<128, 1, 0, 1>(d); run<256, 1, 0, 1>(d);
  run<64, 1, 1, 0>(d); run<64, 2, 1, 0>(d); run<128, 1, 1, 0>(d); run<256, 1, 1, 0>(d);
  run<32, 1, 0, 0>(d); run<32, 1, 1, 0>(d); run<16, 1, 1, 0>(d);
  return 0;
}
