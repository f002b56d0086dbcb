// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM for W warps, alone and under a concurrent tcgen05.mma stream.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
namespace ns { void set_error(const char*, ...) {} }
#include "ns_sm100.cuh"
using namespace ns::sm100;

// mode 0: ld.x32 + wait each; 1: st.x32 + wait each; 2: ld.x32 (x2 in flight) ; 3: ld x16 packed + st x16 (softmax-like, no math)
template <int MODE, int MMA>
__global__ void __launch_bounds__(1024, 1) bench(int nwarps, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 160 * 1024;
  const uint32_t slot = bar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  __shared__ long long tmax[32];
  long long dt = 0;
  if (warp == 31) {
    if (MMA) {
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
        const uint64_t bd = umma_smem_desc(base + 65536, 16, 1024);
        const long long t0 = clock64();
        for (int it = 0; it < iters * 4; ++it) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_f16_ts(tmem + 384u, tmem + 448u + 8u * (k & 3), bd + 2u * (k & 3), idesc, 1);
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        dt = clock64() - t0;
        if (blockIdx.x == 0) out[1] = dt;
      }
      __syncwarp();
    }
  } else if (warp < nwarps) {
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t col = 32u * ((warp >> 2) & 7);
    uint32_t v[32], w[32];
    uint32_t accu = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i + lane;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) { tmem_ld32(tmem + lane_addr + col, v); tmem_ld_wait(); accu += v[it & 31]; }
      if (MODE == 1) { v[0] = accu + it; tmem_st32(tmem + lane_addr + col, v); tmem_st_wait(); }
      if (MODE == 2) { tmem_ld32(tmem + lane_addr + col, v); tmem_ld32(tmem + lane_addr + ((col + 32u) & 255u), w); tmem_ld_wait(); accu += v[it & 31] + w[it & 31]; }
      if (MODE == 3) {
        tmem_ld32(tmem + lane_addr + col, v); tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = v[2 * i] ^ v[2 * i + 1];
        tmem_st16(tmem + lane_addr + col, pk); tmem_st_wait();
      }
    }
    dt = clock64() - t0;
    if (accu == 0x12345) out[3] = accu;
  }
  if (lane == 0) tmax[warp] = (warp < nwarps) ? dt : 0;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long m = 0;
    for (int i = 0; i < nwarps && i < 31; ++i) m = tmax[i] > m ? tmax[i] : m;
    out[0] = m;
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE, int MMA>
void run(long long* d, int nwarps) {
  auto k = bench<MODE, MMA>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  const int iters = 2048;
  cudaMemset(d, 0, 32);
  k<<<148, 1024, 170 * 1024>>>(nwarps, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long r[4] = {0, 0, 0, 0}; cudaMemcpy(r, d, 32, cudaMemcpyDeviceToHost);
  const double bytes_per_it = (MODE == 2 ? 2.0 : 1.0) * 4096.0 * nwarps + (MODE == 3 ? 2048.0 * nwarps : 0.0);
  printf("mode %d mma %d warps %2d : %7.1f cycles/iter/warp-slot, %7.1f B/cycle/SM", MODE, MMA, nwarps, (double)r[0] / iters, bytes_per_it * iters / (double)r[0]);
  if (MMA) printf(", MMA %6.1f cycles each (ideal 32)", (double)r[1] / (iters * 4 * 8));
  printf("  [%s]\n", cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 32);
  for (int nw : {4, 8, 16, 24}) run<0, 0>(d, nw);
  for (int nw : {4, 8, 16}) run<1, 0>(d, nw);
  for (int nw : {4, 8, 16}) run<2, 0>(d, nw);
  for (int nw : {4, 8, 16}) run<3, 0>(d, nw);
  for (int nw : {4, 8, 16}) run<0, 1>(d, nw);
  for (int nw : {8, 16}) run<3, 1>(d, nw);
  return 0;
}
