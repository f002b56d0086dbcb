// Throughput of the legacy tensor path on sm_100a: mma.sync.m16n8k16 bf16, W warps per SM, A independent accumulators per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/hmma_bench tools/ubench/hmma_bench.cu ; prints cycles per MMA per SM sub-core.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int ACC>
__global__ void k(int iters, float* out, long long* cyc) {
  float c[ACC][4];
  for (int a = 0; a < ACC; ++a) for (int i = 0; i < 4; ++i) c[a][i] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 11, b1 = 13;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int a = 0; a < ACC; ++a)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0; for (int a = 0; a < ACC; ++a) for (int i = 0; i < 4; ++i) s += c[a][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16, 32}) {
    for (int acc : {1, 4, 8}) {
      if (acc == 1) k<1><<<148, warps * 32>>>(iters, out, cyc);
      if (acc == 4) k<4><<<148, warps * 32>>>(iters, out, cyc);
      if (acc == 8) k<8><<<148, warps * 32>>>(iters, out, cyc);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double per_sm = (double)h[0] / ((double)iters * acc * warps);     // cycles per MMA per SM
      printf("warps/SM %2d  indep acc %d : %.2f cycles/MMA/SM  (%.1f cyc per sub-core MMA; %.0f TFLOP/s chip at 1.9 GHz)\n", warps, acc, per_sm,
             per_sm * 4, 4096.0 / per_sm * 148 * 1.9e9 / 1e12);
    }
  }
  return 0;
}
