// Plain SIMT (CUDA-core, fp32 accumulate) GEMM family.  Used for NS_F32 storage (the fp32 parity mode) and for shapes the
// tcgen05 path does not take.  Same "segment" formulation as the tcgen05 kernels: a row of the A operand of segment s for
// output row (b, t) is   A_s[b*a_bs + (t + a_off)*a_rs + a_add, :]   (zero when t + a_off is outside [0, a_rows)).
#include "ns_common.cuh"
#include "ns_gemm.cuh"

namespace ns {

template <typename T>
__global__ void __launch_bounds__(256) gemm_nt_simt_kernel(const SimtProg p) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 1];
  __shared__ float Ws[TK][TN + 1];
  const int tiles_per_batch = (p.tout + TM - 1) / TM;
  const int b = blockIdx.y / tiles_per_batch;
  const int t0 = (blockIdx.y % tiles_per_batch) * TM;
  const int n0 = blockIdx.x * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int s = 0; s < p.nseg; ++s) {
    const SimtSeg sg = p.seg[s];
    const T* A = reinterpret_cast<const T*>(sg.A);
    const T* W = reinterpret_cast<const T*>(sg.W);
    const int kbase = sg.a_ngrp > 0 ? (n0 / sg.a_ngrp) * sg.a_kstep : 0;
    for (int k0 = 0; k0 < sg.K; k0 += TK) {
      // load A tile: 64 rows x 16 k  (256 threads: 4 elements each)
      for (int e = threadIdx.x; e < TM * TK; e += 256) {
        const int r = e / TK, kk = e % TK;
        const int t = t0 + r;
        const int ar = t + sg.a_off;
        float v = 0.f;
        if (t < p.tout && ar >= 0 && ar < sg.a_rows && k0 + kk < sg.K)
          v = to_f<T>(A[(static_cast<long long>(b) * sg.a_bs + static_cast<long long>(ar) * sg.a_rs + sg.a_add) * sg.lda + kbase + k0 + kk]);
        As[kk][r] = v;
      }
      for (int e = threadIdx.x; e < TN * TK; e += 256) {
        const int r = e / TK, kk = e % TK;
        const int n = n0 + r;
        float v = 0.f;
        if (n < p.N && k0 + kk < sg.K) v = to_f<T>(W[static_cast<long long>(n) * sg.ldw + k0 + kk]);
        Ws[kk][r] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float a[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  const EpiDev& e = p.epi;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= p.tout) continue;
    const long long row = static_cast<long long>(b) * p.out_bs + static_cast<long long>(t) * p.out_rs + p.out_off;
    const long long res_row = e.res_mod > 0 ? ((static_cast<long long>(t) * p.out_rs + p.out_off) % e.res_mod) : row;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= p.N) continue;
      const float x = epi_apply<T>(e, acc[i][j], row, res_row, col);
      if (e.out_f32) reinterpret_cast<float*>(p.D)[row * p.ldd + col] = x;
      else reinterpret_cast<T*>(p.D)[row * p.ldd + col] = from_f<T>(x);
    }
  }
}

int launch_nt_simt(int dtype, const SimtProg& p, cudaStream_t st) {
  if (p.batches <= 0 || p.tout <= 0 || p.N <= 0) return NS_OK;
  dim3 grid((p.N + 63) / 64, static_cast<unsigned>(p.batches) * ((p.tout + 63) / 64));
  if (grid.y > 65535u * 16u) { set_error("SIMT GEMM: too many row tiles"); return NS_ERR_UNSUPPORTED; }
  // grid.y limit is 65535: fold into x if needed
  if (grid.y > 65535u) { set_error("SIMT GEMM: M too large (%u row tiles)", grid.y); return NS_ERR_UNSUPPORTED; }
  if (dtype == NS_F32) gemm_nt_simt_kernel<float><<<grid, 256, 0, st>>>(p);
  else gemm_nt_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  NS_LAUNCH_CHECK();
  count(C_GEMM_SIMT);
  return NS_OK;
}

// ---------------------------------------------------------------------------------------------- TN (wgrad)
template <typename T>
__global__ void __launch_bounds__(256) gemm_tn_simt_kernel(const SimtTnProg p) {
  constexpr int TI = 64, TJ = 64, TK = 16;
  __shared__ float Xs[TK][TI + 1];
  __shared__ float Ys[TK][TJ + 1];
  const int i0 = blockIdx.x * TI;
  const int j_tiles = (p.J + TJ - 1) / TJ;
  const int j0 = (blockIdx.y % j_tiles) * TJ;
  const int tap = blockIdx.y / j_tiles;
  const long long total = static_cast<long long>(p.batches) * p.tout;
  const long long m_begin = static_cast<long long>(blockIdx.z) * p.chunk;
  const long long m_end = min(total, m_begin + p.chunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const T* X = reinterpret_cast<const T*>(p.X);
  const T* Y = reinterpret_cast<const T*>(p.Y);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long m0 = m_begin; m0 < m_end; m0 += TK) {
    for (int e = threadIdx.x; e < TK * TI; e += 256) {
      const int kk = e / TI, r = e % TI;
      const long long m = m0 + kk;
      float v = 0.f;
      if (m < m_end && i0 + r < p.I) {
        const long long b = m / p.tout, t = m % p.tout;
        v = to_f<T>(X[(b * p.x_bs + t) * p.ldx + i0 + r]);
      }
      Xs[kk][r] = v;
    }
    for (int e = threadIdx.x; e < TK * TJ; e += 256) {
      const int kk = e / TJ, r = e % TJ;
      const long long m = m0 + kk;
      float v = 0.f;
      if (m < m_end && j0 + r < p.J) {
        const long long b = m / p.tout, t = m % p.tout;
        const long long yr = t + p.y_off[tap];
        if (yr >= 0 && yr < p.y_rows) v = to_f<T>(Y[(b * p.y_bs + yr * p.y_rs + p.y_add[tap]) * p.ldy + j0 + r]);
      }
      Ys[kk][r] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Xs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ys[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = i0 + ty * 4 + i;
    if (ii >= p.I) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = j0 + tx * 4 + j;
      if (jj >= p.J) continue;
      atomicAdd(p.G + tap * p.stap + ii * p.si + jj * p.sj, p.alpha * acc[i][j]);
    }
  }
}

int launch_tn_simt(int dtype, SimtTnProg& p, cudaStream_t st) {
  if (p.I <= 0 || p.J <= 0 || p.batches <= 0 || p.tout <= 0) return NS_OK;
  const long long total = static_cast<long long>(p.batches) * p.tout;
  const int tiles = ((p.I + 63) / 64) * ((p.J + 63) / 64) * p.ntaps;
  long long nsplit = (4LL * sm_count() + tiles - 1) / tiles;
  const long long max_split = (total + 255) / 256;
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 65535) nsplit = 65535;
  long long chunk = (total + nsplit - 1) / nsplit;
  chunk = (chunk + 15) / 16 * 16;
  p.chunk = static_cast<int>(chunk);
  nsplit = (total + chunk - 1) / chunk;
  dim3 grid((p.I + 63) / 64, ((p.J + 63) / 64) * p.ntaps, static_cast<unsigned>(nsplit));
  if (dtype == NS_F32) gemm_tn_simt_kernel<float><<<grid, 256, 0, st>>>(p);
  else gemm_tn_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  NS_LAUNCH_CHECK();
  count(C_GEMM_SIMT);
  return NS_OK;
}

// column sums: out[n] += sum_rows x[r, n]   (conv bias gradient)
template <typename T>
__global__ void colsum_kernel(long long rows, int N, const T* x, long long ld, float* out, int rows_per_block) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += to_f<T>(x[r * ld + n]);
  atomicAdd(out + n, s);
}

int launch_colsum(int dtype, long long rows, int N, const void* x, long long ld, float* out, cudaStream_t st) {
  if (rows <= 0 || N <= 0) return NS_OK;
  int rpb = 256;
  long long by = (rows + rpb - 1) / rpb;
  while (by > 65535) { rpb *= 2; by = (rows + rpb - 1) / rpb; }
  dim3 grid((N + 127) / 128, static_cast<unsigned>(by));
  if (dtype == NS_F32) colsum_kernel<float><<<grid, 128, 0, st>>>(rows, N, reinterpret_cast<const float*>(x), ld, out, rpb);
  else colsum_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(rows, N, reinterpret_cast<const __nv_bfloat16*>(x), ld, out, rpb);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

}  // namespace ns
