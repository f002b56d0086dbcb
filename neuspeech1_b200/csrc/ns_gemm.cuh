// Shared declarations of the GEMM family (tcgen05 fast path + SIMT path).
#pragma once
#include "ns_common.cuh"

namespace ns {

struct SimtSeg {
  const void* A;
  const void* W;
  long long lda, ldw;
  long long a_bs, a_rs, a_add;   // element offsets in units of rows (multiplied by lda)
  int a_off, a_rows;
  int K;
  int a_ngrp, a_kstep;
};

struct SimtProg {
  int nseg;
  SimtSeg seg[4];
  int batches, tout, N;
  long long out_bs, out_rs, out_off, ldd;
  void* D;
  EpiDev epi;
};

struct SimtTnProg {
  const void* X; const void* Y;
  long long ldx, ldy;
  int batches, tout;
  long long x_bs;                         // X row = b*x_bs + t
  long long y_bs, y_rs, y_add[3];         // Y row = b*y_bs + (t + y_off)*y_rs + y_add
  int y_off[3], y_rows;
  int ntaps, I, J;
  long long si, sj, stap;
  float* G;
  float alpha;
  int chunk;                              // contraction rows per CTA
};

// implemented in ns_gemm_sm100.cu / ns_gemm_simt.cu
int gemm_nt_fast(long long M, int N, int K, const void* A, long long lda, const void* W, long long ldw, void* D,
                 long long ldd, const EpiDev& epi, const void* A2, long long lda2, const void* W2, long long ldw2,
                 int K2, int a2_ngrp, cudaStream_t st);
int conv3_fwd_fast(int B, int Tin, int Cp, int N, int stride, const void* x, const void* w, void* y, const EpiDev& epi,
                   cudaStream_t st);
int conv3_dgrad_fast(int B, int Tin, int Cp, int N, int stride, const void* dz, const void* wt, void* dx,
                     const EpiDev& epi, cudaStream_t st);
int gemm_tn_fast(long long M, int I, int J, const void* X, long long ldx, const void* Y, long long ldy, float* G,
                 long long si, long long sj, float alpha, cudaStream_t st, const uint32_t* xbits = nullptr, long long xbits_ld = 0);
int gemm_tn_grouped_fast(long long M, int I, int J, int groups, const void* X, long long ldx, const void* Y, long long ldy, float* G,
                         long long si, long long sj, const float* alphas, cudaStream_t st);
// ns_lora_bwd.cu: dt = alpha' dy B and dB += dy^T t in one pass over dy
int lora_bwd_b_fast(long long M, int N, int r, int groups, const void* dy, long long lddy, const void* Bt, long long ldbt, const void* t,
                    long long ldt, void* dt, long long lddt, float* dB, const float* alpha_dt, const float* alpha_db, void* workspace,
                    long long workspace_bytes, cudaStream_t st);
long long lora_bwd_b_workspace_bytes(long long M, int N, int r, int groups);
int conv3_wgrad_fast(int B, int Tin, int Cp, int N, int stride, const void* dz, const void* x, float* dw, cudaStream_t st);

// ns_skinny.cu: D = epi(LN(x) W^T) for M <= 128 rows (decoder step); NS_ERR_UNSUPPORTED when the shape does not qualify
int skinny_gemm(long long M, int N, int K, const void* x, long long ldx, const float* gamma, const float* beta, float eps,
                const void* w, long long ldw, void* d, long long ldd, const ns_epilogue* ep, cudaStream_t st);

int launch_nt_simt(int dtype, const SimtProg& p, cudaStream_t st);
int launch_tn_simt(int dtype, SimtTnProg& p, cudaStream_t st);
int launch_colsum(int dtype, long long rows, int N, const void* x, long long ld, float* out, cudaStream_t st);


}  // namespace ns
