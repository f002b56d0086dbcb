// SIMT (CUDA-core, fp32) multi-head attention forward / backward with online softmax; any Lq/Lk, head_dim 32 or 64,
// optional causal mask.  Used for NS_F32 storage (fp32 parity mode), for the small decoder attentions and as the checker of
// the tensor-core attention.  q is pre-scaled by the caller (HF modeling_whisper.py:310), so scores are plain q.k.
//
// forward : grid (ceil(Lq/16), H, B), 8 warps, each warp owns 2 queries; K/V tiles of 64 keys staged in shared memory.
// backward: delta = rowsum(dO*O); dq kernel (same structure as forward); dk/dv kernel (one warp per 2 keys, loops queries).
#include "ns_common.cuh"

#include <stdlib.h>

namespace ns {

constexpr int kQPB = 16;    // queries per block
constexpr int kKT = 64;     // keys per shared tile

template <typename T, int DH>
__global__ void __launch_bounds__(256) attn_fwd_simt_kernel(const ns_attn_shape s, const T* __restrict__ q, const T* __restrict__ k,
                                                            const T* __restrict__ v, T* __restrict__ o, float* __restrict__ lse) {
  constexpr int DPL = DH / 32;
  __shared__ float Ks[kKT][DH + 1];
  __shared__ float Vs[kKT][DH + 1];
  __shared__ float Qs[kQPB][DH];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQPB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int off = s.Lk - s.Lq;
  for (int e = threadIdx.x; e < kQPB * DH; e += 256) {
    const int r = e / DH, d = e % DH;
    const int qi = q0 + r;
    Qs[r][d] = qi < s.Lq ? to_f<T>(q[b * s.q_bs + static_cast<long long>(qi) * s.q_rs + h * DH + d]) : 0.f;
  }
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float acc[2][DPL];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[i][d] = 0.f;
  // keys needed by this block (causal): up to q0 + kQPB - 1 + off
  const int k_end = s.causal ? min(s.Lk, q0 + kQPB + off) : s.Lk;
  for (int k0 = 0; k0 < k_end; k0 += kKT) {
    __syncthreads();
    for (int e = threadIdx.x; e < kKT * DH; e += 256) {
      const int r = e / DH, d = e % DH;
      const int kj = k0 + r;
      float kv = 0.f, vv = 0.f;
      if (kj < s.Lk) {
        kv = to_f<T>(k[b * s.k_bs + static_cast<long long>(kj) * s.k_rs + h * DH + d]);
        vv = to_f<T>(v[b * s.v_bs + static_cast<long long>(kj) * s.v_rs + h * DH + d]);
      }
      Ks[r][d] = kv; Vs[r][d] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = warp * 2 + i;
      const int qi = q0 + r;
      if (qi >= s.Lq) continue;   // warp-uniform
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int d = 0; d < DH; ++d) {
        const float qd = Qs[r][d];
        s0 = fmaf(qd, Ks[lane][d], s0);
        s1 = fmaf(qd, Ks[lane + 32][d], s1);
      }
      const int j0 = k0 + lane, j1 = k0 + lane + 32;
      const int lim = s.causal ? qi + off : s.Lk - 1;
      if (j0 >= s.Lk || j0 > lim) s0 = -INFINITY;
      if (j1 >= s.Lk || j1 > lim) s1 = -INFINITY;
      const float mn = fmaxf(m[i], warp_max(fmaxf(s0, s1)));
      if (mn == -INFINITY) continue;
      const float corr = __expf(m[i] - mn);
      const float p0 = __expf(s0 - mn), p1 = __expf(s1 - mn);
      l[i] = l[i] * corr + warp_sum(p0 + p1);
      m[i] = mn;
#pragma unroll
      for (int d = 0; d < DPL; ++d) acc[i][d] *= corr;
      for (int j = 0; j < 32; ++j) {
        const float pa = __shfl_sync(0xffffffffu, p0, j);
        const float pb = __shfl_sync(0xffffffffu, p1, j);
#pragma unroll
        for (int d = 0; d < DPL; ++d) {
          acc[i][d] = fmaf(pa, Vs[j][lane + 32 * d], acc[i][d]);
          acc[i][d] = fmaf(pb, Vs[j + 32][lane + 32 * d], acc[i][d]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int qi = q0 + warp * 2 + i;
    if (qi >= s.Lq) continue;
    const float inv = l[i] > 0.f ? 1.0f / l[i] : 0.f;
#pragma unroll
    for (int d = 0; d < DPL; ++d)
      o[b * s.o_bs + static_cast<long long>(qi) * s.o_rs + h * DH + lane + 32 * d] = from_f<T>(acc[i][d] * inv);
    if (lse && lane == 0) lse[(static_cast<long long>(b) * s.H + h) * s.Lq + qi] = m[i] + logf(l[i]);
  }
}

// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]
template <typename T, int DH>
__global__ void attn_delta_kernel(const ns_attn_shape s, const T* __restrict__ o, const T* __restrict__ d_o, float* __restrict__ delta) {
  const long long idx = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(s.B) * s.H * s.Lq;
  if (idx >= total) return;
  const int qi = static_cast<int>(idx % s.Lq);
  const int h = static_cast<int>((idx / s.Lq) % s.H);
  const int b = static_cast<int>(idx / (static_cast<long long>(s.Lq) * s.H));
  float a = 0.f;
  for (int d = lane; d < DH; d += 32) {
    const long long off = b * s.o_bs + static_cast<long long>(qi) * s.o_rs + h * DH + d;
    a = fmaf(to_f<T>(o[off]), to_f<T>(d_o[off]), a);
  }
  a = warp_sum(a);
  if (lane == 0) delta[idx] = a;
}

// dq_i = sum_j p_ij (dp_ij - delta_i) k_j,   p_ij = exp(s_ij - lse_i),  dp_ij = dO_i . v_j
template <typename T, int DH>
__global__ void __launch_bounds__(256) attn_bwd_dq_simt_kernel(const ns_attn_shape s, const T* __restrict__ q, const T* __restrict__ k,
                                                               const T* __restrict__ v, const T* __restrict__ d_o,
                                                               const float* __restrict__ lse, const float* __restrict__ delta,
                                                               T* __restrict__ dq) {
  constexpr int DPL = DH / 32;
  __shared__ float Ks[kKT][DH + 1];
  __shared__ float Vs[kKT][DH + 1];
  __shared__ float Qs[kQPB][DH];
  __shared__ float Os[kQPB][DH];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQPB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int off = s.Lk - s.Lq;
  for (int e = threadIdx.x; e < kQPB * DH; e += 256) {
    const int r = e / DH, d = e % DH;
    const int qi = q0 + r;
    Qs[r][d] = qi < s.Lq ? to_f<T>(q[b * s.q_bs + static_cast<long long>(qi) * s.q_rs + h * DH + d]) : 0.f;
    Os[r][d] = qi < s.Lq ? to_f<T>(d_o[b * s.o_bs + static_cast<long long>(qi) * s.o_rs + h * DH + d]) : 0.f;
  }
  float acc[2][DPL];
  float ls[2], dl[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int qi = q0 + warp * 2 + i;
    const long long idx = (static_cast<long long>(b) * s.H + h) * s.Lq + qi;
    ls[i] = qi < s.Lq ? lse[idx] : 0.f;
    dl[i] = qi < s.Lq ? delta[idx] : 0.f;
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[i][d] = 0.f;
  }
  const int k_end = s.causal ? min(s.Lk, q0 + kQPB + off) : s.Lk;
  for (int k0 = 0; k0 < k_end; k0 += kKT) {
    __syncthreads();
    for (int e = threadIdx.x; e < kKT * DH; e += 256) {
      const int r = e / DH, d = e % DH;
      const int kj = k0 + r;
      float kv = 0.f, vv = 0.f;
      if (kj < s.Lk) {
        kv = to_f<T>(k[b * s.k_bs + static_cast<long long>(kj) * s.k_rs + h * DH + d]);
        vv = to_f<T>(v[b * s.v_bs + static_cast<long long>(kj) * s.v_rs + h * DH + d]);
      }
      Ks[r][d] = kv; Vs[r][d] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = warp * 2 + i;
      const int qi = q0 + r;
      if (qi >= s.Lq) continue;
      float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;
#pragma unroll 8
      for (int d = 0; d < DH; ++d) {
        const float qd = Qs[r][d], od = Os[r][d];
        s0 = fmaf(qd, Ks[lane][d], s0); s1 = fmaf(qd, Ks[lane + 32][d], s1);
        p0 = fmaf(od, Vs[lane][d], p0); p1 = fmaf(od, Vs[lane + 32][d], p1);
      }
      const int j0 = k0 + lane, j1 = k0 + lane + 32;
      const int lim = s.causal ? qi + off : s.Lk - 1;
      const float ds0 = (j0 >= s.Lk || j0 > lim) ? 0.f : __expf(s0 - ls[i]) * (p0 - dl[i]);
      const float ds1 = (j1 >= s.Lk || j1 > lim) ? 0.f : __expf(s1 - ls[i]) * (p1 - dl[i]);
      for (int j = 0; j < 32; ++j) {
        const float a0 = __shfl_sync(0xffffffffu, ds0, j);
        const float a1 = __shfl_sync(0xffffffffu, ds1, j);
#pragma unroll
        for (int d = 0; d < DPL; ++d) {
          acc[i][d] = fmaf(a0, Ks[j][lane + 32 * d], acc[i][d]);
          acc[i][d] = fmaf(a1, Ks[j + 32][lane + 32 * d], acc[i][d]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int qi = q0 + warp * 2 + i;
    if (qi >= s.Lq) continue;
#pragma unroll
    for (int d = 0; d < DPL; ++d)
      dq[b * s.q_bs + static_cast<long long>(qi) * s.q_rs + h * DH + lane + 32 * d] = from_f<T>(acc[i][d]);
  }
}

// dv_j = sum_i p_ij dO_i ;  dk_j = sum_i p_ij (dp_ij - delta_i) q_i      (block = 16 keys, loops over query tiles of 64)
template <typename T, int DH>
__global__ void __launch_bounds__(256) attn_bwd_dkv_simt_kernel(const ns_attn_shape s, const T* __restrict__ q, const T* __restrict__ k,
                                                                const T* __restrict__ v, const T* __restrict__ d_o,
                                                                const float* __restrict__ lse, const float* __restrict__ delta,
                                                                T* __restrict__ dk, T* __restrict__ dv) {
  constexpr int DPL = DH / 32;
  __shared__ float Qs[kKT][DH + 1];
  __shared__ float Os[kKT][DH + 1];
  __shared__ float Ls[kKT], Ds[kKT];
  __shared__ float Kb[kQPB][DH];
  __shared__ float Vb[kQPB][DH];
  const int b = blockIdx.z, h = blockIdx.y, kb0 = blockIdx.x * kQPB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int off = s.Lk - s.Lq;
  for (int e = threadIdx.x; e < kQPB * DH; e += 256) {
    const int r = e / DH, d = e % DH;
    const int kj = kb0 + r;
    Kb[r][d] = kj < s.Lk ? to_f<T>(k[b * s.k_bs + static_cast<long long>(kj) * s.k_rs + h * DH + d]) : 0.f;
    Vb[r][d] = kj < s.Lk ? to_f<T>(v[b * s.v_bs + static_cast<long long>(kj) * s.v_rs + h * DH + d]) : 0.f;
  }
  float acck[2][DPL], accv[2][DPL];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int d = 0; d < DPL; ++d) { acck[i][d] = 0.f; accv[i][d] = 0.f; }
  // causal: key j is seen by queries i >= j - off
  const int q_begin = s.causal ? max(0, kb0 - off) : 0;
  for (int qt = (q_begin / kKT) * kKT; qt < s.Lq; qt += kKT) {
    __syncthreads();
    for (int e = threadIdx.x; e < kKT * DH; e += 256) {
      const int r = e / DH, d = e % DH;
      const int qi = qt + r;
      float qv = 0.f, ov = 0.f;
      if (qi < s.Lq) {
        qv = to_f<T>(q[b * s.q_bs + static_cast<long long>(qi) * s.q_rs + h * DH + d]);
        ov = to_f<T>(d_o[b * s.o_bs + static_cast<long long>(qi) * s.o_rs + h * DH + d]);
      }
      Qs[r][d] = qv; Os[r][d] = ov;
    }
    if (threadIdx.x < kKT) {
      const int qi = qt + threadIdx.x;
      const long long idx = (static_cast<long long>(b) * s.H + h) * s.Lq + qi;
      Ls[threadIdx.x] = qi < s.Lq ? lse[idx] : 0.f;
      Ds[threadIdx.x] = qi < s.Lq ? delta[idx] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = warp * 2 + i;
      const int kj = kb0 + r;
      if (kj >= s.Lk) continue;
      float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;   // lane handles queries qt+lane and qt+lane+32
#pragma unroll 8
      for (int d = 0; d < DH; ++d) {
        const float kd = Kb[r][d], vd = Vb[r][d];
        s0 = fmaf(kd, Qs[lane][d], s0); s1 = fmaf(kd, Qs[lane + 32][d], s1);
        p0 = fmaf(vd, Os[lane][d], p0); p1 = fmaf(vd, Os[lane + 32][d], p1);
      }
      const int i0 = qt + lane, i1 = qt + lane + 32;
      const bool ok0 = i0 < s.Lq && (!s.causal || kj <= i0 + off);
      const bool ok1 = i1 < s.Lq && (!s.causal || kj <= i1 + off);
      const float pr0 = ok0 ? __expf(s0 - Ls[lane]) : 0.f;
      const float pr1 = ok1 ? __expf(s1 - Ls[lane + 32]) : 0.f;
      const float ds0 = pr0 * (p0 - Ds[lane]);
      const float ds1 = pr1 * (p1 - Ds[lane + 32]);
      for (int j = 0; j < 32; ++j) {
        const float a0 = __shfl_sync(0xffffffffu, pr0, j), a1 = __shfl_sync(0xffffffffu, pr1, j);
        const float c0 = __shfl_sync(0xffffffffu, ds0, j), c1 = __shfl_sync(0xffffffffu, ds1, j);
#pragma unroll
        for (int d = 0; d < DPL; ++d) {
          accv[i][d] = fmaf(a0, Os[j][lane + 32 * d], accv[i][d]);
          accv[i][d] = fmaf(a1, Os[j + 32][lane + 32 * d], accv[i][d]);
          acck[i][d] = fmaf(c0, Qs[j][lane + 32 * d], acck[i][d]);
          acck[i][d] = fmaf(c1, Qs[j + 32][lane + 32 * d], acck[i][d]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int kj = kb0 + warp * 2 + i;
    if (kj >= s.Lk) continue;
#pragma unroll
    for (int d = 0; d < DPL; ++d) {
      dk[b * s.k_bs + static_cast<long long>(kj) * s.k_rs + h * DH + lane + 32 * d] = from_f<T>(acck[i][d]);
      dv[b * s.v_bs + static_cast<long long>(kj) * s.v_rs + h * DH + lane + 32 * d] = from_f<T>(accv[i][d]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tiny self-attention (Lq == Lk <= 32): the teacher-forced decoder self-attention of the training step (L = 32 label
// positions, causal; utils/load_model.py:512-532).  ONE WARP per (batch, head), lane = query row; K / V (and in the
// backward Q / dO, P / dS) live in shared memory and are read as broadcasts, so the softmax needs no cross-lane traffic.
// The general kernels above spend ~30 us (forward) / ~120 us (backward, three launches) on this 32 x 32 problem.
// shared-memory rows are DH floats (16-byte aligned) and always read as float4 BROADCASTS (every lane the same row), so one
// LDS.128 feeds four FMAs; each lane's own q / dO row lives in registers.
template <typename T, int DH>
__device__ __forceinline__ void tiny_load_tile(float* dst, const T* __restrict__ src, long long rs, int L, int lane) {
  // 16-byte loads, all issued before the first use (a single warp has nobody to hide a load latency behind)
  constexpr int EPV = 16 / sizeof(T);              // elements per vector
  constexpr int VPR = DH / EPV;                    // vectors per row
  constexpr int NIT = 32 * VPR / 32;
  uint4 buf[NIT];
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    const int e = i * 32 + lane, r = e / VPR, c = e % VPR;
    buf[i] = r < L ? __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(r) * rs + c * EPV)) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    const int e = i * 32 + lane, r = e / VPR, c = e % VPR;
    float* d = dst + r * DH + c * EPV;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<uint4*>(d) = buf[i];
    } else {
      float2 f;
      f = unpack_bf16x2(buf[i].x); d[0] = f.x; d[1] = f.y;
      f = unpack_bf16x2(buf[i].y); d[2] = f.x; d[3] = f.y;
      f = unpack_bf16x2(buf[i].z); d[4] = f.x; d[5] = f.y;
      f = unpack_bf16x2(buf[i].w); d[6] = f.x; d[7] = f.y;
    }
  }
}
template <typename T, int DH>
__device__ __forceinline__ void tiny_store_row(T* __restrict__ p, const float (&a)[DH], float scale) {   // 16-byte stores
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int d = 0; d < DH; d += 4) *reinterpret_cast<float4*>(p + d) = make_float4(a[d] * scale, a[d + 1] * scale, a[d + 2] * scale, a[d + 3] * scale);
  } else {
#pragma unroll
    for (int d = 0; d < DH; d += 8) {
      uint4 u;
      u.x = pack_bf16x2(a[d] * scale, a[d + 1] * scale); u.y = pack_bf16x2(a[d + 2] * scale, a[d + 3] * scale);
      u.z = pack_bf16x2(a[d + 4] * scale, a[d + 5] * scale); u.w = pack_bf16x2(a[d + 6] * scale, a[d + 7] * scale);
      *reinterpret_cast<uint4*>(p + d) = u;
    }
  }
}
template <typename T, int DH>
__device__ __forceinline__ void tiny_load_row(float (&a)[DH], const T* __restrict__ p) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(p + d));
      a[d] = x.x; a[d + 1] = x.y; a[d + 2] = x.z; a[d + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int d = 0; d < DH; d += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p + d));
      float2 f;
      f = unpack_bf16x2(u.x); a[d] = f.x; a[d + 1] = f.y;
      f = unpack_bf16x2(u.y); a[d + 2] = f.x; a[d + 3] = f.y;
      f = unpack_bf16x2(u.z); a[d + 4] = f.x; a[d + 5] = f.y;
      f = unpack_bf16x2(u.w); a[d + 6] = f.x; a[d + 7] = f.y;
    }
  }
}
template <int DH>
__device__ __forceinline__ float tiny_dot(const float (&a)[DH], const float* __restrict__ row) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int d = 0; d < DH; d += 8) {
    const float4 x = *reinterpret_cast<const float4*>(row + d), y = *reinterpret_cast<const float4*>(row + d + 4);
    s0 = fmaf(a[d], x.x, s0); s0 = fmaf(a[d + 1], x.y, s0); s0 = fmaf(a[d + 2], x.z, s0); s0 = fmaf(a[d + 3], x.w, s0);
    s1 = fmaf(a[d + 4], y.x, s1); s1 = fmaf(a[d + 5], y.y, s1); s1 = fmaf(a[d + 6], y.z, s1); s1 = fmaf(a[d + 7], y.w, s1);
  }
  return s0 + s1;
}
template <int DH>
__device__ __forceinline__ void tiny_axpy(float (&acc)[DH], float w, const float* __restrict__ row) {
#pragma unroll
  for (int d = 0; d < DH; d += 4) {
    const float4 x = *reinterpret_cast<const float4*>(row + d);
    acc[d] = fmaf(w, x.x, acc[d]); acc[d + 1] = fmaf(w, x.y, acc[d + 1]); acc[d + 2] = fmaf(w, x.z, acc[d + 2]); acc[d + 3] = fmaf(w, x.w, acc[d + 3]);
  }
}

template <typename T, int DH>
__global__ void __launch_bounds__(32) attn_tiny_fwd_kernel(const ns_attn_shape s, const T* __restrict__ q, const T* __restrict__ k,
                                                           const T* __restrict__ v, T* __restrict__ o, float* __restrict__ lse) {
  __shared__ __align__(16) float Ks[32 * DH];
  __shared__ __align__(16) float Vs[32 * DH];
  const int h = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
  const int L = s.Lq;
  tiny_load_tile<T, DH>(Ks, k + b * s.k_bs + h * DH, s.k_rs, L, lane);
  tiny_load_tile<T, DH>(Vs, v + b * s.v_bs + h * DH, s.v_rs, L, lane);
  float qr[DH], acc[DH];
  const bool act = lane < L;
  const T* qp = q + b * s.q_bs + static_cast<long long>(act ? lane : 0) * s.q_rs + h * DH;
  tiny_load_row<T, DH>(qr, qp);
#pragma unroll
  for (int d = 0; d < DH; ++d) acc[d] = 0.f;
  __syncwarp();
  float m = -INFINITY, l = 0.f;
  const int jend = s.causal ? lane + 1 : L;       // Lq == Lk: query i sees keys j <= i
  for (int j = 0; j < L; ++j) {
    const float sc = tiny_dot<DH>(qr, Ks + j * DH);
    const bool on = j < jend && act;
    const float mn = on ? fmaxf(m, sc) : m;
    const float corr = on ? __expf(m - mn) : 1.f, pj = on ? __expf(sc - mn) : 0.f;
    l = l * corr + pj;
    m = mn;
#pragma unroll
    for (int d = 0; d < DH; ++d) acc[d] *= corr;
    tiny_axpy<DH>(acc, pj, Vs + j * DH);
  }
  if (act) {
    const float inv = 1.0f / l;
    tiny_store_row<T, DH>(o + b * s.o_bs + static_cast<long long>(lane) * s.o_rs + h * DH, acc, inv);
    if (lse) lse[(static_cast<long long>(b) * s.H + h) * L + lane] = m + logf(l);
  }
}

template <typename T, int DH>
__global__ void __launch_bounds__(32) attn_tiny_bwd_kernel(const ns_attn_shape s, const T* __restrict__ q, const T* __restrict__ k,
                                                           const T* __restrict__ v, const T* __restrict__ d_o, const float* __restrict__ lse,
                                                           float* __restrict__ delta_out, T* __restrict__ dq, T* __restrict__ dk,
                                                           T* __restrict__ dv) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;
  float* Vs = Ks + 32 * DH;
  float* Qs = Vs + 32 * DH;
  float* Os = Qs + 32 * DH;                       // dO
  float (*Ps)[33] = reinterpret_cast<float (*)[33]>(Os + 32 * DH);
  float (*Ss)[33] = Ps + 32;                      // dP, then dS
  const int h = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
  const int L = s.Lq;
  tiny_load_tile<T, DH>(Ks, k + b * s.k_bs + h * DH, s.k_rs, L, lane);
  tiny_load_tile<T, DH>(Vs, v + b * s.v_bs + h * DH, s.v_rs, L, lane);
  tiny_load_tile<T, DH>(Qs, q + b * s.q_bs + h * DH, s.q_rs, L, lane);
  tiny_load_tile<T, DH>(Os, d_o + b * s.o_bs + h * DH, s.o_rs, L, lane);
  __syncwarp();
  const bool act = lane < L;
  const int jend = s.causal ? lane + 1 : L;
  const float my_lse = act ? lse[(static_cast<long long>(b) * s.H + h) * L + lane] : 0.f;
  float dlt = 0.f;
  {
    // ---- phase A1 (lane = query i): P and dP rows, delta_i = sum_j p_ij dp_ij (= dO_i . O_i)
    float qr[DH], orow[DH];
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      const float4 a = *reinterpret_cast<const float4*>(Qs + lane * DH + d), c = *reinterpret_cast<const float4*>(Os + lane * DH + d);
      qr[d] = a.x; qr[d + 1] = a.y; qr[d + 2] = a.z; qr[d + 3] = a.w;
      orow[d] = c.x; orow[d + 1] = c.y; orow[d + 2] = c.z; orow[d + 3] = c.w;
    }
    for (int j = 0; j < L; ++j) {
      const float sc = tiny_dot<DH>(qr, Ks + j * DH);
      const float dp = tiny_dot<DH>(orow, Vs + j * DH);
      const float pj = (act && j < jend) ? __expf(sc - my_lse) : 0.f;
      Ps[lane][j] = pj;
      Ss[lane][j] = dp;
      dlt = fmaf(pj, dp, dlt);
    }
  }
  if (act && delta_out) delta_out[(static_cast<long long>(b) * s.H + h) * L + lane] = dlt;
  float acc[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) acc[d] = 0.f;
  // ---- phase A2: dS row, dq_i = sum_j ds_ij k_j
  for (int j = 0; j < L; ++j) {
    const float ds = Ps[lane][j] * (Ss[lane][j] - dlt);
    Ss[lane][j] = ds;
    tiny_axpy<DH>(acc, ds, Ks + j * DH);
  }
  if (act) tiny_store_row<T, DH>(dq + b * s.q_bs + static_cast<long long>(lane) * s.q_rs + h * DH, acc, 1.0f);
  __syncwarp();
  // ---- phase B (lane = key j): dv_j = sum_i p_ij dO_i ; dk_j = sum_i ds_ij q_i
  float av[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { acc[d] = 0.f; av[d] = 0.f; }
  for (int i = 0; i < L; ++i) {
    tiny_axpy<DH>(av, Ps[i][lane], Os + i * DH);
    tiny_axpy<DH>(acc, Ss[i][lane], Qs + i * DH);
  }
  if (act) {
    tiny_store_row<T, DH>(dk + b * s.k_bs + static_cast<long long>(lane) * s.k_rs + h * DH, acc, 1.0f);
    tiny_store_row<T, DH>(dv + b * s.v_bs + static_cast<long long>(lane) * s.v_rs + h * DH, av, 1.0f);
  }
}

template <typename T, int DH>
static bool tiny_eligible(const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o) {
  constexpr long long epv = 16 / sizeof(T);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return s.Lq == s.Lk && s.Lq <= 32 && s.H <= 65535 && s.B <= 65535 && al(q) && al(k) && al(v) && al(o) && s.q_rs % epv == 0 &&
         s.k_rs % epv == 0 && s.v_rs % epv == 0 && s.o_rs % epv == 0 && s.q_bs % epv == 0 && s.k_bs % epv == 0 && s.v_bs % epv == 0 &&
         s.o_bs % epv == 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Single-query attention (Lq == 1): the per-token decoder step of greedy generation against the self-attention cache
// (Lk = t <= 448) and the cross-attention cache (Lk = 1500) -- utils/load_model.py:1332-1351, HF modeling_whisper.py:314-336.
// Pure streaming: every K and V row is read once with 16-byte loads (8 lanes per 128-byte bf16 row, 4 keys per warp
// instruction), scores / weights stay in registers, one online-softmax state per warp, partial results of the 8 warps are
// merged in shared memory.  HBM-bound by construction: B*H*Lk*Dh*2 elements per call.
// WARPS = 8 (256 threads, 3 CTAs per SM) or 4 (128 threads, 7 CTAs per SM).  At B * H = 1024 (batch 128) the 8-warp shape
// needs 2.3 waves of 444 CTAs -- the third wave is a third full and the kernel reaches 58 % of the copy bandwidth -- while the
// 4-warp shape holds all 1024 CTAs (1036 slots) at once: no tail.  Small batches keep 8 warps per (sample, head).
template <typename T, int DH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? 7 : 3) attn_decode_kernel(const ns_attn_shape s, const T* __restrict__ q, const T* __restrict__ k,
                                                          const T* __restrict__ v, T* __restrict__ o, float* __restrict__ lse,
                                                          const int* __restrict__ kv_row, long long kv_ld) {
  constexpr int EPV = 16 / sizeof(T);              // elements per 16-byte vector
  constexpr int CPR = DH / EPV;                    // vectors (lanes) per row
  constexpr int KPW = 32 / CPR;                    // keys per warp instruction
  __shared__ float sm_m[WARPS], sm_l[WARPS];
  __shared__ __align__(16) float sm_o[WARPS][DH];
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane % CPR, kk = lane / CPR;
  float qv[EPV];
  {
    const T* qp = q + b * s.q_bs + h * DH + c * EPV;       // Lq == 1: row 0
#pragma unroll
    for (int e = 0; e < EPV; ++e) qv[e] = to_f<T>(qp[e]);
  }
  float m = -INFINITY, l = 0.f, acc[EPV];
#pragma unroll
  for (int e = 0; e < EPV; ++e) acc[e] = 0.f;
  // kv_row (beam search): key/value j of batch row b lives in cache row kv_row[b][j] -- the beam reorder permutes this small
  // table instead of copying the cache (utils/load_model.py:1353-1360 _reorder_cache index_selects every layer's K and V)
  const int* rowtab = kv_row ? kv_row + b * kv_ld : nullptr;
  const T* kb = k + (rowtab ? 0 : b * s.k_bs) + h * DH + c * EPV;
  const T* vb = v + (rowtab ? 0 : b * s.v_bs) + h * DH + c * EPV;
  // UNR key groups per loop trip, all their 16-byte loads issued before the first use (memory-level parallelism is the whole
  // game here: ~64 KB must be in flight per SM to cover the HBM latency)
  constexpr int UNR = 4;
  for (int j0 = warp * KPW; j0 < s.Lk; j0 += WARPS * KPW * UNR) {
    uint4 ku[UNR], vu[UNR];
    bool ok[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + u * WARPS * KPW + kk;
      ok[u] = j < s.Lk;
      const long long pr = (rowtab && ok[u]) ? __ldg(rowtab + j) : 0;
      ku[u] = ok[u] ? __ldg(reinterpret_cast<const uint4*>(kb + pr * s.k_bs + static_cast<long long>(j) * s.k_rs)) : make_uint4(0, 0, 0, 0);
      vu[u] = ok[u] ? __ldg(reinterpret_cast<const uint4*>(vb + pr * s.v_bs + static_cast<long long>(j) * s.v_rs)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (j0 + u * WARPS * KPW >= s.Lk) break;           // warp-uniform: this group has no valid key
      float kx[EPV], vx[EPV];
      if constexpr (sizeof(T) == 4) {
        kx[0] = __uint_as_float(ku[u].x); kx[1] = __uint_as_float(ku[u].y); kx[2] = __uint_as_float(ku[u].z); kx[3] = __uint_as_float(ku[u].w);
        vx[0] = __uint_as_float(vu[u].x); vx[1] = __uint_as_float(vu[u].y); vx[2] = __uint_as_float(vu[u].z); vx[3] = __uint_as_float(vu[u].w);
      } else {
        float2 f;
        f = unpack_bf16x2(ku[u].x); kx[0] = f.x; kx[1] = f.y; f = unpack_bf16x2(ku[u].y); kx[2] = f.x; kx[3] = f.y;
        f = unpack_bf16x2(ku[u].z); kx[4] = f.x; kx[5] = f.y; f = unpack_bf16x2(ku[u].w); kx[6] = f.x; kx[7] = f.y;
        f = unpack_bf16x2(vu[u].x); vx[0] = f.x; vx[1] = f.y; f = unpack_bf16x2(vu[u].y); vx[2] = f.x; vx[3] = f.y;
        f = unpack_bf16x2(vu[u].z); vx[4] = f.x; vx[5] = f.y; f = unpack_bf16x2(vu[u].w); vx[6] = f.x; vx[7] = f.y;
      }
      float sc = 0.f;
#pragma unroll
      for (int e = 0; e < EPV; ++e) sc = fmaf(qv[e], kx[e], sc);
#pragma unroll
      for (int off = CPR / 2; off > 0; off >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, off);   // over the lanes of one row
      if (!ok[u]) sc = -INFINITY;
      float mt = sc;                                                                             // max over the KPW keys
#pragma unroll
      for (int off = CPR; off < 32; off <<= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, off));
      const float mn = fmaxf(m, mt);               // finite: the first key of this group is valid
      const float corr = __expf(m - mn), pj = __expf(sc - mn);
      m = mn;
      l = l * corr + pj;                           // per key group; groups are summed at the end
#pragma unroll
      for (int e = 0; e < EPV; ++e) acc[e] = fmaf(pj, vx[e], acc[e] * corr);
    }
  }
  // merge the KPW key groups of this warp (they share m), then the 8 warps
#pragma unroll
  for (int off = CPR; off < 32; off <<= 1) {
    l += __shfl_xor_sync(0xffffffffu, l, off);
#pragma unroll
    for (int e = 0; e < EPV; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
  }
  if (kk == 0) {
#pragma unroll
    for (int e = 0; e < EPV; ++e) sm_o[warp][c * EPV + e] = acc[e];
    if (c == 0) { sm_m[warp] = m; sm_l[warp] = l; }
  }
  __syncthreads();
  if (threadIdx.x < DH) {
    float mg = -INFINITY;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) mg = fmaxf(mg, sm_m[w]);
    float lg = 0.f, og = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const float sc = sm_m[w] == -INFINITY ? 0.f : __expf(sm_m[w] - mg);
      lg = fmaf(sm_l[w], sc, lg);
      og = fmaf(sm_o[w][threadIdx.x], sc, og);
    }
    o[b * s.o_bs + h * DH + threadIdx.x] = from_f<T>(og / lg);
    if (lse && threadIdx.x == 0) lse[static_cast<long long>(b) * s.H + h] = mg + logf(lg);
  }
}

// 4 warps per (sample, head) once the CTAs alone saturate the GPU (see attn_decode_kernel); NS_DECODE_WARPS=4|8 overrides
static bool decode_four_warps(const ns_attn_shape& s) {
  static const int forced = getenv("NS_DECODE_WARPS") ? atoi(getenv("NS_DECODE_WARPS")) : 0;
  if (forced == 4) return true;
  if (forced == 8) return false;
  return static_cast<long long>(s.B) * s.H > 3LL * sm_count();
}

template <typename T, int DH>
static bool decode_eligible(const ns_attn_shape& s, const void* q, const void* k, const void* v) {
  constexpr long long epv = 16 / sizeof(T);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  // Lq == 1: with or without the causal flag the query sees every key (j <= 0 + Lk - 1)
  return s.Lq == 1 && s.Lk >= 1 && s.H <= 65535 && s.B <= 65535 && al(k) && al(v) && s.k_rs % epv == 0 && s.v_rs % epv == 0 &&
         s.k_bs % epv == 0 && s.v_bs % epv == 0 && q != nullptr;
}

template <typename T, int DH>
static int attn_fwd_simt_t(const ns_attn_shape& s, const void* q, const void* k, const void* v, void* o, float* lse, cudaStream_t st) {
  if (decode_eligible<T, DH>(s, q, k, v)) {
    if (decode_four_warps(s))
      NS_CUDA(launch_pdl(attn_decode_kernel<T, DH, 4>, dim3(s.H, s.B), dim3(128), 0, st, s, reinterpret_cast<const T*>(q), reinterpret_cast<const T*>(k),
                         reinterpret_cast<const T*>(v), reinterpret_cast<T*>(o), lse, static_cast<const int*>(nullptr), 0LL));
    else
      NS_CUDA(launch_pdl(attn_decode_kernel<T, DH, 8>, dim3(s.H, s.B), dim3(256), 0, st, s, reinterpret_cast<const T*>(q), reinterpret_cast<const T*>(k),
                         reinterpret_cast<const T*>(v), reinterpret_cast<T*>(o), lse, static_cast<const int*>(nullptr), 0LL));
    NS_LAUNCH_CHECK();
    count(C_ATTN_SIMT);
    return NS_OK;
  }
  if (tiny_eligible<T, DH>(s, q, k, v, o)) {
    attn_tiny_fwd_kernel<T, DH><<<dim3(s.H, s.B), 32, 0, st>>>(s, reinterpret_cast<const T*>(q), reinterpret_cast<const T*>(k),
                                                             reinterpret_cast<const T*>(v), reinterpret_cast<T*>(o), lse);
    NS_LAUNCH_CHECK();
    count(C_ATTN_SIMT);
    return NS_OK;
  }
  dim3 grid((s.Lq + kQPB - 1) / kQPB, s.H, s.B);
  attn_fwd_simt_kernel<T, DH><<<grid, 256, 0, st>>>(s, reinterpret_cast<const T*>(q), reinterpret_cast<const T*>(k),
                                                    reinterpret_cast<const T*>(v), reinterpret_cast<T*>(o), lse);
  NS_LAUNCH_CHECK();
  count(C_ATTN_SIMT);
  return NS_OK;
}

template <typename T, int DH>
static int attn_bwd_simt_t(const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                           const float* lse, float* delta, void* dq, void* dk, void* dv, cudaStream_t st) {
  if (tiny_eligible<T, DH>(s, q, k, v, d_o) && (reinterpret_cast<uintptr_t>(dq) & 15) == 0) {
    constexpr int smem = (4 * 32 * DH + 2 * 32 * 33) * 4;
    static bool attr_done = false;
    if (!attr_done) {
      NS_CUDA(cudaFuncSetAttribute(attn_tiny_bwd_kernel<T, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_done = true;
    }
    attn_tiny_bwd_kernel<T, DH><<<dim3(s.H, s.B), 32, smem, st>>>(s, reinterpret_cast<const T*>(q), reinterpret_cast<const T*>(k),
                                                                 reinterpret_cast<const T*>(v), reinterpret_cast<const T*>(d_o), lse, delta,
                                                                 reinterpret_cast<T*>(dq), reinterpret_cast<T*>(dk), reinterpret_cast<T*>(dv));
    NS_LAUNCH_CHECK();
    count(C_ATTN_SIMT);
    return NS_OK;
  }
  const long long total = static_cast<long long>(s.B) * s.H * s.Lq;
  attn_delta_kernel<T, DH><<<static_cast<unsigned>((total + 7) / 8), 256, 0, st>>>(s, reinterpret_cast<const T*>(o),
                                                                                    reinterpret_cast<const T*>(d_o), delta);
  NS_LAUNCH_CHECK();
  dim3 gq((s.Lq + kQPB - 1) / kQPB, s.H, s.B);
  attn_bwd_dq_simt_kernel<T, DH><<<gq, 256, 0, st>>>(s, reinterpret_cast<const T*>(q), reinterpret_cast<const T*>(k),
                                                     reinterpret_cast<const T*>(v), reinterpret_cast<const T*>(d_o), lse, delta,
                                                     reinterpret_cast<T*>(dq));
  NS_LAUNCH_CHECK();
  dim3 gk((s.Lk + kQPB - 1) / kQPB, s.H, s.B);
  attn_bwd_dkv_simt_kernel<T, DH><<<gk, 256, 0, st>>>(s, reinterpret_cast<const T*>(q), reinterpret_cast<const T*>(k),
                                                      reinterpret_cast<const T*>(v), reinterpret_cast<const T*>(d_o), lse, delta,
                                                      reinterpret_cast<T*>(dk), reinterpret_cast<T*>(dv));
  NS_LAUNCH_CHECK();
  count(C_ATTN_SIMT, 3);
  return NS_OK;
}

int attention_delta(int dtype, const ns_attn_shape& s, const void* o, const void* d_o, float* delta, cudaStream_t st) {
  const long long total = static_cast<long long>(s.B) * s.H * s.Lq;
  const unsigned grid = static_cast<unsigned>((total + 7) / 8);
  if (dtype == NS_F32) {
    if (s.Dh == 64) attn_delta_kernel<float, 64><<<grid, 256, 0, st>>>(s, reinterpret_cast<const float*>(o), reinterpret_cast<const float*>(d_o), delta);
    else attn_delta_kernel<float, 32><<<grid, 256, 0, st>>>(s, reinterpret_cast<const float*>(o), reinterpret_cast<const float*>(d_o), delta);
  } else {
    if (s.Dh == 64) attn_delta_kernel<__nv_bfloat16, 64><<<grid, 256, 0, st>>>(s, reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(d_o), delta);
    else attn_delta_kernel<__nv_bfloat16, 32><<<grid, 256, 0, st>>>(s, reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(d_o), delta);
  }
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

template <typename T, int DH>
static int attn_decode_rows_t(const ns_attn_shape& s, const void* q, const void* k, const void* v, void* o, const int* kv_row, long long kv_ld,
                              cudaStream_t st) {
  if (!decode_eligible<T, DH>(s, q, k, v)) {
    set_error("ns_attention_decode_rows: needs Lq == 1 and 16-byte aligned K/V rows");
    return NS_ERR_UNSUPPORTED;
  }
  if (decode_four_warps(s))
    attn_decode_kernel<T, DH, 4><<<dim3(s.H, s.B), 128, 0, st>>>(s, reinterpret_cast<const T*>(q), reinterpret_cast<const T*>(k),
                                                                reinterpret_cast<const T*>(v), reinterpret_cast<T*>(o), nullptr, kv_row, kv_ld);
  else
    attn_decode_kernel<T, DH, 8><<<dim3(s.H, s.B), 256, 0, st>>>(s, reinterpret_cast<const T*>(q), reinterpret_cast<const T*>(k),
                                                                reinterpret_cast<const T*>(v), reinterpret_cast<T*>(o), nullptr, kv_row, kv_ld);
  NS_LAUNCH_CHECK();
  count(C_ATTN_SIMT);
  return NS_OK;
}
int attention_decode_rows(int dtype, const ns_attn_shape& s, const void* q, const void* k, const void* v, void* o, const int* kv_row,
                          long long kv_ld, cudaStream_t st) {
  if (dtype == NS_F32)
    return s.Dh == 64 ? attn_decode_rows_t<float, 64>(s, q, k, v, o, kv_row, kv_ld, st) : attn_decode_rows_t<float, 32>(s, q, k, v, o, kv_row, kv_ld, st);
  return s.Dh == 64 ? attn_decode_rows_t<__nv_bfloat16, 64>(s, q, k, v, o, kv_row, kv_ld, st)
                    : attn_decode_rows_t<__nv_bfloat16, 32>(s, q, k, v, o, kv_row, kv_ld, st);
}

int attention_fwd_simt(int dtype, const ns_attn_shape& s, const void* q, const void* k, const void* v, void* o, float* lse, cudaStream_t st) {
  if (dtype == NS_F32) return s.Dh == 64 ? attn_fwd_simt_t<float, 64>(s, q, k, v, o, lse, st) : attn_fwd_simt_t<float, 32>(s, q, k, v, o, lse, st);
  return s.Dh == 64 ? attn_fwd_simt_t<__nv_bfloat16, 64>(s, q, k, v, o, lse, st) : attn_fwd_simt_t<__nv_bfloat16, 32>(s, q, k, v, o, lse, st);
}
int attention_bwd_simt(int dtype, const ns_attn_shape& s, const void* q, const void* k, const void* v, const void* o, const void* d_o,
                       const float* lse, float* delta, void* dq, void* dk, void* dv, cudaStream_t st) {
  if (dtype == NS_F32)
    return s.Dh == 64 ? attn_bwd_simt_t<float, 64>(s, q, k, v, o, d_o, lse, delta, dq, dk, dv, st)
                      : attn_bwd_simt_t<float, 32>(s, q, k, v, o, d_o, lse, delta, dq, dk, dv, st);
  return s.Dh == 64 ? attn_bwd_simt_t<__nv_bfloat16, 64>(s, q, k, v, o, d_o, lse, delta, dq, dk, dv, st)
                    : attn_bwd_simt_t<__nv_bfloat16, 32>(s, q, k, v, o, d_o, lse, delta, dq, dk, dv, st);
}

}  // namespace ns
