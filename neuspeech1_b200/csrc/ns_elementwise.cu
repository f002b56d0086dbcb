// HBM-bound kernels: LayerNorm fwd/bwd, embedding, cross-entropy, greedy pick, EEG augmentation pass, casts/transposes,
// conv-weight re-layouts, fused clip + AdamW.  One warp per row with 16-byte vector accesses where the shape allows.
#include "ns_common.cuh"

namespace ns {

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per PAIR of rows (both rows' 16-byte loads are in flight together: with one row per warp the kernel ran at
// 4.5 TB/s, latency-bound).  NV = number of 8-element vectors per lane (bf16, d = 256*NV), register-resident.
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_bf16_kernel(long long rows, const __nv_bfloat16* __restrict__ x,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                                                          float* __restrict__ rstd_out, float eps) {
  constexpr int d = NV * 256;
  pdl_launch_dependents();
  pdl_wait();
  const long long row0 = (static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * 2;
  const int lane = threadIdx.x & 31;
  if (row0 >= rows) return;
  const bool two = row0 + 1 < rows;
  uint4 u[2][NV];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (row0 + ((r == 1 && two) ? 1 : 0)) * d);
#pragma unroll
    for (int i = 0; i < NV; ++i) u[r][i] = __ldg(xr + i * 32 + lane);
  }
  float v[2][NV * 8], s[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float2 f;
      f = unpack_bf16x2(u[r][i].x); v[r][8 * i + 0] = f.x; v[r][8 * i + 1] = f.y;
      f = unpack_bf16x2(u[r][i].y); v[r][8 * i + 2] = f.x; v[r][8 * i + 3] = f.y;
      f = unpack_bf16x2(u[r][i].z); v[r][8 * i + 4] = f.x; v[r][8 * i + 5] = f.y;
      f = unpack_bf16x2(u[r][i].w); v[r][8 * i + 6] = f.x; v[r][8 * i + 7] = f.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) s[r] += v[r][8 * i + j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s[0] += __shfl_xor_sync(0xffffffffu, s[0], o); s[1] += __shfl_xor_sync(0xffffffffu, s[1], o); }
  const float mean[2] = {s[0] * (1.0f / d), s[1] * (1.0f / d)};
  float q[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) { const float c = v[r][i] - mean[r]; q[r] = fmaf(c, c, q[r]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { q[0] += __shfl_xor_sync(0xffffffffu, q[0], o); q[1] += __shfl_xor_sync(0xffffffffu, q[1], o); }
  const float rstd[2] = {rsqrtf(q[0] * (1.0f / d) + eps), rsqrtf(q[1] * (1.0f / d) + eps)};
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r == 1 && !two) break;
      float o[8];
      o[0] = (v[r][8 * i + 0] - mean[r]) * rstd[r] * g0.x + b0.x; o[1] = (v[r][8 * i + 1] - mean[r]) * rstd[r] * g0.y + b0.y;
      o[2] = (v[r][8 * i + 2] - mean[r]) * rstd[r] * g0.z + b0.z; o[3] = (v[r][8 * i + 3] - mean[r]) * rstd[r] * g0.w + b0.w;
      o[4] = (v[r][8 * i + 4] - mean[r]) * rstd[r] * g1.x + b1.x; o[5] = (v[r][8 * i + 5] - mean[r]) * rstd[r] * g1.y + b1.y;
      o[6] = (v[r][8 * i + 6] - mean[r]) * rstd[r] * g1.z + b1.z; o[7] = (v[r][8 * i + 7] - mean[r]) * rstd[r] * g1.w + b1.w;
      uint4 w;
      w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]); w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
      reinterpret_cast<uint4*>(y + (row0 + r) * d)[i * 32 + lane] = w;
    }
  }
  if (lane == 0) {
    if (mean_out) { mean_out[row0] = mean[0]; if (two) mean_out[row0 + 1] = mean[1]; }
    if (rstd_out) { rstd_out[row0] = rstd[0]; if (two) rstd_out[row0 + 1] = rstd[1]; }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) ln_fwd_generic_kernel(long long rows, int d, const T* __restrict__ x,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             T* __restrict__ y, float* mean_out, float* rstd_out, float eps) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xr = x + row * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += to_f<T>(xr[c]);
  const float mean = warp_sum(s) / d;
  float q = 0.f;
  for (int c = lane; c < d; c += 32) { const float t = to_f<T>(xr[c]) - mean; q = fmaf(t, t, q); }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  T* yr = y + row * d;
  for (int c = lane; c < d; c += 32) yr[c] = from_f<T>((to_f<T>(xr[c]) - mean) * rstd * gamma[c] + beta[c]);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) (+ dres)
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(long long rows, int d, const T* __restrict__ dy, const T* __restrict__ x,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const T* __restrict__ dres,
                                                     T* __restrict__ dx) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* dyr = dy + row * d;
  const T* xr = x + row * d;
  const float mu = mean[row], rs = rstd[row];
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float g = to_f<T>(dyr[c]) * gamma[c];
    const float xh = (to_f<T>(xr[c]) - mu) * rs;
    s1 += g; s2 = fmaf(g, xh, s2);
  }
  s1 = warp_sum(s1) / d; s2 = warp_sum(s2) / d;
  T* dxr = dx + row * d;
  for (int c = lane; c < d; c += 32) {
    const float g = to_f<T>(dyr[c]) * gamma[c];
    const float xh = (to_f<T>(xr[c]) - mu) * rs;
    float v = rs * (g - s1 - xh * s2);
    if (dres) v += to_f<T>(dres[row * d + c]);
    dxr[c] = from_f<T>(v);
  }
}

// bf16 register-resident variant of the backward (d = 256*NV)
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_bf16_kernel(long long rows, const __nv_bfloat16* __restrict__ dy,
                                                          const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx) {
  constexpr int d = NV * 256;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint4* dyr = reinterpret_cast<const uint4*>(dy + row * d);
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * d);
  const float mu = mean[row], rs = rstd[row];
  float g[NV * 8], xh[NV * 8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint4 a = __ldg(dyr + i * 32 + lane);
    const uint4 b = __ldg(xr + i * 32 + lane);
    const int c0 = (i * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, bu[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack_bf16x2(au[j]), fb = unpack_bf16x2(bu[j]);
      g[8 * i + 2 * j] = fa.x * gm[2 * j]; g[8 * i + 2 * j + 1] = fa.y * gm[2 * j + 1];
      xh[8 * i + 2 * j] = (fb.x - mu) * rs; xh[8 * i + 2 * j + 1] = (fb.y - mu) * rs;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1 += g[8 * i + j]; s2 = fmaf(g[8 * i + j], xh[8 * i + j], s2); }
  }
  s1 = warp_sum(s1) * (1.0f / d); s2 = warp_sum(s2) * (1.0f / d);
  uint4* dxr = reinterpret_cast<uint4*>(dx + row * d);
  const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + row * d) : nullptr;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = rs * (g[8 * i + j] - s1 - xh[8 * i + j] * s2);
    if (rr) {
      const uint4 r = __ldg(rr + i * 32 + lane);
      const uint32_t ru[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(ru[j]); o[2 * j] += f.x; o[2 * j + 1] += f.y; }
    }
    uint4 u;
    u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]); u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
    dxr[i * 32 + lane] = u;
  }
}

// ------------------------------------------------------------------------------------------------ embedding
template <typename T>
__global__ void embed_kernel(int L, int d, const long long* __restrict__ ids, const T* __restrict__ E, const T* __restrict__ P,
                             int pos0, T* __restrict__ h) {
  pdl_launch_dependents();
  pdl_wait();
  const long long tok = blockIdx.x;   // b*L + l
  const int l = static_cast<int>(tok % L);
  const long long id = ids[tok];
  const T* e = E + id * d;
  const T* p = P + static_cast<long long>(pos0 + l) * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) h[tok * d + c] = from_f<T>(to_f<T>(e[c]) + to_f<T>(p[c]));
}

// ------------------------------------------------------------------------------------------------ cross entropy
__global__ void ce_count_kernel(long long rows, const long long* __restrict__ labels, int* n_valid, float* loss_sum) {
  __shared__ int sh[32];
  int c = 0;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) c += (labels[r] != -100);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
    *n_valid = t;
    if (loss_sum) *loss_sum = 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(512) ce_kernel(int V, long long ld, T* __restrict__ logits, const long long* __restrict__ labels,
                                                 float* __restrict__ row_loss, float* loss_sum, const int* __restrict__ n_valid,
                                                 int write_grad, float grad_scale) {
  __shared__ float sh[32];
  __shared__ float bc;
  const long long row = blockIdx.x;
  T* lr = logits + row * ld;
  const long long label = labels[row];
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < V; c += blockDim.x) m = fmaxf(m, to_f<T>(lr[c]));
  m = warp_max(m);
  if (lane == 0) sh[warp] = m;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? sh[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) bc = t;
  }
  __syncthreads();
  m = bc;
  float s = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x) s += __expf(to_f<T>(lr[c]) - m);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? sh[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) bc = t;
  }
  __syncthreads();
  s = bc;
  const float lse = m + logf(s);
  const bool valid = label != -100;
  if (threadIdx.x == 0) {
    const float loss = valid ? lse - to_f<T>(lr[label]) : 0.f;
    if (row_loss) row_loss[row] = loss;
    if (loss_sum && valid) atomicAdd(loss_sum, loss);
  }
  if (write_grad) {
    __syncthreads();   // label logit read above before it is overwritten
    const int nv = *n_valid;
    const float sc = (valid && nv > 0) ? grad_scale / nv : 0.f;
    const float inv = 1.0f / s;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) {
      float g = 0.f;
      if (c < V) {
        g = __expf(to_f<T>(lr[c]) - m) * inv;
        if (c == label) g -= 1.0f;
        g *= sc;
      }
      lr[c] = from_f<T>(g);
    }
  }
}


// bf16 rows with 16-byte alignment: 8 logits per load, ONE pass for (max, sum) with the online-softmax update, then (for the
// gradient) one more read + one write.  The scalar kernel above reads every logit three times, two bytes at a time.
__global__ void __launch_bounds__(512) ce_vec_kernel(int V, long long ld, __nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels,
                                                     float* __restrict__ row_loss, float* loss_sum, const int* __restrict__ n_valid,
                                                     int write_grad, float grad_scale) {
  constexpr float kL2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  __shared__ float shm[16], shs[16];
  __shared__ float bcm, bcs;
  const long long row = blockIdx.x;
  __nv_bfloat16* lr = logits + row * ld;
  uint4* lv = reinterpret_cast<uint4*>(lr);
  const long long label = labels[row];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = V / 8;                       // full vectors; the tail (< 8 logits) is handled by thread 0
  float m = -INFINITY, s = 0.f;                 // s = sum exp2((x - m) * log2e)
  auto upd = [&](float x) {
    if (x > m) { s *= exp2f((m - x) * kL2e); m = x; }
    s += exp2f((x - m) * kL2e);
  };
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 u = lv[i];
    float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
    const float vm = fmaxf(fmaxf(fmaxf(f0.x, f0.y), fmaxf(f1.x, f1.y)), fmaxf(fmaxf(f2.x, f2.y), fmaxf(f3.x, f3.y)));
    if (vm > m) { s *= exp2f((m - vm) * kL2e); m = vm; }
    const float mb = m * kL2e;
    s += exp2f(fmaf(f0.x, kL2e, -mb)) + exp2f(fmaf(f0.y, kL2e, -mb)) + exp2f(fmaf(f1.x, kL2e, -mb)) + exp2f(fmaf(f1.y, kL2e, -mb)) +
         exp2f(fmaf(f2.x, kL2e, -mb)) + exp2f(fmaf(f2.y, kL2e, -mb)) + exp2f(fmaf(f3.x, kL2e, -mb)) + exp2f(fmaf(f3.y, kL2e, -mb));
  }
  if (threadIdx.x == 0)
    for (int c = nvec * 8; c < V; ++c) upd(__bfloat162float(lr[c]));
  // combine (m, s) pairs: warp, then block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    s = (mn == -INFINITY) ? 0.f : s * exp2f((m - mn) * kL2e) + s2 * exp2f((m2 - mn) * kL2e);
    m = mn;
  }
  if (lane == 0) { shm[warp] = m; shs[warp] = s; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    m = lane < nw ? shm[lane] : -INFINITY;
    s = lane < nw ? shs[lane] : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      const float mn = fmaxf(m, m2);
      s = (mn == -INFINITY) ? 0.f : s * exp2f((m - mn) * kL2e) + s2 * exp2f((m2 - mn) * kL2e);
      m = mn;
    }
    if (lane == 0) { bcm = m; bcs = s; }
  }
  __syncthreads();
  m = bcm; s = bcs;
  const bool valid = label != -100;
  if (threadIdx.x == 0) {
    const float lse = m + log2f(s) * kLn2;
    const float loss = valid ? lse - __bfloat162float(lr[label]) : 0.f;
    if (row_loss) row_loss[row] = loss;
    if (loss_sum && valid) atomicAdd(loss_sum, loss);
  }
  if (write_grad) {
    __syncthreads();   // label logit read above before it is overwritten
    const int nv = *n_valid;
    const float sc = (valid && nv > 0) ? grad_scale / nv : 0.f;
    const float k = sc / s;
    const float mb = m * kL2e;
    const int nvec_ld = static_cast<int>(ld / 8);
    const int lab = valid ? static_cast<int>(label) : -1;
    for (int i = threadIdx.x; i < nvec_ld; i += blockDim.x) {
      const int c0 = i * 8;
      const uint4 u = lv[i];
      float x[8];
      float2 f;
      f = unpack_bf16x2(u.x); x[0] = f.x; x[1] = f.y;
      f = unpack_bf16x2(u.y); x[2] = f.x; x[3] = f.y;
      f = unpack_bf16x2(u.z); x[4] = f.x; x[5] = f.y;
      f = unpack_bf16x2(u.w); x[6] = f.x; x[7] = f.y;
      float gq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gg = (c0 + j < V) ? exp2f(fmaf(x[j], kL2e, -mb)) * k : 0.f;
        if (c0 + j == lab) gg -= sc;
        gq[j] = gg;
      }
      uint4 o;
      o.x = pack_bf16x2(gq[0], gq[1]); o.y = pack_bf16x2(gq[2], gq[3]); o.z = pack_bf16x2(gq[4], gq[5]); o.w = pack_bf16x2(gq[6], gq[7]);
      lv[i] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------ greedy pick
// One CTA per row, 16-byte loads (the scalar version with a global suppress-list lookup per element took 53 us for 128 rows of
// 51865 logits: latency-bound at 0.25 TB/s).  Ties go to the smallest index, as torch.argmax.
constexpr int kMaxSuppress = 64;
template <typename T>
__global__ void __launch_bounds__(512) greedy_kernel(int V, long long ld, const T* __restrict__ logits, const int* __restrict__ suppress,
                                                     int n_suppress, int eos, int pad, unsigned char* finished,
                                                     long long* __restrict__ next_ids, long long* __restrict__ out, long long out_ld) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float shv[32];
  __shared__ int shi[32];
  __shared__ int s_sup[kMaxSuppress];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const T* lr = logits + static_cast<long long>(b) * ld;
  if (threadIdx.x < n_suppress) s_sup[threadIdx.x] = suppress[threadIdx.x];
  __syncthreads();
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(lr) & 15) == 0;
  const int Vv = vec_ok ? (V / VEC) * VEC : 0;
  for (int c0 = threadIdx.x * VEC; c0 < Vv; c0 += blockDim.x * VEC) {
    float v[VEC];
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(lr + c0));
    if constexpr (sizeof(T) == 2) {
      float2 f;
      f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y; f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
      f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y; f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
    } else {
      v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
    for (int s = 0; s < n_suppress; ++s) {
      const unsigned dlt = static_cast<unsigned>(s_sup[s] - c0);
      if (dlt < static_cast<unsigned>(VEC)) {
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (static_cast<unsigned>(j) == dlt) v[j] = -INFINITY;
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j)
      if (v[j] > best) { best = v[j]; bi = c0 + j; }            // ascending indices inside a thread: strict > keeps the first
  }
  for (int c = Vv + threadIdx.x; c < V; c += blockDim.x) {      // tail (and rows that are not 16-byte aligned)
    float v = to_f<T>(lr[c]);
    for (int s = 0; s < n_suppress; ++s)
      if (s_sup[s] == c) v = -INFINITY;
    if (v > best || (v == best && c < bi)) { best = v; bi = c; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { shv[warp] = best; shi[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < nw ? shv[lane] : -INFINITY;
    bi = lane < nw ? shi[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
      int tok = bi;
      if (finished) {
        if (finished[b]) tok = pad;
        else if (tok == eos) finished[b] = 1;
      }
      next_ids[b] = tok;
      if (out) out[b * out_ld] = tok;
    }
  }
}

// ------------------------------------------------------------------------------------------------ augmentation pass
__device__ __forceinline__ void store8(__nv_bfloat16* dst, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  uint4 u;
  u.x = pack_bf16x2(a0, a1); u.y = pack_bf16x2(a2, a3); u.z = pack_bf16x2(a4, a5); u.w = pack_bf16x2(a6, a7);
  *reinterpret_cast<uint4*>(dst) = u;
}
__device__ __forceinline__ void store8(float* dst, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  reinterpret_cast<float4*>(dst)[0] = make_float4(a0, a1, a2, a3);
  reinterpret_cast<float4*>(dst)[1] = make_float4(a4, a5, a6, a7);
}
__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t* hi) {
  const unsigned long long p = static_cast<unsigned long long>(a) * b;
  *hi = static_cast<uint32_t>(p >> 32);
  return static_cast<uint32_t>(p);
}
// Philox4x32-10 -> one standard normal (Box-Muller on the first two words)
__device__ float philox_normal(unsigned long long seed, unsigned long long idx) {
  uint32_t c0 = static_cast<uint32_t>(idx), c1 = static_cast<uint32_t>(idx >> 32), c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, h1;
    const uint32_t l0 = mulhilo(0xD2511F53u, c0, &h0);
    const uint32_t l1 = mulhilo(0xCD9E8D57u, c2, &h1);
    const uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float u1 = (static_cast<float>(c0 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (static_cast<float>(c1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

struct AugDev {
  int B, C, Tin, T, Cp;
  const int* n; const int* shift; const int* e0; const int* e1; const int* flags;
  const unsigned char* grid; long long grid_stride; const int* gl; const int* rep_c; const int* rep_t;
  const float* sigma; unsigned long long seed;
  const long long* src_off; const int* src_ld;     // ragged sample store: row c of sample b starts at x + src_off[b] + c * src_ld[b]
};
// source row of (sample b, channel c) and its readable length
__device__ __forceinline__ long long aug_row(const AugDev& a, int b, int c, int& ld) {
  if (a.src_off) { ld = a.src_ld[b]; return a.src_off[b] + static_cast<long long>(c) * ld; }
  ld = a.Tin;
  return (static_cast<long long>(b) * a.C + c) * a.Tin;
}
template <typename TI> __device__ __forceinline__ void load4(const TI* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 f = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <typename TI>
__device__ __forceinline__ float aug_value(const AugDev& a, const TI* __restrict__ x, int b, int c, int tau) {
  const int sh = a.shift ? a.shift[b] : 0;
  int ld;
  const long long row = aug_row(a, b, c, ld);
  const int n = a.n ? min(a.n[b], ld) : ld;
  const int t = tau - sh;
  if (t < 0 || t >= n) return 0.f;
  float v = to_f<TI>(x[row + t]);
  const int fl = a.flags ? a.flags[b] : 0;
  if (fl & 2) {   // gaussian noise: reference returns signal + (signal + noise)
    const float sg = a.sigma[static_cast<long long>(b) * a.C + c];
    v = 2.0f * v + sg * philox_normal(a.seed, (static_cast<unsigned long long>(b) * a.C + c) * a.Tin + t);
  }
  if ((fl & 1) && a.grid) {
    const int gc = c / a.rep_c[b], gt = t / a.rep_t[b];
    if (!a.grid[b * a.grid_stride + static_cast<long long>(gc) * a.gl[b] + gt]) v = 0.f;
  }
  const int e0 = a.e0 ? a.e0[b] : 0, e1 = a.e1 ? a.e1[b] : 0;
  if (t < e0 || t >= n - e1) v = 0.f;
  return v;
}

// layout 0: (B,C,T) -> (B,C,T)
template <typename TI, typename TO>
__global__ void aug_bct_kernel(const AugDev a, const TI* __restrict__ x, TO* __restrict__ y) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int tau = blockIdx.x * blockDim.x + threadIdx.x;
  if (tau >= a.T) return;
  y[(static_cast<long long>(b) * a.C + c) * a.T + tau] = from_f<TO>(aug_value(a, x, b, c, tau));
}

// layout 1: (B,C,T) -> (B,T,Cp) channels-last through a 128(t) x 64(c) shared tile; pad channels written as zero.
// HBM-bound (4 B in + 2 B out per element): every lane reads 16 bytes (four consecutive samples of one channel, a warp covers
// 512 contiguous bytes) and writes 16 bytes (eight channels of one sample).  The per-sample decisions (length, shift, edge
// zeroing, mask grid geometry) are read once per block; the grid cell is looked up once per run of samples inside it.
constexpr int kAugTT = 128, kAugTC = 64;
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) aug_btc_kernel(const AugDev a, const TI* __restrict__ x, TO* __restrict__ y) {
  // tile[channel][sample], 16-byte groups of 4 samples XOR-swizzled with (channel / 8): the 16-byte stores of the read phase
  // and the scalar reads of the write phase (lanes = 4 samples x 8 channel groups) are both bank-conflict free
  __shared__ __align__(16) float tile[kAugTC][kAugTT];
  auto at = [&](int cc, int t) -> float& { return tile[cc][((((t >> 2) ^ (cc >> 3)) & 31) << 2) | (t & 3)]; };
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kAugTC;
  const int t0 = blockIdx.x * kAugTT;
  const int sh = a.shift ? a.shift[b] : 0;
  int sld;
  const long long row0 = aug_row(a, b, 0, sld);
  const int n = a.n ? min(a.n[b], sld) : sld;
  const int fl = a.flags ? a.flags[b] : 0;
  const int e0 = a.e0 ? a.e0[b] : 0, e1 = a.e1 ? a.e1[b] : 0;
  const int lo = max(e0, 0), hi = min(n, n - e1);            // samples outside [lo, hi) of the source are zero
  const bool masked = (fl & 1) && a.grid && n > 0;
  const int rep_c = masked ? a.rep_c[b] : 1, rep_t = masked ? a.rep_t[b] : 1, gl = masked ? a.gl[b] : 1;
  const unsigned char* grid = masked ? a.grid + b * a.grid_stride : nullptr;
  // vector loads (4 samples) need the SOURCE index t = tau - shift of a lane's first sample to be a multiple of 4 and rows that
  // start on a multiple of 4 elements
  const bool vec = ((sh & 3) == 0) && ((sld & 3) == 0) && ((row0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (int e = threadIdx.x; e < kAugTC * (kAugTT / 4); e += 256) {
    const int cc = e / (kAugTT / 4), t4 = (e % (kAugTT / 4)) * 4;
    const int c = c0 + cc, tau = t0 + t4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int t = tau - sh;
    if (c < a.C && t + 3 >= lo && t < hi) {
      const TI* xr = x + row0 + static_cast<long long>(c) * sld;
      if (vec && t >= 0 && t + 3 < sld) {
        load4<TI>(xr + t, v);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (t + k >= 0 && t + k < n) v[k] = to_f<TI>(xr[t + k]);
      }
      if (fl & 2) {   // gaussian noise: the reference returns signal + (signal + noise)
        const float sg = a.sigma[static_cast<long long>(b) * a.C + c];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (t + k >= 0 && t + k < n)
            v[k] = 2.0f * v[k] + sg * philox_normal(a.seed, (static_cast<unsigned long long>(b) * a.C + c) * a.Tin + (t + k));
      }
      if (masked) {
        const unsigned char* grow = grid + static_cast<long long>(c / rep_c) * gl;
        const int tb = max(t, 0);
        int gq = tb / rep_t, gr = tb - gq * rep_t;
        bool keep = grow[gq] != 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (t + k < tb) continue;                           // before the signal starts: already zero
          if (t + k > tb) { if (++gr == rep_t) { gr = 0; ++gq; keep = (t + k < n) ? grow[gq] != 0 : false; } }
          if (!keep) v[k] = 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t + k < lo || t + k >= hi) v[k] = 0.f;
    }
    *reinterpret_cast<float4*>(&at(cc, t4)) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
  // write: 8 channels (16 bytes of bf16, 32 of fp32) of one sample per lane; the 64 channels of a sample are contiguous
  const bool vec_out = (a.Cp % 8 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  for (int e = threadIdx.x; e < kAugTT * (kAugTC / 8); e += 256) {
    const int tt = e / (kAugTC / 8), ch = (e % (kAugTC / 8)) * 8;
    const int c = c0 + ch, tau = t0 + tt;
    if (tau >= a.T || c >= a.Cp) continue;
    TO* dst = y + (static_cast<long long>(b) * a.T + tau) * a.Cp + c;
    if (vec_out && c + 8 <= a.Cp) {
      store8(dst, at(ch, tt), at(ch + 1, tt), at(ch + 2, tt), at(ch + 3, tt), at(ch + 4, tt), at(ch + 5, tt), at(ch + 6, tt), at(ch + 7, tt));
    } else {
      for (int k = 0; k < 8 && c + k < a.Cp; ++k) dst[k] = from_f<TO>(at(ch + k, tt));
    }
  }
}

template <typename TI>
__global__ void channel_meansq_kernel(int C, int Tin, const int* __restrict__ n, const TI* __restrict__ x, float* __restrict__ ms,
                                      const long long* __restrict__ src_off, const int* __restrict__ src_ld) {
  __shared__ float sh[8];
  const int b = blockIdx.y, c = blockIdx.x;
  const int ld = src_off ? src_ld[b] : Tin;
  const int len = n ? min(n[b], ld) : ld;
  const TI* xr = x + (src_off ? src_off[b] + static_cast<long long>(c) * ld : (static_cast<long long>(b) * C + c) * Tin);
  float s = 0.f;
  for (int t = threadIdx.x; t < len; t += blockDim.x) { const float v = to_f<TI>(xr[t]); s = fmaf(v, v, s); }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
    ms[static_cast<long long>(b) * C + c] = len > 0 ? t / len : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ utilities
template <typename TS, typename TD>
__global__ void cast_kernel(long long n, const TS* __restrict__ s, TD* __restrict__ d) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    d[i] = from_f<TD>(to_f<TS>(s[i]));
}

template <typename TS, typename TD>
__global__ void transpose_kernel(int rows, int cols, const TS* __restrict__ s, long long lds, TD* __restrict__ d, long long ldd, float scale) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? to_f<TS>(s[static_cast<long long>(r) * lds + c]) * scale : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;   // dst[c][r]
    if (c < cols && r < ldd) d[static_cast<long long>(c) * ldd + r] = from_f<TD>(tile[tx][i]);
  }
}

// Many small transposes in one launch (the per-step refresh of the LoRA A/B operand layouts): blockIdx.y picks the job, the
// blocks of a job stride over ITS 32x32 tiles.  (A grid of max_rows x max_cols tiles per job launched 245 k blocks for the 60
// rank-32 operands of the training step -- (32, 2048) and (2048, 32) share no tile beyond the first row / column -- and spent
// 157 us dispatching blocks that returned at once.)
template <typename TS, typename TD>
__global__ void transpose_batched_kernel(const ns_transpose_job* __restrict__ jobs) {
  __shared__ float tile[32][33];
  const ns_transpose_job j = jobs[blockIdx.y];
  const TS* s = reinterpret_cast<const TS*>(j.src);
  TD* d = reinterpret_cast<TD*>(j.dst);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tiles_c = (j.cols + 31) / 32, tiles_r = (static_cast<int>(j.ldd) + 31) / 32;
  for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {      // block-uniform trip count
    const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      tile[i][tx] = (r < j.rows && c < j.cols) ? to_f<TS>(s[static_cast<long long>(r) * j.lds + c]) * j.scale : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;   // dst[c][r]
      if (c < j.cols && r < j.ldd) d[static_cast<long long>(c) * j.ldd + r] = from_f<TD>(tile[tx][i]);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void conv_pack_kernel(int N, int C, int Cp, const float* __restrict__ w, T* __restrict__ wt, T* __restrict__ wtt) {
  const long long total = 3LL * N * Cp;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cp);
    const int n = static_cast<int>((i / Cp) % N);
    const int k = static_cast<int>(i / (static_cast<long long>(Cp) * N));
    const float v = c < C ? w[(static_cast<long long>(n) * C + c) * 3 + k] : 0.f;
    if (wt) wt[i] = from_f<T>(v);
    if (wtt) wtt[(static_cast<long long>(k) * Cp + c) * N + n] = from_f<T>(v);
  }
}
__global__ void conv_unpack_grad_kernel(int N, int C, int Cp, const float* __restrict__ dwt, float* __restrict__ dw) {
  const long long total = static_cast<long long>(N) * C * 3;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % 3);
    const int c = static_cast<int>((i / 3) % C);
    const int n = static_cast<int>(i / (3LL * C));
    dw[i] = dwt[(static_cast<long long>(k) * N + n) * Cp + c];
  }
}

template <typename T>
__global__ void add_kernel(long long n, const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = from_f<T>(to_f<T>(a[i]) + to_f<T>(b[i]));
}

template <typename T>
__global__ void dgelu_mul_kernel(long long n, const T* __restrict__ dy, const T* __restrict__ z, T* __restrict__ dz) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    dz[i] = from_f<T>(to_f<T>(dy[i]) * dgelu_erf(to_f<T>(z[i])));
}
// bf16 storage, 16-byte aligned, n % 8 == 0: eight elements per thread and trip, the derivative in the fitted-tanh form every bf16
// GEMM epilogue uses (ns_common.cuh dgelu_fast: 1.3e-4 from the erf form, below a bf16 ulp).  The scalar erf loop above took
// 108 us for the stem's (64, 1500, 512) gradient, 2.4x the time of its 295 MB of traffic.
__global__ void __launch_bounds__(256) dgelu_mul_bf16x8_kernel(long long n8, const uint4* __restrict__ dy, const uint4* __restrict__ z,
                                                                uint4* __restrict__ dz) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 a = __ldg(dy + i), b = __ldg(z + i);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&av[k]));
      const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bv[k]));
      const __nv_bfloat162 r = __floats2bfloat162_rn(g.x * dgelu_fast(x.x), g.y * dgelu_fast(x.y));
      o[k] = *reinterpret_cast<const uint32_t*>(&r);
    }
    dz[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void sumsq_kernel(long long n, const float* __restrict__ g, float* out) {
  __shared__ float sh[32];
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    s = fmaf(g[i], g[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

__global__ void adamw_kernel(long long n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             const float* __restrict__ sumsq, float gscale, float max_norm, float lr, float b1, float b2, float eps,
                             float wd, float bc1, float bc2_sqrt) {
  float coef = gscale;
  if (sumsq && max_norm > 0.f) {
    const float norm = sqrtf(*sumsq) * fabsf(gscale);
    coef *= fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gr = g[i] * coef;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gr;
    const float vi = b2 * v[i] + (1.0f - b2) * gr * gr;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

static int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace ns

using namespace ns;
typedef __nv_bfloat16 bf16;

extern "C" {

int ns_layernorm_fwd(int dtype, long long rows, int d, const void* x, const float* gamma, const float* beta, void* y,
                     float* mean, float* rstd, float eps, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && rows >= 0 && d > 0 && x && gamma && beta && y, "ns_layernorm_fwd: bad arguments");
  if (rows == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                    reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  if (dtype == NS_BF16 && al && d % 256 == 0 && d <= 1280) {
    const bf16* xi = reinterpret_cast<const bf16*>(x);
    bf16* yo = reinterpret_cast<bf16*>(y);
    const unsigned grid = static_cast<unsigned>((rows + 15) / 16);    // 8 warps x 2 rows per block
    switch (d / 256) {
      case 1: NS_CUDA(launch_pdl(ln_fwd_bf16_kernel<1>, grid, 256, 0, st, rows, xi, gamma, beta, yo, mean, rstd, eps)); break;
      case 2: NS_CUDA(launch_pdl(ln_fwd_bf16_kernel<2>, grid, 256, 0, st, rows, xi, gamma, beta, yo, mean, rstd, eps)); break;
      case 3: NS_CUDA(launch_pdl(ln_fwd_bf16_kernel<3>, grid, 256, 0, st, rows, xi, gamma, beta, yo, mean, rstd, eps)); break;
      case 4: NS_CUDA(launch_pdl(ln_fwd_bf16_kernel<4>, grid, 256, 0, st, rows, xi, gamma, beta, yo, mean, rstd, eps)); break;
      default: NS_CUDA(launch_pdl(ln_fwd_bf16_kernel<5>, grid, 256, 0, st, rows, xi, gamma, beta, yo, mean, rstd, eps)); break;
    }
  } else if (dtype == NS_BF16) {
    ln_fwd_generic_kernel<bf16><<<grid, 256, 0, st>>>(rows, d, reinterpret_cast<const bf16*>(x), gamma, beta,
                                                      reinterpret_cast<bf16*>(y), mean, rstd, eps);
  } else {
    ln_fwd_generic_kernel<float><<<grid, 256, 0, st>>>(rows, d, reinterpret_cast<const float*>(x), gamma, beta,
                                                       reinterpret_cast<float*>(y), mean, rstd, eps);
  }
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_layernorm_bwd(int dtype, long long rows, int d, const void* dy, const void* x, const float* gamma,
                     const float* mean, const float* rstd, const void* dres, void* dx, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && rows >= 0 && d > 0 && dy && x && gamma && mean && rstd && dx, "ns_layernorm_bwd: bad arguments");
  if (rows == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                    reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(dres)) & 15) == 0;
  if (dtype == NS_BF16 && al && d % 256 == 0 && d <= 1280) {
    const bf16* a = reinterpret_cast<const bf16*>(dy);
    const bf16* b = reinterpret_cast<const bf16*>(x);
    const bf16* r = reinterpret_cast<const bf16*>(dres);
    bf16* o = reinterpret_cast<bf16*>(dx);
    switch (d / 256) {
      case 1: ln_bwd_bf16_kernel<1><<<grid, 256, 0, st>>>(rows, a, b, gamma, mean, rstd, r, o); break;
      case 2: ln_bwd_bf16_kernel<2><<<grid, 256, 0, st>>>(rows, a, b, gamma, mean, rstd, r, o); break;
      case 3: ln_bwd_bf16_kernel<3><<<grid, 256, 0, st>>>(rows, a, b, gamma, mean, rstd, r, o); break;
      case 4: ln_bwd_bf16_kernel<4><<<grid, 256, 0, st>>>(rows, a, b, gamma, mean, rstd, r, o); break;
      default: ln_bwd_bf16_kernel<5><<<grid, 256, 0, st>>>(rows, a, b, gamma, mean, rstd, r, o); break;
    }
  } else if (dtype == NS_BF16) {
    ln_bwd_kernel<bf16><<<grid, 256, 0, st>>>(rows, d, reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(x), gamma,
                                              mean, rstd, reinterpret_cast<const bf16*>(dres), reinterpret_cast<bf16*>(dx));
  } else {
    ln_bwd_kernel<float><<<grid, 256, 0, st>>>(rows, d, reinterpret_cast<const float*>(dy), reinterpret_cast<const float*>(x),
                                               gamma, mean, rstd, reinterpret_cast<const float*>(dres), reinterpret_cast<float*>(dx));
  }
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_embed(int dtype, int B, int L, int d, const long long* ids, const void* E, const void* P, int pos0, void* h, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && B >= 0 && L >= 0 && d > 0 && ids && E && P && h && pos0 >= 0, "ns_embed: bad arguments");
  if (B * L == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == NS_BF16)
    NS_CUDA(launch_pdl(embed_kernel<bf16>, dim3(B * L), dim3(128), 0, st, L, d, ids, reinterpret_cast<const bf16*>(E), reinterpret_cast<const bf16*>(P), pos0,
                       reinterpret_cast<bf16*>(h)));
  else
    NS_CUDA(launch_pdl(embed_kernel<float>, dim3(B * L), dim3(128), 0, st, L, d, ids, reinterpret_cast<const float*>(E), reinterpret_cast<const float*>(P), pos0,
                       reinterpret_cast<float*>(h)));
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_cross_entropy(int dtype, long long rows, int V, long long ld, void* logits, const long long* labels, float* row_loss,
                     float* loss_sum_out, int* n_valid_out, int write_grad, float grad_scale, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && rows >= 0 && V > 0 && ld >= V && logits && labels && n_valid_out, "ns_cross_entropy: bad arguments");
  if (rows == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ce_count_kernel<<<1, 256, 0, st>>>(rows, labels, n_valid_out, loss_sum_out);
  NS_LAUNCH_CHECK();
  if (dtype == NS_BF16 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0)
    ce_vec_kernel<<<static_cast<unsigned>(rows), 512, 0, st>>>(V, ld, reinterpret_cast<bf16*>(logits), labels, row_loss, loss_sum_out, n_valid_out, write_grad, grad_scale);
  else if (dtype == NS_BF16)
    ce_kernel<bf16><<<static_cast<unsigned>(rows), 512, 0, st>>>(V, ld, reinterpret_cast<bf16*>(logits), labels, row_loss, loss_sum_out, n_valid_out, write_grad, grad_scale);
  else
    ce_kernel<float><<<static_cast<unsigned>(rows), 512, 0, st>>>(V, ld, reinterpret_cast<float*>(logits), labels, row_loss, loss_sum_out, n_valid_out, write_grad, grad_scale);
  NS_LAUNCH_CHECK();
  count(C_OTHER, 2);
  return NS_OK;
}

int ns_greedy_pick(int dtype, int B, int V, long long ld, const void* logits, const int* suppress, int n_suppress, int eos,
                   int pad, unsigned char* finished, long long* next_ids, long long* out, long long out_ld, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && B >= 0 && V > 0 && ld >= V && logits && next_ids && (n_suppress == 0 || suppress), "ns_greedy_pick: bad arguments");
  NS_CHECK_ARG(n_suppress >= 0 && n_suppress <= kMaxSuppress, "ns_greedy_pick: at most %d suppressed ids", kMaxSuppress);
  if (B == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == NS_BF16)
    NS_CUDA(launch_pdl(greedy_kernel<bf16>, dim3(B), dim3(512), 0, st, V, ld, reinterpret_cast<const bf16*>(logits), suppress, n_suppress, eos, pad,
                       finished, next_ids, out, out_ld));
  else
    NS_CUDA(launch_pdl(greedy_kernel<float>, dim3(B), dim3(512), 0, st, V, ld, reinterpret_cast<const float*>(logits), suppress, n_suppress, eos, pad,
                       finished, next_ids, out, out_ld));
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_aug_pass(const ns_aug_args* a, const void* x, void* y, void* stream) {
  NS_CHECK_ARG(a && x && y, "ns_aug_pass: null argument");
  NS_CHECK_ARG(a->B > 0 && a->C > 0 && a->Tin > 0 && a->T > 0 && valid_dtype(a->out_dtype) && valid_dtype(a->in_dtype), "ns_aug_pass: bad shape");
  NS_CHECK_ARG(a->layout == 0 || (a->layout == 1 && a->Cp >= a->C), "ns_aug_pass: bad layout / Cp");
  NS_CHECK_ARG(!a->grid || (a->gl && a->rep_c && a->rep_t && a->flags), "ns_aug_pass: grid needs gl/rep_c/rep_t/flags");
  NS_CHECK_ARG((a->src_off == nullptr) == (a->src_ld == nullptr) && (!a->src_off || a->n), "ns_aug_pass: src_off, src_ld and n go together");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  AugDev d{a->B, a->C, a->Tin, a->T, a->Cp, a->n, a->shift, a->e0, a->e1, a->flags, a->grid, a->grid_stride, a->gl,
           a->rep_c, a->rep_t, a->sigma, a->seed, a->src_off, a->src_ld};
  const bool ibf = a->in_dtype == NS_BF16, obf = a->out_dtype == NS_BF16;
  const float* xf = static_cast<const float*>(x);
  const bf16* xb = static_cast<const bf16*>(x);
  if (a->layout == 0) {
    dim3 grid((a->T + 255) / 256, a->C, a->B);
    if (ibf && obf) aug_bct_kernel<bf16, bf16><<<grid, 256, 0, st>>>(d, xb, reinterpret_cast<bf16*>(y));
    else if (ibf) aug_bct_kernel<bf16, float><<<grid, 256, 0, st>>>(d, xb, reinterpret_cast<float*>(y));
    else if (obf) aug_bct_kernel<float, bf16><<<grid, 256, 0, st>>>(d, xf, reinterpret_cast<bf16*>(y));
    else aug_bct_kernel<float, float><<<grid, 256, 0, st>>>(d, xf, reinterpret_cast<float*>(y));
  } else {
    dim3 grid((a->T + kAugTT - 1) / kAugTT, (a->Cp + kAugTC - 1) / kAugTC, a->B);
    if (ibf && obf) aug_btc_kernel<bf16, bf16><<<grid, 256, 0, st>>>(d, xb, reinterpret_cast<bf16*>(y));
    else if (ibf) aug_btc_kernel<bf16, float><<<grid, 256, 0, st>>>(d, xb, reinterpret_cast<float*>(y));
    else if (obf) aug_btc_kernel<float, bf16><<<grid, 256, 0, st>>>(d, xf, reinterpret_cast<bf16*>(y));
    else aug_btc_kernel<float, float><<<grid, 256, 0, st>>>(d, xf, reinterpret_cast<float*>(y));
  }
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_channel_meansq(int dtype, int B, int C, int Tin, const int* n, const void* x, float* ms, const long long* src_off, const int* src_ld,
                      void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && B > 0 && C > 0 && Tin > 0 && x && ms, "ns_channel_meansq: bad arguments");
  NS_CHECK_ARG((src_off == nullptr) == (src_ld == nullptr) && (!src_off || n), "ns_channel_meansq: src_off, src_ld and n go together");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == NS_BF16) channel_meansq_kernel<bf16><<<dim3(C, B), 256, 0, st>>>(C, Tin, n, static_cast<const bf16*>(x), ms, src_off, src_ld);
  else channel_meansq_kernel<float><<<dim3(C, B), 256, 0, st>>>(C, Tin, n, static_cast<const float*>(x), ms, src_off, src_ld);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_cast(int sdt, int ddt, long long n, const void* src, void* dst, void* stream) {
  NS_CHECK_ARG(valid_dtype(sdt) && valid_dtype(ddt) && n >= 0 && src && dst, "ns_cast: bad arguments");
  if (n == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int g = grid_for(n, 256);
  if (sdt == NS_F32 && ddt == NS_BF16) cast_kernel<float, bf16><<<g, 256, 0, st>>>(n, reinterpret_cast<const float*>(src), reinterpret_cast<bf16*>(dst));
  else if (sdt == NS_BF16 && ddt == NS_F32) cast_kernel<bf16, float><<<g, 256, 0, st>>>(n, reinterpret_cast<const bf16*>(src), reinterpret_cast<float*>(dst));
  else if (sdt == NS_F32) cast_kernel<float, float><<<g, 256, 0, st>>>(n, reinterpret_cast<const float*>(src), reinterpret_cast<float*>(dst));
  else cast_kernel<bf16, bf16><<<g, 256, 0, st>>>(n, reinterpret_cast<const bf16*>(src), reinterpret_cast<bf16*>(dst));
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_transpose(int sdt, int ddt, int rows, int cols, const void* src, long long lds, void* dst, long long ldd, float scale, void* stream) {
  NS_CHECK_ARG(valid_dtype(sdt) && valid_dtype(ddt) && rows > 0 && cols > 0 && src && dst && lds >= cols && ldd >= rows, "ns_transpose: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((cols + 31) / 32, static_cast<unsigned>((ldd + 31) / 32));
  if (sdt == NS_F32 && ddt == NS_BF16) transpose_kernel<float, bf16><<<grid, 256, 0, st>>>(rows, cols, reinterpret_cast<const float*>(src), lds, reinterpret_cast<bf16*>(dst), ldd, scale);
  else if (sdt == NS_F32 && ddt == NS_F32) transpose_kernel<float, float><<<grid, 256, 0, st>>>(rows, cols, reinterpret_cast<const float*>(src), lds, reinterpret_cast<float*>(dst), ldd, scale);
  else if (sdt == NS_BF16 && ddt == NS_BF16) transpose_kernel<bf16, bf16><<<grid, 256, 0, st>>>(rows, cols, reinterpret_cast<const bf16*>(src), lds, reinterpret_cast<bf16*>(dst), ldd, scale);
  else transpose_kernel<bf16, float><<<grid, 256, 0, st>>>(rows, cols, reinterpret_cast<const bf16*>(src), lds, reinterpret_cast<float*>(dst), ldd, scale);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_transpose_batched(int sdt, int ddt, int n_jobs, int max_rows_pad, int max_cols, const ns_transpose_job* jobs_device, void* stream) {
  NS_CHECK_ARG(valid_dtype(sdt) && valid_dtype(ddt) && n_jobs >= 0 && n_jobs <= 65535 && max_rows_pad > 0 && max_cols > 0 && jobs_device,
               "ns_transpose_batched: bad arguments");
  if (n_jobs == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long tiles = static_cast<long long>((max_cols + 31) / 32) * ((max_rows_pad + 31) / 32);
  dim3 grid(static_cast<unsigned>(tiles < 256 ? tiles : 256), n_jobs);
  if (sdt == NS_F32 && ddt == NS_BF16) transpose_batched_kernel<float, bf16><<<grid, 256, 0, st>>>(jobs_device);
  else if (sdt == NS_F32 && ddt == NS_F32) transpose_batched_kernel<float, float><<<grid, 256, 0, st>>>(jobs_device);
  else if (sdt == NS_BF16 && ddt == NS_BF16) transpose_batched_kernel<bf16, bf16><<<grid, 256, 0, st>>>(jobs_device);
  else transpose_batched_kernel<bf16, float><<<grid, 256, 0, st>>>(jobs_device);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_conv_weight_pack(int dtype, int N, int C, int Cp, const float* w, void* w_tap, void* w_tap_t, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && N > 0 && C > 0 && Cp >= C && w && (w_tap || w_tap_t), "ns_conv_weight_pack: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int g = grid_for(3LL * N * Cp, 256);
  if (dtype == NS_BF16) conv_pack_kernel<bf16><<<g, 256, 0, st>>>(N, C, Cp, w, reinterpret_cast<bf16*>(w_tap), reinterpret_cast<bf16*>(w_tap_t));
  else conv_pack_kernel<float><<<g, 256, 0, st>>>(N, C, Cp, w, reinterpret_cast<float*>(w_tap), reinterpret_cast<float*>(w_tap_t));
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_conv_weight_unpack_grad(int N, int C, int Cp, const float* dw_tap, float* dw, void* stream) {
  NS_CHECK_ARG(N > 0 && C > 0 && Cp >= C && dw_tap && dw, "ns_conv_weight_unpack_grad: bad arguments");
  conv_unpack_grad_kernel<<<grid_for(3LL * N * C, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(N, C, Cp, dw_tap, dw);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_add(int dtype, long long n, const void* a, const void* b, void* y, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && n >= 0 && a && b && y, "ns_add: bad arguments");
  if (n == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == NS_BF16) add_kernel<bf16><<<grid_for(n, 256), 256, 0, st>>>(n, reinterpret_cast<const bf16*>(a), reinterpret_cast<const bf16*>(b), reinterpret_cast<bf16*>(y));
  else add_kernel<float><<<grid_for(n, 256), 256, 0, st>>>(n, reinterpret_cast<const float*>(a), reinterpret_cast<const float*>(b), reinterpret_cast<float*>(y));
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_dgelu_mul(int dtype, long long n, const void* dy, const void* z, void* dz, void* stream) {
  NS_CHECK_ARG(valid_dtype(dtype) && n >= 0 && dy && z && dz, "ns_dgelu_mul: bad arguments");
  if (n == 0) return NS_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = dtype == NS_BF16 && n % 8 == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0;
  if (vec) dgelu_mul_bf16x8_kernel<<<grid_for(n / 8, 256), 256, 0, st>>>(n / 8, reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(z), reinterpret_cast<uint4*>(dz));
  else if (dtype == NS_BF16) dgelu_mul_kernel<bf16><<<grid_for(n, 256), 256, 0, st>>>(n, reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(z), reinterpret_cast<bf16*>(dz));
  else dgelu_mul_kernel<float><<<grid_for(n, 256), 256, 0, st>>>(n, reinterpret_cast<const float*>(dy), reinterpret_cast<const float*>(z), reinterpret_cast<float*>(dz));
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_sumsq(long long n, const float* g, float* out, void* stream) {
  NS_CHECK_ARG(n >= 0 && g && out, "ns_sumsq: bad arguments");
  if (n == 0) return NS_OK;
  sumsq_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(n, g, out);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

int ns_adamw_clip(long long n, float* p, const float* g, float* m, float* v, const float* sumsq, float gscale, float max_norm,
                  float lr, float beta1, float beta2, float eps, float wd, int step, void* stream) {
  NS_CHECK_ARG(n >= 0 && p && g && m && v && step >= 1, "ns_adamw_clip: bad arguments");
  if (n == 0) return NS_OK;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2s = sqrtf(1.0f - powf(beta2, static_cast<float>(step)));
  adamw_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(n, p, g, m, v, sumsq, gscale, max_norm, lr,
                                                                                      beta1, beta2, eps, wd, bc1, bc2s);
  NS_LAUNCH_CHECK();
  count(C_OTHER);
  return NS_OK;
}

}  // extern "C"
