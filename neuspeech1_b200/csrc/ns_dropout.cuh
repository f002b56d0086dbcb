// Counter-based mask of the LoRA-branch dropout (finetune.py:210; specification: oracle/whisper_eeg.py lora_dropout_plane).
// Shared by the plane generator (ns_lora.cu ns_dropout_bits) and the mask stage of the rank-32 down product that draws its
// plane itself (ns_gemm_sm100.cu, ns_epilogue::drop_mode 2).
#pragma once
#include <stdint.h>

namespace ns {

__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
constexpr uint32_t kRowMul = 0x9E3779B1u, kColMul = 0x85EBCA77u, kIdxMul = 0xC2B2AE35u;
// the 32 dropped flags of (row, 32-column block): see the header of ns_lora.cu
__device__ __forceinline__ uint32_t mix1(uint32_t x) { x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; return x; }
__device__ __forceinline__ uint32_t drop_plane_word(uint32_t row, uint32_t w, uint32_t module_seed, uint32_t thr) {
  const uint32_t km = lowbias32((row * kRowMul) ^ (w * kColMul) ^ module_seed);
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t r = mix1(km + static_cast<uint32_t>(i + 1) * kIdxMul);
    d = ((thr >> i) & 1u) ? (d | r) : (d & r);
  }
  return d;
}
// P(dropped) = thr / 65536
inline uint32_t drop_thr16(float p) {
  const double v = static_cast<double>(p) * 65536.0 + 0.5;
  return v <= 0 ? 0u : (v >= 65535.0 ? 65535u : static_cast<uint32_t>(v));
}

}  // namespace ns
