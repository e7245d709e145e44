// Shared device helpers for the CAT-Seg B200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CATSEG_HID 128   // aggregator hidden dim (configs/vitb_384.yaml:25) — kernels are specialised for it
#define CATSEG_HEADS 4
#define CATSEG_HD 32

namespace catseg {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// LayerNorm over 128 features held as 4 per lane across one warp (eps 1e-5, biased variance),
// matching torch.nn.LayerNorm(128) (model.py:152,158,233,368-369).
__device__ __forceinline__ float4 warp_layernorm128(float4 x, float4 gamma, float4 beta) {
  float mean = warp_sum(x.x + x.y + x.z + x.w) * (1.0f / 128.0f);
  float dx = x.x - mean, dy = x.y - mean, dz = x.z - mean, dw = x.w - mean;
  float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / 128.0f);
  float rstd = 1.0f / sqrtf(var + 1e-5f);
  return make_float4(dx * rstd * gamma.x + beta.x, dy * rstd * gamma.y + beta.y,
                     dz * rstd * gamma.z + beta.z, dw * rstd * gamma.w + beta.w);
}

// Same, for the bf16 paths: MUFU rsqrt instead of the IEEE sqrt + divide sequence (the result is rounded to bf16).
__device__ __forceinline__ float4 warp_layernorm128_fast(float4 x, float4 gamma, float4 beta) {
  float mean = warp_sum(x.x + x.y + x.z + x.w) * (1.0f / 128.0f);
  float dx = x.x - mean, dy = x.y - mean, dz = x.z - mean, dw = x.w - mean;
  float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / 128.0f);
  float rstd = rsqrtf(var + 1e-5f);
  return make_float4(fmaf(dx * rstd, gamma.x, beta.x), fmaf(dy * rstd, gamma.y, beta.y),
                     fmaf(dz * rstd, gamma.z, beta.z), fmaf(dw * rstd, gamma.w, beta.w));
}

__device__ __forceinline__ float gelu_erf(float x) {   // nn.GELU() default (exact erf), model.py:139
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float elu1(float x) {        // elu(x) + 1, model.py:256-257
  return x > 0.0f ? x + 1.0f : expf(x);
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

}  // namespace catseg
