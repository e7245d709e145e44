#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__global__ void k_scalar(float* out, int iters, long long* cyc) {
  float a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
  float t = out[0];
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], t, a[(i + 1) & 7]);
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[threadIdx.x + blockIdx.x * blockDim.x + 1] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_packed(float* out, int iters, long long* cyc) {
  float2 a[4]; for (int i = 0; i < 4; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  float2 t = make_float2(out[0], out[0]);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = fma2(a[i], t, a[(i + 1) & 3]);
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 4; ++i) s += a[i].x + a[i].y;
  out[threadIdx.x + blockIdx.x * blockDim.x + 1] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMemset(out, 0, 1 << 22); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    long long c1, c2;
    k_scalar<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c1, cyc, 8, cudaMemcpyDeviceToHost);
    k_packed<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c2, cyc, 8, cudaMemcpyDeviceToHost);
    printf("PROBE %2d warps/SM: scalar FFMA %.2f cycles per 8 fp32 FMAs per warp, packed f32x2 %.2f (error %s)\n", warps, (double)c1 / iters, (double)c2 / iters, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
