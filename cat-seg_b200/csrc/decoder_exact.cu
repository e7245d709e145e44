// EXACT (fp32 CUDA-core) guided upsampling decoder.
//
// Reference: Aggregator.conv_decoder (model.py:674-681), Up.forward (:549-555), DoubleConv (:520-537),
// head (:634).  Per (image, class) slice, NHWC activations:
//   ConvT(k2,s2) 128->96 @48^2 | cat guidance(32) -> conv3x3 128->64 -> GN(4) -> ReLU -> conv3x3 64->64 -> GN -> ReLU
//   ConvT(k2,s2)  64->48 @96^2 | cat guidance(16) -> conv3x3  64->32 -> GN(2) -> ReLU -> conv3x3 32->32 -> GN -> ReLU
//   conv3x3 32->1 + bias -> logits[b][class id]
// Every conv runs on the shared fp32 implicit-GEMM engine; GroupNorm is "statistics pass + normalise
// in the next consumer's loader".  Slices are processed in chunks so the scratch stays bounded.
#include "igemm.cuh"
#include "internal.h"

namespace catseg {

// ---- dense rows with optional GroupNorm+ReLU on load (input of the second transposed conv)
struct DenseGN_A {
  static constexpr bool kMFastest = false;
  const float* A; int K; int rows_per_slice;
  const float* stats;   // [S][G][2] mean, rstd  (nullptr: raw)
  const float *gamma, *beta; int G;
  __device__ float operator()(int, int m, int k) const {
    float v = __ldg(A + (long long)m * K + k);
    if (stats) {
      int s = m / rows_per_slice, g = k >> 4;
      const float* st = stats + ((long long)s * G + g) * 2;
      v = (v - __ldg(st)) * __ldg(st + 1) * __ldg(gamma + k) + __ldg(beta + k);
      v = fmaxf(v, 0.0f);
    }
    return v;
  }
};

// ---- ConvTranspose2d(k=2,s=2) epilogue: n = (dy*2+dx)*Co + co  ->  out[s][(2y+dy)*(2W) + 2x+dx][co]
struct ConvTStore {
  float* out; const float* bias; int H, W, Co;
  __device__ void operator()(int, int m, int n, float acc) const {
    int hw = H * W;
    int s = m / hw, p = m - s * hw;
    int y = p / W, x = p - y * W;
    int q = n / Co, co = n - q * Co;
    int dy = q >> 1, dx = q & 1;
    long long o = ((long long)s * 4 * hw + (long long)(2 * y + dy) * (2 * W) + 2 * x + dx) * Co + co;
    out[o] = acc + __ldg(bias + co);
  }
};

// ---- 3x3 conv (pad 1) over NHWC: channels [0,C1) from the per-slice tensor (optionally GN+ReLU'd on
// load), channels [C1,C1+C2) from the per-image projected guidance (model.py:552-554).  k = tap*Cin + ci
struct ConvNHWC_A {
  static constexpr bool kMFastest = false;
  const float* in1; int C1;
  const float* in2; int C2;
  int slice0, Te, H, W;
  const float* stats; const float *gamma, *beta; int G;
  __device__ float operator()(int, int m, int k) const {
    int hw = H * W, Cin = C1 + C2;
    int s = m / hw, p = m - s * hw;
    int tap = k / Cin, ci = k - tap * Cin;
    int y = p / W + tap / 3 - 1, x = p % W + tap % 3 - 1;
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.0f;     // zero padding of the conv input
    if (ci < C1) {
      float v = __ldg(in1 + ((long long)s * hw + y * W + x) * C1 + ci);
      if (stats) {
        const float* st = stats + ((long long)s * G + (ci >> 4)) * 2;
        v = (v - __ldg(st)) * __ldg(st + 1) * __ldg(gamma + ci) + __ldg(beta + ci);
        v = fmaxf(v, 0.0f);
      }
      return v;
    }
    int img = (slice0 + s) / Te;
    return __ldg(in2 + ((long long)img * hw + y * W + x) * C2 + (ci - C1));
  }
};

struct PlainStore {
  float* out; int ldo;
  __device__ void operator()(int, int m, int n, float acc) const { out[(long long)m * ldo + n] = acc; }
};

// ---- GroupNorm statistics (16 channels per group, eps 1e-5, biased variance): block per (slice, group)
__global__ void gn_stats_kernel(const float* __restrict__ x, float* __restrict__ stats, int npix, int C, int G) {
  __shared__ float red[32];
  __shared__ float s_mean;
  int s = blockIdx.x / G, g = blockIdx.x % G;
  const float* base = x + (long long)s * npix * C + g * 16;
  int n4 = npix * 4;   // float4 items
  float sum = 0.0f;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 v = ld4(base + (long long)(i >> 2) * C + (i & 3) * 4);
    sum += (v.x + v.y) + (v.z + v.w);
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    v = warp_sum(v);
    if (threadIdx.x == 0) s_mean = v / (float)(npix * 16);
  }
  __syncthreads();
  float mean = s_mean, sq = 0.0f;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 v = ld4(base + (long long)(i >> 2) * C + (i & 3) * 4);
    float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
  sq = warp_sum(sq);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      stats[(long long)blockIdx.x * 2] = mean;
      stats[(long long)blockIdx.x * 2 + 1] = 1.0f / sqrtf(v / (float)(npix * 16) + 1e-5f);
    }
  }
}

// ---- materialise relu(GN(x)) (only used when a parity tap is requested)
__global__ void gn_relu_apply_kernel(const float* __restrict__ x, float* __restrict__ out,
                                     const float* __restrict__ stats, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, long long total, int npix, int C, int G) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long s = i / ((long long)npix * C);
  const float* st = stats + (s * G + (c >> 4)) * 2;
  out[i] = fmaxf((x[i] - st[0]) * st[1] * gamma[c] + beta[c], 0.0f);
}

// ---- head: conv3x3 32->1 + bias on relu(GN(x)), scattered to logits[b][class id]   (model.py:634, 679)
__global__ void head_conv_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ wt, const float* __restrict__ bias,
                                 const int32_t* __restrict__ classes, float* __restrict__ logits, int slice0,
                                 int nslice, int Te, int T, int H, int W, int C, int G) {
  extern __shared__ float sw[];      // [9][C] weights, then gamma, beta
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) sw[i] = wt[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sw[9 * C + i] = gamma[i]; sw[10 * C + i] = beta[i]; }
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int hw = H * W;
  if (i >= (long long)nslice * hw) return;
  int s = (int)(i / hw), p = (int)(i % hw);
  int y0 = p / W, x0 = p % W;
  const float* st = stats + (long long)s * G * 2;
  float acc = 0.0f;
  for (int tap = 0; tap < 9; ++tap) {
    int y = y0 + tap / 3 - 1, xx = x0 + tap % 3 - 1;
    if (y < 0 || y >= H || xx < 0 || xx >= W) continue;
    const float* px = x + ((long long)s * hw + y * W + xx) * C;
    for (int c = 0; c < C; c += 4) {
      float4 v = ld4(px + c);
      float mean = st[(c >> 4) * 2], rstd = st[(c >> 4) * 2 + 1];
      float v4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float t = fmaxf((v4[q] - mean) * rstd * sw[9 * C + c + q] + sw[10 * C + c + q], 0.0f);
        acc = fmaf(t, sw[tap * C + c + q], acc);
      }
    }
  }
  int gs = slice0 + s;
  int b = gs / Te;
  int cls = classes[gs];
  logits[((long long)b * T + cls) * hw + p] = acc + bias[0];
}

size_t decoder_exact_scratch_floats(const DecoderDims& d, int chunk) {
  size_t hw = (size_t)d.H * d.W;
  size_t f = 0;
  f += (size_t)chunk * 4 * hw * d.U1;        // u1
  f += 2 * (size_t)chunk * 4 * hw * d.D1;    // c1a, c1b
  f += (size_t)chunk * 16 * hw * d.U2;       // u2
  f += 2 * (size_t)chunk * 16 * hw * d.D2;   // c2a, c2b
  f += 4 * (size_t)chunk * 8 * 2;            // GN stats (<= 8 groups)
  return f;
}

#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; ++nl; } while (0)

cudaError_t run_decoder_exact(const float* X, const float* dg0, const float* dg1, const int32_t* classes,
                              float* logits, int B, int T, int Te, const DecoderDims& d, const DecoderW& w,
                              float* scratch, int chunk, float* tap_up1, float* tap_up2, int* launches,
                              cudaStream_t st) {
  const int hw = d.H * d.W, hw1 = 4 * hw, hw2 = 16 * hw;
  const int H1 = 2 * d.H, W1 = 2 * d.W, H2 = 4 * d.H, W2 = 4 * d.W;
  const int G1 = d.D1 / 16, G2 = d.D2 / 16;
  float* u1 = scratch;
  float* c1a = u1 + (size_t)chunk * hw1 * d.U1;
  float* c1b = c1a + (size_t)chunk * hw1 * d.D1;
  float* u2 = c1b + (size_t)chunk * hw1 * d.D1;
  float* c2a = u2 + (size_t)chunk * hw2 * d.U2;
  float* c2b = c2a + (size_t)chunk * hw2 * d.D2;
  float* s1a = c2b + (size_t)chunk * hw2 * d.D2;
  float* s1b = s1a + (size_t)chunk * 16;
  float* s2a = s1b + (size_t)chunk * 16;
  float* s2b = s2a + (size_t)chunk * 16;
  int nl = 0;
  const int nslice = B * Te;
  for (int s0 = 0; s0 < nslice; s0 += chunk) {
    const int n = nslice - s0 < chunk ? nslice - s0 : chunk;
    // decoder1.up
    {
      DenseGN_A a{X + (long long)s0 * hw * d.C0, d.C0, hw, nullptr, nullptr, nullptr, 0};
      ConvTStore e{u1, w.up1_b, d.H, d.W, d.U1};
      CK(launch_igemm(a, w.up1_wt, 0, 1, n * hw, 4 * d.U1, d.C0, e, st));
    }
    // decoder1.conv: conv -> GN -> ReLU -> conv -> GN -> ReLU
    {
      ConvNHWC_A a{u1, d.U1, dg0, d.G1, s0, Te, H1, W1, nullptr, nullptr, nullptr, 0};
      CK(launch_igemm(a, w.c1a_wt, 0, 1, n * hw1, d.D1, 9 * (d.U1 + d.G1), PlainStore{c1a, d.D1}, st));
      gn_stats_kernel<<<n * G1, 256, 0, st>>>(c1a, s1a, hw1, d.D1, G1);
      CK(cudaGetLastError());
      ConvNHWC_A a2{c1a, d.D1, nullptr, 0, s0, Te, H1, W1, s1a, w.gn1a_g, w.gn1a_b, G1};
      CK(launch_igemm(a2, w.c1b_wt, 0, 1, n * hw1, d.D1, 9 * d.D1, PlainStore{c1b, d.D1}, st));
      gn_stats_kernel<<<n * G1, 256, 0, st>>>(c1b, s1b, hw1, d.D1, G1);
      CK(cudaGetLastError());
      if (tap_up1) {
        long long total = (long long)n * hw1 * d.D1;
        gn_relu_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
            c1b, tap_up1 + (long long)s0 * hw1 * d.D1, s1b, w.gn1b_g, w.gn1b_b, total, hw1, d.D1, G1);
        CK(cudaGetLastError());
      }
    }
    // decoder2.up on relu(GN(c1b))
    {
      DenseGN_A a{c1b, d.D1, hw1, s1b, w.gn1b_g, w.gn1b_b, G1};
      ConvTStore e{u2, w.up2_b, H1, W1, d.U2};
      CK(launch_igemm(a, w.up2_wt, 0, 1, n * hw1, 4 * d.U2, d.D1, e, st));
    }
    {
      ConvNHWC_A a{u2, d.U2, dg1, d.G2, s0, Te, H2, W2, nullptr, nullptr, nullptr, 0};
      CK(launch_igemm(a, w.c2a_wt, 0, 1, n * hw2, d.D2, 9 * (d.U2 + d.G2), PlainStore{c2a, d.D2}, st));
      gn_stats_kernel<<<n * G2, 256, 0, st>>>(c2a, s2a, hw2, d.D2, G2);
      CK(cudaGetLastError());
      ConvNHWC_A a2{c2a, d.D2, nullptr, 0, s0, Te, H2, W2, s2a, w.gn2a_g, w.gn2a_b, G2};
      CK(launch_igemm(a2, w.c2b_wt, 0, 1, n * hw2, d.D2, 9 * d.D2, PlainStore{c2b, d.D2}, st));
      gn_stats_kernel<<<n * G2, 256, 0, st>>>(c2b, s2b, hw2, d.D2, G2);
      CK(cudaGetLastError());
      if (tap_up2) {
        long long total = (long long)n * hw2 * d.D2;
        gn_relu_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
            c2b, tap_up2 + (long long)s0 * hw2 * d.D2, s2b, w.gn2b_g, w.gn2b_b, total, hw2, d.D2, G2);
        CK(cudaGetLastError());
      }
    }
    {
      long long total = (long long)n * hw2;
      size_t sm = (size_t)11 * d.D2 * 4;
      head_conv_kernel<<<(unsigned)((total + 255) / 256), 256, sm, st>>>(c2b, s2b, w.gn2b_g, w.gn2b_b, w.head_w,
                                                                         w.head_b, classes, logits, s0, n, Te, T,
                                                                         H2, W2, d.D2, G2);
      CK(cudaGetLastError());
    }
  }
  if (launches) *launches += nl;
  return cudaSuccess;
}

}  // namespace catseg
