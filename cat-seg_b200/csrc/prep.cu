// Per-forward preparation kernels (EXACT fp32 path): L2 normalisation, cost volume, top-k class
// selection, text/appearance guidance projections, 7x7 cost embedding.
// Reference: cat_seg/modeling/transformer/model.py:648-659, 694-715.
#include "igemm.cuh"
#include "internal.h"

namespace catseg {

// ---- F.normalize(img, dim=1) on [B,C,HW] (model.py:649): x / max(||x||_2, 1e-12) per pixel
__global__ void normalize_img_kernel(const float* __restrict__ img, float* __restrict__ out, int C, int HW) {
  int b = blockIdx.y;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float* src = img + (long long)b * C * HW + p;
  float ss = 0.0f;
  for (int c = 0; c < C; ++c) { float v = __ldg(src + (long long)c * HW); ss = fmaf(v, v, ss); }
  float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  float* dst = out + (long long)b * C * HW + p;
  for (int c = 0; c < C; ++c) dst[(long long)c * HW] = __ldg(src + (long long)c * HW) * inv;
}
cudaError_t launch_normalize_img(const float* img, float* out, int B, int C, int HW, cudaStream_t st) {
  dim3 grid((HW + 127) / 128, B);
  normalize_img_kernel<<<grid, 128, 0, st>>>(img, out, C, HW);
  return cudaGetLastError();
}

// ---- F.normalize(text, dim=-1) on rows of length C (model.py:650): one warp per row
__global__ void normalize_rows_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int C) {
  long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* src = in + r * C;
  float ss = 0.0f;
  for (int c = lane; c < C; c += 32) { float v = __ldg(src + c); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  for (int c = lane; c < C; c += 32) out[r * C + c] = __ldg(src + c) * inv;
}
cudaError_t launch_normalize_rows(const float* in, float* out, long long rows, int C, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  normalize_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(in, out, rows, C);
  return cudaGetLastError();
}

// ---- 1 / max(||row||_2, 1e-12) of rows of length C (the scale F.normalize applies, model.py:650): one warp per row
__global__ void inv_norm_rows_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int C) {
  long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* src = in + r * C;
  float ss = 0.0f;
  for (int c = lane; c < C; c += 32) { float v = __ldg(src + c); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  if (lane == 0) out[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
}
cudaError_t launch_inv_norm_rows(const float* in, float* out, long long rows, int C, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  inv_norm_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(in, out, rows, C);
  return cudaGetLastError();
}
// ---- the same for the pixels of img [B][C][HW] (model.py:649): out[b][p]
__global__ void inv_norm_pixels_kernel(const float* __restrict__ img, float* __restrict__ out, int C, int HW) {
  int b = blockIdx.y;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float* src = img + (long long)b * C * HW + p;
  float ss = 0.0f;
  for (int c = 0; c < C; ++c) { float v = __ldg(src + (long long)c * HW); ss = fmaf(v, v, ss); }
  out[(long long)b * HW + p] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
}
cudaError_t launch_inv_norm_pixels(const float* img, float* out, int B, int C, int HW, cudaStream_t st) {
  dim3 grid((HW + 127) / 128, B);
  inv_norm_pixels_kernel<<<grid, 128, 0, st>>>(img, out, C, HW);
  return cudaGetLastError();
}
// ---- dst[i][:] = src[idx[i]][:] (rows of `width` floats, width % 4 == 0): the per-image kept classes pick their rows of
// a per-vocabulary table
__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, float* __restrict__ dst,
                                   long long n, int width4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width4) return;
  long long r = i / width4;
  int c = (int)(i % width4);
  st4(dst + (r * width4 + c) * 4, ld4(src + ((long long)idx[r] * width4 + c) * 4));
}
cudaError_t launch_gather_rows(const float* src, const int32_t* idx, float* dst, long long n, int width, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  long long tot = n * (width / 4);
  gather_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, idx, dst, n, width / 4);
  return cudaGetLastError();
}

// ---- einsum('bchw,btpc->bpthw') (model.py:651) as a batched GEMM: corr[b][t*P+p][hw]
cudaError_t launch_cost_volume(const float* textn, long long text_batch_stride, const float* imgn, float* corr, int B, int TP,
                               int C, int HW, cudaStream_t st) {
  DenseA a{textn, text_batch_stride, C};
  BiasActStore e{corr, (long long)TP * HW, HW, nullptr, 0};
  return launch_igemm(a, imgn, (long long)C * HW, B, TP, HW, C, e, st);
}

// ---- per-(b,t) max over (P,HW) (model.py:695): one warp per row of length n
__global__ void class_max_kernel(const float* __restrict__ corr, float* __restrict__ cmax, long long rows, int n) {
  long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float m = -INFINITY;
  for (int i = lane; i < n; i += 32) m = fmaxf(m, __ldg(corr + r * n + i));
  m = warp_max(m);
  if (lane == 0) cmax[r] = m;
}
cudaError_t launch_class_max(const float* corr, float* cmax, long long rows, int n, cudaStream_t st) {
  class_max_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(corr, cmax, rows, n);
  return cudaGetLastError();
}

// ---- top-Te class selection (model.py:696).  topk(sorted=False) leaves the order of the kept set
// unspecified and the result is invariant to it: we keep ascending class id; ties
// at the cut go to the lower id.  One block of 1024 threads per image: rank by counting (the scores sit in shared memory
// and are compared four at a time), then an ordered compaction by warp ballots and a scan of the warp counts.
__global__ void __launch_bounds__(1024) select_classes_kernel(const float* __restrict__ cmax, int32_t* __restrict__ classes, int T,
                                                              int Te) {
  extern __shared__ __align__(16) float sc[];  // [T rounded up to 4], padded with -inf (never counted: -inf > v is false)
  __shared__ int s_wcnt[32], s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T4 = (T + 3) & ~3;
  for (int t = tid; t < T4; t += 1024) sc[t] = t < T ? cmax[(long long)b * T + t] : -INFINITY;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < T; t0 += 1024) {       // classes in ascending order, 1024 per pass
    const int t = t0 + tid;
    bool keep = false;
    if (t < T) {
      const float v = sc[t];
      int rank = 0;
      for (int u = 0; u < T4; u += 4) {
        const float4 w = *reinterpret_cast<const float4*>(sc + u);
        rank += (w.x > v) || (w.x == v && u < t);
        rank += (w.y > v) || (w.y == v && u + 1 < t);
        rank += (w.z > v) || (w.z == v && u + 2 < t);
        rank += (w.w > v) || (w.w == v && u + 3 < t);
      }
      keep = rank < Te;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int pos = s_base;
    for (int w = 0; w < warp; ++w) pos += s_wcnt[w];
    pos += __popc(bal & ((1u << lane) - 1u));
    if (keep && pos < Te) classes[(long long)b * Te + pos] = t;
    __syncthreads();
    if (tid == 0) { int n = 0; for (int w = 0; w < 32; ++w) n += s_wcnt[w]; s_base += n; }
    __syncthreads();
  }
}
cudaError_t launch_select_classes(const float* cmax, int32_t* classes, int B, int T, int Te, cudaStream_t st) {
  select_classes_kernel<<<B, 1024, (size_t)((T + 3) & ~3) * 4, st>>>(cmax, classes, T, Te);
  return cudaGetLastError();
}
__global__ void iota_classes_kernel(int32_t* classes, int n, int Te) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) classes[i] = i % Te;
}
cudaError_t launch_iota_classes(int32_t* classes, int B, int Te, cudaStream_t st) {
  int n = B * Te;
  iota_classes_kernel<<<(n + 255) / 256, 256, 0, st>>>(classes, n, Te);
  return cudaGetLastError();
}

__global__ void iota_range_kernel(int32_t* ids, int n, int Te, int offset) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = offset + i % Te;
}
cudaError_t launch_iota_range(int32_t* ids, int B, int Te, int offset, cudaStream_t st) {
  int n = B * Te;
  iota_range_kernel<<<(n + 255) / 256, 256, 0, st>>>(ids, n, Te, offset);
  return cudaGetLastError();
}

__global__ void slice_classes_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ dst, int B, int Te, int first, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * n) dst[i] = src[(i / n) * Te + first + i % n];
}
cudaError_t launch_slice_classes(const int32_t* src, int32_t* dst, int B, int Te, int first, int n, cudaStream_t st) {
  slice_classes_kernel<<<(B * n + 255) / 256, 256, 0, st>>>(src, dst, B, Te, first, n);
  return cudaGetLastError();
}

// ---- text guidance input (model.py:713-714): mean over P of the kept classes' text rows, then
// divide by the L2 norm (no eps).  src is the raw text when not truncated and the normalised text
// when truncated (model.py:701).  One block per (b, i).
__global__ void text_mean_kernel(const float* __restrict__ src, const int32_t* __restrict__ classes,
                                 float* __restrict__ out, int T, int Te, int P, int C) {
  __shared__ float red[32];
  int bi = blockIdx.x;
  int b = bi / Te;
  int t = classes[bi];
  const float* row = src + ((long long)b * T + t) * P * C;
  float ss = 0.0f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.0f;
    for (int p = 0; p < P; ++p) s += __ldg(row + (long long)p * C + c);
    s = s / (float)P;
    out[(long long)bi * C + c] = s;
    ss = fmaf(s, s, ss);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  float nrm = sqrtf(red[0]);
  for (int c = threadIdx.x; c < C; c += blockDim.x) out[(long long)bi * C + c] = out[(long long)bi * C + c] / nrm;
}
cudaError_t launch_text_mean(const float* src, const int32_t* classes, float* out, int B, int T, int Te, int P,
                             int C, cudaStream_t st) {
  text_mean_kernel<<<B * Te, 128, 0, st>>>(src, classes, out, T, Te, P, C);
  return cudaGetLastError();
}

// ---- LayerNorm over rows of 128 (guidance_norm, model.py:233,249): one warp per row
__global__ void layernorm128_kernel(const float* __restrict__ in, float* __restrict__ out,
                                    const float* __restrict__ g, const float* __restrict__ b, long long rows) {
  long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float4 x = ld4(in + r * 128 + lane * 4);
  float4 y = warp_layernorm128(x, ld4(g + lane * 4), ld4(b + lane * 4));
  st4(out + r * 128 + lane * 4, y);
}
cudaError_t launch_layernorm128(const float* in, float* out, const float* g, const float* b, long long rows,
                                cudaStream_t st) {
  layernorm128_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(in, out, g, b, rows);
  return cudaGetLastError();
}

// ---- out[M][N] = act(A[M][K] @ Wt[K][N] + bias)
cudaError_t launch_linear(const float* A, const float* Wt, const float* bias, float* out, long long M, int N,
                          int K, int relu, cudaStream_t st) {
  DenseA a{A, 0, K};
  BiasActStore e{out, 0, N, bias, relu};
  return launch_igemm(a, Wt, 0, 1, (int)M, N, K, e, st);
}

// ---- 3x3 conv (pad 1) on NCHW input, NHWC output, + bias + ReLU.  k = ci*9 + ky*3 + kx.
struct ConvNCHW_A {
  static constexpr bool kMFastest = true;
  const float* in; int Ci, H, W;
  __device__ float operator()(int b, int m, int k) const {
    int ci = k / 9, tap = k - ci * 9;
    int y = m / W + tap / 3 - 1, x = m % W + tap % 3 - 1;
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.0f;
    return __ldg(in + (((long long)b * Ci + ci) * H + y) * W + x);
  }
};
cudaError_t launch_conv3x3_nchw(const float* in, const float* Wt, const float* bias, float* out, int B, int Ci,
                                int H, int W, int Co, cudaStream_t st) {
  ConvNCHW_A a{in, Ci, H, W};
  BiasActStore e{out, (long long)H * W * Co, Co, bias, 1};
  return launch_igemm(a, Wt, 0, B, H * W, Co, Ci * 9, e, st);
}

// ---- 7x7 cost embedding: batch = (b,i) slice, A(m=pixel, k=p*49+ky*7+kx) gathered from corr
struct EmbedA {
  static constexpr bool kMFastest = true;
  const float* corr; const int32_t* classes; int T, Te, P, H, W;
  __device__ float operator()(int s, int m, int k) const {
    int p = k / 49, tap = k - p * 49;
    int y = m / W + tap / 7 - 3, x = m % W + tap % 7 - 3;
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.0f;
    int b = s / Te;
    int t = __ldg(classes + s);
    return __ldg(corr + (((long long)b * T + t) * P + p) * H * W + y * W + x);
  }
};
cudaError_t launch_cost_embed(const float* corr, const int32_t* classes, const float* Wt, const float* bias,
                              float* X, int B, int T, int Te, int P, int H, int W, cudaStream_t st) {
  EmbedA a{corr, classes, T, Te, P, H, W};
  BiasActStore e{X, (long long)H * W * CATSEG_HID, CATSEG_HID, bias, 0};
  int nslice = B * Te;
  if (nslice > 65535) return cudaErrorInvalidValue;  // gridDim.z limit; cfg4 has 4096 slices
  return launch_igemm(a, Wt, 0, nslice, H * W, CATSEG_HID, P * 49, e, st);
}

// ---- dst[k][dst_col0 + r] = src[r*lds + src_col0 + k]  (Linear weight [out][in] -> [in][out])
__global__ void transpose_pack_kernel(float* __restrict__ dst, int ldd, int dst_col0, const float* __restrict__ src,
                                      int lds, int src_col0, int rows, int cols) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  int r = i % rows, k = i / rows;
  dst[(long long)k * ldd + dst_col0 + r] = src[(long long)r * lds + src_col0 + k];
}
cudaError_t launch_transpose_pack(float* dst, int ldd, int dst_col0, const float* src, int lds, int src_col0,
                                  int rows, int cols, cudaStream_t st) {
  int n = rows * cols;
  transpose_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(dst, ldd, dst_col0, src, lds, src_col0, rows, cols);
  return cudaGetLastError();
}

__global__ void fill_kernel(float* p, float v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}
cudaError_t launch_fill(float* p, float v, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, v, n);
  return cudaGetLastError();
}

}  // namespace catseg
