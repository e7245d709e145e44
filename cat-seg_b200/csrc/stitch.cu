// Sliding-window stitch + postprocess + argmax, fused (no [T,384,384] / [T,640,640] intermediates).
//
// Reference: CATSeg.forward sliding-window branch, cat_seg/cat_seg_model.py:204-218:
//   outputs = interpolate(logits[nwin+1,T,S,S], size=kernel, bilinear, align_corners=False).sigmoid()
//   global  = interpolate(outputs[-1:], size=out_res)                                  (:209-210)
//   tiles   = fold(outputs[:-1].flatten(1).T) / fold(unfold(ones))                     (:212)
//   out     = (tiles + global) / 2                                                     (:213)
//   sem_seg = sem_seg_postprocess(out[0], out_res, height, width)   -> bilinear to (height,width)
//   label   = sem_seg.argmax(dim=0)                                 (train_net.py:58)
// Index logic follows PyTorch exactly: bilinear source index = max(scale*(dst+0.5)-0.5, 0) with
// scale = in/out in fp32; Fold accumulates covering tiles in (ky,kx)-ascending order, i.e. tiles in
// descending (ty,tx); the count map is the integer number of covering tiles; argmax keeps the first
// maximum.
#include "common.cuh"
#include "internal.h"

namespace catseg {

namespace {

struct Lerp { int i0, i1; float l0, l1; };

__device__ __forceinline__ Lerp make_lerp(int dst, float scale, int in_size) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.0f) src = 0.0f;
  Lerp r;
  r.i0 = (int)src;
  if (r.i0 > in_size - 1) r.i0 = in_size - 1;
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.0f - r.l1;
  return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// sigmoid(bilinear S->kernel) of one window plane at (ky,kx) on the kernel grid
__device__ __forceinline__ float win_prob(const float* __restrict__ plane, int S, float sc, int ky, int kx) {
  Lerp a = make_lerp(ky, sc, S), b = make_lerp(kx, sc, S);
  float v00 = __ldg(plane + a.i0 * S + b.i0), v01 = __ldg(plane + a.i0 * S + b.i1);
  float v10 = __ldg(plane + a.i1 * S + b.i0), v11 = __ldg(plane + a.i1 * S + b.i1);
  float v = a.l0 * (b.l0 * v00 + b.l1 * v01) + a.l1 * (b.l0 * v10 + b.l1 * v11);
  return sigmoidf_(v);
}

// stitched probability of class plane set `cls` at (Y,X) on the out_res grid
__device__ float stitched(const float* __restrict__ win_logits, int T, int t, int S, int kernel, int stride,
                          int out_res, int ntile, int Y, int X) {
  const long long plane = (long long)S * S;
  const float sc_win = (float)S / (float)kernel;
  float acc = 0.0f;
  int cnt = 0;
  for (int ty = ntile - 1; ty >= 0; --ty) {
    int ky = Y - ty * stride;
    if (ky < 0 || ky >= kernel) continue;
    for (int tx = ntile - 1; tx >= 0; --tx) {
      int kx = X - tx * stride;
      if (kx < 0 || kx >= kernel) continue;
      const float* p = win_logits + ((long long)(ty * ntile + tx) * T + t) * plane;
      acc += win_prob(p, S, sc_win, ky, kx);
      ++cnt;
    }
  }
  acc = acc / (float)cnt;
  // global view: bilinear kernel -> out_res of sigmoid(bilinear S -> kernel)
  const float* gp = win_logits + ((long long)(ntile * ntile) * T + t) * plane;
  const float sc_g = (float)kernel / (float)out_res;
  Lerp a = make_lerp(Y, sc_g, kernel), b = make_lerp(X, sc_g, kernel);
  float g00 = win_prob(gp, S, sc_win, a.i0, b.i0), g01 = win_prob(gp, S, sc_win, a.i0, b.i1);
  float g10 = win_prob(gp, S, sc_win, a.i1, b.i0), g11 = win_prob(gp, S, sc_win, a.i1, b.i1);
  float g = a.l0 * (b.l0 * g00 + b.l1 * g01) + a.l1 * (b.l0 * g10 + b.l1 * g11);
  return (acc + g) / 2.0f;
}

}  // namespace

__global__ void stitch_kernel(const float* __restrict__ win_logits, int T, int S, int kernel, int stride,
                              int out_res, int ntile, int height, int width, float* __restrict__ probs_out,
                              int32_t* __restrict__ labels_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)height * width) return;
  int oy = (int)(i / width), ox = (int)(i % width);
  const bool identity = (height == out_res && width == out_res);
  Lerp a, b;
  if (!identity) {
    a = make_lerp(oy, (float)out_res / (float)height, out_res);
    b = make_lerp(ox, (float)out_res / (float)width, out_res);
  }
  float best = -INFINITY;
  int best_t = 0;
  for (int t = 0; t < T; ++t) {
    float v;
    if (identity) {
      v = stitched(win_logits, T, t, S, kernel, stride, out_res, ntile, oy, ox);
    } else {
      float v00 = stitched(win_logits, T, t, S, kernel, stride, out_res, ntile, a.i0, b.i0);
      float v01 = stitched(win_logits, T, t, S, kernel, stride, out_res, ntile, a.i0, b.i1);
      float v10 = stitched(win_logits, T, t, S, kernel, stride, out_res, ntile, a.i1, b.i0);
      float v11 = stitched(win_logits, T, t, S, kernel, stride, out_res, ntile, a.i1, b.i1);
      v = a.l0 * (b.l0 * v00 + b.l1 * v01) + a.l1 * (b.l0 * v10 + b.l1 * v11);
    }
    if (probs_out) probs_out[(long long)t * height * width + i] = v;
    if (v > best) { best = v; best_t = t; }
  }
  if (labels_out) labels_out[i] = best_t;
}

cudaError_t launch_stitch(const float* win_logits, int T, int S, int kernel, int stride, int out_res, int height,
                          int width, float* probs_out, int32_t* labels_out, cudaStream_t st) {
  if (kernel > out_res || stride <= 0 || (out_res - kernel) % stride != 0) return cudaErrorInvalidValue;
  int ntile = (out_res - kernel) / stride + 1;
  long long n = (long long)height * width;
  stitch_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(win_logits, T, S, kernel, stride, out_res, ntile,
                                                             height, width, probs_out, labels_out);
  return cudaGetLastError();
}

// labels[i] = argmax_t scores[t][i], first maximum wins (torch.argmax semantics on distinct values;
// NaN is not expected on this path).  One CTA owns 32 pixels; its 8 warps split the class axis (t = w, w+8, ...),
// 8 independent 128-byte loads in flight per warp, then a fixed-order merge (lower class index wins ties).
__global__ void __launch_bounds__(256) argmax_kernel(const float* __restrict__ scores, int T, long long npix,
                                                     int32_t* __restrict__ labels) {
  __shared__ float s_best[8][32];
  __shared__ int s_bt[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + lane;
  float best = -INFINITY;
  int bt = 0x7fffffff;
  if (i < npix) {
    int t = w;
    for (; t + 56 < T; t += 64) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(scores + (long long)(t + 8 * u) * npix + i);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (v[u] > best || bt == 0x7fffffff) { best = v[u]; bt = t + 8 * u; }
    }
    for (; t < T; t += 8) {
      float v = __ldg(scores + (long long)t * npix + i);
      if (v > best || bt == 0x7fffffff) { best = v; bt = t; }
    }
  }
  s_best[w][lane] = best;
  s_bt[w][lane] = bt;
  __syncthreads();
  if (w == 0 && i < npix) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      float v = s_best[k][lane];
      int t = s_bt[k][lane];
      if (t != 0x7fffffff && (bt == 0x7fffffff || v > best || (v == best && t < bt))) { best = v; bt = t; }
    }
    labels[i] = bt == 0x7fffffff ? 0 : bt;
  }
}
cudaError_t launch_argmax(const float* scores, int T, long long npix, int32_t* labels, cudaStream_t st) {
  argmax_kernel<<<(unsigned)((npix + 31) / 32), 256, 0, st>>>(scores, T, npix, labels);
  return cudaGetLastError();
}

}  // namespace catseg
