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

// Tiled version (the one launched): a CTA owns 16x16 output pixels and walks the classes.  Per class it first
// evaluates the global-view probabilities on the patch of the `kernel` grid its pixels can touch (phase A), then the
// stitched value on the patch of the out_res grid they can touch (phase B: covering tiles + bilinear of the phase-A
// patch), then each thread interpolates its own output pixel from that patch (phase C).  Every value is computed with
// the same fp32 expression order as stitch_kernel above; what changes is that the four global-view samples per
// stitched value (and, when postprocess resizes, the four stitched values per output pixel) are shared through shared
// memory instead of being recomputed by every thread: ~2 instead of 8 sigmoid(bilinear) evaluations per pixel and class.
constexpr int ST_TILE = 16, ST_RMAX = 36, ST_GMAX = 26, ST_MAXT = 4, ST_G = 8;

// sigmoid on the MUFU pipe: ex2.approx + rcp.approx.  |error| <= ~1e-7 on a probability (relative 3e-7 on exp for
// |x| <= 10, damped by p(1-p) <= 1/4); the parity tolerance on stitched probabilities is 2e-6 (tests/test_gpu_parity.py).
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

struct LerpE { int o0, o1; float l0, l1; };      // like Lerp, with the row indices pre-multiplied where useful

// Which (window, class) planes are dropped classes?  The Aggregator writes exactly -100.0 into every pixel of a class
// that the top-256 truncation dropped (model.py:721-724), bilinear interpolation of a constant plane stays within
// rounding of -100, and sigmoid of that is exactly 0.0 in fp32 (exp(100) overflows): a plane whose S*S values ALL equal
// -100.0 contributes exactly nothing.  With T = 847 about 70 % of the planes are such planes.  One CTA per class scans its
// nwin planes (the whole logits tensor is read once: 156 MB for five windows of A-847) and writes a bit mask.
__global__ void __launch_bounds__(256) plane_live_kernel(const float* __restrict__ win_logits, int T, int nwin, int plane4,
                                                         uint32_t* __restrict__ mask) {
  const int t = blockIdx.x;
  __shared__ uint32_t s_m;
  if (threadIdx.x == 0) s_m = 0u;
  __syncthreads();
  uint32_t m = 0u;
  for (int w = 0; w < nwin; ++w) {
    const float4* p = reinterpret_cast<const float4*>(win_logits) + ((long long)w * T + t) * plane4;
    bool live = false;
    for (int i = threadIdx.x; i < plane4; i += 256) {
      const float4 v = __ldg(p + i);
      live |= (v.x != -100.0f) | (v.y != -100.0f) | (v.z != -100.0f) | (v.w != -100.0f);
    }
    if (__any_sync(0xffffffffu, live)) m |= 1u << w;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicOr(&s_m, m);
  __syncthreads();
  if (threadIdx.x == 0) mask[t] = s_m;
}

// `mask` (optional): bit w of mask[t] = plane (w, t) is live (plane_live_kernel); nullptr = every plane is evaluated.
__global__ void __launch_bounds__(ST_TILE * ST_TILE, 3)
stitch_tiled_kernel(const float* __restrict__ win_logits, int T, int S, int kernel, int stride, int out_res, int ntile,
                    int height, int width, float* __restrict__ probs_out, int32_t* __restrict__ labels_out,
                    const uint32_t* __restrict__ mask) {
  extern __shared__ __align__(16) float st_dyn[];                     // s_g | s_v | live class ids [T] (uint16)
  float* s_g = st_dyn;
  float* s_v = st_dyn + ST_G * ST_GMAX * ST_GMAX;
  uint16_t* s_live = reinterpret_cast<uint16_t*>(s_v + ST_G * ST_RMAX * ST_RMAX);
  __shared__ int s_nlive;
  // class-independent interpolation tables (built once per CTA)
  __shared__ LerpE s_gy[ST_GMAX], s_gx[ST_GMAX];                     // kernel-grid patch row/col -> source rows (x S) / cols
  __shared__ LerpE s_ty[ST_RMAX][ST_MAXT], s_tx[ST_RMAX][ST_MAXT];   // out_res patch row/col x tile -> source rows (x S) / cols; o0 < 0: not covered
  __shared__ LerpE s_vy[ST_RMAX], s_vx[ST_RMAX];                     // out_res patch row/col -> rows (x GX) / cols of the kernel-grid patch
  const int tid = threadIdx.x, lx = tid % ST_TILE, ly = tid / ST_TILE;
  const int oy0 = blockIdx.y * ST_TILE, ox0 = blockIdx.x * ST_TILE;
  const int oy1 = min(oy0 + ST_TILE, height) - 1, ox1 = min(ox0 + ST_TILE, width) - 1;
  const int oy = oy0 + ly, ox = ox0 + lx;
  const bool live = oy < height && ox < width;
  const bool identity = (height == out_res && width == out_res);
  const float sc_o_y = (float)out_res / (float)height, sc_o_x = (float)out_res / (float)width;
  // patch of the out_res grid touched by this CTA (bilinear source indices are monotone in the destination index)
  int Ya, Yb, Xa, Xb;
  if (identity) { Ya = oy0; Yb = oy1; Xa = ox0; Xb = ox1; }
  else {
    Ya = make_lerp(oy0, sc_o_y, out_res).i0; Yb = make_lerp(oy1, sc_o_y, out_res).i1;
    Xa = make_lerp(ox0, sc_o_x, out_res).i0; Xb = make_lerp(ox1, sc_o_x, out_res).i1;
  }
  const int RY = Yb - Ya + 1, RX = Xb - Xa + 1;
  // patch of the kernel grid touched by the global view of that patch
  const float sc_g = (float)kernel / (float)out_res;
  const int Ga = make_lerp(Ya, sc_g, kernel).i0, Gb = make_lerp(Yb, sc_g, kernel).i1;
  const int Ha = make_lerp(Xa, sc_g, kernel).i0, Hb = make_lerp(Xb, sc_g, kernel).i1;
  const int GY = Gb - Ga + 1, GX = Hb - Ha + 1;
  const long long plane = (long long)S * S;
  const float sc_win = (float)S / (float)kernel;
  for (int i = tid; i < GY; i += ST_TILE * ST_TILE) { Lerp a = make_lerp(Ga + i, sc_win, S); s_gy[i] = LerpE{a.i0 * S, a.i1 * S, a.l0, a.l1}; }
  for (int i = tid; i < GX; i += ST_TILE * ST_TILE) { Lerp a = make_lerp(Ha + i, sc_win, S); s_gx[i] = LerpE{a.i0, a.i1, a.l0, a.l1}; }
  for (int i = tid; i < RY * ntile; i += ST_TILE * ST_TILE) {
    const int r = i / ntile, ty = i % ntile, ky = Ya + r - ty * stride;
    LerpE e{-1, -1, 0.f, 0.f};
    if (ky >= 0 && ky < kernel) { Lerp a = make_lerp(ky, sc_win, S); e = LerpE{a.i0 * S, a.i1 * S, a.l0, a.l1}; }
    s_ty[r][ty] = e;
  }
  for (int i = tid; i < RX * ntile; i += ST_TILE * ST_TILE) {
    const int r = i / ntile, tx = i % ntile, kx = Xa + r - tx * stride;
    LerpE e{-1, -1, 0.f, 0.f};
    if (kx >= 0 && kx < kernel) { Lerp a = make_lerp(kx, sc_win, S); e = LerpE{a.i0, a.i1, a.l0, a.l1}; }
    s_tx[r][tx] = e;
  }
  for (int i = tid; i < RY; i += ST_TILE * ST_TILE) { Lerp a = make_lerp(Ya + i, sc_g, kernel); s_vy[i] = LerpE{(a.i0 - Ga) * GX, (a.i1 - Ga) * GX, a.l0, a.l1}; }
  for (int i = tid; i < RX; i += ST_TILE * ST_TILE) { Lerp a = make_lerp(Xa + i, sc_g, kernel); s_vx[i] = LerpE{a.i0 - Ha, a.i1 - Ha, a.l0, a.l1}; }
  Lerp pa, pb;                                   // this thread's output pixel on the out_res grid
  if (live && !identity) { pa = make_lerp(oy, sc_o_y, out_res); pb = make_lerp(ox, sc_o_x, out_res); }
  __syncthreads();
  auto sample = [&](const float* __restrict__ pl, const LerpE& a, const LerpE& b) {
    const float v00 = __ldg(pl + a.o0 + b.o0), v01 = __ldg(pl + a.o0 + b.o1);
    const float v10 = __ldg(pl + a.o1 + b.o0), v11 = __ldg(pl + a.o1 + b.o1);
    return sigmoid_fast(a.l0 * (b.l0 * v00 + b.l1 * v01) + a.l1 * (b.l0 * v10 + b.l1 * v11));
  };
  // per-thread work lists (class independent): <= 3 phase-A samples, <= 6 phase-B patch pixels
  constexpr int NA = (ST_GMAX * ST_GMAX + ST_TILE * ST_TILE - 1) / (ST_TILE * ST_TILE);
  constexpr int NB = (ST_RMAX * ST_RMAX + ST_TILE * ST_TILE - 1) / (ST_TILE * ST_TILE);
  int ea[NA], eb[NB];
  float einv[NB];                                // 1/cnt when the tile count is a power of two (exact), else -cnt
#pragma unroll
  for (int k = 0; k < NA; ++k) { const int i = tid + k * ST_TILE * ST_TILE; ea[k] = i < GY * GX ? ((i / GX) << 8 | (i % GX)) : -1; }
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int i = tid + k * ST_TILE * ST_TILE;
    eb[k] = -1; einv[k] = 1.0f;
    if (i < RY * RX) {
      const int ry = i / RX, rx = i % RX;
      int cy = 0, cx = 0;
      for (int ty = 0; ty < ntile; ++ty) cy += s_ty[ry][ty].o0 >= 0;
      for (int tx = 0; tx < ntile; ++tx) cx += s_tx[rx][tx].o0 >= 0;
      const int cnt = cy * cx;
      eb[k] = ry << 8 | rx;
      einv[k] = (cnt & (cnt - 1)) == 0 ? 1.0f / (float)cnt : -(float)cnt;
    }
  }
  // ---- live classes of this image (mask != 0), ascending; dropped classes have stitched value exactly 0 everywhere
  int nlive = 0;
  if (mask != nullptr) {
    if (tid == 0) {
      int n = 0;
      for (int t = 0; t < T; ++t) if (__ldg(mask + t) != 0u) s_live[n++] = (uint16_t)t;
      s_nlive = n;
    }
    __syncthreads();
    nlive = s_nlive;
  } else {
    nlive = T;
  }
  float best = -INFINITY;
  int best_t = 0;
  if (mask != nullptr) {
    // first-maximum rule with the skipped classes: their value 0.0 can only win from the front of the class axis (every
    // stitched value is >= 0, and a later class needs a strictly larger one)
    if (nlive == 0 || s_live[0] != 0) { best = 0.0f; best_t = 0; }
    if (probs_out != nullptr && live) {
      int k = 0;
      for (int t = 0; t < T; ++t) {
        if (k < nlive && s_live[k] == t) { ++k; continue; }
        probs_out[(long long)t * height * width + (long long)oy * width + ox] = 0.0f;
      }
    }
  }
  // ST_G classes per round: one barrier pair per round instead of per class, and ST_G times as many independent loads in
  // flight (the loop is latency bound: every class brings new planes from L2)
  for (int l0 = 0; l0 < nlive; l0 += ST_G) {
    int tt[ST_G];
    uint32_t mm[ST_G];
#pragma unroll
    for (int c = 0; c < ST_G; ++c) {
      const int l = l0 + c;
      tt[c] = l < nlive ? (mask != nullptr ? (int)s_live[l] : l) : -1;
      mm[c] = tt[c] < 0 ? 0u : (mask != nullptr ? __ldg(mask + tt[c]) : 0xffffffffu);
    }
    // ---- phase A: global view on the kernel grid (not needed when the global window dropped the class).  The interpolation
    //      entry of a sample is class independent: it is fetched once and applied to the ST_G classes of the round
    const long long gplane0 = (long long)(ntile * ntile) * T * plane;
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      if (ea[k] < 0) continue;
      const LerpE a = s_gy[ea[k] >> 8], b = s_gx[ea[k] & 255];
#pragma unroll
      for (int c = 0; c < ST_G; ++c)
        if ((mm[c] >> (ntile * ntile)) & 1u)
          s_g[c * (ST_GMAX * ST_GMAX) + tid + k * ST_TILE * ST_TILE] = sample(win_logits + gplane0 + (long long)tt[c] * plane, a, b);
    }
    __syncthreads();
    // ---- phase B: stitched value on the out_res grid
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      if (eb[k] < 0) continue;
      const int ry = eb[k] >> 8, rx = eb[k] & 255;
      float acc[ST_G];
#pragma unroll
      for (int c = 0; c < ST_G; ++c) acc[c] = 0.0f;
      for (int ty = ntile - 1; ty >= 0; --ty) {
        const LerpE a = s_ty[ry][ty];
        if (a.o0 < 0) continue;
        for (int tx = ntile - 1; tx >= 0; --tx) {
          const LerpE b = s_tx[rx][tx];
          if (b.o0 < 0) continue;
          const int w = ty * ntile + tx;
          const float* wp = win_logits + (long long)w * T * plane;
#pragma unroll
          for (int c = 0; c < ST_G; ++c)
            if ((mm[c] >> w) & 1u) acc[c] += sample(wp + (long long)tt[c] * plane, a, b);     // a dropped plane adds exactly 0.0
        }
      }
      const LerpE a = s_vy[ry], b = s_vx[rx];
#pragma unroll
      for (int c = 0; c < ST_G; ++c) {
        if (mm[c] == 0u) continue;
        const float* sg = s_g + c * (ST_GMAX * ST_GMAX);
        const float av = einv[k] > 0.0f ? acc[c] * einv[k] : acc[c] / -einv[k];
        const float g = ((mm[c] >> (ntile * ntile)) & 1u)
                            ? a.l0 * (b.l0 * sg[a.o0 + b.o0] + b.l1 * sg[a.o0 + b.o1]) + a.l1 * (b.l0 * sg[a.o1 + b.o0] + b.l1 * sg[a.o1 + b.o1])
                            : 0.0f;
        s_v[c * (ST_RMAX * ST_RMAX) + tid + k * ST_TILE * ST_TILE] = (av + g) * 0.5f;
      }
    }
    __syncthreads();
    // ---- phase C: this thread's output pixel
    if (live) {
#pragma unroll
      for (int c = 0; c < ST_G; ++c) {
        if (mm[c] == 0u) continue;
        const float* sv = s_v + c * (ST_RMAX * ST_RMAX);
        float v;
        if (identity) {
          v = sv[(oy - Ya) * RX + (ox - Xa)];
        } else {
          const float* v0 = sv + (pa.i0 - Ya) * RX - Xa;
          const float* v1 = sv + (pa.i1 - Ya) * RX - Xa;
          v = pa.l0 * (pb.l0 * v0[pb.i0] + pb.l1 * v0[pb.i1]) + pa.l1 * (pb.l0 * v1[pb.i0] + pb.l1 * v1[pb.i1]);
        }
        if (probs_out) probs_out[(long long)tt[c] * height * width + (long long)oy * width + ox] = v;
        if (v > best) { best = v; best_t = tt[c]; }
      }
    }
    // (the next round's phase A writes s_g only, which phase B of this round has finished reading before the barrier above;
    //  its phase B writes s_v after the next barrier, by which time every thread has left this phase C)
  }
  if (live && labels_out) labels_out[(long long)oy * width + ox] = best_t;
}

cudaError_t launch_stitch(const float* win_logits, int T, int S, int kernel, int stride, int out_res, int height,
                          int width, float* probs_out, int32_t* labels_out, uint32_t* scratch_mask, cudaStream_t st) {
  if (kernel > out_res || stride <= 0 || (out_res - kernel) % stride != 0) return cudaErrorInvalidValue;
  int ntile = (out_res - kernel) / stride + 1;
  long long n = (long long)height * width;
  // patch bounds of the tiled kernel: 16 output pixels span at most 16*scale + 2 source pixels per axis
  const double so = (double)out_res / (double)(height < width ? height : width);
  const bool identity = height == out_res && width == out_res;
  const int rmax = identity ? ST_TILE : (int)(ST_TILE * so) + 3;
  const int gmax = (int)(rmax * (double)kernel / (double)out_res) + 3;
  if (rmax <= ST_RMAX && gmax <= ST_GMAX && ntile <= ST_MAXT) {
    dim3 grid((unsigned)((width + ST_TILE - 1) / ST_TILE), (unsigned)((height + ST_TILE - 1) / ST_TILE));
    const int nwin = ntile * ntile + 1;
    const bool use_mask = scratch_mask != nullptr && nwin <= 32 && T <= 6000 && (S * S) % 4 == 0 && (reinterpret_cast<uintptr_t>(win_logits) & 15) == 0;
    if (use_mask) plane_live_kernel<<<T, 256, 0, st>>>(win_logits, T, nwin, S * S / 4, scratch_mask);
    const size_t dyn = (size_t)ST_G * (ST_GMAX * ST_GMAX + ST_RMAX * ST_RMAX) * sizeof(float) + (use_mask ? (size_t)T * 2 : 0);
    cudaError_t e = cudaFuncSetAttribute(stitch_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    stitch_tiled_kernel<<<grid, ST_TILE * ST_TILE, dyn, st>>>(win_logits, T, S, kernel, stride, out_res, ntile, height, width,
                                                             probs_out, labels_out, use_mask ? scratch_mask : nullptr);
  } else {                                       // extreme down-scaling: one thread per pixel recomputes everything
    stitch_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(win_logits, T, S, kernel, stride, out_res, ntile,
                                                               height, width, probs_out, labels_out);
  }
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
  scores += (long long)blockIdx.y * T * npix;                      // batch of independent [T, npix] problems
  labels += (long long)blockIdx.y * npix;
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
cudaError_t launch_argmax(const float* scores, int batch, int T, long long npix, int32_t* labels, cudaStream_t st) {
  argmax_kernel<<<dim3((unsigned)((npix + 31) / 32), (unsigned)batch), 256, 0, st>>>(scores, T, npix, labels);
  return cudaGetLastError();
}

}  // namespace catseg
