// EXACT (fp32 CUDA-core) class aggregation: linear attention over the class axis at every pixel.
//
// Reference: ClassTransformerLayer.forward (model.py:387-424), AttentionLayer.forward (:338-354),
// LinearAttention.forward (:266-286).  The layer is split at the only point where classes couple:
//   class_state : per (image, pixel)  KV[h] = sum_t phi(k_t)^T (v_t / S),  Ksum = sum_t phi(k_t)
//   class_apply : per token  x1 = x + phi(q) KV / (phi(q).Ksum + eps) * S ; x2 = x1 + MLP(LN2 x1)
// so a class-sharded multi-GPU run only has to all-reduce the 4x(32x32+32) state per pixel
//  and so padding tokens — identical at every pixel and discarded afterwards
// (model.py:401-402, 419-421) — reduce to a constant added to the state.
// The guidance half of q/k depends only on (image, class): cg_qk[b][t][0:128]=q term, [128:256]=k
// term, biases included.
#include "common.cuh"
#include "internal.h"

namespace catseg {

namespace {
constexpr int CW = 8;            // warps per CTA
constexpr int CT = 8;            // tokens per warp per chunk
constexpr int CHUNK = CW * CT;   // 64 tokens
}

// ---------------------------------------------------------------------------------------------
// Contribution of the (pad_len - Te) identical padding tokens to the state.  One block, 128 thr.
__global__ void class_pad_state_kernel(ClassLayerW w, int Tg, float* __restrict__ pad_state, int n_pad, int S) {
  __shared__ float xn[128], kp[128], vp[128];
  int t = threadIdx.x;
  if (n_pad <= 0) {
    for (int i = t; i < kStateFloats; i += 128) pad_state[i] = 0.0f;
    return;
  }
  // LN1(padding_tokens) by warp 0
  if (t < 32) {
    float4 y = warp_layernorm128(ld4(w.pad_tok + t * 4), ld4(w.ln1_g + t * 4), ld4(w.ln1_b + t * 4));
    st4(xn + t * 4, y);
  }
  __syncthreads();
  float k = 0.0f, v = 0.0f;
  for (int i = 0; i < 128; ++i) {
    k = fmaf(xn[i], w.wqkv_t[i * 384 + 128 + t], k);
    v = fmaf(xn[i], w.wqkv_t[i * 384 + 256 + t], v);
  }
  float kg = w.bqk[128 + t];
  for (int i = 0; i < Tg; ++i) kg = fmaf(w.pad_g[i], w.wg_qk_t[i * 256 + 128 + t], kg);
  kp[t] = elu1(k + kg);
  vp[t] = (v + w.bv[t]) / (float)S;
  __syncthreads();
  float np = (float)n_pad;
  for (int i = t; i < 4 * 32 * 32; i += 128) {
    int h = i >> 10, d = (i >> 5) & 31, vv = i & 31;
    pad_state[i] = np * (kp[h * 32 + d] * vp[h * 32 + vv]);
  }
  pad_state[4096 + t] = np * kp[t];
}
cudaError_t launch_class_pad_state(const ClassLayerW& w, int Tg, float* pad_state, int n_pad, int S,
                                   cudaStream_t st) {
  class_pad_state_kernel<<<1, 128, 0, st>>>(w, Tg, pad_state, n_pad, S);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// state[b][p] = sum over the Te real classes.  One CTA per (image, pixel).
__global__ void __launch_bounds__(CW * 32)
class_state_exact_kernel(const float* __restrict__ X, const float* __restrict__ cg_qk, float* __restrict__ state,
                         int Te, int npix, int S, ClassLayerW w) {
  extern __shared__ __align__(16) float smem[];
  float* xn = smem;                       // [CW][CT][128]
  float* Ks = smem + CW * CT * 128;       // [CHUNK][128]
  float* Vs = Ks + CHUNK * 128;           // [CHUNK][128]
  const int b = blockIdx.x / npix, p = blockIdx.x % npix;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  const float invS = 1.0f / (float)S;

  // accumulation mapping: thread -> head h, key dim d, 16 value dims
  const int ah = t >> 6, ad = (t & 63) >> 1, av0 = (t & 1) * 16;
  float kv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kv[i] = 0.0f;
  float ksum = 0.0f;

  float4 g1 = ld4(w.ln1_g + lane * 4), be1 = ld4(w.ln1_b + lane * 4), bv = ld4(w.bv + lane * 4);
  float* xw = xn + warp * CT * 128;

  for (int c0 = 0; c0 < Te; c0 += CHUNK) {
    // LN1 of this warp's tokens
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      int tk = c0 + warp * CT + j;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tk < Te) x = ld4(X + (((long long)b * Te + tk) * npix + p) * 128 + lane * 4);
      st4(xw + j * 128 + lane * 4, warp_layernorm128(x, g1, be1));
    }
    __syncwarp();
    float ak[CT][4], avv[CT][4];
#pragma unroll
    for (int j = 0; j < CT; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) ak[j][i] = avv[j][i] = 0.0f;
    for (int k = 0; k < 128; k += 4) {
      float4 wk[4], wv[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        wk[kk] = ld4(w.wqkv_t + (k + kk) * 384 + 128 + lane * 4);
        wv[kk] = ld4(w.wqkv_t + (k + kk) * 384 + 256 + lane * 4);
      }
#pragma unroll
      for (int j = 0; j < CT; ++j) {
        float4 a = ld4(xw + j * 128 + k);
        float a4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          ak[j][0] = fmaf(a4[kk], wk[kk].x, ak[j][0]); ak[j][1] = fmaf(a4[kk], wk[kk].y, ak[j][1]);
          ak[j][2] = fmaf(a4[kk], wk[kk].z, ak[j][2]); ak[j][3] = fmaf(a4[kk], wk[kk].w, ak[j][3]);
          avv[j][0] = fmaf(a4[kk], wv[kk].x, avv[j][0]); avv[j][1] = fmaf(a4[kk], wv[kk].y, avv[j][1]);
          avv[j][2] = fmaf(a4[kk], wv[kk].z, avv[j][2]); avv[j][3] = fmaf(a4[kk], wv[kk].w, avv[j][3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      int tk = c0 + warp * CT + j;
      int row = warp * CT + j;
      if (tk < Te) {
        float4 gk = ld4(cg_qk + ((long long)b * Te + tk) * 256 + 128 + lane * 4);
        st4(Ks + row * 128 + lane * 4, make_float4(elu1(ak[j][0] + gk.x), elu1(ak[j][1] + gk.y),
                                                   elu1(ak[j][2] + gk.z), elu1(ak[j][3] + gk.w)));
        st4(Vs + row * 128 + lane * 4, make_float4((avv[j][0] + bv.x) * invS, (avv[j][1] + bv.y) * invS,
                                                   (avv[j][2] + bv.z) * invS, (avv[j][3] + bv.w) * invS));
      } else {
        st4(Ks + row * 128 + lane * 4, make_float4(0.f, 0.f, 0.f, 0.f));
        st4(Vs + row * 128 + lane * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    __syncthreads();
    for (int r = 0; r < CHUNK; ++r) {
      float kd = Ks[r * 128 + ah * 32 + ad];
      const float* vr = Vs + r * 128 + ah * 32 + av0;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 v4 = ld4(vr + i);
        kv[i] = fmaf(kd, v4.x, kv[i]); kv[i + 1] = fmaf(kd, v4.y, kv[i + 1]);
        kv[i + 2] = fmaf(kd, v4.z, kv[i + 2]); kv[i + 3] = fmaf(kd, v4.w, kv[i + 3]);
      }
      if (t < 128) ksum += Ks[r * 128 + t];
    }
    __syncthreads();
  }
  float* out = state + (long long)blockIdx.x * kStateFloats;
#pragma unroll
  for (int i = 0; i < 16; i += 4)
    st4(out + ah * 1024 + ad * 32 + av0 + i, make_float4(kv[i], kv[i + 1], kv[i + 2], kv[i + 3]));
  if (t < 128) out[4096 + t] = ksum;
}

cudaError_t launch_class_state_exact(const float* X, const float* cg_qk, float* state, int B, int Te, int npix,
                                     int S, const ClassLayerW& w, cudaStream_t st) {
  size_t smem = (size_t)(CW * CT * 128 + 2 * CHUNK * 128) * 4;
  {   // per-device function attribute: set on every launch (cheap), a process-wide flag would miss other devices
    cudaError_t e = cudaFuncSetAttribute(class_state_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
  }
  class_state_exact_kernel<<<B * npix, CW * 32, smem, st>>>(X, cg_qk, state, Te, npix, S, w);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Token-wise second half.  One CTA per (image, pixel), looping over the class axis in chunks.
__global__ void __launch_bounds__(CW * 32)
class_apply_exact_kernel(const float* __restrict__ X, float* __restrict__ Xout, const float* __restrict__ cg_qk,
                         const float* __restrict__ state, const float* __restrict__ pad_state, int Te, int npix,
                         int S, int out_mode, ClassLayerW w) {
  extern __shared__ __align__(16) float smem[];
  float* st_s = smem;                         // [4224] KV then Ksum (state + padding contribution)
  float* xn = smem + kStateFloats;            // [CW][CT][128]  LN output (GEMM A operand)
  float* qb = xn + CW * CT * 128;             // [CW][CT][128]  phi(q), later the MLP hidden chunk
  const int b = blockIdx.x / npix, p = blockIdx.x % npix;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kStateFloats; i += blockDim.x)
    st_s[i] = state[(long long)blockIdx.x * kStateFloats + i] + pad_state[i];
  __syncthreads();

  const float fS = (float)S;
  float4 g1 = ld4(w.ln1_g + lane * 4), be1 = ld4(w.ln1_b + lane * 4);
  float4 g2 = ld4(w.ln2_g + lane * 4), be2 = ld4(w.ln2_b + lane * 4);
  float* xw = xn + warp * CT * 128;
  float* qw = qb + warp * CT * 128;
  const int hh = lane >> 3;                   // head of this lane's 4 output features
  const float* kvh = st_s + hh * 1024 + (lane & 7) * 4;   // KV[hh][d][v0..v0+3] at + d*32
  const float* ksh = st_s + 4096 + hh * 32;

  for (int c0 = 0; c0 < Te; c0 += CHUNK) {
    const int tk0 = c0 + warp * CT;
    if (tk0 >= Te) continue;                  // warp-uniform; no block-level sync inside the loop
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      int tk = tk0 + j;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tk < Te) x = ld4(X + (((long long)b * Te + tk) * npix + p) * 128 + lane * 4);
      st4(xw + j * 128 + lane * 4, warp_layernorm128(x, g1, be1));
    }
    __syncwarp();
    // ---- q = LN1(x) Wq^T + guidance term ; phi(q)
    float acc[CT][4];
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
    for (int k = 0; k < 128; k += 4) {
      float4 w4[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) w4[kk] = ld4(w.wqkv_t + (k + kk) * 384 + lane * 4);
#pragma unroll
      for (int j = 0; j < CT; ++j) {
        float4 a = ld4(xw + j * 128 + k);
        float a4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          acc[j][0] = fmaf(a4[kk], w4[kk].x, acc[j][0]); acc[j][1] = fmaf(a4[kk], w4[kk].y, acc[j][1]);
          acc[j][2] = fmaf(a4[kk], w4[kk].z, acc[j][2]); acc[j][3] = fmaf(a4[kk], w4[kk].w, acc[j][3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      int tk = tk0 + j;
      float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tk < Te) gq = ld4(cg_qk + ((long long)b * Te + tk) * 256 + lane * 4);
      st4(qw + j * 128 + lane * 4, make_float4(elu1(acc[j][0] + gq.x), elu1(acc[j][1] + gq.y),
                                               elu1(acc[j][2] + gq.z), elu1(acc[j][3] + gq.w)));
    }
    __syncwarp();
    // ---- apply the state: num = phi(q) KV ; den = phi(q) . Ksum + eps      (model.py:283-284)
    float x1[CT][4];
    {
      float num[CT][4], den[CT];
#pragma unroll
      for (int j = 0; j < CT; ++j) { num[j][0] = num[j][1] = num[j][2] = num[j][3] = 0.0f; den[j] = 0.0f; }
      for (int d = 0; d < 32; ++d) {
        float4 kv4 = ld4(kvh + d * 32);
        float ks = ksh[d];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          float qd = qw[j * 128 + hh * 32 + d];
          num[j][0] = fmaf(qd, kv4.x, num[j][0]); num[j][1] = fmaf(qd, kv4.y, num[j][1]);
          num[j][2] = fmaf(qd, kv4.z, num[j][2]); num[j][3] = fmaf(qd, kv4.w, num[j][3]);
          den[j] = fmaf(qd, ks, den[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < CT; ++j) {
        int tk = tk0 + j;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tk < Te) x = ld4(X + (((long long)b * Te + tk) * npix + p) * 128 + lane * 4);
        float z = 1.0f / (den[j] + 1e-6f);
        x1[j][0] = x.x + num[j][0] * z * fS; x1[j][1] = x.y + num[j][1] * z * fS;
        x1[j][2] = x.z + num[j][2] * z * fS; x1[j][3] = x.w + num[j][3] * z * fS;
        st4(xw + j * 128 + lane * 4,
            warp_layernorm128(make_float4(x1[j][0], x1[j][1], x1[j][2], x1[j][3]), g2, be2));
      }
    }
    __syncwarp();
    // ---- MLP 128 -> 512 (ReLU) -> 128 (model.py:362-366)
    float acc2[CT][4];
#pragma unroll
    for (int j = 0; j < CT; ++j) acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.0f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int j = 0; j < CT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
      for (int k = 0; k < 128; k += 4) {
        float4 w4[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) w4[kk] = ld4(w.w1_t + (k + kk) * 512 + c * 128 + lane * 4);
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          float4 a = ld4(xw + j * 128 + k);
          float a4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            acc[j][0] = fmaf(a4[kk], w4[kk].x, acc[j][0]); acc[j][1] = fmaf(a4[kk], w4[kk].y, acc[j][1]);
            acc[j][2] = fmaf(a4[kk], w4[kk].z, acc[j][2]); acc[j][3] = fmaf(a4[kk], w4[kk].w, acc[j][3]);
          }
        }
      }
      float4 b1 = ld4(w.b1 + c * 128 + lane * 4);
#pragma unroll
      for (int j = 0; j < CT; ++j)
        st4(qw + j * 128 + lane * 4, make_float4(fmaxf(acc[j][0] + b1.x, 0.f), fmaxf(acc[j][1] + b1.y, 0.f),
                                                 fmaxf(acc[j][2] + b1.z, 0.f), fmaxf(acc[j][3] + b1.w, 0.f)));
      __syncwarp();
      for (int k = 0; k < 128; k += 4) {
        float4 w4[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) w4[kk] = ld4(w.w2_t + (c * 128 + k + kk) * 128 + lane * 4);
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          float4 a = ld4(qw + j * 128 + k);
          float a4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            acc2[j][0] = fmaf(a4[kk], w4[kk].x, acc2[j][0]); acc2[j][1] = fmaf(a4[kk], w4[kk].y, acc2[j][1]);
            acc2[j][2] = fmaf(a4[kk], w4[kk].z, acc2[j][2]); acc2[j][3] = fmaf(a4[kk], w4[kk].w, acc2[j][3]);
          }
        }
      }
      __syncwarp();
    }
    float4 b2 = ld4(w.b2 + lane * 4);
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      int tk = tk0 + j;
      if (tk >= Te) continue;
      long long off = (((long long)b * Te + tk) * npix + p) * 128 + lane * 4;
      float4 x2 = make_float4(x1[j][0] + (acc2[j][0] + b2.x), x1[j][1] + (acc2[j][1] + b2.y),
                              x1[j][2] + (acc2[j][2] + b2.z), x1[j][3] + (acc2[j][3] + b2.w));
      if (out_mode == 0) {   // pooling 1x1: avg-pool and the align_corners resize are identities
        float4 x = ld4(X + off);
        x2 = f4add(x, x2);
      }
      st4(Xout + off, x2);
    }
  }
}

cudaError_t launch_class_apply_exact(const float* X, float* Xout, const float* cg_qk, const float* state,
                                     const float* pad_state, int B, int Te, int npix, int S, int out_mode,
                                     const ClassLayerW& w, cudaStream_t st) {
  size_t smem = (size_t)(kStateFloats + 2 * CW * CT * 128) * 4;
  {   // per-device function attribute: set on every launch (cheap), a process-wide flag would miss other devices
    cudaError_t e = cudaFuncSetAttribute(class_apply_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
  }
  class_apply_exact_kernel<<<B * npix, CW * 32, smem, st>>>(X, Xout, cg_qk, state, pad_state, Te, npix, S,
                                                           out_mode, w);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// AvgPool2d(ph,pw) on token-major slices (model.py:375-385): Xp[s][py*Wp+px][c]
__global__ void avgpool_tokens_kernel(const float* __restrict__ X, float* __restrict__ Xp, long long nslice, int H,
                                      int W, int ph, int pw) {
  int Hp = H / ph, Wp = W / pw;
  long long total = nslice * Hp * Wp * 32;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c4 = (int)(i & 31);
  long long r = i >> 5;
  int pp = (int)(r % (Hp * Wp));
  long long s = r / (Hp * Wp);
  int py = pp / Wp, px = pp % Wp;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int dy = 0; dy < ph; ++dy)
    for (int dx = 0; dx < pw; ++dx)
      a = f4add(a, ld4(X + ((s * H + py * ph + dy) * W + px * pw + dx) * 128 + c4 * 4));
  float d = (float)(ph * pw);
  st4(Xp + r * 128 + c4 * 4, make_float4(a.x / d, a.y / d, a.z / d, a.w / d));
}
cudaError_t launch_avgpool_tokens(const float* X, float* Xp, long long nslice, int H, int W, int ph, int pw,
                                  cudaStream_t st) {
  long long total = nslice * (H / ph) * (W / pw) * 32;
  avgpool_tokens_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, Xp, nslice, H, W, ph, pw);
  return cudaGetLastError();
}

// X += bilinear(Xp, size=(H,W), align_corners=True)   (model.py:416, 423)
__global__ void upsample_add_kernel(float* __restrict__ X, const float* __restrict__ Xp, long long nslice, int H,
                                    int W, int Hp, int Wp) {
  long long total = nslice * H * W * 32;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c4 = (int)(i & 31);
  long long r = i >> 5;
  int pix = (int)(r % (H * W));
  long long s = r / (H * W);
  int y = pix / W, x = pix % W;
  float sh = H > 1 ? (float)(Hp - 1) / (float)(H - 1) : 0.0f;
  float sw = W > 1 ? (float)(Wp - 1) / (float)(W - 1) : 0.0f;
  float fy = sh * y, fx = sw * x;
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = y0 + (y0 < Hp - 1 ? 1 : 0), x1 = x0 + (x0 < Wp - 1 ? 1 : 0);
  float ly1 = fy - y0, lx1 = fx - x0, ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
  const float* base = Xp + s * Hp * Wp * 128 + c4 * 4;
  float4 v00 = ld4(base + (y0 * Wp + x0) * 128), v01 = ld4(base + (y0 * Wp + x1) * 128);
  float4 v10 = ld4(base + (y1 * Wp + x0) * 128), v11 = ld4(base + (y1 * Wp + x1) * 128);
  float4 cur = ld4(X + r * 128 + c4 * 4);
  float4 o;
  o.x = cur.x + (ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x));
  o.y = cur.y + (ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y));
  o.z = cur.z + (ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z));
  o.w = cur.w + (ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w));
  st4(X + r * 128 + c4 * 4, o);
}
cudaError_t launch_upsample_add(float* X, const float* Xp, long long nslice, int H, int W, int Hp, int Wp,
                                cudaStream_t st) {
  long long total = nslice * H * W * 32;
  upsample_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, Xp, nslice, H, W, Hp, Wp);
  return cudaGetLastError();
}

}  // namespace catseg
