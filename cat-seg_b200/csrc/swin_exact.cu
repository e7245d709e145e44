// EXACT (fp32 CUDA-core) Swin block with appearance guidance: one CTA per (slice, window).
//
// Reference: SwinTransformerBlock.forward (model.py:185-225), WindowAttention.forward (:86-114),
// shifted-window mask (:161-183), window_partition/reverse (:18-47).  The whole block for one
// 12x12 window (144 tokens x 128 ch) runs out of shared memory:
//   LN1 -> per head: q,k,v = LN(x) W^T + class-independent guidance term -> softmax(q k^T + mask) v
//       -> proj accumulated head by head -> + shortcut -> LN2 -> fc1/GELU/fc2 -> + residual.
// roll(-s)/window_partition/window_reverse/roll(+s) are index arithmetic: the token at window
// position (wy,wx,ly,lx) lives at pixel ((wy*12+ly+s)%24, (wx*12+lx+s)%24).
// The guidance half of the q/k projections only depends on (image, pixel); it is precomputed once
// per image and block (ag_qk, including the q/k biases) — SURVEY.md §7.2.
#include "common.cuh"
#include "internal.h"

namespace catseg {

namespace {
constexpr int WIN = 12, GRID = 24, NTOK = 144, NWARP = 16, TPW = 9;   // 16 warps x 9 tokens
constexpr int QPAD = 33;
constexpr int SM_BUFA = NTOK * 128;                 // floats
constexpr int SM_QKV = 3 * NTOK * QPAD;
constexpr int SM_O = NWARP * TPW * 32;
constexpr int SM_UNION = (SM_QKV + SM_O) > (NWARP * TPW * 128) ? (SM_QKV + SM_O) : (NWARP * TPW * 128);
constexpr size_t SWIN_SMEM = (size_t)(SM_BUFA + SM_UNION) * 4 + NTOK * 2 * sizeof(int);
}  // namespace

template <bool WITH_MLP>
__global__ void __launch_bounds__(512, 1)
swin_block_exact_kernel(float* __restrict__ X, const float* __restrict__ ag_qk, int Te, int shift, SwinBlockW w) {
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem;                         // [144][128]  LN1(x), later LN2(x1)
  float* bufQ = smem + SM_BUFA;               // [144][33]
  float* bufK = bufQ + NTOK * QPAD;
  float* bufV = bufK + NTOK * QPAD;
  float* bufO = bufV + NTOK * QPAD;           // [16][9][32]
  float* hbuf = smem + SM_BUFA;               // [16][9][128]  (aliases q/k/v/o after attention)
  int* tokpix = reinterpret_cast<int*>(smem + SM_BUFA + SM_UNION);   // [144]
  int* tokreg = tokpix + NTOK;                                        // [144]

  const int slice = blockIdx.x >> 2, win = blockIdx.x & 3;
  const int wy = win >> 1, wx = win & 1;
  const int b = slice / Te;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Xs = X + (long long)slice * (GRID * GRID) * 128;
  const float* ag = ag_qk + (long long)b * (GRID * GRID) * 256;

  if (threadIdx.x < NTOK) {
    int li = threadIdx.x;
    int sy = wy * WIN + li / WIN, sx = wx * WIN + li % WIN;       // coordinates on the shifted grid
    int oy = (sy + shift) % GRID, ox = (sx + shift) % GRID;        // roll(-shift): shifted[i] = x[(i+shift)%n]
    tokpix[li] = oy * GRID + ox;
    int rh = sy < GRID - WIN ? 0 : (sy < GRID - shift ? 1 : 2);   // model.py:166-173
    int rw = sx < GRID - WIN ? 0 : (sx < GRID - shift ? 1 : 2);
    tokreg[li] = rh * 3 + rw;
  }
  __syncthreads();

  // ---- LN1
  {
    float4 g = ld4(w.ln1_g + lane * 4), be = ld4(w.ln1_b + lane * 4);
#pragma unroll
    for (int j = 0; j < TPW; ++j) {
      int li = warp * TPW + j;
      float4 x = ld4(Xs + (long long)tokpix[li] * 128 + lane * 4);
      st4(bufA + li * 128 + lane * 4, warp_layernorm128(x, g, be));
    }
  }
  __syncwarp();

  float accp[TPW][4];
#pragma unroll
  for (int j = 0; j < TPW; ++j) accp[j][0] = accp[j][1] = accp[j][2] = accp[j][3] = 0.0f;
  const float scale = 0.17677669529663688110f;   // head_dim ** -0.5, head_dim = 32 (model.py:75)

  for (int h = 0; h < CATSEG_HEADS; ++h) {
    // ---- q, k, v for head h: lane = feature within the head
    {
      float aq[TPW], ak[TPW], av[TPW];
#pragma unroll
      for (int j = 0; j < TPW; ++j) aq[j] = ak[j] = av[j] = 0.0f;
      const float* wq = w.wqkv_t + h * 32 + lane;
      for (int k = 0; k < 128; k += 4) {
        float wqv[4], wkv[4], wvv[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          wqv[kk] = __ldg(wq + (k + kk) * 384);
          wkv[kk] = __ldg(wq + (k + kk) * 384 + 128);
          wvv[kk] = __ldg(wq + (k + kk) * 384 + 256);
        }
#pragma unroll
        for (int j = 0; j < TPW; ++j) {
          float4 a = ld4(bufA + (warp * TPW + j) * 128 + k);
          float av4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            aq[j] = fmaf(av4[kk], wqv[kk], aq[j]);
            ak[j] = fmaf(av4[kk], wkv[kk], ak[j]);
            av[j] = fmaf(av4[kk], wvv[kk], av[j]);
          }
        }
      }
      float bvv = __ldg(w.bv + h * 32 + lane);
#pragma unroll
      for (int j = 0; j < TPW; ++j) {
        int li = warp * TPW + j;
        const float* agp = ag + (long long)tokpix[li] * 256 + h * 32 + lane;
        bufQ[li * QPAD + lane] = (aq[j] + __ldg(agp)) * scale;
        bufK[li * QPAD + lane] = ak[j] + __ldg(agp + 128);
        bufV[li * QPAD + lane] = av[j] + bvv;
      }
    }
    __syncthreads();

    // ---- attention for this warp's 9 queries, 3 at a time; lane owns keys lane+32m
#pragma unroll 1
    for (int g3 = 0; g3 < 3; ++g3) {
      const int q0 = warp * TPW + g3 * 3;
      float s[3][5];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int m = 0; m < 5; ++m) s[i][m] = 0.0f;
      int krow[5];
#pragma unroll
      for (int m = 0; m < 5; ++m) { int r = lane + 32 * m; krow[m] = r < NTOK ? r : NTOK - 1; }
      for (int d = 0; d < 32; ++d) {
        float kd[5];
#pragma unroll
        for (int m = 0; m < 5; ++m) kd[m] = bufK[krow[m] * QPAD + d];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float qd = bufQ[(q0 + i) * QPAD + d];
#pragma unroll
          for (int m = 0; m < 5; ++m) s[i][m] = fmaf(qd, kd[m], s[i][m]);
        }
      }
      float p[3][5];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        int rq = tokreg[q0 + i];
        float mx = -INFINITY;
#pragma unroll
        for (int m = 0; m < 5; ++m) {
          int r = lane + 32 * m;
          float v = s[i][m];
          if (shift > 0 && tokreg[krow[m]] != rq) v += -100.0f;      // model.py:179 (-100, not -inf)
          if (r >= NTOK) v = -INFINITY;
          s[i][m] = v;
          mx = fmaxf(mx, v);
        }
        mx = warp_max(mx);
        float sum = 0.0f;
#pragma unroll
        for (int m = 0; m < 5; ++m) { p[i][m] = expf(s[i][m] - mx); sum += p[i][m]; }
        sum = warp_sum(sum);
#pragma unroll
        for (int m = 0; m < 5; ++m) p[i][m] = p[i][m] / sum;
      }
      float o[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const int cnt = m < 4 ? 32 : NTOK - 128;
        for (int jj = 0; jj < cnt; ++jj) {
          float vj = bufV[(m * 32 + jj) * QPAD + lane];
#pragma unroll
          for (int i = 0; i < 3; ++i) o[i] = fmaf(__shfl_sync(0xffffffffu, p[i][m], jj), vj, o[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) bufO[(warp * TPW + g3 * 3 + i) * 32 + lane] = o[i];
    }
    __syncwarp();
    // ---- proj, accumulated per head: y[:, n] += o_h[:, d] * Wp^T[h*32+d][n]
    for (int d = 0; d < 32; ++d) {
      float4 wp = ld4(w.wproj_t + (h * 32 + d) * 128 + lane * 4);
#pragma unroll
      for (int j = 0; j < TPW; ++j) {
        float ov = bufO[(warp * TPW + j) * 32 + d];
        accp[j][0] = fmaf(ov, wp.x, accp[j][0]);
        accp[j][1] = fmaf(ov, wp.y, accp[j][1]);
        accp[j][2] = fmaf(ov, wp.z, accp[j][2]);
        accp[j][3] = fmaf(ov, wp.w, accp[j][3]);
      }
    }
    __syncthreads();   // q/k/v of this head are dead before the next head overwrites them
  }

  // ---- x1 = shortcut + proj(attn) ; LN2(x1) -> bufA (rows are warp-private)
  {
    float4 bp = ld4(w.bproj + lane * 4);
    float4 g = ld4(w.ln2_g + lane * 4), be = ld4(w.ln2_b + lane * 4);
#pragma unroll
    for (int j = 0; j < TPW; ++j) {
      int li = warp * TPW + j;
      float* xp = Xs + (long long)tokpix[li] * 128 + lane * 4;
      float4 x = ld4(xp);
      float4 x1 = make_float4(x.x + (accp[j][0] + bp.x), x.y + (accp[j][1] + bp.y), x.z + (accp[j][2] + bp.z),
                              x.w + (accp[j][3] + bp.w));
      st4(xp, x1);
      st4(bufA + li * 128 + lane * 4, warp_layernorm128(x1, g, be));
    }
  }
  if (!WITH_MLP) return;
  __syncwarp();

  // ---- MLP 128 -> 512 (GELU) -> 128 in four hidden chunks of 128
  float acc2[TPW][4];
#pragma unroll
  for (int j = 0; j < TPW; ++j) acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.0f;
  float* hb = hbuf + warp * TPW * 128;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    float acc1[TPW][4];
#pragma unroll
    for (int j = 0; j < TPW; ++j) acc1[j][0] = acc1[j][1] = acc1[j][2] = acc1[j][3] = 0.0f;
    for (int k = 0; k < 128; k += 4) {
      float4 w4[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) w4[kk] = ld4(w.w1_t + (k + kk) * 512 + c * 128 + lane * 4);
#pragma unroll
      for (int j = 0; j < TPW; ++j) {
        float4 a = ld4(bufA + (warp * TPW + j) * 128 + k);
        float av4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          acc1[j][0] = fmaf(av4[kk], w4[kk].x, acc1[j][0]);
          acc1[j][1] = fmaf(av4[kk], w4[kk].y, acc1[j][1]);
          acc1[j][2] = fmaf(av4[kk], w4[kk].z, acc1[j][2]);
          acc1[j][3] = fmaf(av4[kk], w4[kk].w, acc1[j][3]);
        }
      }
    }
    float4 b1 = ld4(w.b1 + c * 128 + lane * 4);
#pragma unroll
    for (int j = 0; j < TPW; ++j)
      st4(hb + j * 128 + lane * 4, make_float4(gelu_erf(acc1[j][0] + b1.x), gelu_erf(acc1[j][1] + b1.y),
                                               gelu_erf(acc1[j][2] + b1.z), gelu_erf(acc1[j][3] + b1.w)));
    __syncwarp();
    for (int k = 0; k < 128; k += 4) {
      float4 w4[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) w4[kk] = ld4(w.w2_t + (c * 128 + k + kk) * 128 + lane * 4);
#pragma unroll
      for (int j = 0; j < TPW; ++j) {
        float4 a = ld4(hb + j * 128 + k);
        float av4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          acc2[j][0] = fmaf(av4[kk], w4[kk].x, acc2[j][0]);
          acc2[j][1] = fmaf(av4[kk], w4[kk].y, acc2[j][1]);
          acc2[j][2] = fmaf(av4[kk], w4[kk].z, acc2[j][2]);
          acc2[j][3] = fmaf(av4[kk], w4[kk].w, acc2[j][3]);
        }
      }
    }
    __syncwarp();
  }
  {
    float4 b2 = ld4(w.b2 + lane * 4);
#pragma unroll
    for (int j = 0; j < TPW; ++j) {
      int li = warp * TPW + j;
      float* xp = Xs + (long long)tokpix[li] * 128 + lane * 4;
      float4 x1 = ld4(xp);
      st4(xp, make_float4(x1.x + (acc2[j][0] + b2.x), x1.y + (acc2[j][1] + b2.y), x1.z + (acc2[j][2] + b2.z),
                          x1.w + (acc2[j][3] + b2.w)));
    }
  }
}

cudaError_t launch_swin_block_exact(float* X, const float* ag_qk, int nslice, int Te, int shift,
                                    const SwinBlockW& w, int with_mlp, cudaStream_t st) {
  {   // per-device function attribute: set on every launch (cheap), a process-wide flag would miss other devices
    cudaError_t e = cudaFuncSetAttribute(swin_block_exact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SWIN_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(swin_block_exact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SWIN_SMEM);
    if (e != cudaSuccess) return e;
  }
  if (with_mlp) swin_block_exact_kernel<true><<<nslice * 4, 512, SWIN_SMEM, st>>>(X, ag_qk, Te, shift, w);
  else swin_block_exact_kernel<false><<<nslice * 4, 512, SWIN_SMEM, st>>>(X, ag_qk, Te, shift, w);
  return cudaGetLastError();
}

}  // namespace catseg
