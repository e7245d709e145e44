// FAST front end on tcgen05: the 7x7 cost embedding and the three 3x3 guidance projections.
//
// Reference: Aggregator.corr_embed (model.py:654-659, conv1 = Conv2d(P, 128, 7, padding 3)) and the guidance
// projections guidance_projection / decoder_guidance_projection (model.py:615-630: Conv2d 3x3 + ReLU).
//
// Both kernels use the "row-shifted view" implicit GEMM of fast_decoder.cu: the input is staged once in shared
// memory as a canonical K-major image whose ROWS are zero-padded raster positions, so a vertical / diagonal
// convolution tap is the same image read from a shifted start row (descriptor address + offset * 16 bytes).
//
// Cost embedding (P = 1, one input channel).  The K axis carries the 7 HORIZONTAL taps: row pp of the image
// holds v[pp-3 .. pp+4]; the 7 vertical taps are 7 row-shifted views.  The cosine similarities feed everything
// downstream, so they are not rounded to bf16: v = hi + lo (two bf16 chunks = one K=16 MMA against [Whi | Whi])
// plus a second MMA of the same rows against [Wlo | 0]; the dropped lo*Wlo term is 2^-18 relative, i.e. the
// embedding is fp32-accurate (accumulation is fp32 in TMEM).  14 MMAs (N = 128) per 128-row tile.
//
// Guidance projections.  NCHW fp32 input, Ci in {1024, 512, 256}: the K loop runs over channel chunks of KC
// channels, double-buffered (the bf16 staging of chunk c+1 and the bulk copy of its 9 weight images overlap the
// MMAs of chunk c).  Output NHWC fp32 with bias + ReLU, exactly the layout the exact path produces.
#include "fast_common.cuh"
#include "internal.h"

namespace catseg {

using namespace fast;

// ================================================================================================ cost embedding
namespace {
constexpr int EM_THREADS = 512;
constexpr int EM_W = 24, EM_PW = EM_W + 6, EM_ROWS = 960, EM_P0 = 3 * EM_PW + 3, EM_NT = 6;
constexpr uint32_t EM_LBO_A = EM_ROWS * 16;                 // hi chunk | lo chunk
constexpr uint32_t EM_BIMG = 128 * 16 * 2;                  // one [128 x 16] bf16 weight image = 4 KiB
constexpr uint32_t EM_SM_B = 0;                             // 14 images: (dy, {hi, lo})
constexpr uint32_t EM_SM_A = EM_SM_B + 14 * EM_BIMG;
constexpr uint32_t EM_SM_RAW = EM_SM_A + 2 * EM_LBO_A;      // 2 x 1024 floats: padded raster, index pp + 8
constexpr uint32_t EM_SM_STAGE = EM_SM_RAW + 2 * 4096;
constexpr uint32_t EM_SM_BIAS = EM_SM_STAGE + STG_BYTES;
constexpr uint32_t EM_SM_BAR = EM_SM_BIAS + 512;
constexpr uint32_t EM_SMEM = EM_SM_BAR + 8 * 8 + 16;
static_assert(EM_P0 + EM_NT * 128 + 3 * EM_PW <= EM_ROWS && EM_ROWS + 8 + 8 <= 1024, "embedding image geometry");
static_assert(EM_SMEM <= 232448, "shared memory budget");
}  // namespace

__global__ void __launch_bounds__(EM_THREADS, 1)
embed_fast_kernel(const float* __restrict__ corr, const int32_t* __restrict__ classes, const __nv_bfloat16* __restrict__ bimg,
                  const float* __restrict__ bias, float* __restrict__ X, int nslice, int T, int Te) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* raw = reinterpret_cast<float*>(smem + EM_SM_RAW);
  float* stage = reinterpret_cast<float*>(smem + EM_SM_STAGE);
  float* s_bias = reinterpret_cast<float*>(smem + EM_SM_BIAS);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + EM_SM_BAR);
  uint64_t* bar_acc = bar_w + 1;                                      // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q = warp & 3, cq = warp >> 2;
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == 0;

  for (int i = tid; i < 2048; i += EM_THREADS) raw[i] = 0.0f;
  if (tid < 128) s_bias[tid] = bias[tid];
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) umma::mbar_init(&bar_w[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t sb = umma::smem_u32(smem);
  const uint32_t lane_addr = tm + ((uint32_t)(q * 32) << 16);
  const uint64_t d_a = umma::make_smem_desc(sb + EM_SM_A, EM_LBO_A, 128);
  const uint64_t d_b = umma::make_smem_desc(sb + EM_SM_B, 128 * 16, 128);
  if (issuer) {
    if (umma::elect_one()) {
      umma::mbar_expect_tx(bar_w, 14 * EM_BIMG);
      umma::bulk_g2s(smem + EM_SM_B, bimg, 14 * EM_BIMG, bar_w);
    }
    __syncwarp();
  }
  auto corr_of = [&](int s) { return corr + ((long long)(s / Te) * T + __ldg(classes + s)) * (EM_W * EM_W); };
  auto raster = [](int i) { return 8 + (i / EM_W + 3) * EM_PW + (i % EM_W) + 3; };
  if ((int)blockIdx.x < nslice) {
    const float* src = corr_of(blockIdx.x);
    for (int i = tid; i < EM_W * EM_W; i += EM_THREADS) raw[raster(i)] = __ldg(src + i);
  }
  __syncthreads();

  // one tile = 7 vertical taps x (hi|lo against Whi, hi against Wlo)
  auto issue_tile = [&](int t, int acc) {
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
      const uint64_t a = d_a + (uint64_t)(uint32_t)(EM_P0 + t * 128 + (dy - 3) * EM_PW);
      umma::mma_bf16_ss(tm + acc * 128, a, d_b + (uint64_t)((2 * dy) * (EM_BIMG >> 4)), IDESC_BF16_128x128, dy > 0 ? 1u : 0u);
      umma::mma_bf16_ss(tm + acc * 128, a, d_b + (uint64_t)((2 * dy + 1) * (EM_BIMG >> 4)), IDESC_BF16_128x128, 1u);
    }
    umma::mma_commit(&bar_acc[acc]);
  };

  uint32_t cnt[4] = {0, 0, 0, 0};
  bool w_ready = false;
  int cur = 0;
  for (int s = blockIdx.x; s < nslice; s += gridDim.x) {
    // ---- the next slice's similarities are fetched now and parked in registers
    const int sn = s + gridDim.x;
    float c0 = 0.0f, c1 = 0.0f;
    if (sn < nslice) {
      const float* src = corr_of(sn);
      c0 = __ldg(src + tid);
      if (tid < EM_W * EM_W - EM_THREADS) c1 = __ldg(src + EM_THREADS + tid);
    }
    // ---- A image: row pp = [hi(v[pp-3..pp+4]) | lo(...)]
    {
      const float* rb = raw + cur * 1024;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int r = tid + it * EM_THREADS;
        if (r < EM_ROWS) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float v0 = rb[r + 5 + 2 * i], v1 = rb[r + 6 + 2 * i];
            const __nv_bfloat16 h0 = __float2bfloat16(v0), h1 = __float2bfloat16(v1);
            hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo[i] = umma::pack_bf16x2(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
          }
          *reinterpret_cast<uint4*>(smem + EM_SM_A + r * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(smem + EM_SM_A + EM_LBO_A + r * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    if (issuer) {
      umma::fence_after_sync();
      if (!w_ready) { umma::mbar_wait(bar_w, 0); w_ready = true; }
      if (umma::elect_one()) {
        issue_tile(0, 0); issue_tile(1, 1); issue_tile(2, 2); issue_tile(3, 3);
      }
      __syncwarp();
    }
    if (sn < nslice) {
      float* rn = raw + (cur ^ 1) * 1024;
      rn[raster(tid)] = c0;
      if (tid < EM_W * EM_W - EM_THREADS) rn[raster(EM_THREADS + tid)] = c1;
    }
    float* Xs = X + (long long)s * (EM_W * EM_W) * 128;
#pragma unroll
    for (int t = 0; t < EM_NT; ++t) {
      const int acc = t & 3;
      umma::mbar_wait(&bar_acc[acc], cnt[acc] & 1u);
      ++cnt[acc];
      umma::fence_after_sync();
      {
        float v[32];
        umma::tmem_ld32(lane_addr + acc * 128 + cq * 32, v);
        const float* bb = s_bias + cq * 32;
        float* sp = stage + (q * 32 + lane) * STG_LD + cq * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          st4(sp + i, make_float4(v[i] + bb[i], v[i + 1] + bb[i + 1], v[i + 2] + bb[i + 2], v[i + 3] + bb[i + 3]));
      }
      umma::fence_before_sync();
      __syncthreads();
      if (t + 4 < EM_NT && issuer) {                // this accumulator is free again: tiles 4 and 5
        umma::fence_after_sync();
        if (umma::elect_one()) issue_tile(t + 4, acc);
        __syncwarp();
      }
      // coalesced copy-out: warp = row, 512 bytes per store instruction; halo rows are skipped
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = warp * 8 + i, pp = EM_P0 + t * 128 + row;
        const int yy = pp / EM_PW - 3, xx = pp % EM_PW - 3;
        if (xx >= 0 && xx < EM_W && yy < EM_W)
          st4(Xs + (long long)(yy * EM_W + xx) * 128 + lane * 4, ld4(stage + row * STG_LD + lane * 4));
      }
      __syncthreads();
    }
    cur ^= 1;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tm);
}

// Wt [49][128] fp32 (k = ky*7 + kx) -> 14 images [128 x 16] bf16: image 2*ky = [Whi(kx 0..6,0) | same], 2*ky+1 = [Wlo | 0]
__global__ void pack_embed_img_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ Wt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 14 * 128 * 16) return;
  const int img = i / 2048, n = (i % 2048) / 16, k = i % 16;
  const int ky = img >> 1, lo = img & 1, kx = k & 7;
  float w = kx < 7 ? Wt[(ky * 7 + kx) * 128 + n] : 0.0f;
  const __nv_bfloat16 hi = __float2bfloat16(w);
  __nv_bfloat16 v;
  if (!lo) v = hi;
  else v = k < 8 ? __float2bfloat16(w - __bfloat162float(hi)) : __float2bfloat16(0.0f);
  dst[(size_t)img * 2048 + (k >> 3) * 1024 + n * 8 + (k & 7)] = v;
}
cudaError_t launch_pack_embed_img(__nv_bfloat16* dst, const float* Wt, cudaStream_t st) {
  pack_embed_img_kernel<<<(14 * 2048 + 255) / 256, 256, 0, st>>>(dst, Wt);
  return cudaGetLastError();
}
cudaError_t launch_cost_embed_fast(const float* corr, const int32_t* classes, const __nv_bfloat16* bimg, const float* bias,
                                   float* X, int B, int T, int Te, int num_sms, cudaStream_t st) {
  {   // per-device function attribute: set on every launch (cheap), a process-wide flag would miss other devices
    cudaError_t e = cudaFuncSetAttribute(embed_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EM_SMEM);
    if (e != cudaSuccess) return e;
  }
  const int nslice = B * Te;
  const int grid = nslice < num_sms ? nslice : num_sms;
  if (grid <= 0) return cudaSuccess;
  embed_fast_kernel<<<grid, EM_THREADS, EM_SMEM, st>>>(corr, classes, bimg, bias, X, nslice, T, Te);
  return cudaGetLastError();
}

// ================================================================================================ guidance 3x3 conv
// SPLIT (PRECISE front end): fp16 hi + lo staging images, one weight image [Wh rows | Wl rows] per tap (2 NOUT rows):
// Ah [Wh | Wl] is one MMA of N = 2 NOUT whose column halves are summed in the epilogue, Al Wh a second one of N = NOUT
// (the scheme of the PRECISE decoder, fast_decoder.cu); fp32-accurate projections at two MMAs per k-step.
template <int NOUT, int WIN_, int BR, int KC, bool SPLIT = false>
struct GConvCfg {
  static constexpr int PW = WIN_ + 2, NP = (BR + 2) * PW, P0 = PW + 1;
  static constexpr int MROWS = (BR - 1) * PW + WIN_, NTILES = (MROWS + 127) / 128;
  static constexpr int ROWS = (P0 + NTILES * 128 + PW + 1 + 7) / 8 * 8;     // rows any tap view may touch
  static constexpr int NW = SPLIT ? 2 : 1;
  static constexpr uint32_t LBO_I = ROWS * 16;
  static constexpr uint32_t A_TERM = (KC / 8) * LBO_I;                      // one staged term (hi or lo)
  static constexpr uint32_t A_BYTES = NW * A_TERM;
  static constexpr uint32_t LBO_WT = NW * NOUT * 16;
  static constexpr uint32_t WIMG = NW * NOUT * KC * 2, W_BYTES = 9 * WIMG;
  static constexpr int NACC = NW * NOUT;
  static constexpr uint32_t SM_A = 0, SM_W = 2 * A_BYTES, SM_BAR = SM_W + 2 * W_BYTES;
  static constexpr uint32_t SMEM = SM_BAR + 8 * 8 + 16;
  static constexpr int NB = WIN_ / BR;
  static constexpr uint32_t IDESC = SPLIT ? umma::make_idesc_f16(128, NACC) : umma::make_idesc_bf16(128, NOUT);
  static constexpr uint32_t IDESC_LO = umma::make_idesc_f16(128, NOUT);
  static_assert(NTILES * NACC <= 256 && SMEM <= 232448 && WIN_ % BR == 0 && KC % 16 == 0, "guidance conv shape");
  static_assert(A_BYTES % 128 == 0 && W_BYTES % 16 == 0, "alignment");
};

template <int NOUT, int WIN_, int BR, int KC, bool SPLIT>
__global__ void __launch_bounds__(256, 1)
gconv_fast_kernel(const float* __restrict__ in, const void* __restrict__ wimg, const float* __restrict__ bias,
                  float* __restrict__ out, int Ci) {
  using C = GConvCfg<NOUT, WIN_, BR, KC, SPLIT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + C::SM_BAR);      // [2] weights of chunk landed
  uint64_t* bar_free = bar_full + 2;                                        // [2] MMAs of chunk done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q4 = warp & 3, tgrp = warp >> 2;
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == 0;
  const int b = blockIdx.x / C::NB, band = blockIdx.x % C::NB;
  const int nchunk = Ci / KC;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) umma::mbar_init(&bar_full[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<256>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t sb = umma::smem_u32(smem);
  const uint32_t lane_addr = tm + ((uint32_t)(q4 * 32) << 16);

  auto issue_wload = [&](int c) {           // one elected thread
    const int s = c & 1;
    umma::mbar_expect_tx(&bar_full[s], C::W_BYTES);
    umma::bulk_g2s(smem + C::SM_W + s * C::W_BYTES, reinterpret_cast<const uint8_t*>(wimg) + (size_t)c * C::W_BYTES,
                   C::W_BYTES, &bar_full[s]);
  };
  if (issuer) {
    if (umma::elect_one()) { issue_wload(0); if (nchunk > 1) issue_wload(1); }
    __syncwarp();
  }
  // rows beyond the staged band (only reachable by discarded accumulator rows) are zeroed once in both buffers
  for (int i = tid; i < 2 * C::NW * (KC / 8) * (C::ROWS - C::NP); i += 256) {
    const int bufk = i / (C::ROWS - C::NP), r = C::NP + i % (C::ROWS - C::NP);
    *reinterpret_cast<uint4*>(smem + C::SM_A + bufk * C::LBO_I + r * 16) = make_uint4(0u, 0u, 0u, 0u);
  }

  const int y_first = band * BR - 1;
  constexpr int NITEM = C::NP * (KC / 8);
  for (int c = 0; c < nchunk; ++c) {
    const int buf = c & 1;
    if (c >= 2) {
      umma::mbar_wait(&bar_free[buf], (uint32_t)(((c - 2) >> 1) & 1));      // MMAs of chunk c-2 have read this buffer
      if (issuer) {
        if (umma::elect_one()) issue_wload(c);
        __syncwarp();
      }
    }
    // ---- stage channels [c*KC, +KC) of the padded band: item = (position, 8-channel group), positions fastest
    const float* src = in + ((long long)b * Ci + (long long)c * KC) * (WIN_ * WIN_);
    uint8_t* abuf = smem + C::SM_A + buf * C::A_BYTES;
#pragma unroll 1
    for (int base = tid; base < NITEM; base += 256 * 2) {
      float v[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int idx = base + u * 256;
        const int pp = idx % C::NP, k8 = idx / C::NP;
        const int yy = y_first + pp / C::PW, xx = pp % C::PW - 1;
        const bool inb = idx < NITEM && yy >= 0 && yy < WIN_ && xx >= 0 && xx < WIN_;
        const float* p0 = src + (long long)(k8 * 8) * (WIN_ * WIN_) + (inb ? yy * WIN_ + xx : 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] = inb ? __ldg(p0 + (long long)j * (WIN_ * WIN_)) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int idx = base + u * 256;
        if (idx < NITEM) {
          const int pp = idx % C::NP, k8 = idx / C::NP;
          if constexpr (SPLIT) {
            uint4 hi, lo;
            umma::split_h2(v[u][0], v[u][1], hi.x, lo.x);
            umma::split_h2(v[u][2], v[u][3], hi.y, lo.y);
            umma::split_h2(v[u][4], v[u][5], hi.z, lo.z);
            umma::split_h2(v[u][6], v[u][7], hi.w, lo.w);
            *reinterpret_cast<uint4*>(abuf + k8 * C::LBO_I + pp * 16) = hi;
            *reinterpret_cast<uint4*>(abuf + C::A_TERM + k8 * C::LBO_I + pp * 16) = lo;
          } else {
            *reinterpret_cast<uint4*>(abuf + k8 * C::LBO_I + pp * 16) =
                make_uint4(umma::pack_bf16x2(v[u][0], v[u][1]), umma::pack_bf16x2(v[u][2], v[u][3]),
                           umma::pack_bf16x2(v[u][4], v[u][5]), umma::pack_bf16x2(v[u][6], v[u][7]));
          }
        }
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    if (issuer) {
      umma::fence_after_sync();
      umma::mbar_wait(&bar_full[buf], (uint32_t)((c >> 1) & 1));
      if (umma::elect_one()) {
        const uint64_t a0 = umma::make_smem_desc(sb + C::SM_A + buf * C::A_BYTES, C::LBO_I, 128);
        const uint64_t b0 = umma::make_smem_desc(sb + C::SM_W + buf * C::W_BYTES, C::LBO_WT, 128);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int off = (tap / 3 - 1) * C::PW + (tap % 3 - 1);
          const uint64_t bd = b0 + (uint64_t)(tap * (C::WIMG >> 4));
#pragma unroll 1
          for (int t = 0; t < C::NTILES; ++t) {
            const uint64_t ad = a0 + (uint64_t)(uint32_t)(C::P0 + off + t * 128);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              umma::mma_bf16_ss(tm + t * C::NACC, ad + (uint64_t)(k * 2 * (C::LBO_I >> 4)),
                                bd + (uint64_t)(k * 2 * (C::LBO_WT >> 4)), C::IDESC, (c > 0 || tap > 0 || k > 0) ? 1u : 0u);
            if constexpr (SPLIT) {
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)      // lo activations x the hi rows of the image
                umma::mma_f16_ss(tm + t * C::NACC, ad + (uint64_t)((C::A_TERM >> 4) + k * 2 * (C::LBO_I >> 4)),
                                 bd + (uint64_t)(k * 2 * (C::LBO_WT >> 4)), C::IDESC_LO, 1u);
            }
          }
        }
        umma::mma_commit(&bar_free[buf]);
      }
      __syncwarp();
    }
  }
  // ---- epilogue: all MMAs done when the last chunk's commit lands (commits complete in issue order)
  {
    const int c = nchunk - 1;
    umma::mbar_wait(&bar_free[c & 1], (uint32_t)((c >> 1) & 1));
    umma::fence_after_sync();
  }
  for (int t = tgrp; t < C::NTILES; t += 2) {
    const int pr = C::P0 + t * 128 + q4 * 32 + lane;
    const int yl = pr / C::PW - 1, xl = pr % C::PW - 1;
    const bool valid = (pr < C::P0 + C::MROWS) && xl >= 0 && xl < WIN_;
    float* o = out + ((long long)b * (WIN_ * WIN_) + (long long)(band * BR + yl) * WIN_ + xl) * NOUT;
#pragma unroll
    for (int c0 = 0; c0 < NOUT; c0 += 16) {
      float v[16];
      umma::tmem_ld16(lane_addr + t * C::NACC + c0, v);
      if constexpr (SPLIT) {
        float v2[16];
        umma::tmem_ld16(lane_addr + t * C::NACC + NOUT + c0, v2);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += v2[i];
      }
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          st4(o + c0 + i, make_float4(fmaxf(v[i] + __ldg(bias + c0 + i), 0.f), fmaxf(v[i + 1] + __ldg(bias + c0 + i + 1), 0.f),
                                      fmaxf(v[i + 2] + __ldg(bias + c0 + i + 2), 0.f), fmaxf(v[i + 3] + __ldg(bias + c0 + i + 3), 0.f)));
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tm);
}

// Wt [Ci*9][Co] fp32 (k = ci*9 + tap) -> per channel chunk c: 9 tap images [Co x KC] bf16 canonical (LBO = Co*16)
__global__ void pack_gconv_img_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ Wt, int Ci, int Co, int KC) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Ci * 9 * Co) return;
  const int k = (int)(i % KC);
  long long r = i / KC;
  const int n = (int)(r % Co); r /= Co;
  const int tap = (int)(r % 9);
  const int c = (int)(r / 9);
  const int ci = c * KC + k;
  dst[((size_t)c * 9 + tap) * ((size_t)Co * KC) + (size_t)(k >> 3) * (Co * 8) + n * 8 + (k & 7)] =
      __float2bfloat16(Wt[((long long)ci * 9 + tap) * Co + n]);
}
// PRECISE: per channel chunk c, 9 tap images [2 Co x KC] fp16: rows [0, Co) = hi term, [Co, 2 Co) = lo term
__global__ void pack_gconv_img_split_kernel(__half* __restrict__ dst, const float* __restrict__ Wt, int Ci, int Co, int KC) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Ci * 9 * Co) return;
  const int k = (int)(i % KC);
  long long r = i / KC;
  const int n = (int)(r % Co); r /= Co;
  const int tap = (int)(r % 9);
  const int c = (int)(r / 9);
  const int ci = c * KC + k;
  const float w = Wt[((long long)ci * 9 + tap) * Co + n];
  const __half h = __float2half_rn(w);
  __half* base = dst + ((size_t)c * 9 + tap) * ((size_t)2 * Co * KC) + (size_t)(k >> 3) * (2 * Co * 8) + (k & 7);
  base[n * 8] = h;
  base[(Co + n) * 8] = __float2half_rn(w - __half2float(h));
}
cudaError_t launch_pack_gconv_img_split(__half* dst, const float* Wt, int Ci, int Co, int KC, cudaStream_t st) {
  const long long n = (long long)Ci * 9 * Co;
  pack_gconv_img_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, Wt, Ci, Co, KC);
  return cudaGetLastError();
}
cudaError_t launch_pack_gconv_img(__nv_bfloat16* dst, const float* Wt, int Ci, int Co, int KC, cudaStream_t st) {
  const long long n = (long long)Ci * 9 * Co;
  pack_gconv_img_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, Wt, Ci, Co, KC);
  return cudaGetLastError();
}

template <int NOUT, int WIN_, int BR, int KC, bool SPLIT = false>
static cudaError_t run_gconv(const float* in, const void* wimg, const float* bias, float* out, int B, int Ci,
                             cudaStream_t st) {
  using C = GConvCfg<NOUT, WIN_, BR, KC, SPLIT>;
  {   // per-device function attribute: set on every launch (cheap), a process-wide flag would miss other devices
    cudaError_t e = cudaFuncSetAttribute(gconv_fast_kernel<NOUT, WIN_, BR, KC, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM);
    if (e != cudaSuccess) return e;
  }
  if (Ci % KC != 0) return cudaErrorInvalidValue;
  gconv_fast_kernel<NOUT, WIN_, BR, KC, SPLIT><<<B * C::NB, 256, C::SMEM, st>>>(in, wimg, bias, out, Ci);
  return cudaGetLastError();
}

int gconv_split_kc(int which) { return which == 0 ? 16 : 32; }
cudaError_t launch_gconv_split(int which, const float* in, const __half* wimg, const float* bias, float* out, int B, int Ci,
                               cudaStream_t st) {
  if (which == 0) return run_gconv<128, 24, 4, 16, true>(in, wimg, bias, out, B, Ci, st);
  if (which == 1) return run_gconv<32, 48, 2, 32, true>(in, wimg, bias, out, B, Ci, st);
  return run_gconv<16, 96, 2, 32, true>(in, wimg, bias, out, B, Ci, st);
}

int gconv_fast_kc(int which) { return which == 0 ? 32 : 64; }
bool gconv_fast_supported(int which, int Ci, int H, int W, int Co) {
  if (H != W || Ci % gconv_fast_kc(which) != 0) return false;
  if (which == 0) return H == 24 && Co == 128;
  if (which == 1) return H == 48 && Co == 32;
  return H == 96 && Co == 16;
}
cudaError_t launch_gconv_fast(int which, const float* in, const __nv_bfloat16* wimg, const float* bias, float* out, int B,
                              int Ci, cudaStream_t st) {
  if (which == 0) return run_gconv<128, 24, 4, 32>(in, wimg, bias, out, B, Ci, st);
  if (which == 1) return run_gconv<32, 48, 2, 64>(in, wimg, bias, out, B, Ci, st);
  return run_gconv<16, 96, 2, 64>(in, wimg, bias, out, B, Ci, st);
}

}  // namespace catseg
