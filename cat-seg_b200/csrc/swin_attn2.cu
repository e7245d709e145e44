// Window-attention half of the Swin block on tcgen05, second generation: one kernel for the FAST (single fp16 operands)
// and the PRECISE (hi + lo fp16 pairs on the value path) modes.
//   x1 = x + proj( softmax( (LN1(x) Wq^T + gq) s (LN1(x) Wk^T + gk)^T + mask ) (LN1(x) Wv^T + bv) )
// Reference: SwinTransformerBlock.forward (model.py:185-222), WindowAttention.forward (:86-114).
//
// One 12x12 window (144 tokens) per pass.  Token-wise GEMMs are TRANSPOSED (features on the UMMA M axis, the 144 tokens
// on N): no M padding, and ALL FOUR heads of q, k, v come out of three M = 128 GEMMs (the first-generation kernel ran
// one M = 128 GEMM per head with 32 dead rows).
//   phase 1   Q^T K^T V^T (128 x 144 each) = Wq|Wk|Wv . LN1(x)^T            TMEM [0,432)
//             q / k: one fp16 term (attention weights, tools/precision_study.py); v: hi+lo pairs, 3 products (PRECISE)
//   epilogue  + guidance term (fp32, read from L2 in coalesced 512-byte pieces: class independent, shared by the 256 classes of an image)
//             -> per-head images [token group][32 d][8 tok]: MN-major A/B operands of S = Q K^T and K-major B of O = P V
//   per head  S (two query tiles; queries 128..143 replicated into all four TMEM lane quarters so that all 16 warps share
//             the second tile) -> softmax with the -100 shift mask by index arithmetic -> P (single term) -> O = P (Vh + Vl)
//             -> O_h / rowsum as hi+lo images -> Y^T += Wp[:, 32h:32h+32] . O_h^T (accumulated over the heads in TMEM)
//   end       x1 = x + Y + b_proj
// Shared memory is time-shared: the weight images of phase 1 land (TMA) in the regions that later hold the Q/K/V images;
// LN1(x) hi/lo shares its region with P, the replicated query tiles and the per-head proj slice; O_h overwrites Q_h / K_h.
// TMEM: phase 1 [0,432); then S0 [0,144) S1 [144,288) O0 [288,320) O1 [320,352) Y^T [352,496).
#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "split_common.cuh"

namespace catseg {

using namespace fast;

namespace {
constexpr int A2_THREADS = 512;    // 16 warps: TMEM lane quarter q4 = warp & 3, work group grp = warp >> 2
constexpr int NTOK = 144, WIN = 12, GRID = 24;
constexpr uint32_t LBO_X = NTOK * 16 + 16;                 // 2320: LN tiles (padded: the LN prologue stores across chunks)
constexpr uint32_t LBO_P = NTOK * 16;                      // 2304: P and O images (thread = row stores)
constexpr uint32_t HEAD_IMG = 18 * 512;                    // 9216: [18 token groups][32 d][8 tok] fp16
constexpr uint32_t R1 = 0;                                 // LN1(x) hi | lo   ->   P | Q1 slots | proj slice
constexpr uint32_t R1_BYTES = 2 * 16 * LBO_X;              // 74240
constexpr uint32_t SM_XH = R1, SM_XL = R1 + 16 * LBO_X;
constexpr uint32_t SM_P = R1;                              // 18 chunks x LBO_P = 41472
constexpr uint32_t SM_Q1 = SM_P + 18 * LBO_P;              // 2 slots x 8192 (head parity)
constexpr uint32_t SM_WP = SM_Q1 + 2 * 8192;               // proj slice of the current head: hi 8192 | lo 8192
static_assert(SM_WP + 16384 <= R1 + R1_BYTES, "region 1");
constexpr uint32_t SM_Q = R1 + R1_BYTES;                   // phase 1: Wq image; then Q images of the 4 heads; then O hi
constexpr uint32_t SM_K = SM_Q + 4 * HEAD_IMG;             //          Wk;       K images;                  then O lo
constexpr uint32_t SM_VH = SM_K + 4 * HEAD_IMG;            //          Wv hi;    V hi images
constexpr uint32_t SM_VL = SM_VH + 4 * HEAD_IMG;           //          Wv lo;    V lo images
constexpr uint32_t SM_MISC = SM_VL + 4 * HEAD_IMG;
// misc: tokpix[144] (int) | red[4][144] | rsum[4][144] | max1[16][16] | sum1[16][16] | ln g,b [256] | bv[128] | bproj[128]
constexpr uint32_t SM_BAR = SM_MISC + (144 + 2 * 4 * 144 + 2 * 256 + 256 + 128 + 128) * 4;
constexpr uint32_t A2_SMEM = SM_BAR + 12 * 8 + 16;
static_assert(A2_SMEM <= 232448, "shared memory budget");
static_assert(WIMG_BYTES <= 4 * HEAD_IMG, "a weight image fits the region it is staged in");
constexpr uint32_t TM_QT = 0, TM_KT = 144, TM_VT = 288;
constexpr uint32_t TM_S0 = 0, TM_S1 = 144, TM_O0 = 288, TM_O1 = 320, TM_Y = 352;
constexpr uint32_t IDESC_T = umma::make_idesc_f16(128, 144, 0, 0);      // W . X^T
constexpr uint32_t IDESC_S = umma::make_idesc_f16(128, 144, 1, 1);      // S = Q K^T (both MN-major images)
constexpr uint32_t IDESC_PV = umma::make_idesc_f16(128, 32, 0, 0);      // O = P V
}  // namespace

template <bool SPLIT>
__global__ void __launch_bounds__(A2_THREADS, 1)
swin_attn2_kernel(float* __restrict__ X, const float* __restrict__ agT, int nwin_total, int Te, int shift, SwinAttn2W w, int dbgmask) {
  extern __shared__ __align__(1024) uint8_t smem[];
  int* tokpix = reinterpret_cast<int*>(smem + SM_MISC);
  float* red = reinterpret_cast<float*>(tokpix + 144);        // row max per key quarter [4][144]
  float* rsum = red + 4 * 144;                                 // row sum per key quarter [4][144]
  float* max1 = rsum + 4 * 144;                                // tile 1: row max per 9-key part [16][16]
  float* sum1 = max1 + 256;
  float* s_g = sum1 + 256;
  float* s_be = s_g + 128;
  float* s_bv = s_be + 128;
  float* s_bp = s_bv + 128;
  uint64_t* bar_wq = reinterpret_cast<uint64_t*>(smem + SM_BAR);   // Wq image landed
  uint64_t* bar_wk = bar_wq + 1;
  uint64_t* bar_wv = bar_wq + 2;                                     // Wv hi (+ lo)
  uint64_t* bar_wp = bar_wq + 3;                                     // proj slice of the current head
  uint64_t* bar_a = bar_wq + 4;                                      // phase-1 GEMMs done
  uint64_t* bar_s = bar_wq + 5;                                      // S(h) done
  uint64_t* bar_o = bar_wq + 6;                                      // O(h) done
  uint64_t* bar_yp = bar_wq + 7;                                     // proj(h), h < 3, done (slice buffer free)
  uint64_t* bar_y = bar_wq + 8;                                      // proj(3) done: Y^T complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_wq + 10);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q4 = warp & 3, grp = warp >> 2;
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == 15;   // least epilogue work (see fast_swin_attn.cu)

  if (tid < 128) { s_g[tid] = w.ln_g[tid]; s_be[tid] = w.ln_b[tid]; s_bv[tid] = w.bv[tid]; s_bp[tid] = w.bproj[tid]; }
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) umma::mbar_init(&bar_wq[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t sb = umma::smem_u32(smem);
  const uint32_t lane_addr = tm + ((uint32_t)(q4 * 32) << 16);
  const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(w.wimg);   // Wq | Wk | Wv hi | Wv lo | 4 x (Wp_h hi 8K | lo 8K)
  const uint64_t d_xh = umma::make_smem_desc(sb + SM_XH, LBO_X, 128), d_xl = umma::make_smem_desc(sb + SM_XL, LBO_X, 128);
  const uint64_t d_wq = umma::make_smem_desc(sb + SM_Q, LBO_W, 128), d_wk = umma::make_smem_desc(sb + SM_K, LBO_W, 128);
  const uint64_t d_wvh = umma::make_smem_desc(sb + SM_VH, LBO_W, 128), d_wvl = umma::make_smem_desc(sb + SM_VL, LBO_W, 128);
  const uint64_t d_qimg = umma::make_smem_desc(sb + SM_Q, 128, 512), d_kimg = umma::make_smem_desc(sb + SM_K, 128, 512);
  const uint64_t d_q1 = umma::make_smem_desc(sb + SM_Q1, 128, 512);
  const uint64_t d_p = umma::make_smem_desc(sb + SM_P, LBO_P, 128);
  const uint64_t d_vh = umma::make_smem_desc(sb + SM_VH, 512, 128), d_vl = umma::make_smem_desc(sb + SM_VL, 512, 128);
  const uint64_t d_oh = umma::make_smem_desc(sb + SM_Q, LBO_P, 128), d_ol = umma::make_smem_desc(sb + SM_K, LBO_P, 128);
  const uint64_t d_wph = umma::make_smem_desc(sb + SM_WP, LBO_W, 128), d_wpl = umma::make_smem_desc(sb + SM_WP + 8192, LBO_W, 128);

  // ---- asynchronous weight loads (one elected lane of the issuing warp)
  auto load_qk = [&]() {
    umma::mbar_expect_tx(bar_wq, WIMG_BYTES);
    umma::bulk_g2s(smem + SM_Q, wsrc, WIMG_BYTES, bar_wq);
    umma::mbar_expect_tx(bar_wk, WIMG_BYTES);
    umma::bulk_g2s(smem + SM_K, wsrc + WIMG_BYTES, WIMG_BYTES, bar_wk);
  };
  auto load_v = [&]() {
    umma::mbar_expect_tx(bar_wv, SPLIT ? 2 * WIMG_BYTES : WIMG_BYTES);
    umma::bulk_g2s(smem + SM_VH, wsrc + 2 * WIMG_BYTES, WIMG_BYTES, bar_wv);
    if (SPLIT) umma::bulk_g2s(smem + SM_VL, wsrc + 3 * WIMG_BYTES, WIMG_BYTES, bar_wv);
  };
  auto load_wp = [&](int h) {
    umma::mbar_expect_tx(bar_wp, SPLIT ? 16384 : 8192);
    umma::bulk_g2s(smem + SM_WP, wsrc + 4 * WIMG_BYTES + h * 16384, SPLIT ? 16384 : 8192, bar_wp);
  };
  if (issuer && (long long)blockIdx.x < nwin_total) {
    if (umma::elect_one()) { load_qk(); load_v(); }
    __syncwarp();
  }
  // the unused row groups of the two Q1 slots are written once per window (the region is shared with the LN tiles)

  uint32_t ph_w = 0, ph_wp = 0, ph_a = 0, ph_s = 0, ph_o = 0, ph_yp = 0, ph_y = 0;
  const float scale = 0.17677669529663688110f;

  for (long long wi = blockIdx.x; wi < nwin_total; wi += gridDim.x) {
    const int slice = (int)(wi >> 2), win = (int)(wi & 3);
    const int wy = win >> 1, wx = win & 1;
    const int b = slice / Te;
    float* Xs = X + (long long)slice * (GRID * GRID) * 128;
    if (tid < NTOK) {
      int sy = wy * WIN + tid / WIN, sx = wx * WIN + tid % WIN;
      tokpix[tid] = ((sy + shift) % GRID) * GRID + (sx + shift) % GRID;
    }
    __syncthreads();
    // ---- LN1 -> XN hi (| lo); the NEXT window's rows are prefetched into L2 meanwhile
    {
      const long long wn = wi + gridDim.x;
      if (wn < nwin_total) {
        const int nsl = (int)(wn >> 2), nwy = (int)((wn >> 1) & 1), nwx = (int)(wn & 1);
        for (int i = tid; i < NTOK * 4; i += A2_THREADS) {
          const int r = i >> 2, sy = nwy * WIN + r / WIN, sx = nwx * WIN + r % WIN;
          const int pix = ((sy + shift) % GRID) * GRID + (sx + shift) % GRID;
          umma::prefetch_l2(X + ((long long)nsl * (GRID * GRID) + pix) * 128 + (i & 3) * 32);
        }
      }
    }
    {
      const float4 g = ld4(s_g + lane * 4), be = ld4(s_be + lane * 4);
      const int r0 = warp * 9;                       // 16 warps x 9 rows, all 9 loads in flight
      float4 x[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) x[i] = ld4(Xs + (long long)tokpix[r0 + i] * 128 + lane * 4);
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const float4 y = warp_layernorm128_fast(x[i], g, be);
        const uint32_t off = (lane >> 1) * LBO_X + (r0 + i) * 16 + (lane & 1) * 8;
        if constexpr (SPLIT) {
          uint2 hi, lo;
          umma::split_h2(y.x, y.y, hi.x, lo.x);
          umma::split_h2(y.z, y.w, hi.y, lo.y);
          *reinterpret_cast<uint2*>(smem + SM_XH + off) = hi;
          *reinterpret_cast<uint2*>(smem + SM_XL + off) = lo;
        } else {
          *reinterpret_cast<uint2*>(smem + SM_XH + off) = make_uint2(umma::pack_h2(y.x, y.y), umma::pack_h2(y.z, y.w));
        }
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- phase 1: Q^T, K^T (single terms), V^T (SPLIT: three products)
    if (issuer) {
      umma::fence_after_sync();
      umma::mbar_wait(bar_wq, ph_w);
      if (umma::elect_one()) issue_gemm_k128_desc(tm + TM_QT, d_wq, LBO_W, d_xh, LBO_X, IDESC_T, false);
      __syncwarp();
      umma::mbar_wait(bar_wk, ph_w);
      if (umma::elect_one()) issue_gemm_k128_desc(tm + TM_KT, d_wk, LBO_W, d_xh, LBO_X, IDESC_T, false);
      __syncwarp();
      umma::mbar_wait(bar_wv, ph_w);
      if (umma::elect_one()) {
        issue_gemm_k128_desc(tm + TM_VT, d_wvh, LBO_W, d_xh, LBO_X, IDESC_T, false);
        if (SPLIT) {
          if (!(dbgmask & 1)) issue_gemm_k128_desc(tm + TM_VT, d_wvh, LBO_W, d_xl, LBO_X, IDESC_T, true);
          if (!(dbgmask & 2)) issue_gemm_k128_desc(tm + TM_VT, d_wvl, LBO_W, d_xh, LBO_X, IDESC_T, true);
        }
        umma::mma_commit(bar_a);
      }
      __syncwarp();
    }
    ph_w ^= 1;
    // ---- q/k/v epilogue.  Lane quarter = head (features 32 q4 .. +32), thread = feature d = lane; the 54 (matrix, token
    //      group) units are dealt round-robin to the four groups.  The guidance terms of the first units are fetched
    //      (fp32, [image][window][256 features][144 tokens]) BEFORE waiting for the MMAs.
    {
      // guidance terms [image][window][18 token groups][2 halves][256 features][4 tokens]: the 32 lanes (= 32 features) of a
      // warp read 512 contiguous bytes per load, 1 KiB per unit (with the token axis innermost per feature every lane touched
      // its own 128-byte line: 64 lines per unit, ~9 K LSU cycles per window; with 8 tokens per feature innermost each 16-byte
      // load still fetched every sector twice)
      const float* ag0 = agT + ((long long)b * 4 + win) * (18 * 256 * 8) + (q4 * 32 + lane) * 4;
      constexpr int NU = 14, D = 4;
      float4 ga[D][2];
      auto fetch = [&](int i, float4* dst) {
        const int u = grp + 4 * i;
        if (u < 36) {                                  // q (u < 18) and k units carry a guidance term
          const float* p = ag0 + ((u % 18) * 512 + (u >= 18 ? 128 : 0)) * 4;
          dst[0] = ld4(p); dst[1] = ld4(p + 256 * 4);
        }
      };
#pragma unroll
      for (int i = 0; i < D - 1; ++i) fetch(i, ga[i]);
      umma::mbar_wait(bar_a, ph_a); ph_a ^= 1;
      umma::fence_after_sync();
      if (issuer) {                                    // LN tiles are dead: the proj slice of head 0 may land in region 1
        if (umma::elect_one()) load_wp(0);
        __syncwarp();
      }
      const float bvv = s_bv[q4 * 32 + lane];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        if (i + D - 1 < NU) fetch(i + D - 1, ga[(i + D - 1) % D]);
        const int u = grp + 4 * i;
        if (u < 54) {                                  // warp-uniform
          const int m = u / 18, tg = u % 18;
          float v[8];
          umma::tmem_ld8(lane_addr + m * 144 + tg * 8, v);
          if (m < 2) {
            const float4 a0 = ga[i % D][0], a1 = ga[i % D][1];
            v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
            v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
            if (m == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] *= scale;
            }
            const uint4 pk = make_uint4(umma::pack_h2(v[0], v[1]), umma::pack_h2(v[2], v[3]), umma::pack_h2(v[4], v[5]),
                                        umma::pack_h2(v[6], v[7]));
            *reinterpret_cast<uint4*>(smem + (m == 0 ? SM_Q : SM_K) + q4 * HEAD_IMG + tg * 512 + lane * 16) = pk;
            if (m == 0 && tg >= 16 && q4 < 2) {        // queries 128..143 of heads 0, 1 -> Q1 slot q4, row groups {4c, 4c+1}
#pragma unroll
              for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(smem + SM_Q1 + q4 * 8192 + (4 * c + tg - 16) * 512 + lane * 16) = pk;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += bvv;
            if constexpr (SPLIT) {
              uint4 hi, lo;
              umma::split_h2(v[0], v[1], hi.x, lo.x);
              umma::split_h2(v[2], v[3], hi.y, lo.y);
              umma::split_h2(v[4], v[5], hi.z, lo.z);
              umma::split_h2(v[6], v[7], hi.w, lo.w);
              *reinterpret_cast<uint4*>(smem + SM_VH + q4 * HEAD_IMG + tg * 512 + lane * 16) = hi;
              *reinterpret_cast<uint4*>(smem + SM_VL + q4 * HEAD_IMG + tg * 512 + lane * 16) = lo;
            } else {
              *reinterpret_cast<uint4*>(smem + SM_VH + q4 * HEAD_IMG + tg * 512 + lane * 16) =
                  make_uint4(umma::pack_h2(v[0], v[1]), umma::pack_h2(v[2], v[3]), umma::pack_h2(v[4], v[5]), umma::pack_h2(v[6], v[7]));
            }
          }
        }
      }
      // Q1 slots: row groups 4c+2, 4c+3 (c = 0..3) are zero (they multiply nothing that is read back)
      for (int i = tid; i < 2 * 8 * 32; i += A2_THREADS) {
        const int slot = i >> 8, r = (i >> 5) & 7, c = i & 31;
        const int g16 = (r >> 1) * 4 + 2 + (r & 1);
        *reinterpret_cast<uint4*>(smem + SM_Q1 + slot * 8192 + g16 * 512 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    if (issuer) {                                      // S(0)
      umma::fence_after_sync();
      if (umma::elect_one()) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma::mma_f16_ss(tm + (mt ? TM_S1 : TM_S0), (mt ? d_q1 : d_qimg) + (uint64_t)((k * 256) >> 4),
                             d_kimg + (uint64_t)((k * 256) >> 4), IDESC_S, k > 0);
        umma::mma_commit(bar_s);
      }
      __syncwarp();
    }

    for (int h = 0; h < 4; ++h) {
      umma::mbar_wait(bar_s, ph_s); ph_s ^= 1;
      umma::fence_after_sync();
      // ---- softmax (see fast_swin_attn.cu for the work split): tile 0 with four threads per query row, tile 1 shared by
      //      all 16 warps through the replicated query image
      {
        const int part = grp * 4 + q4, key0 = part * 9, r1 = lane & 15;
        const int qrow = 128 + r1, qly = qrow / WIN, qlx = qrow % WIN;
        auto load_tile1 = [&](float* tv) {
          umma::tmem_ld16(lane_addr + TM_S1 + key0, tv);
          if (shift > 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) {
              const int kly = (key0 + i) / WIN, klx = (key0 + i) % WIN;
              const bool masked = (wy == 1 && ((qly >= 6) != (kly >= 6))) || (wx == 1 && ((qlx >= 6) != (klx >= 6)));
              tv[i] += masked ? -100.0f : 0.0f;
            }
          }
        };
        {
          float tv[16];
          load_tile1(tv);
          float mx1 = tv[0];
#pragma unroll
          for (int i = 1; i < 9; ++i) mx1 = fmaxf(mx1, tv[i]);
          if (lane < 16) max1[part * 16 + r1] = mx1;
        }
        const int row = q4 * 32 + lane;
        const int kq = grp;
        const uint32_t s_addr = lane_addr + TM_S0 + kq * 36;
        float add0 = 0.0f, add1 = 0.0f;
        if (shift > 0) {
          const int ly = row / WIN, lx = row % WIN;
          const bool rowmask = (wy == 1) && ((ly >= 6) != (kq >= 2));
          add0 = (rowmask || (wx == 1 && lx >= 6)) ? -100.0f : 0.0f;
          add1 = (rowmask || (wx == 1 && lx < 6)) ? -100.0f : 0.0f;
        }
        float sv[36];
        float mx = -INFINITY;
        umma::tmem_ld32(s_addr, sv);
        umma::tmem_ld4(s_addr + 32, sv + 32);
        if (shift > 0) {
#pragma unroll
          for (int i = 0; i < 36; ++i) sv[i] += ((i % 12) >= 6) ? add1 : add0;
        }
#pragma unroll
        for (int i = 0; i < 36; ++i) mx = fmaxf(mx, sv[i]);
        red[kq * 144 + row] = mx;
        __syncthreads();                                      // exchange of the row maxima of both tiles
        {
          mx = fmaxf(fmaxf(red[row], red[144 + row]), fmaxf(red[288 + row], red[432 + row]));
          const float mb = mx * 1.4426950408889634f;
          // The normaliser is the sum of the ROUNDED weights (what the P V MMAs multiply): numerator and denominator then see
          // the same fp16 values and their rounding errors act on (v - o) only (tools/precision_study.py; split_common.cuh)
          float sum = 0.0f;
#pragma unroll
          for (int i = 0; i < 36; ++i) sv[i] = umma::ex2_approx(fmaf(sv[i], 1.4426950408889634f, -mb));
          uint8_t* prow = smem + SM_P + row * 16;
#pragma unroll
          for (int i = 0; i < 36; i += 4) {
            const int key = kq * 36 + i;
            const uint32_t p01 = umma::pack_h2(sv[i], sv[i + 1]), p23 = umma::pack_h2(sv[i + 2], sv[i + 3]);
            *reinterpret_cast<uint2*>(prow + (key >> 3) * LBO_P + (key & 7) * 2) = make_uint2(p01, p23);
            const float2 r01 = umma::unpack_h2(p01), r23 = umma::unpack_h2(p23);
            sum += (r01.x + r01.y) + (r23.x + r23.y);
          }
          rsum[kq * 144 + row] = sum;
        }
        {
          float tv[16];
          load_tile1(tv);
          const int ph0 = (lane >> 4) * 8;
          float mx1 = max1[ph0 * 16 + r1];
#pragma unroll
          for (int pp = 1; pp < 8; ++pp) mx1 = fmaxf(mx1, max1[(ph0 + pp) * 16 + r1]);
          mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 16));
          const float mb = mx1 * 1.4426950408889634f;
          float sum = 0.0f;
#pragma unroll
          for (int i = 0; i < 9; ++i) tv[i] = umma::ex2_approx(fmaf(tv[i], 1.4426950408889634f, -mb));
          if (lane < 16) {
            uint8_t* prow = smem + SM_P + qrow * 16;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
              const int key = key0 + i;
              const __half ph = __float2half_rn(tv[i]);
              *reinterpret_cast<__half*>(prow + (key >> 3) * LBO_P + (key & 7) * 2) = ph;
              sum += __half2float(ph);                       // sum of the rounded weights (see tile 0)
            }
            sum1[part * 16 + r1] = sum;
          }
        }
      }
      umma::fence_proxy_async();
      umma::fence_before_sync();
      __syncthreads();
      // ---- O_h = P (Vh [+ Vl]) : K = 144 keys (9 k-steps), N = 32; then S(h+1) right behind it (the S columns are free)
      if (issuer) {
        umma::fence_after_sync();
        if (h > 0) {                                           // proj(h-1) completed long ago: its slice buffer takes slice h
          umma::mbar_wait(bar_yp, ph_yp); ph_yp ^= 1;
          if (umma::elect_one()) load_wp(h);
          __syncwarp();
        }
        if (umma::elect_one()) {
          const uint64_t vh = d_vh + (uint64_t)((h * HEAD_IMG) >> 4), vl = d_vl + (uint64_t)((h * HEAD_IMG) >> 4);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int k = 0; k < 9; ++k)
              umma::mma_f16_ss(tm + (mt ? TM_O1 : TM_O0), d_p + (uint64_t)((mt * 128 * 16 + k * 2 * LBO_P) >> 4),
                               vh + (uint64_t)((k * 1024) >> 4), IDESC_PV, k > 0);
            if (SPLIT && !(dbgmask & 4)) {
#pragma unroll
              for (int k = 0; k < 9; ++k)
                umma::mma_f16_ss(tm + (mt ? TM_O1 : TM_O0), d_p + (uint64_t)((mt * 128 * 16 + k * 2 * LBO_P) >> 4),
                                 vl + (uint64_t)((k * 1024) >> 4), IDESC_PV, 1u);
            }
          }
          umma::mma_commit(bar_o);
          if (h < 3) {
            const uint64_t qi = d_qimg + (uint64_t)(((h + 1) * HEAD_IMG) >> 4), ki = d_kimg + (uint64_t)(((h + 1) * HEAD_IMG) >> 4);
            const uint64_t q1 = d_q1 + (uint64_t)((((h + 1) & 1) * 8192) >> 4);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma::mma_f16_ss(tm + (mt ? TM_S1 : TM_S0), (mt ? q1 : qi) + (uint64_t)((k * 256) >> 4), ki + (uint64_t)((k * 256) >> 4),
                                 IDESC_S, k > 0);
            umma::mma_commit(bar_s);
          }
        }
        __syncwarp();
      }
      umma::mbar_wait(bar_o, ph_o); ph_o ^= 1;
      umma::fence_after_sync();
      // ---- O epilogue: thread = query row (tile 0: group 0; tile 1: warp 4, lanes < 16) -> O_h hi/lo over the Q_h / K_h
      //      images (S(h) has completed).  Warp 8 meanwhile copies queries 128..143 of head h+2 into the Q1 slot that S(h)
      //      has released.
      {
        const int mt = grp;
        const int row = mt * 128 + q4 * 32 + lane;
        if (grp == 0 || (grp == 1 && q4 == 0)) {              // warp-uniform
          float v[32];
          umma::tmem_ld32(lane_addr + (mt ? TM_O1 : TM_O0), v);
          if (row < NTOK) {
            float tot;
            if (mt == 0) {
              tot = (rsum[row] + rsum[144 + row]) + (rsum[288 + row] + rsum[432 + row]);
            } else {
              tot = 0.0f;
#pragma unroll
              for (int pp = 0; pp < 16; ++pp) tot += sum1[pp * 16 + row - 128];
            }
            const float inv = 1.0f / tot;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = v[c * 8 + j] * inv;
              const uint32_t off = h * HEAD_IMG + c * LBO_P + row * 16;
              if constexpr (SPLIT) {
                uint4 hi, lo;
                umma::split_h2(o[0], o[1], hi.x, lo.x);
                umma::split_h2(o[2], o[3], hi.y, lo.y);
                umma::split_h2(o[4], o[5], hi.z, lo.z);
                umma::split_h2(o[6], o[7], hi.w, lo.w);
                *reinterpret_cast<uint4*>(smem + SM_Q + off) = hi;
                *reinterpret_cast<uint4*>(smem + SM_K + off) = lo;
              } else {
                *reinterpret_cast<uint4*>(smem + SM_Q + off) =
                    make_uint4(umma::pack_h2(o[0], o[1]), umma::pack_h2(o[2], o[3]), umma::pack_h2(o[4], o[5]), umma::pack_h2(o[6], o[7]));
              }
            }
          }
        } else if (warp == 8 && h + 2 < 4) {
          // Q image of head h+2, token groups 16, 17 (2 x 512 B) -> slot (h & 1), row groups {4c, 4c+1}
          const uint8_t* src = smem + SM_Q + (h + 2) * HEAD_IMG + 16 * 512;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint4 val = *reinterpret_cast<const uint4*>(src + j * 512 + lane * 16);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(smem + SM_Q1 + (h & 1) * 8192 + (4 * c + j) * 512 + lane * 16) = val;
          }
        }
      }
      umma::fence_proxy_async();
      umma::fence_before_sync();
      __syncthreads();
      // ---- Y^T (+)= Wp[:, 32h .. 32h+32] . O_h^T   (K = 32: two k-steps; SPLIT: three products)
      if (issuer) {
        umma::fence_after_sync();
        umma::mbar_wait(bar_wp, ph_wp); ph_wp ^= 1;
        if (umma::elect_one()) {
          const uint64_t oh = d_oh + (uint64_t)((h * HEAD_IMG) >> 4), ol = d_ol + (uint64_t)((h * HEAD_IMG) >> 4);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma::mma_f16_ss(tm + TM_Y, d_wph + (uint64_t)(k * 2 * (LBO_W >> 4)), oh + (uint64_t)(k * 2 * (LBO_P >> 4)), IDESC_T, (h > 0 || k > 0) ? 1u : 0u);
          if (SPLIT && !(dbgmask & 8)) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma::mma_f16_ss(tm + TM_Y, d_wph + (uint64_t)(k * 2 * (LBO_W >> 4)), ol + (uint64_t)(k * 2 * (LBO_P >> 4)), IDESC_T, 1u);
          }
          if (SPLIT && !(dbgmask & 16)) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma::mma_f16_ss(tm + TM_Y, d_wpl + (uint64_t)(k * 2 * (LBO_W >> 4)), oh + (uint64_t)(k * 2 * (LBO_P >> 4)), IDESC_T, 1u);
          }
          umma::mma_commit(h < 3 ? bar_yp : bar_y);
        }
        __syncwarp();
        if (h == 3 && wi + gridDim.x < nwin_total) {              // next window: the V regions are dead (PV(3) has completed)
          if (umma::elect_one()) load_v();
          __syncwarp();
        }
      }
    }
    // ---- x1 = x + Y + bproj : thread = feature, token groups by work group; the shortcut loads are issued BEFORE the
    //      wait for the last projection MMAs
    {
      const int tg0 = grp * 18 / 4, tg1 = (grp + 1) * 18 / 4;
      const int f = q4 * 32 + lane;
      const float bp = s_bp[f];
      float xv[40];
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (tg0 + j < tg1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) xv[j * 8 + i] = Xs[(long long)tokpix[(tg0 + j) * 8 + i] * 128 + f];
        }
      umma::mbar_wait(bar_y, ph_y); ph_y ^= 1;
      umma::fence_after_sync();
      if (issuer && wi + gridDim.x < nwin_total) {             // O images are dead: next window's Wq / Wk
        if (umma::elect_one()) load_qk();
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (tg0 + j < tg1) {
          float v[8];
          umma::tmem_ld8(lane_addr + TM_Y + (tg0 + j) * 8, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) Xs[(long long)tokpix[(tg0 + j) * 8 + i] * 128 + f] = xv[j * 8 + i] + (v[i] + bp);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (warp == 0) umma::tmem_dealloc<512>(tm);
}

cudaError_t launch_swin_attn2(float* X, const float* agT, int nslice, int Te, int shift, const SwinAttn2W& w, bool split,
                              int num_sms, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(swin_attn2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(swin_attn2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_SMEM);
  if (e != cudaSuccess) return e;
  const int nwin = nslice * 4;
  const int grid = nwin < num_sms ? nwin : num_sms;
  if (grid <= 0) return cudaSuccess;
  static int dbgmask = -1;
  if (dbgmask < 0) { const char* e2 = getenv("CATSEG_A2_DBG"); dbgmask = e2 ? atoi(e2) : 0; }
  static int force = -1;
  if (force < 0) { const char* e3 = getenv("CATSEG_A2_FORCE_SPLIT"); force = e3 ? atoi(e3) : 0; }
  if (force == 1) split = true;
  if (force == 2) split = false;
  if (split) swin_attn2_kernel<true><<<grid, A2_THREADS, A2_SMEM, st>>>(X, agT, nwin, Te, shift, w, dbgmask);
  else swin_attn2_kernel<false><<<grid, A2_THREADS, A2_SMEM, st>>>(X, agT, nwin, Te, shift, w, dbgmask);
  return cudaGetLastError();
}

// ag_qk fp32 [B][576][256] (q 128 | k 128 guidance terms per pixel, biases folded in) -> fp32 [B][4 windows][18 token groups][2][256][4 tok]
// in window-token order with the cyclic shift applied (SwinTransformerBlock.forward, model.py:195-205)
__global__ void pack_ag_windows_T_kernel(const float* __restrict__ ag, float* __restrict__ out, int B, int shift) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * 4 * 256 * NTOK) return;
  // output index i = ((((b * 4 + win) * 18 + tg) * 2 + h) * 256 + f) * 4 + t4   (token = 8 tg + 4 h + t4)
  const int t4 = (int)(i & 3);
  long long r = i >> 2;
  const int f = (int)(r & 255); r >>= 8;
  const int h = (int)(r & 1); r >>= 1;
  const int tg = (int)(r % 18); r /= 18;
  const int win = (int)(r & 3);
  const int b = (int)(r >> 2);
  const int tok = tg * 8 + h * 4 + t4;
  const int sy = (win >> 1) * WIN + tok / WIN, sx = (win & 1) * WIN + tok % WIN;
  const int pix = ((sy + shift) % GRID) * GRID + (sx + shift) % GRID;
  out[i] = ag[((long long)b * (GRID * GRID) + pix) * 256 + f];
}
cudaError_t launch_pack_ag_windows_T(const float* ag_qk, float* out, int B, int shift, cudaStream_t st) {
  const long long n = (long long)B * 4 * 256 * NTOK;
  pack_ag_windows_T_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ag_qk, out, B, shift);
  return cudaGetLastError();
}

// proj slice of head h: rows = 128 output features, K = the 32 input features of head h (4 chunks x 2048 B), hi and lo
__global__ void pack_proj_slices_kernel(__half* __restrict__ dst, const float* __restrict__ Wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 128) return;
  const int r = i >> 7, k = i & 127, h = k >> 5, kk = k & 31;
  const float v = Wp[(long long)r * 128 + k];
  const __half hi = __float2half_rn(v);
  __half* base = dst + (size_t)h * 8192;                       // 16384 B per head
  const int o = (kk >> 3) * (128 * 8) + r * 8 + (kk & 7);
  base[o] = hi;
  base[4096 + o] = __float2half_rn(v - __half2float(hi));
}
cudaError_t pack_swin_attn2(__half* dst, const float* Wq, const float* Wk, const float* Wv, const float* Wp, int ldqk,
                            cudaStream_t st) {
  const size_t kImg = 128 * 128;
  cudaError_t e;
  if ((e = launch_pack_wimg_split(dst, nullptr, Wq, ldqk, 0, 0, st)) != cudaSuccess) return e;
  if ((e = launch_pack_wimg_split(dst + kImg, nullptr, Wk, ldqk, 0, 0, st)) != cudaSuccess) return e;
  if ((e = launch_pack_wimg_split(dst + 2 * kImg, dst + 3 * kImg, Wv, 128, 0, 0, st)) != cudaSuccess) return e;
  pack_proj_slices_kernel<<<64, 256, 0, st>>>(dst + 4 * kImg, Wp);
  return cudaGetLastError();
}

}  // namespace catseg
