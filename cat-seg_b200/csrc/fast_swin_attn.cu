// FAST window-attention half of the Swin block on tcgen05 (one 12x12 window = 144 tokens per pass):
//   x1 = x + proj( softmax( (LN1(x) Wq^T + gq) s (LN1(x) Wk^T + gk)^T + mask ) (LN1(x) Wv^T + bv) )
// Reference: SwinTransformerBlock.forward (model.py:185-222), WindowAttention.forward (:86-114).
//
// Orientation.  Token-wise GEMMs are computed TRANSPOSED (features on the UMMA M axis, the 144
// window tokens on the N axis) so the 144-token window needs no M padding:
//   per head h :  [q_h; k_h; v_h; 0]^T (128 x 144) = Wqkv_h (128 x 128) . LN1(x)^T        8 MMAs N=144
//   the TMEM rows of q_h/k_h/v_h are read by "feature" threads and stored as [token-group][d][8 tok]
//   images, which are simultaneously the MN-major A/B operands of  S = Q K^T  (M = queries: 2 tiles,
//   the second holds only 16 valid rows) and the K-major B operand of  O = P V.
//   softmax: tile 0 with four threads per query row (36 keys each); tile 1 (queries 128..143) is replicated
//   into all four TMEM lane quarters (Q1 image) and shared by all 16 warps (9 keys each).  P goes back to
//   shared memory as the K-major A operand, O_h (TMEM) is normalised and stored into the [token][128] image
//   that is the B operand of the transposed projection  Y^T (128 x 144) = Wp . O^T.
// The guidance half of q/k (class independent, biases included) arrives per (window, head) as one bulk copy of
// a fp16 [144 tok][q 32 | k 32] tile prepared by pack_ag_windows (window order, shift applied).
// Weights stream from L2 through a 3-slot ring of 16 KiB half images (K halves), 10 halves per window.
//
// Software pipeline over the heads (every phase of this kernel is issue/MUFU bound, so the point is to keep
// the 16 warps busy while the tensor pipe works):
//   S(h) and QKV(h+1) are queued together; softmax(h) runs meanwhile;
//   PV(h) is queued; the q/k epilogue of head h+1 runs meanwhile (the v epilogue follows as soon as PV(h), the
//   reader of the v image, has completed); then the O epilogue of head h.
// TMEM columns: [0,144) QKV_h^T / Y^T   [160,304) S tile 0   [304,448) S tile 1   [448,512) O_h tiles.
#include <cstdio>
#include <cstdlib>

#include "fast_common.cuh"
#include "internal.h"

namespace catseg {

using namespace fast;

namespace {
constexpr int SA_THREADS = 512;    // 16 warps: TMEM lane quarter q4 = warp & 3, work group grp = warp >> 2
constexpr int NTOK = 144, WIN = 12, GRID = 24;
constexpr uint32_t LBO_X = NTOK * 16 + 16;                 // 2320: thread-written 144-row K-major tiles
constexpr uint32_t WHALF = WIMG_BYTES / 2;                 // 16 KiB: K chunks 0..7 or 8..15 of a weight image
constexpr uint32_t AG_BYTES = NTOK * 64 * 2;               // fp16 [144][q 32 | k 32]
constexpr uint32_t SM_RING = 0;                            // 3 x 16 KiB
constexpr uint32_t SM_XN = SM_RING + 3 * WHALF;            // LN1(x): [144 tok x 128] K-major   16 chunks
constexpr uint32_t SM_QH = SM_XN + 16 * LBO_X;             // q_h image  [18][32][8] fp16 = 9216 B
constexpr uint32_t SM_KH = SM_QH + 9216;
constexpr uint32_t SM_VH = SM_KH + 9216;
constexpr uint32_t SM_P = SM_VH + 9216;                    // P: [rows x 144 keys] K-major, 18 chunks
constexpr uint32_t SM_O = SM_P + 18 * LBO_X;               // O: [144 tok x 128] K-major, 16 chunks
constexpr uint32_t SM_Q1 = SM_O + 16 * LBO_X;              // queries 128..143 replicated into all four TMEM lane quarters: 16 groups x 512 B
constexpr uint32_t SM_AG = SM_Q1 + 16 * 512;               // guidance tile of the head whose q/k epilogue comes next
constexpr uint32_t SM_MISC = SM_AG + AG_BYTES;
// misc: tokpix[144] (int) | red[2][4][144] (float) | red1[2][16][16] | ln g,b [256] | bv[128] | bproj[128]
constexpr uint32_t SM_BAR = SM_MISC + (144 + 2 * 4 * 144 + 2 * 256 + 256 + 128 + 128) * 4;
constexpr uint32_t SA_SMEM = SM_BAR + 10 * 8 + 16;
constexpr uint32_t TM_QKV = 0, TM_S0 = 160, TM_S1 = 304, TM_O0 = 448, TM_O1 = 480;
constexpr uint32_t IDESC_T = umma::make_idesc_f16(128, 144, 0, 0);     // QKV^T, proj^T
constexpr uint32_t IDESC_S = umma::make_idesc_f16(128, 144, 1, 1);     // S = Q K^T (both MN-major images)
constexpr uint32_t IDESC_PV = umma::make_idesc_f16(128, 32, 0, 0);     // O = P V
static_assert(SA_SMEM <= 232448, "shared memory budget");
}  // namespace

// Optional phase timing (CATSEG_PHASE_TIMING=1): thread 0 of CTA 0 accumulates clock64() deltas per phase.
#define PH(i) do { if (dbg != nullptr && blockIdx.x == 0 && tid == 0) { long long _t = clock64(); pacc##i += _t - t_last; t_last = _t; } } while (0)

__global__ void __launch_bounds__(SA_THREADS, 1)
swin_attn_fast_kernel(float* __restrict__ X, const __half* __restrict__ agw, int nwin_total, int Te, int shift,
                      SwinAttnFastW w, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  int* tokpix = reinterpret_cast<int*>(smem + SM_MISC);
  float* red = reinterpret_cast<float*>(tokpix + 144);        // row max per key quarter [4][144]
  float* rsum = red + 4 * 144;                                 // row sum per key quarter [4][144]
  float* max1 = rsum + 4 * 144;                                // tile 1 (queries 128..143): row max per 9-key part [16][16]
  float* sum1 = max1 + 256;                                    //                            row sum per part     [16][16]
  float* s_g = sum1 + 256;
  float* s_be = s_g + 128;
  float* s_bv = s_be + 128;
  float* s_bp = s_bv + 128;
  const __half* agb = reinterpret_cast<const __half*>(smem + SM_AG);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + SM_BAR);   // [3] weight halves
  uint64_t* bar_ag = bar_full + 3;                                    // guidance tile landed
  uint64_t* bar_a = bar_full + 4;                                     // QKV_h done
  uint64_t* bar_s = bar_full + 5;                                     // S done
  uint64_t* bar_o = bar_full + 6;                                     // O_h done
  uint64_t* bar_y = bar_full + 7;                                     // proj done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q4 = warp & 3, grp = warp >> 2;
  // Issuing warp = warp 15 (lane quarter 3, group 3): it has no work in the q/k/v and O epilogues, which is when the long
  // MMA batches are queued (the MMA queue is shallow: the issuing thread is blocked while its MMAs execute).
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == 15;   // warp-uniform issue region, one elected lane

  long long mine = 0;
  for (long long i = blockIdx.x; i < nwin_total; i += gridDim.x) ++mine;
  const long long total_halves = mine * 10;

  if (tid < 128) { s_g[tid] = w.ln_g[tid]; s_be[tid] = w.ln_b[tid]; s_bv[tid] = w.bv[tid]; s_bp[tid] = w.bproj[tid]; }
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) umma::mbar_init(&bar_full[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  for (int i = tid; i < 16 * 512 / 16; i += SA_THREADS) reinterpret_cast<uint4*>(smem + SM_Q1)[i] = make_uint4(0u, 0u, 0u, 0u);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t sb = umma::smem_u32(smem);
  const uint32_t lane_addr = tm + ((uint32_t)(q4 * 32) << 16);
  // base descriptors (built once; the issuing thread only adds to the address field afterwards)
  const uint64_t d_xn = umma::make_smem_desc(sb + SM_XN, LBO_X, 128), d_o = umma::make_smem_desc(sb + SM_O, LBO_X, 128);
  const uint64_t d_qh = umma::make_smem_desc(sb + SM_QH, 128, 512), d_kh = umma::make_smem_desc(sb + SM_KH, 128, 512);
  const uint64_t d_p = umma::make_smem_desc(sb + SM_P, LBO_X, 128), d_vh = umma::make_smem_desc(sb + SM_VH, 512, 128);
  const uint64_t d_ring = umma::make_smem_desc(sb + SM_RING, LBO_W, 128);
  const uint64_t d_q1 = umma::make_smem_desc(sb + SM_Q1, 128, 512);

  // ---- weight ring: half n (10 per window: 2h, 2h+1 = head h; 8, 9 = proj) lives in slot n % 3
  long long nissued = 0, nconsumed = 0;          // warp-uniform counters (only the issuing warp uses them)
  auto refill = [&]() {                          // all lanes of the issuing warp
    while (nissued < nconsumed + 3) {
      if (nissued < total_halves && umma::elect_one()) {
        const int s = (int)(nissued % 3);
        umma::mbar_expect_tx(&bar_full[s], WHALF);
        umma::bulk_g2s(smem + SM_RING + s * WHALF, reinterpret_cast<const uint8_t*>(w.wimg) + (nissued % 10) * WHALF, WHALF,
                       &bar_full[s]);
      }
      __syncwarp();
      ++nissued;
    }
  };
  // [128 x 144]^T GEMM with the two weight halves n0, n0+1 as A and a 144-row K-major tile as B (all lanes of warp 0)
  auto issue_wgemm = [&](long long n0, uint64_t b_desc, uint64_t* bar) {
    umma::mbar_wait(&bar_full[n0 % 3], (uint32_t)((n0 / 3) & 1));
    umma::mbar_wait(&bar_full[(n0 + 1) % 3], (uint32_t)(((n0 + 1) / 3) & 1));
    const uint64_t wa = d_ring + (uint64_t)((uint32_t)(n0 % 3) * (WHALF >> 4));
    const uint64_t wb = d_ring + (uint64_t)((uint32_t)((n0 + 1) % 3) * (WHALF >> 4));
    if (umma::elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma::mma_f16_ss(tm + TM_QKV, wa + (uint64_t)(k * 2 * (LBO_W >> 4)), b_desc + (uint64_t)(k * 2 * (LBO_X >> 4)), IDESC_T, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma::mma_f16_ss(tm + TM_QKV, wb + (uint64_t)(k * 2 * (LBO_W >> 4)), b_desc + (uint64_t)((k + 4) * 2 * (LBO_X >> 4)), IDESC_T, 1u);
      umma::mma_commit(bar);
    }
    __syncwarp();
  };
  auto issue_ag = [&](long long wi, int h) {     // one elected lane
    const int slice = (int)(wi >> 2), win = (int)(wi & 3), b = slice / Te;
    umma::mbar_expect_tx(bar_ag, AG_BYTES);
    umma::bulk_g2s(smem + SM_AG, agw + ((long long)(b * 4 + win) * 4 + h) * (NTOK * 64), AG_BYTES, bar_ag);
  };
  if (issuer) {
    refill();
    if ((long long)blockIdx.x < nwin_total && umma::elect_one()) issue_ag(blockIdx.x, 0);
    __syncwarp();
  }

  long long t_last = clock64();
  long long pacc0 = 0, pacc1 = 0, pacc2 = 0, pacc3 = 0, pacc4 = 0, pacc5 = 0, pacc6 = 0, pacc7 = 0, nwin_dbg = 0;
  long long nwin_done = 0;
  uint32_t ph_a = 0, ph_s = 0, ph_o = 0, ph_y = 0, ph_ag = 0;
  const float scale = 0.17677669529663688110f;

  // q/k/v epilogue pieces (thread = feature d of q, k or v; token groups [tg0, tg1) of 8 tokens)
  const int tg0 = grp * 18 / 4, tg1 = (grp + 1) * 18 / 4;
  auto qk_epilogue = [&]() {                     // warps with q4 < 2
    const int d = lane;
    uint8_t* img = smem + (q4 == 0 ? SM_QH : SM_KH);
    const __half* agp = agb + (q4 == 1 ? 32 : 0) + d;
#pragma unroll 1
    for (int tg = tg0; tg < tg1; ++tg) {
      float v[8];
      umma::tmem_ld8(lane_addr + TM_QKV + tg * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += __half2float(agp[(tg * 8 + i) * 64]);
      if (q4 == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= scale;
      }
      const uint4 pk = make_uint4(umma::pack_h2(v[0], v[1]), umma::pack_h2(v[2], v[3]), umma::pack_h2(v[4], v[5]),
                                  umma::pack_h2(v[6], v[7]));
      *reinterpret_cast<uint4*>(img + tg * 512 + d * 16) = pk;
      if (q4 == 0 && tg >= 16) {                 // queries 128..143 also go to row groups {4c, 4c+1} of the Q1 image
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(smem + SM_Q1 + (4 * c + tg - 16) * 512 + d * 16) = pk;
      }
    }
  };
  auto v_epilogue = [&](int h) {                 // warps with q4 == 2: v_h + bias -> v image (only once its reader PV is done)
    const float bvv = s_bv[h * 32 + lane];
#pragma unroll 1
    for (int tg = tg0; tg < tg1; ++tg) {
      float v[8];
      umma::tmem_ld8(lane_addr + TM_QKV + tg * 8, v);
      *reinterpret_cast<uint4*>(smem + SM_VH + tg * 512 + lane * 16) =
          make_uint4(umma::pack_h2(v[0] + bvv, v[1] + bvv), umma::pack_h2(v[2] + bvv, v[3] + bvv),
                     umma::pack_h2(v[4] + bvv, v[5] + bvv), umma::pack_h2(v[6] + bvv, v[7] + bvv));
    }
  };

  for (long long wi = blockIdx.x; wi < nwin_total; wi += gridDim.x, ++nwin_done) {
    const int slice = (int)(wi >> 2), win = (int)(wi & 3);
    const int wy = win >> 1, wx = win & 1;
    float* Xs = X + (long long)slice * (GRID * GRID) * 128;
    const long long hbase = nwin_done * 10;
    if (tid < NTOK) {
      int sy = wy * WIN + tid / WIN, sx = wx * WIN + tid % WIN;
      tokpix[tid] = ((sy + shift) % GRID) * GRID + (sx + shift) % GRID;
    }
    __syncthreads();
    // ---- LN1 -> XN (warp per row, 9 rows in flight); the NEXT window's rows are prefetched into L2 meanwhile
    {
      const long long wn = wi + gridDim.x;
      if (wn < nwin_total) {
        const int nsl = (int)(wn >> 2), nwy = (int)((wn >> 1) & 1), nwx = (int)(wn & 1);
        for (int i = tid; i < NTOK * 4; i += SA_THREADS) {          // 576 lines of 128 bytes
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
        float4 y = warp_layernorm128_fast(x[i], g, be);
        uint2 pk = make_uint2(umma::pack_h2(y.x, y.y), umma::pack_h2(y.z, y.w));
        *reinterpret_cast<uint2*>(smem + SM_XN + (lane >> 1) * LBO_X + (r0 + i) * 16 + (lane & 1) * 8) = pk;
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    PH(0);
    // ---- prologue: head 0 projections and images
    if (issuer) {
      umma::fence_after_sync();
      issue_wgemm(hbase, d_xn, bar_a);
    }
    umma::mbar_wait(bar_a, ph_a); ph_a ^= 1;
    umma::fence_after_sync();
    if (issuer) { nconsumed += 2; refill(); }
    umma::mbar_wait(bar_ag, ph_ag); ph_ag ^= 1;
    if (q4 < 2) qk_epilogue();
    if (q4 == 2) v_epilogue(0);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    PH(1);

    for (int h = 0; h < 4; ++h) {
      // ---- queue S(h) = Q K^T (two query tiles, K = 32) and, right behind it, the projections of head h+1
      if (issuer) {
        umma::fence_after_sync();
        if (umma::elect_one()) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma::mma_f16_ss(tm + (mt ? TM_S1 : TM_S0), (mt ? d_q1 : d_qh) + (uint64_t)((k * 256) >> 4),
                                d_kh + (uint64_t)((k * 256) >> 4), IDESC_S, k > 0);
          umma::mma_commit(bar_s);
        }
        __syncwarp();
        if (h < 3) {
          issue_wgemm(hbase + 2 * (h + 1), d_xn, bar_a);
          if (umma::elect_one()) issue_ag(wi, h + 1);
        } else {
          if (wi + gridDim.x < nwin_total && umma::elect_one()) issue_ag(wi + gridDim.x, 0);
        }
        __syncwarp();
      }
      umma::mbar_wait(bar_s, ph_s); ph_s ^= 1;
      umma::fence_after_sync();
      PH(2);
      // ---- softmax.  Tile 0 (queries 0..127): FOUR threads per query row (key quarter kq = grp: keys [36 kq, +36) =
      //      3 window rows).  The shifted-window mask needs no lookup there: within a key quarter the vertical band is
      //      constant (ly < 6 <=> kq < 2) and the horizontal band of key j is the compile-time pattern (j % 12 >= 6).
      //      Tile 1 (queries 128..143) exists in all four TMEM lane quarters (Q1 image), so all 16 warps share it:
      //      warp (q4, grp) takes the 9 keys of part grp*4 + q4 -- a quarter-0-only pass would keep one scheduler (and
      //      its MUFU) busy for as long as the whole of tile 0.
      {
        // tile 1, pass 1: local max of this warp's 9 keys (the values are re-read from TMEM after the barrier: cheaper
        // than keeping them live next to the 36 values of tile 0)
        const int part = grp * 4 + q4, key0 = part * 9, r1 = lane & 15;     // lanes 16..31 mirror 0..15 (results unused)
        const int qrow = 128 + r1, qly = qrow / WIN, qlx = qrow % WIN;
        auto load_tile1 = [&](float* tv) {
          umma::tmem_ld16(lane_addr + TM_S1 + key0, tv);                    // 9 keys used; the rest are other parts' columns
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
        // tile 0, pass 1
        const int row = q4 * 32 + lane;
        const int kq = grp;
        const uint32_t s_addr = lane_addr + TM_S0 + kq * 36;
        float add0 = 0.0f, add1 = 0.0f;                       // additive mask for keys with lx < 6 / lx >= 6
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
        if (shift > 0) {                                      // (uniform) only shifted windows carry the additive mask
#pragma unroll
          for (int i = 0; i < 36; ++i) sv[i] += ((i % 12) >= 6) ? add1 : add0;
        }
#pragma unroll
        for (int i = 0; i < 36; ++i) mx = fmaxf(mx, sv[i]);
        red[kq * 144 + row] = mx;
        __syncthreads();                                      // the only exchange barrier: row maxima of both tiles
        {
          mx = fmaxf(fmaxf(red[row], red[144 + row]), fmaxf(red[288 + row], red[432 + row]));
          const float mb = mx * 1.4426950408889634f;
          // (packed FFMA2 / FADD2 for the exponent arguments and the row sum were measured: 8 % slower here)
          float sum = 0.0f;
#pragma unroll
          for (int i = 0; i < 36; ++i) sv[i] = umma::ex2_approx(fmaf(sv[i], 1.4426950408889634f, -mb));
          // 36 keys = 4.5 chunks of 8: quarter kq starts at chunk 4.5 kq -> write 72-byte span as 8-byte pieces.
          // The normaliser is the sum of the ROUNDED fp16 weights the P V MMAs multiply (see swin_attn2.cu)
          uint8_t* prow = smem + SM_P + row * 16;
#pragma unroll
          for (int i = 0; i < 36; i += 4) {
            const int key = kq * 36 + i;                      // multiple of 4
            const uint32_t p01 = umma::pack_h2(sv[i], sv[i + 1]), p23 = umma::pack_h2(sv[i + 2], sv[i + 3]);
            *reinterpret_cast<uint2*>(prow + (key >> 3) * LBO_X + (key & 7) * 2) = make_uint2(p01, p23);
            const float2 r01 = umma::unpack_h2(p01), r23 = umma::unpack_h2(p23);
            sum += (r01.x + r01.y) + (r23.x + r23.y);
          }
          rsum[kq * 144 + row] = sum;
        }
        // tile 1, pass 2
        {
          float tv[16];
          load_tile1(tv);
          // row maximum over the 16 parts: each half of the warp reads 8 of them, one exchange with the mirror lane
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
              *reinterpret_cast<__half*>(prow + (key >> 3) * LBO_X + (key & 7) * 2) = ph;
              sum += __half2float(ph);
            }
            sum1[part * 16 + r1] = sum;
          }
        }
      }
      umma::fence_proxy_async();
      umma::fence_before_sync();
      __syncthreads();
      PH(3);
      // ---- queue O_h = P V : K = 144 keys (9 k-steps), N = 32; meanwhile the q/k/v epilogue of head h+1
      if (issuer) {
        umma::fence_after_sync();
        if (umma::elect_one()) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int k = 0; k < 9; ++k)
              umma::mma_f16_ss(tm + (mt ? TM_O1 : TM_O0), d_p + (uint64_t)((mt * 128 * 16 + k * 2 * LBO_X) >> 4),
                                d_vh + (uint64_t)((k * 1024) >> 4), IDESC_PV, k > 0);
          umma::mma_commit(bar_o);
        }
        __syncwarp();
      }
      if (h < 3) {
        umma::mbar_wait(bar_a, ph_a); ph_a ^= 1;              // QKV(h+1), queued behind S(h)
        umma::fence_after_sync();
        if (issuer) { nconsumed += 2; refill(); }
        umma::mbar_wait(bar_ag, ph_ag); ph_ag ^= 1;
        if (q4 < 2) qk_epilogue();                            // the q/k images are free: S(h) has completed
      }
      umma::mbar_wait(bar_o, ph_o); ph_o ^= 1;
      umma::fence_after_sync();
      PH(4);
      if (h < 3 && q4 == 2) v_epilogue(h + 1);                // the v image is free: PV(h) has completed
      // ---- O epilogue: thread = query row (tile 0: warps 0-3, tile 1: warp 4 lanes < 16)
      {
        const int mt = grp;                                   // group 0: tile 0, group 1 (lane quarter 0): tile 1
        const int row = mt * 128 + q4 * 32 + lane;
        if (grp == 0 || (grp == 1 && q4 == 0)) {              // warp-uniform (tcgen05.ld is .sync.aligned)
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
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(smem + SM_O + (h * 4 + c) * LBO_X + row * 16) =
                  make_uint4(umma::pack_h2(v[c * 8 + 0] * inv, v[c * 8 + 1] * inv), umma::pack_h2(v[c * 8 + 2] * inv, v[c * 8 + 3] * inv),
                             umma::pack_h2(v[c * 8 + 4] * inv, v[c * 8 + 5] * inv), umma::pack_h2(v[c * 8 + 6] * inv, v[c * 8 + 7] * inv));
          }
        }
      }
      umma::fence_proxy_async();
      umma::fence_before_sync();
      __syncthreads();
      PH(5);
    }
    // ---- Y^T = Wp . O^T
    if (issuer) {
      umma::fence_after_sync();
      issue_wgemm(hbase + 8, d_o, bar_y);
    }
    // ---- x1 = x + Y + bproj : thread = feature, 36-40 tokens each; a warp touches 128 contiguous bytes per token.
    //      The shortcut loads are issued BEFORE the wait for the projection MMAs (all in flight under them).
    {
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
      if (issuer) { nconsumed += 2; refill(); }
      PH(6);
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (tg0 + j < tg1) {                                         // warp-uniform
          float v[8];
          umma::tmem_ld8(lane_addr + TM_QKV + (tg0 + j) * 8, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) Xs[(long long)tokpix[(tg0 + j) * 8 + i] * 128 + f] = xv[j * 8 + i] + (v[i] + bp);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    PH(7);
    ++nwin_dbg;
  }
  if (dbg != nullptr && blockIdx.x == 0 && tid == 0) {
    dbg[0] = pacc0; dbg[1] = pacc1; dbg[2] = pacc2; dbg[3] = pacc3; dbg[4] = pacc4; dbg[5] = pacc5; dbg[6] = pacc6; dbg[7] = pacc7;
    dbg[15] = nwin_dbg;
  }
  if (warp == 0) umma::tmem_dealloc<512>(tm);
}

cudaError_t launch_swin_attn_fast(float* X, const __half* agw, int nslice, int Te, int shift,
                                  const SwinAttnFastW& w, int num_sms, cudaStream_t st) {
  {   // per-device function attribute: set on every launch (cheap), a process-wide flag would miss other devices
    cudaError_t e = cudaFuncSetAttribute(swin_attn_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_SMEM);
    if (e != cudaSuccess) return e;
  }
  int nwin = nslice * 4;
  int grid = nwin < num_sms ? nwin : num_sms;
  if (grid <= 0) return cudaSuccess;
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("CATSEG_PHASE_TIMING");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on) { cudaMalloc(&dbg, 16 * sizeof(long long)); cudaMemset(dbg, 0, 16 * sizeof(long long)); }
  }
  swin_attn_fast_kernel<<<grid, SA_THREADS, SA_SMEM, st>>>(X, agw, nwin, Te, shift, w, dbg_on ? dbg : nullptr);
  if (dbg_on) {
    long long hbuf[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(hbuf, dbg, sizeof(hbuf), cudaMemcpyDeviceToHost);
    cudaMemset(dbg, 0, 16 * sizeof(long long));
    double n = hbuf[15] > 0 ? (double)hbuf[15] : 1.0;
    fprintf(stderr, "[swin_attn phases, cycles/window over %lld windows] LN %.0f | head-0 qkv+epi %.0f | per head: S-wait %.0f softmax %.0f "
            "qkv-epi(h+1)+PV-wait %.0f O-epi %.0f | proj-mma %.0f Y-epi %.0f\n", hbuf[15], hbuf[0] / n, hbuf[1] / n, hbuf[2] / n / 4, hbuf[3] / n / 4,
            hbuf[4] / n / 4, hbuf[5] / n / 4, hbuf[6] / n, hbuf[7] / n);
  }
  return cudaGetLastError();
}

// ag_qk fp32 [B][576][256] (q 128 | k 128 guidance terms per pixel) -> fp16 tiles [B][4 windows][4 heads][144 tok][q_h 32 | k_h 32]
// in window-token order with the cyclic shift applied (SwinTransformerBlock.forward, model.py:195-205).
__global__ void pack_ag_windows_kernel(const float* __restrict__ ag, __half* __restrict__ out, int B, int shift) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * 16 * NTOK * 64) return;
  const int c = (int)(i & 63);
  long long r = i >> 6;
  const int tok = (int)(r % NTOK); r /= NTOK;
  const int h = (int)(r & 3); r >>= 2;
  const int win = (int)(r & 3);
  const int b = (int)(r >> 2);
  const int sy = (win >> 1) * WIN + tok / WIN, sx = (win & 1) * WIN + tok % WIN;
  const int pix = ((sy + shift) % GRID) * GRID + (sx + shift) % GRID;
  out[i] = __float2half_rn(ag[((long long)b * (GRID * GRID) + pix) * 256 + (c < 32 ? h * 32 + c : 128 + h * 32 + (c - 32))]);
}
cudaError_t launch_pack_ag_windows(const float* ag_qk, __half* out, int B, int shift, cudaStream_t st) {
  const long long n = (long long)B * 16 * NTOK * 64;
  pack_ag_windows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ag_qk, out, B, shift);
  return cudaGetLastError();
}

// Per-head QKV weight image: rows [0,32) = Wq[h*32..] (LN(x) columns), [32,64) = Wk, [64,96) = Wv, [96,128) = 0
__global__ void pack_qkv_head_img_kernel(__half* __restrict__ dst, const float* __restrict__ Wq,
                                         const float* __restrict__ Wk, const float* __restrict__ Wv, int ldqk, int h) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 128) return;
  int r = i >> 7, k = i & 127;
  float v = 0.0f;
  if (r < 32) v = Wq[(long long)(h * 32 + r) * ldqk + k];
  else if (r < 64) v = Wk[(long long)(h * 32 + r - 32) * ldqk + k];
  else if (r < 96) v = Wv[(long long)(h * 32 + r - 64) * 128 + k];
  dst[(k >> 3) * (128 * 8) + r * 8 + (k & 7)] = __float2half_rn(v);
}
cudaError_t launch_pack_qkv_head_img(__half* dst, const float* Wq, const float* Wk, const float* Wv, int ldqk,
                                     int h, cudaStream_t st) {
  pack_qkv_head_img_kernel<<<64, 256, 0, st>>>(dst, Wq, Wk, Wv, ldqk, h);
  return cudaGetLastError();
}

}  // namespace catseg
