// FAST class aggregation (linear attention over the class axis at every pixel) on tcgen05.
//
// Reference: ClassTransformerLayer.forward (model.py:387-424), AttentionLayer.forward (:338-354),
// LinearAttention.forward (:266-286).  Same split as the exact path (class_exact.cu):
//   class_state_fast : per (image, pixel)   KV[h] = sum_t phi(k_t)^T (v_t/S),  Ksum = sum_t phi(k_t)
//   class_apply_fast : per (image, pixel, 128-class tile)
//        q -> phi(q) [KV | Ksum] -> x1 = x + att -> LN2 -> MLP (ReLU) -> out
// Everything that is a contraction runs on the tensor cores:
//   * q/k projections use K = 256 = [LN1(x) | text guidance]: the guidance half of the A operand is a
//     ready-made fp16 canonical image per (image, 128-class tile) (built once per forward, fetched by
//     one bulk copy), so the "concat" of the reference costs no thread work (biases in the epilogue);
//   * KV = K^T [V | 1] contracts over the TOKEN axis: the [token][feature] images written by the
//     epilogue are consumed as MN-major A and B operands (the ones column yields Ksum);
//   * the state is applied as phi(q) . Bstate with Bstate = [blockdiag(KV_h) | Ksum_h columns] (N = 144).
// Padding classes (T < pad_len) enter as the constant pad_state.
#include <cstdio>
#include <cstdlib>

#include "fast_common.cuh"
#include "internal.h"

namespace catseg {

using namespace fast;

namespace {
constexpr uint32_t LBO_V = 128 * 16 + 16;                          // thread-written 128-row tiles (= LBO_T)
constexpr uint32_t IDESC_N144 = umma::make_idesc_f16(128, 144, 0, 0);
constexpr uint32_t IDESC_KV = umma::make_idesc_f16(128, 144, 1, 1);     // K^T [V|1]: both operands MN-major
constexpr uint32_t IDESC_APPLY = umma::make_idesc_f16(128, 144, 0, 1);  // phi(q) (K-major) x Bstate (MN-major)

// ---------------------------------------------------------------- state kernel layout
constexpr uint32_t ST_W = 0;                                       // Wk_x, Wk_g, Wv images: 3 x 32 KiB resident
constexpr uint32_t ST_XN = ST_W + 3 * WIMG_BYTES;                  // LN1(x) tile; later the phi(k) image
constexpr uint32_t ST_G = ST_XN + TILE_BYTES_T;                    // guidance tile (dense, TMA); later the V image
constexpr uint32_t ST_GV_BYTES = 18 * LBO_V;                       // V image has 18 chunks (128 v + ones + pad)
constexpr uint32_t ST_PAR = ST_G + ST_GV_BYTES;                    // ln g,b [256] bk[128] bv[128]
constexpr uint32_t ST_BAR = ST_PAR + 512 * 4;
constexpr uint32_t ST_SMEM = ST_BAR + 6 * 8 + 16;
constexpr uint32_t ST_TM_KV = 0, ST_TM_ACC = 256;                  // k|v raw: 256 cols; KV accumulator: 144 cols

// ---------------------------------------------------------------- apply kernel layout
constexpr uint32_t AP_RING = 0;                                    // 2 x 32 KiB
constexpr uint32_t AP_XN = AP_RING + 2 * WIMG_BYTES;               // LN1(x) -> later LN2(x1)
constexpr uint32_t AP_GH = AP_XN + TILE_BYTES_T;                   // guidance tile (TMA) -> later MLP hidden chunk
constexpr uint32_t AP_Q = AP_GH + TILE_BYTES_T;                    // phi(q) image
constexpr uint32_t AP_BST = AP_Q + TILE_BYTES_T;                   // Bstate [128 k x 144 n] MN-major, 18 n-groups
constexpr uint32_t AP_PAR = AP_BST + 18 * LBO_V;                   // ln1 g,b ln2 g,b [512] bq[128] b1[512] b2[128] red[512]
constexpr uint32_t AP_BAR = AP_PAR + (512 + 128 + 512 + 128 + 1024) * 4;
static_assert(STG_BYTES <= TILE_BYTES_T + 18 * LBO_V, "staging tile must fit over the phi(q) + Bstate buffers");
constexpr uint32_t AP_SMEM = AP_BAR + 10 * 8 + 16;
constexpr uint32_t AP_TM_Q = 0, AP_TM_ND = 128, AP_TM_Y = 272;     // q / H: [0,128)  num|den: [128,272)  Y: [272,400)
static_assert(ST_SMEM <= 232448 && AP_SMEM <= 232448, "shared memory budget");
}  // namespace

// ================================================================================================
// 16 warps: TMEM lane quarter q4 = warp & 3; group grp = warp >> 2: groups 0,1 take the two 64-column halves of k,
// groups 2,3 those of v.
__global__ void __launch_bounds__(512, 1)
class_state_fast_kernel(const float* __restrict__ X, const __half* __restrict__ timg, float* __restrict__ state,
                        int B, int Te, int npix, int S, ClassFastW w, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* s_g = reinterpret_cast<float*>(smem + ST_PAR);
  float* s_be = s_g + 128;
  float* s_bk = s_be + 128;
  float* s_bv = s_bk + 128;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + ST_BAR);
  uint64_t* bar_g = bar_w + 1;
  uint64_t* bar_m1 = bar_w + 2;
  uint64_t* bar_m2 = bar_w + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q4 = warp & 3, grp = warp >> 2;
  const int half = grp >> 1, chalf = grp & 1;
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == 0;   // warp-uniform issue region, one elected lane
  const int row = q4 * 32 + lane;
  const int ntile = (Te + 127) / 128;

  if (tid < 128) { s_g[tid] = w.ln1_g[tid]; s_be[tid] = w.ln1_b[tid]; s_bk[tid] = w.bqk[128 + tid]; s_bv[tid] = w.bv[tid]; }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) umma::mbar_init(&bar_w[i], 1);
    umma::mbar_fence_init();
    umma::mbar_expect_tx(bar_w, 3 * WIMG_BYTES);
    umma::bulk_g2s(smem + ST_W, w.wimg_kv, 3 * WIMG_BYTES, bar_w);
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot, sb = umma::smem_u32(smem);
  const uint32_t lane_addr = tm + ((uint32_t)(q4 * 32) << 16);
  uint32_t ph_g = 0, ph_m1 = 0, ph_m2 = 0;
  bool w_ready = false;
  const uint64_t d_xn = umma::make_smem_desc(sb + ST_XN, LBO_T, 128), d_g = umma::make_smem_desc(sb + ST_G, LBO_W, 128);
  const uint64_t d_w = umma::make_smem_desc(sb + ST_W, LBO_W, 128);
  const uint64_t d_kT = umma::make_smem_desc(sb + ST_XN, 128, LBO_V), d_vT = umma::make_smem_desc(sb + ST_G, 128, LBO_V);
  const float invS = 1.0f / (float)S;

  const long long nitems = (long long)B * npix;
  long long t_last = clock64(), sacc0 = 0, sacc1 = 0, sacc2 = 0, sacc3 = 0, sacc4 = 0, nt_dbg = 0;
#define SPH(i) do { if (dbg != nullptr && blockIdx.x == 0 && tid == 0) { long long _t = clock64(); sacc##i += _t - t_last; t_last = _t; } } while (0)
  for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int b = (int)it / npix, pix = (int)it % npix;         // B * npix fits 31 bits: no 64-bit divisions
    {   // the next item's token rows (Te rows of 512 bytes, one per class) are prefetched into L2 meanwhile
      const long long itn = it + gridDim.x;
      if (itn < nitems) {
        const int bn = (int)itn / npix, pn = (int)itn % npix;
        for (int i = tid; i < Te * 4; i += 512)
          umma::prefetch_l2(X + (((long long)bn * Te + (i >> 2)) * npix + pn) * 128 + (i & 3) * 32);
      }
    }
    for (int tl = 0; tl < ntile; ++tl) {
      const int t0 = tl * 128;
      const int nvalid = Te - t0 < 128 ? Te - t0 : 128;
      // guidance tile of (b, tl) -> ST_G (the previous V image there is dead: its MMA2 was waited)
      if (issuer) {
        if (umma::elect_one()) {
          umma::mbar_expect_tx(bar_g, WIMG_BYTES);
          umma::bulk_g2s(smem + ST_G, timg + ((long long)b * ntile + tl) * (128 * 128), WIMG_BYTES, bar_g);
        }
        __syncwarp();
      }
      ln_rows_to_tile(X + (((long long)b * Te + t0) * npix + pix) * 128, (long long)npix * 128, nvalid, smem + ST_XN, s_g,
                      s_be, warp, 16, lane);
      umma::fence_proxy_async();
      umma::fence_before_sync();
      __syncthreads();
      SPH(0);
      // ---- k = [xn | g] [Wk_x | Wk_g]^T (cols 0..127), v = xn Wv^T (cols 128..255)
      if (issuer) {
        umma::fence_after_sync();
        if (!w_ready) { umma::mbar_wait(bar_w, 0); w_ready = true; }
        umma::mbar_wait(bar_g, ph_g);
        if (umma::elect_one()) {
          issue_gemm_k128_desc(tm + ST_TM_KV, d_xn, LBO_T, d_w, LBO_W, IDESC_128x128, false);
          issue_gemm_k128_desc(tm + ST_TM_KV, d_g, LBO_W, d_w + (uint64_t)(WIMG_BYTES >> 4), LBO_W, IDESC_128x128, true);
          issue_gemm_k128_desc(tm + ST_TM_KV + 128, d_xn, LBO_T, d_w + (uint64_t)(2 * (WIMG_BYTES >> 4)), LBO_W, IDESC_128x128, false);
          umma::mma_commit(bar_m1);
        }
        __syncwarp();
      }
      ph_g ^= 1;
      umma::mbar_wait(bar_m1, ph_m1); ph_m1 ^= 1;
      umma::fence_after_sync();
      SPH(1);
      // ---- epilogue: groups 0,1 -> phi(k) image (over the LN tile), groups 2,3 -> [v/S | 1] image (over the g tile)
      {
        const bool live = row < nvalid;
        uint8_t* img = smem + (half == 0 ? ST_XN : ST_G);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int cc = chalf * 2 + c2;
          float v[32];
          umma::tmem_ld32(lane_addr + ST_TM_KV + half * 128 + cc * 32, v);
          const float* bb = (half == 0 ? s_bk : s_bv) + cc * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float a = v[i] + bb[i];
            const float ex = __expf(fminf(a, 0.0f));          // unconditional: a per-element branch around the exp costs far more than the MUFU
            a = half == 0 ? (a > 0.0f ? a + 1.0f : ex) : a * invS;
            v[i] = live ? a : 0.0f;
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(img + (cc * 4 + c) * LBO_V + row * 16) =
                make_uint4(umma::pack_h2(v[c * 8], v[c * 8 + 1]), umma::pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                           umma::pack_h2(v[c * 8 + 4], v[c * 8 + 5]), umma::pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
        }
        if (grp == 2) {    // ones column (n = 128) and zero padding (n = 129..143)
          *reinterpret_cast<uint4*>(smem + ST_G + 16 * LBO_V + row * 16) = make_uint4(live ? 0x00003C00u : 0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(smem + ST_G + 17 * LBO_V + row * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      umma::fence_proxy_async();
      umma::fence_before_sync();
      __syncthreads();
      SPH(2);
      // ---- KV (+)= K^T [V | 1]  (contract over the 128 tokens of this tile)
      if (issuer) {
        umma::fence_after_sync();
        if (umma::elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + ST_TM_ACC, d_kT + (uint64_t)(k * 16), d_vT + (uint64_t)(k * 16), IDESC_KV, (tl > 0 || k > 0) ? 1u : 0u);
          umma::mma_commit(bar_m2);
        }
        __syncwarp();
      }
      umma::mbar_wait(bar_m2, ph_m2); ph_m2 ^= 1;
      umma::fence_after_sync();
      SPH(3);
      ++nt_dbg;
    }
    // ---- state[b][pix]: thread = k-feature (h, d) x 8 of the 32 columns KV[h][d][grp*8..+8]; group 0 also Ksum[h*32+d]
    {
      const int h = row >> 5;
      float v[8];
      umma::tmem_ld8(lane_addr + ST_TM_ACC + h * 32 + grp * 8, v);
      float* o = state + it * kStateFloats;
      st4(o + row * 32 + grp * 8, make_float4(v[0], v[1], v[2], v[3]));
      st4(o + row * 32 + grp * 8 + 4, make_float4(v[4], v[5], v[6], v[7]));
      if (grp == 0) {
        float ks[8];
        umma::tmem_ld8(lane_addr + ST_TM_ACC + 128, ks);
        o[4096 + row] = ks[0];
      }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    SPH(4);
  }
#undef SPH
  if (dbg != nullptr && blockIdx.x == 0 && tid == 0) { dbg[0] = sacc0; dbg[1] = sacc1; dbg[2] = sacc2; dbg[3] = sacc3; dbg[4] = sacc4; dbg[5] = nt_dbg; }
  if (warp == 0) umma::tmem_dealloc<512>(tm);
}

// ================================================================================================
// Weight ring of the apply kernel: 10 images per item in consumption order
//   0: Wq_x  1: Wq_g  2: W1_0  3: W2_0  4: W1_1  5: W2_1  6: W1_2  7: W2_2  8: W1_3  9: W2_3
// 16 warps: TMEM lane quarter q4 = warp & 3 (rows), column quarter cq = warp >> 2 (32 columns = one head).
// Residual I/O is coalesced: x is fetched warp-per-row into an fp32 staging tile (over the dead phi(q) /
// Bstate buffers), updated in place by the row threads, and written back warp-per-row once per item.
__global__ void __launch_bounds__(512, 1)
class_apply_fast_kernel(const float* __restrict__ X, float* __restrict__ Xout, const __half* __restrict__ timg,
                        const float* __restrict__ state, const float* __restrict__ pad_state, int B, int Te, int npix,
                        int S, int out_mode, ClassFastW w, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* s_g1 = reinterpret_cast<float*>(smem + AP_PAR);
  float* s_be1 = s_g1 + 128;
  float* s_g2 = s_be1 + 128;
  float* s_be2 = s_g2 + 128;
  float* s_bq = s_be2 + 128;
  float* s_b1 = s_bq + 128;
  float* s_b2 = s_b1 + 512;
  float* s_red = s_b2 + 128;                       // [4 column quarters][128 rows][2]
  float* stage = reinterpret_cast<float*>(smem + AP_Q);   // fp32 [128][STG_LD] over AP_Q + AP_BST
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + AP_BAR);    // [2]
  uint64_t* bar_empty = bar_full + 2;                                  // [2]
  uint64_t* bar_g = bar_full + 4;
  uint64_t* bar_acc = bar_full + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q4 = warp & 3, cq = warp >> 2;
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == 0;   // warp-uniform issue region, one elected lane
  const int row = q4 * 32 + lane;
  const int ntile = (Te + 127) / 128;
  const long long nitems = (long long)B * npix * ntile;
  long long mine = 0;
  for (long long i = blockIdx.x; i < nitems; i += gridDim.x) ++mine;
  const long long total_loads = mine * 10;

  if (tid < 128) {
    s_g1[tid] = w.ln1_g[tid]; s_be1[tid] = w.ln1_b[tid]; s_g2[tid] = w.ln2_g[tid]; s_be2[tid] = w.ln2_b[tid];
    s_bq[tid] = w.bqk[tid]; s_b2[tid] = w.b2[tid];
  }
  s_b1[tid] = w.b1[tid];
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) umma::mbar_init(&bar_full[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot, sb = umma::smem_u32(smem);
  const uint32_t lane_addr = tm + ((uint32_t)(q4 * 32) << 16);

  auto issue_load = [&](long long n) {       // thread 0
    if (n < total_loads) {
      int s = (int)(n & 1);
      umma::mbar_expect_tx(&bar_full[s], WIMG_BYTES);
      umma::bulk_g2s(smem + AP_RING + s * WIMG_BYTES, reinterpret_cast<const uint8_t*>(w.wimg_apply) + (n % 10) * WIMG_BYTES,
                     WIMG_BYTES, &bar_full[s]);
    }
  };
  // thread 0: consume ring image `nimg` with an 8-k-step GEMM, then refill the other slot
  long long nimg = 0;
  const uint64_t d_ring = umma::make_smem_desc(sb + AP_RING, LBO_W, 128);
  const uint64_t d_axn = umma::make_smem_desc(sb + AP_XN, LBO_T, 128), d_agd = umma::make_smem_desc(sb + AP_GH, LBO_W, 128);
  const uint64_t d_agh = umma::make_smem_desc(sb + AP_GH, LBO_V, 128);
  const uint64_t d_aq = umma::make_smem_desc(sb + AP_Q, LBO_V, 128), d_abst = umma::make_smem_desc(sb + AP_BST, 128, LBO_V);
  // called by ALL lanes of warp 0 (uniform state); the issue instructions themselves run on one elected lane
  auto ring_gemm = [&](uint32_t d_tmem, uint64_t a_desc, uint32_t lbo_a, bool acc) {
    const int slot = (int)(nimg & 1);
    umma::mbar_wait(&bar_full[slot], (uint32_t)((nimg >> 1) & 1));
    const uint64_t wd = d_ring + (uint64_t)((uint32_t)slot * (WIMG_BYTES >> 4));
    if (umma::elect_one()) {
      issue_gemm_k128_desc(d_tmem, a_desc, lbo_a, wd, LBO_W, IDESC_128x128, acc);
      umma::mma_commit(&bar_empty[slot]);
    }
    __syncwarp();
    if (nimg > 0 && nimg + 1 < total_loads) {
      umma::mbar_wait(&bar_empty[(nimg - 1) & 1], (uint32_t)(((nimg - 1) >> 1) & 1));
      if (umma::elect_one()) issue_load(nimg + 1);
      __syncwarp();
    }
    ++nimg;
  };
  if (issuer) {
    if (umma::elect_one()) { issue_load(0); issue_load(1); }
    __syncwarp();
  }
  uint32_t ph_g = 0, ph_acc = 0;
  const float fS = (float)S;
  long long t_last = clock64(), pacc0 = 0, pacc1 = 0, pacc2 = 0, pacc3 = 0, pacc4 = 0, pacc5 = 0, pacc6 = 0, pacc7 = 0, nit_dbg = 0;
#define CPH(i) do { if (dbg != nullptr && blockIdx.x == 0 && tid == 0) { long long _t = clock64(); pacc##i += _t - t_last; t_last = _t; } } while (0)

  for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int it32 = (int)it;                                // B * npix * ntile fits 31 bits: no 64-bit divisions
    const int tl = it32 % ntile;
    const int bp = it32 / ntile;
    const int b = bp / npix, pix = bp % npix;
    const int t0 = tl * 128;
    const int nvalid = Te - t0 < 128 ? Te - t0 : 128;
    const long long rstride = (long long)npix * 128;
    const long long row0off = (((long long)b * Te + t0) * npix + pix) * 128;
    {   // the next item's token rows (one 512-byte row per class, 295 KiB apart) and state are prefetched into L2 meanwhile
      const long long itn = it + gridDim.x;
      if (itn < nitems) {
        const int tln = (int)itn % ntile;
        const int bpn = (int)itn / ntile;
        const int bn = bpn / npix, pn = bpn % npix;
        const int nv = Te - tln * 128 < 128 ? Te - tln * 128 : 128;
        if (tid < nv * 4) umma::prefetch_l2(X + (((long long)bn * Te + tln * 128 + (tid >> 2)) * npix + pn) * 128 + (tid & 3) * 32);
        if (tid < 132) umma::prefetch_l2(state + (long long)bpn * kStateFloats + tid * 32);
      }
    }
    if (issuer) {
      if (umma::elect_one()) {
        umma::mbar_expect_tx(bar_g, WIMG_BYTES);
        umma::bulk_g2s(smem + AP_GH, timg + ((long long)b * ntile + tl) * (128 * 128), WIMG_BYTES, bar_g);
      }
      __syncwarp();
    }
    // ---- Bstate [128 k x 144 n] (MN-major, 18 n-groups).  All 512 threads: thread (k, j) = (tid >> 2, tid & 3) loads the
    //      8 values KV[k][8j .. 8j+8) (consecutive threads read consecutive 32 bytes: coalesced; the earlier form, one
    //      128-byte row per thread of four warps, took 4.4 K cycles per item) BEFORE the LayerNorm prologue, which covers
    //      their latency, and afterwards writes chunk gq = 4h + j of row k plus three of the twelve zero chunks.
    const int bk = tid >> 2, bj = tid & 3, bh = bk >> 5;
    float kv[8];
    float ks = 0.0f;
    {
      const float* sp = state + (long long)bp * kStateFloats + bk * 32 + bj * 8;
      const float* pp = pad_state + bk * 32 + bj * 8;
      const float4 a0 = ld4(sp), a1 = ld4(sp + 4), p0 = ld4(pp), p1 = ld4(pp + 4);
      kv[0] = a0.x + p0.x; kv[1] = a0.y + p0.y; kv[2] = a0.z + p0.z; kv[3] = a0.w + p0.w;
      kv[4] = a1.x + p1.x; kv[5] = a1.y + p1.y; kv[6] = a1.z + p1.z; kv[7] = a1.w + p1.w;
      if (bj == 0) ks = state[(long long)bp * kStateFloats + 4096 + bk] + pad_state[4096 + bk];
    }
    ln_rows_to_tile(X + row0off, rstride, nvalid, smem + AP_XN, s_g1, s_be1, warp, 16, lane);
    {
      const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
      uint8_t* brow = smem + AP_BST + bk * 16;
      *reinterpret_cast<uint4*>(brow + (4 * bh + bj) * LBO_V) =
          make_uint4(umma::pack_h2(kv[0], kv[1]), umma::pack_h2(kv[2], kv[3]), umma::pack_h2(kv[4], kv[5]),
                     umma::pack_h2(kv[6], kv[7]));
#pragma unroll
      for (int z = 0; z < 3; ++z) {                           // the 12 chunks of the other three heads are zero
        const int oh = (bh + 1 + z) & 3;
        *reinterpret_cast<uint4*>(brow + (4 * oh + bj) * LBO_V) = zero4;
      }
      if (bj == 0) {
        uint32_t kb = (uint32_t)__half_as_ushort(__float2half_rn(ks));
        uint4 dz = zero4;                                     // n = 128 + h holds Ksum for the rows of head h
        if (bh == 0) dz.x = kb; else if (bh == 1) dz.x = kb << 16; else if (bh == 2) dz.y = kb; else dz.y = kb << 16;
        *reinterpret_cast<uint4*>(brow + 16 * LBO_V) = dz;
      } else if (bj == 1) {
        *reinterpret_cast<uint4*>(brow + 17 * LBO_V) = zero4;
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    CPH(0);
    // ---- q = [xn | g] [Wq_x | Wq_g]^T
    if (issuer) {
      umma::fence_after_sync();
      ring_gemm(tm + AP_TM_Q, d_axn, LBO_T, false);
      umma::mbar_wait(bar_g, ph_g);
      ring_gemm(tm + AP_TM_Q, d_agd, LBO_W, true);
      if (umma::elect_one()) umma::mma_commit(bar_acc);
      __syncwarp();
    }
    ph_g ^= 1;
    umma::mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
    umma::fence_after_sync();
    CPH(1);
    // ---- phi(q) -> Q image (thread = (row, 32-column quarter))
    {
      float v[32];
      umma::tmem_ld32(lane_addr + AP_TM_Q + cq * 32, v);
      const float* bb = s_bq + cq * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) { float a = v[i] + bb[i]; const float ex = __expf(fminf(a, 0.0f)); v[i] = a > 0.0f ? a + 1.0f : ex; }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(smem + AP_Q + (cq * 4 + c) * LBO_V + row * 16) =
            make_uint4(umma::pack_h2(v[c * 8], v[c * 8 + 1]), umma::pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                       umma::pack_h2(v[c * 8 + 4], v[c * 8 + 5]), umma::pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    CPH(2);
    // ---- [num | den] = phi(q) Bstate.  The shortcut rows (warp per row, coalesced) are fetched under these MMAs.
    float4 xv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int r = warp * 8 + i;
      xv[i] = r < nvalid ? ld4(X + row0off + (long long)r * rstride + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (issuer) {
      umma::fence_after_sync();
      if (umma::elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma::mma_f16_ss(tm + AP_TM_ND, d_aq + (uint64_t)(k * 2 * (LBO_V >> 4)), d_abst + (uint64_t)(k * 16), IDESC_APPLY, k > 0);
        umma::mma_commit(bar_acc);
      }
      __syncwarp();
    }
    umma::mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
    umma::fence_after_sync();
    CPH(3);
    // ---- x tile -> staging (phi(q) and Bstate are dead now)
#pragma unroll
    for (int i = 0; i < 8; ++i) st4(stage + (warp * 8 + i) * STG_LD + lane * 4, xv[i]);
    __syncthreads();
    CPH(4);
    // ---- x1 = x + num/(den+eps)*S ; z -> staging ; LN2 statistics
    float x1[32];
    const bool live = row < nvalid;
    {
      float den[8], v[32];
      umma::tmem_ld8(lane_addr + AP_TM_ND + 128, den);
      umma::tmem_ld32(lane_addr + AP_TM_ND + cq * 32, v);
      const float dsel = cq == 0 ? den[0] : (cq == 1 ? den[1] : (cq == 2 ? den[2] : den[3]));
      const float zs = fS / (dsel + 1e-6f);
      float* sp = stage + row * STG_LD + cq * 32;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 x = ld4(sp + i);
        float a0 = fmaf(v[i], zs, x.x), a1 = fmaf(v[i + 1], zs, x.y), a2 = fmaf(v[i + 2], zs, x.z), a3 = fmaf(v[i + 3], zs, x.w);
        x1[i] = a0; x1[i + 1] = a1; x1[i + 2] = a2; x1[i + 3] = a3;
        s1 += (a0 + a1) + (a2 + a3);
        s2 += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        if (out_mode == 0) st4(sp + i, make_float4(x.x + a0, x.y + a1, x.z + a2, x.w + a3));
        else st4(sp + i, make_float4(a0, a1, a2, a3));
      }
      s_red[(cq * 128 + row) * 2] = s1;
      s_red[(cq * 128 + row) * 2 + 1] = s2;
    }
    __syncthreads();
    {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) { s1 += s_red[(c * 128 + row) * 2]; s2 += s_red[(c * 128 + row) * 2 + 1]; }
      float mean = s1 * (1.0f / 128.0f);
      float var = fmaxf(s2 * (1.0f / 128.0f) - mean * mean, 0.0f);
      float rstd = rsqrtf(var + 1e-5f);
      const float* gg = s_g2 + cq * 32;
      const float* bb = s_be2 + cq * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = live ? (x1[c * 8 + i] - mean) * rstd * gg[c * 8 + i] + bb[c * 8 + i] : 0.0f;
        *reinterpret_cast<uint4*>(smem + AP_XN + (cq * 4 + c) * LBO_V + row * 16) =
            make_uint4(umma::pack_h2(y[0], y[1]), umma::pack_h2(y[2], y[3]), umma::pack_h2(y[4], y[5]),
                       umma::pack_h2(y[6], y[7]));
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    CPH(5);
    // ---- MLP 128 -> 512 (ReLU) -> 128, hidden chunks of 128 (H reuses the q columns)
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      if (issuer) {
        umma::fence_after_sync();
        ring_gemm(tm + AP_TM_Q, d_axn, LBO_T, false);
        if (umma::elect_one()) umma::mma_commit(bar_acc);
        __syncwarp();
      }
      umma::mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
      umma::fence_after_sync();
      {
        float v[32];
        umma::tmem_ld32(lane_addr + AP_TM_Q + cq * 32, v);
        const float* bb = s_b1 + j * 128 + cq * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bb[i], 0.0f);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(smem + AP_GH + (cq * 4 + c) * LBO_V + row * 16) =
              make_uint4(umma::pack_h2(v[c * 8], v[c * 8 + 1]), umma::pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                         umma::pack_h2(v[c * 8 + 4], v[c * 8 + 5]), umma::pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
      }
      umma::fence_proxy_async();
      umma::fence_before_sync();
      __syncthreads();
      if (issuer) {
        umma::fence_after_sync();
        ring_gemm(tm + AP_TM_Y, d_agh, LBO_V, j > 0);
        if (umma::elect_one()) umma::mma_commit(bar_acc);
        __syncwarp();
      }
      umma::mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;     // h (and H) may be overwritten by the next chunk
      umma::fence_after_sync();
    }
    CPH(6);
    // ---- out = z + (Y + b2): finish in the staging tile, then one coalesced store per row
    {
      float v[32];
      umma::tmem_ld32(lane_addr + AP_TM_Y + cq * 32, v);
      const float* bb = s_b2 + cq * 32;
      float* sp = stage + row * STG_LD + cq * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 z = ld4(sp + i);
        st4(sp + i, make_float4(z.x + (v[i] + bb[i]), z.y + (v[i + 1] + bb[i + 1]), z.z + (v[i + 2] + bb[i + 2]),
                                z.w + (v[i + 3] + bb[i + 3])));
      }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int r = warp * 8 + i;
      if (r < nvalid) st4(Xout + row0off + (long long)r * rstride + lane * 4, ld4(stage + r * STG_LD + lane * 4));
    }
    __syncthreads();      // the staging tile (phi(q) / Bstate buffers) is rebuilt by the next item
    CPH(7);
    ++nit_dbg;
  }
#undef CPH
  if (dbg != nullptr && blockIdx.x == 0 && tid == 0) {
    dbg[0] = pacc0; dbg[1] = pacc1; dbg[2] = pacc2; dbg[3] = pacc3; dbg[4] = pacc4; dbg[5] = pacc5; dbg[6] = pacc6; dbg[7] = pacc7; dbg[8] = nit_dbg;
  }
  if (warp == 0) umma::tmem_dealloc<512>(tm);
}

// ================================================================================================
cudaError_t launch_class_state_fast(const float* X, const __half* timg, float* state, int B, int Te, int npix,
                                    int S, const ClassFastW& w, int num_sms, cudaStream_t st) {
  {   // per-device function attribute: set on every launch (cheap), a process-wide flag would miss other devices
    cudaError_t e = cudaFuncSetAttribute(class_state_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM);
    if (e != cudaSuccess) return e;
  }
  long long n = (long long)B * npix;
  int grid = (int)(n < num_sms ? n : num_sms);
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("CATSEG_PHASE_TIMING");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on) { cudaMalloc(&dbg, 8 * sizeof(long long)); cudaMemset(dbg, 0, 8 * sizeof(long long)); }
  }
  class_state_fast_kernel<<<grid, 512, ST_SMEM, st>>>(X, timg, state, B, Te, npix, S, w, dbg_on ? dbg : nullptr);
  if (dbg_on) {
    long long hb[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    double nn = hb[5] > 0 ? (double)hb[5] : 1.0;
    fprintf(stderr, "[class_state phases, cycles per 128-class tile over %lld tiles] guidance TMA + LN1 %.0f | k,v mma %.0f | phi/pack epilogue %.0f "
            "| KV mma %.0f | state store (per tile share) %.0f\n", hb[5], hb[0] / nn, hb[1] / nn, hb[2] / nn, hb[3] / nn, hb[4] / nn);
  }
  return cudaGetLastError();
}

cudaError_t launch_class_apply_fast(const float* X, float* Xout, const __half* timg, const float* state,
                                    const float* pad_state, int B, int Te, int npix, int S, int out_mode,
                                    const ClassFastW& w, int num_sms, cudaStream_t st) {
  {   // per-device function attribute: set on every launch (cheap), a process-wide flag would miss other devices
    cudaError_t e = cudaFuncSetAttribute(class_apply_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AP_SMEM);
    if (e != cudaSuccess) return e;
  }
  long long n = (long long)B * npix * ((Te + 127) / 128);
  int grid = (int)(n < num_sms ? n : num_sms);
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("CATSEG_PHASE_TIMING");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on) { cudaMalloc(&dbg, 16 * sizeof(long long)); cudaMemset(dbg, 0, 16 * sizeof(long long)); }
  }
  class_apply_fast_kernel<<<grid, 512, AP_SMEM, st>>>(X, Xout, timg, state, pad_state, B, Te, npix, S, out_mode, w, dbg_on ? dbg : nullptr);
  if (dbg_on) {
    long long hb[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    double nn = hb[8] > 0 ? (double)hb[8] : 1.0;
    fprintf(stderr, "[class_apply phases, cycles/item(128 classes of one pixel) over %lld items] LN1+Bstate %.0f | q-mma %.0f | phi(q) %.0f | apply-mma %.0f "
            "| x staging %.0f | x1+LN2 %.0f | MLP (4 chunks) %.0f | out %.0f\n", hb[8], hb[0] / nn, hb[1] / nn, hb[2] / nn, hb[3] / nn, hb[4] / nn,
            hb[5] / nn, hb[6] / nn, hb[7] / nn);
  }
  return cudaGetLastError();
}

// text guidance [B][Te][128] fp32 -> per (image, 128-class tile) canonical dense fp16 images (rows >= Te are zero)
__global__ void pack_text_img_kernel(const float* __restrict__ tg, __half* __restrict__ timg, int B, int Te, int ntile) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * ntile * 128 * 128;
  if (i >= total) return;
  int k = (int)(i & 127), r = (int)((i >> 7) & 127);
  long long bt = i >> 14;
  int tl = (int)(bt % ntile), b = (int)(bt / ntile);
  int t = tl * 128 + r;
  float v = t < Te ? tg[((long long)b * Te + t) * 128 + k] : 0.0f;
  timg[bt * (128 * 128) + (k >> 3) * (128 * 8) + r * 8 + (k & 7)] = __float2half_rn(v);
}
cudaError_t launch_pack_text_img(const float* tg, __half* timg, int B, int Te, cudaStream_t st) {
  int ntile = (Te + 127) / 128;
  long long total = (long long)B * ntile * 128 * 128;
  pack_text_img_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(tg, timg, B, Te, ntile);
  return cudaGetLastError();
}

}  // namespace catseg
