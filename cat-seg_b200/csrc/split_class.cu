// PRECISE class aggregation (linear attention over the class axis at every pixel) on tcgen05.
//
// Reference: ClassTransformerLayer.forward (model.py:387-424), AttentionLayer.forward (:338-354),
// LinearAttention.forward (:266-286).  Same stage split as fast_class.cu, with the operand scheme of
// split_common.cuh: the q / k projections and phi(q), phi(k) are attention weights (single fp16 term); the
// value path -- v projection, the state KV = phi(k)^T [v/S | 1], its application phi(q) [KV | Ksum] -- carries
// hi + lo fp16 pairs.  Numerator and normaliser come out of the same MMAs, so both see the same rounded phi(q) / phi(k).
//   class_state_split : per (image, pixel)   KV[h], Ksum   (fp32 state, identical layout to the fast kernel)
//   class_apply_split : per (image, pixel, 128-class tile)   x1 = x + phi(q) [KV | Ksum]  ->  X1
//   the MLP half  X = x + x1 + MLP(LN2(x1))  runs in split_mlp.cu (ReLU) on the X1 buffer.
#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "split_common.cuh"

namespace catseg {

using namespace fast;

namespace {
constexpr uint32_t LBO_V = 128 * 16 + 16;
constexpr uint32_t IDESC_KV144 = umma::make_idesc_f16(128, 144, 1, 1);    // K^T [Vh | 1]: both operands MN-major
constexpr uint32_t IDESC_KV128 = umma::make_idesc_f16(128, 128, 1, 1);    // K^T Vl
constexpr uint32_t IDESC_APPLY = umma::make_idesc_f16(128, 144, 0, 1);    // phi(q) (K-major) x Bstate (MN-major)

// ---------------------------------------------------------------- state kernel layout
constexpr uint32_t SS_RING = 0;                                    // 3 x 32 KiB: Wk_x, Wk_g, Wv hi, Wv lo per tile
constexpr uint32_t SS_XH = SS_RING + 3 * WIMG_BYTES;               // LN1(x) hi -> later the phi(k) image
constexpr uint32_t SS_XL = SS_XH + TILE_BYTES_T;                   // LN1(x) lo -> later the V lo image (16 chunks)
constexpr uint32_t SS_G = SS_XL + TILE_BYTES_T;                    // guidance tile (dense, TMA) -> later the [V hi | 1] image
constexpr uint32_t SS_PAR = SS_G + 18 * LBO_V;                     // ln g,b [256] bk[128] bv[128]
constexpr uint32_t SS_BAR = SS_PAR + 512 * 4;                      // full[3] empty[3] g m1 m2 + tmem ptr
constexpr uint32_t SS_SMEM = SS_BAR + 10 * 8 + 16;
constexpr uint32_t SS_TM_KV = 0, SS_TM_ACC = 256;

// ---------------------------------------------------------------- apply kernel layout
constexpr uint32_t SA_W = 0;                                       // Wq_x, Wq_g resident (2 x 32 KiB)
constexpr uint32_t SA_XN = SA_W + 2 * WIMG_BYTES;                  // LN1(x) (single) -> later the phi(q) image
constexpr uint32_t SA_G = SA_XN + TILE_BYTES_T;                    // guidance tile (TMA)
constexpr uint32_t SA_BH = SA_G + WIMG_BYTES;                      // Bstate hi [128 k x 144 n] MN-major, 18 n-groups
constexpr uint32_t SA_BL = SA_BH + 18 * LBO_V;                     // Bstate lo
constexpr uint32_t SA_PAR = SA_BL + 18 * LBO_V;                    // ln1 g,b [256] bq[128]
constexpr uint32_t SA_BAR = SA_PAR + 384 * 4;
constexpr uint32_t SA_SMEM = SA_BAR + 4 * 8 + 16;
static_assert(STG_BYTES <= 2 * 18 * LBO_V, "the staging tile must fit over the two Bstate images");
constexpr uint32_t SA_TM_Q = 0, SA_TM_ND = 128;
static_assert(SS_SMEM <= 232448 && SA_SMEM <= 232448, "shared memory budget");
}  // namespace

// ================================================================================================
// 16 warps: TMEM lane quarter q4 = warp & 3; group grp = warp >> 2: groups 0,1 take the two 64-column halves of k,
// groups 2,3 those of v.  Warp 0 also issues.
__global__ void __launch_bounds__(512, 1)
class_state_split_kernel(const float* __restrict__ X, const __half* __restrict__ timg, float* __restrict__ state,
                         int B, int Te, int npix, int S, ClassSplitW w) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* s_g = reinterpret_cast<float*>(smem + SS_PAR);
  float* s_be = s_g + 128;
  float* s_bk = s_be + 128;
  float* s_bv = s_bk + 128;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + SS_BAR);
  uint64_t* bar_empty = bar_full + 3;
  uint64_t* bar_g = bar_full + 6;
  uint64_t* bar_m1 = bar_full + 7;
  uint64_t* bar_m2 = bar_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 9);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q4 = warp & 3, grp = warp >> 2;
  const int half = grp >> 1, chalf = grp & 1;
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == 0;
  const int row = q4 * 32 + lane;
  const int ntile = (Te + 127) / 128;

  if (tid < 128) { s_g[tid] = w.ln1_g[tid]; s_be[tid] = w.ln1_b[tid]; s_bk[tid] = w.bqk[128 + tid]; s_bv[tid] = w.bv[tid]; }
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) umma::mbar_init(&bar_full[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot, sb = umma::smem_u32(smem);
  const uint32_t lane_addr = tm + ((uint32_t)(q4 * 32) << 16);
  uint32_t ph_g = 0, ph_m1 = 0, ph_m2 = 0;
  const uint64_t d_xh = umma::make_smem_desc(sb + SS_XH, LBO_T, 128), d_xl = umma::make_smem_desc(sb + SS_XL, LBO_T, 128);
  const uint64_t d_g = umma::make_smem_desc(sb + SS_G, LBO_W, 128);
  const uint64_t d_w0 = umma::make_smem_desc(sb + SS_RING, LBO_W, 128);
  const uint64_t d_kT = umma::make_smem_desc(sb + SS_XH, 128, LBO_V), d_vhT = umma::make_smem_desc(sb + SS_G, 128, LBO_V);
  const uint64_t d_vlT = umma::make_smem_desc(sb + SS_XL, 128, LBO_V);
  const float invS = 1.0f / (float)S;

  const long long nitems = (long long)B * npix;
  long long mine = 0;
  for (long long i = blockIdx.x; i < nitems; i += gridDim.x) ++mine;
  split::WeightRing<3, WIMG_BYTES> ring;
  ring.init(smem + SS_RING, bar_full, bar_empty, w.wimg_kv, 4, mine * ntile * 4);
  if (issuer) ring.prime();

  for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int b = (int)it / npix, pix = (int)it % npix;
    {   // the next item's token rows are prefetched into L2 meanwhile
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
      if (issuer) {
        if (umma::elect_one()) {
          umma::mbar_expect_tx(bar_g, WIMG_BYTES);
          umma::bulk_g2s(smem + SS_G, timg + ((long long)b * ntile + tl) * (128 * 128), WIMG_BYTES, bar_g);
        }
        __syncwarp();
      }
      split::ln_rows_to_tile_split(X + (((long long)b * Te + t0) * npix + pix) * 128, (long long)npix * 128, nvalid,
                                   smem + SS_XH, smem + SS_XL, s_g, s_be, warp, 16, lane);
      umma::fence_proxy_async();
      umma::fence_before_sync();
      __syncthreads();
      // ---- k = [xh | g] [Wk_x | Wk_g]^T (single terms), v = (xh + xl) (Wvh + Wvl)^T (3 products)
      if (issuer) {
        umma::fence_after_sync();
        uint32_t off = ring.acquire();
        if (umma::elect_one()) issue_gemm_k128_desc(tm + SS_TM_KV, d_xh, LBO_T, d_w0 + (uint64_t)(off >> 4), LBO_W, IDESC_128x128, false);
        __syncwarp();
        ring.release();
        umma::mbar_wait(bar_g, ph_g);
        off = ring.acquire();
        if (umma::elect_one()) issue_gemm_k128_desc(tm + SS_TM_KV, d_g, LBO_W, d_w0 + (uint64_t)(off >> 4), LBO_W, IDESC_128x128, true);
        __syncwarp();
        ring.release();
        off = ring.acquire();
        if (umma::elect_one()) {
          issue_gemm_k128_desc(tm + SS_TM_KV + 128, d_xh, LBO_T, d_w0 + (uint64_t)(off >> 4), LBO_W, IDESC_128x128, false);
          issue_gemm_k128_desc(tm + SS_TM_KV + 128, d_xl, LBO_T, d_w0 + (uint64_t)(off >> 4), LBO_W, IDESC_128x128, true);
        }
        __syncwarp();
        ring.release();
        off = ring.acquire();
        if (umma::elect_one()) {
          issue_gemm_k128_desc(tm + SS_TM_KV + 128, d_xh, LBO_T, d_w0 + (uint64_t)(off >> 4), LBO_W, IDESC_128x128, true);
          umma::mma_commit(bar_m1);
        }
        __syncwarp();
        ring.release();
      }
      ph_g ^= 1;
      umma::mbar_wait(bar_m1, ph_m1); ph_m1 ^= 1;
      umma::fence_after_sync();
      // ---- epilogue: groups 0,1 -> phi(k) image (over LN hi), groups 2,3 -> [v/S hi | 1] (over the g tile) and v/S lo (over LN lo)
      {
        const bool live = row < nvalid;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int cc = chalf * 2 + c2;
          float v[32];
          umma::tmem_ld32(lane_addr + SS_TM_KV + half * 128 + cc * 32, v);
          const float* bb = (half == 0 ? s_bk : s_bv) + cc * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float a = v[i] + bb[i];
            const float ex = __expf(fminf(a, 0.0f));
            a = half == 0 ? (a > 0.0f ? a + 1.0f : ex) : a * invS;
            v[i] = live ? a : 0.0f;
          }
          if (half == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(smem + SS_XH + (cc * 4 + c) * LBO_V + row * 16) =
                  make_uint4(umma::pack_h2(v[c * 8], v[c * 8 + 1]), umma::pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                             umma::pack_h2(v[c * 8 + 4], v[c * 8 + 5]), umma::pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 hi, lo;
              umma::split_h2(v[c * 8], v[c * 8 + 1], hi.x, lo.x);
              umma::split_h2(v[c * 8 + 2], v[c * 8 + 3], hi.y, lo.y);
              umma::split_h2(v[c * 8 + 4], v[c * 8 + 5], hi.z, lo.z);
              umma::split_h2(v[c * 8 + 6], v[c * 8 + 7], hi.w, lo.w);
              *reinterpret_cast<uint4*>(smem + SS_G + (cc * 4 + c) * LBO_V + row * 16) = hi;
              *reinterpret_cast<uint4*>(smem + SS_XL + (cc * 4 + c) * LBO_V + row * 16) = lo;
            }
          }
        }
        if (grp == 2) {    // ones column (n = 128) and zero padding (n = 129..143) of the hi image
          *reinterpret_cast<uint4*>(smem + SS_G + 16 * LBO_V + row * 16) = make_uint4(live ? 0x00003C00u : 0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(smem + SS_G + 17 * LBO_V + row * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      umma::fence_proxy_async();
      umma::fence_before_sync();
      __syncthreads();
      // ---- KV (+)= K^T [Vh | 1] + K^T Vl   (contract over the 128 tokens of this tile)
      if (issuer) {
        umma::fence_after_sync();
        if (umma::elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + SS_TM_ACC, d_kT + (uint64_t)(k * 16), d_vhT + (uint64_t)(k * 16), IDESC_KV144, (tl > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + SS_TM_ACC, d_kT + (uint64_t)(k * 16), d_vlT + (uint64_t)(k * 16), IDESC_KV128, 1u);
          umma::mma_commit(bar_m2);
        }
        __syncwarp();
      }
      umma::mbar_wait(bar_m2, ph_m2); ph_m2 ^= 1;
      umma::fence_after_sync();
    }
    // ---- state[b][pix]
    {
      const int h = row >> 5;
      float v[8];
      umma::tmem_ld8(lane_addr + SS_TM_ACC + h * 32 + grp * 8, v);
      float* o = state + it * kStateFloats;
      st4(o + row * 32 + grp * 8, make_float4(v[0], v[1], v[2], v[3]));
      st4(o + row * 32 + grp * 8 + 4, make_float4(v[4], v[5], v[6], v[7]));
      if (grp == 0) {
        float ks[8];
        umma::tmem_ld8(lane_addr + SS_TM_ACC + 128, ks);
        o[4096 + row] = ks[0];
      }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (warp == 0) umma::tmem_dealloc<512>(tm);
}

// ================================================================================================
// 16 warps: TMEM lane quarter q4 = warp & 3 (rows), column quarter cq = warp >> 2 (32 columns = one head).
__global__ void __launch_bounds__(512, 1)
class_apply_split_kernel(const float* __restrict__ X, float* __restrict__ X1, const __half* __restrict__ timg,
                         const float* __restrict__ state, const float* __restrict__ pad_state, int B, int Te, int npix,
                         int S, ClassSplitW w) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* s_g1 = reinterpret_cast<float*>(smem + SA_PAR);
  float* s_be1 = s_g1 + 128;
  float* s_bq = s_be1 + 128;
  float* stage = reinterpret_cast<float*>(smem + SA_BH);      // fp32 [128][STG_LD] over the two Bstate images
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + SA_BAR);
  uint64_t* bar_g = bar_w + 1;
  uint64_t* bar_acc = bar_w + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q4 = warp & 3, cq = warp >> 2;
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == 0;
  const int row = q4 * 32 + lane;
  const int ntile = (Te + 127) / 128;
  const long long nitems = (long long)B * npix * ntile;

  if (tid < 128) { s_g1[tid] = w.ln1_g[tid]; s_be1[tid] = w.ln1_b[tid]; s_bq[tid] = w.bqk[tid]; }
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) umma::mbar_init(&bar_w[i], 1);
    umma::mbar_fence_init();
    umma::mbar_expect_tx(bar_w, 2 * WIMG_BYTES);
    umma::bulk_g2s(smem + SA_W, w.wimg_q, 2 * WIMG_BYTES, bar_w);
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot, sb = umma::smem_u32(smem);
  const uint32_t lane_addr = tm + ((uint32_t)(q4 * 32) << 16);
  const uint64_t d_w = umma::make_smem_desc(sb + SA_W, LBO_W, 128);
  const uint64_t d_xn = umma::make_smem_desc(sb + SA_XN, LBO_T, 128), d_g = umma::make_smem_desc(sb + SA_G, LBO_W, 128);
  const uint64_t d_q = umma::make_smem_desc(sb + SA_XN, LBO_V, 128);
  const uint64_t d_bh = umma::make_smem_desc(sb + SA_BH, 128, LBO_V), d_bl = umma::make_smem_desc(sb + SA_BL, 128, LBO_V);
  uint32_t ph_g = 0, ph_acc = 0;
  bool w_ready = false;
  const float fS = (float)S;

  for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int it32 = (int)it;
    const int tl = it32 % ntile;
    const int bp = it32 / ntile;
    const int b = bp / npix, pix = bp % npix;
    const int t0 = tl * 128;
    const int nvalid = Te - t0 < 128 ? Te - t0 : 128;
    const long long rstride = (long long)npix * 128;
    const long long row0off = (((long long)b * Te + t0) * npix + pix) * 128;
    {
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
        umma::bulk_g2s(smem + SA_G, timg + ((long long)b * ntile + tl) * (128 * 128), WIMG_BYTES, bar_g);
      }
      __syncwarp();
    }
    // ---- Bstate hi / lo [128 k x 144 n] (MN-major): thread (k, j) = (tid >> 2, tid & 3) owns KV[k][8j .. 8j+8)
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
    ln_rows_to_tile(X + row0off, rstride, nvalid, smem + SA_XN, s_g1, s_be1, warp, 16, lane);
    {
      const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
      uint8_t* bhi = smem + SA_BH + bk * 16;
      uint8_t* blo = smem + SA_BL + bk * 16;
      uint4 hi, lo;
      umma::split_h2(kv[0], kv[1], hi.x, lo.x);
      umma::split_h2(kv[2], kv[3], hi.y, lo.y);
      umma::split_h2(kv[4], kv[5], hi.z, lo.z);
      umma::split_h2(kv[6], kv[7], hi.w, lo.w);
      *reinterpret_cast<uint4*>(bhi + (4 * bh + bj) * LBO_V) = hi;
      *reinterpret_cast<uint4*>(blo + (4 * bh + bj) * LBO_V) = lo;
#pragma unroll
      for (int z = 0; z < 3; ++z) {                           // the 12 chunks of the other three heads are zero
        const int oh = (bh + 1 + z) & 3;
        *reinterpret_cast<uint4*>(bhi + (4 * oh + bj) * LBO_V) = zero4;
        *reinterpret_cast<uint4*>(blo + (4 * oh + bj) * LBO_V) = zero4;
      }
      if (bj == 0) {
        const __half kh = __float2half_rn(ks);
        const uint32_t kbh = (uint32_t)__half_as_ushort(kh);
        const uint32_t kbl = (uint32_t)__half_as_ushort(__float2half_rn(ks - __half2float(kh)));
        uint4 dh = zero4, dl = zero4;                         // n = 128 + h holds Ksum for the rows of head h
        if (bh == 0) { dh.x = kbh; dl.x = kbl; } else if (bh == 1) { dh.x = kbh << 16; dl.x = kbl << 16; }
        else if (bh == 2) { dh.y = kbh; dl.y = kbl; } else { dh.y = kbh << 16; dl.y = kbl << 16; }
        *reinterpret_cast<uint4*>(bhi + 16 * LBO_V) = dh;
        *reinterpret_cast<uint4*>(blo + 16 * LBO_V) = dl;
      } else if (bj == 1) {
        *reinterpret_cast<uint4*>(bhi + 17 * LBO_V) = zero4;
        *reinterpret_cast<uint4*>(blo + 17 * LBO_V) = zero4;
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- q = [xn | g] [Wq_x | Wq_g]^T   (single terms: attention weights)
    if (issuer) {
      umma::fence_after_sync();
      if (!w_ready) { umma::mbar_wait(bar_w, 0); w_ready = true; }
      umma::mbar_wait(bar_g, ph_g);
      if (umma::elect_one()) {
        issue_gemm_k128_desc(tm + SA_TM_Q, d_xn, LBO_T, d_w, LBO_W, IDESC_128x128, false);
        issue_gemm_k128_desc(tm + SA_TM_Q, d_g, LBO_W, d_w + (uint64_t)(WIMG_BYTES >> 4), LBO_W, IDESC_128x128, true);
        umma::mma_commit(bar_acc);
      }
      __syncwarp();
    }
    ph_g ^= 1;
    umma::mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
    umma::fence_after_sync();
    // ---- phi(q) -> Q image over the LN tile (its reader, the q GEMM, has completed)
    {
      float v[32];
      umma::tmem_ld32(lane_addr + SA_TM_Q + cq * 32, v);
      const float* bb = s_bq + cq * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) { float a = v[i] + bb[i]; const float ex = __expf(fminf(a, 0.0f)); v[i] = a > 0.0f ? a + 1.0f : ex; }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(smem + SA_XN + (cq * 4 + c) * LBO_V + row * 16) =
            make_uint4(umma::pack_h2(v[c * 8], v[c * 8 + 1]), umma::pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                       umma::pack_h2(v[c * 8 + 4], v[c * 8 + 5]), umma::pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- [num | den] = phi(q) (Bstate hi + Bstate lo).  The shortcut rows are fetched under these MMAs.
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
          umma::mma_f16_ss(tm + SA_TM_ND, d_q + (uint64_t)(k * 2 * (LBO_V >> 4)), d_bh + (uint64_t)(k * 16), IDESC_APPLY, k > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma::mma_f16_ss(tm + SA_TM_ND, d_q + (uint64_t)(k * 2 * (LBO_V >> 4)), d_bl + (uint64_t)(k * 16), IDESC_APPLY, 1u);
        umma::mma_commit(bar_acc);
      }
      __syncwarp();
    }
    umma::mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
    umma::fence_after_sync();
    // ---- x tile -> staging (the Bstate images are dead now)
#pragma unroll
    for (int i = 0; i < 8; ++i) st4(stage + (warp * 8 + i) * STG_LD + lane * 4, xv[i]);
    __syncthreads();
    // ---- x1 = x + num / (den + eps) * S   (model.py:283-284, :412)
    {
      float den[8], v[32];
      umma::tmem_ld8(lane_addr + SA_TM_ND + 128, den);
      umma::tmem_ld32(lane_addr + SA_TM_ND + cq * 32, v);
      const float dsel = cq == 0 ? den[0] : (cq == 1 ? den[1] : (cq == 2 ? den[2] : den[3]));
      const float zs = fS / (dsel + 1e-6f);
      float* sp = stage + row * STG_LD + cq * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 x = ld4(sp + i);
        st4(sp + i, make_float4(fmaf(v[i], zs, x.x), fmaf(v[i + 1], zs, x.y), fmaf(v[i + 2], zs, x.z), fmaf(v[i + 3], zs, x.w)));
      }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int r = warp * 8 + i;
      if (r < nvalid) st4(X1 + row0off + (long long)r * rstride + lane * 4, ld4(stage + r * STG_LD + lane * 4));
    }
    __syncthreads();      // the staging tile (Bstate images) is rebuilt by the next item
  }
  if (warp == 0) umma::tmem_dealloc<512>(tm);
}

// ================================================================================================
cudaError_t launch_class_state_split(const float* X, const __half* timg, float* state, int B, int Te, int npix, int S,
                                     const ClassSplitW& w, int num_sms, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(class_state_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SS_SMEM);
  if (e != cudaSuccess) return e;
  const long long n = (long long)B * npix;
  const int grid = (int)(n < num_sms ? n : num_sms);
  if (grid <= 0) return cudaSuccess;
  class_state_split_kernel<<<grid, 512, SS_SMEM, st>>>(X, timg, state, B, Te, npix, S, w);
  return cudaGetLastError();
}

cudaError_t launch_class_apply_split(const float* X, float* X1, const __half* timg, const float* state, const float* pad_state,
                                     int B, int Te, int npix, int S, const ClassSplitW& w, int num_sms, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(class_apply_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SA_SMEM);
  if (e != cudaSuccess) return e;
  const long long n = (long long)B * npix * ((Te + 127) / 128);
  const int grid = (int)(n < num_sms ? n : num_sms);
  if (grid <= 0) return cudaSuccess;
  class_apply_split_kernel<<<grid, 512, SA_SMEM, st>>>(X, X1, timg, state, pad_state, B, Te, npix, S, w);
  return cudaGetLastError();
}

}  // namespace catseg
