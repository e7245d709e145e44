// Building blocks of the PRECISE tensor-core kernels (fp16 hi + lo operand pairs, fp32 accumulation).
//
// Why a split.  north_star asks for >= 99.9 % argmax agreement with the fp32 reference; with random-init weights the
// top-1/top-2 margin is ~1e-5 at the 0.1 % quantile, so single 16-bit operands cannot reach it
// (bf16 87 %, fp16 98 %).  tools/precision_study.py measured, on the CPU oracle with emulated operand rounding, which
// contractions need more than one term:
//   * the q / k projections, Q K^T, phi(q), phi(k) (everything that only shapes ATTENTION WEIGHTS) are fine with a
//     single fp16 term, provided numerator and normaliser see the same rounded weights (P V and its row sum, phi(q) KV and
//     phi(q) Ksum come out of the same MMA);
//   * everything on the VALUE path (v projection, P V's V operand, proj, both FFN GEMMs, the linear-attention state, the
//     class MLP, every decoder convolution) needs x = hi + lo on both operands: A B ~ Ah Bh + Al Bh + Ah Bl
//     (the lo*lo term is below 2^-22).  Emulated result of this scheme: max-abs 2.5e-5, raw argmax >= 99.989 %.
// A weight therefore ships as two canonical images (hi, lo); an activation tile is written twice by its producer
// (umma::split_h2: 6 instructions per element pair).
#pragma once
#include "fast_common.cuh"

namespace catseg {
namespace split {

using namespace fast;

// Weight ring of the issuing warp: images of IMG_BYTES stream from global memory, in CONSUMPTION order and cyclically
// (`per_item` images per work item), through NS shared-memory slots.  Image n lives in slot n % NS; it is re-filled
// with image n + NS once the MMAs that read it have completed (tcgen05.commit -> bar_empty).  All methods are called by
// every lane of the issuing warp (uniform state in registers); the asynchronous instructions run on one elected lane.
template <int NS, uint32_t IMG_BYTES>
struct WeightRing {
  uint8_t* slots;
  uint64_t *bar_full, *bar_empty;
  const uint8_t* src;
  int per_item;
  long long total;       // images this CTA will consume over its lifetime
  long long n;           // images consumed so far
  int slot, par;         // slot / parity of image n
  int pslot, ppar;       // ... of image n - 1
  long long nload;       // next image to request
  int lslot, limg;       // its slot and its index within the item
  int dbg_nostream;      // measurement aid (CATSEG_DBG_NOSTREAM=1): only the first NS images are ever fetched (WRONG results)

  __device__ __forceinline__ void init(uint8_t* slots_, uint64_t* full, uint64_t* empty, const void* src_, int per_item_,
                                       long long total_) {
    slots = slots_; bar_full = full; bar_empty = empty; src = reinterpret_cast<const uint8_t*>(src_);
    per_item = per_item_; total = total_;
    n = 0; slot = 0; par = 0; pslot = NS - 1; ppar = 1; nload = 0; lslot = 0; limg = 0; dbg_nostream = 0;
  }
  __device__ __forceinline__ void load_next() {
    if (nload < total && !(dbg_nostream && nload >= NS - 1) && umma::elect_one()) {
      umma::mbar_expect_tx(&bar_full[lslot], IMG_BYTES);
      umma::bulk_g2s(slots + (uint32_t)lslot * IMG_BYTES, src + (size_t)limg * IMG_BYTES, IMG_BYTES, &bar_full[lslot]);
    }
    __syncwarp();
    ++nload;
    lslot = lslot + 1 == NS ? 0 : lslot + 1;
    limg = limg + 1 == per_item ? 0 : limg + 1;
  }
  __device__ __forceinline__ void prime() {
    for (int i = 0; i < NS - 1; ++i) load_next();
  }
  // waits until image n is resident; returns its shared-memory byte offset from `slots`
  __device__ __forceinline__ uint32_t acquire() {
    if (dbg_nostream) { if (n < NS - 1) umma::mbar_wait(&bar_full[slot], 0u); return (uint32_t)(n < NS - 1 ? slot : 0) * IMG_BYTES; }
    umma::mbar_wait(&bar_full[slot], (uint32_t)par);
    return (uint32_t)slot * IMG_BYTES;
  }
  // the MMAs reading image n have been issued (by the elected lane): track their completion, re-fill the slot of image n - 1
  __device__ __forceinline__ void release() {
    if (umma::elect_one()) umma::mma_commit(&bar_empty[slot]);
    __syncwarp();
    if (nload < total) {
      if (n > 0 && !dbg_nostream) umma::mbar_wait(&bar_empty[pslot], (uint32_t)ppar);
      load_next();
    } else {
      ++nload;
    }
    pslot = slot; ppar = par;
    ++n;
    slot = slot + 1 == NS ? 0 : slot + 1;
    if (slot == 0) par ^= 1;
  }
};

// GELU with the exact erf (nn.GELU() default, model.py:139) for the PRECISE path:  x Phi(x),  Phi(x) = erfc(-x / sqrt 2) / 2.
// erfc(z), z >= 0, from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z);
// the complementary form has no cancellation for negative x.  Branch free, two MUFU operations (rcp, ex2) and ten FMA-pipe
// instructions per element; measured max-abs error against float64 over [-8, 8]: 4.2e-7 (torch's fp32 F.gelu: 1.2e-6).
// erff() costs about three times as many issue slots (two divergent branches), and this epilogue is ALU bound.
__device__ __forceinline__ float gelu_precise(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float q = 0.5f * poly * t * umma::ex2_approx(x * x * -0.72134752044448170368f);
  return x * (x < 0.0f ? q : 1.0f - q);
}

// The same function on two values at once, on packed fp32 pairs (FFMA2 / FMUL2 on sm_100a: half the issue slots of the
// scalar form, which is what bounds this epilogue), rearranged as  GELU(x) = max(x, 0) - |x| q(|x|),  q = erfc(|x| / sqrt 2) / 2:
// no compare / select / 1 - q, the 1/2 folded into the polynomial.  14 issue slots + 4 MUFU per PAIR (scalar form: 15 + 2 per
// element); max-abs error against float64 over [-8, 8]: 3.3e-7.  (b0, b1) is the bias to add first; x0 / x1 are updated in place.
__device__ __forceinline__ void gelu_precise_pair(float& x0, float& x1, float b0, float b1) {
  const unsigned long long X = add2(pk2(x0, x1), pk2(b0, b1));
  float a0, a1;
  unpk2(X, a0, a1);
  const unsigned long long NA = pk2(-fabsf(a0), -fabsf(a1));                       // -|x|
  constexpr float pz = -0.3275911f * 0.70710678118654752440f;
  float d0, d1, t0, t1;
  unpk2(fma2(NA, pk2(pz, pz), pk2(1.0f, 1.0f)), d0, d1);                           // 1 + p |x| / sqrt 2
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const unsigned long long T = pk2(t0, t1);
  constexpr float c5 = 0.5f * 1.061405429f, c4 = 0.5f * -1.453152027f, c3 = 0.5f * 1.421413741f, c2 = 0.5f * -0.284496736f,
                  c1 = 0.5f * 0.254829592f, ke = -0.72134752044448170368f;
  unsigned long long P = fma2(T, pk2(c5, c5), pk2(c4, c4));
  P = fma2(P, T, pk2(c3, c3));
  P = fma2(P, T, pk2(c2, c2));
  P = fma2(P, T, pk2(c1, c1));
  float e0, e1;
  unpk2(mul2(mul2(X, X), pk2(ke, ke)), e0, e1);                                    // -x^2 / 2 * log2(e)
  const unsigned long long E = pk2(umma::ex2_approx(e0), umma::ex2_approx(e1));
  const unsigned long long Q = mul2(mul2(P, T), E);                                // q(|x|)
  unpk2(fma2(NA, Q, pk2(fmaxf(a0, 0.0f), fmaxf(a1, 0.0f))), x0, x1);
}

}  // namespace split
}  // namespace catseg
