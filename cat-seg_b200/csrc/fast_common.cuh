// Building blocks shared by the FAST (fp16 tcgen05, fp32 accumulate) kernels.
//
// Conventions
//   * A "token tile" is 128 rows (tokens) x 128 features.  For the epilogues one thread owns one row:
//     warp w of a CTA may only touch TMEM lanes 32*(w%4)..+31, so warps w and w+4 share rows and
//     split the columns.
//   * fp16 UMMA operands live in shared memory in the canonical no-swizzle K-major layout of
//     umma.cuh.  Tiles written by threads use a PADDED chunk stride (LBO_T = rows*16 + 16) so that
//     the 8-byte stores of the warp-per-row LayerNorm prologue are bank-conflict free; weight
//     images are pre-packed in global memory with the dense stride and arrive by bulk (TMA) copy.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace catseg {
namespace fast {

constexpr int TILE = 128;                       // rows per UMMA M tile
constexpr uint32_t LBO_W = 128 * 16;            // weight images: 128 rows, dense
constexpr uint32_t LBO_T = 128 * 16 + 16;       // thread-written 128-row tiles, padded
constexpr uint32_t TILE_BYTES_T = 16 * LBO_T;   // 128 features = 16 chunks
constexpr uint32_t WIMG_BYTES = 16 * LBO_W;     // one 128x128 fp16 weight image = 32 KiB
constexpr uint32_t IDESC_128x128 = umma::make_idesc_f16(128, 128);
constexpr uint32_t IDESC_BF16_128x128 = umma::make_idesc_bf16(128, 128);   // the 7x7 embedding keeps its hi+lo bf16 split (fast_prep.cu)

// 8 k-steps of a [128 x 128] x [N x 128]^T GEMM from PRE-BUILT base descriptors: the issuing thread only adds the
// K-step offset to the address field (building a descriptor from scratch costs ~50-100 cycles in one thread,
// longer than the MMA itself; measured with the phase-timing hook).
__device__ __forceinline__ void issue_gemm_k128_desc(uint32_t d_tmem, uint64_t a_desc, uint32_t lbo_a, uint64_t b_desc,
                                                     uint32_t lbo_b, uint32_t idesc, bool acc_first) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    umma::mma_f16_ss(d_tmem, a_desc + (uint64_t)(k * 2 * (lbo_a >> 4)), b_desc + (uint64_t)(k * 2 * (lbo_b >> 4)), idesc,
                      (k > 0 || acc_first) ? 1u : 0u);
}

// 8 k-steps of a [128 x 128] x [N=128 x 128]^T GEMM; A/B are shared-memory byte addresses
__device__ __forceinline__ void issue_gemm_k128(uint32_t d_tmem, uint32_t a_addr, uint32_t lbo_a, uint32_t b_addr,
                                                uint32_t lbo_b, uint32_t idesc, bool acc_first) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint64_t da = umma::make_smem_desc(a_addr + k * 2 * lbo_a, lbo_a, 128);
    uint64_t db = umma::make_smem_desc(b_addr + k * 2 * lbo_b, lbo_b, 128);
    umma::mma_f16_ss(d_tmem, da, db, idesc, (k > 0 || acc_first) ? 1u : 0u);
  }
}

// LayerNorm(128) of `nrows` token rows (row i at src + i*row_stride floats), written as fp16 into a
// canonical 128-row tile.  Warp-per-row: coalesced 512-byte loads, shuffle reductions.  Rows
// >= nvalid are zero-filled.  Called by all `nwarps` warps of the CTA.
__device__ __forceinline__ void ln_rows_to_tile(const float* __restrict__ src, long long row_stride, int nvalid,
                                                uint8_t* tile, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, int warp, int nwarps, int lane) {
  const float4 g = ld4(gamma + lane * 4), be = ld4(beta + lane * 4);
  constexpr int R = 8;                                   // rows in flight per warp (one memory latency per 8 rows)
  for (int r0 = warp * R; r0 < TILE; r0 += nwarps * R) {
    float4 x[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      int r = r0 + i;
      x[i] = (r < nvalid) ? ld4(src + (long long)r * row_stride + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      int r = r0 + i;
      float4 y = warp_layernorm128_fast(x[i], g, be);
      if (r >= nvalid) y = make_float4(0.f, 0.f, 0.f, 0.f);
      uint2 p = make_uint2(umma::pack_h2(y.x, y.y), umma::pack_h2(y.z, y.w));
      *reinterpret_cast<uint2*>(tile + (lane >> 1) * LBO_T + r * 16 + (lane & 1) * 8) = p;
    }
  }
}

// PRECISE variant: the normalised row is written as a hi + lo fp16 pair into two tiles (umma::split_h2); tile_lo may be
// nullptr when only the single-term operand is needed (q/k path, tools/precision_study.py).
__device__ __forceinline__ void ln_rows_to_tile_split(const float* __restrict__ src, long long row_stride, int nvalid,
                                                      uint8_t* tile_hi, uint8_t* tile_lo, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, int warp, int nwarps, int lane,
                                                      int tile_rows = TILE) {
  const float4 g = ld4(gamma + lane * 4), be = ld4(beta + lane * 4);
  constexpr int R = 8;
  for (int r0 = warp * R; r0 < tile_rows; r0 += nwarps * R) {
    float4 x[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      int r = r0 + i;
      x[i] = (r < nvalid) ? ld4(src + (long long)r * row_stride + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      int r = r0 + i;
      float4 y = warp_layernorm128_fast(x[i], g, be);
      if (r >= nvalid) y = make_float4(0.f, 0.f, 0.f, 0.f);
      uint2 hi, lo;
      umma::split_h2(y.x, y.y, hi.x, lo.x);
      umma::split_h2(y.z, y.w, hi.y, lo.y);
      const uint32_t off = (lane >> 1) * LBO_T + r * 16 + (lane & 1) * 8;
      *reinterpret_cast<uint2*>(tile_hi + off) = hi;
      if (tile_lo != nullptr) *reinterpret_cast<uint2*>(tile_lo + off) = lo;
    }
  }
}

// Same for NT consecutive 128-row tiles handled by 16 warps (tile t at tile + t*TILE_BYTES_T), software pipelined: a warp
// owns 8*NT consecutive rows and works in batches of BATCH rows, the loads of the next DEPTH-1 batches in flight while a
// batch is normalised.
// Measured in the FFN kernel (256 rows): loads alone 3.6 K cycles, arithmetic alone 4 K, load-all-then-compute 10.5 K --
// with only four warps per scheduler nothing else overlaps the two.
template <int NT, int BATCH = 2, int DEPTH = 4>
__device__ __forceinline__ void ln_rows_to_tiles_pipelined(const float* __restrict__ src, long long row_stride, int nvalid,
                                                           uint8_t* tile, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, int warp, int lane) {
  const float4 g = ld4(gamma + lane * 4), be = ld4(beta + lane * 4);
  constexpr int NB = 8 * NT / BATCH;                       // batches per warp; DEPTH batches are in flight
  const int rw = warp * (8 * NT);
  float4 x[DEPTH][BATCH];
  auto load = [&](int b, float4* v) {
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int r = rw + b * BATCH + i;
      v[i] = (r < nvalid) ? ld4(src + (long long)r * row_stride + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
#pragma unroll
  for (int b = 0; b < DEPTH - 1 && b < NB; ++b) load(b, x[b]);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (b + DEPTH - 1 < NB) load(b + DEPTH - 1, x[(b + DEPTH - 1) % DEPTH]);
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int r = rw + b * BATCH + i;
      float4 y = warp_layernorm128_fast(x[b % DEPTH][i], g, be);
      if (r >= nvalid) y = make_float4(0.f, 0.f, 0.f, 0.f);
      const uint2 pk = make_uint2(umma::pack_h2(y.x, y.y), umma::pack_h2(y.z, y.w));
      *reinterpret_cast<uint2*>(tile + (r >> 7) * TILE_BYTES_T + (lane >> 1) * LBO_T + (r & 127) * 16 + (lane & 1) * 8) = pk;
    }
  }
}

// erf-GELU for the fp16 path.  The result is rounded to fp16 (relative 2^-9) before the next MMA, so erf only
// needs ~1e-4 absolute accuracy: odd minimax polynomial directly in x, erf(x/sqrt 2) ~ x Q(x^2) on |x| <= 3.8
// (clamped beyond, where 1 - erf < 1.5e-4); max |erf error| 1.3e-4, GELU error <= 2.5e-4 (1.6e-4 for |x| < 2.5).
// 12 FMA-pipe instructions, no MUFU (the FFN epilogue is issue bound: ncu, profiles/README.md).
__device__ __forceinline__ float gelu_fast(float x) {
  const float xc = fminf(fmaxf(x, -3.8f), 3.8f);
  const float t = xc * xc;
  // gelu = x (0.5 + xc Q(t) / 2): the 1/2 is folded into the coefficients (exact), one instruction fewer than h + h xc Q
  float q = fmaf(0.5f * 7.331543372e-08f, t, 0.5f * -4.544918738e-06f);
  q = fmaf(q, t, 0.5f * 1.213695141e-04f);
  q = fmaf(q, t, 0.5f * -1.863094512e-03f);
  q = fmaf(q, t, 0.5f * 1.863326877e-02f);
  q = fmaf(q, t, 0.5f * -1.314395666e-01f);
  q = fmaf(q, t, 0.5f * 7.973535061e-01f);
  return x * fmaf(xc, q, 0.5f);
}

// ---- packed fp32 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 on register pairs).  Same FLOP rate as the scalar forms
// (tools/probes/ffma2_probe.cu) but half the issue slots, which is what bounds the activation epilogues.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_fast on two values at once: the same operations in the same order per element (bit-identical results),
// 7 instead of 12 instructions per element.  x0/x1 are updated in place; (b0, b1) is the bias to add first.
__device__ __forceinline__ void gelu_fast_pair(float& x0, float& x1, float b0, float b1) {
  const unsigned long long X = add2(pk2(x0, x1), pk2(b0, b1));
  float a0, a1;
  unpk2(X, a0, a1);
  const unsigned long long XC = pk2(fminf(fmaxf(a0, -3.8f), 3.8f), fminf(fmaxf(a1, -3.8f), 3.8f));
  const unsigned long long T = mul2(XC, XC);
  constexpr float c6 = 0.5f * 7.331543372e-08f, c5 = 0.5f * -4.544918738e-06f, c4 = 0.5f * 1.213695141e-04f,
                  c3 = 0.5f * -1.863094512e-03f, c2 = 0.5f * 1.863326877e-02f, c1 = 0.5f * -1.314395666e-01f,
                  c0 = 0.5f * 7.973535061e-01f;
  unsigned long long Q = fma2(pk2(c6, c6), T, pk2(c5, c5));
  Q = fma2(Q, T, pk2(c4, c4));
  Q = fma2(Q, T, pk2(c3, c3));
  Q = fma2(Q, T, pk2(c2, c2));
  Q = fma2(Q, T, pk2(c1, c1));
  Q = fma2(Q, T, pk2(c0, c0));
  unpk2(mul2(X, fma2(XC, Q, pk2(0.5f, 0.5f))), x0, x1);
}

// fp32 staging tile [128 rows][128 cols] with a 132-float row stride: conflict-free both for "thread = row"
// float4 accesses and for "warp = row" coalesced accesses.  Used to turn per-thread-row TMEM epilogues into
// coalesced global loads/stores (a per-thread-row global read-modify-write stalls for microseconds: ncu).
constexpr int STG_LD = 132;
constexpr uint32_t STG_BYTES = 128 * STG_LD * 4;

}  // namespace fast
}  // namespace catseg
