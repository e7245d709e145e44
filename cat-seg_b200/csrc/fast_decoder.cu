// FAST guided upsampling decoder on tcgen05: implicit-GEMM "band convolution".
//
// Reference: Aggregator.conv_decoder (model.py:674-681), Up.forward (:549-555), DoubleConv (:520-537).
//
// One CTA processes a band of BR input rows of one (image, class) slice.  The band (plus a 1-pixel
// halo, zero padded) is staged ONCE in shared memory as a fp16 canonical K-major image whose rows
// are the padded raster positions; every convolution tap is then just a row-shifted view of that
// image (descriptor start address + offset*16 bytes), so a 3x3 conv is 9 x (CIN/16) accumulating
// tcgen05.mma per 128-row tile with no im2col copy.  GroupNorm+ReLU of the producer is applied
// while staging (statistics come from the producer's epilogue as per-band partial sums, reduced in
// a fixed order: deterministic); this kernel's epilogue emits the statistics of its own output.
//
// Results-preserving algebra (verified in fp64 there):
//   * ConvTranspose2d(k2,s2) followed by conv3x3 is composed at load time into, per output parity
//     (a,b), a 2x2-tap convolution on the LOW-resolution grid with CIN = channels before the
//     transposed conv (128 -> 64 at 48^2, 64 -> 32 at 96^2): 43 % fewer MACs and no upsampled tensor;
//   * the guidance channels of the concatenated input and the transposed-conv bias only depend on
//     the image (resp. only on the weights): they enter as an additive per-image map E.
// Stages: D1 x(24^2,128) -> c1a(48^2,64) | D2 c1a -> c1b | D3 c1b -> c2a(96^2,32) | D4 c2a -> c2b |
//         D5 head c2b -> logits (N padded to 16, column 0 real).
#include <cstdio>
#include <cstdlib>

#include "fast_common.cuh"
#include "igemm.cuh"
#include "internal.h"

namespace catseg {

using namespace fast;

static __device__ unsigned kBackoffNs = 200;     // CATSEG_DEC_BACKOFF_NS (A/B)

struct BandConvParams {
  const void* in;            // [S][Win*Win][CIN]  fp32 (IN_F32) or fp16
  const float* in_stats;     // [S][nb_in][G_in][2] partial (sum, sumsq); nullptr: no GroupNorm on the input
  const float* in_ss;        // [S][CIN][2] GroupNorm (scale, shift) of the input, from gn_finalize_kernel (set when in_stats is)
  int nb_in;
  float in_count;            // elements per (slice, group)
  const float *gamma, *beta; // [CIN]
  const __half* wimg; // NIMG images [NOUT x CIN], canonical dense
  const float* emap;         // composed stages: the additive map re-laid out per accumulator tile (relayout_emap_kernel):
                             // [B][NB][4 parities][NTILES][NREAL/4][128 rows][4] fp32, else nullptr
  int Te;
  __half* out;        // [S][Wout*Wout][NREAL]
  float* out32;       // PRECISE: the same tensor in fp32
  float* out_stats;          // [S][NB][G_out][2]
  float* logits;             // HEAD: [B][T][Wout*Wout]
  const int32_t* classes;    // HEAD: [B*Te] (absolute slice index)
  int T;
  float head_bias;
  int slice0;                // absolute index of local slice 0 (for b = (slice0+s)/Te and classes)
  int nslice;                // local slices in this launch
  long long* dbg;            // optional phase timing (CATSEG_PHASE_TIMING=1)
};

// SPLIT = PRECISE mode (split_common.cuh): the staged band exists as a hi and a lo fp16 image, every weight image as a hi
// and a lo image, each tap issues three products (hi*hi, lo*hi, hi*lo), and the intermediates in HBM are fp32.
template <int CIN, int NOUT, int NREAL, bool UPS, bool IN_F32, int WIN_, int BR, bool HEAD, int CTAS = 2, bool SPLIT = false>
struct BandCfg {
  static constexpr int PW = WIN_ + 2, NP = (BR + 2) * PW, P0 = PW + 1;
  static constexpr int MROWS = (BR - 1) * PW + WIN_, NTILES = (MROWS + 127) / 128;
  static constexpr int KCH = CIN / 8, KSTEPS = CIN / 16;
  static constexpr uint32_t LBO_I = NP * 16 + ((16 + 128 - (NP * 16) % 128) % 128);   // LBO_I % 128 == 16
  static constexpr int OVER = P0 + NTILES * 128 + PW + 2 - NP;
  static constexpr uint32_t IMG_BYTES = ((KCH * LBO_I + (OVER > 0 ? OVER * 16 : 0)) + 127) / 128 * 128;
  static constexpr int NTAP = UPS ? 4 : 9, NPG = UPS ? 4 : 1, NIMG = NTAP * NPG;
  static constexpr int NW = SPLIT ? 2 : 1;                               // terms per weight / per staged band (hi, lo)
  // SPLIT: the hi and lo terms of a tap's weights form ONE image with 2 NOUT rows (B operand rows = N): the product
  // Ah [Wh | Wl] is a single MMA of N = 2 NOUT (its two column halves are summed in the epilogue), Al Wh a second one of
  // N = NOUT on the first NOUT rows: two MMAs instead of three per k-step, both at a wider N (the N <= 64 MMAs of this
  // decoder sit at the ~46-cycle operand-read floor), and half as many ring images to wait for.
  static constexpr int NACC = NW * NOUT;                                 // accumulator columns per M tile
  static constexpr uint32_t WBYTES = NW * NOUT * CIN * 2, LBO_WT = NW * NOUT * 16;
  // two CTAs per SM (half-height bands, <= 113 KiB and <= 256 TMEM columns each): one CTA's staging / epilogue
  // overlaps the other's MMAs.  Small weight sets stay resident, larger ones stream through a ring.
  // (the whole weight set may stay in shared memory when, together with the band image, it fits half an SM)
  static constexpr bool RESIDENT = NIMG * WBYTES <= 24 * 1024 ||
                                   NW * IMG_BYTES + NIMG * WBYTES <= (CTAS == 2 ? 108 : (SPLIT ? 214 : 200)) * 1024;
  // streaming ring: 48 KiB deep (96 KiB when the CTA owns the SM), so that the prefetch distance (in MMA time) exceeds the
  // ~1 us L2->SMEM latency
  // (a band image that leaves less than 96 KiB gets a 64 KiB ring: taller bands amortise more than the shallower prefetch costs)
  static constexpr int RING_KB = CTAS == 2 ? 48 : (NW * IMG_BYTES + 98 * 1024 <= 227 * 1024 ? 96 : 64);
  static constexpr int NSLOT = RESIDENT ? NIMG : (int)(RING_KB * 1024 / WBYTES);
  // worker warps (staging + epilogues): 8 per CTA with two CTAs per SM, 16 when the CTA owns the SM; one more warp issues
  static constexpr int NWW = CTAS == 2 ? 8 : 16, NWT = NWW * 32, THREADS = NWT + 32, NTG = NWW / 4;
  static constexpr int NB = WIN_ / BR;                 // bands per slice
  static constexpr int WOUT = UPS ? 2 * WIN_ : WIN_;
  static constexpr int GOUT = HEAD ? 1 : NREAL / 16;
  static constexpr uint32_t SM_W = NW * IMG_BYTES;
  static constexpr uint32_t SM_SC = SM_W + NSLOT * WBYTES;            // scale[CIN], shift[CIN]
  static constexpr uint32_t SM_ST = SM_SC + 2 * CIN * 4;              // [NWW warps][GOUT][2]
  static constexpr uint32_t SM_BAR = (SM_ST + NWW * GOUT * 2 * 4 + 15) / 16 * 16;
  static constexpr uint32_t SMEM = SM_BAR + (2 * NSLOT + 33) * 8 + 16;   // ring barriers + [2 sets][16 tiles] accumulator barriers
  static constexpr uint32_t IDESC = umma::make_idesc_f16(128, NACC);     // A (hi) x the whole image
  static constexpr uint32_t IDESC_LO = umma::make_idesc_f16(128, NOUT);  // SPLIT: A (lo) x the hi rows
  // CTAS = CTAs per SM: 2 (half-height bands; one CTA's staging / epilogue overlaps the other's MMAs) or 1 (a band as
  // tall as the shared memory allows: fewer M-tile remainders and one pass over a streamed weight set per band)
  static constexpr int TMEM_COLS = CTAS == 2 ? 256 : 512;
  static_assert(NTILES * NACC <= TMEM_COLS && NTILES <= 16, "TMEM columns");
  static_assert(SMEM <= (CTAS == 2 ? 113 * 1024 : 227 * 1024), "shared memory budget");
  static_assert(!SPLIT || IN_F32, "PRECISE stages read fp32 activations");
  static_assert(WIN_ % BR == 0 && CIN % 16 == 0 && NOUT % 16 == 0, "shape");
};

// Roles: warps 0-7 stage the band and run the epilogues (TMEM lane quarter q4 = warp & 3, the two warp sets take
// alternate tiles); warp 8 only issues (tcgen05.mma, weight ring).  The MMA queue is shallow, so the issuing thread is
// blocked while its MMAs execute: a worker that also issued would join every epilogue late and hold up its barrier.
// The composed (UPS) stages compute four output parities from the same staged image: with two accumulator sets the MMAs
// of parity pg+1 run under the epilogue of parity pg.

template <int CIN, int NOUT, int NREAL, bool UPS, bool IN_F32, int WIN_, int BR, bool HEAD, int CTAS, bool SPLIT>
__global__ void __launch_bounds__((CTAS == 2 ? 8 : 16) * 32 + 32, CTAS) band_conv_kernel(BandConvParams p) {
  using C = BandCfg<CIN, NOUT, NREAL, UPS, IN_F32, WIN_, BR, HEAD, CTAS, SPLIT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* s_scale = reinterpret_cast<float*>(smem + C::SM_SC);
  float* s_shift = s_scale + CIN;
  float* s_part = reinterpret_cast<float*>(smem + C::SM_ST);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + C::SM_BAR);   // [NSLOT]
  uint64_t* bar_empty = bar_full + C::NSLOT;                             // [NSLOT]
  uint64_t* bar_acc = bar_empty + C::NSLOT;                              // [2 accumulator sets][8]: one barrier per M tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, q4 = warp & 3, tgrp = (warp >> 2) % C::NTG;
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == C::NWW;           // warp-uniform role
  constexpr int ACC_COLS = C::NTILES * C::NACC;
  constexpr int SETS = (C::NPG > 1 && 2 * ACC_COLS <= C::TMEM_COLS) ? 2 : 1;

  const long long nitems = (long long)p.nslice * C::NB;
  long long mine = 0;
  for (long long i = blockIdx.x; i < nitems; i += gridDim.x) ++mine;
  const long long total_loads = C::RESIDENT ? 1 : mine * C::NIMG;

  if (tid == 0) {
    for (int i = 0; i < 2 * C::NSLOT + 32; ++i) umma::mbar_init(&bar_full[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t sb = umma::smem_u32(smem);
  const uint32_t lane_addr = tm + ((uint32_t)(q4 * 32) << 16);

  auto issue_load = [&](long long n) {      // one elected lane of the issuing warp (streaming mode)
    if (n < total_loads) {
      int s = (int)(n % C::NSLOT);
      umma::mbar_expect_tx(&bar_full[s], C::WBYTES);
      umma::bulk_g2s(smem + C::SM_W + s * C::WBYTES,
                     reinterpret_cast<const uint8_t*>(p.wimg) + (n % C::NIMG) * C::WBYTES, C::WBYTES, &bar_full[s]);
    }
  };
  if (issuer) {
    if (umma::elect_one()) {
      if (C::RESIDENT) {
        umma::mbar_expect_tx(&bar_full[0], C::NIMG * C::WBYTES);
        umma::bulk_g2s(smem + C::SM_W, p.wimg, C::NIMG * C::WBYTES, &bar_full[0]);
      } else {
        for (int i = 0; i < C::NSLOT - 1; ++i) issue_load(i);
      }
    }
    __syncwarp();
  }
  long long nimg = 0;          // streaming: images consumed so far by this CTA (issuing warp, all lanes)
  long long iacc_full = 0, iacc_issue = 0, iacc_empty = 0;   // CATSEG_PHASE_TIMING: issuing warp, streaming mode
  uint32_t ph_acc[2] = {0, 0};
  bool w_ready = false;
  long long t_last = clock64(), pacc0 = 0, pacc1 = 0, pacc2 = 0, pacc3 = 0, pacc4 = 0, nit_dbg = 0;
#define BPH(i) do { if (p.dbg != nullptr && blockIdx.x == 0 && tid == 0) { long long _t = clock64(); pacc##i += _t - t_last; t_last = _t; } } while (0)

  // all MMAs of one parity group into accumulator set `set` (every lane of the issuing warp runs the control flow and
  // keeps the ring state; one elected lane issues)
  auto issue_group = [&](int pg, int set) {
    const int pa = pg >> 1, pb = pg & 1;
    if (C::RESIDENT && !w_ready) { umma::mbar_wait(&bar_full[0], 0); w_ready = true; }
    const uint64_t a_desc0 = umma::make_smem_desc(sb, C::LBO_I, 128);
    if constexpr (C::RESIDENT) {
      // resident weights: tile-outer order with one commit per M tile, so the epilogue of tile t runs under the MMAs of
      // the tiles behind it
      if (umma::elect_one()) {
#pragma unroll 1
        for (int t = 0; t < C::NTILES; ++t) {
#pragma unroll 1
          for (int tap = 0; tap < C::NTAP; ++tap) {
            int off;
            if (UPS) off = ((tap >> 1) + pa - 1) * C::PW + ((tap & 1) + pb - 1);
            else off = (tap / 3 - 1) * C::PW + (tap % 3 - 1);
            const uint64_t b_desc = umma::make_smem_desc(sb + C::SM_W + (uint32_t)(pg * C::NTAP + tap) * C::WBYTES, C::LBO_WT, 128);
            const uint64_t a_tile = a_desc0 + (uint64_t)(uint32_t)(C::P0 + off + t * 128);
#pragma unroll
            for (int k = 0; k < C::KSTEPS; ++k)
              umma::mma_f16_ss(tm + set * ACC_COLS + t * C::NACC, a_tile + (uint64_t)(k * 2 * (C::LBO_I >> 4)),
                                b_desc + (uint64_t)(k * 2 * (C::LBO_WT >> 4)), C::IDESC, (tap > 0 || k > 0) ? 1u : 0u);
            if constexpr (SPLIT) {
#pragma unroll
              for (int k = 0; k < C::KSTEPS; ++k)      // lo activations x hi weights (the first NOUT rows of the image)
                umma::mma_f16_ss(tm + set * ACC_COLS + t * C::NACC, a_tile + (uint64_t)((C::IMG_BYTES >> 4) + k * 2 * (C::LBO_I >> 4)),
                                 b_desc + (uint64_t)(k * 2 * (C::LBO_WT >> 4)), C::IDESC_LO, 1u);
            }
          }
          umma::mma_commit(&bar_acc[set * 16 + t]);
        }
      }
      __syncwarp();
      return;
    }
#pragma unroll 1
    for (int tap = 0; tap < C::NTAP; ++tap) {
      int off;
      if (UPS) off = ((tap >> 1) + pa - 1) * C::PW + ((tap & 1) + pb - 1);
      else off = (tap / 3 - 1) * C::PW + (tap % 3 - 1);
      const int slot = (int)(nimg % C::NSLOT);
      const long long ti0 = clock64();
      umma::mbar_wait(&bar_full[slot], (uint32_t)((nimg / C::NSLOT) & 1));
      const long long ti1 = clock64();
      const uint32_t wb = sb + C::SM_W + (uint32_t)slot * C::WBYTES;
      // descriptors are built once and advanced by integer adds on the (address >> 4) field
      const uint64_t b_desc = umma::make_smem_desc(wb, C::LBO_WT, 128);
      if (umma::elect_one()) {
        uint64_t a_tile = a_desc0 + (uint64_t)(uint32_t)(C::P0 + off);
#pragma unroll 1
        for (int t = 0; t < C::NTILES; ++t) {
#pragma unroll
          for (int k = 0; k < C::KSTEPS; ++k)
            umma::mma_f16_ss(tm + set * ACC_COLS + t * C::NACC, a_tile + (uint64_t)(k * 2 * (C::LBO_I >> 4)),
                              b_desc + (uint64_t)(k * 2 * (C::LBO_WT >> 4)), C::IDESC, (tap > 0 || k > 0) ? 1u : 0u);
          if (SPLIT) {
#pragma unroll
            for (int k = 0; k < C::KSTEPS; ++k)
              umma::mma_f16_ss(tm + set * ACC_COLS + t * C::NACC, a_tile + (uint64_t)((C::IMG_BYTES >> 4) + k * 2 * (C::LBO_I >> 4)),
                               b_desc + (uint64_t)(k * 2 * (C::LBO_WT >> 4)), C::IDESC_LO, 1u);
          }
          a_tile += 128;
        }
        umma::mma_commit(&bar_empty[slot]);
      }
      __syncwarp();
      const long long ti2 = clock64();
      {
        // refill the slot of the PREVIOUS image (its MMAs were committed one step ago)
        const long long nn = nimg + C::NSLOT - 1;
        if (nimg > 0 && nn < total_loads)
          umma::mbar_wait(&bar_empty[(nimg - 1) % C::NSLOT], (uint32_t)(((nimg - 1) / C::NSLOT) & 1));
        if (umma::elect_one()) issue_load(nn);
        __syncwarp();
        ++nimg;
      }
      iacc_full += ti1 - ti0; iacc_issue += ti2 - ti1; iacc_empty += clock64() - ti2;
    }
    if (umma::elect_one()) {
#pragma unroll
      for (int t = 0; t < C::NTILES; ++t) umma::mma_commit(&bar_acc[set * 16 + t]);
    }
    __syncwarp();
  };

  // band statistics: per-warp partials in s_part, summed in fixed order by 2 GOUT threads one barrier later (while the
  // other warps already wait for the next band's MMAs)
  long long stats_item = -1;
  auto flush_stats = [&]() {
    if (!HEAD && stats_item >= 0 && tid < C::GOUT * 2) {
      float a = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < C::NWW; ++w8) a += s_part[w8 * C::GOUT * 2 + tid];
      p.out_stats[stats_item * (C::GOUT * 2) + tid] = a;
    }
  };
  float2 ss_next = make_float2(0.f, 0.f);
  if (p.in_stats != nullptr && tid < CIN && (long long)blockIdx.x < nitems)
    ss_next = __ldg(reinterpret_cast<const float2*>(p.in_ss) + ((long long)blockIdx.x / C::NB) * CIN + tid);
  for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int sl = (int)(it / C::NB), band = (int)(it % C::NB);
    const int gslice = p.slice0 + sl;
    const int b = gslice / p.Te;
    if (!issuer) {   // the input rows of this CTA's NEXT band (one contiguous span) are prefetched into L2 meanwhile
      const long long itn = it + gridDim.x;
      if (itn < nitems) {
        const int sln = (int)(itn / C::NB), bandn = (int)(itn % C::NB);
        const int y0 = bandn * BR - 1 < 0 ? 0 : bandn * BR - 1, y1 = bandn * BR + BR + 1 > WIN_ ? WIN_ : bandn * BR + BR + 1;
        constexpr int ESZ = IN_F32 ? 4 : 2;
        const char* base = reinterpret_cast<const char*>(p.in) + (((long long)sln * WIN_ + y0) * WIN_) * CIN * ESZ;
        const int nlines = (y1 - y0) * WIN_ * CIN * ESZ / 128;
        for (int i = tid; i < nlines; i += C::NWT) umma::prefetch_l2(base + (long long)i * 128);
      }
    }
    // ---- GroupNorm parameters of the input (fixed-order reduction of the producer's band partials)
    //      (fetched one band ahead: the dependent L2 round trip used to cost ~1.5 K cycles per band)
    if (p.in_stats != nullptr) {
      if (tid < CIN) {
        s_scale[tid] = ss_next.x;
        s_shift[tid] = ss_next.y;
        const long long itn = it + gridDim.x;
        if (itn < nitems) ss_next = __ldg(reinterpret_cast<const float2*>(p.in_ss) + (itn / C::NB) * CIN + tid);
      }
      __syncthreads();
    }
    BPH(4);
    // ---- stage the padded band image (fp16, canonical K-major, rows = padded raster positions).
    //      Explicitly software-pipelined: U independent 16/32-byte loads are issued before any is consumed
    //      (ncu showed the compiler serialising load -> convert -> store per chunk: one latency per chunk).
    //      256 % KCH == 0, so a thread always handles the same 8-channel group: its GroupNorm scale/shift live in registers.
    if (!issuer) {
      const int y_first = band * BR - 1;
      constexpr int U = 8;                      // chunks in flight per thread (fp32 input: 2 x 16 bytes each)
      constexpr int NCHUNK = C::NP * C::KCH;
      static_assert(C::NWT % C::KCH == 0, "a thread keeps its channel group");
      const int c = tid % C::KCH;
      float sc[8], sh[8];
      if (p.in_stats != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = s_scale[c * 8 + j]; sh[j] = s_shift[c * 8 + j]; }
      }
#pragma unroll 1
      for (int base = tid; base < NCHUNK; base += C::NWT * U) {
        uint4 raw[U][IN_F32 ? 2 : 1];
        bool inb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int idx = base + u * C::NWT;
          const int pp = idx / C::KCH;
          const int yy = y_first + pp / C::PW, xx = pp % C::PW - 1;
          inb[u] = idx < NCHUNK && yy >= 0 && yy < WIN_ && xx >= 0 && xx < WIN_;
          const long long off = inb[u] ? (((long long)sl * WIN_ + yy) * WIN_ + xx) * CIN + c * 8 : 0;
          if (IN_F32) {
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.in) + off);
            raw[u][0] = __ldg(src);
            raw[u][IN_F32 ? 1 : 0] = __ldg(src + 1);
          } else {
            raw[u][0] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.in) + off));
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int idx = base + u * C::NWT;
          if (idx >= NCHUNK) continue;
          const int pp = idx / C::KCH;
          uint4 val = make_uint4(0u, 0u, 0u, 0u), vlo = make_uint4(0u, 0u, 0u, 0u);
          if (inb[u]) {
            float v[8];
            if (IN_F32) {
              const float* f = reinterpret_cast<const float*>(&raw[u][0]);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = f[j];
            } else {
              const __half2* h2 = reinterpret_cast<const __half2*>(&raw[u][0]);
#pragma unroll
              for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h2[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
            }
            if (p.in_stats != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.0f);
            }
            if constexpr (SPLIT) {
              umma::split_h2(v[0], v[1], val.x, vlo.x);
              umma::split_h2(v[2], v[3], val.y, vlo.y);
              umma::split_h2(v[4], v[5], val.z, vlo.z);
              umma::split_h2(v[6], v[7], val.w, vlo.w);
            } else {
              val = make_uint4(umma::pack_h2(v[0], v[1]), umma::pack_h2(v[2], v[3]), umma::pack_h2(v[4], v[5]),
                               umma::pack_h2(v[6], v[7]));
            }
          }
          *reinterpret_cast<uint4*>(smem + c * C::LBO_I + pp * 16) = val;
          if constexpr (SPLIT) *reinterpret_cast<uint4*>(smem + C::IMG_BYTES + c * C::LBO_I + pp * 16) = vlo;
        }
      }
      umma::fence_proxy_async();
    }
    umma::fence_before_sync();
    __syncthreads();
    BPH(0);
    flush_stats();
    if (issuer) {
      umma::fence_after_sync();
      issue_group(0, 0);
      if (SETS == 2) issue_group(1, 1);
    }

    float st_sum[C::GOUT], st_sq[C::GOUT];
#pragma unroll
    for (int g = 0; g < C::GOUT; ++g) st_sum[g] = st_sq[g] = 0.0f;

#pragma unroll 1
    for (int pg = 0; pg < C::NPG; ++pg) {
      const int pa = pg >> 1, pb = pg & 1;
      const int set = pg % SETS;
      const uint32_t acc_addr = lane_addr + set * ACC_COLS;
      if (!issuer) {
        // ---- epilogue: thread = padded raster row; the two warp sets take alternate tiles.  The additive map
        //      of the first tile is fetched BEFORE waiting for the accumulators (hides the global latency).
        auto tile_geom = [&](int t, bool& valid, long long& opix) {
          const int pr = C::P0 + t * 128 + q4 * 32 + lane;
          const int yl = pr / C::PW - 1, xl = pr % C::PW - 1;
          valid = (pr < C::P0 + C::MROWS) && xl >= 0 && xl < WIN_;
          int Yo = band * BR + yl, Xo = xl;
          if (UPS) { Yo = 2 * Yo + pa; Xo = 2 * Xo + pb; }
          opix = (long long)Yo * C::WOUT + Xo;
        };
#pragma unroll 1
        // work units: whole tiles over the two warp sets, or (tile, 32-channel chunk) pairs over four warp sets
        constexpr int NCH = HEAD ? 1 : NREAL / 32;
        constexpr bool SPLIT_CH = C::NTG > 2 && NCH > 1;
        for (int t = SPLIT_CH ? (tgrp / NCH) % (C::NTG / NCH) : tgrp; t < C::NTILES; t += SPLIT_CH ? C::NTG / NCH : C::NTG) {
          if (SPLIT) umma::mbar_wait_backoff(&bar_acc[set * 16 + t], ph_acc[set], kBackoffNs);
          else umma::mbar_wait(&bar_acc[set * 16 + t], ph_acc[set]);
          umma::fence_after_sync();
          if (t == 0) BPH(1);
          bool valid; long long opix;
          tile_geom(t, valid, opix);
          if constexpr (HEAD) {
            float v[8];
            umma::tmem_ld8(acc_addr + t * C::NACC, v);
            if constexpr (SPLIT) {
              float v2[8];
              umma::tmem_ld8(acc_addr + t * C::NACC + NOUT, v2);
              v[0] += v2[0];
            }
            if (valid) {
              int cls = p.classes[gslice];
              p.logits[((long long)b * p.T + cls) * (C::WOUT * C::WOUT) + opix] = v[0] + p.head_bias;
            }
          } else {
#pragma unroll
            for (int c0 = 0; c0 < NREAL; c0 += 32) {
              if (SPLIT_CH && (c0 / 32) != tgrp % NCH) continue;     // warp-uniform
              float v[32];
              float4 e4[UPS ? 8 : 1];                          // additive map (composed stages): issued before the TMEM load
              if constexpr (UPS) {
                // tile-ordered copy of the map: chunk-major inside a tile, so the 32 lanes (= 32 accumulator rows) of one load
                // read 512 contiguous bytes (one 128-byte row per THREAD cost 1.7 ms per step: 32 lines per instruction)
                const float* e = p.emap + ((((long long)b * C::NB + band) * 4 + pg) * C::NTILES + t) * (NREAL * 128) +
                                 ((c0 >> 2) * 128 + q4 * 32 + lane) * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) e4[i] = ld4(e + i * 512);
              }
              umma::tmem_ld32(acc_addr + t * C::NACC + c0, v);
              if constexpr (SPLIT) {                           // + Ah Wl, accumulated in the second column half
                float v2[32];
                umma::tmem_ld32(acc_addr + t * C::NACC + NOUT + c0, v2);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += v2[i];
              }
              if (valid) {
                if constexpr (UPS) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    v[4 * i] += e4[i].x; v[4 * i + 1] += e4[i].y; v[4 * i + 2] += e4[i].z; v[4 * i + 3] += e4[i].w;
                  }
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                  float s = 0.f, ss = 0.f;
#pragma unroll
                  for (int i = 0; i < 16; ++i) { float x = v[g * 16 + i]; s += x; ss = fmaf(x, x, ss); }
                  st_sum[c0 / 16 + g] += s; st_sq[c0 / 16 + g] += ss;
                }
                if constexpr (SPLIT) {
                  float* o = p.out32 + ((long long)sl * (C::WOUT * C::WOUT) + opix) * NREAL + c0;
#pragma unroll
                  for (int i = 0; i < 32; i += 4) st4(o + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
                } else {
                  __half* o = p.out + ((long long)sl * (C::WOUT * C::WOUT) + opix) * NREAL + c0;
#pragma unroll
                  for (int i = 0; i < 32; i += 8)
                    *reinterpret_cast<uint4*>(o + i) =
                        make_uint4(umma::pack_h2(v[i], v[i + 1]), umma::pack_h2(v[i + 2], v[i + 3]),
                                   umma::pack_h2(v[i + 4], v[i + 5]), umma::pack_h2(v[i + 6], v[i + 7]));
                }
              }
            }
          }
        }
      }
      ph_acc[set] ^= 1;
      umma::fence_before_sync();
      __syncthreads();        // accumulator set (and, after the last parity, the image) may be overwritten
      BPH(2);
      if (issuer && pg + SETS < C::NPG) {
        umma::fence_after_sync();
        issue_group(pg + SETS, set);
      }
    }
    // ---- output statistics of this band: warp shuffle -> per-warp slots -> fixed-order sum
    if (!HEAD) {
      if (!issuer) {
#pragma unroll
        for (int g = 0; g < C::GOUT; ++g) {
          float s = warp_sum(st_sum[g]), ss = warp_sum(st_sq[g]);
          if (lane == 0) { s_part[(warp * C::GOUT + g) * 2] = s; s_part[(warp * C::GOUT + g) * 2 + 1] = ss; }
        }
      }
      stats_item = it;      // summed and stored behind the next band's staging barrier (or after the loop)
    }
    BPH(3);
    ++nit_dbg;
  }
  if (p.dbg != nullptr && blockIdx.x == 0 && tid == 0) { p.dbg[0] = pacc0; p.dbg[1] = pacc1; p.dbg[2] = pacc2; p.dbg[3] = pacc3; p.dbg[4] = nit_dbg; p.dbg[5] = pacc4; }
  if (p.dbg != nullptr && blockIdx.x == 0 && tid == C::NWT) { p.dbg[6] = iacc_full; p.dbg[7] = iacc_issue; p.dbg[8] = iacc_empty; }
#undef BPH
  umma::fence_before_sync();
  __syncthreads();
  flush_stats();
  if (warp == 0) umma::tmem_dealloc<C::TMEM_COLS>(tm);
}

// GroupNorm coefficients of a producer's output, once per slice: fixed-order sum of the band partials, then
// scale = rstd * gamma, shift = beta - mean * scale (the band kernels used to redo this per band: ~2 K cycles each).
__global__ void gn_finalize_kernel(const float* __restrict__ stats, int nb, int cin, float count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ ss, int nslice) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nslice * cin) return;
  const int sl = i / cin, c = i % cin, g = c >> 4, G = cin / 16;
  float s = 0.f, q = 0.f;
  for (int b = 0; b < nb; ++b) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(stats + (((long long)sl * nb + b) * G + g) * 2));
    s += v.x; q += v.y;
  }
  const float mean = s / count;
  const float var = fmaxf(q / count - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float sc = rstd * gamma[c];
  reinterpret_cast<float2*>(ss)[i] = make_float2(sc, beta[c] - mean * sc);
}
static cudaError_t launch_gn_finalize(const float* stats, int nb, int cin, float count, const float* gamma, const float* beta,
                                      float* ss, int nslice, cudaStream_t st) {
  const int n = nslice * cin;
  gn_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(stats, nb, cin, count, gamma, beta, ss, nslice);
  return cudaGetLastError();
}

// Additive map E [B][Wout*Wout][NREAL] -> accumulator-tile order of the consuming composed stage (see BandConvParams::emap):
// rows that are halo columns or beyond the band are zero.
template <int CIN, int NOUT, int NREAL, bool UPS, bool IN_F32, int WIN_, int BR, bool HEAD, int CTAS, bool SPLIT>
__global__ void relayout_emap_kernel(const float* __restrict__ E, float* __restrict__ Et, int B) {
  using C = BandCfg<CIN, NOUT, NREAL, UPS, IN_F32, WIN_, BR, HEAD, CTAS, SPLIT>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;          // one float4 each
  const long long total = (long long)B * C::NB * 4 * C::NTILES * (NREAL / 4) * 128;
  if (i >= total) return;
  const int row = (int)(i % 128);
  long long r = i / 128;
  const int c = (int)(r % (NREAL / 4)); r /= (NREAL / 4);
  const int t = (int)(r % C::NTILES); r /= C::NTILES;
  const int pg = (int)(r % 4); r /= 4;
  const int band = (int)(r % C::NB);
  const int b = (int)(r / C::NB);
  const int pr = C::P0 + t * 128 + row;
  const int yl = pr / C::PW - 1, xl = pr % C::PW - 1;
  const bool valid = (pr < C::P0 + C::MROWS) && xl >= 0 && xl < WIN_;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
    const int Yo = 2 * (band * BR + yl) + (pg >> 1), Xo = 2 * xl + (pg & 1);
    v = ld4(E + ((long long)b * (C::WOUT * C::WOUT) + (long long)Yo * C::WOUT + Xo) * NREAL + c * 4);
  }
  st4(Et + i * 4, v);
}
template <int CIN, int NOUT, int NREAL, bool UPS, bool IN_F32, int WIN_, int BR, bool HEAD, int CTAS = 2, bool SPLIT = false>
static size_t emap_tiled_floats(int B) {
  using C = BandCfg<CIN, NOUT, NREAL, UPS, IN_F32, WIN_, BR, HEAD, CTAS, SPLIT>;
  return (size_t)B * C::NB * 4 * C::NTILES * NREAL * 128;
}
template <int CIN, int NOUT, int NREAL, bool UPS, bool IN_F32, int WIN_, int BR, bool HEAD, int CTAS = 2, bool SPLIT = false>
static cudaError_t launch_relayout_emap(const float* E, float* Et, int B, cudaStream_t st) {
  const long long total = (long long)emap_tiled_floats<CIN, NOUT, NREAL, UPS, IN_F32, WIN_, BR, HEAD, CTAS, SPLIT>(B) / 4;
  relayout_emap_kernel<CIN, NOUT, NREAL, UPS, IN_F32, WIN_, BR, HEAD, CTAS, SPLIT><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(E, Et, B);
  return cudaGetLastError();
}

template <int CIN, int NOUT, int NREAL, bool UPS, bool IN_F32, int WIN_, int BR, bool HEAD, int CTAS = 2, bool SPLIT = false>
static cudaError_t launch_band(const BandConvParams& p, int num_sms, cudaStream_t st) {
  using C = BandCfg<CIN, NOUT, NREAL, UPS, IN_F32, WIN_, BR, HEAD, CTAS, SPLIT>;
  auto kern = band_conv_kernel<CIN, NOUT, NREAL, UPS, IN_F32, WIN_, BR, HEAD, CTAS, SPLIT>;
  {   // per-device function attribute: set on every launch (cheap), a process-wide flag would miss other devices
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
  }
  long long nitems = (long long)p.nslice * C::NB;
  int grid = (int)(nitems < (long long)CTAS * num_sms ? nitems : (long long)CTAS * num_sms);
  if (grid <= 0) return cudaSuccess;
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("CATSEG_PHASE_TIMING");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on) { cudaMalloc(&dbg, 16 * sizeof(long long)); cudaMemset(dbg, 0, 16 * sizeof(long long)); }
    if (const char* b = getenv("CATSEG_DEC_BACKOFF_NS")) { unsigned v = (unsigned)atoi(b); cudaMemcpyToSymbol(kBackoffNs, &v, sizeof(v)); }
  }
  BandConvParams q = p;
  q.dbg = dbg_on ? dbg : nullptr;
  kern<<<grid, C::THREADS, C::SMEM, st>>>(q);
  if (dbg_on) {
    long long hb[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    double n = hb[4] > 0 ? (double)hb[4] : 1.0;
    if (!C::RESIDENT) fprintf(stderr, "[issuing warp, per band: wait-weights %.0f | issue %.0f | wait-prev-image+refill %.0f] ", hb[6] / n, hb[7] / n, hb[8] / n);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, C::SMEM);
    fprintf(stderr, "[occupancy %d CTAs/SM, smem %u] ", occ, (unsigned)C::SMEM);
    fprintf(stderr, "[band_conv<%d,%d,ups=%d,W=%d> cycles/band over %lld bands, thread 0 (issuer)] stage %.0f | mma issue+wait %.0f (x%d parity groups) "
            "| epilogue %.0f | stats %.0f | (GN params of the input %.0f, before 'stage')\n", CIN, NOUT, (int)UPS, WIN_, hb[4], hb[0] / n, hb[1] / n, C::NPG, hb[2] / n, hb[3] / n, hb[5] / n);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// PRECISE head: conv3x3 32 -> 1 (+ bias) on relu(GroupNorm(c2b)), fp32 on the CUDA cores (model.py:634, 679).
// One output channel makes this stage a 288-term dot product per pixel: on the tensor pipe (band_conv_kernel<.., HEAD>) it
// pays for an N = 16 MMA tile plus the hi/lo operand pairs and ran at 0.85 ms per 1024 slices; as an fp32 stencil it is bound
// by reading c2b once from HBM.  CTA = (slice, band of 6 rows): the band + halo is staged normalised in shared memory
// ([8 rows][98 cols][32 ch] fp32, 16-byte chunks XOR-swizzled by bit 3 of the column so that the two pixel groups of an
// LDS.128 phase hit disjoint banks); thread = 8 pixels x 8 channels with its 72 weights in registers, four threads reduce.
namespace {
constexpr int HB_BR = 6, HB_W = 96, HB_C = 32, HB_PW = 98, HB_THREADS = 288, HB_NB = HB_W / HB_BR;
constexpr int HB_SMEM = (HB_BR + 2) * HB_PW * HB_C * 4;
}  // namespace

__global__ void __launch_bounds__(HB_THREADS, 2)
head_conv_f32_kernel(const float* __restrict__ in, const float* __restrict__ in_ss, const float* __restrict__ wt, float bias,
                     const int32_t* __restrict__ classes, float* __restrict__ logits, int slice0, int nslice, int Te, int T,
                     PeerPtrs lpeers, int nlp) {
  extern __shared__ float4 hb_sm[];                    // [(BR + 2) * PW pixels][8 chunks of 4 channels]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cg = lane & 3, G = warp * 8 + (lane >> 2);   // channel group (8 channels), pixel group (8 pixels of one row)
  const int ry = G / 12, x0 = (G % 12) * 8;
  float w[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) w[t][k] = __ldg(wt + t * HB_C + cg * 8 + k);
  const int q = tid & 7;                               // staging: this thread always handles channel chunk q (288 % 8 == 0)
  const long long nitems = (long long)nslice * HB_NB;
  for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int sl = (int)(it / HB_NB), band = (int)(it % HB_NB);
    const int y_first = band * HB_BR - 1;
    {
      // raw rows -> shared memory with cp.async (every chunk of the band in flight at once: a register-staged loop kept one
      // load per thread in flight and was latency bound), then GroupNorm + ReLU in place; conv padding is zero AFTER the ReLU
      const float4* ss = reinterpret_cast<const float4*>(in_ss + ((long long)sl * HB_C + q * 4) * 2);
      const float4 s01 = __ldg(ss), s23 = __ldg(ss + 1);                 // (scale, shift) of channels 4q .. 4q+3
      // chunk i = tid + 288 k  <->  pixel p = (tid >> 3) + 36 k: (row, column) advance incrementally, no division in the loop
      constexpr int NPIX = (HB_BR + 2) * HB_PW, STEP = HB_THREADS / 8;
      const int p0 = tid >> 3;
      const float* src_sl = in + (long long)sl * HB_W * HB_W * HB_C + q * 4;
      {
        int r = p0 / HB_PW, c = p0 - r * HB_PW;
        for (int p = p0; p < NPIX; p += STEP) {
          const int y = y_first + r, x = c - 1;
          if ((unsigned)y < (unsigned)HB_W && (unsigned)x < (unsigned)HB_W) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&hb_sm[p * 8 + (q ^ ((c >> 3) & 1))]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src_sl + (y * HB_W + x) * HB_C) : "memory");
          }
          c += STEP;
          if (c >= HB_PW) { c -= HB_PW; ++r; }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      {
        int r = p0 / HB_PW, c = p0 - r * HB_PW;
        for (int p = p0; p < NPIX; p += STEP) {                          // the same thread owns the same chunks: no barrier needed
          const int y = y_first + r, x = c - 1;
          float4* cell = &hb_sm[p * 8 + (q ^ ((c >> 3) & 1))];
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((unsigned)y < (unsigned)HB_W && (unsigned)x < (unsigned)HB_W) {
            v = *cell;
            v.x = fmaxf(fmaf(v.x, s01.x, s01.y), 0.0f); v.y = fmaxf(fmaf(v.y, s01.z, s01.w), 0.0f);
            v.z = fmaxf(fmaf(v.z, s23.x, s23.y), 0.0f); v.w = fmaxf(fmaf(v.w, s23.z, s23.w), 0.0f);
          }
          *cell = v;
          c += STEP;
          if (c >= HB_PW) { c -= HB_PW; ++r; }
        }
      }
    }
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int tr = 0; tr < 3; ++tr) {
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        const int col = x0 + j, p = (ry + tr) * HB_PW + col, s = (col >> 3) & 1;
        const float4 a = hb_sm[p * 8 + ((2 * cg) ^ s)], b = hb_sm[p * 8 + ((2 * cg + 1) ^ s)];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int px = j - kx;                       // output pixel whose tap (tr, kx) reads padded column x0 + j
          if (px >= 0 && px < 8) {
            const float* ww = w[tr * 3 + kx];
            float d = acc[px];
            d = fmaf(a.x, ww[0], d); d = fmaf(a.y, ww[1], d); d = fmaf(a.z, ww[2], d); d = fmaf(a.w, ww[3], d);
            d = fmaf(b.x, ww[4], d); d = fmaf(b.y, ww[5], d); d = fmaf(b.z, ww[6], d); d = fmaf(b.w, ww[7], d);
            acc[px] = d;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 2);
    }
    if (cg == 0) {
      const int gs = slice0 + sl, b = gs / Te, cls = classes[gs];
      const long long off = ((long long)b * T + cls) * (HB_W * HB_W) + (band * HB_BR + ry) * HB_W + x0;
      const float4 o0 = make_float4(acc[0] + bias, acc[1] + bias, acc[2] + bias, acc[3] + bias);
      const float4 o1 = make_float4(acc[4] + bias, acc[5] + bias, acc[6] + bias, acc[7] + bias);
      if (nlp == 0) {
        st4(logits + off, o0); st4(logits + off + 4, o1);
      } else {                                           // class-sharded, peer-direct: the plane goes to every rank's full buffer
        for (int r = 0; r < nlp; ++r) { st4(lpeers.p[r] + off, o0); st4(lpeers.p[r] + off + 4, o1); }
      }
    }
    __syncthreads();                                   // the band image is rebuilt by the next item
  }
}

static cudaError_t launch_head_conv_f32(const float* in, const float* in_ss, const float* wt, float bias, const int32_t* classes,
                                        float* logits, int slice0, int nslice, int Te, int T, int num_sms, const PeerPtrs* lpeers,
                                        int nlp, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(head_conv_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HB_SMEM);
  if (e != cudaSuccess) return e;
  const long long nitems = (long long)nslice * HB_NB;
  const int grid = (int)(nitems < 2LL * num_sms ? nitems : 2LL * num_sms);
  if (grid <= 0) return cudaSuccess;
  PeerPtrs lp{};
  if (lpeers != nullptr) lp = *lpeers;
  head_conv_f32_kernel<<<grid, HB_THREADS, HB_SMEM, st>>>(in, in_ss, wt, bias, classes, logits, slice0, nslice, Te, T, lp,
                                                          lpeers != nullptr ? nlp : 0);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// weight preparation (runs once in catseg_finalize_params)

// image[(k/8)*NOUT*8 + n*8 + k%8] = W3[n][ci0 + k][tap]   (n < nreal, else 0)
// Element (n, k) of weight image `img` with `nout` output rows: nw = 1: one fp16 image [nout x cin]; nw = 2 (PRECISE): one
// image with 2 nout rows, rows [0, nout) = hi term, [nout, 2 nout) = lo term (BandCfg::WBYTES).
__device__ __forceinline__ void store_wimg(__half* dst, long long img, int nout, int cin, int n, int k, float v, int nw) {
  const __half h = __float2half_rn(v);
  __half* base = dst + img * (long long)(nw * nout * cin) + (k >> 3) * (nw * nout * 8) + (k & 7);
  base[n * 8] = h;
  if (nw == 2) base[(nout + n) * 8] = __float2half_rn(v - __half2float(h));
}
__global__ void pack_tap_img_kernel(__half* dst, const float* W3, int Cin3, int ci0, int CIN, int NOUT, int nreal, int nw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int per = NOUT * CIN;
  if (i >= 9 * per) return;
  int tap = i / per, r = i % per, n = r / CIN, k = r % CIN;
  float v = n < nreal ? W3[((long long)n * Cin3 + ci0 + k) * 9 + tap] : 0.0f;
  store_wimg(dst, tap, NOUT, CIN, n, k, v, nw);
}

// composed ConvTranspose(k2,s2) o conv3x3:  image index (a*2+b)*4 + (u*2+v), element [co][ci]
//   Wc = sum_{dy,dx -> (u,v)} sum_cu Wup[ci][cu][a'][b'] * W3[co][cu][dy+1][dx+1]
__global__ void compose_up_img_kernel(__half* dst, const float* Wup, const float* W3, int Ci, int Cup, int Cin3,
                                      int Co, int nw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int per = Co * Ci;
  if (i >= 16 * per) return;
  int img = i / per, r = i % per, co = r / Ci, ci = r % Ci;
  int a = img >> 3, bq = (img >> 2) & 1, u = (img >> 1) & 1, v = img & 1;
  float acc = 0.0f;
  for (int dy = -1; dy <= 1; ++dy) {
    int ty = a + dy, fy = ty < 0 ? -1 : ty >> 1, ay = ty & 1;
    if (fy - (a - 1) != u) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      int tx = bq + dx, fx = tx < 0 ? -1 : tx >> 1, ax = tx & 1;
      if (fx - (bq - 1) != v) continue;
      for (int cu = 0; cu < Cup; ++cu)
        acc = fmaf(Wup[(((long long)ci * Cup + cu) * 2 + ay) * 2 + ax], W3[((long long)co * Cin3 + cu) * 9 + (dy + 1) * 3 + dx + 1], acc);
    }
  }
  store_wimg(dst, img, Co, Ci, co, ci, acc, nw);
}

// bias map of the composed conv: Bmap[Y][X][co] = sum_{valid dy,dx} sum_cu bup[cu] W3[co][cu][dy+1][dx+1]
__global__ void up_bias_map_kernel(float* dst, const float* bup, const float* W3, int Cup, int Cin3, int Co, int Wout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Wout * Wout * Co) return;
  int co = i % Co, pix = i / Co, Y = pix / Wout, X = pix % Wout;
  float acc = 0.0f;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      if (Y + dy < 0 || Y + dy >= Wout || X + dx < 0 || X + dx >= Wout) continue;
      for (int cu = 0; cu < Cup; ++cu) acc = fmaf(bup[cu], W3[((long long)co * Cin3 + cu) * 9 + (dy + 1) * 3 + dx + 1], acc);
    }
  dst[i] = acc;
}

// guidance part of the first conv of an Up block, packed for the fp32 implicit GEMM: Wg[(tap*Cg + cg)][co]
__global__ void pack_guid_w_kernel(float* dst, const float* W3, int Cup, int Cg, int Cin3, int Co) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * Cg * Co) return;
  int co = i % Co, r = i / Co, cg = r % Cg, tap = r / Cg;
  dst[i] = W3[((long long)co * Cin3 + Cup + cg) * 9 + tap];
}

size_t decoder_fast_weight_bytes(const DecoderDims& d, int nw) {
  size_t b = 0;
  b += (size_t)16 * d.D1 * d.C0 * 2 * nw;      // D1 composed
  b += (size_t)9 * d.D1 * d.D1 * 2 * nw;       // D2
  b += (size_t)16 * d.D2 * d.D1 * 2 * nw;      // D3 composed
  b += (size_t)9 * d.D2 * d.D2 * 2 * nw;       // D4
  b += (size_t)9 * 16 * d.D2 * 2 * nw;         // D5 head (N padded to 16)
  b += (size_t)(4 * d.H * d.W * d.D1 + 16 * d.H * d.W * d.D2) * 4;   // bias maps
  b += (size_t)(9 * d.G1 * d.D1 + 9 * d.G2 * d.D2) * 4;              // guidance conv weights
  return (b + 255) / 256 * 256;
}

cudaError_t decoder_fast_pack(const DecoderDims& d, const float* up1_w, const float* up1_b, const float* c1a_w,
                              const float* c1b_w, const float* up2_w, const float* up2_b, const float* c2a_w,
                              const float* c2b_w, const float* head_w, void* storage, DecoderFastW* out, int nw,
                              cudaStream_t st) {
  if (d.C0 != 128 || d.D1 != 64 || d.D2 != 32 || d.H != 24 || d.W != 24 || d.G1 % 4 || d.G2 % 4)
    return cudaErrorInvalidValue;      // the band kernels are instantiated for the shipped decoder geometry
  uint8_t* ptr = reinterpret_cast<uint8_t*>(storage);
  auto take = [&](size_t bytes) { uint8_t* r = ptr; ptr += bytes; return r; };
  __half* w1 = reinterpret_cast<__half*>(take((size_t)16 * d.D1 * d.C0 * 2 * nw));
  __half* w2 = reinterpret_cast<__half*>(take((size_t)9 * d.D1 * d.D1 * 2 * nw));
  __half* w3 = reinterpret_cast<__half*>(take((size_t)16 * d.D2 * d.D1 * 2 * nw));
  __half* w4 = reinterpret_cast<__half*>(take((size_t)9 * d.D2 * d.D2 * 2 * nw));
  __half* w5 = reinterpret_cast<__half*>(take((size_t)9 * 16 * d.D2 * 2 * nw));
  float* bm1 = reinterpret_cast<float*>(take((size_t)4 * d.H * d.W * d.D1 * 4));
  float* bm2 = reinterpret_cast<float*>(take((size_t)16 * d.H * d.W * d.D2 * 4));
  float* wg1 = reinterpret_cast<float*>(take((size_t)9 * d.G1 * d.D1 * 4));
  float* wg2 = reinterpret_cast<float*>(take((size_t)9 * d.G2 * d.D2 * 4));
  auto blocks = [](long long n) { return (unsigned)((n + 255) / 256); };
  compose_up_img_kernel<<<blocks(16LL * d.D1 * d.C0), 256, 0, st>>>(w1, up1_w, c1a_w, d.C0, d.U1, d.U1 + d.G1, d.D1, nw);
  pack_tap_img_kernel<<<blocks(9LL * d.D1 * d.D1), 256, 0, st>>>(w2, c1b_w, d.D1, 0, d.D1, d.D1, d.D1, nw);
  compose_up_img_kernel<<<blocks(16LL * d.D2 * d.D1), 256, 0, st>>>(w3, up2_w, c2a_w, d.D1, d.U2, d.U2 + d.G2, d.D2, nw);
  pack_tap_img_kernel<<<blocks(9LL * d.D2 * d.D2), 256, 0, st>>>(w4, c2b_w, d.D2, 0, d.D2, d.D2, d.D2, nw);
  pack_tap_img_kernel<<<blocks(9LL * 16 * d.D2), 256, 0, st>>>(w5, head_w, d.D2, 0, d.D2, 16, 1, nw);
  up_bias_map_kernel<<<blocks(4LL * d.H * d.W * d.D1), 256, 0, st>>>(bm1, up1_b, c1a_w, d.U1, d.U1 + d.G1, d.D1, 2 * d.W);
  up_bias_map_kernel<<<blocks(16LL * d.H * d.W * d.D2), 256, 0, st>>>(bm2, up2_b, c2a_w, d.U2, d.U2 + d.G2, d.D2, 4 * d.W);
  pack_guid_w_kernel<<<blocks(9LL * d.G1 * d.D1), 256, 0, st>>>(wg1, c1a_w, d.U1, d.G1, d.U1 + d.G1, d.D1);
  pack_guid_w_kernel<<<blocks(9LL * d.G2 * d.D2), 256, 0, st>>>(wg2, c2a_w, d.U2, d.G2, d.U2 + d.G2, d.D2);
  out->w1 = w1; out->w2 = w2; out->w3 = w3; out->w4 = w4; out->w5 = w5;
  out->bmap1 = bm1; out->bmap2 = bm2; out->wg1 = wg1; out->wg2 = wg2;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// per-image additive maps  E = Bmap + conv3x3(projected guidance, Wg)   (fp32 implicit GEMM)
struct GuidConvA {
  static constexpr bool kMFastest = false;
  const float* g; int Cg, H, W;
  __device__ float operator()(int, int m, int k) const {
    int hw = H * W, img = m / hw, pix = m - img * hw;
    int tap = k / Cg, cg = k - tap * Cg;
    int y = pix / W + tap / 3 - 1, x = pix % W + tap % 3 - 1;
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.0f;
    return __ldg(g + ((long long)img * hw + y * W + x) * Cg + cg);
  }
};
struct MapAddStore {
  float* out; const float* bmap; int hw, N;
  __device__ void operator()(int, int m, int n, float acc) const {
    out[(long long)m * N + n] = acc + __ldg(bmap + (long long)(m % hw) * N + n);
  }
};

// Stage configurations <CIN, NOUT, NREAL, UPS, IN_F32, WIN, BR, HEAD, CTAS>: N = two 9-warp CTAs per SM were intended (but only
// one fits the register file, see BandCfg::NWW), W = one 17-warp CTA per SM with bands twice as tall.  CATSEG_DEC_WIDE is a
// bit mask (bit i = stage D(i+1) uses W), read once; it exists for A/B measurements.
#define D1N 128, 64, 64, true, true, 24, 6, false, 2
#define D1W 128, 64, 64, true, true, 24, 12, false, 1
#define D2N 64, 64, 64, false, false, 48, 6, false, 2
#define D2W 64, 64, 64, false, false, 48, 12, false, 1
#define D3N 64, 32, 32, true, false, 48, 4, false, 2
#define D3W 64, 32, 32, true, false, 48, 16, false, 1
#define D4N 32, 32, 32, false, false, 96, 6, false, 2
#define D4W 32, 32, 32, false, false, 96, 12, false, 1
#define D5N 32, 16, 16, false, false, 96, 6, true, 2
#define D5W 32, 16, 16, false, false, 96, 12, true, 1
// PRECISE shapes: one 17-warp CTA per SM, fp32 activations in and out
#define D1S 128, 64, 64, true, true, 24, 8, false, 1, true
#define D2S 64, 64, 64, false, true, 48, 8, false, 1, true
#define D3S 64, 32, 32, true, true, 48, 8, false, 1, true
#define D4S 32, 32, 32, false, true, 96, 8, false, 1, true
#define D5S 32, 16, 16, false, true, 96, 6, true, 1, true
// ... and two 9-warp CTAs per SM with short bands (one CTA's staging / epilogue under the other's MMAs); CATSEG_DECS_NARROW is
// a bit mask (bit i = stage D(i+1) uses the narrow shape), read once, for A/B measurements
#define D2T 64, 64, 64, false, true, 48, 2, false, 2, true
#define D3T 64, 32, 32, true, true, 48, 2, false, 2, true
#define D4T 32, 32, 32, false, true, 96, 3, false, 2, true
#define D5T 32, 16, 16, false, true, 96, 3, true, 2, true
static int decs_narrow_mask() {
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("CATSEG_DECS_NARROW");
    m = e ? atoi(e) & 28 : 0;
  }
  return m;
}
static int dec_wide_mask() {
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("CATSEG_DEC_WIDE");
    m = e ? atoi(e) & 31 : 6;
  }
  return m;
}
static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

size_t decoder_fast_scratch_bytes(const DecoderDims& d, int B, int chunk) {
  size_t hw = (size_t)d.H * d.W;
  size_t b = 0;
  b += (size_t)B * (4 * hw * d.D1 + 16 * hw * d.D2) * 4;                 // E1, E2
  b += (max_sz(emap_tiled_floats<D1N>(B), emap_tiled_floats<D1W>(B)) + max_sz(emap_tiled_floats<D3N>(B), emap_tiled_floats<D3W>(B))) * 4 + 512;   // tile-ordered copies
  b += (size_t)chunk * (2 * 4 * hw * d.D1 + 2 * 16 * hw * d.D2) * 2;     // c1a c1b c2a c2b (fp16)
  b += (size_t)chunk * (4 * 4 + 8 * 4 + 12 * 2 + 16 * 2) * 2 * 4 + 4096; // band statistics
  b += (size_t)chunk * 64 * 2 * 4 + 256;                                  // GroupNorm (scale, shift) of the current producer
  return (b + 255) / 256 * 256;
}

size_t decoder_split_scratch_bytes(const DecoderDims& d, int B, int chunk) {
  size_t hw = (size_t)d.H * d.W;
  size_t b = 0;
  b += (size_t)B * (4 * hw * d.D1 + 16 * hw * d.D2) * 4;                 // E1, E2
  b += (emap_tiled_floats<D1S>(B) + max_sz(emap_tiled_floats<D3S>(B), emap_tiled_floats<D3T>(B))) * 4 + 512;
  b += (size_t)chunk * (2 * 4 * hw * d.D1 + 2 * 16 * hw * d.D2) * 4;     // c1a c1b c2a c2b (fp32)
  b += (size_t)chunk * (4 * 4 + 24 * 4 + 24 * 2 + 32 * 2) * 2 * 4 + 4096; // band statistics (max NB * G per stage over both shapes)
  b += (size_t)chunk * 64 * 2 * 4 + 256;
  return (b + 255) / 256 * 256;
}

#define CKF(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; ++nl; } while (0)

cudaError_t run_decoder_fast(const float* X, const float* dg0, const float* dg1, const int32_t* classes,
                             float* logits, int B, int T, int Te, const DecoderDims& d, const DecoderFastW& w,
                             const DecoderW& wx, float head_bias, void* scratch, int chunk, int num_sms,
                             int* launches, cudaStream_t st) {
  const int hw = d.H * d.W;
  int nl = 0;
  const int wide = dec_wide_mask();
  uint8_t* ptr = reinterpret_cast<uint8_t*>(scratch);
  auto take = [&](size_t bytes) { uint8_t* r = ptr; ptr += (bytes + 255) / 256 * 256; return r; };
  float* E1 = reinterpret_cast<float*>(take((size_t)B * 4 * hw * d.D1 * 4));
  float* E2 = reinterpret_cast<float*>(take((size_t)B * 16 * hw * d.D2 * 4));
  float* E1t = reinterpret_cast<float*>(take(max_sz(emap_tiled_floats<D1N>(B), emap_tiled_floats<D1W>(B)) * 4));
  float* E2t = reinterpret_cast<float*>(take(max_sz(emap_tiled_floats<D3N>(B), emap_tiled_floats<D3W>(B)) * 4));
  __half* c1a = reinterpret_cast<__half*>(take((size_t)chunk * 4 * hw * d.D1 * 2));
  __half* c1b = reinterpret_cast<__half*>(take((size_t)chunk * 4 * hw * d.D1 * 2));
  __half* c2a = reinterpret_cast<__half*>(take((size_t)chunk * 16 * hw * d.D2 * 2));
  __half* c2b = reinterpret_cast<__half*>(take((size_t)chunk * 16 * hw * d.D2 * 2));
  float* s1a = reinterpret_cast<float*>(take((size_t)chunk * 4 * 4 * 2 * 4));    // NB=4 G=4
  float* s1b = reinterpret_cast<float*>(take((size_t)chunk * 8 * 4 * 2 * 4));    // NB=8 G=4
  float* s2a = reinterpret_cast<float*>(take((size_t)chunk * 12 * 2 * 2 * 4));   // NB=12 G=2
  float* s2b = reinterpret_cast<float*>(take((size_t)chunk * 16 * 2 * 2 * 4));   // NB=16 G=2
  float* gss = reinterpret_cast<float*>(take((size_t)chunk * 64 * 2 * 4));       // (scale, shift) [chunk][<=64][2]

  {
    GuidConvA a{dg0, d.G1, 2 * d.H, 2 * d.W};
    CKF(launch_igemm(a, w.wg1, 0, 1, B * 4 * hw, d.D1, 9 * d.G1, MapAddStore{E1, w.bmap1, 4 * hw, d.D1}, st));
    GuidConvA a2{dg1, d.G2, 4 * d.H, 4 * d.W};
    CKF(launch_igemm(a2, w.wg2, 0, 1, B * 16 * hw, d.D2, 9 * d.G2, MapAddStore{E2, w.bmap2, 16 * hw, d.D2}, st));
    if (wide & 1) CKF((launch_relayout_emap<D1W>(E1, E1t, B, st))); else CKF((launch_relayout_emap<D1N>(E1, E1t, B, st)));
    if (wide & 4) CKF((launch_relayout_emap<D3W>(E2, E2t, B, st))); else CKF((launch_relayout_emap<D3N>(E2, E2t, B, st)));
  }
  const int nslice = B * Te;
  for (int s0 = 0; s0 < nslice; s0 += chunk) {
    const int n = nslice - s0 < chunk ? nslice - s0 : chunk;
    BandConvParams p{};
    p.Te = Te; p.slice0 = s0; p.nslice = n; p.T = T; p.classes = classes; p.logits = logits; p.head_bias = head_bias;
    // D1: x (24^2 x 128, fp32) -> c1a (48^2 x 64), composed transposed conv + conv
    p.in = X + (long long)s0 * hw * d.C0; p.in_stats = nullptr; p.wimg = w.w1; p.emap = E1t; p.out = c1a; p.out_stats = s1a;
    if (wide & 1) CKF((launch_band<D1W>(p, num_sms, st))); else CKF((launch_band<D1N>(p, num_sms, st)));
    const int nb1 = (wide & 1) ? BandCfg<D1W>::NB : BandCfg<D1N>::NB;
    // D2: c1a -> c1b, 3x3 64 -> 64 on relu(gn(c1a))
    CKF(launch_gn_finalize(s1a, nb1, 64, (float)(4 * hw * 16), wx.gn1a_g, wx.gn1a_b, gss, n, st));
    p.in = c1a; p.in_stats = s1a; p.in_ss = gss; p.nb_in = nb1; p.in_count = (float)(4 * hw * 16); p.gamma = wx.gn1a_g; p.beta = wx.gn1a_b;
    p.wimg = w.w2; p.emap = nullptr; p.out = c1b; p.out_stats = s1b;
    if (wide & 2) CKF((launch_band<D2W>(p, num_sms, st))); else CKF((launch_band<D2N>(p, num_sms, st)));
    const int nb2 = (wide & 2) ? BandCfg<D2W>::NB : BandCfg<D2N>::NB;
    // D3: c1b -> c2a (96^2 x 32), composed
    CKF(launch_gn_finalize(s1b, nb2, 64, (float)(4 * hw * 16), wx.gn1b_g, wx.gn1b_b, gss, n, st));
    p.in = c1b; p.in_stats = s1b; p.nb_in = nb2; p.gamma = wx.gn1b_g; p.beta = wx.gn1b_b;
    p.wimg = w.w3; p.emap = E2t; p.out = c2a; p.out_stats = s2a;
    if (wide & 4) CKF((launch_band<D3W>(p, num_sms, st))); else CKF((launch_band<D3N>(p, num_sms, st)));   // the 64 KiB weight set stays resident
    const int nb3 = (wide & 4) ? BandCfg<D3W>::NB : BandCfg<D3N>::NB;
    // D4: c2a -> c2b, 3x3 32 -> 32
    CKF(launch_gn_finalize(s2a, nb3, 32, (float)(16 * hw * 16), wx.gn2a_g, wx.gn2a_b, gss, n, st));
    p.in = c2a; p.in_stats = s2a; p.nb_in = nb3; p.in_count = (float)(16 * hw * 16); p.gamma = wx.gn2a_g; p.beta = wx.gn2a_b;
    p.wimg = w.w4; p.emap = nullptr; p.out = c2b; p.out_stats = s2b;
    if (wide & 8) CKF((launch_band<D4W>(p, num_sms, st))); else CKF((launch_band<D4N>(p, num_sms, st)));
    const int nb4 = (wide & 8) ? BandCfg<D4W>::NB : BandCfg<D4N>::NB;
    // D5: head 3x3 32 -> 1 (+ bias), scattered to logits[b][class]
    CKF(launch_gn_finalize(s2b, nb4, 32, (float)(16 * hw * 16), wx.gn2b_g, wx.gn2b_b, gss, n, st));
    p.in = c2b; p.in_stats = s2b; p.nb_in = nb4; p.gamma = wx.gn2b_g; p.beta = wx.gn2b_b;
    p.wimg = w.w5; p.emap = nullptr; p.out = nullptr; p.out_stats = nullptr;
    if (wide & 16) CKF((launch_band<D5W>(p, num_sms, st))); else CKF((launch_band<D5N>(p, num_sms, st)));
  }
  if (launches) *launches += nl;
  return cudaSuccess;
}

// Per-image additive maps of the PRECISE decoder (guidance convolution + transposed-conv bias, tile ordered): they depend
// on the decoder guidance only, so the caller computes them on its internal stream beside the aggregation layers.  The
// map area [E1 | E2 | E1t | E2t] (laid out for B images, image major) is the head of the decoder scratch, or -- class split
// with the guidance sharded by image -- a peer-visible buffer: then only images [b0, b0 + Bl) are computed here (dg0 / dg1
// point at image b0) and the E1t / E2t slices of the other images arrive from their owners.
namespace {
struct EmapArea { float *E1, *E2, *E1t, *E2t; size_t bytes, e1t_off, e2t_off, per1, per2; };
EmapArea emap_area(void* base, int B, const DecoderDims& d) {
  const int hw = d.H * d.W;
  const int narrow = decs_narrow_mask();
  uint8_t* ptr = reinterpret_cast<uint8_t*>(base);
  uint8_t* p0 = ptr;
  auto take = [&](size_t bytes) { uint8_t* r = ptr; ptr += (bytes + 255) / 256 * 256; return r; };
  EmapArea a{};
  a.E1 = reinterpret_cast<float*>(take((size_t)B * 4 * hw * d.D1 * 4));
  a.E2 = reinterpret_cast<float*>(take((size_t)B * 16 * hw * d.D2 * 4));
  a.E1t = reinterpret_cast<float*>(take(emap_tiled_floats<D1S>(B) * 4));
  a.E2t = reinterpret_cast<float*>(take(max_sz(emap_tiled_floats<D3S>(B), emap_tiled_floats<D3T>(B)) * 4));
  a.bytes = (size_t)(ptr - p0);
  a.e1t_off = (size_t)(reinterpret_cast<uint8_t*>(a.E1t) - p0) / 4;
  a.e2t_off = (size_t)(reinterpret_cast<uint8_t*>(a.E2t) - p0) / 4;
  a.per1 = emap_tiled_floats<D1S>(1);
  a.per2 = (narrow & 4) ? emap_tiled_floats<D3T>(1) : emap_tiled_floats<D3S>(1);
  return a;
}
}  // namespace

size_t decoder_split_emap_bytes(const DecoderDims& d, int B) { return emap_area(nullptr, B, d).bytes; }
void decoder_split_emap_slices(const DecoderDims& d, int B, size_t* e1t_off, size_t* e2t_off, size_t* per1, size_t* per2) {
  const EmapArea a = emap_area(nullptr, B, d);
  *e1t_off = a.e1t_off; *e2t_off = a.e2t_off; *per1 = a.per1; *per2 = a.per2;
}

cudaError_t decoder_split_prepare(const float* dg0, const float* dg1, int B, int b0, int Bl, const DecoderDims& d,
                                  const DecoderFastW& w, void* area, int* launches, cudaStream_t st) {
  const int hw = d.H * d.W;
  int nl = 0;
  const int narrow = decs_narrow_mask();
  const EmapArea a = emap_area(area, B, d);
  float* E1 = a.E1 + (size_t)b0 * 4 * hw * d.D1;
  float* E2 = a.E2 + (size_t)b0 * 16 * hw * d.D2;
  GuidConvA ga{dg0, d.G1, 2 * d.H, 2 * d.W};
  CKF(launch_igemm(ga, w.wg1, 0, 1, Bl * 4 * hw, d.D1, 9 * d.G1, MapAddStore{E1, w.bmap1, 4 * hw, d.D1}, st));
  GuidConvA ga2{dg1, d.G2, 4 * d.H, 4 * d.W};
  CKF(launch_igemm(ga2, w.wg2, 0, 1, Bl * 16 * hw, d.D2, 9 * d.G2, MapAddStore{E2, w.bmap2, 16 * hw, d.D2}, st));
  CKF((launch_relayout_emap<D1S>(E1, a.E1t + (size_t)b0 * a.per1, Bl, st)));
  if (narrow & 4) CKF((launch_relayout_emap<D3T>(E2, a.E2t + (size_t)b0 * a.per2, Bl, st)));
  else CKF((launch_relayout_emap<D3S>(E2, a.E2t + (size_t)b0 * a.per2, Bl, st)));
  if (launches) *launches += nl;
  return cudaSuccess;
}

// PRECISE decoder: the same five band-convolution stages with hi + lo fp16 operand pairs and fp32 intermediates
// (decoder_split_prepare must have run on the same scratch area)
cudaError_t run_decoder_split(const float* X, const float* dg0, const float* dg1, const int32_t* classes,
                              float* logits, int B, int T, int Te, const DecoderDims& d, const DecoderFastW& w,
                              const DecoderW& wx, float head_bias, void* scratch, int chunk, int num_sms,
                              int* launches, const PeerPtrs* lpeers, int nlp, const void* ext_area, cudaStream_t st) {
  const int hw = d.H * d.W;
  int nl = 0;
  const int narrow = decs_narrow_mask();
  uint8_t* ptr = reinterpret_cast<uint8_t*>(scratch);
  auto take = [&](size_t bytes) { uint8_t* r = ptr; ptr += (bytes + 255) / 256 * 256; return r; };
  // the additive maps were computed by decoder_split_prepare into the head of the scratch area or into `ext_area`
  const EmapArea ea = emap_area(ext_area != nullptr ? const_cast<void*>(ext_area) : scratch, B, d);
  ptr += ea.bytes;
  const float* E1t = ea.E1t;
  const float* E2t = ea.E2t;
  float* c1a = reinterpret_cast<float*>(take((size_t)chunk * 4 * hw * d.D1 * 4));
  float* c1b = reinterpret_cast<float*>(take((size_t)chunk * 4 * hw * d.D1 * 4));
  float* c2a = reinterpret_cast<float*>(take((size_t)chunk * 16 * hw * d.D2 * 4));
  float* c2b = reinterpret_cast<float*>(take((size_t)chunk * 16 * hw * d.D2 * 4));
  const int nb1 = BandCfg<D1S>::NB;
  const int nb2 = (narrow & 2) ? BandCfg<D2T>::NB : BandCfg<D2S>::NB;
  const int nb3 = (narrow & 4) ? BandCfg<D3T>::NB : BandCfg<D3S>::NB;
  const int nb4 = (narrow & 8) ? BandCfg<D4T>::NB : BandCfg<D4S>::NB;
  float* s1a = reinterpret_cast<float*>(take((size_t)chunk * nb1 * 4 * 2 * 4));
  float* s1b = reinterpret_cast<float*>(take((size_t)chunk * nb2 * 4 * 2 * 4));
  float* s2a = reinterpret_cast<float*>(take((size_t)chunk * nb3 * 2 * 2 * 4));
  float* s2b = reinterpret_cast<float*>(take((size_t)chunk * nb4 * 2 * 2 * 4));
  float* gss = reinterpret_cast<float*>(take((size_t)chunk * 64 * 2 * 4));
  (void)dg0; (void)dg1;
  const int nslice = B * Te;
  for (int s0 = 0; s0 < nslice; s0 += chunk) {
    const int n = nslice - s0 < chunk ? nslice - s0 : chunk;
    BandConvParams p{};
    p.Te = Te; p.slice0 = s0; p.nslice = n; p.T = T; p.classes = classes; p.logits = logits; p.head_bias = head_bias;
    p.in = X + (long long)s0 * hw * d.C0; p.in_stats = nullptr; p.wimg = w.w1; p.emap = E1t; p.out32 = c1a; p.out_stats = s1a;
    CKF((launch_band<D1S>(p, num_sms, st)));
    CKF(launch_gn_finalize(s1a, nb1, 64, (float)(4 * hw * 16), wx.gn1a_g, wx.gn1a_b, gss, n, st));
    p.in = c1a; p.in_stats = s1a; p.in_ss = gss; p.nb_in = nb1; p.in_count = (float)(4 * hw * 16); p.gamma = wx.gn1a_g; p.beta = wx.gn1a_b;
    p.wimg = w.w2; p.emap = nullptr; p.out32 = c1b; p.out_stats = s1b;
    if (narrow & 2) CKF((launch_band<D2T>(p, num_sms, st))); else CKF((launch_band<D2S>(p, num_sms, st)));
    CKF(launch_gn_finalize(s1b, nb2, 64, (float)(4 * hw * 16), wx.gn1b_g, wx.gn1b_b, gss, n, st));
    p.in = c1b; p.in_stats = s1b; p.nb_in = nb2; p.gamma = wx.gn1b_g; p.beta = wx.gn1b_b;
    p.wimg = w.w3; p.emap = E2t; p.out32 = c2a; p.out_stats = s2a;
    if (narrow & 4) CKF((launch_band<D3T>(p, num_sms, st))); else CKF((launch_band<D3S>(p, num_sms, st)));
    CKF(launch_gn_finalize(s2a, nb3, 32, (float)(16 * hw * 16), wx.gn2a_g, wx.gn2a_b, gss, n, st));
    p.in = c2a; p.in_stats = s2a; p.nb_in = nb3; p.in_count = (float)(16 * hw * 16); p.gamma = wx.gn2a_g; p.beta = wx.gn2a_b;
    p.wimg = w.w4; p.emap = nullptr; p.out32 = c2b; p.out_stats = s2b;
    if (narrow & 8) CKF((launch_band<D4T>(p, num_sms, st))); else CKF((launch_band<D4S>(p, num_sms, st)));
    CKF(launch_gn_finalize(s2b, nb4, 32, (float)(16 * hw * 16), wx.gn2b_g, wx.gn2b_b, gss, n, st));
    p.in = c2b; p.in_stats = s2b; p.nb_in = nb4; p.gamma = wx.gn2b_g; p.beta = wx.gn2b_b;
    p.wimg = w.w5; p.emap = nullptr; p.out32 = nullptr; p.out_stats = nullptr;
    static const bool head_tc = getenv("CATSEG_DEC_HEAD_TC") != nullptr;       // A/B: the tensor-core head stage
    if (lpeers != nullptr && (head_tc || !(d.D2 == HB_C && 4 * d.W == HB_W && 4 * d.H == HB_W))) return cudaErrorNotSupported;
    if (d.D2 == HB_C && 4 * d.W == HB_W && 4 * d.H == HB_W && !head_tc)
      CKF(launch_head_conv_f32(c2b, gss, wx.head_w, head_bias, classes, logits, s0, n, Te, T, num_sms, lpeers, nlp, st));
    else if (narrow & 16) CKF((launch_band<D5T>(p, num_sms, st)));
    else CKF((launch_band<D5S>(p, num_sms, st)));
  }
  if (launches) *launches += nl;
  return cudaSuccess;
}

}  // namespace catseg
