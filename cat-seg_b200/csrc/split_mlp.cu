// PRECISE token-wise MLP block on tcgen05 (fp16 hi + lo operand pairs, 3 MMAs per product, fp32 accumulate):
//     Xout[row] = Xin[row] (+ Xres[row]) + fc2(act(fc1(LayerNorm(Xin[row]))))
// Reference: the FFN half of SwinTransformerBlock.forward (model.py:223, Mlp 128 -> 512 GELU -> 128; Xin = Xout, no Xres)
// and the MLP of ClassTransformerLayer.forward (model.py:413 + the outer residual :423, ReLU; Xres = the layer input).
//
// One pass = 128 token rows (one UMMA M tile).  Shared memory holds LN(x) and one hidden chunk as hi/lo tile pairs
// (4 x 32.25 KiB) next to a 3-slot ring of 32 KiB weight images (16 images per pass: W1 hi/lo and W2 hi/lo of the four
// hidden chunks of 128, stored in consumption order).  TMEM: H0 H1 Y (3 x 128 columns): with two hidden accumulators the
// fc1 MMAs of chunk j+1 and the fc2 MMAs of chunk j-1 run under the activation epilogue of chunk j:
//     sync(j-1) -> issue MMA2(j-1), MMA1(j+1) | workers: epilogue(j) (TMEM H[j&1] -> act -> hi/lo tiles) -> sync(j) ...
// Roles as in fast_mlp.cu: warps 0-15 = LayerNorm prologue + epilogues, warp 16 = issuing warp (MMAs and the weight ring).
#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "split_common.cuh"

namespace catseg {

using namespace fast;

namespace {
constexpr int SP_EPI_WARPS = 16;
constexpr int SP_EPI_THREADS = SP_EPI_WARPS * 32;
constexpr int SP_THREADS = SP_EPI_THREADS + 32;
constexpr uint32_t SP_RING = 0;                                   // 3 x 32 KiB weight ring
constexpr uint32_t SP_XH = SP_RING + 3 * WIMG_BYTES;              // LN(x) hi
constexpr uint32_t SP_XL = SP_XH + TILE_BYTES_T;                  // LN(x) lo
constexpr uint32_t SP_HH = SP_XL + TILE_BYTES_T;                  // hidden chunk hi
constexpr uint32_t SP_HL = SP_HH + TILE_BYTES_T;                  // hidden chunk lo
constexpr uint32_t SP_PAR = SP_HL + TILE_BYTES_T;                 // b2[128] g[128] b[128] floats (b1 is read from global)
constexpr uint32_t SP_BAR = SP_PAR + 384 * 4;                     // full[3] empty[3] h[2] y[1] + tmem ptr
constexpr uint32_t SP_SMEM = SP_BAR + 10 * 8 + 16;
static_assert(SP_SMEM <= 232448, "shared memory budget");
static_assert(STG_BYTES <= 4 * TILE_BYTES_T, "the fp32 staging tile aliases the operand tiles");
constexpr uint32_t TM_H0 = 0, TM_Y = 256;
}  // namespace

// Q0FREE: the tcgen05.mma issue rate collapses when the issuing warp shares its scheduler with ALU-bound warps (probe:
// 16 MMAs take 11 K cycles to ISSUE next to three busy warps on the same scheduler, 1.2 K alone) -- and TMEM lane quarter q
// can only be read by warps with warp % 4 == q, i.e. by warps of scheduler q.  With Q0FREE a pass holds 96 tokens in tile
// rows 32..127: the warps of scheduler 0 have no rows and only keep the barriers company, the issuing warp (16 % 4 == 0) has
// that scheduler to itself.  A quarter of every MMA is dead rows; the tensor pipe has the head-room (it is ~30 % busy).
template <int ACT, bool Q0FREE, bool TIMING>   // ACT: 0 = GELU (Swin), 1 = ReLU (class layer); TIMING: phase-timing build
__global__ void __launch_bounds__(SP_THREADS, 1)
mlp_split_kernel(const float* __restrict__ Xin, const float* __restrict__ Xres, float* __restrict__ Xout, long long ntok,
                 MlpSplitW w, int dbg_nostream, long long* dbg) {
  static_assert(SP_EPI_WARPS == 16, "warp 16 (the issuing warp) must land on scheduler 0");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + SP_BAR);        // [3]
  uint64_t* bar_empty = bar_full + 3;                                      // [3]
  uint64_t* bar_h = bar_full + 6;                                          // [2] H0 / H1 ready
  uint64_t* bar_y = bar_full + 8;                                          // an fc2 chain has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 9);
  float* s_b2 = reinterpret_cast<float*>(smem + SP_PAR);
  float* s_g = s_b2 + 128;
  float* s_be = s_g + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == SP_EPI_WARPS;   // warp-uniform role

  constexpr int ROW0 = Q0FREE ? 32 : 0, TP = 128 - ROW0;      // first tile row in use, tokens per pass
  const long long npass = (ntok + TP - 1) / TP;
  long long my_pass = 0;
  for (long long p = blockIdx.x; p < npass; p += gridDim.x) ++my_pass;

  if (tid < 128) { s_b2[tid] = w.b2[tid]; s_g[tid] = w.ln_g[tid]; s_be[tid] = w.ln_b[tid]; }
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) umma::mbar_init(&bar_full[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t sb = umma::smem_u32(smem);
  const uint64_t d_xh = umma::make_smem_desc(sb + SP_XH, LBO_T, 128), d_xl = umma::make_smem_desc(sb + SP_XL, LBO_T, 128);
  const uint64_t d_hh = umma::make_smem_desc(sb + SP_HH, LBO_T, 128), d_hl = umma::make_smem_desc(sb + SP_HL, LBO_T, 128);
  const uint64_t d_w0 = umma::make_smem_desc(sb + SP_RING, LBO_W, 128);

  split::WeightRing<3, WIMG_BYTES> ring;
  ring.init(smem + SP_RING, bar_full, bar_empty, w.wimg, 16, my_pass * 16);
  ring.dbg_nostream = dbg_nostream;
  if (issuer) ring.prime();
  // one product chain A W^T with A = a_hi + a_lo and W = (next two ring images: hi, lo): 24 MMAs
  auto chain = [&](uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, bool acc) {
    uint32_t off = ring.acquire();
    uint64_t wd = d_w0 + (uint64_t)(off >> 4);
    if (umma::elect_one()) {
      issue_gemm_k128_desc(d_tmem, a_hi, LBO_T, wd, LBO_W, IDESC_128x128, acc);
      issue_gemm_k128_desc(d_tmem, a_lo, LBO_T, wd, LBO_W, IDESC_128x128, true);
    }
    __syncwarp();
    ring.release();
    off = ring.acquire();
    wd = d_w0 + (uint64_t)(off >> 4);
    if (umma::elect_one()) issue_gemm_k128_desc(d_tmem, a_hi, LBO_T, wd, LBO_W, IDESC_128x128, true);
    __syncwarp();
    ring.release();
  };

  const int q = warp & 3, cq = (warp >> 2) & 3;          // TMEM lane quarter, column quarter (32 columns)
  const int row = q * 32 + lane;
  const bool worker = !issuer && (!Q0FREE || q != 0);     // warps that own tile rows
  const int aw = Q0FREE ? (q - 1) + 3 * cq : warp;        // index among the NAW working warps (row-wise phases: 8 rows each)
  constexpr int NAW = Q0FREE ? 12 : 16;
  float* stage = reinterpret_cast<float*>(smem + SP_XH);
  const uint32_t lane_addr = tm + ((uint32_t)(q * 32) << 16);
  uint32_t ph_h[2] = {0u, 0u}, ph_y = 0u;

  // CATSEG_PHASE_TIMING=1: cycles per phase, accumulated by thread 32 (a working warp) and by the issuing warp of CTA 0
  long long pt[TIMING ? 10 : 1] = {}, t_last = TIMING ? clock64() : 0, npass_dbg = 0;
  const bool timer = TIMING && dbg != nullptr && blockIdx.x == 0 && (tid == 32 || tid == SP_EPI_THREADS);
#define MPH(i) do { if constexpr (TIMING) { if (timer) { const long long _t = clock64(); pt[i] += _t - t_last; t_last = _t; } } } while (0)
  for (long long p = blockIdx.x; p < npass; p += gridDim.x) {
    const long long row0 = p * TP;
    if constexpr (TIMING) ++npass_dbg;
    if (!issuer) {
      const long long nrow0 = (p + gridDim.x) * TP;       // next pass: 4 lines of 128 bytes per token row into L2
      const long long r = nrow0 + (tid >> 2);
      if ((tid >> 2) < TP && r < ntok) {
        umma::prefetch_l2(Xin + r * 128 + (tid & 3) * 32);
        if (Xres != nullptr) umma::prefetch_l2(Xres + r * 128 + (tid & 3) * 32);   // read once, in the Y epilogue
      }
    }
    if (worker) {
      const long long nv = ntok - row0;
      split::ln_rows_to_tile_split(Xin + row0 * 128, 128, nv >= TP ? TP : (int)nv, smem + SP_XH + ROW0 * 16, smem + SP_XL + ROW0 * 16,
                                   s_g, s_be, aw, NAW, lane, TP);
      umma::fence_proxy_async();
    }
    MPH(0);                                                // LayerNorm prologue
    umma::fence_before_sync();
    __syncthreads();
    MPH(1);                                                // barrier after the prologue
    if (issuer) {
      umma::fence_after_sync();
      chain(tm + TM_H0, d_xh, d_xl, false);
      if (umma::elect_one()) umma::mma_commit(&bar_h[0]);
      __syncwarp();
      chain(tm + TM_H0 + 128, d_xh, d_xl, false);
      if (umma::elect_one()) umma::mma_commit(&bar_h[1]);
      __syncwarp();
    }
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int hb = j & 1;
      if (worker) {
        umma::mbar_wait(&bar_h[hb], ph_h[hb]); ph_h[hb] ^= 1u;
        umma::fence_after_sync();
        MPH(2);                                            // wait for H[j]
        uint4 phi[4], plo[4];
        {
          float v[32];
          umma::tmem_ld32(lane_addr + TM_H0 + hb * 128 + cq * 32, v);
          const float4* bb = reinterpret_cast<const float4*>(w.b1 + j * 128 + cq * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(bb + i);
            if constexpr (ACT == 0) {                         // GELU on packed pairs (bias added inside)
              split::gelu_precise_pair(v[4 * i], v[4 * i + 1], b4.x, b4.y);
              split::gelu_precise_pair(v[4 * i + 2], v[4 * i + 3], b4.z, b4.w);
            } else {
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
          }
          if constexpr (ACT != 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            umma::split_h2(v[c * 8 + 0], v[c * 8 + 1], phi[c].x, plo[c].x);
            umma::split_h2(v[c * 8 + 2], v[c * 8 + 3], phi[c].y, plo[c].y);
            umma::split_h2(v[c * 8 + 4], v[c * 8 + 5], phi[c].z, plo[c].z);
            umma::split_h2(v[c * 8 + 6], v[c * 8 + 7], phi[c].w, plo[c].w);
          }
        }
        MPH(3);                                            // activation + split
        if (j > 0) {                                       // the hidden tiles are free once fc2 of chunk j-1 has completed
          umma::mbar_wait(bar_y, ph_y); ph_y ^= 1u;
          umma::fence_after_sync();
        }
        MPH(4);                                            // wait for fc2(j-1)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          *reinterpret_cast<uint4*>(smem + SP_HH + (cq * 4 + c) * LBO_T + row * 16) = phi[c];
          *reinterpret_cast<uint4*>(smem + SP_HL + (cq * 4 + c) * LBO_T + row * 16) = plo[c];
        }
        umma::fence_proxy_async();
        MPH(5);                                            // hidden tile stores
      }
      umma::fence_before_sync();
      __syncthreads();                                     // hidden tiles written; every warp has read H[hb]
      MPH(6);                                              // chunk barrier (issuing warp: waiting for the workers)
      if (issuer) {
        umma::fence_after_sync();
        chain(tm + TM_Y, d_hh, d_hl, j > 0);
        if (umma::elect_one()) umma::mma_commit(bar_y);
        __syncwarp();
        if (j + 2 < 4) {                                   // H[hb] is free: fc1 of chunk j+2
          chain(tm + TM_H0 + hb * 128, d_xh, d_xl, false);
          if (umma::elect_one()) umma::mma_commit(&bar_h[hb]);
          __syncwarp();
        }
        MPH(7);                                            // issuing warp: issuing the chains of this step
      }
    }
    // ---- Y epilogue: residual rows (coalesced, warp per row) are fetched before waiting for the last fc2 chain
    float4 xres[8];
    if (worker) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long r = row0 + aw * 8 + i;
        // (the second residual, Xres, is fetched in the store loop below: added here, every row's pair of loads had to land
        //  before the next pair was issued -- 8 serialised L2 round trips, 6 K cycles per pass in the class MLP)
        xres[i] = r < ntok ? ld4(Xin + r * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      umma::mbar_wait(bar_y, ph_y); ph_y ^= 1u;
      umma::fence_after_sync();
      MPH(8);                                              // residual loads + wait for the last fc2 chain
      float v[32];
      umma::tmem_ld32(lane_addr + TM_Y + cq * 32, v);
      const float* bb = s_b2 + cq * 32;
      float* sp = stage + row * STG_LD + cq * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) st4(sp + i, make_float4(v[i] + bb[i], v[i + 1] + bb[i + 1], v[i + 2] + bb[i + 2], v[i + 3] + bb[i + 3]));
    }
    umma::fence_before_sync();
    __syncthreads();
    if (worker) {
      if (Xres != nullptr) {
        float4 x2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long r = row0 + aw * 8 + i;
          x2[i] = r < ntok ? ld4(Xres + r * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) xres[i] = f4add(xres[i], x2[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long r = row0 + aw * 8 + i;
        if (r < ntok) st4(Xout + r * 128 + lane * 4, f4add(xres[i], ld4(stage + (ROW0 + aw * 8 + i) * STG_LD + lane * 4)));
      }
    }
    __syncthreads();
    umma::fence_after_sync();          // TMEM and the operand tiles may be overwritten by the next pass
    MPH(9);                                                // Y epilogue: staging, barrier, coalesced stores, barrier
  }
#undef MPH
  if constexpr (TIMING) {
    if (timer) {
      long long* o = dbg + (tid == 32 ? 0 : 16);
      for (int i = 0; i < 10; ++i) o[i] = pt[i];
      o[10] = npass_dbg;
    }
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tm);
}

cudaError_t launch_mlp_split(const float* Xin, const float* Xres, float* Xout, long long ntok, const MlpSplitW& w, int act,
                             int num_sms, cudaStream_t st) {
  static long long* tdbg = nullptr;
  static int timing = -1;
  if (timing < 0) {
    const char* e4 = getenv("CATSEG_PHASE_TIMING");
    timing = (e4 && e4[0] == '1') ? 1 : 0;
    if (timing) { cudaMalloc(&tdbg, 32 * sizeof(long long)); cudaMemset(tdbg, 0, 32 * sizeof(long long)); }
  }
  static int q0free = -1;
  if (q0free < 0) { const char* e3 = getenv("CATSEG_MLP_Q0FREE"); q0free = e3 ? atoi(e3) : 0; }
  cudaError_t e = cudaSuccess;
  for (auto* fn : {(const void*)mlp_split_kernel<0, false, false>, (const void*)mlp_split_kernel<1, false, false>,
                   (const void*)mlp_split_kernel<0, true, false>, (const void*)mlp_split_kernel<1, true, false>,
                   (const void*)mlp_split_kernel<0, false, true>, (const void*)mlp_split_kernel<1, false, true>})
    if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_SMEM)) != cudaSuccess) return e;
  const int tp = (q0free && !timing) ? 96 : 128;
  const long long npass = (ntok + tp - 1) / tp;
  const int grid = (int)(npass < num_sms ? npass : num_sms);
  if (grid <= 0) return cudaSuccess;
  static int dbg = -1;
  if (dbg < 0) { const char* e2 = getenv("CATSEG_DBG_NOSTREAM"); dbg = e2 ? atoi(e2) : 0; }
  if (timing) {
    if (act == 0) mlp_split_kernel<0, false, true><<<grid, SP_THREADS, SP_SMEM, st>>>(Xin, Xres, Xout, ntok, w, dbg, tdbg);
    else mlp_split_kernel<1, false, true><<<grid, SP_THREADS, SP_SMEM, st>>>(Xin, Xres, Xout, ntok, w, dbg, tdbg);
  } else if (q0free) {
    if (act == 0) mlp_split_kernel<0, true, false><<<grid, SP_THREADS, SP_SMEM, st>>>(Xin, Xres, Xout, ntok, w, dbg, tdbg);
    else mlp_split_kernel<1, true, false><<<grid, SP_THREADS, SP_SMEM, st>>>(Xin, Xres, Xout, ntok, w, dbg, tdbg);
  } else {
    if (act == 0) mlp_split_kernel<0, false, false><<<grid, SP_THREADS, SP_SMEM, st>>>(Xin, Xres, Xout, ntok, w, dbg, tdbg);
    else mlp_split_kernel<1, false, false><<<grid, SP_THREADS, SP_SMEM, st>>>(Xin, Xres, Xout, ntok, w, dbg, tdbg);
  }
  if (timing) {
    long long hb[32];
    cudaStreamSynchronize(st);
    cudaMemcpy(hb, tdbg, sizeof(hb), cudaMemcpyDeviceToHost);
    const double n = hb[10] > 0 ? (double)hb[10] : 1.0;
    fprintf(stderr, "[mlp_split act=%d q0free=%d, cycles per pass over %lld passes] worker: LN %.0f | bar %.0f | wait-H %.0f | act+split %.0f | "
            "wait-fc2 %.0f | H stores %.0f | chunk bar %.0f | resid+wait-Y %.0f | Y epilogue %.0f  || issuer: idle-before-pass %.0f | bar %.0f | "
            "chunk-bar wait %.0f | issue %.0f | tail %.0f\n", act, q0free, hb[10], hb[0] / n, hb[1] / n, hb[2] / n, hb[3] / n, hb[4] / n,
            hb[5] / n, hb[6] / n, hb[8] / n, hb[9] / n, hb[16] / n, hb[17] / n, hb[22] / n, hb[23] / n, hb[25] / n);
  }
  return cudaGetLastError();
}

// ---- weight image packing: hi / lo fp16 images (canonical dense 128x128) <- W[r0 + r][c0 + k], ld = row stride
__global__ void pack_wimg_split_kernel(__half* __restrict__ dhi, __half* __restrict__ dlo, const float* __restrict__ W, int ld,
                                       int r0, int c0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 128) return;
  const int r = i >> 7, k = i & 127;
  const float v = W[(long long)(r0 + r) * ld + c0 + k];
  const __half h = __float2half_rn(v);
  const int o = (k >> 3) * (128 * 8) + r * 8 + (k & 7);
  dhi[o] = h;
  if (dlo != nullptr) dlo[o] = __float2half_rn(v - __half2float(h));
}
cudaError_t launch_pack_wimg_split(__half* dhi, __half* dlo, const float* W, int ld, int r0, int c0, cudaStream_t st) {
  pack_wimg_split_kernel<<<64, 256, 0, st>>>(dhi, dlo, W, ld, r0, c0);
  return cudaGetLastError();
}

cudaError_t pack_mlp_split(__half* dst, const float* W1, const float* W2, cudaStream_t st) {
  // consumption order of mlp_split_kernel: (fc1, chunk) and (fc2, chunk) pairs, each as a hi and a lo image
  static const int kind[8] = {1, 1, 2, 1, 2, 1, 2, 2}, chunk[8] = {0, 1, 0, 2, 1, 3, 2, 3};
  for (int i = 0; i < 8; ++i) {
    __half* hi = dst + (size_t)(2 * i) * 128 * 128;
    __half* lo = hi + 128 * 128;
    cudaError_t e = kind[i] == 1 ? launch_pack_wimg_split(hi, lo, W1, 128, chunk[i] * 128, 0, st)
                                 : launch_pack_wimg_split(hi, lo, W2, 512, 0, chunk[i] * 128, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace catseg
