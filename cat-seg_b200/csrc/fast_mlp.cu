// FAST token-wise MLP block on tcgen05:  X[row] += fc2(act(fc1(LayerNorm(X[row]))))   (in place)
//
// Reference: the FFN half of SwinTransformerBlock.forward (model.py:223, Mlp 128->512 GELU ->128).
// Persistent CTAs; each pass handles 256 token rows (two 128-row UMMA tiles) so that the 256 KiB of
// fp16 weights, streamed from L2 through a 3-slot ring of 32 KiB bulk (TMA) copies, are fetched
// once per 256 tokens.  Per hidden chunk j (128 of the 512 hidden units) and tile t:
//   MMA1: H_t = LN(x_t) W1_j^T   -> TMEM          (8 x tcgen05.mma M128 N128 K16)
//   epilogue: TMEM -> +b1 -> GELU -> fp16 -> shared memory (canonical K-major A operand)
//   MMA2: Y_t += h W2_j^T        -> TMEM
// and finally Y_t + b2 + x -> global through an fp32 staging tile (coalesced rows).
// TMEM: H_a H_b Y_a Y_b = 4 x 128 columns.
//
// Roles.  Warps 0-15 do the LayerNorm prologue and the epilogues (4 TMEM lane quarters x 4 column quarters).
// Warp 16 only issues: tcgen05.mma and the weight ring.  Measured (profiles/README.md): the MMA
// queue is shallow, so the issuing thread is blocked for the whole duration of the MMAs it issues, and a warp that
// joins an ALU-bound phase late is starved by the three warps sharing its scheduler until they finish; an epilogue
// warp that also issues therefore delays every __syncthreads of the pass.  MMA1 of chunk j+1 is queued right behind
// MMA2 of chunk j (as soon as H_t has been read), so the tensor pipe runs under the other tile's epilogue.
#include <cstdio>
#include <cstdlib>

#include "fast_common.cuh"
#include "internal.h"

namespace catseg {

using namespace fast;

namespace {
constexpr int MLP_EPI_WARPS = 16;
constexpr int MLP_EPI_THREADS = MLP_EPI_WARPS * 32;
constexpr int MLP_THREADS = MLP_EPI_THREADS + 32;                // + the issuing warp
constexpr uint32_t SM_RING = 0;                                   // 3 x 32 KiB weight ring
constexpr uint32_t SM_XN = SM_RING + 3 * WIMG_BYTES;              // 2 tiles LN(x) fp16
constexpr uint32_t SM_H = SM_XN + 2 * TILE_BYTES_T;               // hidden chunk fp16
constexpr uint32_t SM_PAR = SM_H + TILE_BYTES_T;                  // b1[512] b2[128] g[128] b[128] floats
constexpr uint32_t SM_BAR = SM_PAR + (512 + 128 + 128 + 128) * 4; // 7 mbarriers + tmem ptr
constexpr uint32_t MLP_SMEM = SM_BAR + 8 * 8 + 16;
static_assert(STG_BYTES <= 3 * TILE_BYTES_T, "the fp32 staging tile aliases the LN tiles + hidden chunk");
}  // namespace

#define PH(i) do { if (dbg != nullptr && blockIdx.x == 0 && tid == 0) { long long _t = clock64(); pacc##i += _t - t_last; t_last = _t; } } while (0)

template <int ACT>   // 0 = GELU (Swin), 1 = ReLU (class layer)
__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_fast_kernel(float* __restrict__ X, long long ntok, MlpFastW w, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + SM_BAR);       // [3]
  uint64_t* bar_mma = bar_full + 3;                                       // [4]: H_a, H_b ready; MMA2_a, MMA2_b done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 4);
  float* s_b1 = reinterpret_cast<float*>(smem + SM_PAR);
  float* s_b2 = s_b1 + 512;
  float* s_g = s_b2 + 128;
  float* s_be = s_g + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = __shfl_sync(0xffffffffu, warp, 0) == MLP_EPI_WARPS;   // warp-uniform role

  const long long npass = (ntok + 255) / 256;
  long long my_pass = 0;
  for (long long p = blockIdx.x; p < npass; p += gridDim.x) ++my_pass;
  const long long total_loads = my_pass * 8;

  for (int i = tid; i < 512; i += MLP_THREADS) s_b1[i] = w.b1[i];
  if (tid < 128) { s_b2[tid] = w.b2[tid]; s_g[tid] = w.ln_g[tid]; s_be[tid] = w.ln_b[tid]; }
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) umma::mbar_init(&bar_full[i], 1);
    for (int i = 0; i < 4; ++i) umma::mbar_init(&bar_mma[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t smem_base = umma::smem_u32(smem);
  // base descriptors, built once
  const uint64_t d_xn0 = umma::make_smem_desc(smem_base + SM_XN, LBO_T, 128);
  const uint64_t d_xn1 = umma::make_smem_desc(smem_base + SM_XN + TILE_BYTES_T, LBO_T, 128);
  const uint64_t d_h = umma::make_smem_desc(smem_base + SM_H, LBO_T, 128);
  const uint64_t d_w0 = umma::make_smem_desc(smem_base + SM_RING, LBO_W, 128);

  // weight image n (ring order W1_0, W2_0, W1_1, ...) lives in slot n % 3; one elected lane of the issuing warp
  auto issue_load = [&](long long n) {
    if (n < total_loads) {
      int s = (int)(n % 3);
      umma::mbar_expect_tx(&bar_full[s], WIMG_BYTES);
      umma::bulk_g2s(smem + SM_RING + s * WIMG_BYTES, reinterpret_cast<const uint8_t*>(w.wimg) + (n % 8) * WIMG_BYTES,
                     WIMG_BYTES, &bar_full[s]);
    }
  };
  auto wait_image = [&](long long n) { umma::mbar_wait(&bar_full[n % 3], (uint32_t)((n / 3) & 1)); };
  auto wdesc = [&](long long n) { return d_w0 + (uint64_t)((uint32_t)(n % 3) * (WIMG_BYTES >> 4)); };
  if (issuer) {
    if (umma::elect_one()) { issue_load(0); issue_load(1); issue_load(2); }
    __syncwarp();
  }

  const int q = warp & 3, cq = (warp >> 2) & 3;          // TMEM lane quarter, column quarter (32 columns)
  const int row = q * 32 + lane;                         // row of the tile owned by this thread
  float* stage = reinterpret_cast<float*>(smem + SM_XN); // fp32 staging tile (LN tiles + h are dead by then)
  const uint32_t lane_addr = tm + ((uint32_t)(q * 32) << 16);
  long long g = 0;                                       // running hidden-chunk counter of this CTA
  long long t_last = clock64();
  long long pacc0 = 0, pacc1 = 0, pacc2 = 0, pacc3 = 0, pacc4 = 0, pacc5 = 0, npass_dbg = 0;

  for (long long p = blockIdx.x; p < npass; p += gridDim.x) {
    const long long row0 = p * 256;
    if (!issuer) {
      {   // the NEXT pass's 256 rows (128 KiB) are prefetched into L2 while this pass computes
        const long long nrow0 = (p + gridDim.x) * 256;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          long long line = (long long)tid + i * MLP_EPI_THREADS;        // 1024 lines of 128 bytes
          long long r = nrow0 + (line >> 2);
          if (r < ntok) umma::prefetch_l2(X + r * 128 + (line & 3) * 32);
        }
      }
      // ---- LN prologue: two tiles, loads of the next 4 rows in flight under the arithmetic of the current 4
      {
        const long long nv = ntok - row0;
        ln_rows_to_tiles_pipelined<2>(X + row0 * 128, 128, nv >= 256 ? 256 : (int)nv, smem + SM_XN, s_g, s_be, warp, lane);
      }
      umma::fence_proxy_async();
    }
    umma::fence_before_sync();
    __syncthreads();
    PH(0);
    if (issuer) {                                          // MMA1 of the first chunk, both tiles
      umma::fence_after_sync();
      wait_image(2 * g);
      if (umma::elect_one()) {
        issue_gemm_k128_desc(tm + 0, d_xn0, LBO_T, wdesc(2 * g), LBO_W, IDESC_128x128, false);
        umma::mma_commit(&bar_mma[0]);
        issue_gemm_k128_desc(tm + 128, d_xn1, LBO_T, wdesc(2 * g), LBO_W, IDESC_128x128, false);
        umma::mma_commit(&bar_mma[1]);
      }
      __syncwarp();
      umma::mbar_wait(&bar_mma[1], (uint32_t)(g & 1));     // W1_g consumed -> its slot takes image 2g+3
      if (umma::elect_one()) issue_load(2 * g + 3);
      __syncwarp();
    }

    for (int j = 0; j < 4; ++j, ++g) {
      const uint32_t par = (uint32_t)(g & 1);
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        if (!issuer) {
          umma::mbar_wait(&bar_mma[t], par);                 // H_t ready
          umma::fence_after_sync();
          PH(1);
          // ---- H_t -> bias -> act -> fp16 (32 columns per thread)
          uint4 packed[4];
          {
            float v[32];
            umma::tmem_ld32(lane_addr + t * 128 + cq * 32, v);
            const float* bb = s_b1 + j * 128 + cq * 32;
            if (ACT == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) gelu_fast_pair(v[i], v[i + 1], bb[i], bb[i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bb[i], 0.0f);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
              packed[c] = make_uint4(umma::pack_h2(v[c * 8 + 0], v[c * 8 + 1]), umma::pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                                     umma::pack_h2(v[c * 8 + 4], v[c * 8 + 5]), umma::pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
          }
          PH(2);
          // ---- the h buffer is free once the previous MMA2 has completed
          if (t == 0) {
            if (j > 0) umma::mbar_wait(&bar_mma[3], (uint32_t)((g - 1) & 1));
          } else {
            umma::mbar_wait(&bar_mma[2], par);
          }
          umma::fence_after_sync();
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(smem + SM_H + (cq * 4 + c) * LBO_T + row * 16) = packed[c];
          umma::fence_proxy_async();
        }
        umma::fence_before_sync();
        __syncthreads();                                     // h written; every warp has read H_t
        PH(3);
        if (issuer) {
          umma::fence_after_sync();
          if (t == 0) {
            wait_image(2 * g + 1);
            if (j < 3) wait_image(2 * g + 2);
          }
          if (umma::elect_one()) {
            issue_gemm_k128_desc(tm + 256 + t * 128, d_h, LBO_T, wdesc(2 * g + 1), LBO_W, IDESC_128x128, j > 0);
            umma::mma_commit(&bar_mma[2 + t]);
            if (j < 3) {                                     // H_t is free: MMA1 of the next chunk runs under the other tile's epilogue
              issue_gemm_k128_desc(tm + t * 128, t == 0 ? d_xn0 : d_xn1, LBO_T, wdesc(2 * g + 2), LBO_W, IDESC_128x128, false);
              umma::mma_commit(&bar_mma[t]);
            }
          }
          __syncwarp();
          if (t == 1) {                                      // ring bookkeeping: images 2g+1 (W2_g) and 2g+2 (W1_g+1) are consumed
            umma::mbar_wait(&bar_mma[3], par);
            if (umma::elect_one()) issue_load(2 * g + 4);
            __syncwarp();
            if (j < 3) {
              umma::mbar_wait(&bar_mma[1], (uint32_t)((g + 1) & 1));
              if (umma::elect_one()) issue_load(2 * g + 5);
              __syncwarp();
            }
          }
        }
      }
    }
    // ---- Y epilogue.  The residual rows of tile a are fetched (coalesced, warp per row, 8 loads in flight) BEFORE
    //      waiting for the last MMAs, those of tile b while tile a is being staged.  Y_t + b2 goes through an fp32
    //      staging tile (thread = row) so that X is read and written with full 512-byte rows.
    //      (A bulk fp32 reduce-add of the staged rows into X was measured 2.4x slower: 8 K cycles per 64 KiB tile.)
    float4 xres[8];
    if (!issuer) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long r = row0 + warp * 8 + i;
        xres[i] = r < ntok ? ld4(X + r * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    PH(4);
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
      if (!issuer) {
        umma::mbar_wait(&bar_mma[2 + t], (uint32_t)((g - 1) & 1));   // Y_t complete (tile b's last MMA2 runs under tile a's epilogue)
        umma::fence_after_sync();
        float v[32];
        umma::tmem_ld32(lane_addr + 256 + t * 128 + cq * 32, v);
        const float* bb = s_b2 + cq * 32;
        float* sp = stage + row * STG_LD + cq * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) st4(sp + i, make_float4(v[i] + bb[i], v[i + 1] + bb[i + 1], v[i + 2] + bb[i + 2], v[i + 3] + bb[i + 3]));
      }
      umma::fence_before_sync();
      __syncthreads();
      if (!issuer) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long r = row0 + t * 128 + warp * 8 + i;
          if (r < ntok) st4(X + r * 128 + lane * 4, f4add(xres[i], ld4(stage + (warp * 8 + i) * STG_LD + lane * 4)));
        }
        if (t == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const long long r = row0 + 128 + warp * 8 + i;
            xres[i] = r < ntok ? ld4(X + r * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      __syncthreads();
    }
    umma::fence_after_sync();          // TMEM and the LN tiles may be overwritten by the next pass
    PH(5);
    ++npass_dbg;
  }
  if (dbg != nullptr && blockIdx.x == 0 && tid == 0) {
    dbg[0] = pacc0; dbg[1] = pacc1; dbg[2] = pacc2; dbg[3] = pacc3; dbg[4] = pacc4; dbg[5] = pacc5; dbg[15] = npass_dbg;
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tm);
}

cudaError_t launch_mlp_fast(float* X, long long ntok, const MlpFastW& w, int act, int num_sms, cudaStream_t st) {
  {   // per-device function attribute: set on every launch (cheap), a process-wide flag would miss other devices
    cudaError_t e = cudaFuncSetAttribute(mlp_fast_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mlp_fast_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM);
    if (e != cudaSuccess) return e;
  }
  long long npass = (ntok + 255) / 256;
  int grid = (int)(npass < num_sms ? npass : num_sms);
  if (grid <= 0) return cudaSuccess;
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("CATSEG_PHASE_TIMING");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on) { cudaMalloc(&dbg, 16 * sizeof(long long)); cudaMemset(dbg, 0, 16 * sizeof(long long)); }
  }
  if (act == 0) mlp_fast_kernel<0><<<grid, MLP_THREADS, MLP_SMEM, st>>>(X, ntok, w, dbg_on ? dbg : nullptr);
  else mlp_fast_kernel<1><<<grid, MLP_THREADS, MLP_SMEM, st>>>(X, ntok, w, dbg_on ? dbg : nullptr);
  if (dbg_on) {
    long long hb[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    cudaMemset(dbg, 0, 16 * sizeof(long long));
    double n = hb[15] > 0 ? (double)hb[15] : 1.0;
    fprintf(stderr, "[mlp phases, cycles/pass(256 tok) over %lld passes, epilogue thread 0] LN %.0f | per (chunk,tile) x8: wait-H %.0f act %.0f "
            "h-free+STS+sync %.0f | wait-last %.0f Y-epi %.0f\n", hb[15], hb[0] / n, hb[1] / n / 8, hb[2] / n / 8, hb[3] / n / 8, hb[4] / n, hb[5] / n);
  }
  return cudaGetLastError();
}

// ---- weight image packing: dst image (fp16, canonical dense 128x128) <- W[r0 + r][c0 + k], ld = row stride
__global__ void pack_wimg_kernel(__half* __restrict__ dst, const float* __restrict__ W, int ld, int r0, int c0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 128) return;
  int r = i >> 7, k = i & 127;
  dst[(k >> 3) * (128 * 8) + r * 8 + (k & 7)] = __float2half_rn(W[(long long)(r0 + r) * ld + c0 + k]);
}
cudaError_t launch_pack_wimg(__half* dst, const float* W, int ld, int r0, int c0, cudaStream_t st) {
  pack_wimg_kernel<<<64, 256, 0, st>>>(dst, W, ld, r0, c0);
  return cudaGetLastError();
}

}  // namespace catseg
