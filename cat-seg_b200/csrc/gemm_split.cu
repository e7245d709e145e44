// General fp32-accurate GEMM on tcgen05:  C[b][m][n] = epilogue( sum_k A[b][m][k] * B[b][n][k] )
//
// fp32 operands with arbitrary element strides are staged by the threads as hi + lo fp16 canonical tiles (split_common.cuh)
// and contracted with three MMAs per product (hi*hi, lo*hi, hi*lo; fp32 accumulation in TMEM): relative error ~2^-21,
// i.e. the results are interchangeable with an fp32 GEMM.  It serves the contractions of the front end that are too small
// or too irregular for a dedicated kernel:
//   * the cost volume  einsum('bchw,btpc->bpthw')  with both L2 normalisations applied as row / column scales in the
//     epilogue and the per-class maximum over the pixels reduced there as well (model.py:648-652, 695);
//   * the token-wise linears of the CLIP dense last block (model_vpt.py:219-240, 306-312) with bias / QuickGELU /
//     residual epilogues.
// One CTA = one 128 x 128 output tile; K is consumed in chunks of 32 through two shared-memory buffers: the threads
// stage chunk c+1 while the tensor pipe works on chunk c; 64 KiB per CTA, so three CTAs share an SM and one CTA's staging
// overlaps the others' MMAs and epilogues.
#include "internal.h"
#include "split_common.cuh"

namespace catseg {

using namespace fast;

namespace {
constexpr int GS_THREADS = 256;
constexpr int GS_KC = 32;                                      // K per chunk: 4 canonical 8-element chunks (64 KiB per CTA: 3 CTAs per SM)
constexpr uint32_t GS_TILE = (GS_KC / 8) * LBO_W;              // one 128 x 32 fp16 tile = 8 KiB
constexpr uint32_t GS_BUF = 4 * GS_TILE;                       // A hi | A lo | B hi | B lo
constexpr uint32_t GS_BAR = 2 * GS_BUF;
constexpr uint32_t GS_SMEM = GS_BAR + 4 * 8 + 16;
}  // namespace

__global__ void __launch_bounds__(GS_THREADS, 3) gemm_split_kernel(GemmSplitParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_free = reinterpret_cast<uint64_t*>(smem + GS_BAR);     // [2] the MMAs that read buffer i have completed
  uint64_t* bar_done = bar_free + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_free + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * 128, m0 = blockIdx.y * 128, b = blockIdx.z;
  const float* A = p.A + (long long)b * p.a_batch;
  const float* B = p.B + (long long)b * p.b_batch;
  const int32_t* aidx = p.a_index != nullptr ? p.a_index + (long long)b * p.ai_batch : nullptr;
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) umma::mbar_init(&bar_free[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<128>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot, sb = umma::smem_u32(smem);
  const int nchunk = (p.K + GS_KC - 1) / GS_KC;

  // thread -> (row, 8-element K piece): a warp covers 32 consecutive rows of one piece, so the loads coalesce when the rows
  // are contiguous in memory (stride_row == 1) and use whole 32-byte sectors when K is (stride_k == 1)
  auto stage = [&](const float* src, long long s_row, long long s_k, int row0, int nrows, int k0, uint8_t* hi, uint8_t* lo,
                   const int32_t* index) {
#pragma unroll
    for (int i = 0; i < GS_KC / 16; ++i) {
      const int item = tid + i * GS_THREADS, r = item & 127, c = item >> 7;
      const int k = k0 + c * 8;
      float v[8];
      const bool rv = row0 + r < nrows;
      const long long srow = (rv && index != nullptr) ? (long long)__ldg(index + row0 + r) : (long long)(row0 + r);
      const float* q = src + srow * s_row + (long long)k * s_k;
      if (rv && s_k == 1 && k + 8 <= p.K && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
        const float4 a = ld4(q), c4 = ld4(q + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c4.x; v[5] = c4.y; v[6] = c4.z; v[7] = c4.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (rv && k + j < p.K) ? __ldg(q + (long long)j * s_k) : 0.0f;
      }
      uint4 h, l;
      umma::split_h2(v[0], v[1], h.x, l.x);
      umma::split_h2(v[2], v[3], h.y, l.y);
      umma::split_h2(v[4], v[5], h.z, l.z);
      umma::split_h2(v[6], v[7], h.w, l.w);
      *reinterpret_cast<uint4*>(hi + c * LBO_W + r * 16) = h;
      *reinterpret_cast<uint4*>(lo + c * LBO_W + r * 16) = l;
    }
  };

  uint32_t ph[2] = {0u, 0u};
  for (int c = 0; c < nchunk; ++c) {
    const int buf = c & 1;
    uint8_t* base = smem + buf * GS_BUF;
    if (c >= 2) { umma::mbar_wait(&bar_free[buf], ph[buf]); ph[buf] ^= 1u; }
    stage(A, p.a_row, p.a_k, m0, p.M, c * GS_KC, base, base + GS_TILE, aidx);
    stage(B, p.b_row, p.b_k, n0, p.N, c * GS_KC, base + 2 * GS_TILE, base + 3 * GS_TILE, nullptr);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t a = sb + buf * GS_BUF;
        const uint64_t d_ah = umma::make_smem_desc(a, LBO_W, 128), d_al = umma::make_smem_desc(a + GS_TILE, LBO_W, 128);
        const uint64_t d_bh = umma::make_smem_desc(a + 2 * GS_TILE, LBO_W, 128), d_bl = umma::make_smem_desc(a + 3 * GS_TILE, LBO_W, 128);
#pragma unroll
        for (int k = 0; k < GS_KC / 16; ++k) {
          const uint64_t o = (uint64_t)(k * 2 * (LBO_W >> 4));
          umma::mma_f16_ss(tm, d_ah + o, d_bh + o, IDESC_128x128, (c > 0 || k > 0) ? 1u : 0u);
          umma::mma_f16_ss(tm, d_al + o, d_bh + o, IDESC_128x128, 1u);
          umma::mma_f16_ss(tm, d_ah + o, d_bl + o, IDESC_128x128, 1u);
        }
        umma::mma_commit(c + 1 < nchunk ? &bar_free[buf] : bar_done);
      }
      __syncwarp();
    }
  }
  umma::mbar_wait(bar_done, 0);
  umma::fence_after_sync();
  // ---- epilogue: thread = row (TMEM lane quarter = warp & 3), column half = warp >> 2
  {
    const int q4 = warp & 3, half = warp >> 2;
    const int m = m0 + q4 * 32 + lane;
    const float rs = (p.row_scale != nullptr && m < p.M) ? p.row_scale[(long long)b * p.rs_batch + (aidx != nullptr ? aidx[m] : m)] : 1.0f;
    float rmax = -INFINITY;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int nb = n0 + half * 64 + cc * 32;
      float v[32];
      umma::tmem_ld32(tm + ((uint32_t)(q4 * 32) << 16) + half * 64 + cc * 32, v);
      if (m < p.M) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int n = nb + i;
          if (n < p.N) {
            float x = v[i] * rs;
            if (p.col_scale != nullptr) x *= __ldg(p.col_scale + (long long)b * p.cs_batch + n);
            if (p.bias != nullptr) x += __ldg(p.bias + n);
            if (p.act == 1) x = fmaxf(x, 0.0f);
            else if (p.act == 2) x = x / (1.0f + expf(-1.702f * x));           // QuickGELU: x * sigmoid(1.702 x)
            if (p.residual != nullptr) x += __ldg(p.residual + (long long)b * p.r_batch + (long long)(p.r_mod > 0 ? m % p.r_mod : m) * p.r_row + n);
            v[i] = x;
            rmax = fmaxf(rmax, x);
          }
        }
        if (p.C != nullptr) {
          float* o = p.C + (long long)b * p.c_batch + (long long)m * p.c_row + nb;
          if (nb + 32 <= p.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) st4(o + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) o[i] = v[i];
          }
        }
      }
    }
    // per-row maximum over this CTA's 64-column half -> row_max[b][m][2 * tile_n + half] (reduced by the consumer)
    if (p.row_max != nullptr && m < p.M) p.row_max[((long long)b * p.M + m) * (2 * gridDim.x) + 2 * blockIdx.x + half] = rmax;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<128>(tm);
}

cudaError_t launch_gemm_split(const GemmSplitParams& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.batch <= 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(gemm_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GS_SMEM);
  if (e != cudaSuccess) return e;
  dim3 grid((p.N + 127) / 128, (p.M + 127) / 128, p.batch);
  if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidValue;
  gemm_split_kernel<<<grid, GS_THREADS, GS_SMEM, st>>>(p);
  return cudaGetLastError();
}

}  // namespace catseg
