// fp32 CUDA-core implicit-GEMM template used by the EXACT precision path.
//
//   C[b][m][n] = sum_k A(b, m, k) * Bw[b*ldb_batch + k*N + n]
//
// A is supplied by a loader functor (dense rows, or an on-the-fly im2col gather), so the same tile
// engine serves the cost volume, the 7x7 cost embedding, the guidance projections, every Linear
// outside the fused kernels, the transposed convs and the 3x3 decoder convs.  The epilogue is a
// functor too (bias / ReLU / scatter).  64x64x16 tiles, 256 threads, 4x4 register micro-tile.
#pragma once
#include "common.cuh"

namespace catseg {

constexpr int IG_BM = 64, IG_BN = 64, IG_BK = 16, IG_THREADS = 256;

// ALoad: __device__ float operator()(int b, int m, int k) const   (must return 0 outside bounds)
//        static constexpr bool kMFastest : lanes sweep m (true) or k (false) for coalescing
// Epi:   __device__ void operator()(int b, int m, int n, float acc) const
template <class ALoad, class Epi>
__global__ void __launch_bounds__(IG_THREADS)
igemm_kernel(ALoad aload, const float* __restrict__ Bw, long long ldb_batch, int M, int N, int K, Epi epi) {
  __shared__ float As[IG_BK][IG_BM + 4];
  __shared__ float Bs[IG_BK][IG_BN + 4];
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * IG_BM, n0 = blockIdx.y * IG_BN;
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const float* Bb = Bw + (long long)b * ldb_batch;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += IG_BK) {
    // ---- A tile: 64 x 16 scalars, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + i * IG_THREADS;
      int mm, kk;
      if (ALoad::kMFastest) { mm = e & 63; kk = e >> 6; } else { kk = e & 15; mm = e >> 4; }
      int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? aload(b, m, k) : 0.0f;
    }
    // ---- B tile: 16 x 64 scalars
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + i * IG_THREADS;
      int nn = e & 63, kk = e >> 6;
      int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < K) ? __ldg(Bb + (long long)k * N + n) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < IG_BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 w = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) epi(b, m, n, acc[i][j]);
    }
  }
}

template <class ALoad, class Epi>
inline cudaError_t launch_igemm(const ALoad& al, const float* Bw, long long ldb_batch, int batch, int M, int N,
                                int K, const Epi& epi, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return cudaSuccess;
  dim3 grid((M + IG_BM - 1) / IG_BM, (N + IG_BN - 1) / IG_BN, batch);
  igemm_kernel<ALoad, Epi><<<grid, IG_THREADS, 0, st>>>(al, Bw, ldb_batch, M, N, K, epi);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- common loaders / epilogues
struct DenseA {            // A[b][m][k] row-major
  static constexpr bool kMFastest = false;
  const float* A; long long batch_stride; int lda;
  __device__ float operator()(int b, int m, int k) const {
    return __ldg(A + b * batch_stride + (long long)m * lda + k);
  }
};

struct BiasActStore {      // out[b][m][n] = act(acc + bias[n])
  float* out; long long batch_stride; int ldo; const float* bias; int relu;
  __device__ void operator()(int b, int m, int n, float acc) const {
    float v = acc + (bias ? __ldg(bias + n) : 0.0f);
    if (relu) v = fmaxf(v, 0.0f);
    out[b * batch_stride + (long long)m * ldo + n] = v;
  }
};

}  // namespace catseg
