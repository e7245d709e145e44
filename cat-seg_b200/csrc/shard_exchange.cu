// Class-sharded <-> pixel-sharded transposition of the residual stream over NVLink peer memory (SURVEY.md 8e row 3:
// north_star's all-to-all around each class-attention layer, model.py:404-413).
//
// Spatial aggregation and the decoder are independent per (image, class): rank r owns the kept classes
// [r Tl, (r+1) Tl) of every image as X_r [B][Tl][HW][128].  Class attention couples all classes of one pixel: rank r owns
// the pixels [r HW/G, (r+1) HW/G) of every image as P_r [B][Te][HW/G][128].  Both buffers of every rank are mapped into
// every process (CUDA IPC); the transposition is ONE kernel per direction whose stores go straight to the owning peer --
// 512-byte token rows, one warp per row, so every NVLink packet is a full 128-byte line.  No staging copy, no NCCL call on
// the data path; a stream-ordered barrier (host callback) separates the peer stores from their consumers.
#include "common.cuh"
#include "internal.h"

namespace catseg {

// X [B][Tl][HW][128] (this rank) -> P_dst[b][rank Tl + j][pix - dst npl][128] for dst = pix / npl
__global__ void shard_c2p_kernel(const float* __restrict__ X, PeerPtrs pb, int B, int Tl, int Te, int HW, int npl, int rank) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * Tl * HW) return;
  const int pix = (int)(row % HW);
  const long long bj = row / HW;
  const int j = (int)(bj % Tl), b = (int)(bj / Tl);
  const int dst = pix / npl;
  const float4 v = ld4(X + row * 128 + lane * 4);
  st4(pb.p[dst] + ((((long long)b * Te + rank * Tl + j) * npl) + (pix - dst * npl)) * 128 + lane * 4, v);
}

// P [B][Te][npl][128] (this rank) -> X_dst[b][t - dst Tl][rank npl + q][128] for dst = t / Tl
__global__ void shard_p2c_kernel(const float* __restrict__ P, PeerPtrs xb, int B, int Tl, int Te, int HW, int npl, int rank) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * Te * npl) return;
  const int q = (int)(row % npl);
  const long long bt = row / npl;
  const int t = (int)(bt % Te), b = (int)(bt / Te);
  const int dst = t / Tl;
  const float4 v = ld4(P + row * 128 + lane * 4);
  st4(xb.p[dst] + ((((long long)b * Tl + (t - dst * Tl)) * HW) + rank * npl + q) * 128 + lane * 4, v);
}

cudaError_t launch_shard_c2p(const float* X, const PeerPtrs& pb, int B, int Tl, int Te, int HW, int rank, int world, cudaStream_t st) {
  const long long rows = (long long)B * Tl * HW;
  shard_c2p_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(X, pb, B, Tl, Te, HW, HW / world, rank);
  return cudaGetLastError();
}
cudaError_t launch_shard_p2c(const float* P, const PeerPtrs& xb, int B, int Tl, int Te, int HW, int rank, int world, cudaStream_t st) {
  const int npl = HW / world;
  const long long rows = (long long)B * Te * npl;
  shard_p2c_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(P, xb, B, Tl, Te, HW, npl, rank);
  return cudaGetLastError();
}

}  // namespace catseg
