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
constexpr int XR = 4;      // rows per warp: XR independent 16-byte loads in flight per thread before the peer stores
__global__ void shard_c2p_kernel(const float* __restrict__ X, PeerPtrs pb, int B, int Tl, int Te, int HW, int npl, int rank) {
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * XR;
  const int lane = threadIdx.x & 31;
  const long long nrows = (long long)B * Tl * HW;
  float4 v[XR];
#pragma unroll
  for (int i = 0; i < XR; ++i)
    if (row0 + i < nrows) v[i] = ld4(X + (row0 + i) * 128 + lane * 4);
#pragma unroll
  for (int i = 0; i < XR; ++i) {
    const long long row = row0 + i;
    if (row >= nrows) break;
    const int pix = (int)(row % HW);
    const long long bj = row / HW;
    const int j = (int)(bj % Tl), b = (int)(bj / Tl);
    const int dst = pix / npl;
    st4(pb.p[dst] + ((((long long)b * Te + rank * Tl + j) * npl) + (pix - dst * npl)) * 128 + lane * 4, v[i]);
  }
}

// P [B][Te][npl][128] (this rank) -> X_dst[b][t - dst Tl][rank npl + q][128] for dst = t / Tl
__global__ void shard_p2c_kernel(const float* __restrict__ P, PeerPtrs xb, int B, int Tl, int Te, int HW, int npl, int rank) {
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * XR;
  const int lane = threadIdx.x & 31;
  const long long nrows = (long long)B * Te * npl;
  float4 v[XR];
#pragma unroll
  for (int i = 0; i < XR; ++i)
    if (row0 + i < nrows) v[i] = ld4(P + (row0 + i) * 128 + lane * 4);
#pragma unroll
  for (int i = 0; i < XR; ++i) {
    const long long row = row0 + i;
    if (row >= nrows) break;
    const int q = (int)(row % npl);
    const long long bt = row / npl;
    const int t = (int)(bt % Te), b = (int)(bt / Te);
    const int dst = t / Tl;
    st4(xb.p[dst] + ((((long long)b * Tl + (t - dst * Tl)) * HW) + rank * npl + q) * 128 + lane * 4, v[i]);
  }
}

// ---- device-side barrier over the shard group through flags in peer memory.  Every rank's flag block is
//   uint32 arrived[kMaxShard] (slot r is written by rank r), uint32 epoch (local counter), uint32 timed_out.
// The ranks issue the same sequence of barriers, so their epochs agree; flags only grow, so a fast peer that is already in
// barrier e + 1 does not confuse a slow one.  Runs as one warp; a rank that waits longer than ~4 s gives up and sets
// timed_out (the result is then garbage, but the GPU is not hung).
__global__ void peer_barrier_kernel(PeerFlags f, int rank, int world) {
  volatile uint32_t* mine = f.p[rank];
  __shared__ uint32_t s_e;
  if (threadIdx.x == 0) { s_e = mine[kMaxShard] + 1; mine[kMaxShard] = s_e; }
  __syncwarp();
  const uint32_t e = s_e;
  const int r = threadIdx.x;
  if (r < world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[r] + rank), "r"(e) : "memory");
    const long long t0 = clock64();
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.p[rank] + r) : "memory");
      if ((int)(v - e) >= 0) break;
      if (clock64() - t0 > 8000000000LL) { mine[kMaxShard + 1] = 1u; break; }
      __nanosleep(200);
    } while (true);
  }
}
cudaError_t launch_peer_barrier(const PeerFlags& f, int rank, int world, cudaStream_t st) {
  peer_barrier_kernel<<<1, 32, 0, st>>>(f, rank, world);
  return cudaGetLastError();
}

// segments of this rank's buffer -> the same offsets of every other rank's buffer (peer stores, 16 bytes per thread and step)
__global__ void peer_bcast_kernel(PeerPtrs bufs, PeerSegs segs, int rank, int world) {
  const float4* src = reinterpret_cast<const float4*>(bufs.p[rank]);
  for (int sgi = 0; sgi < segs.nseg; ++sgi) {
    const long long o4 = segs.off[sgi] >> 2, n4 = segs.n[sgi] >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = src[o4 + i];
      for (int r = 0; r < world; ++r)
        if (r != rank) reinterpret_cast<float4*>(bufs.p[r])[o4 + i] = v;
    }
  }
}
cudaError_t launch_peer_bcast(const PeerPtrs& bufs, const PeerSegs& segs, int rank, int world, cudaStream_t st) {
  peer_bcast_kernel<<<296, 256, 0, st>>>(bufs, segs, rank, world);
  return cudaGetLastError();
}

// local per-class maxima [B][Tr] -> columns [t0, t0 + Tr) of every rank's [B][T] table
__global__ void shard_put_cmax_kernel(const float* __restrict__ loc, PeerPtrs dst, int B, int Tr, int T, int t0, int world) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Tr) return;
  const int b = i / Tr, j = i - b * Tr;
  const float v = loc[i];
  for (int r = 0; r < world; ++r) dst.p[r][(long long)b * T + t0 + j] = v;
}
cudaError_t launch_shard_put_cmax(const float* loc, const PeerPtrs& dst, int B, int Tr, int T, int t0, int world, cudaStream_t st) {
  shard_put_cmax_kernel<<<(B * Tr + 255) / 256, 256, 0, st>>>(loc, dst, B, Tr, T, t0, world);
  return cudaGetLastError();
}

cudaError_t launch_shard_c2p(const float* X, const PeerPtrs& pb, int B, int Tl, int Te, int HW, int rank, int world, cudaStream_t st) {
  const long long rows = (long long)B * Tl * HW;
  shard_c2p_kernel<<<(unsigned)((rows + 8 * XR - 1) / (8 * XR)), 256, 0, st>>>(X, pb, B, Tl, Te, HW, HW / world, rank);
  return cudaGetLastError();
}
cudaError_t launch_shard_p2c(const float* P, const PeerPtrs& xb, int B, int Tl, int Te, int HW, int rank, int world, cudaStream_t st) {
  const int npl = HW / world;
  const long long rows = (long long)B * Te * npl;
  shard_p2c_kernel<<<(unsigned)((rows + 8 * XR - 1) / (8 * XR)), 256, 0, st>>>(P, xb, B, Tl, Te, HW, npl, rank);
  return cudaGetLastError();
}


// ---- final assembly of the class-sharded logits (model.py:721-724): gathered [world][B][Tl][npix] (rank-major local planes),
// kept [B][world Tl] class ids -> out [B][T][npix], -100 for classes that were not kept.  One pass: every output plane is
// either a copy of one gathered plane or a constant fill.
__global__ void invert_kept_kernel(const int32_t* __restrict__ kept, int32_t* __restrict__ pos, int B, int Te, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Te) return;
  const int b = i / Te, c = kept[i];
  if (c >= 0 && c < T) pos[b * T + c] = i - b * Te;
}
__global__ void assemble_planes_kernel(const float* __restrict__ gathered, const int32_t* __restrict__ pos, float* __restrict__ out,
                                       int B, int Tl, int T, long long npix4) {
  const int plane = blockIdx.y;                        // b * T + t
  const int b = plane / T;
  const int j = pos[plane];
  float4* o = reinterpret_cast<float4*>(out) + (long long)plane * npix4;
  if (j < 0) {
    const float4 f = make_float4(-100.f, -100.f, -100.f, -100.f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix4; i += (long long)gridDim.x * blockDim.x) o[i] = f;
  } else {
    const int r = j / Tl, jl = j - r * Tl;
    const float4* g = reinterpret_cast<const float4*>(gathered) + (((long long)r * B + b) * Tl + jl) * npix4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix4; i += (long long)gridDim.x * blockDim.x) o[i] = g[i];
  }
}
cudaError_t launch_assemble_class_sharded(const float* gathered, const int32_t* kept, int32_t* pos_scratch, float* out, int world,
                                          int B, int Tl, int T, long long npix, cudaStream_t st) {
  if (npix % 4 || (long long)B * T > 65535) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(pos_scratch, 0xff, (size_t)B * T * sizeof(int32_t), st);
  if (e != cudaSuccess) return e;
  const int Te = world * Tl;
  invert_kept_kernel<<<(B * Te + 255) / 256, 256, 0, st>>>(kept, pos_scratch, B, Te, T);
  const long long npix4 = npix / 4;
  const int gx = (int)((npix4 + 1023) / 1024) < 4 ? (int)((npix4 + 1023) / 1024) : 4;
  assemble_planes_kernel<<<dim3(gx > 0 ? gx : 1, B * T), 256, 0, st>>>(gathered, pos_scratch, out, B, Tl, T, npix4);
  return cudaGetLastError();
}

}  // namespace catseg
