// Guidance pyramid producers: the step immediately before the boundary operator.
//
// Reference: CATSeg.forward, cat_seg/cat_seg_model.py:176-185 with the modules built at :80-82
//   image_features = clip_features[:, 1:, :]                      -> res3 "B (H W) C -> B C H W"
//   res4 = upsample1(rearrange(layers[0][1:], "(H W) B C -> B C H W"))   ConvTranspose2d(width, 256, k2, s2)
//   res5 = upsample2(rearrange(layers[1][1:], "(H W) B C -> B C H W"))   ConvTranspose2d(width, 128, k4, s4)
// A transposed convolution whose stride equals its kernel has no overlapping taps: it is the GEMM
//   out[b][co][k y + a][k x + c] = bias[co] + sum_ci tok[1 + y G + x][b][ci] * W[ci][co][a][c]
// with M = B G^2 token rows read straight from the hooked CLIP layer output [1 + G^2][B][width] (CLS row skipped,
// no rearranged copy), K = width, N = cout k^2 (the ConvTranspose2d weight [ci][co][a][c] IS the row-major [K][N]
// operand) and a pixel-shuffle store into NCHW -- the layout catseg_forward takes for g1 / g2.
// fp32 on the CUDA cores through the implicit-GEMM tile engine (igemm.cuh): the producers stay exact in every
// precision mode; they are 0.6 + 1.2 GMAC per image against 370 GMAC for the boundary operator.
#include "igemm.cuh"
#include "internal.h"

namespace catseg {

namespace {
struct TokenRowsA {          // A(m, k) = tok[1 + pix][b][k],  m = b * npix + pix
  static constexpr bool kMFastest = false;
  const float* tok; int B, width, npix;
  __device__ float operator()(int, int m, int k) const {
    const int b = m / npix, pix = m - b * npix;
    return __ldg(tok + ((long long)(1 + pix) * B + b) * width + k);
  }
};
struct PixelShuffleStore {   // n = (co, a, c) -> out[b][co][ks y + a][ks x + c]
  float* out; const float* bias; int cout, ks, grid, npix;
  __device__ void operator()(int, int m, int n, float acc) const {
    const int b = m / npix, pix = m - b * npix, y = pix / grid, x = pix - y * grid;
    const int kk = ks * ks, co = n / kk, r = n - co * kk, a = r / ks, c = r - a * ks;
    const int W = grid * ks;
    out[(((long long)b * cout + co) * W + (y * ks + a)) * W + (x * ks + c)] = acc + __ldg(bias + co);
  }
};

// feats [B][1 + P][C] -> out [B][C][P] (CLS row dropped): 32 x 32 tiles through shared memory, both sides coalesced
__global__ void __launch_bounds__(256) strip_cls_nchw_kernel(const float* __restrict__ feats, float* __restrict__ out, int P, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = feats + ((long long)b * (1 + P) + 1) * C;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    tile[i][tx] = (p < P && c < C) ? __ldg(src + (long long)p * C + c) : 0.0f;
  }
  __syncthreads();
  float* dst = out + (long long)b * C * P;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (c < C && p < P) dst[(long long)c * P + p] = tile[tx][i];
  }
}
}  // namespace

cudaError_t launch_guidance_upsample(const float* tokens, const float* weight, const float* bias, float* out, int B,
                                     int width, int cout, int ks, int grid, cudaStream_t st) {
  const int npix = grid * grid;
  TokenRowsA a{tokens, B, width, npix};
  return launch_igemm(a, weight, 0, 1, B * npix, cout * ks * ks, width, PixelShuffleStore{out, bias, cout, ks, grid, npix}, st);
}

cudaError_t launch_strip_cls_nchw(const float* feats, float* out, int B, int C, int grid, cudaStream_t st) {
  const int P = grid * grid;
  dim3 g((P + 31) / 32, (C + 31) / 32, B);
  strip_cls_nchw_kernel<<<g, 256, 0, st>>>(feats, out, P, C);
  return cudaGetLastError();
}

}  // namespace catseg
