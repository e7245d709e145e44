// CLIP dense last block: the producer of img_feats on the other side of the boundary.
//
// Reference: ResidualAttentionBlock.forward_dense (cat_seg/third_party/model_vpt.py:219-240) followed by the dense branch
// of VisualTransformer.forward (:300-312): x.permute(1,0,2) -> ln_post -> @ proj.
//
//   y  = ln_1(x)                                         x [L][N][D]  (LND, the transformer's layout)
//   v  = out_proj(in_proj_v(y))                          the q / k thirds of in_proj are computed and DISCARDED by the
//                                                        reference (:221-233): only the v third (attn.v_proj_weight) is used here
//   v  = v + x[0]                                        the CLS row of every image, broadcast over the L tokens (:235)
//   v  = v + c_proj(QuickGELU(c_fc(ln_2(v))))            (:237)
//   v  = cat(v[0:1], v[prompt+1:])                       prompt tokens dropped (:238-239); every op above is token-wise and
//                                                        x[0] is kept, so the dropped rows are simply never computed
//   f  = ln_post(v.permute(1,0,2)) @ proj                [N][L'][out_dim]
//
// The five linears run on the fp32-accurate tcgen05 GEMM (gemm_split.cu: hi + lo fp16 operand pairs, fp32 accumulate) with
// bias / QuickGELU / residual epilogues; the three LayerNorms are warp-per-row fp32 kernels, the first and last of which
// also do the row re-mapping (prompt drop, LND -> NLD transpose) so that no permute copy exists.
#include "common.cuh"
#include "internal.h"

namespace catseg {

namespace {

// out row r <- LayerNorm(in row map(r)), eps 1e-5, biased variance, two-pass (mean, then centred sum of squares).
//   mode 0: rows are (l', n) in LND order; source l = l' == 0 ? 0 : l' + prompt
//   mode 1: out rows are (n, l') (NLD), source row (l', n)
__global__ void layernorm_rows_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ g,
                                      const float* __restrict__ b, long long rows, int D, int N, int Lp, int prompt, int mode) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  long long src;
  if (mode == 0) {
    const long long lp = r / N, n = r % N;
    src = (lp == 0 ? 0 : lp + prompt) * N + n;
  } else {
    const long long n = r / Lp, lp = r % Lp;
    src = lp * N + n;
  }
  const float* x = in + src * D;
  float* y = out + r * D;
  float s = 0.0f;
  if ((D & 3) == 0) {
    for (int i = lane * 4; i < D; i += 128) { const float4 v = ld4(x + i); s += (v.x + v.y) + (v.z + v.w); }
  } else {
    for (int i = lane; i < D; i += 32) s += x[i];
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.0f;
  if ((D & 3) == 0) {
    for (int i = lane * 4; i < D; i += 128) {
      const float4 v = ld4(x + i);
      const float a = v.x - mean, c = v.y - mean, d = v.z - mean, e = v.w - mean;
      q += (a * a + c * c) + (d * d + e * e);
    }
  } else {
    for (int i = lane; i < D; i += 32) { const float a = x[i] - mean; q += a * a; }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + 1e-5f);
  if ((D & 3) == 0) {
    for (int i = lane * 4; i < D; i += 128) {
      const float4 v = ld4(x + i), gg = ld4(g + i), bb = ld4(b + i);
      st4(y + i, make_float4((v.x - mean) * rstd * gg.x + bb.x, (v.y - mean) * rstd * gg.y + bb.y,
                             (v.z - mean) * rstd * gg.z + bb.z, (v.w - mean) * rstd * gg.w + bb.w));
    }
  } else {
    for (int i = lane; i < D; i += 32) y[i] = (x[i] - mean) * rstd * g[i] + b[i];
  }
}

cudaError_t launch_layernorm_rows(const float* in, float* out, const float* g, const float* b, long long rows, int D, int N,
                                  int Lp, int prompt, int mode, cudaStream_t st) {
  layernorm_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(in, out, g, b, rows, D, N, Lp, prompt, mode);
  return cudaGetLastError();
}

}  // namespace

size_t clip_dense_workspace_floats(int L, int N, int D, int prompt) {
  const size_t M = (size_t)(L - prompt) * N;
  return M * D * 4 + M * (size_t)(4 * D) + 64;           // y | a | v1 | v2 | h
}

#define CKD(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

cudaError_t run_clip_dense_block(const ClipDenseW& w, const float* x, int L, int N, int prompt, float* block_out,
                                 float* feats_out, float* ws, cudaStream_t st) {
  const int D = w.width, Lp = L - prompt;
  const long long M = (long long)Lp * N;
  float* y = ws;
  float* a = y + M * D;
  float* v1 = a + M * D;
  float* v2 = block_out != nullptr ? block_out : v1 + M * D;
  float* hbuf = v1 + 2 * M * D;
  auto gemm = [&](const float* A, int K, const float* Bw, long long b_row, long long b_k, const float* bias, int Nout, float* C,
                  int act, const float* res, long long r_row, int r_mod) {
    GemmSplitParams g{};
    g.A = A; g.a_row = K; g.a_k = 1; g.a_batch = 0;
    g.B = Bw; g.b_row = b_row; g.b_k = b_k; g.b_batch = 0;
    g.C = C; g.c_row = Nout; g.c_batch = 0;
    g.M = (int)M; g.N = Nout; g.K = K; g.batch = 1;
    g.bias = bias; g.act = act;
    g.residual = res; g.r_row = r_row; g.r_batch = 0; g.r_mod = r_mod;
    return launch_gemm_split(g, st);
  };
  CKD(launch_layernorm_rows(x, y, w.ln1_g, w.ln1_b, M, D, N, Lp, prompt, 0, st));                         // :220
  CKD(gemm(y, D, w.v_w, D, 1, w.v_b, D, a, 0, nullptr, 0, 0));    // :221-229 (v third)
  CKD(gemm(a, D, w.out_proj_w, D, 1, w.out_proj_b, D, v1, 0, x, D, N));                                   // :230, :235 (+ x[0][n])
  CKD(launch_layernorm_rows(v1, y, w.ln2_g, w.ln2_b, M, D, N, Lp, 0, 0, st));                             // :237
  CKD(gemm(y, D, w.c_fc_w, D, 1, w.c_fc_b, 4 * D, hbuf, 2, nullptr, 0, 0));
  CKD(gemm(hbuf, 4 * D, w.c_proj_w, 4 * D, 1, w.c_proj_b, D, v2, 0, v1, D, 0));
  if (feats_out != nullptr) {
    CKD(launch_layernorm_rows(v2, y, w.ln_post_g, w.ln_post_b, M, D, N, Lp, 0, 1, st));                   // :301-307
    CKD(gemm(y, D, w.proj, 1, w.out_dim, nullptr, w.out_dim, feats_out, 0, nullptr, 0, 0));               // :309-310 (x @ proj)
  }
  return cudaSuccess;
}

}  // namespace catseg
