// Internal declarations shared by the CAT-Seg B200 translation units (not part of the C ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace catseg {

// Packed parameters of one Swin block (model.py:117-225).  All fp32 device pointers.
struct SwinBlockW {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const float* wqkv_t;   // [128][384]  LN(x)-part of q | k | v, transposed so lanes sweep outputs
  const float* bv;       // [128]
  const float* wproj_t;  // [128][128]
  const float* bproj;    // [128]
  const float* w1_t;     // [128][512]
  const float* b1;       // [512]
  const float* w2_t;     // [512][128]
  const float* b2;       // [128]
  const float* wg_qk_t;  // [Ag][256]   guidance-part of q | k (class independent)
  const float* bqk;      // [256]       q bias | k bias (folded into the guidance term)
};

// Packed parameters of one class-aggregation layer (model.py:357-424).
struct ClassLayerW {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const float* wqkv_t;   // [128][384]
  const float* bv;       // [128]
  const float* wg_qk_t;  // [Tg][256]
  const float* bqk;      // [256]
  const float* w1_t;     // [128][512]
  const float* b1;
  const float* w2_t;     // [512][128]
  const float* b2;
  const float* pad_tok;  // [128]  padding_tokens
  const float* pad_g;    // [Tg]   padding_guidance
};

// FAST path: bf16 weight images for the token MLP (fast_mlp.cu).  wimg = 8 canonical 128x128 images
// in ring order W1_0, W2_0, W1_1, W2_1, ... (hidden chunks of 128).
struct MlpFastW {
  const __half* wimg;
  const float *ln_g, *ln_b, *b1, *b2;
};

// PRECISE path (split_mlp.cu): 16 fp16 images in consumption order
//   W1h0 W1l0 W1h1 W1l1 W2h0 W2l0 W1h2 W1l2 W2h1 W2l1 W1h3 W1l3 W2h2 W2l2 W2h3 W2l3   (h = hi, l = lo term)
struct MlpSplitW {
  const __half* wimg;
  const float *ln_g, *ln_b, *b1, *b2;
};

// FAST window attention (fast_swin_attn.cu): 5 images per block: Wqkv_h (h = 0..3), Wproj.
struct SwinAttnFastW {
  const __half* wimg;
  const float *ln_g, *ln_b, *bv, *bproj;
};

// Second-generation window attention (swin_attn2.cu), FAST and PRECISE: fp16 images Wq | Wk | Wv hi | Wv lo (4 x 32 KiB),
// then per head the proj slice hi | lo (4 x 16 KiB).
struct SwinAttn2W {
  const __half* wimg;
  const float *ln_g, *ln_b, *bv, *bproj;
};
constexpr size_t kSwinAttn2Halfs = 4 * 128 * 128 + 4 * 8192;

constexpr int kMaxShard = 8;
struct PeerPtrs { float* p[kMaxShard]; };      // one buffer per rank of the shard group (peer-mapped device pointers)

constexpr int kStateFloats = 4 * 32 * 32 + 128;   // per (image, pixel): KV[4][32][32] then Ksum[128]

// ---------------------------------------------------------------- prep.cu
cudaError_t launch_normalize_img(const float* img, float* out, int B, int C, int HW, cudaStream_t st);
cudaError_t launch_normalize_rows(const float* in, float* out, long long rows, int C, cudaStream_t st);
// text_batch_stride = 0: one set of class embeddings for every image (vocabulary mode)
cudaError_t launch_cost_volume(const float* textn, long long text_batch_stride, const float* imgn, float* corr, int B, int TP,
                               int C, int HW, cudaStream_t st);
cudaError_t launch_iota_range(int32_t* ids, int B, int Te, int offset, cudaStream_t st);
cudaError_t launch_inv_norm_rows(const float* in, float* out, long long rows, int C, cudaStream_t st);
cudaError_t launch_inv_norm_pixels(const float* img, float* out, int B, int C, int HW, cudaStream_t st);
cudaError_t launch_gather_rows(const float* src, const int32_t* idx, float* dst, long long n, int width, cudaStream_t st);
cudaError_t launch_class_max(const float* corr, float* cmax, long long rows, int n, cudaStream_t st);
cudaError_t launch_select_classes(const float* cmax, int32_t* classes, int B, int T, int Te, cudaStream_t st);
cudaError_t launch_text_mean(const float* src, const int32_t* classes, float* out, int B, int T, int Te, int P,
                             int C, cudaStream_t st);
cudaError_t launch_layernorm128(const float* in, float* out, const float* g, const float* b, long long rows,
                                cudaStream_t st);
cudaError_t launch_linear(const float* A, const float* Wt, const float* bias, float* out, long long M, int N,
                          int K, int relu, cudaStream_t st);
// 3x3 conv, NCHW input -> NHWC output, + bias + ReLU (guidance projections, model.py:615-630)
cudaError_t launch_conv3x3_nchw(const float* in, const float* Wt, const float* bias, float* out, int B, int Ci,
                                int H, int W, int Co, cudaStream_t st);
// 7x7 cost embedding (model.py:654-659): corr [B,T,P,HW] + classes -> X [B,Te,HW,128]
cudaError_t launch_cost_embed(const float* corr, const int32_t* classes, const float* Wt, const float* bias,
                              float* X, int B, int T, int Te, int P, int H, int W, cudaStream_t st);
cudaError_t launch_transpose_pack(float* dst, int ldd, int dst_col0, const float* src, int lds, int src_col0,
                                  int rows, int cols, cudaStream_t st);
cudaError_t launch_fill(float* p, float v, long long n, cudaStream_t st);
cudaError_t launch_iota_classes(int32_t* classes, int B, int Te, cudaStream_t st);
// dst[b][j] = src[b][first + j], j < n (class-sharded mode: this rank's slice of the kept-class list)
cudaError_t launch_slice_classes(const int32_t* src, int32_t* dst, int B, int Te, int first, int n, cudaStream_t st);

// ---------------------------------------------------------------- gemm_split.cu
// C[b][m][n] = epilogue(sum_k A[b][rowidx(m)][k] * B[b][n][k]), fp32 in / out, hi+lo fp16 operand pairs on tcgen05 (fp32-accurate)
struct GemmSplitParams {
  const float* A; long long a_row, a_k, a_batch;     // element strides (floats)
  const float* B; long long b_row, b_k, b_batch;
  float* C; long long c_row, c_batch;                // may be nullptr (only row_max wanted)
  int M, N, K, batch;
  const int32_t* a_index; long long ai_batch;        // optional row indirection of A (and of row_scale): row m -> a_index[b][m]
  const float* bias;                                 // [N]
  const float* row_scale; long long rs_batch;        // [batch][rows of A]
  const float* col_scale; long long cs_batch;        // [batch][N]
  const float* residual; long long r_row, r_batch;   // added after the activation; r_row = 0 broadcasts one row
  int r_mod;                                         // > 0: the residual row is m % r_mod (a block of rows broadcast cyclically)
  int act;                                           // 0 none, 1 ReLU, 2 QuickGELU
  float* row_max;                                    // [batch][M][2 * ceil(N / 128)] partial maxima over 64-column spans
};
cudaError_t launch_gemm_split(const GemmSplitParams& p, cudaStream_t st);

// ---------------------------------------------------------------- shard_exchange.cu
struct PeerFlags { uint32_t* p[kMaxShard]; };   // every rank's barrier flag block (kMaxShard + 2 words)
cudaError_t launch_peer_barrier(const PeerFlags& f, int rank, int world, cudaStream_t st);
// copies nseg segments (float offsets / counts relative to the buffer base, multiples of 4) of this rank's buffer to every peer
struct PeerSegs { long long off[8]; long long n[8]; int nseg; };
cudaError_t launch_peer_bcast(const PeerPtrs& bufs, const PeerSegs& segs, int rank, int world, cudaStream_t st);
cudaError_t launch_shard_put_cmax(const float* loc, const PeerPtrs& dst, int B, int Tr, int T, int t0, int world, cudaStream_t st);
cudaError_t launch_shard_c2p(const float* X, const PeerPtrs& pb, int B, int Tl, int Te, int HW, int rank, int world, cudaStream_t st);
cudaError_t launch_shard_p2c(const float* P, const PeerPtrs& xb, int B, int Tl, int Te, int HW, int rank, int world, cudaStream_t st);

cudaError_t launch_assemble_class_sharded(const float* gathered, const int32_t* kept, int32_t* pos_scratch, float* out, int world,
                                          int B, int Tl, int T, long long npix, cudaStream_t st);

// ---------------------------------------------------------------- clip_dense.cu
// CLIP ResidualAttentionBlock.forward_dense + ln_post + proj (model_vpt.py:219-240, 300-312); all fp32 device pointers in the
// reference's own parameter layouts
struct ClipDenseW {
  int width, out_dim;
  const float *ln1_g, *ln1_b;
  const float *v_w, *v_b;                  // [width][width], [width]: the v third of the attention input projection
  const float *out_proj_w, *out_proj_b;    // [width][width], [width]
  const float *ln2_g, *ln2_b;
  const float *c_fc_w, *c_fc_b;            // [4 width][width], [4 width]
  const float *c_proj_w, *c_proj_b;        // [width][4 width], [width]
  const float *ln_post_g, *ln_post_b;
  const float* proj;                       // [width][out_dim]
};
size_t clip_dense_workspace_floats(int L, int N, int D, int prompt);
// x [L][N][width]; block_out [L - prompt][N][width] (optional); feats_out [N][L - prompt][out_dim] (optional)
cudaError_t run_clip_dense_block(const ClipDenseW& w, const float* x, int L, int N, int prompt, float* block_out,
                                 float* feats_out, float* ws, cudaStream_t st);

// ---------------------------------------------------------------- fast_prep.cu
// 7x7 cost embedding for P = 1 on a 24x24 grid; bimg = 14 images packed by launch_pack_embed_img from Wt [49][128]
cudaError_t launch_pack_embed_img(__nv_bfloat16* dst, const float* Wt, cudaStream_t st);
cudaError_t launch_cost_embed_fast(const float* corr, const int32_t* classes, const __nv_bfloat16* bimg, const float* bias,
                                   float* X, int B, int T, int Te, int num_sms, cudaStream_t st);
// 3x3 guidance projections: which = 0 (24^2 -> 128), 1 (48^2 -> 32), 2 (96^2 -> 16); NCHW fp32 in, NHWC fp32 out
int gconv_fast_kc(int which);
bool gconv_fast_supported(int which, int Ci, int H, int W, int Co);
cudaError_t launch_pack_gconv_img(__nv_bfloat16* dst, const float* Wt, int Ci, int Co, int KC, cudaStream_t st);
cudaError_t launch_gconv_fast(int which, const float* in, const __nv_bfloat16* wimg, const float* bias, float* out, int B,
                              int Ci, cudaStream_t st);
// PRECISE: fp16 hi+lo operand pairs (fp32-accurate); images from launch_pack_gconv_img_split with KC = gconv_split_kc(which)
int gconv_split_kc(int which);
cudaError_t launch_pack_gconv_img_split(__half* dst, const float* Wt, int Ci, int Co, int KC, cudaStream_t st);
cudaError_t launch_gconv_split(int which, const float* in, const __half* wimg, const float* bias, float* out, int B, int Ci,
                               cudaStream_t st);

// ---------------------------------------------------------------- swin_exact.cu
// with_mlp = 0 stops after x1 = shortcut + proj(attn) (the FFN half then runs in fast_mlp.cu)
cudaError_t launch_swin_block_exact(float* X, const float* ag_qk, int nslice, int Te, int shift,
                                    const SwinBlockW& w, int with_mlp, cudaStream_t st);

// ---------------------------------------------------------------- fast_mlp.cu
cudaError_t launch_mlp_fast(float* X, long long ntok, const MlpFastW& w, int act, int num_sms, cudaStream_t st);
cudaError_t launch_pack_wimg(__half* dst, const float* W, int ld, int r0, int c0, cudaStream_t st);

// ---------------------------------------------------------------- split_mlp.cu (PRECISE: fp16 hi + lo operands)
// Xout[row] = Xin[row] (+ Xres[row]) + fc2(act(fc1(LN(Xin[row]))));  Xres may be nullptr, Xout may alias Xin
cudaError_t launch_mlp_split(const float* Xin, const float* Xres, float* Xout, long long ntok, const MlpSplitW& w, int act,
                             int num_sms, cudaStream_t st);
cudaError_t launch_pack_wimg_split(__half* dhi, __half* dlo, const float* W, int ld, int r0, int c0, cudaStream_t st);
// packs the 16 images of MlpSplitW from fc1 [512][128] and fc2 [128][512]
cudaError_t pack_mlp_split(__half* dst, const float* W1, const float* W2, cudaStream_t st);

// ---------------------------------------------------------------- fast_swin_attn.cu
// agw: bf16 guidance tiles [B][4 windows][4 heads][144 tok][64] of this block (launch_pack_ag_windows, same shift)
cudaError_t launch_swin_attn_fast(float* X, const __half* agw, int nslice, int Te, int shift,
                                  const SwinAttnFastW& w, int num_sms, cudaStream_t st);
cudaError_t launch_pack_ag_windows(const float* ag_qk, __half* out, int B, int shift, cudaStream_t st);
cudaError_t launch_pack_qkv_head_img(__half* dst, const float* Wq, const float* Wk, const float* Wv, int ldqk,
                                     int h, cudaStream_t st);

// ---------------------------------------------------------------- swin_attn2.cu
// agT: fp32 guidance terms [B][4 windows][18 token groups][2 halves][256 features][4 tok] of this block (launch_pack_ag_windows_T, same shift)
cudaError_t launch_swin_attn2(float* X, const float* agT, int nslice, int Te, int shift, const SwinAttn2W& w, bool split,
                              int num_sms, cudaStream_t st);
cudaError_t launch_pack_ag_windows_T(const float* ag_qk, float* out, int B, int shift, cudaStream_t st);
cudaError_t pack_swin_attn2(__half* dst, const float* Wq, const float* Wk, const float* Wv, const float* Wp, int ldqk,
                            cudaStream_t st);

// ---------------------------------------------------------------- class_exact.cu
cudaError_t launch_class_pad_state(const ClassLayerW& w, int Tg, float* pad_state, int n_pad, int S,
                                   cudaStream_t st);
cudaError_t launch_class_state_exact(const float* X, const float* cg_qk, float* state, int B, int Te, int npix,
                                     int S, const ClassLayerW& w, cudaStream_t st);
// out_mode 0: Xout[b][t][p] = x + x_pool'   (pool 1x1: model.py:423 fused);  1: Xout = x_pool'
cudaError_t launch_class_apply_exact(const float* X, float* Xout, const float* cg_qk, const float* state,
                                     const float* pad_state, int B, int Te, int npix, int S, int out_mode,
                                     const ClassLayerW& w, cudaStream_t st);
cudaError_t launch_avgpool_tokens(const float* X, float* Xp, long long nslice, int H, int W, int ph, int pw,
                                  cudaStream_t st);
cudaError_t launch_upsample_add(float* X, const float* Xp, long long nslice, int H, int W, int Hp, int Wp,
                                cudaStream_t st);

// ---------------------------------------------------------------- decoder_exact.cu
struct DecoderW {
  const float* up1_wt;   // [128][4*96]   n = (dy*2+dx)*96 + co
  const float* up1_b;    // [96]
  const float* c1a_wt;   // [9*128][64]   k = tap*128 + ci
  const float *gn1a_g, *gn1a_b;
  const float* c1b_wt;   // [9*64][64]
  const float *gn1b_g, *gn1b_b;
  const float* up2_wt;   // [64][4*48]
  const float* up2_b;
  const float* c2a_wt;   // [9*64][32]
  const float *gn2a_g, *gn2a_b;
  const float* c2b_wt;   // [9*32][32]
  const float *gn2b_g, *gn2b_b;
  const float* head_w;   // [9][32]
  const float* head_b;   // [1]
};
struct DecoderDims { int H, W, C0, U1, G1, D1, U2, G2, D2; };  // C0=128 U1=96 G1=32 D1=64 U2=48 G2=16 D2=32
size_t decoder_exact_scratch_floats(const DecoderDims& d, int chunk);
// X [nslice][HW][C0]; dg0 [B][4HW][G1]; dg1 [B][16HW][G2]; logits [B][T][16HW] written at class ids.
cudaError_t run_decoder_exact(const float* X, const float* dg0, const float* dg1, const int32_t* classes,
                              float* logits, int B, int T, int Te, const DecoderDims& d, const DecoderW& w,
                              float* scratch, int chunk, float* tap_up1, float* tap_up2, int* launches,
                              cudaStream_t st);

// ---------------------------------------------------------------- fast_class.cu
struct ClassFastW {
  const __half* wimg_kv;      // 3 images: Wk (LN(x) part), Wk (guidance part), Wv
  const __half* wimg_apply;   // 10 images: Wq_x, Wq_g, then W1_j, W2_j for j = 0..3
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *bqk, *bv, *b1, *b2;
};
cudaError_t launch_class_state_fast(const float* X, const __half* timg, float* state, int B, int Te, int npix,
                                    int S, const ClassFastW& w, int num_sms, cudaStream_t st);
cudaError_t launch_class_apply_fast(const float* X, float* Xout, const __half* timg, const float* state,
                                    const float* pad_state, int B, int Te, int npix, int S, int out_mode,
                                    const ClassFastW& w, int num_sms, cudaStream_t st);
cudaError_t launch_pack_text_img(const float* tg, __half* timg, int B, int Te, cudaStream_t st);

// ---------------------------------------------------------------- split_class.cu (PRECISE)
struct ClassSplitW {
  const __half* wimg_kv;   // 4 images in ring order: Wk (LN(x) part), Wk (guidance part), Wv hi, Wv lo
  const __half* wimg_q;    // 2 images: Wq (LN(x) part), Wq (guidance part)
  const float *ln1_g, *ln1_b, *bqk, *bv;
};
cudaError_t launch_class_state_split(const float* X, const __half* timg, float* state, int B, int Te, int npix, int S,
                                     const ClassSplitW& w, int num_sms, cudaStream_t st);
// X1[b][t][p] = x + attention (model.py:412); the MLP half follows in launch_mlp_split
cudaError_t launch_class_apply_split(const float* X, float* X1, const __half* timg, const float* state, const float* pad_state,
                                     int B, int Te, int npix, int S, const ClassSplitW& w, int num_sms, cudaStream_t st);

// ---------------------------------------------------------------- fast_decoder.cu
struct DecoderFastW {
  const __half *w1, *w2, *w3, *w4, *w5;   // UMMA weight images per stage (fast_decoder.cu)
  const float *bmap1, *bmap2;                    // [4HW][D1], [16HW][D2] transposed-conv bias maps
  const float *wg1, *wg2;                        // [9*G1][D1], [9*G2][D2] guidance conv weights (fp32 GEMM)
};
size_t decoder_fast_weight_bytes(const DecoderDims& d, int nw);   // nw = 2: hi + lo images (PRECISE)
cudaError_t decoder_fast_pack(const DecoderDims& d, const float* up1_w, const float* up1_b, const float* c1a_w,
                              const float* c1b_w, const float* up2_w, const float* up2_b, const float* c2a_w,
                              const float* c2b_w, const float* head_w, void* storage, DecoderFastW* out, int nw,
                              cudaStream_t st);
size_t decoder_fast_scratch_bytes(const DecoderDims& d, int B, int chunk);
cudaError_t run_decoder_fast(const float* X, const float* dg0, const float* dg1, const int32_t* classes,
                             float* logits, int B, int T, int Te, const DecoderDims& d, const DecoderFastW& w,
                             const DecoderW& wx, float head_bias, void* scratch, int chunk, int num_sms,
                             int* launches, cudaStream_t st);

size_t decoder_split_scratch_bytes(const DecoderDims& d, int B, int chunk);
// additive-map area [E1 | E2 | E1t | E2t] for B images (the head of the decoder scratch, or a peer-visible buffer)
size_t decoder_split_emap_bytes(const DecoderDims& d, int B);
void decoder_split_emap_slices(const DecoderDims& d, int B, size_t* e1t_off, size_t* e2t_off, size_t* per1, size_t* per2);
// computes the maps of images [b0, b0 + Bl) (dg0 / dg1 point at image b0) into `area`
cudaError_t decoder_split_prepare(const float* dg0, const float* dg1, int B, int b0, int Bl, const DecoderDims& d,
                                  const DecoderFastW& w, void* area, int* launches, cudaStream_t st);
// lpeers (optional, class-sharded peer-direct mode): every rank's full logits buffer; the head kernel stores into all of them
cudaError_t run_decoder_split(const float* X, const float* dg0, const float* dg1, const int32_t* classes,
                              float* logits, int B, int T, int Te, const DecoderDims& d, const DecoderFastW& w,
                              const DecoderW& wx, float head_bias, void* scratch, int chunk, int num_sms,
                              int* launches, const PeerPtrs* lpeers, int nlp, const void* ext_area, cudaStream_t st);

// ---------------------------------------------------------------- stitch.cu
// scratch_mask: T uint32 of device scratch (enables the skipping of dropped-class planes), or nullptr
cudaError_t launch_stitch(const float* win_logits, int T, int S, int kernel, int stride, int out_res, int height,
                          int width, float* probs_out, int32_t* labels_out, uint32_t* scratch_mask, cudaStream_t st);
cudaError_t launch_argmax(const float* scores, int batch, int T, long long npix, int32_t* labels, cudaStream_t st);
// guidance_pyramid.cu: ConvTranspose2d(stride == kernel) from hooked CLIP tokens, CLS strip + NCHW transpose
cudaError_t launch_guidance_upsample(const float* tokens, const float* weight, const float* bias, float* out, int B,
                                     int width, int cout, int ks, int grid, cudaStream_t st);
cudaError_t launch_strip_cls_nchw(const float* feats, float* out, int B, int C, int grid, cudaStream_t st);

}  // namespace catseg
