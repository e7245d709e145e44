/*
 * catseg_b200.h — C ABI of the B200-native CAT-Seg cost-aggregation hot path.
 *
 * The reference is pure Python and has no FFI; the operator this library replaces is
 *   Aggregator.forward(img_feats, text_feats, appearance_guidance) -> logits
 *   (/root/reference/cat_seg/modeling/transformer/model.py:683-725), constructed at
 *   cat_seg/modeling/transformer/cat_seg_predictor.py:97-113 and called at :161,
 * plus the sliding-window stitch + argmax of cat_seg/cat_seg_model.py:204-218 and
 * train_net.py:58.  A maintainer binds these entry points with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are contiguous fp32 in the reference's own layouts at the boundary:
 *       img_feats [B,C,H,W]   text_feats [B,T,P,C]   g0 [B,Cg,H,W]  g1 [B,D0,2H,2W]  g2 [B,D1,4H,4W]
 *       logits [B,T,4H,4W];
 *   - the caller owns every buffer including the workspace; the handle owns only packed weights;
 *   - calls are asynchronous on the given stream; a handle belongs to one device, one thread;
 *   - return value 0 = OK, negative = error (catseg_last_error gives the text).  There is no CPU
 *     fallback: unsupported configurations fail.
 */
#ifndef CATSEG_B200_H
#define CATSEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CATSEG_OK 0
#define CATSEG_ERR_INVALID (-1)     /* bad argument / shape */
#define CATSEG_ERR_UNSUPPORTED (-2) /* configuration outside what the kernels implement */
#define CATSEG_ERR_WORKSPACE (-3)   /* workspace too small */
#define CATSEG_ERR_CUDA (-4)        /* CUDA runtime / launch error */
#define CATSEG_ERR_WEIGHTS (-5)     /* unknown / missing / mis-sized parameter */

/* precision: 0 = EXACT (fp32 CUDA-core arithmetic end to end); otherwise a bit mask of the stages that
 * run on the tensor cores (tcgen05, fp16 operands, fp32 accumulate, fp32 residual stream).
 *   FAST    = one fp16 term per operand (11 significant bits).
 *   PRECISE = CATSEG_PRECISE_SPLIT set: the value path uses hi + lo fp16 operand pairs (3 MMAs per product,
 *             ~22 bits), the attention-weight path stays single (cat-seg_b200/csrc/split_common.cuh):
 *             the mode that meets north_star's >= 99.9 % argmax agreement with the fp32 reference.
 * CATSEG_PRECISION_FAST / CATSEG_PRECISION_PRECISE select every stage that has a kernel of that kind; a
 * stage without one runs the EXACT kernel. */
#define CATSEG_PRECISION_EXACT 0
#define CATSEG_FAST_SWIN_MLP 1   /* FFN half of the Swin blocks */
#define CATSEG_FAST_SWIN_ATTN 2  /* window-attention half of the Swin blocks */
#define CATSEG_FAST_CLASS 4      /* class-aggregation layers */
#define CATSEG_FAST_DECODER 8    /* upsampling decoder */
#define CATSEG_FAST_PREP 16      /* 7x7 cost embedding (fp32-accurate split) and the 3x3 guidance projections */
#define CATSEG_PRECISE_SPLIT 0x100 /* modifier: selected stages use hi+lo fp16 operand pairs */
#define CATSEG_PRECISION_FAST 0x1f
#define CATSEG_PRECISION_PRECISE (0x1f | CATSEG_PRECISE_SPLIT)

typedef struct catseg_handle catseg_handle;
typedef void* catseg_stream; /* cudaStream_t */

/* Mirrors Aggregator.__init__ kwargs (model.py:559-576). */
typedef struct catseg_config {
  int32_t text_guidance_dim;
  int32_t text_guidance_proj_dim;
  int32_t appearance_guidance_dim;
  int32_t appearance_guidance_proj_dim;
  int32_t decoder_dims[2];
  int32_t decoder_guidance_dims[2];
  int32_t decoder_guidance_proj_dims[2];
  int32_t num_layers;
  int32_t nheads;
  int32_t hidden_dim;
  int32_t pooling_size[2];
  int32_t feature_resolution[2];
  int32_t window_size;
  int32_t attention_type; /* 0 = "linear"; "full" is dead code in every shipped config */
  int32_t prompt_channel;
  int32_t pad_len;
  int32_t precision;      /* CATSEG_PRECISION_* */
} catseg_config;

/* Optional taps for stage-wise parity tests: any non-NULL pointer receives a copy of that stage's
 * output in the kernels' token-major layout.  l = layer index (0..num_layers-1, at most 4). */
typedef struct catseg_taps {
  float* corr;           /* [B,T,P,HW]      model.py:648-652 */
  int32_t* classes;      /* [B,Te]          kept class ids, ascending (model.py:694-696) */
  float* embed;          /* [B,Te,HW,hid]   model.py:654-659 */
  float* app_guidance;   /* [B,HW,Ag]       model.py:708 */
  float* text_guidance;  /* [B,Te,Tg]       model.py:712-715 */
  float* dec_guidance0;  /* [B,4HW,Dp0]     model.py:710 (NHWC) */
  float* dec_guidance1;  /* [B,16HW,Dp1] */
  float* swin_b1[4];     /* [B,Te,HW,hid]   after block_1 of layer l (model.py:250) */
  float* swin_b2[4];     /* after block_2 (model.py:251) */
  float* class_out[4];   /* after the class layer (model.py:494) */
  float* up1;            /* [B,Te,4HW,dec0]  after decoder1 (model.py:677) */
  float* up2;            /* [B,Te,16HW,dec1] after decoder2 (model.py:678) */
} catseg_taps;

/* Stage ids for catseg_stage_times */
enum {
  CATSEG_STAGE_PREP = 0,   /* normalise, cost volume, top-k, guidance/text projections */
  CATSEG_STAGE_EMBED = 1,  /* 7x7 cost embedding */
  CATSEG_STAGE_SWIN = 2,   /* window-attention half of all Swin blocks (whole blocks on the exact path) */
  CATSEG_STAGE_CLASS = 3,  /* all class-attention layers */
  CATSEG_STAGE_DECODER = 4,
  CATSEG_STAGE_SWIN_MLP = 5, /* FFN half of all Swin blocks when it runs as its own kernel (fast path) */
  CATSEG_STAGE_EXCHANGE = 6, /* class-sharded modes: the state all-reduce, or the peer transpositions + barriers */
  CATSEG_STAGE_COUNT = 7
};

int catseg_create(const catseg_config* cfg, catseg_handle** out);
int catseg_destroy(catseg_handle* h);
const char* catseg_last_error(const catseg_handle* h); /* h may be NULL: last create error */

/* Parameter table: names and sizes are exactly the reference state_dict. */
int catseg_num_params(const catseg_handle* h);
const char* catseg_param_name(const catseg_handle* h, int i);
int64_t catseg_param_numel(const catseg_handle* h, int i);
/* Copies `numel` fp32 values for parameter `name`; src may be a host or a device pointer. */
int catseg_set_param(catseg_handle* h, const char* name, const float* src, int64_t numel, int src_is_device);
/* Packs every parameter into kernel layouts; fails if a parameter is missing. */
int catseg_finalize_params(catseg_handle* h, catseg_stream stream);

/* Persistent per-vocabulary text object (cat_seg_predictor.py:190-224 caches the class embeddings
 * of a vocabulary, model.py:650,701,712-715 re-derives from them on every call).  text_feats [T,P,C] fp32 (device) is
 * copied; its normalisation scales, the text-guidance projection and the guidance half of the class-attention q/k
 * projections are derived once (re-derived after a weight update).  Afterwards catseg_forward* may be called with
 * text_feats == NULL and the same T: per call the kept classes only gather their rows.  T = 0 forgets the vocabulary. */
int catseg_set_vocabulary(catseg_handle* h, const float* text_feats, int T, catseg_stream stream);

/* Number of kept classes for T input classes: min(T, pad_len) when pad_len > 0 (model.py:694). */
int catseg_kept_classes(const catseg_handle* h, int T);
size_t catseg_workspace_bytes(const catseg_handle* h, int B, int T);

/* The boundary call: Aggregator.forward (model.py:683-725). */
int catseg_forward(catseg_handle* h, const float* img_feats, const float* text_feats, const float* g0,
                   const float* g1, const float* g2, float* logits, void* workspace, size_t workspace_bytes,
                   int B, int T, catseg_stream stream);
/* Same, copying intermediates into `taps` (may be NULL). */
int catseg_forward_taps(catseg_handle* h, const float* img_feats, const float* text_feats, const float* g0,
                        const float* g1, const float* g2, float* logits, void* workspace,
                        size_t workspace_bytes, int B, int T, const catseg_taps* taps, catseg_stream stream);

/* Class-sharded multi-GPU mode (spatial aggregation and the decoder are independent per class; the
 * class layers couple classes only through the linear-attention state, model.py:282-283).  Every rank of the shard
 * group receives the SAME images and text; rank r computes the kept classes [r*Te/world, (r+1)*Te/world) of the
 * ascending kept-class list (Te = catseg_kept_classes(T) must be a multiple of world).  Between the state and apply
 * kernels of each class layer the library calls `allreduce(ctx, state, count, stream)`, which must sum `count` fp32
 * values in place over the shard group, ordered on `stream` (e.g. ncclAllReduce / torch.distributed.all_reduce).
 *   logits_local     [B, Te/world, 4H, 4W] fp32: this rank's kept classes, compact.
 *   kept_classes_out [B, Te] int32 (may be NULL): the kept class ids, identical on every rank; the caller all-gathers
 *                    logits_local over the group and scatters plane (r, j) to class kept[b][r*Te/world + j]; classes not
 *                    kept are -100 (model.py:721-724). */
typedef int (*catseg_allreduce_fn)(void* ctx, float* buf, size_t count, catseg_stream stream);
int catseg_forward_class_sharded(catseg_handle* h, const float* img_feats, const float* text_feats, const float* g0,
                                 const float* g1, const float* g2, float* logits_local, int32_t* kept_classes_out,
                                 void* workspace, size_t workspace_bytes, int B, int T, int shard_rank, int shard_world,
                                 catseg_allreduce_fn allreduce, void* ctx, catseg_stream stream);

/* The same class split with north_star's exchange: instead of summing the state, the residual stream is
 * transposed class-sharded [B, Te/world, HW, 128] <-> pixel-sharded [B, Te, HW/world, 128] around each class layer
 * (model.py:404-413 couples all classes of one pixel).  The transposition is one kernel per direction that stores straight
 * into the owning peer's buffer over NVLink: every rank allocates two buffers of catseg_exchange_buffer_bytes() with
 * catseg_peer_alloc, publishes them with catseg_peer_export (64-byte CUDA IPC handles, exchanged by the host, e.g.
 * torch.distributed.all_gather_object) and maps the others' with catseg_peer_open.
 *   xbuf_peers / pbuf_peers  [world] device pointers (HOST arrays), entry r = rank r's class-sharded / pixel-sharded buffer
 *                            (entry shard_rank = this rank's own allocation).
 *   barrier(ctx, stream)     NULL: the library synchronises the group itself with a flag barrier through the peer buffers
 *                            (one single-warp kernel per barrier).  Otherwise it must enqueue, on `stream`, a barrier over
 *                            the shard group's streams (e.g. a one-element ncclAllReduce).  A barrier orders the peer
 *                            stores against their consumers: one after the sharded class selection, two per class layer.
 * With T > pad_len the first cost-volume pass is sharded over the raw classes too: every rank reduces the maxima of T/world
 * classes and stores them into every rank's table (model.py:695-696 then selects from identical tables).
 * PRECISE precision, pooling_size [1,1], HW % world == 0, world <= 8. */
/*   logits_peers             NULL: the result is logits_local (compact planes, as above).  Otherwise [world] device pointers
 *                            (HOST array) to every rank's FULL logits buffer [B, T, 4H, 4W] (catseg_exchange_logits_bytes,
 *                            catseg_peer_alloc): the head kernel stores each of this rank's planes into all of them at its
 *                            class id, the library pre-fills its own buffer with -100 and ends with a barrier -- every rank
 *                            holds the complete logits when the call's work on `stream` is done, with no all-gather and no
 *                            assembly pass.  logits_local may then be NULL. */
typedef int (*catseg_barrier_fn)(void* ctx, catseg_stream stream);
size_t catseg_exchange_buffer_bytes(const catseg_handle* h, int B, int T, int shard_world);
size_t catseg_exchange_logits_bytes(const catseg_handle* h, int B, int T);
/*   guidance_peers           NULL, or [world] device pointers (HOST array) to every rank's guidance buffer
 *                            (catseg_exchange_guidance_bytes, catseg_peer_alloc): the class-independent front end (guidance
 *                            projections, Swin guidance terms, decoder additive maps) is then sharded by IMAGE -- rank r
 *                            computes images [r B/world, (r+1) B/world) and pushes its slices to every peer -- instead of
 *                            being repeated for all B images on every rank.  Needs B % world == 0 and barrier == NULL. */
size_t catseg_exchange_guidance_bytes(const catseg_handle* h, int B, int T);
/* *timed_out = 1 if one of this rank's flag barriers ever gave up waiting (~4 s: a peer died): results since then are invalid.
 * pbuf_self / gbuf_self: this rank's own pixel-sharded and (optional) guidance buffers.  Synchronous; not for the hot path. */
int catseg_exchange_timed_out(const catseg_handle* h, const float* pbuf_self, const float* gbuf_self, int B, int T,
                              int shard_world, int* timed_out);
int catseg_forward_class_sharded_a2a(catseg_handle* h, const float* img_feats, const float* text_feats, const float* g0,
                                     const float* g1, const float* g2, float* logits_local, int32_t* kept_classes_out,
                                     void* workspace, size_t workspace_bytes, int B, int T, int shard_rank, int shard_world,
                                     float* const* xbuf_peers, float* const* pbuf_peers, size_t buf_bytes,
                                     float* const* logits_peers, float* const* guidance_peers, size_t guidance_bytes,
                                     catseg_barrier_fn barrier, void* ctx, catseg_stream stream);
/* Final assembly after the all-gather of the local planes: gathered [world][B][T_local][npix] (rank-major), kept_classes
 * [B][world*T_local] -> logits [B][T][npix] with -100 for classes that were not kept (model.py:721-724).  pos_scratch: B*T
 * int32 of device scratch.  npix % 4 == 0, B*T <= 65535. */
int catseg_assemble_class_sharded(const float* gathered, const int32_t* kept_classes, int32_t* pos_scratch, float* logits,
                                  int shard_world, int B, int T_local, int T, int64_t npix, catseg_stream stream);
/* Peer-visible device memory (cudaMalloc + CUDA IPC): alloc / free on the owner, export -> 64-byte handle, open / close on
 * the other processes of the node (peer access is enabled lazily by the open). */
int catseg_peer_alloc(size_t bytes, void** ptr_out);
int catseg_peer_free(void* ptr);
int catseg_peer_export(const void* ptr, uint8_t handle_out[64]);
int catseg_peer_open(const uint8_t handle[64], void** ptr_out);
int catseg_peer_close(void* ptr);

/* Per-stage CUDA-event timing of catseg_forward on its own stream (for the roofline report). */
int catseg_set_profiling(catseg_handle* h, int enable);
/* Synchronises the recorded events; ms[CATSEG_STAGE_COUNT] = total ms per stage since the last
 * reset, *calls = number of forwards accumulated. */
int catseg_stage_times(catseg_handle* h, float* ms, int* calls, int reset);
/* Number of kernel launches issued by the last catseg_forward. */
int catseg_last_launch_count(const catseg_handle* h);

/* Sliding-window stitch (cat_seg_model.py:204-218) + postprocess resize + argmax (train_net.py:58).
 *   win_logits [nwin+1, T, S, S]: nwin tile windows in row-major tile order, then the global view.
 *   kernel/stride/out_res as at cat_seg_model.py:158-164 (384 / 256 / 640).
 *   probs_out (optional) [T, height, width] fp32: the 'sem_seg' tensor; labels_out (optional)
 *   [height, width] int32: argmax over classes, first maximum wins.  No [T,kernel,kernel] or
 *   [T,out_res,out_res] intermediate is materialised. */
int catseg_stitch_argmax(const float* win_logits, int T, int S, int kernel, int stride, int out_res,
                         int height, int width, float* probs_out, int32_t* labels_out,
                         catseg_stream stream);
/* Same result, faster when classes were dropped (T > pad_len): with T uint32 of caller-owned scratch the library first marks
 * the (window, class) planes that hold -100.0 everywhere -- the Aggregator's encoding of a dropped class, whose sigmoid is
 * exactly 0.0 -- and skips them (bit-identical output; ~70 % of the planes at T = 847). */
size_t catseg_stitch_scratch_bytes(int T);
int catseg_stitch_argmax_ws(const float* win_logits, int T, int S, int kernel, int stride, int out_res, int height,
                            int width, float* probs_out, int32_t* labels_out, void* scratch, size_t scratch_bytes,
                            catseg_stream stream);
/* Plain per-pixel argmax over the class axis of [T, H*W] (first maximum wins). */
int catseg_argmax(const float* scores, int T, int64_t npix, int32_t* labels_out, catseg_stream stream);
/* Same for a batch of independent [T, npix] score sets laid out [batch, T, npix] -> labels [batch, npix] (one launch). */
int catseg_argmax_batched(const float* scores, int batch, int T, int64_t npix, int32_t* labels_out, catseg_stream stream);

/* Guidance pyramid producers, the step before the boundary (cat_seg_model.py:80-82 modules, :176-185 use).
 *   catseg_guidance_upsample: nn.ConvTranspose2d(width, cout, kernel_size=k, stride=k) applied to a hooked CLIP layer
 *   output.  tokens [1 + grid*grid][B][width] fp32 (row 0 = CLS, skipped: the "(H W) B C -> B C H W" rearrange of
 *   :183-184 is index arithmetic), weight [width][cout][k][k] and bias [cout] exactly as the module stores them,
 *   out [B][cout][grid*k][grid*k] fp32 NCHW = the g1 / g2 argument of catseg_forward.  fp32 arithmetic.
 *   catseg_strip_cls_nchw: clip_features[:, 1:, :] + "B (H W) C -> B C H W" (:179,182 and cat_seg_head.py:2009):
 *   feats [B][1 + grid*grid][C] -> out [B][C][grid][grid] = the img_feats / g0 argument. */
int catseg_guidance_upsample(const float* tokens, const float* weight, const float* bias, float* out, int B, int width,
                             int cout, int kernel, int grid, catseg_stream stream);
int catseg_strip_cls_nchw(const float* feats, float* out, int B, int C, int grid, catseg_stream stream);

/* CLIP dense last block, the producer of img_feats on the other side of the boundary:
 * ResidualAttentionBlock.forward_dense (cat_seg/third_party/model_vpt.py:219-240) and the dense branch of
 * VisualTransformer.forward after the transformer (:300-312: permute, ln_post, @ proj).
 *   weights: device fp32 pointers in the reference's own parameter layouts, named after the state_dict keys
 *            (transformer.resblocks.<last>.{ln_1,attn.v_proj_weight,attn.in_proj_bias,attn.out_proj,ln_2,mlp.c_fc,mlp.c_proj},
 *            ln_post, proj); of the attention input projection only the v third is read (the reference computes q and k and
 *            discards them, :231-233): v_proj_weight = attn.v_proj_weight (:177; rows [2 width, 3 width) of a stock CLIP
 *            in_proj_weight), v_proj_bias = attn.in_proj_bias + 2 width.
 *   x          [L][N][width] fp32: the input of the last resblock (LND).
 *   prompt     number of VPT prompt tokens after CLS that the block drops (:238-239; 0 in the shipped configs).
 *   block_out  optional [L - prompt][N][width]: forward_dense's return value.
 *   feats_out  optional [N][L - prompt][out_dim]: ln_post + proj of it = `clip_features` of cat_seg_model.py:176.
 * All GEMMs run on tcgen05 with hi + lo fp16 operand pairs (fp32-accurate).  Stateless: no handle. */
typedef struct catseg_clip_dense_weights {
  int32_t width, out_dim;
  const float *ln_1_weight, *ln_1_bias;
  const float *v_proj_weight, *v_proj_bias;       /* [width][width], [width] */
  const float *out_proj_weight, *out_proj_bias;   /* [width][width], [width] */
  const float *ln_2_weight, *ln_2_bias;
  const float *c_fc_weight, *c_fc_bias;           /* [4 width][width], [4 width] */
  const float *c_proj_weight, *c_proj_bias;       /* [width][4 width], [width] */
  const float *ln_post_weight, *ln_post_bias;     /* may be NULL when feats_out is NULL */
  const float* proj;                              /* [width][out_dim] */
} catseg_clip_dense_weights;
size_t catseg_clip_dense_workspace_bytes(int L, int N, int width, int prompt);
int catseg_clip_dense_last_block(const catseg_clip_dense_weights* w, const float* x, int L, int N, int prompt,
                                 float* block_out, float* feats_out, void* workspace, size_t workspace_bytes,
                                 catseg_stream stream);

/* Library identity, e.g. "catseg_b200 0.1 sm_100a". */
const char* catseg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CATSEG_B200_H */
