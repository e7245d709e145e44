// C ABI of the CAT-Seg B200 hot path: handle, parameter table, packing, workspace carving and the
// stage orchestration of Aggregator.forward (cat_seg/modeling/transformer/model.py:683-725).
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/catseg_b200.h"
#include "internal.h"

using namespace catseg;

namespace {

struct Param {
  std::string name;
  int64_t numel;
  size_t offset;   // floats into raw
  bool set;
};

thread_local std::string g_create_error;

constexpr int kImplementedFast = CATSEG_FAST_SWIN_MLP | CATSEG_FAST_SWIN_ATTN | CATSEG_FAST_DECODER | CATSEG_FAST_CLASS | CATSEG_FAST_PREP;   // stages that have a tcgen05 kernel in this build
// stages that have a PRECISE (hi + lo fp16 operand pair) kernel; the others run the EXACT kernel in that mode
constexpr int kImplementedSplit = CATSEG_FAST_SWIN_MLP | CATSEG_FAST_SWIN_ATTN | CATSEG_FAST_CLASS | CATSEG_FAST_DECODER | CATSEG_FAST_PREP;
constexpr int kMaxProfForwards = 64;
constexpr int kMaxSegments = 48;

}  // namespace

struct catseg_handle {
  catseg_config cfg;
  std::vector<Param> params;
  float* raw = nullptr;
  size_t raw_floats = 0;
  float* packed = nullptr;
  size_t packed_floats = 0;
  __half* wimg = nullptr;            // fp16 UMMA weight images (FAST path)
  size_t wimg_elems = 0;
  int fast_mask = 0;
  bool split = false;                // PRECISE: the stages in fast_mask run their hi+lo operand-pair kernels
  __half* wimg_split = nullptr;      // hi/lo fp16 weight images of the PRECISE kernels
  size_t wimg_split_elems = 0;
  std::vector<MlpSplitW> swin_mlp_split;   // [L*2]
  std::vector<SwinAttn2W> swin_attn2;      // [L*2]  second-generation window attention (both modes)
  __half* wimg_attn2 = nullptr;
  int attn_version = 1;                    // 1: fast_swin_attn.cu (FAST only), 2: swin_attn2.cu
  std::vector<ClassSplitW> class_split;    // [L]
  std::vector<MlpSplitW> class_mlp_split;  // [L]
  int num_sms = 148;
  std::vector<MlpFastW> swin_mlp_fast;   // [L*2]
  std::vector<SwinAttnFastW> swin_attn_fast;   // [L*2]
  std::vector<ClassFastW> class_fast;   // [L]
  void* dec_fast_store = nullptr;
  DecoderFastW dec_fast{};
  float head_bias_host = 0.0f;
  cudaStream_t aux_stream = nullptr;            // fork/join inside catseg_forward: guidance projections run beside the cost volume
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr, ev_sel = nullptr;   // join2: decoder guidance + additive maps ready
  __nv_bfloat16* prep_img = nullptr;            // FAST_PREP: embedding images, then the three guidance-conv image sets
  const __nv_bfloat16 *embed_img = nullptr, *gconv_img[3] = {nullptr, nullptr, nullptr};   // nullptr: shape not covered -> fp32 kernel
  __half* gconv_split_store = nullptr;          // PRECISE: fp16 [Wh | Wl] images of the three guidance convolutions
  const __half* gconv_split_img[3] = {nullptr, nullptr, nullptr};
  // persistent per-vocabulary text object (catseg_set_vocabulary; cat_seg_predictor.py:190-224 caches the class
  // embeddings once per vocabulary, the reference Aggregator then re-derives everything below from them on every call)
  struct Vocab {
    int T = 0;
    bool valid = false;        // false after a weight update: re-derived by the next forward
    float* store = nullptr;    // text [T,P,C] | textn [T,P,C] | inv_norm [T*P] | tmean [T,Ct] | text_g [T,128] | cg_qk [L][T,256]
    size_t floats = 0;
    float *text = nullptr, *textn = nullptr, *inv_norm = nullptr, *tmean = nullptr, *text_g = nullptr, *cg_qk = nullptr;
    int32_t* iota = nullptr;   // [T]
  } vocab;
  bool finalized = false;
  std::string err;
  int device = 0;

  // packed views
  std::vector<SwinBlockW> swin;     // [L*2]
  std::vector<ClassLayerW> cls;     // [L]
  std::vector<const float*> gnorm_g, gnorm_b;   // [L]
  DecoderW dec;
  const float *conv1_wt = nullptr, *conv1_b = nullptr;       // [P*49][128]
  const float *gproj_wt = nullptr, *gproj_b = nullptr;       // [Cg*9][Ag]
  const float *tproj_wt = nullptr, *tproj_b = nullptr;       // [Ct][Tg]
  const float *dgp_wt[2] = {nullptr, nullptr}, *dgp_b[2] = {nullptr, nullptr};

  // profiling
  bool profiling = false;
  std::vector<cudaEvent_t> ev;      // pairs
  std::vector<int> ev_stage;
  int ev_used = 0;                  // segments recorded
  int prof_forwards = 0;
  int last_launches = 0;
};

static int fail(catseg_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CUDA_OK(h, call)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(h, CATSEG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------------
static void add_param(catseg_handle* h, const std::string& name, int64_t numel) {
  Param p{name, numel, h->raw_floats, false};
  h->raw_floats += (size_t)((numel + 3) / 4 * 4);
  h->params.push_back(p);
}

static void build_param_table(catseg_handle* h) {
  const catseg_config& c = h->cfg;
  const int hid = c.hidden_dim, ag = c.appearance_guidance_proj_dim, tg = c.text_guidance_proj_dim;
  char b[160];
  for (int l = 0; l < c.num_layers; ++l) {
    for (int k = 1; k <= 2; ++k) {
      snprintf(b, sizeof(b), "layers.%d.swin_block.block_%d", l, k);
      std::string q(b);
      add_param(h, q + ".norm1.weight", hid); add_param(h, q + ".norm1.bias", hid);
      add_param(h, q + ".attn.q.weight", (int64_t)hid * (hid + ag)); add_param(h, q + ".attn.q.bias", hid);
      add_param(h, q + ".attn.k.weight", (int64_t)hid * (hid + ag)); add_param(h, q + ".attn.k.bias", hid);
      add_param(h, q + ".attn.v.weight", (int64_t)hid * hid); add_param(h, q + ".attn.v.bias", hid);
      add_param(h, q + ".attn.proj.weight", (int64_t)hid * hid); add_param(h, q + ".attn.proj.bias", hid);
      add_param(h, q + ".norm2.weight", hid); add_param(h, q + ".norm2.bias", hid);
      add_param(h, q + ".mlp.fc1.weight", (int64_t)4 * hid * hid); add_param(h, q + ".mlp.fc1.bias", 4 * hid);
      add_param(h, q + ".mlp.fc2.weight", (int64_t)4 * hid * hid); add_param(h, q + ".mlp.fc2.bias", hid);
    }
    snprintf(b, sizeof(b), "layers.%d.swin_block.guidance_norm", l);
    add_param(h, std::string(b) + ".weight", ag); add_param(h, std::string(b) + ".bias", ag);
    snprintf(b, sizeof(b), "layers.%d.attention", l);
    std::string a(b);
    if (c.pad_len > 0) {        // the reference registers them as None when pad_len == 0 (model.py:371-373): absent from its state_dict
      add_param(h, a + ".padding_tokens", hid);
      add_param(h, a + ".padding_guidance", tg);
    }
    add_param(h, a + ".attention.q.weight", (int64_t)hid * (hid + tg)); add_param(h, a + ".attention.q.bias", hid);
    add_param(h, a + ".attention.k.weight", (int64_t)hid * (hid + tg)); add_param(h, a + ".attention.k.bias", hid);
    add_param(h, a + ".attention.v.weight", (int64_t)hid * hid); add_param(h, a + ".attention.v.bias", hid);
    add_param(h, a + ".MLP.0.weight", (int64_t)4 * hid * hid); add_param(h, a + ".MLP.0.bias", 4 * hid);
    add_param(h, a + ".MLP.2.weight", (int64_t)4 * hid * hid); add_param(h, a + ".MLP.2.bias", hid);
    add_param(h, a + ".norm1.weight", hid); add_param(h, a + ".norm1.bias", hid);
    add_param(h, a + ".norm2.weight", hid); add_param(h, a + ".norm2.bias", hid);
  }
  add_param(h, "conv1.weight", (int64_t)hid * c.prompt_channel * 49); add_param(h, "conv1.bias", hid);
  add_param(h, "guidance_projection.0.weight", (int64_t)ag * c.appearance_guidance_dim * 9);
  add_param(h, "guidance_projection.0.bias", ag);
  add_param(h, "text_guidance_projection.0.weight", (int64_t)tg * c.text_guidance_dim);
  add_param(h, "text_guidance_projection.0.bias", tg);
  for (int i = 0; i < 2; ++i) {
    snprintf(b, sizeof(b), "decoder_guidance_projection.%d.0", i);
    add_param(h, std::string(b) + ".weight", (int64_t)c.decoder_guidance_proj_dims[i] * c.decoder_guidance_dims[i] * 9);
    add_param(h, std::string(b) + ".bias", c.decoder_guidance_proj_dims[i]);
  }
  int cin = hid;
  for (int i = 0; i < 2; ++i) {
    int cout = c.decoder_dims[i], gp = c.decoder_guidance_proj_dims[i];
    snprintf(b, sizeof(b), "decoder%d", i + 1);
    std::string d(b);
    add_param(h, d + ".up.weight", (int64_t)cin * (cin - gp) * 4); add_param(h, d + ".up.bias", cin - gp);
    add_param(h, d + ".conv.double_conv.0.weight", (int64_t)cout * cin * 9);
    add_param(h, d + ".conv.double_conv.1.weight", cout); add_param(h, d + ".conv.double_conv.1.bias", cout);
    add_param(h, d + ".conv.double_conv.3.weight", (int64_t)cout * cout * 9);
    add_param(h, d + ".conv.double_conv.4.weight", cout); add_param(h, d + ".conv.double_conv.4.bias", cout);
    cin = cout;
  }
  add_param(h, "head.weight", (int64_t)cin * 9); add_param(h, "head.bias", 1);
}

static const float* raw_of(const catseg_handle* h, const std::string& name) {
  for (const Param& p : h->params)
    if (p.name == name) return h->raw + p.offset;
  return nullptr;
}

// ------------------------------------------------------------------------------------------------
extern "C" const char* catseg_version(void) { return "catseg_b200 0.1 sm_100a"; }

extern "C" const char* catseg_last_error(const catseg_handle* h) {
  return h ? h->err.c_str() : g_create_error.c_str();
}

extern "C" int catseg_create(const catseg_config* cfg, catseg_handle** out) {
  if (!cfg || !out) return fail(nullptr, CATSEG_ERR_INVALID, "null argument");
  *out = nullptr;
  const catseg_config& c = *cfg;
  // The kernels are specialised for the shipped CAT-Seg geometry; anything else fails loudly.
  if (c.hidden_dim != 128 || c.nheads != 4)
    return fail(nullptr, CATSEG_ERR_UNSUPPORTED, "hidden_dim=%d nheads=%d: kernels are built for 128/4", c.hidden_dim, c.nheads);
  if (c.feature_resolution[0] != 24 || c.feature_resolution[1] != 24 || c.window_size != 12)
    return fail(nullptr, CATSEG_ERR_UNSUPPORTED, "feature_resolution %dx%d window %d: kernels are built for 24x24 / 12",
                c.feature_resolution[0], c.feature_resolution[1], c.window_size);
  if (c.attention_type != 0)
    return fail(nullptr, CATSEG_ERR_UNSUPPORTED, "attention_type 'full' is dead code in every shipped config and is not implemented");
  if (c.appearance_guidance_proj_dim != 128 || c.text_guidance_proj_dim != 128)
    return fail(nullptr, CATSEG_ERR_UNSUPPORTED, "guidance projection dims must be 128");
  if (c.num_layers < 1 || c.num_layers > 4) return fail(nullptr, CATSEG_ERR_UNSUPPORTED, "num_layers must be 1..4");
  if (c.pooling_size[0] < 1 || c.pooling_size[1] < 1 || 24 % c.pooling_size[0] || 24 % c.pooling_size[1])
    return fail(nullptr, CATSEG_ERR_UNSUPPORTED, "pooling_size must divide 24");
  if (c.prompt_channel < 1 || c.pad_len < 0) return fail(nullptr, CATSEG_ERR_INVALID, "bad prompt_channel / pad_len");
  if (c.decoder_dims[0] % 16 || c.decoder_dims[1] % 16 || c.decoder_dims[0] > 128 || c.decoder_dims[1] > 128)
    return fail(nullptr, CATSEG_ERR_UNSUPPORTED, "decoder_dims must be multiples of 16 (GroupNorm groups of 16)");
  if (c.decoder_guidance_proj_dims[0] % 4 || c.decoder_guidance_proj_dims[1] % 4)
    return fail(nullptr, CATSEG_ERR_UNSUPPORTED, "decoder_guidance_proj_dims must be multiples of 4");
  if (c.precision < 0) return fail(nullptr, CATSEG_ERR_INVALID, "precision must be 0 (exact) or a mask of CATSEG_FAST_*");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
    cudaGetLastError();
    return fail(nullptr, CATSEG_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  }
  catseg_handle* h = new catseg_handle();
  h->cfg = c;
  h->split = (c.precision & CATSEG_PRECISE_SPLIT) != 0;
  {
    const char* e = getenv("CATSEG_ATTN_V");          // A/B switch for the FAST mode; the PRECISE mode only exists in version 2
    h->attn_version = (h->split || (e && e[0] == '2')) ? 2 : 1;
  }
  h->fast_mask = c.precision & (h->split ? kImplementedSplit : kImplementedFast);
  cudaGetDevice(&h->device);
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
  build_param_table(h);
  if (cudaMalloc(&h->raw, h->raw_floats * sizeof(float)) != cudaSuccess) {
    delete h;
    return fail(nullptr, CATSEG_ERR_CUDA, "cudaMalloc of the parameter store failed");
  }
  cudaMemset(h->raw, 0, h->raw_floats * sizeof(float));
  // everything catseg_forward needs besides the caller's buffers is created here: the library does not allocate on the hot path
  if (cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_sel, cudaEventDisableTiming) != cudaSuccess) {
    catseg_destroy(h);
    return fail(nullptr, CATSEG_ERR_CUDA, "stream / event creation failed");
  }
  h->ev.resize((size_t)kMaxProfForwards * kMaxSegments * 2, nullptr);     // timing events are created by catseg_set_profiling
  *out = h;
  return CATSEG_OK;
}

extern "C" int catseg_destroy(catseg_handle* h) {
  if (!h) return CATSEG_OK;
  for (cudaEvent_t e : h->ev) if (e) cudaEventDestroy(e);
  if (h->wimg_split) cudaFree(h->wimg_split);
  if (h->wimg_attn2) cudaFree(h->wimg_attn2);
  if (h->gconv_split_store) cudaFree(h->gconv_split_store);
  if (h->vocab.store) cudaFree(h->vocab.store);
  if (h->vocab.iota) cudaFree(h->vocab.iota);
  if (h->raw) cudaFree(h->raw);
  if (h->packed) cudaFree(h->packed);
  if (h->wimg) cudaFree(h->wimg);
  if (h->dec_fast_store) cudaFree(h->dec_fast_store);
  if (h->prep_img) cudaFree(h->prep_img);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_join2) cudaEventDestroy(h->ev_join2);
  if (h->ev_sel) cudaEventDestroy(h->ev_sel);
  delete h;
  return CATSEG_OK;
}

extern "C" int catseg_num_params(const catseg_handle* h) { return h ? (int)h->params.size() : 0; }
extern "C" const char* catseg_param_name(const catseg_handle* h, int i) {
  return (h && i >= 0 && i < (int)h->params.size()) ? h->params[i].name.c_str() : nullptr;
}
extern "C" int64_t catseg_param_numel(const catseg_handle* h, int i) {
  return (h && i >= 0 && i < (int)h->params.size()) ? h->params[i].numel : -1;
}

extern "C" int catseg_set_param(catseg_handle* h, const char* name, const float* src, int64_t numel, int src_is_device) {
  if (!h || !name || !src) return fail(h, CATSEG_ERR_INVALID, "null argument");
  for (Param& p : h->params) {
    if (p.name != name) continue;
    if (p.numel != numel)
      return fail(h, CATSEG_ERR_WEIGHTS, "parameter %s: expected %lld values, got %lld", name, (long long)p.numel, (long long)numel);
    CUDA_OK(h, cudaMemcpy(h->raw + p.offset, src, (size_t)numel * sizeof(float),
                          src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    // a device-to-device cudaMemcpy does not block the host and the legacy stream is not ordered against the
    // (non-blocking) stream catseg_finalize_params packs on: wait here, so the caller may also free `src` on return
    if (src_is_device) CUDA_OK(h, cudaStreamSynchronize(cudaStreamLegacy));
    p.set = true;
    h->finalized = false;
    return CATSEG_OK;
  }
  return fail(h, CATSEG_ERR_WEIGHTS, "unknown parameter %s", name);
}

// ------------------------------------------------------------------------------------------------
namespace {

struct Packer {
  catseg_handle* h;
  cudaStream_t st;
  size_t used = 0;
  bool dry;
  cudaError_t err = cudaSuccess;
  float* alloc(size_t n) {
    size_t o = used;
    used += (n + 3) / 4 * 4;
    return dry ? nullptr : h->packed + o;
  }
  void note(cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; }
  const float* copy(const std::string& name, size_t n) {
    float* d = alloc(n);
    if (!dry) note(cudaMemcpyAsync(d, raw_of(h, name), n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return d;
  }
  // Linear weight [out][ld] columns [col0, col0+in) -> transposed into dst[in][ldd] at column dst_col0
  void tr(float* dst, int ldd, int dst_col0, const std::string& name, int lds, int src_col0, int out, int in) {
    if (!dry) note(launch_transpose_pack(dst, ldd, dst_col0, raw_of(h, name), lds, src_col0, out, in, st));
  }
};

__global__ void pack_conv3x3_kernel(float* dst, const float* src, int Co, int Ci) {
  // dst[(tap*Ci + ci)*Co + co] = src[((co*Ci + ci)*9) + tap]
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Co * Ci * 9) return;
  int co = i % Co, r = i / Co, ci = r % Ci, tap = r / Ci;
  dst[i] = src[((long long)co * Ci + ci) * 9 + tap];
}
__global__ void pack_convT_kernel(float* dst, const float* src, int Ci, int Co) {
  // dst[ci*(4*Co) + q*Co + co] = src[((ci*Co + co)*4) + q],  q = dy*2+dx
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ci * Co * 4) return;
  int co = i % Co, r = i / Co, q = r % 4, ci = r / 4;
  dst[i] = src[((long long)ci * Co + co) * 4 + q];
}

void pack_all(Packer& pk) {
  catseg_handle* h = pk.h;
  const catseg_config& c = h->cfg;
  const int ag = c.appearance_guidance_proj_dim, tg = c.text_guidance_proj_dim;
  char b[160];
  h->swin.assign(c.num_layers * 2, SwinBlockW{});
  h->cls.assign(c.num_layers, ClassLayerW{});
  h->gnorm_g.assign(c.num_layers, nullptr);
  h->gnorm_b.assign(c.num_layers, nullptr);
  for (int l = 0; l < c.num_layers; ++l) {
    for (int k = 0; k < 2; ++k) {
      snprintf(b, sizeof(b), "layers.%d.swin_block.block_%d", l, k + 1);
      std::string q(b);
      SwinBlockW& w = h->swin[l * 2 + k];
      w.ln1_g = pk.copy(q + ".norm1.weight", 128); w.ln1_b = pk.copy(q + ".norm1.bias", 128);
      w.ln2_g = pk.copy(q + ".norm2.weight", 128); w.ln2_b = pk.copy(q + ".norm2.bias", 128);
      float* wqkv = pk.alloc(128 * 384);
      pk.tr(wqkv, 384, 0, q + ".attn.q.weight", 128 + ag, 0, 128, 128);
      pk.tr(wqkv, 384, 128, q + ".attn.k.weight", 128 + ag, 0, 128, 128);
      pk.tr(wqkv, 384, 256, q + ".attn.v.weight", 128, 0, 128, 128);
      w.wqkv_t = wqkv;
      w.bv = pk.copy(q + ".attn.v.bias", 128);
      float* wp = pk.alloc(128 * 128);
      pk.tr(wp, 128, 0, q + ".attn.proj.weight", 128, 0, 128, 128);
      w.wproj_t = wp;
      w.bproj = pk.copy(q + ".attn.proj.bias", 128);
      float* w1 = pk.alloc(128 * 512);
      pk.tr(w1, 512, 0, q + ".mlp.fc1.weight", 128, 0, 512, 128);
      w.w1_t = w1;
      w.b1 = pk.copy(q + ".mlp.fc1.bias", 512);
      float* w2 = pk.alloc(512 * 128);
      pk.tr(w2, 128, 0, q + ".mlp.fc2.weight", 512, 0, 128, 512);
      w.w2_t = w2;
      w.b2 = pk.copy(q + ".mlp.fc2.bias", 128);
      float* wg = pk.alloc((size_t)ag * 256);
      pk.tr(wg, 256, 0, q + ".attn.q.weight", 128 + ag, 128, 128, ag);
      pk.tr(wg, 256, 128, q + ".attn.k.weight", 128 + ag, 128, 128, ag);
      w.wg_qk_t = wg;
      float* bqk = pk.alloc(256);
      if (!pk.dry) {
        pk.note(cudaMemcpyAsync(bqk, raw_of(h, q + ".attn.q.bias"), 512, cudaMemcpyDeviceToDevice, pk.st));
        pk.note(cudaMemcpyAsync(bqk + 128, raw_of(h, q + ".attn.k.bias"), 512, cudaMemcpyDeviceToDevice, pk.st));
      }
      w.bqk = bqk;
    }
    snprintf(b, sizeof(b), "layers.%d.swin_block.guidance_norm", l);
    h->gnorm_g[l] = pk.copy(std::string(b) + ".weight", ag);
    h->gnorm_b[l] = pk.copy(std::string(b) + ".bias", ag);
    snprintf(b, sizeof(b), "layers.%d.attention", l);
    std::string a(b);
    ClassLayerW& w = h->cls[l];
    w.ln1_g = pk.copy(a + ".norm1.weight", 128); w.ln1_b = pk.copy(a + ".norm1.bias", 128);
    w.ln2_g = pk.copy(a + ".norm2.weight", 128); w.ln2_b = pk.copy(a + ".norm2.bias", 128);
    float* wqkv = pk.alloc(128 * 384);
    pk.tr(wqkv, 384, 0, a + ".attention.q.weight", 128 + tg, 0, 128, 128);
    pk.tr(wqkv, 384, 128, a + ".attention.k.weight", 128 + tg, 0, 128, 128);
    pk.tr(wqkv, 384, 256, a + ".attention.v.weight", 128, 0, 128, 128);
    w.wqkv_t = wqkv;
    w.bv = pk.copy(a + ".attention.v.bias", 128);
    float* wg = pk.alloc((size_t)tg * 256);
    pk.tr(wg, 256, 0, a + ".attention.q.weight", 128 + tg, 128, 128, tg);
    pk.tr(wg, 256, 128, a + ".attention.k.weight", 128 + tg, 128, 128, tg);
    w.wg_qk_t = wg;
    float* bqk = pk.alloc(256);
    if (!pk.dry) {
      pk.note(cudaMemcpyAsync(bqk, raw_of(h, a + ".attention.q.bias"), 512, cudaMemcpyDeviceToDevice, pk.st));
      pk.note(cudaMemcpyAsync(bqk + 128, raw_of(h, a + ".attention.k.bias"), 512, cudaMemcpyDeviceToDevice, pk.st));
    }
    w.bqk = bqk;
    float* w1 = pk.alloc(128 * 512);
    pk.tr(w1, 512, 0, a + ".MLP.0.weight", 128, 0, 512, 128);
    w.w1_t = w1;
    w.b1 = pk.copy(a + ".MLP.0.bias", 512);
    float* w2 = pk.alloc(512 * 128);
    pk.tr(w2, 128, 0, a + ".MLP.2.weight", 512, 0, 128, 512);
    w.w2_t = w2;
    w.b2 = pk.copy(a + ".MLP.2.bias", 128);
    w.pad_tok = c.pad_len > 0 ? pk.copy(a + ".padding_tokens", 128) : nullptr;    // n_pad is 0 without them
    w.pad_g = c.pad_len > 0 ? pk.copy(a + ".padding_guidance", tg) : nullptr;
  }
  // conv1: [128][P*49] -> [P*49][128]
  {
    int K = c.prompt_channel * 49;
    float* d = pk.alloc((size_t)K * 128);
    pk.tr(d, 128, 0, "conv1.weight", K, 0, 128, K);
    h->conv1_wt = d;
    h->conv1_b = pk.copy("conv1.bias", 128);
  }
  {
    int K = c.appearance_guidance_dim * 9;
    float* d = pk.alloc((size_t)K * ag);
    pk.tr(d, ag, 0, "guidance_projection.0.weight", K, 0, ag, K);
    h->gproj_wt = d;
    h->gproj_b = pk.copy("guidance_projection.0.bias", ag);
  }
  {
    int K = c.text_guidance_dim;
    float* d = pk.alloc((size_t)K * tg);
    pk.tr(d, tg, 0, "text_guidance_projection.0.weight", K, 0, tg, K);
    h->tproj_wt = d;
    h->tproj_b = pk.copy("text_guidance_projection.0.bias", tg);
  }
  for (int i = 0; i < 2; ++i) {
    snprintf(b, sizeof(b), "decoder_guidance_projection.%d.0", i);
    int K = c.decoder_guidance_dims[i] * 9, N = c.decoder_guidance_proj_dims[i];
    float* d = pk.alloc((size_t)K * N);
    pk.tr(d, N, 0, std::string(b) + ".weight", K, 0, N, K);
    h->dgp_wt[i] = d;
    h->dgp_b[i] = pk.copy(std::string(b) + ".bias", N);
  }
  // decoder
  {
    DecoderW& d = h->dec;
    int cin = 128;
    for (int i = 0; i < 2; ++i) {
      int cout = c.decoder_dims[i], gp = c.decoder_guidance_proj_dims[i], up = cin - gp;
      snprintf(b, sizeof(b), "decoder%d", i + 1);
      std::string p(b);
      float* upw = pk.alloc((size_t)cin * 4 * up);
      if (!pk.dry) {
        int n = cin * up * 4;
        pack_convT_kernel<<<(n + 255) / 256, 256, 0, pk.st>>>(upw, raw_of(h, p + ".up.weight"), cin, up);
        pk.note(cudaGetLastError());
      }
      const float* upb = pk.copy(p + ".up.bias", up);
      float* ca = pk.alloc((size_t)9 * cin * cout);
      float* cb = pk.alloc((size_t)9 * cout * cout);
      if (!pk.dry) {
        int n = cout * cin * 9;
        pack_conv3x3_kernel<<<(n + 255) / 256, 256, 0, pk.st>>>(ca, raw_of(h, p + ".conv.double_conv.0.weight"), cout, cin);
        pk.note(cudaGetLastError());
        n = cout * cout * 9;
        pack_conv3x3_kernel<<<(n + 255) / 256, 256, 0, pk.st>>>(cb, raw_of(h, p + ".conv.double_conv.3.weight"), cout, cout);
        pk.note(cudaGetLastError());
      }
      const float* ga = pk.copy(p + ".conv.double_conv.1.weight", cout);
      const float* ba = pk.copy(p + ".conv.double_conv.1.bias", cout);
      const float* gb = pk.copy(p + ".conv.double_conv.4.weight", cout);
      const float* bb = pk.copy(p + ".conv.double_conv.4.bias", cout);
      if (i == 0) {
        d.up1_wt = upw; d.up1_b = upb; d.c1a_wt = ca; d.c1b_wt = cb;
        d.gn1a_g = ga; d.gn1a_b = ba; d.gn1b_g = gb; d.gn1b_b = bb;
      } else {
        d.up2_wt = upw; d.up2_b = upb; d.c2a_wt = ca; d.c2b_wt = cb;
        d.gn2a_g = ga; d.gn2a_b = ba; d.gn2b_g = gb; d.gn2b_b = bb;
      }
      cin = cout;
    }
    float* hw = pk.alloc((size_t)9 * cin);
    if (!pk.dry) {
      int n = cin * 9;
      pack_conv3x3_kernel<<<(n + 255) / 256, 256, 0, pk.st>>>(hw, raw_of(h, "head.weight"), 1, cin);
      pk.note(cudaGetLastError());
    }
    d.head_w = hw;
    d.head_b = pk.copy("head.bias", 1);
  }
}

}  // namespace

extern "C" int catseg_finalize_params(catseg_handle* h, catseg_stream stream) {
  if (!h) return CATSEG_ERR_INVALID;
  for (const Param& p : h->params)
    if (!p.set) return fail(h, CATSEG_ERR_WEIGHTS, "parameter %s was never set", p.name.c_str());
  cudaStream_t st = (cudaStream_t)stream;
  Packer dry{h, st, 0, true};
  pack_all(dry);
  if (!h->packed || h->packed_floats < dry.used) {
    if (h->packed) cudaFree(h->packed);
    h->packed = nullptr;
    CUDA_OK(h, cudaMalloc(&h->packed, dry.used * sizeof(float)));
    h->packed_floats = dry.used;
  }
  Packer pk{h, st, 0, false};
  pack_all(pk);
  if (pk.err != cudaSuccess) return fail(h, CATSEG_ERR_CUDA, "packing failed: %s", cudaGetErrorString(pk.err));
  // ---- bf16 UMMA weight images for the FAST kernels
  if (h->fast_mask && !h->split) {
    const int L = h->cfg.num_layers;
    const size_t kImg = 128 * 128;
    size_t need = (size_t)L * 2 * (8 + 5) * kImg + (size_t)L * 13 * kImg;
    if (!h->wimg || h->wimg_elems < need) {
      if (h->wimg) cudaFree(h->wimg);
      h->wimg = nullptr;
      CUDA_OK(h, cudaMalloc(&h->wimg, need * sizeof(__half)));
      h->wimg_elems = need;
    }
    h->swin_mlp_fast.assign(L * 2, MlpFastW{});
    h->swin_attn_fast.assign(L * 2, SwinAttnFastW{});
    const int ag = h->cfg.appearance_guidance_proj_dim;
    char b[160];
    for (int l = 0; l < L; ++l)
      for (int k = 0; k < 2; ++k) {
        snprintf(b, sizeof(b), "layers.%d.swin_block.block_%d", l, k + 1);
        std::string q(b);
        __half* img = h->wimg + (size_t)(l * 2 + k) * 13 * kImg;
        for (int j = 0; j < 4; ++j) {
          CUDA_OK(h, launch_pack_wimg(img + (size_t)(2 * j) * 128 * 128, raw_of(h, q + ".mlp.fc1.weight"), 128, j * 128, 0, st));
          CUDA_OK(h, launch_pack_wimg(img + (size_t)(2 * j + 1) * 128 * 128, raw_of(h, q + ".mlp.fc2.weight"), 512, 0, j * 128, st));
        }
        const SwinBlockW& sw = h->swin[l * 2 + k];
        h->swin_mlp_fast[l * 2 + k] = MlpFastW{img, sw.ln2_g, sw.ln2_b, sw.b1, sw.b2};
        __half* aimg = img + 8 * kImg;
        for (int hh = 0; hh < 4; ++hh)
          CUDA_OK(h, launch_pack_qkv_head_img(aimg + (size_t)hh * kImg, raw_of(h, q + ".attn.q.weight"),
                                              raw_of(h, q + ".attn.k.weight"), raw_of(h, q + ".attn.v.weight"),
                                              128 + ag, hh, st));
        CUDA_OK(h, launch_pack_wimg(aimg + 4 * kImg, raw_of(h, q + ".attn.proj.weight"), 128, 0, 0, st));
        h->swin_attn_fast[l * 2 + k] = SwinAttnFastW{aimg, sw.ln1_g, sw.ln1_b, sw.bv, sw.bproj};
      }
  }
  if ((h->fast_mask & CATSEG_FAST_SWIN_ATTN) && h->attn_version == 2) {
    const int L = h->cfg.num_layers, ag = h->cfg.appearance_guidance_proj_dim;
    if (!h->wimg_attn2) CUDA_OK(h, cudaMalloc(&h->wimg_attn2, (size_t)L * 2 * kSwinAttn2Halfs * sizeof(__half)));
    h->swin_attn2.assign(L * 2, SwinAttn2W{});
    char b[160];
    for (int l = 0; l < L; ++l)
      for (int k = 0; k < 2; ++k) {
        snprintf(b, sizeof(b), "layers.%d.swin_block.block_%d", l, k + 1);
        std::string q(b);
        __half* img = h->wimg_attn2 + (size_t)(l * 2 + k) * kSwinAttn2Halfs;
        CUDA_OK(h, pack_swin_attn2(img, raw_of(h, q + ".attn.q.weight"), raw_of(h, q + ".attn.k.weight"),
                                   raw_of(h, q + ".attn.v.weight"), raw_of(h, q + ".attn.proj.weight"), 128 + ag, st));
        const SwinBlockW& sw = h->swin[l * 2 + k];
        h->swin_attn2[l * 2 + k] = SwinAttn2W{img, sw.ln1_g, sw.ln1_b, sw.bv, sw.bproj};
      }
  }
  if (h->split && h->fast_mask) {
    const int L = h->cfg.num_layers;
    const size_t kImg = 128 * 128;
    const size_t need = (size_t)L * 2 * 16 * kImg + (size_t)L * 22 * kImg;
    if (!h->wimg_split || h->wimg_split_elems < need) {
      if (h->wimg_split) cudaFree(h->wimg_split);
      h->wimg_split = nullptr;
      CUDA_OK(h, cudaMalloc(&h->wimg_split, need * sizeof(__half)));
      h->wimg_split_elems = need;
    }
    h->swin_mlp_split.assign(L * 2, MlpSplitW{});
    char b[160];
    for (int l = 0; l < L; ++l)
      for (int k = 0; k < 2; ++k) {
        snprintf(b, sizeof(b), "layers.%d.swin_block.block_%d", l, k + 1);
        std::string q(b);
        __half* img = h->wimg_split + (size_t)(l * 2 + k) * 16 * kImg;
        CUDA_OK(h, pack_mlp_split(img, raw_of(h, q + ".mlp.fc1.weight"), raw_of(h, q + ".mlp.fc2.weight"), st));
        const SwinBlockW& sw = h->swin[l * 2 + k];
        h->swin_mlp_split[l * 2 + k] = MlpSplitW{img, sw.ln2_g, sw.ln2_b, sw.b1, sw.b2};
      }
    const int tg = h->cfg.text_guidance_proj_dim;
    h->class_split.assign(L, ClassSplitW{});
    h->class_mlp_split.assign(L, MlpSplitW{});
    for (int l = 0; l < L; ++l) {
      snprintf(b, sizeof(b), "layers.%d.attention", l);
      std::string a(b);
      __half* img = h->wimg_split + (size_t)L * 2 * 16 * kImg + (size_t)l * 22 * kImg;
      CUDA_OK(h, launch_pack_wimg_split(img + 0 * kImg, nullptr, raw_of(h, a + ".attention.k.weight"), 128 + tg, 0, 0, st));
      CUDA_OK(h, launch_pack_wimg_split(img + 1 * kImg, nullptr, raw_of(h, a + ".attention.k.weight"), 128 + tg, 0, 128, st));
      CUDA_OK(h, launch_pack_wimg_split(img + 2 * kImg, img + 3 * kImg, raw_of(h, a + ".attention.v.weight"), 128, 0, 0, st));
      CUDA_OK(h, launch_pack_wimg_split(img + 4 * kImg, nullptr, raw_of(h, a + ".attention.q.weight"), 128 + tg, 0, 0, st));
      CUDA_OK(h, launch_pack_wimg_split(img + 5 * kImg, nullptr, raw_of(h, a + ".attention.q.weight"), 128 + tg, 0, 128, st));
      CUDA_OK(h, pack_mlp_split(img + 6 * kImg, raw_of(h, a + ".MLP.0.weight"), raw_of(h, a + ".MLP.2.weight"), st));
      const ClassLayerW& cw = h->cls[l];
      h->class_split[l] = ClassSplitW{img, img + 4 * kImg, cw.ln1_g, cw.ln1_b, cw.bqk, cw.bv};
      h->class_mlp_split[l] = MlpSplitW{img + 6 * kImg, cw.ln2_g, cw.ln2_b, cw.b1, cw.b2};
    }
  }
  if ((h->fast_mask & CATSEG_FAST_CLASS) && !h->split) {
    const int L = h->cfg.num_layers, tg = h->cfg.text_guidance_proj_dim;
    const size_t kImg = 128 * 128;
    h->class_fast.assign(L, ClassFastW{});
    char b[160];
    for (int l = 0; l < L; ++l) {
      snprintf(b, sizeof(b), "layers.%d.attention", l);
      std::string a(b);
      __half* img = h->wimg + (size_t)L * 2 * 13 * kImg + (size_t)l * 13 * kImg;
      CUDA_OK(h, launch_pack_wimg(img + 0 * kImg, raw_of(h, a + ".attention.k.weight"), 128 + tg, 0, 0, st));
      CUDA_OK(h, launch_pack_wimg(img + 1 * kImg, raw_of(h, a + ".attention.k.weight"), 128 + tg, 0, 128, st));
      CUDA_OK(h, launch_pack_wimg(img + 2 * kImg, raw_of(h, a + ".attention.v.weight"), 128, 0, 0, st));
      __half* ap = img + 3 * kImg;
      CUDA_OK(h, launch_pack_wimg(ap + 0 * kImg, raw_of(h, a + ".attention.q.weight"), 128 + tg, 0, 0, st));
      CUDA_OK(h, launch_pack_wimg(ap + 1 * kImg, raw_of(h, a + ".attention.q.weight"), 128 + tg, 0, 128, st));
      for (int j = 0; j < 4; ++j) {
        CUDA_OK(h, launch_pack_wimg(ap + (size_t)(2 + 2 * j) * kImg, raw_of(h, a + ".MLP.0.weight"), 128, j * 128, 0, st));
        CUDA_OK(h, launch_pack_wimg(ap + (size_t)(3 + 2 * j) * kImg, raw_of(h, a + ".MLP.2.weight"), 512, 0, j * 128, st));
      }
      const ClassLayerW& cw = h->cls[l];
      h->class_fast[l] = ClassFastW{img, ap, cw.ln1_g, cw.ln1_b, cw.ln2_g, cw.ln2_b, cw.bqk, cw.bv, cw.b1, cw.b2};
    }
  }
  if (h->fast_mask & CATSEG_FAST_DECODER) {
    const catseg_config& c = h->cfg;
    const int nw = h->split ? 2 : 1;
    DecoderDims dd{c.feature_resolution[0], c.feature_resolution[1], 128, 128 - c.decoder_guidance_proj_dims[0],
                   c.decoder_guidance_proj_dims[0], c.decoder_dims[0], c.decoder_dims[0] - c.decoder_guidance_proj_dims[1],
                   c.decoder_guidance_proj_dims[1], c.decoder_dims[1]};
    if (dd.D1 != 64 || dd.D2 != 32)
      return fail(h, CATSEG_ERR_UNSUPPORTED, "the fast decoder is built for decoder_dims (64, 32)");
    if (!h->dec_fast_store) CUDA_OK(h, cudaMalloc(&h->dec_fast_store, decoder_fast_weight_bytes(dd, nw)));
    cudaError_t e = decoder_fast_pack(dd, raw_of(h, "decoder1.up.weight"), raw_of(h, "decoder1.up.bias"),
                                      raw_of(h, "decoder1.conv.double_conv.0.weight"),
                                      raw_of(h, "decoder1.conv.double_conv.3.weight"), raw_of(h, "decoder2.up.weight"),
                                      raw_of(h, "decoder2.up.bias"), raw_of(h, "decoder2.conv.double_conv.0.weight"),
                                      raw_of(h, "decoder2.conv.double_conv.3.weight"), raw_of(h, "head.weight"),
                                      h->dec_fast_store, &h->dec_fast, nw, st);
    if (e != cudaSuccess) return fail(h, CATSEG_ERR_CUDA, "decoder_fast_pack: %s", cudaGetErrorString(e));
    CUDA_OK(h, cudaMemcpyAsync(&h->head_bias_host, raw_of(h, "head.bias"), sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  if (h->fast_mask & CATSEG_FAST_PREP) {
    const catseg_config& c = h->cfg;
    const int H = c.feature_resolution[0], W = c.feature_resolution[1];
    const int ci[3] = {c.appearance_guidance_dim, c.decoder_guidance_dims[0], c.decoder_guidance_dims[1]};
    const int co[3] = {c.appearance_guidance_proj_dim, c.decoder_guidance_proj_dims[0], c.decoder_guidance_proj_dims[1]};
    const float* wt[3] = {h->gproj_wt, h->dgp_wt[0], h->dgp_wt[1]};
    size_t off[4] = {14 * 2048, 0, 0, 0}, total = 14 * 2048;
    for (int i = 0; i < 3; ++i) { off[i] = total; total += (size_t)ci[i] * 9 * co[i]; }
    if (!h->prep_img) CUDA_OK(h, cudaMalloc(&h->prep_img, total * sizeof(__nv_bfloat16)));
    h->embed_img = nullptr;
    if (H == 24 && W == 24 && c.prompt_channel == 1 && c.hidden_dim == 128) {
      CUDA_OK(h, launch_pack_embed_img(h->prep_img, h->conv1_wt, st));
      h->embed_img = h->prep_img;
    }
    for (int i = 0; i < 3; ++i) {
      h->gconv_img[i] = nullptr;
      if (ci[i] > 0 && co[i] > 0 && gconv_fast_supported(i, ci[i], H << i, W << i, co[i])) {
        CUDA_OK(h, launch_pack_gconv_img(h->prep_img + off[i], wt[i], ci[i], co[i], gconv_fast_kc(i), st));
        h->gconv_img[i] = h->prep_img + off[i];
      }
    }
    if (h->split) {
      size_t soff[4] = {0, 0, 0, 0};
      for (int i = 0; i < 3; ++i) soff[i + 1] = soff[i] + (size_t)2 * ci[i] * 9 * co[i];
      if (!h->gconv_split_store) CUDA_OK(h, cudaMalloc(&h->gconv_split_store, soff[3] * sizeof(__half)));
      for (int i = 0; i < 3; ++i) {
        h->gconv_split_img[i] = nullptr;
        if (ci[i] > 0 && co[i] > 0 && gconv_fast_supported(i, ci[i], H << i, W << i, co[i]) && ci[i] % gconv_split_kc(i) == 0) {
          CUDA_OK(h, launch_pack_gconv_img_split(h->gconv_split_store + soff[i], wt[i], ci[i], co[i], gconv_split_kc(i), st));
          h->gconv_split_img[i] = h->gconv_split_store + soff[i];
        }
      }
    }
  }
  CUDA_OK(h, cudaStreamSynchronize(st));
  h->finalized = true;
  h->vocab.valid = false;          // text_g / cg_qk depend on the weights: re-derived by the next forward
  return CATSEG_OK;
}

// ------------------------------------------------------------------------------------------------
namespace {

struct Plan {
  int B, T, Te, P, C, Ct, Cg, H, W, HW, Hp, Wp, npix, S, n_pad, L;
  bool truncated, pooled;
  int dec_chunk;
  DecoderDims dd;
  // workspace offsets (floats)
  size_t imgn, textn, corr, cmax, classes, tmean, text_g, cg_qk, pad_state, app_g, app_gn, ag_qk, dg0, dg1, X,
      Xp, Xp2, X1, state, timg, dec, agw, classes_loc, inv_t, inv_i, rmaxp, total;
};

Plan make_plan(const catseg_handle* h, int B, int T) {
  const catseg_config& c = h->cfg;
  Plan p{};
  p.B = B; p.T = T; p.P = c.prompt_channel; p.Ct = c.text_guidance_dim; p.C = c.text_guidance_dim;
  p.Cg = c.appearance_guidance_dim;
  p.H = c.feature_resolution[0]; p.W = c.feature_resolution[1]; p.HW = p.H * p.W; p.L = c.num_layers;
  p.truncated = c.pad_len > 0 && T > c.pad_len;
  p.Te = p.truncated ? c.pad_len : T;
  p.S = (c.pad_len > 0 && p.Te < c.pad_len) ? c.pad_len : p.Te;
  p.n_pad = p.S - p.Te;
  p.pooled = c.pooling_size[0] > 1 || c.pooling_size[1] > 1;
  p.Hp = p.H / c.pooling_size[0]; p.Wp = p.W / c.pooling_size[1]; p.npix = p.Hp * p.Wp;
  p.dd = DecoderDims{p.H, p.W, 128, 128 - c.decoder_guidance_proj_dims[0], c.decoder_guidance_proj_dims[0],
                     c.decoder_dims[0], c.decoder_dims[0] - c.decoder_guidance_proj_dims[1],
                     c.decoder_guidance_proj_dims[1], c.decoder_dims[1]};
  int nslice = B * p.Te;
  p.dec_chunk = nslice < 96 ? nslice : 96;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
  p.imgn = take((size_t)B * p.C * p.HW);
  p.textn = take((size_t)B * T * p.P * p.Ct);
  p.corr = take((size_t)B * T * p.P * p.HW);
  p.cmax = take((size_t)B * T);
  p.inv_t = take((size_t)B * T * p.P);                     // 1 / ||text row||   (tcgen05 cost volume: normalisation as scales)
  p.inv_i = take((size_t)B * p.HW);                        // 1 / ||img pixel||
  p.rmaxp = take((size_t)B * T * p.P * 2 * ((p.HW + 127) / 128));   // per-class partial maxima from the GEMM epilogue
  p.classes = take((size_t)B * p.Te);
  p.classes_loc = take((size_t)2 * B * p.Te);   // class-sharded mode: local slice of the kept list, then local plane ids
  p.tmean = take((size_t)B * p.Te * p.Ct);
  p.text_g = take((size_t)B * p.Te * 128);
  p.cg_qk = take((size_t)p.L * B * p.Te * 256);
  p.pad_state = take((size_t)p.L * kStateFloats);
  p.app_g = take((size_t)B * p.HW * 128);
  p.app_gn = take((size_t)B * p.HW * 128);
  p.ag_qk = take((size_t)p.L * 2 * B * p.HW * 256);
  // guidance terms in window order: version 1 = fp16 tiles [B][4][4 heads][144][64], version 2 = fp32 [B][4][256][144]
  p.agw = (h->fast_mask & CATSEG_FAST_SWIN_ATTN) ? take(h->attn_version == 2 ? (size_t)p.L * 2 * B * 4 * 256 * 144
                                                                               : (size_t)p.L * 2 * B * 16 * 144 * 64 / 2) : 0;
  p.dg0 = take((size_t)B * 4 * p.HW * p.dd.G1);
  p.dg1 = take((size_t)B * 16 * p.HW * p.dd.G2);
  p.X = take((size_t)nslice * p.HW * 128);
  if (p.pooled) {
    p.Xp = take((size_t)nslice * p.npix * 128);
    p.Xp2 = take((size_t)nslice * p.npix * 128);
  }
  if (h->split && (h->fast_mask & CATSEG_FAST_CLASS) && !p.pooled)
    p.X1 = take((size_t)nslice * p.npix * 128);            // PRECISE class layer: x1 = x + attention, input of its MLP kernel
  p.state = take((size_t)B * p.npix * kStateFloats);
  p.timg = take((size_t)B * ((p.Te + 127) / 128) * 8192);   // bf16 text-guidance images (FAST class path)
  {
    size_t fe = decoder_exact_scratch_floats(p.dd, p.dec_chunk);
    if (h->fast_mask & CATSEG_FAST_DECODER) {
      static const int chunk_env = [] { const char* e = getenv("CATSEG_DEC_CHUNK"); return e ? atoi(e) : 0; }();   // A/B
      // one chunk for up to 4096 slices (fp32 intermediates: 3.5 MB per slice, 14.5 GB at 4096 -- B200 has 180 GB): fewer launch
      // tails; smaller chunks that would keep the intermediates in the 126 MB L2 were measured SLOWER (148: 19.6 ms, 4096: 17.65 ms)
      const int want = chunk_env > 0 ? chunk_env : 4096;
      p.dec_chunk = nslice < want ? nslice : want;
      fe = ((h->split ? decoder_split_scratch_bytes(p.dd, B, p.dec_chunk) : decoder_fast_scratch_bytes(p.dd, B, p.dec_chunk)) + 3) / 4;
    }
    p.dec = take(fe);
  }
  p.total = o;
  return p;
}

struct Seg {   // RAII-less helper for per-stage event timing
  catseg_handle* h; cudaStream_t st; bool on;
  void begin(int stage) {
    if (!on) return;
    if (h->ev_used >= kMaxProfForwards * kMaxSegments) { on = false; return; }
    if ((int)h->ev_stage.size() <= h->ev_used) h->ev_stage.resize(h->ev_used + 1);
    h->ev_stage[h->ev_used] = stage;
    cudaEventRecord(h->ev[(size_t)h->ev_used * 2], st);
  }
  void end() {
    if (!on) return;
    cudaEventRecord(h->ev[(size_t)h->ev_used * 2 + 1], st);
    ++h->ev_used;
  }
};

}  // namespace

extern "C" int catseg_kept_classes(const catseg_handle* h, int T) {
  if (!h || T <= 0) return -1;
  return (h->cfg.pad_len > 0 && T > h->cfg.pad_len) ? h->cfg.pad_len : T;
}

extern "C" size_t catseg_workspace_bytes(const catseg_handle* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return make_plan(h, B, T).total * sizeof(float);
}

#define RUN(call)                                                                                           \
  do {                                                                                                      \
    cudaError_t _e = (call);                                                                                \
    if (_e != cudaSuccess)                                                                                  \
      return fail(h, CATSEG_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    ++nl;                                                                                                   \
  } while (0)
#define TAP(dst, src, nfloats)                                                                              \
  do {                                                                                                      \
    if (taps && (dst))                                                                                      \
      CUDA_OK(h, cudaMemcpyAsync((dst), (src), (size_t)(nfloats) * sizeof(float), cudaMemcpyDeviceToDevice, st)); \
  } while (0)

// Everything Aggregator.forward derives from the class embeddings alone (model.py:650, 701, 712-715 and the guidance
// half of the class-attention q/k projections), computed ONCE per vocabulary for all T classes.
static int derive_vocabulary(catseg_handle* h, cudaStream_t st, int* launches) {
  const catseg_config& c = h->cfg;
  auto& v = h->vocab;
  const int T = v.T, P = c.prompt_channel, Ct = c.text_guidance_dim, L = c.num_layers;
  const bool truncated = c.pad_len > 0 && T > c.pad_len;
  int nl = 0;
  RUN(launch_inv_norm_rows(v.text, v.inv_norm, (long long)T * P, Ct, st));
  RUN(launch_normalize_rows(v.text, v.textn, (long long)T * P, Ct, st));
  RUN(launch_iota_classes(v.iota, 1, T, st));
  RUN(launch_text_mean(truncated ? v.textn : v.text, v.iota, v.tmean, 1, T, T, P, Ct, st));
  RUN(launch_linear(v.tmean, h->tproj_wt, h->tproj_b, v.text_g, T, 128, Ct, 1, st));
  for (int l = 0; l < L; ++l)
    RUN(launch_linear(v.text_g, h->cls[l].wg_qk_t, h->cls[l].bqk, v.cg_qk + (size_t)l * T * 256, T, 256, 128, 0, st));
  v.valid = true;
  if (launches) *launches += nl;
  return CATSEG_OK;
}

extern "C" int catseg_set_vocabulary(catseg_handle* h, const float* text_feats, int T, catseg_stream stream) {
  if (!h || T < 0) return CATSEG_ERR_INVALID;
  auto& v = h->vocab;
  if (T == 0 || !text_feats) { v.T = 0; v.valid = false; return CATSEG_OK; }           // forget the vocabulary
  if (T > 6144) return fail(h, CATSEG_ERR_UNSUPPORTED, "T > 6144 classes");
  const catseg_config& c = h->cfg;
  const size_t P = c.prompt_channel, Ct = c.text_guidance_dim, L = c.num_layers;
  auto al = [](size_t n) { return (n + 63) / 64 * 64; };
  const size_t need = 2 * al(T * P * Ct) + al(T * P) + al(T * Ct) + al((size_t)T * 128) + al(L * T * 256);
  if (v.floats < need) {
    if (v.store) cudaFree(v.store);
    v.store = nullptr; v.floats = 0;
    CUDA_OK(h, cudaMalloc(&v.store, need * sizeof(float)));
    v.floats = need;
    if (v.iota) cudaFree(v.iota);
    v.iota = nullptr;
    CUDA_OK(h, cudaMalloc(&v.iota, 6144 * sizeof(int32_t)));
  }
  float* q = v.store;
  v.text = q; q += al(T * P * Ct);
  v.textn = q; q += al(T * P * Ct);
  v.inv_norm = q; q += al(T * P);
  v.tmean = q; q += al(T * Ct);
  v.text_g = q; q += al((size_t)T * 128);
  v.cg_qk = q;
  CUDA_OK(h, cudaMemcpyAsync(v.text, text_feats, (size_t)T * P * Ct * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  v.T = T;
  v.valid = false;                 // derived on the first forward (needs finalised weights), ordered on that call's stream
  return CATSEG_OK;
}

// The forward proper.  shard_world == 1: logits is [B,T,4H,4W].  shard_world > 1 (class-sharded): logits is the compact
// local buffer [B, Te/world, 4H, 4W], every per-(image, class) stage runs on this rank's slice of the kept classes and the
// linear-attention state is summed over the group through `allreduce` between the state and apply kernels.
// Exchange step of the class-sharded mode.  allreduce: activations stay class-sharded, the linear-attention state is summed.
// a2a (north_star's prescription): activations are transposed class-sharded <-> pixel-sharded around each
// class layer by kernels that store straight into the peers' buffers over NVLink (xb / pb: every rank's buffers, indexed by
// rank, opened through CUDA IPC by the host); `barrier` orders those peer stores against their consumers.
struct ShardExchange {
  catseg_allreduce_fn allreduce; void* ar_ctx;
  float* const* xb; float* const* pb; size_t buf_bytes; catseg_barrier_fn barrier; void* bar_ctx;
  float* const* lb;      // optional: every rank's FULL logits buffer [B][T][16 HW]; the head kernel stores into all of them
  float* const* gb;      // optional: every rank's guidance buffer [Swin guidance terms | decoder additive maps | flags]: the
  size_t gb_bytes;       // class-independent front end is then sharded by image and the results are pushed to the peers
};

static size_t guidance_agq_floats(const Plan& p) { return ((size_t)p.L * 2 * p.B * p.HW * 256 + 63) / 64 * 64; }
// peer buffer = [residual-stream data][per-class maxima B*T floats][barrier flag block]
static size_t exchange_data_floats(const Plan& p, int world) { return ((size_t)p.B * (p.Te / world) * p.HW * 128 + 63) / 64 * 64; }
static size_t exchange_cmax_floats(int B, int T) { return ((size_t)B * T + 63) / 64 * 64; }
static size_t exchange_extra_bytes(int B, int T) { return exchange_cmax_floats(B, T) * sizeof(float) + 256; }

static int forward_impl(catseg_handle* h, const float* img, const float* text, const float* g0, const float* g1,
                        const float* g2, float* logits, void* workspace, size_t workspace_bytes, int B, int T,
                        const catseg_taps* taps, int shard_rank, int shard_world, const ShardExchange* xc,
                        int32_t* kept_out, catseg_stream stream) {
  if (!h) return CATSEG_ERR_INVALID;
  if (!img || !g0 || !g1 || !g2 || !workspace || (!logits && !(xc && xc->lb))) return fail(h, CATSEG_ERR_INVALID, "null tensor pointer");
  if (B <= 0 || T <= 0) return fail(h, CATSEG_ERR_INVALID, "B and T must be positive (got %d, %d)", B, T);
  if (!h->finalized) return fail(h, CATSEG_ERR_WEIGHTS, "catseg_finalize_params has not been called");
  const bool use_vocab = text == nullptr;
  if (use_vocab && (h->vocab.T != T || h->vocab.store == nullptr))
    return fail(h, CATSEG_ERR_INVALID, "text_feats is NULL but no vocabulary of %d classes is set (catseg_set_vocabulary)", T);
  if (h->cfg.text_guidance_dim != h->cfg.appearance_guidance_dim)
    return fail(h, CATSEG_ERR_UNSUPPORTED, "img_feats channels (appearance_guidance_dim) must equal text_guidance_dim");
  const Plan p = make_plan(h, B, T);
  if (workspace_bytes < p.total * sizeof(float))
    return fail(h, CATSEG_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", p.total * sizeof(float), workspace_bytes);
  if ((long long)B * p.Te > 65535) return fail(h, CATSEG_ERR_UNSUPPORTED, "B*Te > 65535 slices per call");
  if (T > 6144) return fail(h, CATSEG_ERR_UNSUPPORTED, "T > 6144 classes");
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = reinterpret_cast<float*>(workspace);
  const catseg_config& c = h->cfg;
  int nl = 0;
  Seg seg{h, st, h->profiling};
  const bool sharded = shard_world > 1;
  if (sharded) {
    if (shard_rank < 0 || shard_rank >= shard_world || !xc || (!xc->allreduce && !xc->xb))
      return fail(h, CATSEG_ERR_INVALID, "bad shard rank/world/exchange");
    if (p.Te % shard_world) return fail(h, CATSEG_ERR_UNSUPPORTED, "kept classes (%d) must be a multiple of the shard group size (%d)", p.Te, shard_world);
    if (taps) return fail(h, CATSEG_ERR_UNSUPPORTED, "taps are not available in class-sharded mode");
  }
  const int Te = p.Te / shard_world;                     // classes processed by this rank
  const int nslice = B * Te;
  const bool a2a = sharded && xc->allreduce == nullptr;
  float* X = ws + p.X;
  float* PB = nullptr;                                   // a2a: this rank's pixel-sharded buffer [B][p.Te][HW / world][128]
  const int npl = p.HW / shard_world;                    // a2a: pixels per image owned by this rank in the class layers
  if (a2a) {
    if (p.pooled) return fail(h, CATSEG_ERR_UNSUPPORTED, "the all-to-all exchange needs pooling_size [1,1]");
    if (p.HW % shard_world) return fail(h, CATSEG_ERR_UNSUPPORTED, "H*W (%d) must be a multiple of the shard group size (%d)", p.HW, shard_world);
    if (shard_world > kMaxShard) return fail(h, CATSEG_ERR_UNSUPPORTED, "at most %d ranks in the all-to-all exchange", kMaxShard);
    if (!(h->split && (h->fast_mask & CATSEG_FAST_CLASS))) return fail(h, CATSEG_ERR_UNSUPPORTED, "the all-to-all exchange is implemented for the PRECISE class layer");
    const size_t need_x = exchange_data_floats(p, shard_world) * sizeof(float) + exchange_extra_bytes(B, T);
    if (!xc->xb || !xc->pb || xc->buf_bytes < need_x)
      return fail(h, CATSEG_ERR_WORKSPACE, "exchange buffers too small: need %zu bytes each", need_x);
    X = xc->xb[shard_rank];                              // the residual stream lives in the peer-visible buffer
    PB = xc->pb[shard_rank];
  }
  // class-independent front end sharded by image (guidance projections, Swin guidance terms, decoder additive maps): rank r
  // computes images [r B/world, (r+1) B/world) into its own peer-visible buffer and pushes the slices to every peer
  const bool gshard = a2a && xc->gb != nullptr && B % shard_world == 0 && h->split && (h->fast_mask & CATSEG_FAST_DECODER) &&
                      (xc->barrier == nullptr);     // the internal stream needs the library's own flag barrier
  const int gBl = gshard ? B / shard_world : B, gb0 = gshard ? shard_rank * gBl : 0;
  PeerPtrs gbp{};
  PeerFlags flagp2{};
  float* agq_base = ws + p.ag_qk;
  void* emap_area_ptr = ws + p.dec;
  if (gshard) {
    const size_t need_g = guidance_agq_floats(p) * sizeof(float) + decoder_split_emap_bytes(p.dd, B) + 256;
    if (xc->gb_bytes < need_g) return fail(h, CATSEG_ERR_WORKSPACE, "guidance exchange buffer too small: need %zu bytes", need_g);
    for (int r = 0; r < shard_world; ++r) {
      gbp.p[r] = xc->gb[r];
      flagp2.p[r] = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(xc->gb[r]) + guidance_agq_floats(p) * sizeof(float) +
                                                decoder_split_emap_bytes(p.dd, B));
    }
    agq_base = xc->gb[shard_rank];
    emap_area_ptr = xc->gb[shard_rank] + guidance_agq_floats(p);
  }
  PeerPtrs cmaxp{};                                      // a2a: every rank's [B][T] table of per-class maxima
  PeerFlags flagp{};
  if (a2a)
    for (int r = 0; r < shard_world; ++r) {
      cmaxp.p[r] = xc->pb[r] + exchange_data_floats(p, shard_world);
      flagp.p[r] = reinterpret_cast<uint32_t*>(cmaxp.p[r] + exchange_cmax_floats(B, T));
    }
  // orders the peer stores of one rank against their consumers on the others: the host's collective (callback) or, without
  // one, a flag barrier through peer memory
  auto xbarrier = [&]() -> int {
    if (xc->barrier) return xc->barrier(xc->bar_ctx, stream);
    if (launch_peer_barrier(flagp, shard_rank, shard_world, st) != cudaSuccess) return 1;
    ++nl;
    return 0;
  };
  int32_t* classes_all = reinterpret_cast<int32_t*>(ws + p.classes);          // [B][p.Te] kept class ids (ascending)
  int32_t* classes = sharded ? reinterpret_cast<int32_t*>(ws + p.classes_loc) : classes_all;   // [B][Te]

  // ---------------- PREP: cost volume, class selection, guidance projections (model.py:693-715)
  seg.begin(CATSEG_STAGE_PREP);
  CUDA_OK(h, cudaEventRecord(h->ev_fork, st));
  const bool prep_fast = (h->fast_mask & CATSEG_FAST_PREP) != 0;
  if (use_vocab && !h->vocab.valid) {                     // first call after catseg_set_vocabulary / a weight update
    int rc = derive_vocabulary(h, st, &nl);
    if (rc != CATSEG_OK) return rc;
  }
  // Cost volume (model.py:648-652).  Tensor-core path (FAST / PRECISE front end): ONE tcgen05 GEMM on the raw embeddings
  // with hi+lo fp16 operand pairs (fp32-accurate, gemm_split.cu); both L2 normalisations are row / column scales of its
  // epilogue, which also reduces the per-class maximum the top-k selection needs (:695).  When classes are truncated and
  // nobody taps the volume, the first pass writes NO volume at all (maxima only) and a second pass computes just the kept
  // classes' slices (the reference recomputes them too, :702): the raw [B,P,T,H,W] volume never reaches HBM.
  const long long text_bs = use_vocab ? 0 : (long long)T * p.P * p.Ct;    // text batch stride (0: one vocabulary for all images)
  const float* text_src = use_vocab ? h->vocab.text : text;
  const float* textn = use_vocab ? h->vocab.textn : ws + p.textn;
  bool corr_compact = false;                                               // corr holds [B][Te] kept slices instead of [B][T]
  if (prep_fast) {
    const int ntn = (p.HW + 127) / 128;
    const float* inv_t = use_vocab ? h->vocab.inv_norm : ws + p.inv_t;
    if (!use_vocab) RUN(launch_inv_norm_rows(text, ws + p.inv_t, (long long)B * T * p.P, p.Ct, st));
    RUN(launch_inv_norm_pixels(img, ws + p.inv_i, B, p.C, p.HW, st));
    GemmSplitParams g{};
    g.A = text_src; g.a_row = p.Ct; g.a_k = 1; g.a_batch = text_bs;
    g.B = img; g.b_row = 1; g.b_k = p.HW; g.b_batch = (long long)p.C * p.HW;
    g.M = T * p.P; g.N = p.HW; g.K = p.C; g.batch = B;
    g.row_scale = inv_t; g.rs_batch = use_vocab ? 0 : (long long)T * p.P;
    g.col_scale = ws + p.inv_i; g.cs_batch = p.HW;
    const bool two_pass = p.truncated && !(taps && taps->corr);
    if (!two_pass) { g.C = ws + p.corr; g.c_row = p.HW; g.c_batch = (long long)T * p.P * p.HW; }
    if (p.truncated) g.row_max = ws + p.rmaxp;
    const float* cmax_all = ws + p.cmax;
    if (a2a && two_pass) {
      // first pass sharded over the RAW classes (SURVEY.md 8e "partitioning"): this rank reduces the maxima of classes
      // [t0, t0 + Tr), stores them into every rank's [B][T] table (peer stores) and all ranks select from the full table
      const int base = T / shard_world, extra = T % shard_world;
      const int t0 = shard_rank * base + (shard_rank < extra ? shard_rank : extra), Tr = base + (shard_rank < extra ? 1 : 0);
      GemmSplitParams g1 = g;
      g1.A = g.A + (long long)t0 * p.P * p.Ct; g1.M = Tr * p.P;
      g1.row_scale = g.row_scale + (long long)t0 * p.P;
      if (Tr > 0) {
        RUN(launch_gemm_split(g1, st));
        RUN(launch_class_max(ws + p.rmaxp, ws + p.cmax, (long long)B * Tr, p.P * 2 * ntn, st));
        RUN(launch_shard_put_cmax(ws + p.cmax, cmaxp, B, Tr, T, t0, shard_world, st));
      }
      if (xbarrier() != 0) return fail(h, CATSEG_ERR_CUDA, "class-shard barrier failed");
      cmax_all = cmaxp.p[shard_rank];
    } else {
      RUN(launch_gemm_split(g, st));
      if (p.truncated) RUN(launch_class_max(ws + p.rmaxp, ws + p.cmax, (long long)B * T, p.P * 2 * ntn, st));
    }
    if (p.truncated) {
      RUN(launch_select_classes(cmax_all, classes_all, B, T, p.Te, st));
      if (two_pass) {                                      // kept slices only: rows gathered through the kept-class list
        if (p.P != 1) return fail(h, CATSEG_ERR_UNSUPPORTED, "two-pass cost volume needs prompt_channel == 1");
        // (class-sharded: only this rank's slice of the kept list; the rows land at their position in the [B][p.Te] block)
        g.a_index = classes_all + (sharded ? shard_rank * Te : 0); g.ai_batch = p.Te; g.M = sharded ? Te : p.Te; g.row_max = nullptr;
        g.C = ws + p.corr + (sharded ? (size_t)shard_rank * Te * p.HW : 0); g.c_row = p.HW; g.c_batch = (long long)p.Te * p.HW;
        RUN(launch_gemm_split(g, st));
        corr_compact = true;
      }
    } else {
      RUN(launch_iota_classes(classes_all, B, p.Te, st));
    }
    if (p.truncated && !use_vocab) RUN(launch_normalize_rows(text, ws + p.textn, (long long)B * T * p.P, p.Ct, st));   // text guidance input (:701)
  } else {
    RUN(launch_normalize_img(img, ws + p.imgn, B, p.C, p.HW, st));
    if (!use_vocab) RUN(launch_normalize_rows(text, ws + p.textn, (long long)B * T * p.P, p.Ct, st));
    RUN(launch_cost_volume(textn, text_bs, ws + p.imgn, ws + p.corr, B, T * p.P, p.C, p.HW, st));
    if (p.truncated) {
      RUN(launch_class_max(ws + p.corr, ws + p.cmax, (long long)B * T, p.P * p.HW, st));
      RUN(launch_select_classes(ws + p.cmax, classes_all, B, T, p.Te, st));
    } else {
      RUN(launch_iota_classes(classes_all, B, p.Te, st));
    }
  }
  if (sharded) RUN(launch_slice_classes(classes_all, classes, B, p.Te, shard_rank * Te, Te, st));
  if (kept_out) CUDA_OK(h, cudaMemcpyAsync(kept_out, classes_all, (size_t)B * p.Te * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  const bool class_fast = (h->fast_mask & CATSEG_FAST_CLASS) != 0;
  __half* timg = reinterpret_cast<__half*>(ws + p.timg);
  CUDA_OK(h, cudaEventRecord(h->ev_sel, st));            // the kept-class list (and the normalised text) are ready
  // The guidance projections only depend on the inputs: they run on an internal stream beside the cost volume / class
  // selection / text chain (small, latency-bound kernels) and are joined before the embedding.  Externally the call is
  // still ordered on `stream`.
  // the single-term guidance projections are class independent (they cancel in the argmax) but would cap the logits
  // parity of the PRECISE mode at ~1e-2: that mode keeps them on the fp32 kernels
  const bool gconv_fast = prep_fast && !h->split;
  {
    cudaStream_t mainst = st;
    CUDA_OK(h, cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));     // ev_fork was recorded at the start of the stage
    cudaStream_t st = h->aux_stream;             // shadows: the launches below go to the internal stream
    const bool gsplit = prep_fast && h->split;
    const float* g0l = g0 + (size_t)gb0 * p.Cg * p.HW;      // guidance sharded by image: this rank's images only
    if (gsplit && h->gconv_split_img[0]) RUN(launch_gconv_split(0, g0l, h->gconv_split_img[0], h->gproj_b, ws + p.app_g, gBl, p.Cg, st));
    else if (gconv_fast && h->gconv_img[0]) RUN(launch_gconv_fast(0, g0l, h->gconv_img[0], h->gproj_b, ws + p.app_g, gBl, p.Cg, st));
    else RUN(launch_conv3x3_nchw(g0l, h->gproj_wt, h->gproj_b, ws + p.app_g, gBl, p.Cg, p.H, p.W, 128, st));
    for (int l = 0; l < p.L; ++l) {
      RUN(launch_layernorm128(ws + p.app_g, ws + p.app_gn, h->gnorm_g[l], h->gnorm_b[l], (long long)gBl * p.HW, st));
      for (int k = 0; k < 2; ++k)
        RUN(launch_linear(ws + p.app_gn, h->swin[l * 2 + k].wg_qk_t, h->swin[l * 2 + k].bqk,
                          agq_base + (size_t)(l * 2 + k) * B * p.HW * 256 + (size_t)gb0 * p.HW * 256, (long long)gBl * p.HW, 256, 128, 0, st));
    }
    if (gshard) {                                          // push this rank's slices of the Swin guidance terms to the peers
      PeerSegs sg{};
      for (int blk = 0; blk < p.L * 2 && blk < 8; ++blk) {
        sg.off[blk] = (long long)((size_t)blk * B * p.HW * 256 + (size_t)gb0 * p.HW * 256);
        sg.n[blk] = (long long)gBl * p.HW * 256;
      }
      sg.nseg = p.L * 2;
      RUN(launch_peer_bcast(gbp, sg, shard_rank, shard_world, st));
      RUN(launch_peer_barrier(flagp2, shard_rank, shard_world, st));
    }
    // the Swin blocks only need the projected appearance guidance: join here; the decoder guidance (and, PRECISE, the
    // decoder's per-image additive maps) stay on the internal stream and are joined right before the decoder
    CUDA_OK(h, cudaEventRecord(h->ev_join, st));
    CUDA_OK(h, cudaStreamWaitEvent(mainst, h->ev_join, 0));
    // the text guidance of the kept classes is first needed by the class layers: it is derived here, off the critical chain
    // (cost volume -> selection -> embedding), once the selection has been made on the caller's stream
    CUDA_OK(h, cudaStreamWaitEvent(st, h->ev_sel, 0));
    if (a2a) {
      // the class layers of this mode run on ALL kept classes at this rank's pixels: their text guidance is derived below
      for (int l = 0; l < p.L; ++l)
        RUN(launch_class_pad_state(h->cls[l], 128, ws + p.pad_state + (size_t)l * kStateFloats, p.n_pad, p.S, st));
    } else if (use_vocab) {
      // the text guidance of a vocabulary is derived once (derive_vocabulary): per call the kept classes only pick their rows
      RUN(launch_gather_rows(h->vocab.text_g, classes, ws + p.text_g, (long long)B * Te, 128, st));
      for (int l = 0; l < p.L; ++l) {
        RUN(launch_gather_rows(h->vocab.cg_qk + (size_t)l * T * 256, classes, ws + p.cg_qk + (size_t)l * B * Te * 256, (long long)B * Te, 256, st));
        RUN(launch_class_pad_state(h->cls[l], 128, ws + p.pad_state + (size_t)l * kStateFloats, p.n_pad, p.S, st));
      }
    } else {
      RUN(launch_text_mean(p.truncated ? ws + p.textn : text, classes, ws + p.tmean, B, T, Te, p.P, p.Ct, st));
      RUN(launch_linear(ws + p.tmean, h->tproj_wt, h->tproj_b, ws + p.text_g, (long long)B * Te, 128, p.Ct, 1, st));
      for (int l = 0; l < p.L; ++l) {
        RUN(launch_linear(ws + p.text_g, h->cls[l].wg_qk_t, h->cls[l].bqk, ws + p.cg_qk + (size_t)l * B * Te * 256,
                          (long long)B * Te, 256, 128, 0, st));
        RUN(launch_class_pad_state(h->cls[l], 128, ws + p.pad_state + (size_t)l * kStateFloats, p.n_pad, p.S, st));
      }
    }
    if (a2a) {
      // the class layers of this mode see ALL kept classes (at this rank's pixels): text guidance of the whole kept list
      if (use_vocab) {
        RUN(launch_gather_rows(h->vocab.text_g, classes_all, ws + p.text_g, (long long)B * p.Te, 128, st));
      } else {
        RUN(launch_text_mean(p.truncated ? ws + p.textn : text, classes_all, ws + p.tmean, B, T, p.Te, p.P, p.Ct, st));
        RUN(launch_linear(ws + p.tmean, h->tproj_wt, h->tproj_b, ws + p.text_g, (long long)B * p.Te, 128, p.Ct, 1, st));
      }
      RUN(launch_pack_text_img(ws + p.text_g, timg, B, p.Te, st));
    } else if (class_fast) RUN(launch_pack_text_img(ws + p.text_g, timg, B, Te, st));
    // peer-direct logits: this rank's full buffer is pre-filled with -100 here (dropped classes, model.py:721); the peers'
    // head kernels store the kept planes much later, after barriers that the second join of this stream precedes
    if (a2a && xc->lb && p.truncated) RUN(launch_fill(xc->lb[shard_rank], -100.0f, (long long)B * T * 16 * p.HW, st));
    const float* g1l = g1 + (size_t)gb0 * c.decoder_guidance_dims[0] * 4 * p.HW;
    const float* g2l = g2 + (size_t)gb0 * c.decoder_guidance_dims[1] * 16 * p.HW;
    if (gsplit && h->gconv_split_img[1])
      RUN(launch_gconv_split(1, g1l, h->gconv_split_img[1], h->dgp_b[0], ws + p.dg0, gBl, c.decoder_guidance_dims[0], st));
    else if (gconv_fast && h->gconv_img[1])
      RUN(launch_gconv_fast(1, g1l, h->gconv_img[1], h->dgp_b[0], ws + p.dg0, gBl, c.decoder_guidance_dims[0], st));
    else
      RUN(launch_conv3x3_nchw(g1l, h->dgp_wt[0], h->dgp_b[0], ws + p.dg0, gBl, c.decoder_guidance_dims[0], 2 * p.H, 2 * p.W,
                              p.dd.G1, st));
    if (gsplit && h->gconv_split_img[2])
      RUN(launch_gconv_split(2, g2l, h->gconv_split_img[2], h->dgp_b[1], ws + p.dg1, gBl, c.decoder_guidance_dims[1], st));
    else if (gconv_fast && h->gconv_img[2])
      RUN(launch_gconv_fast(2, g2l, h->gconv_img[2], h->dgp_b[1], ws + p.dg1, gBl, c.decoder_guidance_dims[1], st));
    else
      RUN(launch_conv3x3_nchw(g2l, h->dgp_wt[1], h->dgp_b[1], ws + p.dg1, gBl, c.decoder_guidance_dims[1], 4 * p.H, 4 * p.W,
                              p.dd.G2, st));
    if ((h->fast_mask & CATSEG_FAST_DECODER) && h->split) {
      cudaError_t e = decoder_split_prepare(ws + p.dg0, ws + p.dg1, B, gb0, gBl, p.dd, h->dec_fast, emap_area_ptr, &nl, st);
      if (e != cudaSuccess) return fail(h, CATSEG_ERR_CUDA, "decoder maps: %s", cudaGetErrorString(e));
      if (gshard) {                                        // push this rank's slices of the tile-ordered additive maps
        size_t o1, o2, per1, per2;
        decoder_split_emap_slices(p.dd, B, &o1, &o2, &per1, &per2);
        PeerSegs sg{};
        sg.off[0] = (long long)(guidance_agq_floats(p) + o1 + (size_t)gb0 * per1); sg.n[0] = (long long)gBl * per1;
        sg.off[1] = (long long)(guidance_agq_floats(p) + o2 + (size_t)gb0 * per2); sg.n[1] = (long long)gBl * per2;
        sg.nseg = 2;
        RUN(launch_peer_bcast(gbp, sg, shard_rank, shard_world, st));
        RUN(launch_peer_barrier(flagp2, shard_rank, shard_world, st));
      }
    }
    CUDA_OK(h, cudaEventRecord(h->ev_join2, st));
    if (taps) CUDA_OK(h, cudaStreamWaitEvent(mainst, h->ev_join2, 0));     // the taps below copy the decoder guidance
  }
  seg.end();
  if (taps) {
    TAP(taps->corr, ws + p.corr, (size_t)B * T * p.P * p.HW);
    TAP(taps->classes, ws + p.classes, (size_t)B * Te);
    TAP(taps->app_guidance, ws + p.app_g, (size_t)B * p.HW * 128);
    TAP(taps->text_guidance, ws + p.text_g, (size_t)B * Te * 128);
    TAP(taps->dec_guidance0, ws + p.dg0, (size_t)B * 4 * p.HW * p.dd.G1);
    TAP(taps->dec_guidance1, ws + p.dg1, (size_t)B * 16 * p.HW * p.dd.G2);
  }

  // ---------------- EMBED (model.py:704)
  seg.begin(CATSEG_STAGE_EMBED);
  {
    // compact volume ([B][p.Te] kept slices): slice s of image b is row (rank offset + j) of that image's block
    const int32_t* eids = classes;
    int eT = T;
    if (corr_compact) {
      int32_t* ids = reinterpret_cast<int32_t*>(ws + p.classes_loc) + (size_t)B * p.Te;     // scratch behind the local class list
      RUN(launch_iota_range(ids, B, Te, sharded ? shard_rank * Te : 0, st));
      eids = ids; eT = p.Te;
    }
    if (prep_fast && h->embed_img)
      RUN(launch_cost_embed_fast(ws + p.corr, eids, h->embed_img, h->conv1_b, X, B, eT, Te, h->num_sms, st));
    else
      RUN(launch_cost_embed(ws + p.corr, eids, h->conv1_wt, h->conv1_b, X, B, eT, Te, p.P, p.H, p.W, st));
  }
  seg.end();
  TAP(taps->embed, X, (size_t)nslice * p.HW * 128);

  // ---------------- aggregation layers (model.py:717-718)
  for (int l = 0; l < p.L; ++l) {
    for (int k = 0; k < 2; ++k) {
      const bool attn_fast = (h->fast_mask & CATSEG_FAST_SWIN_ATTN) != 0;
      const bool mlp_fast = attn_fast || (h->fast_mask & CATSEG_FAST_SWIN_MLP) != 0;
      const float* agk = agq_base + (size_t)(l * 2 + k) * B * p.HW * 256;
      const int shift = k == 0 ? 0 : c.window_size / 2;
      seg.begin(CATSEG_STAGE_SWIN);
      if (attn_fast && h->attn_version == 2) {
        float* agT = ws + p.agw + (size_t)(l * 2 + k) * B * 4 * 256 * 144;
        RUN(launch_pack_ag_windows_T(agk, agT, B, shift, st));
        RUN(launch_swin_attn2(X, agT, nslice, Te, shift, h->swin_attn2[l * 2 + k], h->split, h->num_sms, st));
      } else if (attn_fast) {
        __half* agw = reinterpret_cast<__half*>(ws + p.agw) + (size_t)(l * 2 + k) * B * 16 * 144 * 64;
        RUN(launch_pack_ag_windows(agk, agw, B, shift, st));
        RUN(launch_swin_attn_fast(X, agw, nslice, Te, shift, h->swin_attn_fast[l * 2 + k], h->num_sms, st));
      }
      else RUN(launch_swin_block_exact(X, agk, nslice, Te, shift, h->swin[l * 2 + k], mlp_fast ? 0 : 1, st));
      seg.end();
      static const bool dbg_skip_mlp = getenv("CATSEG_DBG_SKIP_MLP") != nullptr;
      if (mlp_fast && !dbg_skip_mlp) {
        seg.begin(CATSEG_STAGE_SWIN_MLP);
        if (h->split) RUN(launch_mlp_split(X, nullptr, X, (long long)nslice * p.HW, h->swin_mlp_split[l * 2 + k], 0, h->num_sms, st));
        else RUN(launch_mlp_fast(X, (long long)nslice * p.HW, h->swin_mlp_fast[l * 2 + k], 0, h->num_sms, st));
        seg.end();
      }
      if (k == 0) TAP(taps->swin_b1[l], X, (size_t)nslice * p.HW * 128);
      else TAP(taps->swin_b2[l], X, (size_t)nslice * p.HW * 128);
    }
    if (l == 0) CUDA_OK(h, cudaStreamWaitEvent(st, h->ev_join2, 0));   // text guidance (internal stream)
    seg.begin(CATSEG_STAGE_CLASS);
    const float* cg = ws + p.cg_qk + (size_t)l * B * Te * 256;
    const float* pad = ws + p.pad_state + (size_t)l * kStateFloats;
    if (a2a) {
      // class-sharded X -> pixel-sharded PB of every rank (peer stores over NVLink), the class layer on all p.Te classes at
      // this rank's pixels (no reduction needed: every class of a pixel is local), and back
      PeerPtrs xbp{}, pbp{};
      for (int r = 0; r < shard_world; ++r) { xbp.p[r] = xc->xb[r]; pbp.p[r] = xc->pb[r]; }
      seg.end();
      seg.begin(CATSEG_STAGE_EXCHANGE);
      RUN(launch_shard_c2p(X, pbp, B, Te, p.Te, p.HW, shard_rank, shard_world, st));
      if (xbarrier() != 0) return fail(h, CATSEG_ERR_CUDA, "class-shard barrier failed");
      seg.end();
      seg.begin(CATSEG_STAGE_CLASS);
      float* x1 = ws + p.X1;
      RUN(launch_class_state_split(PB, timg, ws + p.state, B, p.Te, npl, p.S, h->class_split[l], h->num_sms, st));
      RUN(launch_class_apply_split(PB, x1, timg, ws + p.state, pad, B, p.Te, npl, p.S, h->class_split[l], h->num_sms, st));
      RUN(launch_mlp_split(x1, PB, PB, (long long)B * p.Te * npl, h->class_mlp_split[l], 1, h->num_sms, st));
      seg.end();
      seg.begin(CATSEG_STAGE_EXCHANGE);
      RUN(launch_shard_p2c(PB, xbp, B, Te, p.Te, p.HW, shard_rank, shard_world, st));
      // peer-direct logits: arriving at the last barrier must imply that this rank's -100 pre-fill (internal stream) is
      // complete, because the peers' head kernels store into this rank's buffer any time after it
      if (xc->lb && l == p.L - 1) CUDA_OK(h, cudaStreamWaitEvent(st, h->ev_join2, 0));
      if (xbarrier() != 0) return fail(h, CATSEG_ERR_CUDA, "class-shard barrier failed");
      seg.end();
      continue;
    }
    if (p.pooled) RUN(launch_avgpool_tokens(X, ws + p.Xp, nslice, p.H, p.W, c.pooling_size[0], c.pooling_size[1], st));
    const float* xin = p.pooled ? ws + p.Xp : X;
    float* xout = p.pooled ? ws + p.Xp2 : X;
    const int omode = p.pooled ? 1 : 0;
    if (class_fast && h->split) RUN(launch_class_state_split(xin, timg, ws + p.state, B, Te, p.npix, p.S, h->class_split[l], h->num_sms, st));
    else if (class_fast) RUN(launch_class_state_fast(xin, timg, ws + p.state, B, Te, p.npix, p.S, h->class_fast[l], h->num_sms, st));
    else RUN(launch_class_state_exact(xin, cg, ws + p.state, B, Te, p.npix, p.S, h->cls[l], st));
    if (sharded) {                                         // the only exchange step of the path: sum of the per-pixel state
      seg.end();
      seg.begin(CATSEG_STAGE_EXCHANGE);
      int rc = xc->allreduce(xc->ar_ctx, ws + p.state, (size_t)B * p.npix * kStateFloats, stream);
      if (rc != 0) return fail(h, CATSEG_ERR_CUDA, "class-shard all-reduce callback failed (%d)", rc);
      seg.end();
      seg.begin(CATSEG_STAGE_CLASS);
    }
    if (class_fast && h->split) {
      // x1 = x + attention -> X1 (pooled: Xp2), then the token MLP kernel adds MLP(LN2(x1)) and the outer residual (model.py:423)
      float* x1 = p.pooled ? ws + p.Xp2 : ws + p.X1;
      RUN(launch_class_apply_split(xin, x1, timg, ws + p.state, pad, B, Te, p.npix, p.S, h->class_split[l], h->num_sms, st));
      RUN(launch_mlp_split(x1, p.pooled ? nullptr : X, xout, (long long)nslice * p.npix, h->class_mlp_split[l], 1, h->num_sms, st));
    } else if (class_fast)
      RUN(launch_class_apply_fast(xin, xout, timg, ws + p.state, pad, B, Te, p.npix, p.S, omode, h->class_fast[l], h->num_sms, st));
    else
      RUN(launch_class_apply_exact(xin, xout, cg, ws + p.state, pad, B, Te, p.npix, p.S, omode, h->cls[l], st));
    if (p.pooled) RUN(launch_upsample_add(X, ws + p.Xp2, nslice, p.H, p.W, p.Hp, p.Wp, st));
    seg.end();
    TAP(taps->class_out[l], X, (size_t)nslice * p.HW * 128);
  }

  // ---------------- decoder + scatter (model.py:720-724)
  CUDA_OK(h, cudaStreamWaitEvent(st, h->ev_join2, 0));       // decoder guidance / additive maps from the internal stream
  seg.begin(CATSEG_STAGE_DECODER);
  // class-sharded: the output is the compact local buffer [B][Te][16 HW]: plane ids are 0..Te-1 and T_out = Te
  const int32_t* out_ids = classes;
  int T_out = T;
  PeerPtrs lpeers{};
  const int nlp = (a2a && xc->lb) ? shard_world : 0;
  for (int r = 0; r < nlp; ++r) lpeers.p[r] = xc->lb[r];
  if (nlp > 0) {
    // every rank's full [B][T][16 HW] buffer receives this rank's planes at their class ids (no all-gather, no assembly)
  } else if (sharded) {
    int32_t* ids = reinterpret_cast<int32_t*>(ws + p.classes_loc) + (size_t)B * p.Te;
    RUN(launch_iota_classes(ids, B, Te, st));
    out_ids = ids;
    T_out = Te;
  } else if (p.truncated) {
    RUN(launch_fill(logits, -100.0f, (long long)B * T * 16 * p.HW, st));
  }
  if (h->fast_mask & CATSEG_FAST_DECODER) {
    if (taps && (taps->up1 || taps->up2))
      return fail(h, CATSEG_ERR_UNSUPPORTED, "up1/up2 taps are only available with the exact decoder");
    cudaError_t e = h->split ? run_decoder_split(X, ws + p.dg0, ws + p.dg1, out_ids, logits, B, T_out, Te, p.dd, h->dec_fast, h->dec,
                                                 h->head_bias_host, ws + p.dec, p.dec_chunk, h->num_sms, &nl, nlp > 0 ? &lpeers : nullptr,
                                                 nlp, gshard ? emap_area_ptr : nullptr, st)
                             : run_decoder_fast(X, ws + p.dg0, ws + p.dg1, out_ids, logits, B, T_out, Te, p.dd, h->dec_fast, h->dec,
                                                h->head_bias_host, ws + p.dec, p.dec_chunk, h->num_sms, &nl, st);
    if (e != cudaSuccess) return fail(h, CATSEG_ERR_CUDA, "fast decoder: %s", cudaGetErrorString(e));
    if (nlp > 0 || gshard) {
      // peer-direct logits: the peers' planes have landed in this rank's buffer after this barrier.  Guidance sharded by
      // image: no peer may start the next call's pushes into this rank's guidance buffer before this decoder has read it
      seg.end();
      seg.begin(CATSEG_STAGE_EXCHANGE);
      if (xbarrier() != 0) return fail(h, CATSEG_ERR_CUDA, "class-shard barrier failed");
    }
  } else {
    cudaError_t e = run_decoder_exact(X, ws + p.dg0, ws + p.dg1, out_ids, logits, B, T_out, Te, p.dd, h->dec,
                                      ws + p.dec, p.dec_chunk, taps ? taps->up1 : nullptr,
                                      taps ? taps->up2 : nullptr, &nl, st);
    if (e != cudaSuccess) return fail(h, CATSEG_ERR_CUDA, "decoder: %s", cudaGetErrorString(e));
  }
  seg.end();
  if (h->profiling) ++h->prof_forwards;
  h->last_launches = nl;
  return CATSEG_OK;
}

extern "C" int catseg_forward_taps(catseg_handle* h, const float* img, const float* text, const float* g0,
                                   const float* g1, const float* g2, float* logits, void* workspace,
                                   size_t workspace_bytes, int B, int T, const catseg_taps* taps,
                                   catseg_stream stream) {
  return forward_impl(h, img, text, g0, g1, g2, logits, workspace, workspace_bytes, B, T, taps, 0, 1, nullptr, nullptr, stream);
}

extern "C" int catseg_forward_class_sharded(catseg_handle* h, const float* img, const float* text, const float* g0,
                                            const float* g1, const float* g2, float* logits_local,
                                            int32_t* kept_classes_out, void* workspace, size_t workspace_bytes, int B, int T,
                                            int shard_rank, int shard_world, catseg_allreduce_fn allreduce, void* ctx,
                                            catseg_stream stream) {
  if (shard_world < 1) return CATSEG_ERR_INVALID;
  ShardExchange xc{allreduce, ctx, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0};
  return forward_impl(h, img, text, g0, g1, g2, logits_local, workspace, workspace_bytes, B, T, nullptr, shard_rank, shard_world,
                      &xc, kept_classes_out, stream);
}

extern "C" size_t catseg_exchange_guidance_bytes(const catseg_handle* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  const Plan p = make_plan(h, B, T);
  return guidance_agq_floats(p) * sizeof(float) + decoder_split_emap_bytes(p.dd, B) + 256;
}

// 1 if a flag barrier of this rank ever gave up waiting (a peer died or fell > ~4 s behind): results since then are invalid.
// Synchronous (one 4-byte device-to-host copy per buffer); meant for the end of a job or a test, not for the hot path.
extern "C" int catseg_exchange_timed_out(const catseg_handle* h, const float* pbuf_self, const float* gbuf_self, int B, int T,
                                         int shard_world, int* timed_out) {
  if (!h || !pbuf_self || !timed_out || B <= 0 || T <= 0 || shard_world < 1) return CATSEG_ERR_INVALID;
  const Plan p = make_plan(h, B, T);
  if (p.Te % shard_world) return CATSEG_ERR_INVALID;
  uint32_t v = 0, v2 = 0;
  const uint32_t* f1 = reinterpret_cast<const uint32_t*>(pbuf_self + exchange_data_floats(p, shard_world) + exchange_cmax_floats(B, T));
  if (cudaMemcpy(&v, f1 + kMaxShard + 1, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return CATSEG_ERR_CUDA;
  if (gbuf_self) {
    const uint32_t* f2 = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(gbuf_self) + guidance_agq_floats(p) * sizeof(float) +
                                                           decoder_split_emap_bytes(p.dd, B));
    if (cudaMemcpy(&v2, f2 + kMaxShard + 1, sizeof(v2), cudaMemcpyDeviceToHost) != cudaSuccess) return CATSEG_ERR_CUDA;
  }
  *timed_out = (v | v2) != 0 ? 1 : 0;
  return CATSEG_OK;
}

extern "C" size_t catseg_exchange_logits_bytes(const catseg_handle* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  const Plan p = make_plan(h, B, T);
  return (size_t)B * T * 16 * p.HW * sizeof(float);
}

extern "C" size_t catseg_exchange_buffer_bytes(const catseg_handle* h, int B, int T, int shard_world) {
  if (!h || B <= 0 || T <= 0 || shard_world < 1) return 0;
  const Plan p = make_plan(h, B, T);
  if (p.Te % shard_world) return 0;
  return exchange_data_floats(p, shard_world) * sizeof(float) + exchange_extra_bytes(B, T);
}

extern "C" int catseg_forward_class_sharded_a2a(catseg_handle* h, const float* img, const float* text, const float* g0,
                                                const float* g1, const float* g2, float* logits_local,
                                                int32_t* kept_classes_out, void* workspace, size_t workspace_bytes, int B, int T,
                                                int shard_rank, int shard_world, float* const* xbuf_peers,
                                                float* const* pbuf_peers, size_t buf_bytes, float* const* logits_peers,
                                                float* const* gbuf_peers, size_t gbuf_bytes, catseg_barrier_fn barrier, void* ctx,
                                                catseg_stream stream) {
  if (shard_world < 1 || !xbuf_peers || !pbuf_peers) return CATSEG_ERR_INVALID;
  ShardExchange xc{nullptr, nullptr, xbuf_peers, pbuf_peers, buf_bytes, barrier, ctx, logits_peers, gbuf_peers, gbuf_bytes};
  return forward_impl(h, img, text, g0, g1, g2, logits_local, workspace, workspace_bytes, B, T, nullptr, shard_rank, shard_world,
                      &xc, kept_classes_out, stream);
}

extern "C" int catseg_assemble_class_sharded(const float* gathered, const int32_t* kept_classes, int32_t* pos_scratch, float* logits,
                                             int shard_world, int B, int T_local, int T, int64_t npix, catseg_stream stream) {
  if (!gathered || !kept_classes || !pos_scratch || !logits || shard_world < 1 || B <= 0 || T_local <= 0 || T <= 0 || npix <= 0)
    return CATSEG_ERR_INVALID;
  cudaError_t e = launch_assemble_class_sharded(gathered, kept_classes, pos_scratch, logits, shard_world, B, T_local, T, npix,
                                                (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return e == cudaErrorInvalidValue ? CATSEG_ERR_INVALID : CATSEG_ERR_CUDA; }
  return CATSEG_OK;
}

// Peer-visible device memory for the all-to-all exchange: plain cudaMalloc allocations (CUDA IPC cannot export a sub-range
// of a caching allocator's pool), exported / opened with the CUDA IPC handle API.
extern "C" int catseg_peer_alloc(size_t bytes, void** out) {
  if (!out || bytes == 0) return CATSEG_ERR_INVALID;
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaSuccess) e = cudaMemset(*out, 0, bytes);       // the barrier flag block starts at epoch 0
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return CATSEG_ERR_CUDA; }
  return CATSEG_OK;
}
extern "C" int catseg_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return CATSEG_ERR_CUDA; }
  return CATSEG_OK;
}
extern "C" int catseg_peer_export(const void* ptr, uint8_t handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr || !handle_out) return CATSEG_ERR_INVALID;
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, const_cast<void*>(ptr));
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return CATSEG_ERR_CUDA; }
  memcpy(handle_out, &hd, 64);
  return CATSEG_OK;
}
extern "C" int catseg_peer_open(const uint8_t handle[64], void** out) {
  if (!handle || !out) return CATSEG_ERR_INVALID;
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return CATSEG_ERR_CUDA; }
  return CATSEG_OK;
}
extern "C" int catseg_peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return CATSEG_ERR_CUDA; }
  return CATSEG_OK;
}

extern "C" int catseg_forward(catseg_handle* h, const float* img, const float* text, const float* g0,
                              const float* g1, const float* g2, float* logits, void* workspace,
                              size_t workspace_bytes, int B, int T, catseg_stream stream) {
  return catseg_forward_taps(h, img, text, g0, g1, g2, logits, workspace, workspace_bytes, B, T, nullptr, stream);
}

extern "C" int catseg_set_profiling(catseg_handle* h, int enable) {
  if (!h) return CATSEG_ERR_INVALID;
  if (enable)                                       // the timing events are created here, not inside catseg_forward
    for (cudaEvent_t& e : h->ev)
      if (!e) CUDA_OK(h, cudaEventCreate(&e));
  h->profiling = enable != 0;
  return CATSEG_OK;
}

extern "C" int catseg_stage_times(catseg_handle* h, float* ms, int* calls, int reset) {
  if (!h || !ms) return CATSEG_ERR_INVALID;
  for (int i = 0; i < CATSEG_STAGE_COUNT; ++i) ms[i] = 0.0f;
  for (int i = 0; i < h->ev_used; ++i) {
    CUDA_OK(h, cudaEventSynchronize(h->ev[(size_t)i * 2 + 1]));
    float t = 0.0f;
    CUDA_OK(h, cudaEventElapsedTime(&t, h->ev[(size_t)i * 2], h->ev[(size_t)i * 2 + 1]));
    ms[h->ev_stage[i]] += t;
  }
  if (calls) *calls = h->prof_forwards;
  if (reset) { h->ev_used = 0; h->prof_forwards = 0; }
  return CATSEG_OK;
}

extern "C" int catseg_last_launch_count(const catseg_handle* h) { return h ? h->last_launches : 0; }

static int stitch_impl(const float* win_logits, int T, int S, int kernel, int stride, int out_res, int height, int width,
                       float* probs_out, int32_t* labels_out, void* scratch, size_t scratch_bytes, catseg_stream stream) {
  if (!win_logits || T <= 0 || S <= 0 || height <= 0 || width <= 0) return CATSEG_ERR_INVALID;
  if (!probs_out && !labels_out) return CATSEG_ERR_INVALID;
  if (scratch && scratch_bytes < (size_t)T * sizeof(uint32_t)) return CATSEG_ERR_WORKSPACE;
  cudaError_t e = launch_stitch(win_logits, T, S, kernel, stride, out_res, height, width, probs_out, labels_out,
                                reinterpret_cast<uint32_t*>(scratch), (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return e == cudaErrorInvalidValue ? CATSEG_ERR_INVALID : CATSEG_ERR_CUDA; }
  return CATSEG_OK;
}

extern "C" int catseg_stitch_argmax(const float* win_logits, int T, int S, int kernel, int stride, int out_res,
                                    int height, int width, float* probs_out, int32_t* labels_out,
                                    catseg_stream stream) {
  return stitch_impl(win_logits, T, S, kernel, stride, out_res, height, width, probs_out, labels_out, nullptr, 0, stream);
}

extern "C" size_t catseg_stitch_scratch_bytes(int T) { return T > 0 ? (size_t)T * sizeof(uint32_t) : 0; }

extern "C" int catseg_stitch_argmax_ws(const float* win_logits, int T, int S, int kernel, int stride, int out_res,
                                       int height, int width, float* probs_out, int32_t* labels_out, void* scratch,
                                       size_t scratch_bytes, catseg_stream stream) {
  if (!scratch) return CATSEG_ERR_INVALID;
  return stitch_impl(win_logits, T, S, kernel, stride, out_res, height, width, probs_out, labels_out, scratch, scratch_bytes, stream);
}

extern "C" int catseg_argmax_batched(const float* scores, int batch, int T, int64_t npix, int32_t* labels_out,
                                     catseg_stream stream) {
  if (!scores || !labels_out || batch <= 0 || batch > 65535 || T <= 0 || npix <= 0) return CATSEG_ERR_INVALID;
  cudaError_t e = launch_argmax(scores, batch, T, npix, labels_out, (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return CATSEG_ERR_CUDA; }
  return CATSEG_OK;
}

extern "C" int catseg_guidance_upsample(const float* tokens, const float* weight, const float* bias, float* out, int B,
                                       int width, int cout, int kernel, int grid, catseg_stream stream) {
  if (!tokens || !weight || !bias || !out || B <= 0 || width <= 0 || cout <= 0 || kernel <= 0 || kernel > 8 || grid <= 0 ||
      (long long)B * grid * grid > 0x7fffffffLL / 2 || (long long)cout * kernel * kernel > 0x7fffffffLL / 2)
    return CATSEG_ERR_INVALID;
  cudaError_t e = launch_guidance_upsample(tokens, weight, bias, out, B, width, cout, kernel, grid, (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return CATSEG_ERR_CUDA; }
  return CATSEG_OK;
}

extern "C" int catseg_strip_cls_nchw(const float* feats, float* out, int B, int C, int grid, catseg_stream stream) {
  if (!feats || !out || B <= 0 || B > 65535 || C <= 0 || grid <= 0) return CATSEG_ERR_INVALID;
  cudaError_t e = launch_strip_cls_nchw(feats, out, B, C, grid, (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return CATSEG_ERR_CUDA; }
  return CATSEG_OK;
}

extern "C" size_t catseg_clip_dense_workspace_bytes(int L, int N, int width, int prompt) {
  if (L <= 0 || N <= 0 || width <= 0 || prompt < 0 || prompt >= L) return 0;
  return clip_dense_workspace_floats(L, N, width, prompt) * sizeof(float);
}

extern "C" int catseg_clip_dense_last_block(const catseg_clip_dense_weights* w, const float* x, int L, int N, int prompt,
                                            float* block_out, float* feats_out, void* workspace, size_t workspace_bytes,
                                            catseg_stream stream) {
  auto bad = [](const char* m) { g_create_error = m; return CATSEG_ERR_INVALID; };
  if (!w || !x || !workspace) return bad("clip_dense: null pointer");
  if (!block_out && !feats_out) return bad("clip_dense: no output requested");
  if (L <= 0 || N <= 0 || prompt < 0 || prompt >= L) return bad("clip_dense: bad L / N / prompt");
  if (w->width <= 0 || w->width % 8 || (feats_out && w->out_dim <= 0)) return bad("clip_dense: width must be a positive multiple of 8");
  if ((long long)(L - prompt) * N > 0x7fffffffLL / 4096) return bad("clip_dense: too many tokens");
  if (!w->ln_1_weight || !w->ln_1_bias || !w->v_proj_weight || !w->v_proj_bias || !w->out_proj_weight || !w->out_proj_bias ||
      !w->ln_2_weight || !w->ln_2_bias || !w->c_fc_weight || !w->c_fc_bias || !w->c_proj_weight || !w->c_proj_bias ||
      (feats_out && (!w->ln_post_weight || !w->ln_post_bias || !w->proj)))
    return bad("clip_dense: missing parameter");
  if (workspace_bytes < catseg_clip_dense_workspace_bytes(L, N, w->width, prompt)) {
    g_create_error = "clip_dense: workspace too small";
    return CATSEG_ERR_WORKSPACE;
  }
  ClipDenseW k{};
  k.width = w->width; k.out_dim = w->out_dim;
  k.ln1_g = w->ln_1_weight; k.ln1_b = w->ln_1_bias;
  k.v_w = w->v_proj_weight; k.v_b = w->v_proj_bias;
  k.out_proj_w = w->out_proj_weight; k.out_proj_b = w->out_proj_bias;
  k.ln2_g = w->ln_2_weight; k.ln2_b = w->ln_2_bias;
  k.c_fc_w = w->c_fc_weight; k.c_fc_b = w->c_fc_bias;
  k.c_proj_w = w->c_proj_weight; k.c_proj_b = w->c_proj_bias;
  k.ln_post_g = w->ln_post_weight; k.ln_post_b = w->ln_post_bias; k.proj = w->proj;
  cudaError_t e = run_clip_dense_block(k, x, L, N, prompt, block_out, feats_out, reinterpret_cast<float*>(workspace), (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return CATSEG_ERR_CUDA; }
  return CATSEG_OK;
}

extern "C" int catseg_argmax(const float* scores, int T, int64_t npix, int32_t* labels_out, catseg_stream stream) {
  return catseg_argmax_batched(scores, 1, T, npix, labels_out, stream);
}
