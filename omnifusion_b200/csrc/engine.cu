// Engine: owns the packed weights, the workspace and the forward schedule of the
// OmniFusion patch network (model/spherical_model_iterative.py:308-456 and
// model/spherical_model.py:238-314 of the reference), launching libofb's kernels on the
// caller's stream.  Host-side orchestration only; no device synchronisation.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "token_tc.cuh"

namespace ofb {

static thread_local std::string g_err;
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}
void count_launch(int n) { g_launches += n; }

int conv_simt(const ofb_conv_desc* d, cudaStream_t s);
int stem_tc(const void* patches, int n, int h, int w, const void* wgt_split, float wgt_unscale, const float* scale,
            const float* shift, void* out, cudaStream_t s);
int conv_tc(const ofb_conv_desc* d, cudaStream_t s);
bool conv_tc_supported(const ofb_conv_desc* d);
int conv_tc_heads(const void* x, int n, int h, int w, const void* wgt_split, float wgt_unscale, float b_pred,
                  float b_conf, int confidence, float* pred_out, float* conf_out, int interleaved, cudaStream_t s);
int attention_tc(const void* qkv, int B, int N, int heads, void* out, cudaStream_t s);
int conv_tc_chain(const ofb_conv_desc* descs, int L, unsigned int* flags, int flags_capacity, cudaStream_t s);
int blend_conf_launch(const float* pred_w, const float* conf, bool interleaved, int B, int N, int Ph, int Pw,
                      const int32_t* rowptr, const uint32_t* idx, const float* w, int He, int We, float* out,
                      cudaStream_t s);
int deinterleave_launch(const float* src_pairs, size_t n, int comp, float* dst, cudaStream_t s);
int zero_stem_pads(void* patches, int imgs, int P, cudaStream_t s);
int token_pack(const void* down, const float* pos_emb, int imgs, int N, int spatial, int cstride, void* tokens, int fmt, cudaStream_t s, int out_fmt = -1);
void token_stack_debug_stamps(long long* p);
int range_launch(const void* p, size_t n, int fmt, unsigned int* out2, cudaStream_t s);
long long* conv_tc_debug_buffer();
int conv_tc_timeline_slots();
void conv_tc_default_debug(int v);
void conv_tc_default_nstack(int v);
const char* conv_tc_last_variant();

int conv_dispatch(const ofb_conv_desc* d, cudaStream_t s) {
  OFB_CHECK(d && d->in0 && (d->wgt || d->wgt_split) && d->out, "conv: null pointer");
  OFB_CHECK(d->n > 0 && d->h > 0 && d->w > 0 && d->k > 0 && d->stride > 0, "conv: bad shape");
  if (d->engine == OFB_ENGINE_TC) {
    OFB_CHECK(conv_tc_supported(d), "conv: shape not supported by the tcgen05 engine");
    return conv_tc(d, s);
  }
  // AUTO only picks the tensor-core engine where it is fp32-accurate (split-half operands);
  // its single-pass TF32 mode on float32 tensors must be requested explicitly
  if (d->engine == OFB_ENGINE_AUTO && d->in_fmt == OFB_FMT_SPLIT16 && conv_tc_supported(d)) return conv_tc(d, s);
  OFB_CHECK(!d->ups2x, "conv: the fused 2x upsample exists only on the tcgen05 engine (split-half format, 32 -> 32 channels)");
  return conv_simt(d, s);
}

struct ConvW {
  float* w = nullptr; float* scale = nullptr; float* shift = nullptr;
  void* ws = nullptr; float unscale = 1.f;     // split-half planes of w * 2^e, and 2^-e
  int cout = 0, cin = 0, k = 0;
};
struct Mlp { float *w1, *s1, *t1, *w2, *s2, *t2; int cin; };
struct Block {
  float *n1g, *n1b, *n2g, *n2b;
  ConvW qkv, proj, fc1, fc2;     // qkv = rows of attn.q.weight followed by attn.kv.weight
};
struct Act { float* p = nullptr; int n = 0, h = 0, w = 0, c = 0, fmt = 0; };
constexpr int kRangeSlots = 32;
constexpr int kChainFlags = 1024;

}  // namespace ofb

using namespace ofb;

struct ofb_handle {
  int device = 0;
  bool single = false, has_geo = false, has_weights = false;
  ofb_geometry geo{};
  std::map<std::string, ConvW> conv;
  Block blk[6];
  float *pos_emb = nullptr, *enc_g = nullptr, *enc_b = nullptr;
  int pos_patches = 0;
  ConvW down;
  int down_cout = 32;              // logical output channels of the token reduction conv (down.cout is padded to 32s)
  float *pred_w = nullptr, *conf_w = nullptr;
  ConvW heads16;                   // both heads as one 16-channel 3x3 conv (0 = pred, 1 = weight_pred) for the tcgen05 engine
  int dbg_blocks = 6;              // timing experiments only: number of transformer blocks executed
  int splitk = 4;                  // K slices of the two 512-wide token linears on the tcgen05 engine (1 = off)
  int heads_tc = 1;                // run the heads on the tensor pipe (split-half format, 128-pixel rows)
  int attn_tc = 1;                 // attention core on the tensor pipe (split-half format; tcgen05 QK^T and PV)
  int conv_splitk = 0;             // 2-way split-K + finish kernel for layer4's 512 -> 512 convs in latency mode (see Ctx).
                                   // Off: measured neutral at 8 panoramas per step (the conv goes from 36 to 30 us per
                                   // launch, the finish kernel takes the difference; same-box A/B 1917 / 1924 vs 1921 / 1905)
  int token_fused = 1;             // the whole transformer stack as one launch, 16 CTAs per panorama (token_tc.cu):
                                   // 0 never, 1 when all panoramas of a chunk are resident at once, 2 whenever supported
  TokStack tok_stack{};            // its kernel argument (weight tensor maps + epilogue constants), built by load_weights
  bool tok_ready = false;
  int tok_groups = -1;             // groups of 16 CTAs of that kernel resident at once (-1: not asked yet)
  int chain = 1;                   // run the same-shape convs of an encoder stage as one image-stationary chain launch
                                   // (1: stages whose dependencies stay inside a CTA pair; 2: also cross-cluster chains)
  int no_point_feat = 0;           // ablation of network_360d.py:325 - layer1 is used without the point-feature add
  int no_transformer = 0;          // ablation of network_360d.py:330-335 - no token path, layer4 goes straight to the decoder
  int check_range = 0;             // after every forward: max |x| / non-finite count of each registered activation
  unsigned int* range_dev = nullptr;            // [kRangeSlots][2]
  unsigned int* chain_flags = nullptr;          // arrival counters of the layer chains (conv_tc.cu), kChainFlags entries
  std::vector<std::string> range_names;         // names of the slots filled by the last forward
  float pred_b = 0.f, conf_b = 0.f;
  Mlp mlp[2]{};
  std::vector<void*> owned;        // device allocations for weights
  // workspace
  // Two "lanes" (option lanes=2, off by default): a batch of >= 4 panoramas is split in two halves that run
  // concurrently on two streams (each with its own workspace) so that the latency-bound token path of one half
  // overlaps the convolutions of the other.  Measured on B200: 1561 vs 1898 panoramas/s at 8 per step, 1945 vs
  // 2136 at 32 - the halves' persistent kernels compete for the L2 fabric and for SMs instead of interleaving.
  struct Lane { float* ws = nullptr; size_t ws_floats = 0; };
  Lane lane[2];
  int cur = 0;                     // lane whose launches are being enqueued (host-side state)
  int lanes = 1;
  TcOptions tc;                    // tcgen05 launch variants of THIS handle (installed for the duration of a forward)
  long long ws_generation = 0;     // bumped whenever a workspace arena is (re)allocated: captured CUDA graphs that
                                   // replay launches into the old arena must be dropped (ofb_workspace_generation)
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int engine = OFB_ENGINE_AUTO, chunk = 0, dedup = 1;
  int fuse_ups = 1;                // fold the last decoder upsample into de_conv4_0's operand producer
  int fmt = OFB_FMT_SPLIT16;       // activation storage inside the network
  std::map<std::string, Act> acts;
  // optional per-launch timing (ofb_profile_enable)
  bool profile = false;
  struct Rec { std::string name; double flops, bytes; cudaEvent_t e0, e1; };
  std::vector<Rec> recs;
};

static void drop_profile_records(ofb_handle* h) {
  for (auto& r : h->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  h->recs.clear();
}

namespace ofb {

static int dev_upload(ofb_handle* h, const std::vector<float>& v, float** out) {
  float* d = nullptr;
  OFB_CUDA(cudaMalloc(&d, v.size() * sizeof(float)));
  h->owned.push_back(d);
  OFB_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

typedef std::map<std::string, const ofb_tensor_desc*> TMap;

static const ofb_tensor_desc* find(const TMap& m, const std::string& name) {
  auto it = m.find(name);
  if (it == m.end()) { set_error("load_weights: missing tensor '%s'", name.c_str()); return nullptr; }
  return it->second;
}
static long long numel(const ofb_tensor_desc* t) {
  long long n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= t->shape[i];
  return n;
}

// (O,I,kh,kw[,1]) or (O,I) -> OHWI, optional zero padding of I to cin_pad
static int pack_conv(ofb_handle* h, const TMap& m, const std::string& wname, ConvW* cw, int cin_pad = 0, int cout_pad = 0) {
  const ofb_tensor_desc* t = find(m, wname);
  if (!t) return -1;
  OFB_CHECK(t->ndim >= 2, "load_weights: '%s' must have >= 2 dims", wname.c_str());
  int O = (int)t->shape[0], I = (int)t->shape[1];
  int kh = t->ndim >= 3 ? (int)t->shape[2] : 1, kw = t->ndim >= 4 ? (int)t->shape[3] : 1;
  OFB_CHECK(kh == kw, "load_weights: '%s' non-square kernel", wname.c_str());
  OFB_CHECK(t->ndim < 5 || t->shape[4] == 1, "load_weights: '%s' trailing dim must be 1", wname.c_str());
  int Ip = cin_pad ? cin_pad : I;
  const int Op = cout_pad > O ? cout_pad : O;          // extra output channels = zero filters
  std::vector<float> p((size_t)Op * kh * kw * Ip, 0.f);
  for (int o = 0; o < O; ++o)
    for (int i = 0; i < I; ++i)
      for (int y = 0; y < kh; ++y)
        for (int x = 0; x < kw; ++x)
          p[(((size_t)o * kh + y) * kw + x) * Ip + i] = t->data[(((size_t)o * I + i) * kh + y) * kw + x];
  cw->cout = Op; cw->cin = Ip; cw->k = kh;
  if (dev_upload(h, p, &cw->w)) return -1;
  // split-half planes for the tcgen05 engine: scale by a power of two so max|w| lands in
  // [2^13, 2^14) and both the hi and the lo plane stay in fp16's normal range
  float mx = 0.f;
  for (float v : p) mx = fmaxf(mx, fabsf(v));
  int e = mx > 0.f ? 13 - (int)floorf(log2f(mx)) : 0;
  cw->unscale = ldexpf(1.f, -e);
  OFB_CUDA(cudaMalloc(&cw->ws, p.size() * 4));
  h->owned.push_back(cw->ws);
  // (legacy default stream: ordered after the synchronous upload above; load_all synchronises once at the end)
  if (ofb_split_f16(cw->w, p.size(), ldexpf(1.f, e), cw->ws, nullptr)) return -1;
  return 0;
}

static int pack_vec(ofb_handle* h, const TMap& m, const std::string& name, float** out, int expect = -1) {
  const ofb_tensor_desc* t = find(m, name);
  if (!t) return -1;
  OFB_CHECK(expect < 0 || numel(t) == expect, "load_weights: '%s' has %lld elements, expected %d", name.c_str(), numel(t), expect);
  std::vector<float> v(t->data, t->data + numel(t));
  return dev_upload(h, v, out);
}

// BatchNorm (eval) -> y = x*scale + shift
static int pack_bn(ofb_handle* h, const TMap& m, const std::string& prefix, int c, float eps, float** scale, float** shift) {
  const ofb_tensor_desc *g = find(m, prefix + ".weight"), *b = find(m, prefix + ".bias"),
                        *mu = find(m, prefix + ".running_mean"), *var = find(m, prefix + ".running_var");
  if (!g || !b || !mu || !var) return -1;
  OFB_CHECK(numel(g) == c && numel(b) == c && numel(mu) == c && numel(var) == c, "load_weights: '%s' BN size mismatch", prefix.c_str());
  std::vector<float> s(c), t(c);
  for (int i = 0; i < c; ++i) {
    double sc = (double)g->data[i] / sqrt((double)var->data[i] + (double)eps);
    s[i] = (float)sc;
    t[i] = (float)((double)b->data[i] - (double)mu->data[i] * sc);
  }
  if (dev_upload(h, s, scale)) return -1;
  return dev_upload(h, t, shift);
}

static int conv_bn(ofb_handle* h, const TMap& m, const std::string& key, const std::string& wname, const std::string& bn) {
  ConvW cw;
  if (pack_conv(h, m, wname, &cw)) return -1;
  if (pack_bn(h, m, bn, cw.cout, 1e-5f, &cw.scale, &cw.shift)) return -1;
  h->conv[key] = cw;
  return 0;
}

static int linear(ofb_handle* h, const TMap& m, const std::string& prefix, bool bias, ConvW* cw, int cout_pad = 0) {
  if (pack_conv(h, m, prefix + ".weight", cw, 0, cout_pad)) return -1;
  if (bias) {
    const ofb_tensor_desc* t = find(m, prefix + ".bias");
    if (!t) return -1;
    OFB_CHECK(numel(t) <= cw->cout, "load_weights: '%s.bias' has %lld elements for %d outputs", prefix.c_str(), numel(t), cw->cout);
    std::vector<float> v(cw->cout, 0.f);
    for (long long i = 0; i < numel(t); ++i) v[i] = t->data[i];
    if (dev_upload(h, v, &cw->shift)) return -1;
  }
  return 0;
}

static int load_mlp(ofb_handle* h, const TMap& m, const std::string& p, Mlp* o) {
  const ofb_tensor_desc* w1 = find(m, p + ".0.weight");
  if (!w1) return -1;
  o->cin = (int)w1->shape[1];
  if (pack_vec(h, m, p + ".0.weight", &o->w1, 16 * o->cin)) return -1;
  if (pack_bn(h, m, p + ".1", 16, 1e-5f, &o->s1, &o->t1)) return -1;
  if (pack_vec(h, m, p + ".3.weight", &o->w2, 64 * 16)) return -1;
  return pack_bn(h, m, p + ".4", 64, 1e-5f, &o->s2, &o->t2);
}

static void free_weights(ofb_handle* h) {
  for (void* p : h->owned) cudaFree(p);
  h->owned.clear();
  h->conv.clear();
  h->has_weights = false;
}

static const int kBlocks[4] = {3, 4, 6, 3};
static const int kChan[4] = {64, 128, 256, 512};

static int load_all(ofb_handle* h, const TMap& m, bool single) {
  // stem: (64,3,7,7,1) -> OHWI with the input channel padded to 4
  ConvW stem;
  {  // (64,3,7,7,1) -> [kh][kw][cin padded to 4][cout]: the layout the stem kernel bulk-copies to smem
    const ofb_tensor_desc* t = find(m, "conv1.weight");
    if (!t) return -1;
    OFB_CHECK(numel(t) == 64 * 3 * 49, "load_weights: conv1.weight must be (64,3,7,7,1)");
    std::vector<float> p(49 * 4 * 64, 0.f);
    for (int o = 0; o < 64; ++o)
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 49; ++k) p[(k * 4 + i) * 64 + o] = t->data[(o * 3 + i) * 49 + k];
    stem.cout = 64; stem.cin = 4; stem.k = 7;
    if (dev_upload(h, p, &stem.w)) return -1;
    // tensor-core stem: [cout][kh][kw' = kw + 1][cin padded to 4], kw' = 0 zero (ofb_stem_tc_f16)
    std::vector<float> q(64 * 7 * 8 * 4, 0.f);
    float mx = 0.f;
    for (int o = 0; o < 64; ++o)
      for (int i = 0; i < 3; ++i)
        for (int kh = 0; kh < 7; ++kh)
          for (int kw = 0; kw < 7; ++kw) {
            float v = t->data[((o * 3 + i) * 7 + kh) * 7 + kw];
            q[((o * 7 + kh) * 8 + kw + 1) * 4 + i] = v;
            mx = fmaxf(mx, fabsf(v));
          }
    int ex = mx > 0.f ? 13 - (int)floorf(log2f(mx)) : 0;
    stem.unscale = ldexpf(1.f, -ex);
    float* tmp = nullptr;
    if (dev_upload(h, q, &tmp)) return -1;
    OFB_CUDA(cudaMalloc(&stem.ws, q.size() * 4));
    h->owned.push_back(stem.ws);
    if (ofb_split_f16(tmp, q.size(), ldexpf(1.f, ex), stem.ws, nullptr)) return -1;
  }
  if (pack_bn(h, m, "bn1", 64, 1e-5f, &stem.scale, &stem.shift)) return -1;
  h->conv["stem"] = stem;
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < kBlocks[l]; ++b) {
      std::string p = "layer" + std::to_string(l + 1) + "." + std::to_string(b);
      if (conv_bn(h, m, p + ".conv1", p + ".conv1.weight", p + ".bn1")) return -1;
      if (conv_bn(h, m, p + ".conv2", p + ".conv2.weight", p + ".bn2")) return -1;
      if (m.count(p + ".downsample.0.weight"))
        if (conv_bn(h, m, p + ".ds", p + ".downsample.0.weight", p + ".downsample.1")) return -1;
    }
  {  // token reduction conv: 512 -> 32 (128x128 patches) or 512 -> 8 (256x256, network_test.py:271); the conv engines
     // work on multiples of 32 output channels, so a narrower one is padded with zero filters (token_pack skips them)
    const ofb_tensor_desc* dw = find(m, single ? "down.weight" : "down1.weight");
    if (!dw) return -1;
    h->down_cout = (int)dw->shape[0];
    if (linear(h, m, single ? "down" : "down1", true, &h->down, (h->down_cout + 31) / 32 * 32)) return -1;
  }
  const ofb_tensor_desc* pe = find(m, "transformer.pos_emb");
  if (!pe) return -1;
  h->pos_patches = (int)pe->shape[1];
  if (pack_vec(h, m, "transformer.pos_emb", &h->pos_emb, h->pos_patches * 512)) return -1;
  if (pack_vec(h, m, "transformer.encoder_norm.weight", &h->enc_g, 512)) return -1;
  if (pack_vec(h, m, "transformer.encoder_norm.bias", &h->enc_b, 512)) return -1;
  for (int i = 0; i < 6; ++i) {
    std::string p = "transformer.layer." + std::to_string(i);
    Block& B = h->blk[i];
    if (pack_vec(h, m, p + ".norm1.weight", &B.n1g, 512) || pack_vec(h, m, p + ".norm1.bias", &B.n1b, 512) ||
        pack_vec(h, m, p + ".norm2.weight", &B.n2g, 512) || pack_vec(h, m, p + ".norm2.bias", &B.n2b, 512))
      return -1;
    {  // q and kv projections share their input: one (1536,512) weight, one launch
      const ofb_tensor_desc *tq = find(m, p + ".attn.q.weight"), *tkv = find(m, p + ".attn.kv.weight");
      if (!tq || !tkv) return -1;
      OFB_CHECK(numel(tq) == 512 * 512 && numel(tkv) == 1024 * 512, "load_weights: attention projection shapes");
      std::vector<float> cat(1536 * 512);
      memcpy(cat.data(), tq->data, 512 * 512 * sizeof(float));
      memcpy(cat.data() + 512 * 512, tkv->data, 1024 * 512 * sizeof(float));
      ofb_tensor_desc fused{};
      std::string nm = p + ".attn.qkv.weight";
      fused.name = nm.c_str(); fused.data = cat.data(); fused.ndim = 2; fused.shape[0] = 1536; fused.shape[1] = 512;
      TMap one; one[nm] = &fused;
      if (pack_conv(h, one, nm, &B.qkv)) return -1;
    }
    if (linear(h, m, p + ".attn.proj", true, &B.proj) || linear(h, m, p + ".mlp.fc1", true, &B.fc1) ||
        linear(h, m, p + ".mlp.fc2", true, &B.fc2))
      return -1;
  }
  const char* dec[9] = {"de_conv0_0", "de_conv0_1", "de_conv1_0", "de_conv1_1", "de_conv2_0",
                        "de_conv2_1", "de_conv3_0", "de_conv3_1", "de_conv4_0"};
  for (int i = 0; i < 9; ++i)
    if (conv_bn(h, m, dec[i], std::string(dec[i]) + ".conv.weight", std::string(dec[i]) + ".bn")) return -1;
  // heads: (1,32,3,3,1) -> (3,3,32)
  ConvW hp, hc;
  if (pack_conv(h, m, "pred.weight", &hp) || pack_conv(h, m, "weight_pred.weight", &hc)) return -1;
  h->pred_w = hp.w; h->conf_w = hc.w;
  {  // (16,32,3,3): rows 0/1 = pred / weight_pred, the rest zero
    const ofb_tensor_desc *tp = find(m, "pred.weight"), *tc = find(m, "weight_pred.weight");
    if (!tp || !tc) return -1;
    OFB_CHECK(numel(tp) == 288 && numel(tc) == 288, "load_weights: head filters must be (1,32,3,3,1)");
    std::vector<float> cat(16 * 288, 0.f);
    memcpy(cat.data(), tp->data, 288 * sizeof(float));
    memcpy(cat.data() + 288, tc->data, 288 * sizeof(float));
    ofb_tensor_desc fused{};
    std::string nm = "heads16.weight";
    fused.name = nm.c_str(); fused.data = cat.data(); fused.ndim = 4;
    fused.shape[0] = 16; fused.shape[1] = 32; fused.shape[2] = 3; fused.shape[3] = 3;
    TMap one; one[nm] = &fused;
    if (pack_conv(h, one, nm, &h->heads16)) return -1;
  }
  const ofb_tensor_desc *pb = find(m, "pred.bias"), *cb = find(m, "weight_pred.bias");
  if (!pb || !cb) return -1;
  h->pred_b = pb->data[0]; h->conf_b = cb->data[0];
  if (single) {
    if (load_mlp(h, m, "mlp_points", &h->mlp[0])) return -1;
  } else {
    if (load_mlp(h, m, "mlp_points1", &h->mlp[0]) || load_mlp(h, m, "mlp_points2", &h->mlp[1])) return -1;
  }
  return 0;
}

// ------------------------------------------------------------------ workspace
struct Plan {
  ofb_handle* h; size_t off = 0;
  float* take(size_t n) {
    size_t a = (off + 63) & ~(size_t)63;   // 256-byte alignment
    off = a + n;
    return h->lane[h->cur].ws ? h->lane[h->cur].ws + a : nullptr;
  }
};

struct Buffers {
  float *patches, *conv1, *pool, *l1t, *l1a, *l1b, *layer1_pre, *layer1;
  float *l2t, *l2a, *l2b, *l2d, *layer2, *l3t, *l3a, *l3b, *l3d, *layer3, *l4t, *l4a, *l4b, *l4d, *layer4;
  float *down, *tok, *ln, *q, *kv, *att, *tok2, *fc1, *enc, *part, *tokx, *cpart;
  float *up0, *d00, *d01, *up1, *d10, *d11, *up2, *d20, *d21, *up3, *d30, *d31, *up4, *d40;
  float *pred, *conf, *depth_p;
};

static size_t plan_buffers(ofb_handle* h, int imgs, int P, Buffers* b) {
  Plan pl{h};
  size_t I = (size_t)imgs;
  int p2 = P / 2, p4 = P / 4, p8 = P / 8, p16 = P / 16, p32 = P / 32;
  b->patches = pl.take(I * P * (P + 8) * 4);       // room for the row-padded stem layout
  b->conv1 = pl.take(I * p2 * p2 * 64);
  b->pool = pl.take(I * p4 * p4 * 64);
  size_t s1 = I * p4 * p4 * 64, s2 = I * p8 * p8 * 128, s3 = I * p16 * p16 * 256, s4 = I * p32 * p32 * 512;
  b->l1t = pl.take(s1); b->l1a = pl.take(s1); b->l1b = pl.take(s1); b->layer1_pre = pl.take(s1); b->layer1 = pl.take(s1);
  b->l2t = pl.take(s2); b->l2a = pl.take(s2); b->l2b = pl.take(s2); b->l2d = pl.take(s2); b->layer2 = pl.take(s2);
  b->l3t = pl.take(s3); b->l3a = pl.take(s3); b->l3b = pl.take(s3); b->l3d = pl.take(s3); b->layer3 = pl.take(s3);
  b->l4t = pl.take(s4); b->l4a = pl.take(s4); b->l4b = pl.take(s4); b->l4d = pl.take(s4); b->layer4 = pl.take(s4);
  b->down = pl.take(I * (size_t)p32 * p32 * 32 > I * 512 ? I * (size_t)p32 * p32 * 32 : I * 512); b->tok = pl.take(I * 512); b->ln = pl.take(I * 512); b->q = pl.take(I * 512);
  b->kv = pl.take(I * 1536); b->att = pl.take(I * 512); b->tok2 = pl.take(I * 512); b->fc1 = pl.take(I * 2048);
  b->enc = pl.take(I * 512);
  b->part = pl.take(I * 512 * 4);                  // split-K partial sums of attn.proj / mlp.fc2 (4 slices)
  b->cpart = pl.take(2 * s4);                      // partial sums of layer4's split-K convs (2 slices)
  b->tokx = pl.take(I * (512 * 17 + 16 * 48));     // exchange buffers of the fused transformer stack (token_stack_scratch_floats, N <= 48)
  b->up0 = pl.take(I * p16 * p16 * 512); b->d00 = pl.take(I * p16 * p16 * 256); b->d01 = pl.take(I * p16 * p16 * 128);
  b->up1 = pl.take(I * p8 * p8 * 128); b->d10 = pl.take(I * p8 * p8 * 128); b->d11 = pl.take(I * p8 * p8 * 64);
  b->up2 = pl.take(I * p4 * p4 * 64); b->d20 = pl.take(I * p4 * p4 * 64); b->d21 = pl.take(I * p4 * p4 * 64);
  b->up3 = pl.take(I * p2 * p2 * 64); b->d30 = pl.take(I * p2 * p2 * 64); b->d31 = pl.take(I * p2 * p2 * 32);
  b->up4 = pl.take(I * P * P * 32); b->d40 = pl.take(I * P * P * 32);
  // head outputs: either two maps (pred | conf, back to back) or one interleaved (pred*conf, conf) pair map
  b->pred = pl.take(2 * I * P * P); b->conf = b->pred ? b->pred + I * P * P : nullptr; b->depth_p = pl.take(I * p4 * p4);
  return pl.off;
}

static int ensure_workspace(ofb_handle* h, int imgs, int P, Buffers* b) {
  Buffers tmp;
  float* saved = h->lane[h->cur].ws;
  h->lane[h->cur].ws = nullptr;
  size_t need = plan_buffers(h, imgs, P, &tmp);
  h->lane[h->cur].ws = saved;
  if (need > h->lane[h->cur].ws_floats) {
    if (h->lane[h->cur].ws) cudaFree(h->lane[h->cur].ws);
    h->lane[h->cur].ws = nullptr; h->lane[h->cur].ws_floats = 0;
    OFB_CUDA(cudaMalloc(&h->lane[h->cur].ws, need * sizeof(float)));
    h->lane[h->cur].ws_floats = need;
    ++h->ws_generation;
  }
  if (!h->chain_flags) OFB_CUDA(cudaMalloc(&h->chain_flags, 2 * kChainFlags * sizeof(unsigned int)));   // one set per lane
  plan_buffers(h, imgs, P, b);
  return 0;
}

// ----------------------------------------------------------------- schedule
struct Ctx { ofb_handle* h; cudaStream_t s; int imgs; bool latency = false; float* cpart = nullptr; };
// latency: the chunk holds so few panoramas (<= 9) that the small layers cannot fill the GPU - the forward then uses
// the fused transformer stack and 2-way split-K for layer4's 512 -> 512 convs (cpart = their partial-sum buffer)

// Brackets one launch with CUDA events on the launch stream when profiling is on.
struct Prof {
  ofb_handle* h; cudaStream_t s; bool on;
  Prof(ofb_handle* h_, cudaStream_t s_, const std::string& name, double flops, double bytes) : h(h_), s(s_), on(h_->profile) {
    if (!on) return;
    ofb_handle::Rec r; r.name = name; r.flops = flops; r.bytes = bytes;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, s);
    h->recs.push_back(r);
  }
  ~Prof() { if (on) cudaEventRecord(h->recs.back().e1, s); }
};

static std::string conv_class(const ofb_conv_desc& d) {
  char buf[96];
  int oh = (d.h + 2 * d.pad - d.k) / d.stride + 1;
  snprintf(buf, sizeof(buf), "conv%dx%ds%d_c%d_o%d_@%d", d.k, d.k, d.stride, d.c0 + d.c1, d.cout, oh);
  return buf;
}
static void conv_work(const ofb_conv_desc& d, double* flops, double* bytes) {
  double oh = (d.h + 2 * d.pad - d.k) / d.stride + 1, ow = (d.w + 2 * d.pad - d.k) / d.stride + 1;
  double M = (double)d.n * oh * ow, K = (double)d.k * d.k * (d.c0 + d.c1);
  *flops = 2.0 * M * d.cout * K;
  *bytes = 4.0 * ((double)d.n * d.h * d.w * (d.c0 + d.c1) + M * d.cout * (d.residual ? 2 : 1) + (double)d.cout * K);
}

static ofb_conv_desc conv_desc(Ctx& c, const ConvW& w, const float* in0, int c0, const float* in1, int c1, int hh, int ww,
                               int stride, int pad, const float* residual, int act, float* out, int ups2x = 0) {
  ofb_conv_desc d{};
  d.ups2x = ups2x;
  d.in0 = in0; d.in1 = in1; d.c0 = c0; d.c1 = c1; d.n = c.imgs; d.h = hh; d.w = ww;
  d.wgt = w.w; d.k = w.k; d.stride = stride; d.pad = pad; d.cout = w.cout;
  d.scale = w.scale; d.shift = w.shift; d.residual = residual; d.act = act; d.out = out;
  d.engine = c.h->engine;
  d.in_fmt = d.out_fmt = c.h->fmt; d.wgt_split = w.ws; d.wgt_unscale = w.unscale;
  return d;
}

static int run_conv(Ctx& c, const ConvW& w, const float* in0, int c0, const float* in1, int c1, int hh, int ww,
                    int stride, int pad, const float* residual, int act, float* out, int ups2x = 0) {
  ofb_conv_desc d = conv_desc(c, w, in0, c0, in1, c1, hh, ww, stride, pad, residual, act, out, ups2x);
  OFB_CHECK(w.w && w.cin == c0 + c1, "forward: conv weight/channel mismatch (%d vs %d+%d)", w.cin, c0, c1);
  double fl, by;
  conv_work(d, &fl, &by);
  if (ups2x) by -= 4.0 * 0.75 * (double)d.n * d.h * d.w * (d.c0 + d.c1);     // reads the low-resolution tensor
  Prof pr(c.h, c.s, conv_class(d) + (ups2x ? "_ups" : ""), fl, by);
  return conv_dispatch(&d, c.s);
}

// can the 2x upsample in front of this conv be folded into its operand producer?
static bool can_fuse_ups(ofb_handle* h, const ConvW& w, int n, int hh, int ww) {
  if (!h->fuse_ups || h->fmt != OFB_FMT_SPLIT16 || h->engine == OFB_ENGINE_SIMT) return false;
  ofb_conv_desc d{};
  d.ups2x = 1; d.c0 = w.cin; d.n = n; d.h = hh; d.w = ww; d.k = w.k; d.stride = 1; d.pad = 1; d.cout = w.cout;
  d.in_fmt = d.out_fmt = h->fmt; d.wgt_split = w.ws;
  return conv_tc_supported(&d);
}

static int run_linear(Ctx& c, const ConvW& w, const float* in, const float* residual, int act, float* out,
                      int ksplit = 1, float* partial = nullptr) {
  ofb_conv_desc d{};
  d.in0 = in; d.c0 = w.cin; d.n = c.imgs; d.h = 1; d.w = 1;
  d.wgt = w.w; d.k = 1; d.stride = 1; d.pad = 0; d.cout = w.cout;
  d.scale = nullptr; d.shift = w.shift; d.residual = residual; d.act = act; d.out = out;
  d.ksplit = ksplit; d.partial = partial;      // ksplit > 1: raw partial sums only (`out` just anchors the unused store maps)
  d.engine = c.h->engine;
  d.in_fmt = d.out_fmt = c.h->fmt; d.wgt_split = w.ws; d.wgt_unscale = w.unscale;
  double fl, by;
  conv_work(d, &fl, &by);
  char nm[64];
  snprintf(nm, sizeof(nm), "linear_k%d_o%d", w.cin, w.cout);
  Prof pr(c.h, c.s, nm, fl, by);
  return conv_dispatch(&d, c.s);
}

// One conv of a residual stage.  In latency mode a 3x3 stride-1 conv with K >= 4608 and few pixels (layer4: 18 M tiles
// at 8 panoramas) runs 2-way split-K: every CTA then fetches half of the activations and weights, which is what
// bounds these launches (L2 -> SM at the slice cap), and a finish kernel applies BN / residual / ReLU.
static int run_stage_conv(Ctx& c, ofb_conv_desc d) {
  double f, b_;
  conv_work(d, &f, &b_);
  const bool sk = c.latency && c.cpart && c.h->conv_splitk && c.h->fmt == OFB_FMT_SPLIT16 && c.h->engine != OFB_ENGINE_SIMT &&
                  d.k == 3 && d.stride == 1 && !d.in1 && !d.ups2x && d.cout >= 256 && 9 * d.c0 >= 4608 && d.h * d.w <= 16;
  if (!sk) {
    Prof pr(c.h, c.s, conv_class(d), f, b_);
    return conv_dispatch(&d, c.s);
  }
  d.ksplit = 2; d.partial = c.cpart;
  { Prof pr(c.h, c.s, conv_class(d) + "_splitk2", f, b_);
    if (conv_dispatch(&d, c.s)) return -1; }
  const long long pixels = (long long)d.n * d.h * d.w;
  Prof pr(c.h, c.s, "splitk_finish_conv", 0.0, 4.0 * pixels * d.cout * (2 + 1 + (d.residual ? 1 : 0)));
  return ofb_splitk_finish_conv_f16(c.cpart, 2, pixels, d.cout, d.scale, d.shift, d.wgt_unscale, d.residual, d.act, d.out, (void*)c.s);
}

// torchvision BasicBlock stack: relu(bn2(conv2(relu(bn1(conv1 x)))) + identity/downsample)
static int run_res_layer(Ctx& c, int l, const float* in, int cin, int hin, float* tmp, float* pa, float* pb,
                         float* ds, float* out_final) {
  int ch = kChan[l], nb = kBlocks[l];
  int stride = l == 0 ? 1 : 2;
  int hout = hin / stride;
  const float* x = in;
  int xc = cin, xh = hin;
  // Image-stationary chain: everything after the first block's strided conv / downsample is a run of same-shape 3x3
  // convs - [b0.conv2, b1.conv1, b1.conv2, ...] - that one launch walks layer by layer (conv_tc.cu: conv_chain_kernel)
  // when the stage's geometry keeps every dependency inside a CTA pair (layer2 at 128x128 patches); same buffers,
  // same arithmetic, bit-identical results.
  const bool try_chain = c.h->chain && c.h->fmt == OFB_FMT_SPLIT16 && c.h->engine != OFB_ENGINE_SIMT &&
                         stride == 2 && nb >= 2;
  if (try_chain) {
    std::vector<ofb_conv_desc> ds_;
    const float* cx = in;
    const float* idn = nullptr;
    bool ok = true;
    for (int b = 0; b < nb && ok; ++b) {
      std::string p = "layer" + std::to_string(l + 1) + "." + std::to_string(b);
      float* y = (b == nb - 1) ? out_final : ((b & 1) ? pb : pa);
      if (b == 0) {
        ok = c.h->conv.count(p + ".ds") != 0;
        idn = ds;
      } else {
        ds_.push_back(conv_desc(c, c.h->conv[p + ".conv1"], cx, ch, nullptr, 0, hout, hout, 1, 1, nullptr, OFB_ACT_RELU, tmp));
        idn = cx;
      }
      ds_.push_back(conv_desc(c, c.h->conv[p + ".conv2"], tmp, ch, nullptr, 0, hout, hout, 1, 1, idn, OFB_ACT_RELU, y));
      cx = y;
    }
    if (ok) {
      // first block's strided conv1 and downsample as ordinary launches, then the chain
      std::string p0 = "layer" + std::to_string(l + 1) + ".0";
      if (run_conv(c, c.h->conv[p0 + ".conv1"], in, cin, nullptr, 0, hin, hin, stride, 1, nullptr, OFB_ACT_RELU, tmp)) return -1;
      if (run_conv(c, c.h->conv[p0 + ".ds"], in, cin, nullptr, 0, hin, hin, stride, 0, nullptr, OFB_ACT_NONE, ds)) return -1;
      double fl = 0, by = 0;
      for (auto& d : ds_) { double f, b_; conv_work(d, &f, &b_); fl += f; by += b_; }
      int rc;
      { Prof pr(c.h, c.s, conv_class(ds_[0]) + "_chain" + std::to_string(ds_.size()), fl, by);
        // chain = 1: only the chains whose dependencies stay inside a cluster; 2: also those that hand images over
        // between clusters through global arrival counters (layer3; measured: no gain in a graph replay)
        rc = conv_tc_chain(ds_.data(), (int)ds_.size(), c.h->chain >= 2 ? c.h->chain_flags + (size_t)c.h->cur * kChainFlags : nullptr,
                           kChainFlags, c.s); }
      if (rc < 0) return -1;
      if (rc == 0) return 0;
      // not chainable for this shape: the launches above are exactly the loop's first two; continue it from conv2
      if (c.h->profile) { cudaEventDestroy(c.h->recs.back().e0); cudaEventDestroy(c.h->recs.back().e1); c.h->recs.pop_back(); }
      for (auto& d : ds_)
        if (run_stage_conv(c, d)) return -1;
      return 0;
    }
  }
  for (int b = 0; b < nb; ++b) {
    std::string p = "layer" + std::to_string(l + 1) + "." + std::to_string(b);
    int st = b == 0 ? stride : 1;
    float* y = (b == nb - 1) ? out_final : ((b & 1) ? pb : pa);
    if (run_stage_conv(c, conv_desc(c, c.h->conv[p + ".conv1"], x, xc, nullptr, 0, xh, xh, st, 1, nullptr, OFB_ACT_RELU, tmp))) return -1;
    const float* idn = x;
    auto it = c.h->conv.find(p + ".ds");
    if (it != c.h->conv.end()) {
      if (run_conv(c, it->second, x, xc, nullptr, 0, xh, xh, st, 0, nullptr, OFB_ACT_NONE, ds)) return -1;
      idn = ds;
    }
    if (run_stage_conv(c, conv_desc(c, c.h->conv[p + ".conv2"], tmp, ch, nullptr, 0, hout, hout, 1, 1, idn, OFB_ACT_RELU, y))) return -1;
    x = y; xc = ch; xh = hout;
  }
  return 0;
}

static void reg(ofb_handle* h, const char* name, float* p, int n, int hh, int ww, int cc, int fmt = -1) {
  Act a; a.p = p; a.n = n; a.h = hh; a.w = ww; a.c = cc; a.fmt = fmt < 0 ? h->fmt : fmt;
  h->acts[name] = a;
}

static int forward_chunk(ofb_handle* h, const float* rgb, int Bc, int iters, int confidence,
                         float* const* outs, size_t out_off, cudaStream_t s) {
  const ofb_geometry& g = h->geo;
  const int N = g.n_patch, P = g.patch, p4 = P / 4, He = g.erp_h, We = g.erp_w;
  const int imgs = Bc * N;
  Buffers b;
  if (ensure_workspace(h, imgs, P, &b)) return -1;
  Ctx c{h, s, imgs};
  if (h->tok_groups < 0) h->tok_groups = token_stack_resident_groups(N <= 32 ? 18 : 46);
  c.latency = Bc <= h->tok_groups;       // few panoramas per chunk: fused transformer stack, split-K layer4 (see Ctx)
  c.cpart = b.cpart;
  void* vs = (void*)s;
  const int F = h->fmt;
  bool pairs = false;          // head outputs of the last iteration are an interleaved pair map

  for (int it = 0; it < iters; ++it) {
    bool reuse = it > 0 && h->dedup;
    if (!reuse) {
      // equi2pers(rgb, P) -> patches; stem; pool; layer1 (spherical_model_iterative.py:315,322-324)
      const ConvW& st = h->conv["stem"];
      const bool stem_tc_path = F == OFB_FMT_SPLIT16 && h->engine != OFB_ENGINE_SIMT;
      // the stem layout's zero row pads: re-zeroed by every forward, inside the captured region (layers.cu)
      if (stem_tc_path && zero_stem_pads(b.patches, imgs, P, s)) return -1;
      { Prof pr(h, s, "e2p_rgb", 0.0, 4.0*((double)Bc*3*He*We + (double)imgs*P*P*4));
      if (ofb_equi2pers_f32(rgb, Bc, 3, He, We, g.grid_hi, N, P, P, b.patches,
                            stem_tc_path ? OFB_LAYOUT_STEM16 : OFB_LAYOUT_FOLDED, vs)) return -1; }
      { Prof pr(h, s, "stem7x7", 2.0*imgs*(P/2)*(P/2)*64*147.0, 4.0*((double)imgs*P*P*4 + (double)imgs*(P/2)*(P/2)*64));
      if (stem_tc_path) {
        if (ofb_stem_tc_f16(b.patches, imgs, P, P, st.ws, st.unscale, st.scale, st.shift, b.conv1, vs)) return -1;
      } else {
        if (ofb_stem_f32(b.patches, imgs, P, P, st.w, st.scale, st.shift, b.conv1, F, vs)) return -1;
      } }
      { Prof pr(h, s, "maxpool", 0.0, 4.0*((double)imgs*(P/2)*(P/2)*64 + (double)imgs*p4*p4*64));
      if (ofb_maxpool3x3s2_f32(b.conv1, imgs, P / 2, P / 2, 64, b.pool, F, vs)) return -1; }
      if (run_res_layer(c, 0, b.pool, 64, p4, b.l1t, b.l1a, b.l1b, nullptr, b.layer1_pre)) return -1;
    }
    // point embedding added to layer1 (:319-320,325 / :385-393)
    const float* depth = nullptr;
    if (it > 0) {
      { Prof pr(h, s, "e2p_depth", 0.0, 4.0*((double)Bc*He*We + (double)imgs*p4*p4));
      if (ofb_equi2pers_f32(outs[it - 1] + out_off, Bc, 1, He, We, g.grid_lo, N, p4, p4, b.depth_p,
                            OFB_LAYOUT_FOLDED, vs)) return -1; }
      depth = b.depth_p;
    }
    float* const layer1 = h->no_point_feat ? b.layer1_pre : b.layer1;
    if (!h->no_point_feat) {
      const Mlp& mp = h->mlp[it > 0 ? 1 : 0];
      OFB_CHECK(mp.cin == g.pts_c, "forward: point table has %d channels, mlp expects %d", g.pts_c, mp.cin);
      { Prof pr(h, s, "point_embed", 0.0, 4.0*((double)imgs*p4*p4*128));
      if (ofb_point_embed_f32(g.pts, N, mp.cin, p4, depth, imgs, mp.w1, mp.s1, mp.t1, mp.w2, mp.s2, mp.t2,
                              b.layer1_pre, b.layer1, F, vs)) return -1; }
    }
    if (run_res_layer(c, 1, layer1, 64, p4, b.l2t, b.l2a, b.l2b, b.l2d, b.layer2)) return -1;
    if (run_res_layer(c, 2, b.layer2, 128, P / 8, b.l3t, b.l3a, b.l3b, b.l3d, b.layer3)) return -1;
    if (run_res_layer(c, 3, b.layer3, 256, P / 16, b.l4t, b.l4a, b.l4b, b.l4d, b.layer4)) return -1;

    // tokens: down1 1x1 conv (+bias) over the SxSx512 map (S = P/32) -> (imgs,512) (:330-331)
    const int S32 = P / 32;
    if (!h->no_transformer) {
    OFB_CHECK(h->down_cout * S32 * S32 == 512, "forward: down conv has %d channels, %dx%d patches need %d (token width 512)",
              h->down_cout, P, P, 512 / (S32 * S32));
    {
      Ctx c16{h, s, imgs * S32 * S32};
      if (run_linear(c16, h->down, b.layer4, nullptr, OFB_ACT_NONE, b.down)) return -1;
    }
    OFB_CHECK(h->pos_patches == N, "forward: pos_emb has %d patches, geometry has %d", h->pos_patches, N);
    const int nblk = h->dbg_blocks;
    bool fused_tokens = h->token_fused && h->tok_ready && F == OFB_FMT_SPLIT16 && h->engine != OFB_ENGINE_SIMT &&
                        nblk > 0 && token_stack_supported(N);
    if (fused_tokens && h->token_fused == 1) {
      // The fused stack is the low-latency form: a group of 16 CTAs per panorama, all groups of a launch resident at
      // once (9 on 148 SMs).  More panoramas than that run in waves that each stream the weights again, and the
      // 48-token instantiation has a two-stage ring: measured slower than the per-layer path there (32 panoramas:
      // 1.42 vs 0.84 ms per step; 16 x 46 tokens: 1.86 vs 1.43 ms), faster below (8 x 18: 0.35 vs 0.53 ms, 1 x 18:
      // -0.14 ms, 8 x 26: -0.13 ms).  Option token_fused = 2 forces it.
      fused_tokens = N <= 32 && c.latency;
    }
    fused_tokens = fused_tokens && Bc <= kTokMaxPanos;
    { Prof pr(h, s, "token_pack", 0.0, 4.0*((double)imgs*1024));
    if (token_pack(b.down, h->pos_emb, imgs, N, S32, h->down.cout, b.tok, F, s, fused_tokens ? OFB_FMT_F32 : F)) return -1; }
    float* x = b.tok;
    float* y = b.tok2;
    const int SK = (F == OFB_FMT_SPLIT16 && h->engine != OFB_ENGINE_SIMT) ? h->splitk : 1;
    if (fused_tokens) {
      // all Transformer_Blocks (model/blocks.py:50-88) + encoder_norm in one launch: a group of 16 CTAs per panorama
      TokStack st = h->tok_stack;
      st.nblk = nblk;
      Prof pr(h, s, "token_stack", 2.0*(double)imgs*nblk*(512.0*1536 + 512.0*512 + 2*512.0*2048), 4.0*nblk*3.15e6);
      if (token_stack_launch(st, x, b.tokx, token_stack_scratch_floats(Bc, N), b.enc, Bc, N, 0, h->cur, h->tc.pdl, s)) return -1;
    } else if (SK > 1 && nblk > 0) {
      // attn.proj and mlp.fc2 run split-K; their finish kernels add bias + residual and apply the NEXT LayerNorm
      { Prof pr(h, s, "layernorm", 0.0, 4.0*((double)imgs*1024));
      if (ofb_layernorm_f32(x, h->blk[0].n1g, h->blk[0].n1b, imgs, 512, 1e-5f, b.ln, F, F, vs)) return -1; }
      for (int i = 0; i < nblk; ++i) {   // Transformer_Block, model/blocks.py:84-88
        Block& B = h->blk[i];
        if (run_linear(c, B.qkv, b.ln, nullptr, OFB_ACT_NONE, b.kv)) return -1;     // (imgs,1536) = [q | k | v]
        { Prof pr(h, s, "attention", 0.0, 4.0*((double)imgs*2048));
        if (h->attn_tc) { if (attention_tc(b.kv, Bc, N, 4, b.att, s)) return -1; }
        else if (ofb_attention_qkv_f32(b.kv, Bc, N, 4, 128, b.att, F, vs)) return -1; }
        if (run_linear(c, B.proj, b.att, nullptr, OFB_ACT_NONE, b.kv, SK, b.part)) return -1;
        { Prof pr(h, s, "splitk_finish_ln", 0.0, 4.0*((double)imgs*512*(SK + 3)));     // y = x + proj(att); ln = norm2(y)
        if (ofb_splitk_finish_ln_f32(b.part, SK, B.proj.unscale, B.proj.shift, x, imgs, 512, y, B.n2g, B.n2b, 1e-5f, b.ln, F, vs)) return -1; }
        if (run_linear(c, B.fc1, b.ln, nullptr, OFB_ACT_GELU, b.fc1)) return -1;
        if (run_linear(c, B.fc2, b.fc1, nullptr, OFB_ACT_NONE, b.kv, SK, b.part)) return -1;
        const bool last = i == nblk - 1;      // x = y + fc2(gelu(fc1)); then norm1 of the next block, or encoder_norm (eps 1e-6)
        { Prof pr(h, s, "splitk_finish_ln", 0.0, 4.0*((double)imgs*512*(SK + 3)));
        if (ofb_splitk_finish_ln_f32(b.part, SK, B.fc2.unscale, B.fc2.shift, y, imgs, 512, x,
                                     last ? h->enc_g : h->blk[i + 1].n1g, last ? h->enc_b : h->blk[i + 1].n1b,
                                     last ? 1e-6f : 1e-5f, last ? b.enc : b.ln, last ? OFB_FMT_F32 : F, vs)) return -1; }
      }
    } else {
    for (int i = 0; i < nblk; ++i) {   // Transformer_Block, model/blocks.py:84-88 (dbg_blocks = 6)
      Block& B = h->blk[i];
      { Prof pr(h, s, "layernorm", 0.0, 4.0*((double)imgs*1024));
      if (ofb_layernorm_f32(x, B.n1g, B.n1b, imgs, 512, 1e-5f, b.ln, F, F, vs)) return -1; }
      if (run_linear(c, B.qkv, b.ln, nullptr, OFB_ACT_NONE, b.kv)) return -1;     // (imgs,1536) = [q | k | v]
      { Prof pr(h, s, "attention", 0.0, 4.0*((double)imgs*2048));
      if (h->attn_tc && F == OFB_FMT_SPLIT16 && h->engine != OFB_ENGINE_SIMT) { if (attention_tc(b.kv, Bc, N, 4, b.att, s)) return -1; }
      else if (ofb_attention_qkv_f32(b.kv, Bc, N, 4, 128, b.att, F, vs)) return -1; }
      if (run_linear(c, B.proj, b.att, x, OFB_ACT_NONE, y)) return -1;          // y = x + proj(att)
      { Prof pr(h, s, "layernorm", 0.0, 4.0*((double)imgs*1024));
      if (ofb_layernorm_f32(y, B.n2g, B.n2b, imgs, 512, 1e-5f, b.ln, F, F, vs)) return -1; }
      if (run_linear(c, B.fc1, b.ln, nullptr, OFB_ACT_GELU, b.fc1)) return -1;
      if (run_linear(c, B.fc2, b.fc1, y, OFB_ACT_NONE, x)) return -1;           // x = y + fc2(gelu(fc1))
    }
    { Prof pr(h, s, "layernorm", 0.0, 4.0*((double)imgs*1024));
    if (ofb_layernorm_f32(x, h->enc_g, h->enc_b, imgs, 512, 1e-6f, b.enc, F, OFB_FMT_F32, vs)) return -1; }
    }
    }   // !no_transformer

    // decoder (:337-369); the token broadcast-add (:334-335) is fused into the first upsample
    { Prof pr(h, s, "upsample2x_c512", 0.0, 4.0*5.0*(double)imgs*(P / 32)*(P / 32)*512);
    if (ofb_upsample2x_f32(b.layer4, h->no_transformer ? nullptr : b.enc, imgs, P / 32, P / 32, 512, b.up0, F, vs)) return -1; }
    if (run_conv(c, h->conv["de_conv0_0"], b.up0, 512, nullptr, 0, P / 16, P / 16, 1, 1, nullptr, OFB_ACT_RELU, b.d00)) return -1;
    if (run_conv(c, h->conv["de_conv0_1"], b.d00, 256, b.layer3, 256, P / 16, P / 16, 1, 1, nullptr, OFB_ACT_RELU, b.d01)) return -1;
    { Prof pr(h, s, "upsample2x_c128", 0.0, 4.0*5.0*(double)imgs*(P / 16)*(P / 16)*128);
    if (ofb_upsample2x_f32(b.d01, nullptr, imgs, P / 16, P / 16, 128, b.up1, F, vs)) return -1; }
    if (run_conv(c, h->conv["de_conv1_0"], b.up1, 128, nullptr, 0, P / 8, P / 8, 1, 1, nullptr, OFB_ACT_RELU, b.d10)) return -1;
    if (run_conv(c, h->conv["de_conv1_1"], b.d10, 128, b.layer2, 128, P / 8, P / 8, 1, 1, nullptr, OFB_ACT_RELU, b.d11)) return -1;
    { Prof pr(h, s, "upsample2x_c64", 0.0, 4.0*5.0*(double)imgs*(P / 8)*(P / 8)*64);
    if (ofb_upsample2x_f32(b.d11, nullptr, imgs, P / 8, P / 8, 64, b.up2, F, vs)) return -1; }
    if (run_conv(c, h->conv["de_conv2_0"], b.up2, 64, nullptr, 0, p4, p4, 1, 1, nullptr, OFB_ACT_RELU, b.d20)) return -1;
    if (run_conv(c, h->conv["de_conv2_1"], b.d20, 64, layer1, 64, p4, p4, 1, 1, nullptr, OFB_ACT_RELU, b.d21)) return -1;
    { Prof pr(h, s, "upsample2x_c64", 0.0, 4.0*5.0*(double)imgs*(p4)*(p4)*64);
    if (ofb_upsample2x_f32(b.d21, nullptr, imgs, p4, p4, 64, b.up3, F, vs)) return -1; }
    if (run_conv(c, h->conv["de_conv3_0"], b.up3, 64, nullptr, 0, P / 2, P / 2, 1, 1, nullptr, OFB_ACT_RELU, b.d30)) return -1;
    if (run_conv(c, h->conv["de_conv3_1"], b.d30, 64, b.conv1, 64, P / 2, P / 2, 1, 1, nullptr, OFB_ACT_RELU, b.d31)) return -1;
    if (can_fuse_ups(h, h->conv["de_conv4_0"], imgs, P, P)) {
      if (run_conv(c, h->conv["de_conv4_0"], b.d31, 32, nullptr, 0, P, P, 1, 1, nullptr, OFB_ACT_RELU, b.d40, 1)) return -1;
    } else {
      { Prof pr(h, s, "upsample2x_c32", 0.0, 4.0*5.0*(double)imgs*(P / 2)*(P / 2)*32);
      if (ofb_upsample2x_f32(b.d31, nullptr, imgs, P / 2, P / 2, 32, b.up4, F, vs)) return -1; }
      if (run_conv(c, h->conv["de_conv4_0"], b.up4, 32, nullptr, 0, P, P, 1, 1, nullptr, OFB_ACT_RELU, b.d40)) return -1;
    }

    // heads + ERP merge (:371-380).  With the tensor-core heads and confidence the two patch maps are written as
    // one interleaved (pred*conf, conf) pair map: every tap of the blend below is then one 8-byte gather.
    const bool tc_heads = h->heads_tc && F == OFB_FMT_SPLIT16 && h->engine != OFB_ENGINE_SIMT && P == 128;
    pairs = tc_heads && confidence;
    { Prof pr(h, s, "heads", 0.0, 4.0*((double)imgs*P*P*34));
    if (tc_heads) {
      if (conv_tc_heads(b.d40, imgs, P, P, h->heads16.ws, h->heads16.unscale, h->pred_b, h->conf_b, confidence, b.pred, b.conf, pairs ? 1 : 0, s)) return -1;
    } else if (ofb_heads_f32(b.d40, imgs, P, P, h->pred_w, h->pred_b, h->conf_w, h->conf_b, confidence, b.pred, b.conf, F, vs)) return -1; }
    float* out = outs[it] + out_off;
    if (confidence) {
      { Prof pr(h, s, "blend_conf", 0.0, 4.0*((double)imgs*P*P*2 + (double)Bc*He*We));
      if (blend_conf_launch(b.pred, b.conf, pairs, Bc, N, P, P, g.blend_rowptr, g.blend_idx, g.blend_w, He, We, out, s)) return -1; }
    } else {
      { Prof pr(h, s, "pers2equi", 0.0, 4.0*((double)imgs*P*P + (double)Bc*He*We));
      if (ofb_pers2equi_f32(b.pred, Bc, 1, N, P, P, OFB_LAYOUT_FOLDED, g.blend_rowptr, g.blend_idx, g.blend_w, He, We, out, vs)) return -1; }
    }
  }
  reg(h, "patches", b.patches, imgs, P, P, 4, 0); reg(h, "conv1", b.conv1, imgs, P / 2, P / 2, 64);
  reg(h, "pool", b.pool, imgs, p4, p4, 64); reg(h, "layer1_pre", b.layer1_pre, imgs, p4, p4, 64);
  reg(h, "layer1", b.layer1, imgs, p4, p4, 64); reg(h, "layer2", b.layer2, imgs, P / 8, P / 8, 128);
  reg(h, "layer3", b.layer3, imgs, P / 16, P / 16, 256); reg(h, "layer4", b.layer4, imgs, P / 32, P / 32, 512);
  reg(h, "tokens", b.down, imgs, P / 32, P / 32, h->down.cout); reg(h, "encoded", b.enc, imgs, 1, 1, 512, 0);
  reg(h, "de_conv0_1", b.d01, imgs, P / 16, P / 16, 128); reg(h, "de_conv1_1", b.d11, imgs, P / 8, P / 8, 64);
  reg(h, "de_conv2_1", b.d21, imgs, p4, p4, 64); reg(h, "de_conv3_1", b.d31, imgs, P / 2, P / 2, 32);
  reg(h, "de_conv4_0", b.d40, imgs, P, P, 32); reg(h, "pred_patch", b.pred, imgs, P, P, 1, pairs ? 2 : 0);
  reg(h, "conf_patch", pairs ? b.pred : b.conf, imgs, P, P, 1, pairs ? 3 : 0);    // fmt 2 / 3: component 0 / 1 of a pair map
  if (h->check_range) {
    if (!h->range_dev) OFB_CUDA(cudaMalloc(&h->range_dev, kRangeSlots * 2 * sizeof(unsigned int)));
    OFB_CUDA(cudaMemsetAsync(h->range_dev, 0, kRangeSlots * 2 * sizeof(unsigned int), s));
    h->range_names.clear();
    static const char* const kChecked[] = {"conv1", "pool", "layer1_pre", "layer1", "layer2", "layer3", "layer4", "tokens",
                                           "encoded", "de_conv0_1", "de_conv1_1", "de_conv2_1", "de_conv3_1", "de_conv4_0"};
    for (const char* nm : kChecked) {
      const Act& a = h->acts[nm];
      if (!a.p || (a.fmt != OFB_FMT_F32 && a.fmt != OFB_FMT_SPLIT16)) continue;
      const size_t slot = h->range_names.size();
      if (range_launch(a.p, (size_t)a.n * a.h * a.w * a.c, a.fmt, h->range_dev + 2 * slot, s)) return -1;
      h->range_names.push_back(nm);
    }
  }
  return 0;
}

}  // namespace ofb

// ------------------------------------------------------------------ C ABI
extern "C" int ofb_version(void) { return OFB_VERSION; }
extern "C" const char* ofb_last_error(void) { return g_err.c_str(); }
extern "C" int64_t ofb_launch_count(int reset) {
  long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

extern "C" int ofb_set_device(int device) {
  OFB_CUDA(cudaSetDevice(device));
  return 0;
}

extern "C" int ofb_heads_tc_f16(const void* x_planes, int imgs, int h, int w, const void* wgt_split, float wgt_unscale,
                                float b_pred, float b_conf, int confidence, float* pred_out, float* conf_out,
                                void* stream) {
  return conv_tc_heads(x_planes, imgs, h, w, wgt_split, wgt_unscale, b_pred, b_conf, confidence, pred_out, conf_out, 0,
                       (cudaStream_t)stream);
}

extern "C" int ofb_heads_tc_pairs_f16(const void* x_planes, int imgs, int h, int w, const void* wgt_split, float wgt_unscale,
                                      float b_pred, float b_conf, float* pred_conf_out, void* stream) {
  return conv_tc_heads(x_planes, imgs, h, w, wgt_split, wgt_unscale, b_pred, b_conf, 1, pred_conf_out, pred_conf_out, 1,
                       (cudaStream_t)stream);
}

// The fused transformer stack on its own (tests): x (B, N, 512) float32 is updated in place, scratch holds the
// exchange buffers [scores (B,4,4,N,N) | attention output (B,N,512) | fc2 partial sums (B,16,N,512)]
extern "C" int ofb_token_stack_f32(ofb_handle* h, float* x, float* scratch, long long scratch_floats, float* enc_out, int B,
                                   int N, int nblk, int stop_phase, void* stream) {
  OFB_CHECK(h && h->has_weights && h->tok_ready, "token_stack: no weights loaded (or no split-half planes)");
  OFB_CHECK(nblk >= 1 && nblk <= 6, "token_stack: 1..6 blocks (got %d)", nblk);
  TokStack st = h->tok_stack;
  st.nblk = nblk;
  return token_stack_launch(st, x, scratch, (size_t)scratch_floats, enc_out, B, N, stop_phase, 0, false, (cudaStream_t)stream);
}
extern "C" int ofb_debug_token_stamps(long long* dev_buf) { token_stack_debug_stamps(dev_buf); return 0; }   // [24 phases][8] clocks, experiments
extern "C" long long ofb_token_stack_scratch_floats(int B, int N) { return (long long)token_stack_scratch_floats(B, N); }
extern "C" int ofb_token_stack_resident_groups(int N) { return token_stack_supported(N) ? token_stack_resident_groups(N) : 0; }

extern "C" int ofb_attention_tc_f16(const void* qkv_planes, int B, int N, int heads, int head_dim, void* out_planes,
                                    void* stream) {
  OFB_CHECK(head_dim == 128, "attention_tc: head_dim must be 128 (got %d)", head_dim);
  return attention_tc(qkv_planes, B, N, heads, out_planes, (cudaStream_t)stream);
}

extern "C" int ofb_stem_tc_f16(const void* patches, int n, int h, int w, const void* wgt_split, float wgt_unscale,
                               const float* scale, const float* shift, void* out, void* stream) {
  return stem_tc(patches, n, h, w, wgt_split, wgt_unscale, scale, shift, out, (cudaStream_t)stream);
}

extern "C" int ofb_conv_f32(const ofb_conv_desc* d, void* stream) { return conv_dispatch(d, (cudaStream_t)stream); }

extern "C" int ofb_create(int device, ofb_handle** out) {
  OFB_CHECK(out, "create: null out pointer");
  int count = 0;
  OFB_CUDA(cudaGetDeviceCount(&count));
  OFB_CHECK(device >= 0 && device < count, "create: device %d out of range (%d visible)", device, count);
  cudaDeviceProp prop;
  OFB_CUDA(cudaGetDeviceProperties(&prop, device));
  OFB_CHECK(prop.major == 10, "create: libofb is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  OFB_CUDA(cudaSetDevice(device));
  ofb_handle* h = new ofb_handle();
  h->device = device;
  *out = h;
  return 0;
}

extern "C" int ofb_destroy(ofb_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  free_weights(h);
  token_stack_release(&h->tok_stack);
  drop_profile_records(h);
  for (int l = 0; l < 2; ++l)
    if (h->lane[l].ws) cudaFree(h->lane[l].ws);
  if (h->range_dev) cudaFree(h->range_dev);
  if (h->chain_flags) cudaFree(h->chain_flags);
  if (h->aux) cudaStreamDestroy(h->aux);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
  return 0;
}

extern "C" int ofb_set_geometry(ofb_handle* h, const ofb_geometry* g) {
  OFB_CHECK(h && g, "set_geometry: null pointer");
  OFB_CHECK(g->grid_hi && g->grid_lo && g->pts && g->blend_rowptr && g->blend_idx && g->blend_w, "set_geometry: null table");
  // 128 is the reference's default; 256 is the network_test.py variant (down1 512 -> 8 over 8x8 positions); 64 works
  // the same way (down 512 -> 128 over 2x2).  The token width 32*(P/32)^2-independent 512 is checked against the
  // loaded down conv at forward time.
  OFB_CHECK(g->patch == 64 || g->patch == 128 || g->patch == 256, "set_geometry: patch must be 64, 128 or 256, got %d", g->patch);
  OFB_CHECK(g->n_patch > 0 && g->n_patch <= 64, "set_geometry: n_patch must be in [1,64], got %d", g->n_patch);
  OFB_CHECK(g->pts_c == 3 || g->pts_c == 5, "set_geometry: pts_c must be 3 or 5");
  h->geo = *g;
  h->has_geo = true;
  return 0;
}

extern "C" int ofb_load_weights(ofb_handle* h, const ofb_tensor_desc* tensors, int count, int single_stage) {
  OFB_CHECK(h && tensors && count > 0, "load_weights: bad arguments");
  OFB_CUDA(cudaSetDevice(h->device));
  free_weights(h);
  TMap m;
  for (int i = 0; i < count; ++i) {
    OFB_CHECK(tensors[i].name && tensors[i].data, "load_weights: tensor %d has null name/data", i);
    m[tensors[i].name] = &tensors[i];
  }
  h->single = single_stage != 0;
  if (load_all(h, m, h->single)) { free_weights(h); return -1; }
  OFB_CUDA(cudaStreamSynchronize(nullptr));      // all split-half conversions of the checkpoint
  {
    TokBlockDesc tb[6];
    bool ok = true;
    for (int i = 0; i < 6; ++i) {
      Block& B = h->blk[i];
      const ConvW* cw[4] = {&B.qkv, &B.proj, &B.fc1, &B.fc2};
      for (int j = 0; j < 4; ++j) {
        tb[i].lin[j].ws = cw[j]->ws; tb[i].lin[j].bias = cw[j]->shift; tb[i].lin[j].unscale = cw[j]->unscale;
        ok = ok && cw[j]->ws && (cw[j]->shift || j == 0);
      }
      tb[i].n1g = B.n1g; tb[i].n1b = B.n1b; tb[i].n2g = B.n2g; tb[i].n2b = B.n2b;
    }
    h->tok_ready = ok && token_stack_prepare(tb, 6, h->enc_g, h->enc_b, &h->tok_stack) == 0;
  }
  h->has_weights = true;
  return 0;
}

extern "C" int ofb_set_option(ofb_handle* h, const char* key, int value) {
  OFB_CHECK(h && key, "set_option: null pointer");
  if (!strcmp(key, "engine")) h->engine = value;
  else if (!strcmp(key, "lanes")) h->lanes = value >= 2 ? 2 : 1;
  else if (!strcmp(key, "chunk")) h->chunk = value;
  else if (!strcmp(key, "dedup")) h->dedup = value;
  else if (!strcmp(key, "fuse_ups")) h->fuse_ups = value;
  else if (!strcmp(key, "heads_tc")) h->heads_tc = value;
  else if (!strcmp(key, "attn_tc")) h->attn_tc = value;
  else if (!strcmp(key, "token_fused")) h->token_fused = value;
  else if (!strcmp(key, "conv_splitk")) h->conv_splitk = value;
  else if (!strcmp(key, "check_range")) h->check_range = value;
  else if (!strcmp(key, "chain")) h->chain = value;
  else if (!strcmp(key, "no_point_feat")) h->no_point_feat = value != 0;
  else if (!strcmp(key, "no_transformer")) h->no_transformer = value != 0;
  else if (!strcmp(key, "splitk")) h->splitk = value == 2 || value == 4 ? value : 1;
  else if (!strcmp(key, "wmc")) h->tc.wmc = value != 0;                 // weight multicast in the kh-reuse kernels
  else if (!strcmp(key, "khr_row64")) h->tc.khr_row64 = value != 0;
  else if (!strcmp(key, "khr_bw")) h->tc.khr_bw = value == 32 ? 32 : 16;    // tile width of the kh-reuse kernels
  else if (!strcmp(key, "dbg_blocks")) h->dbg_blocks = value < 0 ? 0 : (value > 6 ? 6 : value);   // timing experiments (wrong results)
  else if (!strcmp(key, "pdl")) h->tc.pdl = value != 0;       // programmatic dependent launch
  else if (!strcmp(key, "store128")) h->tc.store128 = value != 0;
  else if (!strcmp(key, "tc_debug")) h->tc.dbg = value;               // timing experiments only (wrong results)
  else if (!strcmp(key, "fill_div")) h->tc.fill_div = value > 0 ? value : 2;   // N-tile shrink threshold (experiments)
  else if (!strcmp(key, "direct32")) h->tc.direct32 = value != 0;        // (experiments)
  else if (!strcmp(key, "nstack")) { h->tc.nstack = value != 0; h->tc.nstack_ups = value >= 2; }   // tap-stacked MMAs: 1 heads, 2 also de_conv4_0
  else if (!strcmp(key, "cta2")) h->tc.cta2 = value != 0;               // cta_group::2 CTA pairs
  else if (!strcmp(key, "format")) {
    OFB_CHECK(value == OFB_FMT_F32 || value == OFB_FMT_SPLIT16, "set_option: format must be 0 (float32) or 1 (split-half)");
    h->fmt = value;
  }
  else OFB_CHECK(false, "set_option: unknown key '%s'", key);
  return 0;
}

extern "C" int ofb_forward_f32(ofb_handle* h, const float* rgb, int B, int iters, int confidence,
                               float* const* out_depth, void* stream) {
  OFB_CHECK(h && rgb && out_depth, "forward: null pointer");
  OFB_CHECK(h->has_geo, "forward: ofb_set_geometry has not been called");
  OFB_CHECK(h->has_weights, "forward: ofb_load_weights has not been called");
  OFB_CHECK(B > 0 && iters >= 1, "forward: bad batch/iters");
  OFB_CHECK(!h->single || iters == 1, "forward: the single-stage model runs exactly one iteration");
  OFB_CUDA(cudaSetDevice(h->device));
  TcOptScope opt_scope(&h->tc);
  for (int i = 0; i < iters; ++i) OFB_CHECK(out_depth[i], "forward: out_depth[%d] is null", i);
  int N = h->geo.n_patch;
  int chunk = h->chunk > 0 ? h->chunk : (576 / N > 0 ? 576 / N : 1);
  size_t plane = (size_t)h->geo.erp_h * h->geo.erp_w;
  cudaStream_t s = (cudaStream_t)stream;
  // panoramas [lo, hi) on lane `l`, in chunks
  auto run_lane = [&](int l, int lo, int hi, cudaStream_t ls) -> int {
    h->cur = l;
    for (int b0 = lo; b0 < hi; b0 += chunk) {
      int Bc = hi - b0 < chunk ? hi - b0 : chunk;
      if (forward_chunk(h, rgb + (size_t)b0 * 3 * plane, Bc, iters, confidence, out_depth, (size_t)b0 * plane, ls)) return -1;
    }
    return 0;
  };
  const bool two = h->lanes == 2 && B >= 4 && !h->profile;
  if (!two) {
    int rc = run_lane(0, 0, B, s);
    h->cur = 0;
    return rc;
  }
  if (!h->aux) {
    OFB_CUDA(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
    OFB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    OFB_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  }
  // both lanes' workspaces exist before anything is enqueued (an allocation must not interleave with a capture)
  const int half = (B + 1) / 2;
  for (int l = 0; l < 2; ++l) {
    h->cur = l;
    Buffers tmp;
    int nb = l == 0 ? half : B - half;
    if (ensure_workspace(h, (nb < chunk ? nb : chunk) * N, h->geo.patch, &tmp)) { h->cur = 0; return -1; }
  }
  OFB_CUDA(cudaEventRecord(h->ev_fork, s));
  OFB_CUDA(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
  h->tc.sm_share = 2;                      // persistent grids take half of the SMs each: the two lanes run side by side
  int rc = run_lane(0, 0, half, s);
  if (!rc) rc = run_lane(1, half, B, h->aux);
  h->tc.sm_share = 1;
  h->cur = 0;
  OFB_CUDA(cudaEventRecord(h->ev_join, h->aux));
  OFB_CUDA(cudaStreamWaitEvent(s, h->ev_join, 0));
  return rc;
}

// timing experiments: TcParams::dbg of tcgen05 convs launched directly through ofb_conv_f32 (tools/probe_mid.py)
extern "C" int ofb_debug_set(int v) { conv_tc_default_debug(v); return 0; }
extern "C" int ofb_debug_nstack(int v) { conv_tc_default_nstack(v); return 0; }   // 0 off, 1 heads, 2 also fused-upsample conv

// timing experiments: copies the clock stamps recorded with tc_debug & 16 (512 x 8 int64) to the host
extern "C" int ofb_debug_stamps(long long* host_dst) {
  OFB_CUDA(cudaDeviceSynchronize());
  OFB_CUDA(cudaMemcpy(host_dst, conv_tc_debug_buffer(), 512 * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

// timing experiments: %globaltimer stamps of CTA 0 of every tcgen05 conv launch since the last call ("tc_debug" & 256):
// copies up to max_slots x 8 int64 to host_dst, resets the slot counter, returns the number of launches recorded
extern "C" int ofb_debug_timeline(long long* host_dst, int max_slots) {
  OFB_CUDA(cudaDeviceSynchronize());
  const int slots = conv_tc_timeline_slots();
  long long* buf = conv_tc_debug_buffer();
  long long count = 0;
  OFB_CUDA(cudaMemcpy(&count, buf + (size_t)slots * 8, sizeof(long long), cudaMemcpyDeviceToHost));
  int n = (int)(count < slots ? count : slots);
  if (n > max_slots) n = max_slots;
  if (host_dst && n > 0) OFB_CUDA(cudaMemcpy(host_dst, buf, (size_t)n * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  OFB_CUDA(cudaMemset(buf, 0, ((size_t)slots * 8 + 8) * sizeof(long long)));
  return n;
}

// timing experiments: the whole stamp buffer (1024 x 8 int64), no reset (tools/chain_timeline.py)
extern "C" int ofb_debug_timeline_raw(long long* host_dst) {
  OFB_CUDA(cudaDeviceSynchronize());
  OFB_CUDA(cudaMemcpy(host_dst, conv_tc_debug_buffer(), (size_t)conv_tc_timeline_slots() * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int ofb_profile_enable(ofb_handle* h, int on) {
  OFB_CHECK(h, "profile_enable: null handle");
  h->profile = on != 0;
  if (!on) drop_profile_records(h);     // records nobody asked a report for
  return 0;
}

// Option "check_range": report of the last forward (last chunk).  Synchronises the device.  Writes one line per checked
// activation, "name max_abs nonfinite\n", into buf; returns the number of activations whose values left the range the
// storage format represents (non-finite, or |x| >= 65504 in the split-half format), or negative on error.
extern "C" int ofb_range_report(ofb_handle* h, char* buf, int capacity) {
  OFB_CHECK(h && buf && capacity > 0, "range_report: bad arguments");
  buf[0] = 0;
  if (!h->range_dev || h->range_names.empty()) return 0;
  OFB_CUDA(cudaDeviceSynchronize());
  unsigned int host[kRangeSlots * 2];
  OFB_CUDA(cudaMemcpy(host, h->range_dev, sizeof(host), cudaMemcpyDeviceToHost));
  int off = 0, bad = 0;
  for (size_t i = 0; i < h->range_names.size(); ++i) {
    float m;
    memcpy(&m, &host[2 * i], sizeof(float));
    const Act& a = h->acts[h->range_names[i]];
    if (host[2 * i + 1] || (a.fmt == OFB_FMT_SPLIT16 && m >= 65504.f)) ++bad;
    int w = snprintf(buf + off, capacity - off, "%s %.6e %u\n", h->range_names[i].c_str(), m, host[2 * i + 1]);
    if (w < 0 || w >= capacity - off) break;
    off += w;
  }
  return bad;
}

extern "C" long long ofb_workspace_generation(ofb_handle* h) { return h ? h->ws_generation : -1; }
extern "C" const char* ofb_last_conv_variant(void) { return conv_tc_last_variant(); }

extern "C" int ofb_profile_report(ofb_handle* h, char* buf, int capacity) {
  OFB_CHECK(h && buf && capacity > 0, "profile_report: bad arguments");
  struct Agg { int n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (auto& r : h->recs) {
    OFB_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    OFB_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    if (!agg.count(r.name)) order.push_back(r.name);
    Agg& a = agg[r.name];
    a.n++; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
  }
  h->recs.clear();
  int off = 0;
  for (auto& k : order) {
    Agg& a = agg[k];
    int w = snprintf(buf + off, capacity - off, "%s %d %.6f %.6e %.6e\n", k.c_str(), a.n, a.ms, a.flops, a.bytes);
    if (w < 0 || w >= capacity - off) break;
    off += w;
  }
  return off;
}

extern "C" int64_t ofb_get_activation(ofb_handle* h, const char* name, float* dst, int64_t capacity, int dims[4],
                                      void* stream) {
  OFB_CHECK(h && name, "get_activation: null pointer");
  auto it = h->acts.find(name);
  OFB_CHECK(it != h->acts.end(), "get_activation: unknown activation '%s'", name);
  const Act& a = it->second;
  int64_t n = (int64_t)a.n * a.h * a.w * a.c;
  if (dims) { dims[0] = a.n; dims[1] = a.h; dims[2] = a.w; dims[3] = a.c; }
  if (dst) {
    OFB_CHECK(capacity >= n, "get_activation: capacity %lld < %lld", (long long)capacity, (long long)n);
    if (a.fmt == OFB_FMT_SPLIT16) {
      if (ofb_merge_f16(a.p, (size_t)n, dst, stream)) return -1;
    } else if (a.fmt == 2 || a.fmt == 3) {
      if (deinterleave_launch(a.p, (size_t)n, a.fmt - 2, dst, (cudaStream_t)stream)) return -1;
    } else {
      OFB_CUDA(cudaMemcpyAsync(dst, a.p, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
  }
  return n;
}
