/*
 * libofb - B200 (sm_100a) kernels for the OmniFusion tangent-patch inference path.
 *
 * C ABI: plain pointers and sizes, no torch types.  Every tensor pointer is a
 * DEVICE pointer unless the parameter says "host"; the caller owns every buffer
 * it passes in; all calls are asynchronous and ordered on the cudaStream_t passed
 * as `stream` (a void* so the header needs no CUDA include); nothing here calls
 * cudaDeviceSynchronize.  Return value: 0 = OK, negative = error, message in
 * ofb_last_error() (thread-local).  A handle is bound to one device and is not
 * concurrently callable from two threads.
 *
 * Each entry point cites the reference interface it replaces (paths relative to
 * the OmniFusion repository).
 *
 * Activation layout inside the library ("folded NHWC"): image index b*N+n
 * (panorama b, patch n), then H, W, C with C contiguous.
 */
#ifndef OFB_H_
#define OFB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFB_VERSION 103

typedef struct ofb_handle ofb_handle;

int ofb_version(void);
const char* ofb_last_error(void);

/* Selects the CUDA device the calling thread's subsequent operator calls launch on
 * (the library links its own CUDA runtime, whose current device is independent of the
 * host framework's).  Engine calls (ofb_forward_f32 ...) use the handle's device. */
int ofb_set_device(int device);

/* ---------------------------------------------------------------- resamplers */

/* Layout selector for patch tensors. */
enum {
  OFB_LAYOUT_REF = 0,    /* reference layout (B,C,Ph,Pw,N), N innermost            */
  OFB_LAYOUT_FOLDED = 1, /* folded NHWC (B*N,Ph,Pw,Cpad), Cpad = C rounded up to 4
                            when C == 3, else C                                      */
  OFB_LAYOUT_STEM16 = 2  /* equi2pers only, C == 3: split-half planes of (B*N,Ph,Pw+8,4), rows
                            padded by 4 pixels on each side (pad left untouched: the caller
                            zeroes the buffer once) - the input format of ofb_stem_tc_f16  */
};

/* equi2pers sampling: equi_pers/equi2pers_v3.py:106-113 (F.grid_sample bilinear /
 * border / align_corners=True on the gnomonic grid, then F.unfold into patches).
 * erp (B,C,He,We); grid (N,Ph,Pw,2) = the reference's normalised [lon,lat] grid,
 * built once on the host (equi2pers_v3.py:86-104); out per `layout`. */
int ofb_equi2pers_f32(const float* erp, int B, int C, int He, int We,
                      const float* grid, int N, int Ph, int Pw,
                      float* out, int layout, void* stream);

/* The integer neighbour indices the sampler above uses (x0 = floor(ix), y0 =
 * floor(iy)); exposed so tests can check them bit-exactly.  x0,y0: (N,Ph,Pw) int32. */
int ofb_equi2pers_taps(const float* grid, int N, int Ph, int Pw, int He, int We,
                       int32_t* x0, int32_t* y0, void* stream);

/* pers2equi apply: equi_pers/pers2equi_v3.py:169-196 (4-tap gather per covering
 * patch, thresholded L1-normalised weights, sum over taps and patches).
 * The dense (N,He,We) tap table of pers2equi_v3.py:109-152 is passed as CSR over
 * ERP pixels holding only the entries whose normalised weight vector is non-zero:
 *   rowptr (He*We+1) int32; idx (nnz) packed n<<24 | y0<<16 | x0<<8 | dy<<1 | dx
 *   with y1=y0+dy, x1=x0+dx (the reference's clamped x1,y1); w (nnz,4) float32 =
 *   normalised [wa,wb,wc,wd] for taps (y0,x0),(y1,x0),(y0,x1),(y1,x1).
 * pers per `layout` (OFB_LAYOUT_FOLDED here means (B*N,Ph,Pw,C)); out (B,C,He,We). */
int ofb_pers2equi_f32(const float* pers, int B, int C, int N, int Ph, int Pw, int layout,
                      const int32_t* rowptr, const uint32_t* idx, const float* w,
                      int He, int We, float* out, void* stream);

/* Backward of the two resamplers (the training direction of SURVEY section 8f-4: losses back-propagate through
 * equi2pers and pers2equi, e.g. train_erp_depth_iterative.py).  Both are linear in their image argument with
 * input-independent taps, so the backward is the transposed gather: a scatter-add (atomicAdd) of the incoming gradient
 * with the forward's weights.  The destination must be zeroed by the caller; reference layouts:
 * grad_pers (B,C,Ph,Pw,N), grad_erp (B,C,He,We). */
int ofb_equi2pers_backward_f32(const float* grad_pers, int B, int C, int He, int We, const float* grid, int N, int Ph,
                               int Pw, float* grad_erp, void* stream);
int ofb_pers2equi_backward_f32(const float* grad_erp, int B, int C, int N, int Ph, int Pw, const int32_t* rowptr,
                               const uint32_t* idx, const float* w, int He, int We, float* grad_pers, void* stream);

/* Confidence merge: model/spherical_model_iterative.py:372-378.
 * pred_w (B*N,Ph,Pw) = relu(pred)*sigmoid(weight), conf (B*N,Ph,Pw) = sigmoid(weight);
 * out (B,1,He,We) = blend(pred_w) / (blend(conf) + 1e-8*[blend(conf) <= 1e-8]). */
int ofb_blend_conf_f32(const float* pred_w, const float* conf, int B, int N, int Ph, int Pw,
                       const int32_t* rowptr, const uint32_t* idx, const float* w,
                       int He, int We, float* out, void* stream);

/* The same with the two patch maps interleaved per pixel: pred_conf (B*N,Ph,Pw,2) = (relu(pred)*sigmoid(weight),
 * sigmoid(weight)), the layout ofb_heads_tc_pairs_f16 writes; every tap is one 8-byte gather.  8-byte aligned. */
int ofb_blend_conf_pairs_f32(const float* pred_conf, int B, int N, int Ph, int Pw, const int32_t* rowptr,
                             const uint32_t* idx, const float* w, int He, int We, float* out, void* stream);

/* ------------------------------------------------------------ network kernels */

enum { OFB_ACT_NONE = 0, OFB_ACT_RELU = 1, OFB_ACT_GELU = 2 };
enum { OFB_ENGINE_AUTO = 0, OFB_ENGINE_SIMT = 1, OFB_ENGINE_TC = 2 };

/* Activation storage formats.  OFB_FMT_F32: plain float32.  OFB_FMT_SPLIT16: "split-half
 * planes" - value = float(hi) + float(lo), stored as one buffer of 2*numel halves, the hi
 * plane followed by the lo plane (same 4 bytes/element as float32, ~22 mantissa bits).  It
 * is what lets the tcgen05 engine reach fp32-level accuracy with three kind::f16 MMAs
 * (hi*hi + lo*hi + hi*lo) instead of one lossy TF32/BF16 MMA. */
enum { OFB_FMT_F32 = 0, OFB_FMT_SPLIT16 = 1 };

/* One 2-D convolution over folded NHWC, replacing a Conv3d((k,k,1)) + BatchNorm3d
 * (+ residual) (+ ReLU) group of model/spherical_model_iterative.py:322-369 or an
 * nn.Linear of model/blocks.py (n = rows, h = w = 1, k = 1).
 *   y = act(conv(cat(in0,in1)) * scale + shift + residual)
 * in0 (n,h,w,c0), in1 (n,h,w,c1) or NULL; wgt (cout, k, k, c0+c1) ("OHWI") float32;
 * scale/shift (cout) or NULL (= 1 / 0); residual (n,oh,ow,cout) or NULL.
 * in_fmt applies to in0/in1, out_fmt to out/residual.  The tcgen05 engine needs
 * in_fmt == out_fmt; with OFB_FMT_SPLIT16 it reads the weights from wgt_split (split-half
 * planes of wgt * 2^e made by ofb_split_f16, wgt_unscale = 2^-e); with OFB_FMT_F32 it runs
 * one kind::tf32 MMA on wgt (TF32 accuracy only - not used by the engine's default path).
 * ups2x != 0 fuses the decoder's F.interpolate(scale 2, bilinear, align_corners=False) of
 * spherical_model_iterative.py:367 into the conv: in0 is then the LOW-resolution tensor
 * (n,h/2,w/2,c0) and h,w stay the conv's input size; the upsampled tensor is never materialised
 * (tcgen05 engine, split-half format, 32 -> 32 channels, 3x3 stride 1, w % 32 == 0, h % 4 == 0). */
typedef struct {
  const void* in0; const void* in1; int c0, c1;
  int n, h, w;
  const float* wgt; int k, stride, pad, cout;
  const float* scale; const float* shift; const void* residual;
  int act;
  void* out;
  int engine;
  int in_fmt, out_fmt;
  const void* wgt_split; float wgt_unscale;
  int ups2x;
  int ksplit;        /* > 1 (tcgen05 engine, split-half format; 1x1 layers and 3x3 stride-1 layers with cout > 64): the
                        channel chunks of K are split over `ksplit` CTAs per tile; the raw float32 partial sums go to
                        `partial` (ksplit, n*h*w, cout) and scale / shift / residual / act / out are NOT applied -
                        ofb_splitk_finish_ln_f32 (token linears) / ofb_splitk_finish_conv_f16 (convs) finish the layer */
  float* partial;
} ofb_conv_desc;
int ofb_conv_f32(const ofb_conv_desc* d, void* stream);

/* float32 -> split-half planes of (src * mul), and back.  dst/src planes: 2*n halves. */
int ofb_split_f16(const float* src, size_t n, float mul, void* dst_planes, void* stream);
int ofb_merge_f16(const void* src_planes, size_t n, float* dst, void* stream);

/* The network kernels below take OFB_FMT_* selectors for their activation tensors (void*);
 * weights, biases, tables and the head outputs are always float32. */

/* Stem: Conv3d(3->64, 7x7 s2 p3) + BN + ReLU, spherical_model_iterative.py:322.
 * in (n,h,w,4) (4th channel ignored), wgt (7,7,4,64) = [kh][kw][cin padded to 4][cout], 16-byte
 * aligned (fetched with one bulk copy per CTA), out (n,h/2,w/2,64). */
int ofb_stem_f32(const float* in, int n, int h, int w, const float* wgt,
                 const float* scale, const float* shift, void* out, int out_fmt, void* stream);

/* The same stem on the tcgen05 engine.  patches: split-half planes of (n,h,w+8,4) - 4 channels per
 * pixel (4th = 0), every row padded with 4 zero pixels left and right (written by
 * ofb_equi2pers_f32 with OFB_LAYOUT_STEM16; the pad must be zero).  wgt_split: split-half planes
 * of (64,7,8,4) = [cout][kh][kw' = kw+1 (kw' = 0 is zero)][cin padded to 4] times 2^e made by
 * ofb_split_f16, wgt_unscale = 2^-e.  out: split-half (n,h/2,w/2,64). */
int ofb_stem_tc_f16(const void* patches, int n, int h, int w, const void* wgt_split, float wgt_unscale,
                    const float* scale, const float* shift, void* out, void* stream);

/* F.max_pool3d((3,3,1), s(2,2,1), p(1,1,0)), spherical_model_iterative.py:323. */
int ofb_maxpool3x3s2_f32(const void* in, int n, int h, int w, int c, void* out, int fmt, void* stream);

/* F.interpolate(bilinear, align_corners=False) x2, spherical_model_iterative.py:338-367.
 * If img_bias (n,c) is given it is added to every source pixel first (the token
 * broadcast-add of :334-335). in (n,h,w,c) -> out (n,2h,2w,c). */
int ofb_upsample2x_f32(const void* in, const float* img_bias, int n, int h, int w, int c,
                       void* out, int fmt, void* stream);

/* Point embedding: mlp_points1/2, spherical_model_iterative.py:290-305,319-320,387-393.
 * pts (N,cin,p,p) NCHW (the xyz table, or [cx,cy,1,cx,cy] for the single-stage model);
 * depth (imgs,p,p) or NULL scales the first 3 channels per image; base (imgs,p,p,64)
 * is added to the result (layer1 + point_feat, :325). out (imgs,p,p,64). */
int ofb_point_embed_f32(const float* pts, int N, int cin, int p, const float* depth, int imgs,
                        const float* w1, const float* s1, const float* t1,
                        const float* w2, const float* s2, const float* t2,
                        const void* base, void* out, int fmt, void* stream);

/* Token packing: spherical_model_iterative.py:330-331 + pos_emb add (:244).
 * down (imgs,4,4,32) -> tokens (imgs,512) with index c*16+i*4+j, + pos_emb[n]. */
int ofb_token_pack_f32(const void* down, const float* pos_emb, int imgs, int N,
                       void* tokens, int fmt, void* stream);

/* nn.LayerNorm over the last dim (model/blocks.py:76,81; encoder_norm eps 1e-6). */
int ofb_layernorm_f32(const void* x, const float* gamma, const float* beta, int rows, int dim,
                      float eps, void* y, int in_fmt, int out_fmt, void* stream);

/* Finish of a split-K linear with 512 outputs + the LayerNorm that follows it in a Transformer_Block
 * (model/blocks.py:84-88): x_out = residual + bias + wscale * sum_s partial[s] (split-half planes, rows x 512),
 * ln_out = LayerNorm(x_out; gamma, beta, eps) in format ln_fmt.  residual / x_out are split-half planes. */
int ofb_splitk_finish_ln_f32(const float* partial, int ksplit, float wscale, const float* bias,
                             const void* residual, int rows, int dim, void* x_out, const float* gamma,
                             const float* beta, float eps, void* ln_out, int ln_fmt, void* stream);

/* Finish of a split-K conv layer (torchvision BasicBlock convs of layer4, spherical_model_iterative.py:328 via
 * resnet.layer4): out = act((sum_s partial[s]) * (scale * wscale) + shift + residual), partial (ksplit, pixels, cout)
 * float32 in slice order, residual / out split-half planes of (pixels, cout), act NONE or RELU.  The engine uses it
 * for the 512 -> 512 convs when a chunk holds at most nine panoramas (option "conv_splitk"). */
int ofb_splitk_finish_conv_f16(const float* partial, int ksplit, long long pixels, int cout, const float* scale,
                               const float* shift, float wscale, const void* residual_planes, int act, void* out_planes,
                               void* stream);

/* Attention core, model/blocks.py:50-62: q (rows,512), kv (rows,1024) [k | v],
 * rows = B*N, heads of 128; softmax(q k^T / sqrt(128)) v -> out (rows,512). */
int ofb_attention_f32(const void* q, const void* kv, int B, int N, int heads, int head_dim,
                      void* out, int fmt, void* stream);

/* Same with q, k, v produced by one fused linear: qkv (rows,1536) = [q | k | v] per row. */
int ofb_attention_qkv_f32(const void* qkv, int B, int N, int heads, int head_dim,
                          void* out, int fmt, void* stream);

/* The same attention core on the tensor pipe (csrc/conv_tc.cu: attention_tc_kernel): qkv_planes = split-half planes
 * of (B*N, 3*heads*128) = [q | k | v] per row, out_planes = split-half planes of (B*N, heads*128).  One CTA per head
 * and tile of 128 / N whole panoramas: QK^T and PV as tcgen05 MMAs (three split-half products, fp32 accumulation in
 * TMEM), softmax per query row out of TMEM.  N <= 64, head_dim == 128. */
int ofb_attention_tc_f16(const void* qkv_planes, int B, int N, int heads, int head_dim, void* out_planes, void* stream);

/* The whole transformer stack of the patch network (model/blocks.py:50-88, Transformer_Block x nblk, followed by
 * transformer.encoder_norm; spherical_model_iterative.py:332-333) as ONE launch (csrc/token_tc.cu): a group of 16
 * CTAs per panorama, weights streamed by TMA, all four linears as tcgen05 MMAs with the weight rows as M and the
 * tokens as N, exchanges inside the group through L2 and per-panorama arrival counters.  Uses the weights loaded into `h`.  x (B*N, 512) float32 is
 * the residual stream, updated in place; enc_out (B*N, 512) float32 receives encoder_norm(x) when stop_phase == 0.
 * stop_phase = k > 0 stops after the first k GEMM phases (4 per block: qkv + attention, proj, fc1, fc2) - tests read
 * the exchange buffers in `scratch` = [partial scores (B,4 heads,4 quarters,N,N) | attention output (B*N,512) |
 * fc2 partial sums (B,16,N,512)], ofb_token_stack_scratch_floats(B, N) floats.  N <= 48.  The engine option
 * "token_fused" (default 1) selects this path inside ofb_forward_f32 for the split-half format. */
int ofb_token_stack_f32(ofb_handle* h, float* x, float* scratch, long long scratch_floats, float* enc_out, int B, int N,
                        int nblk, int stop_phase, void* stream);
long long ofb_token_stack_scratch_floats(int B, int N);
/* Panoramas (groups of 16 CTAs, one CTA per SM) of that kernel the device runs at once: 9 on a 148-SM B200. */
int ofb_token_stack_resident_groups(int N);
/* experiments (tools/probe_token.py): device buffer of 24 x 8 int64 that thread 0 of CTA 0 of the following
 * token-stack launches fills with clock64 stamps per GEMM phase; NULL switches it off */
int ofb_debug_token_stamps(long long* dev_buf);

/* The same heads on the tensor pipe (split-half format only): x_planes = split-half planes of (imgs,h,128,32),
 * wgt_split = split-half planes of the (16,3,3,32) filter bank [pred; weight_pred; 14 zero rows] scaled by
 * 1 / wgt_unscale (ofb_split_f16).  Rolling-row tcgen05 kernel, see csrc/conv_tc.cu (conv_tc_heads). */
int ofb_heads_tc_f16(const void* x_planes, int imgs, int h, int w, const void* wgt_split, float wgt_unscale,
                     float b_pred, float b_conf, int confidence, float* pred_out, float* conf_out, void* stream);

/* Same (confidence on) with both outputs interleaved per pixel: pred_conf_out (imgs,h,w,2) = (pred*conf, conf). */
int ofb_heads_tc_pairs_f16(const void* x_planes, int imgs, int h, int w, const void* wgt_split, float wgt_unscale,
                           float b_pred, float b_conf, float* pred_conf_out, void* stream);

/* Heads: pred / weight_pred 3x3 convs + relu / sigmoid / product,
 * spherical_model_iterative.py:371-374.  x (imgs,h,w,32); w_pred,w_conf (3,3,32);
 * pred_out = relu(pred) * (confidence ? sigmoid(conf) : 1); conf_out = sigmoid(conf)
 * (written only when confidence != 0). */
int ofb_heads_f32(const void* x, int imgs, int h, int w,
                  const float* w_pred, float b_pred, const float* w_conf, float b_conf,
                  int confidence, float* pred_out, float* conf_out, int in_fmt, void* stream);

/* Loader-side input conversion, dataset_loader_stanford.py:52,79 (rgb.astype(float32) / 255, HWC -> CHW; channel
 * order untouched): src (B,H,W,C) uint8 device -> dst (B,C,H,W) float32, bit-identical to the numpy expression. */
int ofb_u8hwc_to_f32chw(const uint8_t* src, int B, int H, int W, int C, float* dst, void* stream);

/* cv2.resize(img, (W/factor, H/factor), interpolation=cv2.INTER_AREA) of the decoded uint8 panorama for integer
 * scale factors (dataset_loader_stanford.py:92-96): src (B,H,W,C) -> dst (B,H/factor,W/factor,C), both uint8 device;
 * bit-identical to OpenCV (mean of the factor x factor block, rounded to nearest even). */
int ofb_area_resize_u8(const uint8_t* src, int B, int H, int W, int C, int factor, uint8_t* dst, void* stream);

/* Lower median (torch.median) of x[mask != 0] by four radix-select passes; no host synchronisation.  state: device
 * scratch of >= 1027 uint32 (zeroed by the call); out: 1 float (NaN when the mask is empty).  x 16-byte aligned. */
int ofb_masked_median_f32(const float* x, const uint8_t* mask, size_t n, void* state, float* out, void* stream);

/* Median scaling of test.py:161-162: out3 = [median(gt[mask]) / median(pred[mask]), median(gt), median(pred)]. */
int ofb_median_scale_f32(const float* pred, const float* gt, const uint8_t* mask, size_t n, void* state,
                         float* out3, void* stream);

/* Point cloud of test.py:205-218: pts (B,He*We,3) = rays (He*We,3) * depth (B,1,He,We); depths above max_depth are
 * zeroed first when max_depth > 0 (test.py:208).  rays: util.py:159-174 coords2uv / uv2xyz, built on the host. */
int ofb_depth_to_points_f32(const float* depth, const float* rays, int B, int He, int We, float max_depth,
                            float* pts, void* stream);

/* Training-side losses (supervision/direct.py:3-27; train_erp_depth_iterative.py:271): mode 1 = reverse Huber
 * (BerHu) loss calculate_berhu_loss(pred, gt, mask, weights), mode 0 = calculate_l1_loss(pred, gt, mask) (weights
 * NULL).  pred / gt / mask / weights: bs x per_sample float32 (mask as mask.float()).  c = max|gt - pred| / 5 over the
 * whole UNMASKED batch (a constant for the gradient), loss = mean_b(sum(loss * mask * weights)_b / sum(mask)_b);
 * degenerate inputs give NaN exactly where the reference does (c == 0, a sample without valid pixels).
 * work: ofb_loss_work_bytes(bs) bytes of device scratch; stats: 1 + bs floats (c, per-sample counts) kept for
 * ofb_depth_loss_backward_f32, which writes d loss / d pred * (*grad_out) into grad_pred. */
long long ofb_loss_work_bytes(int bs);
int ofb_depth_loss_f32(const float* pred, const float* gt, const float* mask, const float* weights, int bs,
                       long long per_sample, int mode, void* work, float* stats, float* loss, void* stream);
int ofb_depth_loss_backward_f32(const float* pred, const float* gt, const float* mask, const float* weights, int bs,
                                long long per_sample, int mode, const float* stats, const float* grad_out,
                                float* grad_pred, void* stream);

/* Abs-Rel partial sums, metrics.py:7-9: out[0] += sum(|p*scale-g|/g over mask), out[1] += count.
 * `out` must be zeroed by the caller. */
int ofb_absrel_partial(const float* pred, const float* gt, const uint8_t* mask, size_t n,
                       float scale, double* out, void* stream);

/* All seven evaluation metrics of metrics.py:7-26 as partial sums (test.py:151-177): out[9] +=
 * [sum |p-g|/g, sum (p-g)^2/g, sum (p-g)^2, sum (ln p - ln g)^2, n_log, n(delta<1.25),
 *  n(delta<1.25^2), n(delta<1.25^3), n] over mask, with p = pred*scale.  `out` zeroed by the caller. */
int ofb_depth_metrics_partial(const float* pred, const float* gt, const uint8_t* mask, size_t n,
                              float scale, double* out, void* stream);
/* Same with the scale read from device memory (*scale_dev, e.g. out3[0] of ofb_median_scale_f32): the whole
 * evaluation step stays on the stream without a device-to-host round trip. */
int ofb_depth_metrics_partial_ds(const float* pred, const float* gt, const uint8_t* mask, size_t n,
                                 const float* scale_dev, double* out, void* stream);

/* ------------------------------------------------------------------- engine */

int ofb_create(int device, ofb_handle** out);
int ofb_destroy(ofb_handle* h);

/* Geometry tables (device pointers, kept alive by the caller until the next
 * ofb_set_geometry / ofb_destroy).  grid_hi (N,P,P,2), grid_lo (N,p,p,2) with p=P/4,
 * pts (N,pts_c,p,p): xyz (pts_c=3, iterative) or [cx,cy,1,cx,cy] (pts_c=5, single-stage). */
typedef struct {
  int n_patch, patch, erp_h, erp_w;
  const float* grid_hi; const float* grid_lo; const float* pts; int pts_c;
  const int32_t* blend_rowptr; const uint32_t* blend_idx; const float* blend_w;
} ofb_geometry;
int ofb_set_geometry(ofb_handle* h, const ofb_geometry* g);

/* Weights in the reference state_dict layout (HOST pointers, float32), by name:
 * model/spherical_model_iterative.py:254-305 / model/spherical_model.py:191-235.
 * Repacked (OHWI, BN folded to scale/shift) and copied to the device. */
typedef struct {
  const char* name; const float* data; int ndim; int64_t shape[5];
} ofb_tensor_desc;
int ofb_load_weights(ofb_handle* h, const ofb_tensor_desc* tensors, int count, int single_stage);

/* spherical_fusion.forward: model/spherical_model_iterative.py:308-456 (single_stage=0,
 * returns `iters` maps) or model/spherical_model.py:238-314 (iters must be 1).
 * rgb (B,3,He,We); out_depth: HOST array of `iters` device pointers, each (B,1,He,We). */
int ofb_forward_f32(ofb_handle* h, const float* rgb, int B, int iters, int confidence,
                    float* const* out_depth, void* stream);

/* Bumped every time the engine (re)allocates its workspace arena (a forward with more images than any before).  A
 * CUDA graph captured around ofb_forward_f32 replays launches into the arena it was captured with: compare this value
 * with the one read at capture time and re-capture when it changed. */
long long ofb_workspace_generation(ofb_handle* h);

/* Engine knobs (key, value): "engine" conv engine (OFB_ENGINE_*), "chunk" panoramas per internal chunk
 * (0 = auto), "dedup" reuse of the iteration-invariant stem/layer1 across iterations (default 1), "format"
 * activation storage (OFB_FMT_*), "fuse_ups" fold the last decoder upsample into de_conv4_0 (default 1),
 * "check_range" see ofb_range_report, "chain" image-stationary layer chains: 0 off, 1 (default) for the encoder stages whose
 * dependencies stay inside a CTA pair (layer2), 2 also stages that hand images over between clusters (layer3), "heads_tc" heads on the tensor pipe (default 1), "attn_tc" attention core on the tensor pipe (default 1), "lanes" 2 = two concurrent half-batches on two streams
 * (default 1), "cta2" / "pdl" / "store128" / "fill_div" / "direct32" / "khr_bw" / "khr_row64" / "wmc" (weight multicast between the two CTAs of a cluster in
 * the kh-reuse kernels, off) / "nstack" (tap-stacked MMAs of the
 * rolling-row kernels: 1 = heads (default), 2 = also the fused-upsample conv) tcgen05 launch variants (per handle);
 * "tc_debug" / "dbg_blocks" switch parts of the pipeline OFF for timing experiments (results are wrong). */
int ofb_set_option(ofb_handle* h, const char* key, int value);

/* Numeric range of the network's activations.  The split-half storage format (OFB_FMT_SPLIT16) holds value = fp16 hi +
 * fp16 lo: |x| above 65504 overflows to inf and |x| below ~6e-5 loses low-order bits to fp16 subnormals (absolute floor
 * ~6e-8); the reference is fp32 and has neither limit.  With option "check_range" = 1 every forward also reduces the main
 * activations (stem, pool, layer1-4, tokens, encoder output, decoder stages) to max |x| and a non-finite count;
 * ofb_range_report synchronises, writes "name max_abs nonfinite" lines and returns how many of them left the
 * representable range (0 = fine).  Checkpoints whose activations do not fit should run with "format" = OFB_FMT_F32. */
int ofb_range_report(ofb_handle* h, char* buf, int capacity);

/* Copies a named intermediate of the last forward (last chunk, last iteration) into
 * dst (device, `capacity` floats); returns the element count, or negative.  Names:
 * patches, conv1, pool, layer1_pre, layer1, layer2, layer3, layer4, tokens, encoded,
 * de_conv0_1, de_conv1_1, de_conv2_1, de_conv3_1, de_conv4_0, pred_patch, conf_patch. */
int64_t ofb_get_activation(ofb_handle* h, const char* name, float* dst, int64_t capacity,
                           int dims[4], void* stream);

/* Per-launch timing: while enabled, ofb_forward_f32 brackets every kernel launch with CUDA
 * events on the launch stream.  ofb_profile_report waits for them and writes one line per
 * kernel class, "name launches total_ms algorithmic_flops algorithmic_bytes\n"; returns the
 * number of bytes written and clears the records. */
int ofb_profile_enable(ofb_handle* h, int on);
int ofb_profile_report(ofb_handle* h, char* buf, int capacity);

/* Test hook: which instantiation of the tcgen05 conv kernel the calling thread's last ofb_conv_f32 / engine conv
 * selected: "cta2" (cta_group::2 CTA pairs), "khr" (kh-reuse boxes), "ups" (rolling-row fused upsample), "splitk",
 * "plain". */
const char* ofb_last_conv_variant(void);

/* Kernel launches issued by this library on the calling thread since the last reset. */
int64_t ofb_launch_count(int reset);

/* Timing experiments: copies the clock stamps an epilogue warp of CTA 0 recorded while "tc_debug" & 16
 * was set (512 tiles x 8 int64) to host_dst (tools/probe_tail.py). */
int ofb_debug_stamps(long long* host_dst);
int ofb_debug_set(int tc_debug);      /* "tc_debug" for convs launched directly through ofb_conv_f32 */
int ofb_debug_nstack(int on);         /* "nstack" for rolling-row kernels launched directly (tools/probe_rolling.py) */
/* Timing experiments: with "tc_debug" & 256 CTA 0 of every tcgen05 conv launch records eight %globaltimer stamps
 * (kernel entry, prologue done, producer past its dependency wait, first operands landed, last MMA issued, first
 * accumulator complete, last store issued, kernel end).  Copies up to max_slots x 8 int64 of the launches since the
 * last call to host_dst and returns their number (tools/timeline.py). */
int ofb_debug_timeline(long long* host_dst, int max_slots);
int ofb_debug_timeline_raw(long long* host_dst);   /* the whole 1024 x 8 stamp buffer, no reset (tools/chain_timeline.py) */

#ifdef __cplusplus
}
#endif
#endif /* OFB_H_ */
