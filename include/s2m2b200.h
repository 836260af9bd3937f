/*
 * s2m2b200 — C-ABI of the Blackwell-native (sm_100a) S2M2 stereo inference kernels.
 *
 * The reference (junhong-3dv/s2m2) is pure Python/PyTorch and has no FFI of its own: its boundary
 * is the nn.Module surface S2M2.forward (src/s2m2/core/model/s2m2.py:136-197).  This header is the
 * operator-level boundary that the replacement module (s2m2_b200/model.py) binds through ctypes;
 * each entry point names the reference code it replaces.  Plain pointers and sizes only — no torch
 * types.  All pointers are DEVICE pointers unless stated otherwise; `stream` is a cudaStream_t.
 *
 * Conventions
 *   - Activations are NHWC.  An S2m2Tensor describes a view: pixel p=(n*H+y)*W+x starts at
 *     ptr + p*ld elements, channels are contiguous.  `ld >= C` lets a view be a channel slice of a
 *     wider buffer (this is how torch.cat is avoided).
 *   - dtype: 0 = float32, 1 = float16.  GEMM-class kernels accumulate in fp32.
 *   - Every function returns 0 on success and a negative code on error; s2m2b200_last_error()
 *     returns a thread-local message.  Kernels are enqueued on `stream`, nothing synchronises.
 *   - Ownership: the caller owns every buffer.  The library keeps no state between calls except
 *     cached cudaFuncAttributes / TMA descriptors (not thread-safe; one host thread per device).
 */
#ifndef S2M2B200_H_
#define S2M2B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2M2_F32 0
#define S2M2_F16 1

#define S2M2_ACT_NONE 0
#define S2M2_ACT_GELU 1     /* exact erf GELU (nn.GELU default) */
#define S2M2_ACT_RELU 2
#define S2M2_ACT_SIGMOID 3
#define S2M2_ACT_TANH 4

typedef struct {
  void* ptr;
  int32_t dtype;
  int32_t N, H, W, C;
  int64_t ld;
} S2m2Tensor;

/* Implicit-GEMM convolution / linear layer:
 *   y[p, co] = act( sum_{ky,kx,ci} xcat[p*stride - pad + (ky,kx), ci] * w[co, (ky*KW+kx)*Cin + ci] + bias[co] ) (+ res[p, co])
 * where xcat = concat(x, x2) along channels (x2 optional: x2.ptr == NULL), zero padding.
 * Replaces every nn.Conv2d / nn.Linear / nn.ConvTranspose2d on the path (attentions.py:24-30,71-74,
 * 239-241,269-275; submodules.py:70-86,104-108,127-135; feature_fusion.py:14-21; refinenet.py:14-20,
 * 47-57,87-122; unet.py:25-37; stacked_MRT.py:22-34; s2m2.py:65-67), with the activation, the residual
 * add and the concat of the callers fused in.
 *   shuffle2x2 != 0: ConvTranspose2d(k=2,s=2) — GEMM columns are (a*2+b)*Cr + co (Cr = Cout/4) and are
 *   scattered to y[n, 2*oy+a, 2*ox+b, co]; y must be (N, 2*Ho, 2*Wo, Cr).
 *   batch > 1: independent GEMMs; x/w/y pointers advance by x_bs/w_bs/y_bs elements per batch entry
 *   (used for the per-scan-line correlation that builds the cost volume, submodules.py:216-217).
 * w dtype = w_dtype (same as x unless x is fp32 scalar maps feeding fp16 weights); bias is fp32. */
typedef struct {
  S2m2Tensor x, x2, y, res;     /* x2.ptr / res.ptr may be NULL */
  const void* w;
  int32_t w_dtype;
  const float* bias;            /* may be NULL */
  int32_t KH, KW, stride, pad_h, pad_w;
  int32_t Cout;                 /* GEMM N (4*Cr when shuffle2x2) */
  int32_t act;
  int32_t shuffle2x2;
  int32_t batch;                /* >= 1 */
  int64_t x_bs, w_bs, y_bs;
  int32_t impl;                 /* 0 = auto, 1 = force CUDA-core fp32 FMA kernel, 2 = force tcgen05 kernel */
  /* Epilogue modes of the tcgen05 path (fp16 in / out), a = act(acc + bias), R1 = res, R2 = res2 at the same pixel:
   *   0  y = a (+ R1)                                   the plain layer with its residual
   *   1  y[:, n] = a * R1[:, n - res_from]  for n >= res_from (columns below res_from: y = a)
   *         ConvGRU: [z | r] = sigmoid(conv([h, x])) with r * h formed in place of r (refinenet.py:25-27,31-33)
   *   2  g = clamp(a, .01, .99);  y = g * R1 + (1 - g) * R2
   *         FeatureFusion's gate applied to its two inputs (feature_fusion.py:27-29); the fusion branch then adds this as its residual
   *   3  y = R1 * a + (1 - R1) * R2
   *         ConvGRU state update h = (1 - z) h + z q, q = tanh(conv([r h, x])), R1 = z, R2 = h, y may alias R2 (refinenet.py:28-29,34-35)
   * res_from: first output column that has an R1 column (mode 1), 0 otherwise.  Modes 1-3 exist on the tcgen05 path only:
   * s2m2b200_conv2d returns 1 (nothing launched) when the layer does not qualify, and the caller composes the step from separate ops. */
  S2m2Tensor res2;              /* R2 of modes 2 / 3 (fp16 view); ptr NULL otherwise */
  int32_t emode, res_from;
} S2m2Conv;
int s2m2b200_conv2d(const S2m2Conv* a, void* stream);

/* Multi-head attention over S sequences of N tokens (F.scaled_dot_product_attention /
 * explicit softmax, attentions.py:33-54,77-96).  qkv is the fused projection output: token t of
 * sequence s is row s*N+t; q,k,v at channel offsets 0, C, 2C; head h uses channels [h*d,(h+1)*d).
 * Keys/values are taken from sequence (s + kv_shift) % S (symmetric cross attention: kv_shift = S/2).
 * If pesum.ptr != NULL the contextual relative positional term of attentions.py:46-47 is produced:
 *   pesum[s,t,h,0:16]  = sum_j P[t,j] * 0.5*pe_x[x_t - x_j + pw - 1]
 *   pesum[s,t,h,16:32] = sum_j P[t,j] * 0.5*pe_y[y_t - y_j + ph - 1],  token t = y*pw + x
 * (the (N,N,32) tensor of utils.py:32-60 is never materialised).
 * work / work_bytes: optional scratch (s2m2b200_attention_workspace_bytes) that enables the tcgen05 whole-row kernel
 * for the scan-line blocks (it holds V^T); NULL / 0 selects the mma.sync / CUDA-core kernels. */
int s2m2b200_attention(const S2m2Tensor* qkv, const S2m2Tensor* out, int32_t S, int32_t N, int32_t heads,
                       int32_t kv_shift, const float* pe_x, const float* pe_y, int32_t ph, int32_t pw,
                       const S2m2Tensor* pesum, void* work, int64_t work_bytes, void* stream);
/* Scratch bytes the tcgen05 attention kernel needs for (S sequences, N tokens, heads, head dim d); 0 = not applicable. */
int64_t s2m2b200_attention_workspace_bytes(int32_t S, int32_t N, int32_t heads, int32_t d, int32_t dtype);

/* All-pairs scan-line correlation = the cost volume of DispInit with its affine LayerNorm fused in front (submodules.py:165,216-217):
 *   cv[b,y,i,j] = <LN(f)[b,y,i,:], LN(f)[B+b,y,j,:]>
 * on tcgen05 with fp32-grade accuracy: every normalised value is split x = hi + lo into two fp16 numbers and the product is
 * evaluated as hi.hi + hi.lo + lo.hi into one fp32 TMEM accumulator, so that the arg-max over the volume stays where the fp32
 * reference puts it.  f: (2B, h, w, C) fp16 NHWC view (left images first), C in {64,128,192,256}, w % 4 == 0, w <= 512;
 * gamma, beta: fp32 (C); cv: (B, h, w, w) fp32 dense; work: device scratch of s2m2b200_row_corr_workspace_bytes (the hi / lo planes).
 * Returns 0 = launched, 1 = not applicable (caller: s2m2b200_layernorm + the fp32 CUDA-core batched GEMM of s2m2b200_conv2d). */
int64_t s2m2b200_row_corr_workspace_bytes(int32_t B, int32_t h, int32_t w, int32_t C);
int s2m2b200_row_corr(const S2m2Tensor* f, const float* gamma, const float* beta, float* cv, void* work, int64_t work_bytes,
                      void* stream);

/* Indicator matrix of an (h x w) token grid for the tcgen05 PE attention (attentions.py:42-48, utils.py:32-60): the two
 * marginals of the attention matrix that the relative-PE term needs are P . E with E[key][c] = (c == x(key)) or
 * (c == w + y(key)), fp16, (ceil(h*w/128)*128) x 128.  The caller fills a HOST buffer of s2m2b200_pe_indicator_bytes(h, w)
 * bytes once per resolution, keeps a 128-byte-aligned device copy and passes it as `work` to s2m2b200_attention together
 * with `pesum` (h + w <= 80, head dim 64); without it the mma.sync kernel evaluates the PE term. */
int64_t s2m2b200_pe_indicator_bytes(int32_t h, int32_t w);
int s2m2b200_pe_indicator(int32_t h, int32_t w, void* host_out);

/* Fused pre-norm FFN (attentions.py:229-250): y = x + W2 . GELU(W0 . LayerNorm(x) + b0) + b2, one tcgen05 kernel that
 * keeps LN(x) and the hidden activations in shared memory.  x, y: (.., C) fp16 views (y may alias x); w0, w2: (C, C)
 * K-major fp16; b0, b2 fp32.  Returns 0 = launched, 1 = not applicable (caller runs layernorm + two conv2d), < 0 = error. */
int s2m2b200_ffn(const S2m2Tensor* x, const S2m2Tensor* y, const void* w0, const float* b0, const void* w2, const float* b2,
                 int32_t w_dtype, void* stream);

/* Fused two-layer point-wise block (two token tiles in flight per CTA, hidden activations never leave the SM):
 *   y = [res +] W2 . act(W0 . (layernorm ? LayerNorm(x) : x) + b0) + b2
 * (layernorm = 1, act = GELU): the pre-norm FFN above with res = x;  (layernorm = 0, act = ReLU): the 1x1 branch of
 * ConvBlock2D (attentions.py:273-281, `convs_1x`) with res = the 3x3 branch's output, so that the block's sum needs no
 * extra pass.  x, res, y: (.., C) fp16 views, C in {128,192,256}; y may alias res / x.  Returns 0 = launched,
 * 1 = not applicable (caller runs the separate launches), < 0 = error. */
int s2m2b200_mlp2(const S2m2Tensor* x, const S2m2Tensor* res, const S2m2Tensor* y, const void* w0, const float* b0,
                  const void* w2, const float* b2, int32_t w_dtype, int32_t layernorm, int32_t act, void* stream);

/* Fused pre-norm projection (attentions.py:119-130,150-161: q/k/v = Linear(LayerNorm(x))): y = W . LayerNorm(x) + b with the
 * normalised tile kept in shared memory as the A operand for every output-channel tile.  x: (.., C) fp16, C in {128,192,256};
 * y: (.., N) fp16, N % 64 == 0; w: (N, C) K-major fp16; bias fp32 or NULL.  Returns 0 = launched, 1 = not applicable (caller
 * runs layernorm + conv2d), < 0 = error. */
int s2m2b200_ln_linear(const S2m2Tensor* x, const S2m2Tensor* y, const void* w, const float* bias, int32_t w_dtype, void* stream);

/* ---- the callers either side of the forward ---- */
/* uint8 ingest (§8(f)1): same as s2m2b200_preprocess / s2m2b200_stem but the NCHW images are uint8 (4x fewer input bytes; the
 * reference converts to float on the host, model_utils.py:72-73 after image_utils.py:20-21).  Bit-identical results. */
int s2m2b200_preprocess_u8(const uint8_t* left, const uint8_t* right, int32_t B, int32_t H, int32_t W, const S2m2Tensor* out,
                           void* stream);
int s2m2b200_stem_u8(const uint8_t* left, const uint8_t* right, int32_t B, int32_t H, int32_t W, const float* w0, const float* b0,
                     const float* w1, const float* b1, const S2m2Tensor* out, void* stream);
/* Batched confidence score (§8(f)2; model_utils.py:93-94 after image_crop, image_utils.py:66-96): out[b] = mean of conf[b] over
 * the centre crop (crop_h, crop_w) shrunk by `margin` on every side.  conf: (B,H,W) fp32; out: B floats (device). */
int s2m2b200_confidence_score(const float* conf, int32_t B, int32_t H, int32_t W, int32_t crop_h, int32_t crop_w, int32_t margin,
                              float* out, void* stream);
/* Disparity -> camera-frame points (§8(f)3; model_utils.py:111-136 with open3d's RGBD back-projection, trust mask of
 * vis_utils.py:62).  disp/conf/occ: (B,H,W) fp32 (conf, occ may be NULL); xyz: (B,H,W,3) fp32; valid: (B,H,W) bytes. */
int s2m2b200_disparity_to_points(const float* disp, const float* conf, const float* occ, int32_t B, int32_t H, int32_t W, float fx, float cx,
                                 float cy, float baseline, float doffs, float depth_scale, float depth_trunc, float* xyz, uint8_t* valid,
                                 void* stream);

/* Relative-PE table of utils.py:32-60 for one axis of length n: HOST output (2n+1, 16) floats. */
int s2m2b200_pe_table(int32_t n, float* host_out);

/* LayerNorm over C, eps 1e-5; gamma/beta NULL = no affine (attentions.py:117,...; submodules.py:165). */
int s2m2b200_layernorm(const S2m2Tensor* x, const S2m2Tensor* y, const float* gamma, const float* beta, void* stream);
/* GroupNorm(groups, C), eps 1e-5, affine (submodules.py:80,90).  Deterministic (atomic-free) statistics.
 * work: 2*N*groups*(1 + ceil(H*W/256)) doubles. */
int s2m2b200_groupnorm(const S2m2Tensor* x, const S2m2Tensor* y, int32_t groups, const float* gamma,
                       const float* beta, double* work, void* stream);
/* AvgPool2d(2) (unet.py:25) and Upsample(x2, bilinear, align_corners=False) (unet.py:32). */
int s2m2b200_avgpool2(const S2m2Tensor* x, const S2m2Tensor* y, void* stream);
int s2m2b200_upsample2x(const S2m2Tensor* x, const S2m2Tensor* y, void* stream);
/* out = f + w*z0 + (1-w)*z1, w = clamp(g, .01, .99) with g already sigmoid-ed (feature_fusion.py:27-29). */
int s2m2b200_gate_blend(const S2m2Tensor* f, const S2m2Tensor* g, const S2m2Tensor* z0, const S2m2Tensor* z1,
                        const S2m2Tensor* out, void* stream);
int s2m2b200_copy_channels(const S2m2Tensor* src, const S2m2Tensor* dst, void* stream);
int s2m2b200_unary(const S2m2Tensor* x, const S2m2Tensor* y, int32_t act, void* stream);
int s2m2b200_mul(const S2m2Tensor* a, const S2m2Tensor* b, const S2m2Tensor* out, void* stream);
/* h = (1-z)*h + z*q in place (refinenet.py:29,35). */
int s2m2b200_gru_update(const S2m2Tensor* z, const S2m2Tensor* q, const S2m2Tensor* h, void* stream);
/* (x/255-.5)*2 of both NCHW fp32 images, stacked [left..., right...] on N, to NHWC (s2m2.py:86-96). */
int s2m2b200_preprocess(const float* left, const float* right, int32_t B, int32_t H, int32_t W,
                        const S2m2Tensor* out, void* stream);

/* Fused stem: normalise -> Conv1x1(3->16) -> GELU -> Conv1x1(16->16) on both NCHW fp32 images (s2m2.py:86-87,
 * submodules.py:70-72,89).  Two output layouts, chosen by the shape of `out`:
 *   (2B,H,W,C>=16)      pixel-major, channels beyond 16 written as zeros;
 *   (2B,H/2,W/2,64) fp16  space-to-depth: channel ((y&1)*2 + (x&1))*16 + c of block (y/2, x/2) — the 5x5 stride-2 conv that
 *                        follows (submodules.py:74) then is a dense 3x3 stride-1 conv over 64 channels (model._pack_d0).  w0 (16,3), w1 (16,16), biases (16): fp32 HOST pointers (they become kernel parameters). */
int s2m2b200_stem(const float* left, const float* right, int32_t B, int32_t H, int32_t W, const float* w0, const float* b0,
                  const float* w1, const float* b1, const S2m2Tensor* out, void* stream);
/* Fused UpsampleMask1x inputs (submodules.py:121-124,138-139): out[...,0:16] = relu(convT3x3(up[...,0])),
 * out[...,16:32] = relu(convT3x3(rgb)); wd (16,9) / wr (16,27) are the equivalent flipped conv weights, tap-major;
 * weights and biases are fp32 HOST pointers. */
int s2m2b200_mask1x_inputs(const float* up, const S2m2Tensor* rgb, const float* wd, const float* bd, const float* wr,
                           const float* br, const S2m2Tensor* out, void* stream);

/* Dust-bin log-Sinkhorn + argmax + 5-tap soft-argmax over the (B,h,w,w) fp32 cost volume
 * (submodules.py:147-243): disp/conf/occ (B,h,w) fp32, amax int32 = prob_max_ind (submodules.py:226). */
int s2m2b200_sinkhorn_softargmax(const float* cv, int32_t B, int32_t h, int32_t w, float* disp, float* conf,
                                 float* occ, int32_t* amax, int32_t use_positivity, int32_t iters, void* stream);
/* CostVolume.__call__ (submodules.py:39-60): out (B,h,w,18) fp32 = 9 taps level 1 | 9 taps pooled level, /16. */
int s2m2b200_cv_lookup(const float* cv, const float* disp, int32_t B, int32_t h, int32_t w, float* out, void* stream);

/* GlobalRefiner glue (refinenet.py:63-66,72): out[...,0] = disp/100*mask, out[...,1] = logit(mask*conf, .1). */
int s2m2b200_global_refiner_prep(const float* disp, const float* conf, int64_t npix, const S2m2Tensor* out, void* stream);
/* disp = mask*disp + (1-mask)*upd[i*upd_ld] (upd pre-scaled by 100), optional clamp >= 0 (s2m2.py:160-161). */
int s2m2b200_global_refiner_merge(float* disp, const float* conf, const float* upd, int64_t upd_ld, int64_t npix,
                                  int32_t use_positivity, void* stream);
/* LocalRefiner glue (refinenet.py:134-135,140): small (npix,4) = disp/100 | logit(conf,.01) | logit(occ,.01) | 0. */
int s2m2b200_local_refiner_prep(const float* disp, const float* conf, const float* occ, int64_t npix,
                                float* small, void* stream);
/* refinenet.py:147-154 + s2m2.py:177-180: upd (npix, upd_ld >= 3) fp32 = [delta disp | delta conf logit | delta occ logit | ...]:
 * disp += upd0; conf/occ = sigmoid(upd1/2 + logit); clamp; occ *= (x - disp >= 0). */
int s2m2b200_local_refiner_update(float* disp, float* conf, float* occ, const float* small, const float* upd,
                                  int64_t upd_ld, int32_t B, int32_t h, int32_t w, int32_t use_positivity, void* stream);

/* RAFT convex 4x upsampling of disp*4 | occ | conf with one set of 9 logits (s2m2.py:101-114,184-187):
 * logits (B,4h,4w,logits_ld >= 9) fp32, first 9 channels used -> up (B,4h,4w,4) fp32 (4th channel 0). */
int s2m2b200_upsample4x_convex(const float* disp, const float* occ, const float* conf, const float* logits,
                               int32_t logits_ld, int32_t B, int32_t h, int32_t w, float* up, void* stream);
/* 9-tap softmax filter at full resolution (s2m2.py:116-133,190-193): out (3,B,1,H,W) planar fp32.
 * upsample2 != 0 = the output_upsample=True branch (s2m2.py:123-127,194-195): out (3,B,1,2H,2W), logits bilinearly
 * interpolated x2 before the softmax, neighbourhood nearest x2, disparity doubled. */
int s2m2b200_filter1x(const float* up, const float* logits, int32_t logits_ld, int32_t B, int32_t H, int32_t W, float* out,
                      int32_t upsample2, void* stream);
/* The same filter writing only the centre window (crop_h x crop_w) of the padded frame = image_crop (image_utils.py:66-96) fused
 * into the last kernel of the forward: out = three dense (B, crop_h, crop_w) fp32 maps [disparity | occlusion | confidence]. */
int s2m2b200_filter1x_crop(const float* up, const float* logits, int32_t logits_ld, int32_t B, int32_t H, int32_t W, int32_t crop_h,
                           int32_t crop_w, float* out, void* stream);
/* image_pad (image_utils.py:27-63; called by run_stereo_matching, model_utils.py:72-73) on the device: centre padding of a
 * (B,3,H,W) uint8 / fp32 image to multiples of `factor`, the padding filled with the reference's blurred copy
 * (adaptive_avg_pool2d to (H/factor, W/factor) of the zero-padded image, bilinear back up), the original pasted over it.
 * out: (B,3,Hn,Wn) fp32, work: 3*B*(H/factor)*(W/factor) floats of device scratch. */
/* Stereo rectification of uint8 images: cv2.remap(src, mapx, mapy, INTER_LINEAR, border constant 0), bit-exact with OpenCV's
 * fixed-point 8-bit path (image_utils.py:108-136 rectify_images; the calibration optimisers call it per candidate rotation,
 * calibration/base.py:15-36, cem.py:64-72).  src: DEVICE (Hs,Ws,3) interleaved uint8 as cv2 reads it; source n starts at
 * src + n*src_stride bytes (0 = all N outputs sample the same image: N candidate rectifications of one frame); mapx / mapy: DEVICE
 * (N,Hd,Wd) fp32 source coordinates (cv2.initUndistortRectifyMap, CV_32FC1); dst: DEVICE uint8, planar (N,3,Hd,Wd) — what
 * s2m2b200_forward / S2M2.forward take — when planar != 0, else interleaved (N,Hd,Wd,3) like cv2's result. */
int s2m2b200_remap_u8(const uint8_t* src, int64_t src_stride, int32_t Hs, int32_t Ws, const float* mapx, const float* mapy, int32_t N,
                      int32_t Hd, int32_t Wd, uint8_t* dst, int32_t planar, void* stream);
int s2m2b200_image_pad(const void* img, int32_t is_u8, int32_t B, int32_t H, int32_t W, int32_t factor, float* out, float* work,
                       void* stream);


/* ---- the whole forward behind one handle (host side in C++: s2m2_b200/csrc/host_forward.cu) --------------------------------
 * Replaces, for a host that is not Python, what s2m2.py:14-78 (constructor, my_load_state_dict) and s2m2.py:136-197 (forward)
 * do around the ops above; model_utils.py:30-43 (load_model) and model_utils.py:72-83 (run_stereo_matching) are its callers.
 *
 *   create        feature_channels / num_transformer / use_positivity / output_upsample / refine_iter as S2M2.__init__
 *                 (dim_expansion is 1, as load_model always passes); precision S2M2_F16 (tensor-core path) or S2M2_F32 (parity mode)
 *   set_param     one state-dict entry under the reference's key (e.g. "refiner.gru.convz1.weight"): HOST fp32 data, copied
 *   finalize      packs every layer into the layout its kernel reads and uploads the blob to the current device; fails naming the
 *                 first missing / mis-shaped key
 *   workspace     bytes of scratch one forward of (B, H, W) needs: the schedule is walked once in plan mode, nothing is launched;
 *                 also builds the shape's constant tables (synchronous; call it outside stream capture).  < 0 on error
 *   forward       left / right: DEVICE (B,3,H,W) planar images, fp32 (values 0..255) or uint8 (in_u8 != 0); H, W multiples of 32;
 *                 out: DEVICE 3 x (B,1,Ho,Wo) fp32 planes [disparity | occlusion | confidence] (s2m2.py:197), Ho x Wo = H x W,
 *                 2H x 2W with output_upsample, or crop_h x crop_w (the centre window image_crop keeps, image_utils.py:66-96) when
 *                 crop_h, crop_w > 0; workspace: DEVICE scratch, 256-byte aligned.  Enqueues on `stream` only: no allocation, no
 *                 synchronisation, capturable in a CUDA graph once the shape has been planned.
 *   graph mode    off by default.  On: forward captures the launches of a shape into a CUDA graph the first time it sees a set of
 *                 buffer addresses (left, right, out, workspace) and replays it on later calls with the same addresses — one graph
 *                 launch instead of ~900 kernel launches (a host that reuses its buffers; other addresses re-capture).
 * All return 0 on success; s2m2b200_last_error() has the message otherwise. */
typedef struct S2m2Model S2m2Model;
int s2m2b200_create(int32_t feature_channels, int32_t num_transformer, int32_t use_positivity, int32_t output_upsample, int32_t refine_iter,
                    int32_t precision, S2m2Model** out);
int s2m2b200_set_param(S2m2Model* m, const char* name, const float* data, const int64_t* shape, int32_t ndim);
int s2m2b200_finalize_params(S2m2Model* m);
int64_t s2m2b200_workspace_bytes(S2m2Model* m, int32_t B, int32_t H, int32_t W, int32_t in_u8, int32_t crop_h, int32_t crop_w);
int s2m2b200_forward(S2m2Model* m, const void* left, const void* right, int32_t in_u8, int32_t B, int32_t H, int32_t W, int32_t crop_h,
                     int32_t crop_w, float* out, void* workspace, int64_t workspace_bytes, void* stream);
int s2m2b200_set_graph_mode(S2m2Model* m, int32_t on);
/* run_stereo_matching (model_utils.py:50-95) behind the handle: image_pad of both images to multiples of 32 (image_utils.py:27-63),
 * the forward, image_crop back to H0 x W0 fused into its last kernel (model_utils.py:88-91) and the interior confidence mean
 * (model_utils.py:93-94).  left / right: DEVICE (B,3,H0,W0) planar, fp32 or uint8, ANY size >= 32; out: DEVICE 3 x (B,1,H0,W0) fp32
 * [disparity | occlusion | confidence]; score: DEVICE (B) fp32 or NULL (NaN when 2*margin leaves no interior); the reference uses
 * margin = 100.  Workspace from s2m2b200_stereo_match_workspace_bytes. */
int64_t s2m2b200_stereo_match_workspace_bytes(S2m2Model* m, int32_t B, int32_t H0, int32_t W0);
int s2m2b200_stereo_match(S2m2Model* m, const void* left, const void* right, int32_t in_u8, int32_t B, int32_t H0, int32_t W0, int32_t margin,
                          float* out, float* score, void* workspace, int64_t workspace_bytes, void* stream);
void s2m2b200_destroy(S2m2Model* m);
/* The packed blob's directory (tests): entry `index` of s2m2b200_packed_count; ptr is device memory when *on_device. */
int64_t s2m2b200_packed_count(const S2m2Model* m);
int s2m2b200_packed_entry(const S2m2Model* m, int64_t index, const char** name, const void** ptr, int64_t* bytes, int32_t* dtype,
                          int32_t* on_device);

const char* s2m2b200_last_error(void);
/* Version / capability string, e.g. "s2m2b200 0.1 sm_100a tcgen05". */
const char* s2m2b200_version(void);
/* Number of kernel launches issued by this library since process start (bench.py's gpu_launches). */
int64_t s2m2b200_launch_count(void);
/* Plan mode (thread-local; returns the previous value): while on, every op above runs its host-side checks and dispatch and
 * returns the code it would return, but enqueues nothing — how s2m2b200_workspace_bytes walks the forward schedule. */
int s2m2b200_set_plan_mode(int on);
/* Op trace (thread-local): between _begin and _end every op above appends one text line (name, tensor views, pointers, scalars);
 * _note adds a caller's own line (allocations, named regions).  _end copies the text out and returns its full length. */
int s2m2b200_trace_begin(void);
void s2m2b200_trace_note(const char* text);
int64_t s2m2b200_trace_end(char* buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* S2M2B200_H_ */
