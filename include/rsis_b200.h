/*
 * rsis_b200 -- C ABI of the B200-native (sm_100a) RSIS hot path.
 *
 * The reference (imatge-upc/rsis) has no FFI/plugin layer: its hot path is a set of Python
 * nn.Modules (`src/modules/{vision,model,clstm}.py`) that dispatch to torch.nn primitives.
 * This header is therefore the boundary *this* repo defines: one entry point per primitive
 * group the reference invokes, each citing the reference call site it replaces.  The Python
 * package `rsis_b200` binds it with ctypes and re-exposes the reference's own module surface
 * (`FeatureExtractor`, `RSIS`, `ConvLSTMCell`, `test()`).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (PyTorch's caching allocator); the library never allocates or frees tensor memory;
 *   - every call is asynchronous on the `stream` passed in (a cudaStream_t), performs no
 *     host synchronisation and is safe under CUDA-graph capture;
 *   - return value: RSIS_OK (0) or a negative rsis_status; `rsis_strerror` describes it.
 *     Nothing throws or aborts across the ABI;
 *   - activations are NHWC ("channels last").  Two element formats exist:
 *       RSIS_FMT_F32         float32, [N][H][W][C]
 *       RSIS_FMT_SPLIT_BF16  two bfloat16 planes hi|lo, [2][N][H][W][C], value = hi + lo
 *                            (the operand format of the split-precision tcgen05 convolutions:
 *                            a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32 accumulate);
 *   - re-entrant; the only process-wide state is an init-once table of function attributes.
 */
#ifndef RSIS_B200_H_
#define RSIS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSIS_ABI_VERSION 21

typedef void* rsis_stream_t; /* cudaStream_t */

typedef enum rsis_status {
  RSIS_OK = 0,
  RSIS_ERR_BAD_ARG = -1,     /* null pointer, negative size, inconsistent shapes */
  RSIS_ERR_UNSUPPORTED = -2, /* shape / format combination this build does not implement */
  RSIS_ERR_CUDA = -3,        /* a CUDA runtime call or kernel launch failed (see rsis_last_cuda_error) */
  RSIS_ERR_ARCH = -4,        /* current device is not sm_100 (B200) */
  RSIS_ERR_ALIGN = -5        /* pointer not aligned as the kernel requires (16 B) */
} rsis_status;

enum { RSIS_FMT_F32 = 0, RSIS_FMT_SPLIT_BF16 = 1 };
enum { RSIS_IMPL_AUTO = 0, RSIS_IMPL_SIMT = 1, RSIS_IMPL_TCGEN05 = 2 };
/* rsis_convlstm_cell only: `impl | RSIS_IMPL_CTA_CAP(n)` limits the persistent tcgen05 kernel to n CTAs (n SMs), so that
 * the cells of different decoder levels -- independent along the (level, step) wavefront -- share the chip side by side
 * instead of each launch taking every SM in turn.  Ignored by launches that split K across CTAs. */
#define RSIS_IMPL_CTA_CAP(n) (((n) & 0xffff) << 8)

/* NHWC activation view.  `cstride` is the pixel pitch in elements (0 means dense, = c): a view with
 * cstride > c is a channel slice [data, data + c) of a wider NHWC buffer -- this is how the decoder's
 * `torch.cat([upsampled hidden, skip, prev_hidden], 1)` (model.py:153, clstm.py:43) exists without a copy: the
 * producers write their slice of one buffer.  For RSIS_FMT_SPLIT_BF16 `data` points at the hi plane;
 * lo = hi + n*h*w*pitch elements.  Kernels that do not take pitched views return RSIS_ERR_UNSUPPORTED. */
typedef struct rsis_tensor {
  void* data;
  int32_t fmt;
  int32_t n, h, w, c;
  int32_t cstride;
} rsis_tensor;

/* Packed convolution weights + folded per-channel affine (bias and eval-mode BatchNorm):
 *   y[co] = acc[co] * scale[co] + shift[co]
 * produced by rsis_conv_pack from the reference's OIHW float32 parameters. */
typedef struct rsis_conv_weights {
  const float* w_kc;  /* SIMT pack: [KH*KW*Cin][cout_pad] float32, cout_pad = ceil(Cout/64)*64, zero padded */
  const void* w_umma; /* tcgen05 pack: [2 planes hi|lo][cout_pad16][k_pad] bfloat16, K-major; may be NULL */
  const float* scale; /* [cout_pad] */
  const float* shift; /* [cout_pad] */
  int32_t cout, cin, kh, kw;
  int32_t gate_interleaved; /* 1: output channel order is (hidden channel, gate) -- ConvLSTM packs */
  const void* w_umma_il;    /* optional: the tcgen05 pack with its two planes interleaved per output channel,
                             * [cout_pad16][2 planes][k_pad] bfloat16 (same values as w_umma).  When given for a
                             * ConvLSTM pack with cout = 32 or 64 and <= 64 input channels, rsis_convlstm_cell uses its
                             * swapped-operand kernel (weights as the M operand, 256 pixels as N) on maps whose width is
                             * a multiple of 8 and height a multiple of 32.  May be NULL. */
} rsis_conv_weights;

/* Precision of the tcgen05 convolution family (process-wide; read when a launch is set up, so a captured CUDA graph
 * keeps the mode it was captured in).
 *   RSIS_PRECISION_SPLIT_BF16 (default): a*b = a_hi*b_hi + a_hi*b_lo + a_lo*b_hi (+ a_lo*b_lo) on bf16 tensor cores with
 *     fp32 accumulation -- fp32-grade products, the mode every 1e-3 parity claim is made in.
 *   RSIS_PRECISION_BF16: single-pass bf16 operands (the hi planes only), fp32 accumulation, fp32 master weights and
 *     gradients -- the "training step bf16" of BASELINE.json configs[3]; loss-level tolerance.
 * rsis_set_precision returns the previous mode (>= 0) or RSIS_ERR_BAD_ARG. */
enum { RSIS_PRECISION_SPLIT_BF16 = 0, RSIS_PRECISION_BF16 = 1 };
int rsis_set_precision(int mode);
int rsis_get_precision(void);

/* Static-weights mode of the tcgen05 convolution family (process-wide; read when a launch is set up, so a captured CUDA
 * graph keeps the mode it was captured in).  on != 0 promises that no kernel enqueued on the launch's stream right
 * before a convolution / cell writes that launch's PACKED WEIGHTS (true for inference: the packs of the reference's
 * nn.Conv2d / BatchNorm2d parameters, model.py:43-54, clstm.py:17, are built once per load_state_dict).  The kernels then
 * start streaming weights into shared memory while the previous kernel of the stream is still running (programmatic
 * dependent launch: only activation reads wait for it).  Off (default) for training steps, which re-pack in-stream.
 * Returns the previous setting. */
int rsis_set_static_weights(int on);

/* ---- library -------------------------------------------------------------------------------------------- */
int rsis_abi_version(void);
const char* rsis_strerror(int status);
const char* rsis_last_cuda_error(void); /* text of the last CUDA error seen by this thread ("" if none) */
int rsis_device_check(void);            /* RSIS_OK iff the current device is compute capability 10.x */
int rsis_has_tcgen05(void);             /* 1 iff this build carries the tcgen05/TMA convolution kernels */

/* ---- weight packing (once per load_state_dict / optimiser step) ---------------------------------------- */
/* Sizes, in bytes, of the packed buffers for one convolution. */
size_t rsis_conv_pack_bytes_simt(int cout, int cin, int kh, int kw);
size_t rsis_conv_pack_bytes_affine(int cout); /* one of scale / shift */
/* Folds nn.Conv2d(bias) [+ nn.BatchNorm2d eval statistics] into (packed weights, scale, shift).
 * Replaces the parameter handling of nn.Conv2d/nn.BatchNorm2d at model.py:43-54, clstm.py:17 and of
 * torchvision's Bottleneck.  bias / bn_* may be NULL (no bias / no BatchNorm); w_kc may be NULL (affine only).
 * gate_interleave != 0 reorders output channels from the reference's [in|remember|out|cell] blocks (clstm.py:47)
 * to (channel, gate) so one thread owns the four gates of a hidden channel. */
int rsis_conv_pack(const float* w_oihw, const float* bias, const float* bn_weight, const float* bn_bias,
                   const float* bn_mean, const float* bn_var, float bn_eps, int cout, int cin, int kh, int kw,
                   int gate_interleave, float* w_kc, float* scale, float* shift, rsis_stream_t stream);
/* tcgen05 pack.  K is laid out [tap][source][64-channel chunk] (each chunk zero padded to 64) to line up with the
 * activation TMA boxes of the n_src inputs (channel counts src_c[0..n_src)), rows are output channels padded to a
 * multiple of 16; two bf16 planes hi|lo. */
int rsis_conv_umma_kpad(int kh, int kw, int n_src, const int32_t* src_c);
int rsis_conv_umma_coutpad(int cout);
size_t rsis_conv_pack_bytes_umma(int cout, int kh, int kw, int n_src, const int32_t* src_c);
int rsis_conv_pack_umma(const float* w_oihw, int cout, int cin, int kh, int kw, int n_src, const int32_t* src_c,
                        int gate_interleave, void* w_umma, rsis_stream_t stream);

/* rsis_conv_pack + rsis_conv_pack_umma in ONE launch (a training step re-packs every convolution: launches matter), and
 * optionally of the DATA-GRADIENT convolution directly from the stored parameter: dgrad != 0 packs the logical
 * weights w'[co'][ci'][i][j] = w[ci'][ci0 + co'][kh-1-i][kw-1-j] (co' in [0, nci), ci' in [0, w_cout)) that
 * rsis_conv_dgrad_weights would materialise (then bias / bn_* / gate_interleave must be unset).  w_oihw is the stored
 * [w_cout][w_cin][kh][kw] tensor.  w_kc and w_umma may each be NULL (that pack is skipped); src_c / n_src describe
 * the tcgen05 K layout of the logical input channels (see rsis_conv_pack_umma). */
int rsis_conv_pack_all(const float* w_oihw, int w_cout, int w_cin, int kh, int kw, int dgrad, int ci0, int nci,
                       const float* bias, const float* bn_weight, const float* bn_bias, const float* bn_mean,
                       const float* bn_var, float bn_eps, int gate_interleave, int n_src, const int32_t* src_c,
                       float* w_kc, float* scale, float* shift, void* w_umma, rsis_stream_t stream);

/* ---- layout / format plumbing ---------------------------------------------------------------------------- */
/* float32 NCHW (the reference's layout, e.g. the image batch of test.py:35) -> NHWC tensor (either format). */
int rsis_nchw_to_nhwc(const float* src_nchw, const rsis_tensor* dst, rsis_stream_t stream);
/* NHWC tensor -> NHWC tensor of the other element format (same shape). */
int rsis_convert(const rsis_tensor* src, const rsis_tensor* dst, rsis_stream_t stream);

/* im2col: y[n,ho,wo,(i*kw + j)*C + c] = x[n, ho*stride - pad + i, wo*stride - pad + j, c] (zero outside the image and in
 * y's channels beyond kh*kw*C).  x float32 dense, y split-bf16 dense with y->c % 8 == 0.  Turns the 3-channel 7x7/s2 stem
 * (vision.py:12) into a 1x1 tensor-core convolution over 147 (+5) channels. */
int rsis_im2col(const rsis_tensor* x, int kh, int kw, int stride, int pad, const rsis_tensor* y, rsis_stream_t stream);

/* ---- workspace ------------------------------------------------------------------------------------------------ */
/* The tcgen05 convolution splits K across co-resident CTAs when a launch has fewer output tiles than SMs; the
 * partial accumulators are exchanged through a caller-owned device buffer of rsis_conv_workspace_bytes() bytes
 * (a fixed size, ~19.4 MB).  It must be zero-filled ONCE after allocation (the kernels leave its counters at zero),
 * 16-byte aligned, and must not be shared by launches that can run concurrently (one per stream).  Passing
 * workspace = NULL is allowed: split-K is then not used. */
size_t rsis_conv_workspace_bytes(void);

/* ---- encoder primitives ----------------------------------------------------------------------------------- */
/* conv (+folded BN/bias) (+residual) (+ReLU).  Replaces nn.Conv2d + nn.BatchNorm2d (+ReLU, + `out += identity`)
 * of vision.py:12-19 / torchvision Bottleneck.forward and the skip heads model.py:59-63.
 * `srcs` are concatenated along C (n_src = 1 for plain convs).  y2 (optional) receives a second copy of the
 * result in another element format (e.g. float32 for the API-visible feature + split-bf16 for the decoder). */
int rsis_conv2d(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const rsis_tensor* residual,
                const rsis_tensor* y, const rsis_tensor* y2, int stride, int pad, int relu, int impl,
                void* workspace, size_t workspace_bytes, rsis_stream_t stream);
/* nn.MaxPool2d(kernel_size=3, stride=2, padding=1) of vision.py:15. */
int rsis_maxpool3x3s2(const rsis_tensor* x, const rsis_tensor* y, rsis_stream_t stream);

/* ---- train-mode BatchNorm (encoder.train(), train.py:71-77) -------------------------------------------------- */
/* nn.BatchNorm2d in training mode = statistics of THIS batch, so it cannot be folded into the convolution that
 * produces its input.  Pipeline: rsis_conv2d with a pack WITHOUT BatchNorm (raw float32 output) ->
 * rsis_bn_train_stats -> rsis_affine_act.
 * rsis_bn_train_stats: x float32 dense NHWC [N,H,W,C], C % 4 == 0.  Computes the biased batch variance / mean over
 * N*H*W, writes scale = weight / sqrt(var + eps) and shift = bias - mean * scale (weight/bias may be NULL = 1/0), and
 * updates running_mean / running_var (unbiased variance) / num_batches_tracked exactly as nn.BatchNorm2d does
 * (momentum < 0 = cumulative average, i.e. momentum=None; running_* may both be NULL).  batch_mean / batch_invstd
 * (optional) receive the batch statistics.  workspace: rsis_bn_workspace_bytes(C) bytes of device memory, zero-filled
 * once (the call leaves it zeroed). */
size_t rsis_bn_workspace_bytes(int channels);
int rsis_bn_train_stats(const rsis_tensor* x, const float* weight, const float* bias, float eps, float momentum,
                        float* running_mean, float* running_var, int64_t* num_batches_tracked, double* workspace,
                        float* scale, float* shift, float* batch_mean, float* batch_invstd, rsis_stream_t stream);
/* y = [relu](x * scale[c] + shift[c] [+ residual]); dense NHWC tensors of one shape, any element formats; y2
 * (optional) receives a second copy in another format.  Replaces the normalisation + ReLU + `out += identity` of a
 * train-mode Bottleneck (torchvision Bottleneck.forward) and the train-mode skip heads (model.py:59-63). */
int rsis_affine_act(const rsis_tensor* x, const float* scale, const float* shift, const rsis_tensor* residual, int relu,
                    const rsis_tensor* y, const rsis_tensor* y2, rsis_stream_t stream);

/* ---- decoder primitives ----------------------------------------------------------------------------------- */
/* One fused ConvLSTM cell step (clstm.py:19-62): gates = conv3x3(cat(srcs)) + bias; i,f,o = sigmoid, g = tanh;
 * c = f*c_prev + i*g; h = o*tanh(c).  The four gate planes never reach HBM.
 *   srcs      inputs concatenated along C in the reference's order [input_ ... | prev_hidden]; prev_hidden may
 *             be omitted (n_src excludes it) when the state is None (clstm.py:26-37: zeros) -- then c_prev=NULL
 *   w         packed with gate_interleave=1
 *   c_prev    float32 NHWC [N,H,W,Ch] or NULL (zero state)
 *   gate_preact  optional float32 [N,H,W,4*Ch] in (hidden channel, gate) order, added to the gate pre-activations:
 *             the contribution of TIME-INVARIANT input channels (the skip feature of model.py:153 and level 0's
 *             x5_skip are the same at every step t), computed once per image by rsis_conv2d with the matching
 *             column block of the gate weights (packed with gate_interleave) and the gate bias -- the per-step
 *             kernel then only contracts over [upsampled hidden | prev_hidden].  tcgen05 kernel family only.
 *   h_out     float32 NHWC; h_split (optional) the same values as split-bf16 for the next consumer
 *   c_out     float32 NHWC
 *   side_max  optional: per-(image, channel) running max of h, as order-preserving uint32 keys, at
 *             side_max[n*side_stride + side_offset + ch] -- the global nn.MaxPool2d of model.py:143.
 *             Must be zero-filled before the step (key 0 sorts below every float). */
int rsis_convlstm_cell(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const float* c_prev,
                       const float* gate_preact, const rsis_tensor* h_out, const rsis_tensor* h_split, const rsis_tensor* c_out,
                       uint32_t* side_max, int side_stride, int side_offset, int impl, void* workspace,
                       size_t workspace_bytes, rsis_stream_t stream);
/* One WAVEFRONT of the decoder in one launch.  The 5-level ConvLSTM stack of model.py:129-165 only has the dependencies
 * cell(level l, step t) <- cell(l-1, t) (through the x2 upsampling) and cell(l, t-1) (its own state), so the cells
 * {(l, t) : l + t = w} of one anti-diagonal are independent: rsis_convlstm_cell_group runs up to
 * rsis_convlstm_cell_group_max() of them (each described like one rsis_convlstm_cell call on its concatenated input
 * buffer `x` = [up(h_below) | h_prev], tcgen05 kernel family) side by side, every cell on a share of the SMs in
 * proportion to its tensor-core work.  Results are identical to the one-by-one calls. */
typedef struct rsis_cell_args {
  const rsis_tensor* x;          /* split-bf16, all w->cin channels */
  const rsis_conv_weights* w;    /* packed with gate_interleave=1 */
  const float* c_prev;           /* or NULL */
  const float* gate_preact;      /* or NULL */
  const rsis_tensor* h_out;
  const rsis_tensor* h_split;    /* or NULL */
  const rsis_tensor* c_out;
  uint32_t* side_max;            /* or NULL */
  int32_t side_stride, side_offset;
} rsis_cell_args;
int rsis_convlstm_cell_group_max(void);
int rsis_convlstm_cell_group(const rsis_cell_args* cells, int n_cells, rsis_stream_t stream);
/* nn.UpsamplingBilinear2d(size=(y.h, y.w)) == bilinear, align_corners=True (model.py:149-150,163-164). */
int rsis_upsample_bilinear(const rsis_tensor* x, const rsis_tensor* y, rsis_stream_t stream);
/* n (<= 4) independent rsis_upsample_bilinear calls xs[i] -> ys[i] in one launch (the upsamplings between two
 * wavefronts of the decoder). */
int rsis_upsample_bilinear_group(const rsis_tensor* xs, const rsis_tensor* ys, int n, rsis_stream_t stream);
/* conv_out (model.py:167): ksize x ksize (1 or 3), Cin -> 1, on an NHWC float32 input; writes logits [N,H,W] float32 (when
 * logits != NULL) and, when prob_out != NULL, sigmoid(logit) at prob_out[n*prob_stride_n + pixel] (the stacking + sigmoid of test.py:46,50). */
int rsis_mask_head(const rsis_tensor* x, const float* w_oihw, const float* bias, int ksize, float* logits,
                   float* prob_out, int64_t prob_stride_n, rsis_stream_t stream);
/* The same result as rsis_upsample_bilinear(h -> out_h x out_w) followed by rsis_mask_head, in one kernel that never
 * writes the upsampled tensor (model.py:163-167: `upsample_match_clstm5` then `conv_out`).  h: float32 dense NHWC with
 * C % 4 == 0 and C <= 16 (RSIS_ERR_UNSUPPORTED otherwise -- use the two-call form). */
int rsis_upsample_mask_head(const rsis_tensor* h, const float* w_oihw, const float* bias, int ksize, int out_h,
                            int out_w, float* logits, float* prob_out, int64_t prob_stride_n, rsis_stream_t stream);
/* rsis_upsample_mask_head for ALL time-steps of a pass in one launch (the `out_masks.append(out_mask)` ... `torch.cat`
 * of test.py:41-46 over the T decoder steps): h holds steps * B images, step-major ([t][b][H][W][C] dense float32);
 * sigmoid(logit) of image (t, b) goes to prob_out[b*prob_stride_n + t*prob_stride_t + pixel]. */
int rsis_upsample_mask_head_steps(const rsis_tensor* h, int steps, const float* w_oihw, const float* bias, int ksize,
                                  int out_h, int out_w, float* prob_out, int64_t prob_stride_n, int64_t prob_stride_t,
                                  rsis_stream_t stream);
/* fc_class + Softmax + fc_stop on the side features (model.py:169-182).  side_max holds the uint32 keys written by
 * rsis_convlstm_cell; feat_out (optional) receives the decoded float features [N, F].  Outputs: class_probs [N,C]
 * written at class_probs[n*class_stride + c], stop logit at stop_logit[n*stop_stride], and optional sigmoid(stop)
 * at stop_prob[n*stop_stride] (test.py:50). */
int rsis_class_stop_heads(const uint32_t* side_max, int n, int f, const float* w_class, const float* b_class,
                          int num_classes, const float* w_stop, const float* b_stop, float* feat_out,
                          float* class_probs, int64_t class_stride, float* stop_logit, float* stop_prob,
                          int64_t stop_stride, rsis_stream_t stream);
/* The same for the side features of ALL time-steps in one launch (test.py:37-50): side_max [steps][n][f] keys;
 * class_probs[b*class_stride + t*class_stride_t + c], sigmoid(stop) at stop_prob[b*stop_stride + t*stop_stride_t]. */
int rsis_class_stop_heads_steps(const uint32_t* side_max, int n, int steps, int f, const float* w_class,
                                const float* b_class, int num_classes, const float* w_stop, const float* b_stop,
                                float* class_probs, int64_t class_stride, int64_t class_stride_t, float* stop_prob,
                                int64_t stop_stride, int64_t stop_stride_t, rsis_stream_t stream);

/* ---- backward primitives: `loss.backward()` of train.py:184 through vision.py / model.py / clstm.py ----------- */
/* All gradients are float32 NHWC.  The DATA gradient of a convolution is itself a forward convolution
 * (rsis_conv2d) of dy with the weights rsis_conv_dgrad_weights produces -- stride 1: pad' = k - 1 - pad; stride 2:
 * rsis_dilate2x(dy) first (3x3) or afterwards (1x1).  Replaces autograd's ConvolutionBackward for nn.Conv2d at
 * vision.py:12, torchvision Bottleneck.conv1-3/downsample, model.py:43-47 (sk*), clstm.py:17 (Gates), model.py:107
 * (conv_out). */
/* out_oihw[ci - ci0][co][kh-1-i][kw-1-j] = w_oihw[co][ci][i][j] for the input channels [ci0, ci0 + nci). */
int rsis_conv_dgrad_weights(const float* w_oihw, int cout, int cin, int kh, int kw, int ci0, int nci, float* out_oihw,
                            rsis_stream_t stream);
/* dw_oihw[co][ci][i][j] (+)= sum_{n,ho,wo} dy[n,ho,wo,co] * x[n, ho*stride - pad + i, wo*stride - pad + j, ci];
 * dbias[co] (+)= sum dy.  x, dy: dense NHWC, either element format.  accumulate = 0 overwrites.  Either of
 * dw_oihw / dbias may be NULL.  Partial sums are combined with float atomics (summation order is not fixed).
 * impl: RSIS_IMPL_SIMT = fp32 CUDA cores (any shape); RSIS_IMPL_AUTO = the tcgen05 kernel (pixels as the contraction
 * dimension, MN-major split-bf16 operands, three products per multiply) for stride-1/2 1x1 / 3x3 convolutions whose x
 * and dy are both RSIS_FMT_SPLIT_BF16, CUDA cores otherwise; RSIS_IMPL_TCGEN05 = that kernel or RSIS_ERR_UNSUPPORTED.
 * workspace: rsis_wgrad_workspace_bytes() bytes, 16-byte aligned, zero-filled ONCE after allocation (the call leaves it
 * zeroed), not shared by launches that may run concurrently; NULL = CUDA cores only. */
size_t rsis_wgrad_workspace_bytes(void);
int rsis_conv2d_wgrad(const rsis_tensor* x, const rsis_tensor* dy, int kh, int kw, int stride, int pad, float* dw_oihw,
                      float* dbias, int accumulate, int impl, void* workspace, size_t workspace_bytes,
                      rsis_stream_t stream);
/* y[n, 2i, 2j, :] = x[n, i, j, :], zero elsewhere; y->h in {2*x->h - 1, 2*x->h} (same for w).  x, y dense, either element
 * format. */
int rsis_dilate2x(const rsis_tensor* x, const rsis_tensor* y, rsis_stream_t stream);
/* Backward of train-mode nn.BatchNorm2d (+ the ReLU that follows it, + the residual branch of `out += identity`):
 *   g = dy * (y_act > 0)  (y_act = the post-activation output; NULL: no ReLU);  dbias = sum g;  dweight = sum g*xhat;
 *   dx = weight * invstd * (g - dbias/M - xhat * dweight/M);  dres (optional) = g.
 * x_raw / dy float32 dense; mean / invstd = the batch statistics rsis_bn_train_stats returned; workspace as for
 * rsis_bn_train_stats; dx either element format.  dweight / dbias receive this call's sums (overwritten);
 * dweight_acc / dbias_acc (optional) are additionally incremented by them (gradient accumulation buffers). */
int rsis_bn_train_bwd(const rsis_tensor* x_raw, const rsis_tensor* y_act, const rsis_tensor* dy, const float* weight,
                      const float* mean, const float* invstd, double* workspace, float* dweight, float* dbias,
                      float* dweight_acc, float* dbias_acc, const rsis_tensor* dx, const rsis_tensor* dres,
                      rsis_stream_t stream);
/* Backward of nn.MaxPool2d(3, 2, 1) (vision.py:15): dy goes to the first maximum of each window. */
int rsis_maxpool3x3s2_bwd(const rsis_tensor* x, const rsis_tensor* dy, const rsis_tensor* dx, rsis_stream_t stream);
/* Training-mode ConvLSTM cell = rsis_conv2d (gate pre-activations, bias folded) + this kernel (clstm.py:47-58).
 * gates: float32 [N,H,W,4*Ch], block order [in|remember|out|cell]; overwritten with the ACTIVATED gates, which the
 * backward needs.  h_out2 (optional, may be a pitched slice, either format) receives a second copy of h. */
int rsis_lstm_gates_fwd(const rsis_tensor* gates, const float* c_prev, const rsis_tensor* h_out, const rsis_tensor* h_out2,
                        const rsis_tensor* c_out, rsis_stream_t stream);
/* Backward of the above: dh = dh_a + dh_b (either may be NULL; float32, pitched slices allowed), dc_next (may be
 * NULL) -> dgates (pre-activation gradients, same layout as gates, either element format) and dc_prev. */
int rsis_lstm_gates_bwd(const rsis_tensor* gates, const float* c_prev, const float* c_new, const rsis_tensor* dh_a,
                        const rsis_tensor* dh_b, const rsis_tensor* dc_next, const rsis_tensor* dgates, float* dc_prev,
                        rsis_stream_t stream);
/* Global nn.MaxPool2d of model.py:143 with its arg-max.  rsis_global_maxpool folds h into side_packed
 * [n*side_stride + side_offset + c] = max over pixels of (order-preserving key << 32 | ~pixel index) -- zero-filled
 * before the step; the FIRST maximum wins ties.  After all levels rsis_global_maxpool_finish unpacks the n*side_stride
 * entries into keys (the format rsis_class_stop_heads reads) and pixel indices. */
int rsis_global_maxpool(const rsis_tensor* h, uint64_t* side_packed, int side_stride, int side_offset,
                        rsis_stream_t stream);
int rsis_global_maxpool_finish(const uint64_t* side_packed, int n, int side_stride, uint32_t* side_keys,
                               int32_t* side_idx, rsis_stream_t stream);
/* dh[n, idx, c] += dside[n*side_stride + side_offset + c]. */
int rsis_global_maxpool_bwd(const float* dside, const int32_t* side_idx, int side_stride, int side_offset,
                            const rsis_tensor* dh, rsis_stream_t stream);
/* Adjoint of rsis_upsample_bilinear: dy float32 (may be a pitched slice) -> dx float32 dense (overwritten). */
int rsis_upsample_bilinear_bwd(const rsis_tensor* dy, const rsis_tensor* dx, rsis_stream_t stream);
/* Backward of rsis_class_stop_heads: feat [N,F] (its feat_out), class_probs [N,C] dense, dclass [N,C] / dstop [N]
 * (either may be NULL = zero); writes dfeat [N,F] and ACCUMULATES dw_class [C,F], db_class [C], dw_stop [F],
 * db_stop [1].  dlogit_scratch: N*(C+1) floats. */
int rsis_class_stop_heads_bwd(const float* feat, const float* class_probs, const float* dclass, const float* dstop,
                              int n, int f, const float* w_class, int num_classes, const float* w_stop,
                              float* dlogit_scratch, float* dfeat, float* dw_class, float* db_class, float* dw_stop,
                              float* db_stop, rsis_stream_t stream);

/* ---- soft-IoU cost / loss of the training loop --------------------------------------- */
/* softIoU of utils/hungarian.py:64-90: s = sigmoid(logit); num = sum(s*y); den = sum(s + y - s*y) + eps;
 * cost = weight * (1 - num/den).
 * rsis_soft_iou_cost: logits float32 [b][hw] against the g ground-truth rows gt [b][g][hw] (float32, or uint8 when
 * gt_is_u8) of the same image, in ONE pass over HBM -- the per-step cost matrix of train.py:96-110 (there:
 * `y_pred_i.repeat`, softIoU, `.cpu()`); with g = 1 and b = rows it is the row-wise softIoU of softIoULoss
 * (utils/objectives.py:27-34).  cost is written at cost[b*cost_stride_b + g*cost_stride_g] (e.g. straight into
 * scores[:, :, t] of train.py:110).  num_out / den_out (optional, [b*g]) keep what rsis_soft_iou_bwd needs.
 * workspace: rsis_soft_iou_workspace_bytes(b, g) bytes, zero-filled once (the call leaves it zeroed).  hw % 4 == 0. */
size_t rsis_soft_iou_workspace_bytes(int b, int g);
int rsis_soft_iou_cost(const float* logits, const void* gt, int gt_is_u8, int b, int g, int64_t hw, float eps,
                       float weight, float* workspace, float* cost, int64_t cost_stride_b, int64_t cost_stride_g,
                       float* num_out, float* den_out, rsis_stream_t stream);
/* dlogits[r][p] = dcost[r] * d cost[r] / d logits[r][p] for row-wise softIoU (logits, gt: [rows][hw]). */
int rsis_soft_iou_bwd(const float* logits, const void* gt, int gt_is_u8, int rows, int64_t hw, const float* num,
                      const float* den, const float* dcost, float weight, float* dlogits, rsis_stream_t stream);

/* Masked class / stop losses (utils/objectives.py:6-25 over utils/hungarian.py:10-59; train.py:159-168).  A row is
 * selected when (uint8)sw != 0, the reference's `sw.byte()` mask.
 * rsis_masked_nll_fwd (MaskedNLLLoss / MaskedNLL): cost[r] = -balance[target[r]] * log(probs[r][target[r]]) (balance
 * optional, [num_classes]); probs float32 [rows][num_classes] dense, target int64 [rows], sw float32 [rows].
 * cost_rows (optional, [rows]): the per-row costs, 0 on unselected rows; sum_count [2] = {sum of the selected costs,
 * number of selected rows} (overwritten) -- the mean of train.py:161 without masked_select.
 * rsis_masked_nll_bwd: dprobs [rows][num_classes] (overwritten) for upstream dcost[r * dcost_stride] (stride 0 = one
 * scalar for every row, e.g. dloss / count).
 * rsis_masked_bce_fwd (MaskedBCELoss / StableBalancedMaskedBCE): target, logits, sw float32 [n]; balance_weight < 0
 * means "None" (sum(target) / n, computed in the kernel); sum_count_bw [3] = {sum, count, balance weight used}.
 * rsis_masked_bce_bwd: balance_weight points at the weight the forward used (sum_count_bw + 2). */
int rsis_masked_nll_fwd(const float* probs, const int64_t* target, const float* sw, const float* balance, int rows,
                        int num_classes, float* cost_rows, float* sum_count, rsis_stream_t stream);
int rsis_masked_nll_bwd(const float* probs, const int64_t* target, const float* sw, const float* balance,
                        const float* dcost, int64_t dcost_stride, int rows, int num_classes, float* dprobs,
                        rsis_stream_t stream);
int rsis_masked_bce_fwd(const float* target, const float* logits, const float* sw, float balance_weight, int64_t n,
                        float* cost_rows, float* sum_count_bw, rsis_stream_t stream);
int rsis_masked_bce_bwd(const float* target, const float* logits, const float* sw, const float* balance_weight,
                        const float* dcost, int64_t dcost_stride, int64_t n, float* dlogits, rsis_stream_t stream);

/* Hungarian matching on the device (replaces `Munkres().compute` per image on the host,
 * utils/hungarian.py:91-125, train.py:137).  cost: float32 [b][rows][cols] with element strides (rows = ground-truth
 * objects, cols = predictions; both <= 32).  perm[b][col] = row matched to column col for col < min(rows, cols), 0
 * elsewhere (perm_len entries per image; the reference's `permute_indices`); total_cost (optional) [b] = the optimal
 * assignment cost.  Minimum-cost assignment of the smaller side, i.e. what Munkres returns for the zero-padded
 * square matrix; when several assignments are optimal (equal costs) any one of them is returned. */
int rsis_hungarian_match(const float* cost, int64_t stride_b, int64_t stride_r, int64_t stride_c, int b, int rows,
                         int cols, int32_t* perm, int perm_len, float* total_cost, rsis_stream_t stream);

/* ---- evaluation post-processing ------------------------------------------------------ */
/* eval.py:97-127 per predicted instance, on the device: segmentation = masks > threshold (`(pred_mask > th)`), zeroed
 * where ignore == 1 (`segmentation[ignore_pixels==1] = 0`; ignore optional, uint8 [n][h][w]), areas[i] =
 * sum(segmentation) (the `min_size` filter of eval.py:119), and the COLUMN-MAJOR run-length encoding of
 * coco/common/maskApi.c:32-41 (`rleEncode`, reached through `mask.encode(np.asfortranarray(...))`): counts[i][0] =
 * number of leading zeros (0 when the mask starts with a one), then alternating run lengths; n_runs[i] runs (when
 * n_runs[i] > max_runs only the first max_runs counts were written).  masks: float32 [n][h][w] row-major.
 * workspace: rsis_rle_workspace_bytes(n, h, w) bytes, 16-byte aligned (no initialisation needed). */
size_t rsis_rle_workspace_bytes(int n, int h, int w);
int rsis_rle_encode(const float* masks, float threshold, const uint8_t* ignore, int n, int h, int w, void* workspace,
                    uint32_t* counts, int max_runs, int32_t* n_runs, uint32_t* areas, rsis_stream_t stream);

/* ---- optimiser step ------------------------------------------------------------------ */
/* torch.optim.Adam as utils/utils.py:72-83 builds it (train.py:185-187), on n contiguous float32 elements: per element,
 * `repeats` sequential updates with step counts step0+1 .. step0+repeats (a parameter listed k times in an optimiser --
 * utils/utils.py:34-52 does that -- is updated k times per step()):
 *   g' = g + weight_decay*p;  m += (g' - m)*(1 - beta1);  v = v*beta2 + (1 - beta2)*g'^2;
 *   p -= lr/(1 - beta1^t) * m / (sqrt(v)/sqrt(1 - beta2^t) + eps).   repeats <= 8. */
int rsis_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int64_t step0, int repeats, rsis_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RSIS_B200_H_ */
