/*
 * hypelcnn_b200 — C ABI of the B200-native HSI+LiDAR patch engine.
 *
 * The reference (aligokalppeker/hypelcnn) has no FFI: its plug-in boundary is four Python
 * ABCs resolved by name (nnmodel/NNModel.py:4-12, importer/DataImporter.py:4-20,
 * loader/DataLoader.py:5-47, common/common_nn_ops.py:23-42).  The Python classes in
 * hypelcnn_b200/{nnmodel,importer,loader,common} keep those interfaces verbatim and call
 * THIS library through ctypes; each entry point below names the reference code it
 * replaces.  See INTEGRATION.md for the reference-side binding.
 *
 * Conventions
 *  - every function returns 0 (HYP_OK) or a negative HYP_E_* code; never throws/aborts;
 *    hyp_last_error() gives a thread-local message for the last failure on this thread.
 *  - all data pointers are DEVICE pointers owned by the caller (torch.Tensor.data_ptr()),
 *    fp32, NHWC, C-contiguous; `stream` is a cudaStream_t passed as void*.
 *  - calls are asynchronous with respect to the host on `stream`.
 *  - a hyp_model is bound to one device and used from one stream at a time.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *    with HYP_E_CUDA.
 */
#ifndef HYPELCNN_B200_H
#define HYPELCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYP_ABI_VERSION 1

enum {
  HYP_OK = 0,
  HYP_E_INVALID = -1,     /* bad argument (shape, null pointer, range) */
  HYP_E_CUDA = -2,        /* CUDA runtime error (message carries cudaGetErrorString) */
  HYP_E_STATE = -3,       /* call order violated (e.g. backward without training forward) */
  HYP_E_UNSUPPORTED = -4  /* valid request this build does not implement */
};

typedef struct hyp_model hyp_model;

int hyp_version(void);
const char* hyp_last_error(void);

/* ---------------------------------------------------------------------------------------
 * Scene preparation + patch gather.
 * Replaces BasicDataSet.__init__ normalisation (common/common_nn_ops.py:62-78),
 * get_data_point_func (common/common_nn_ops.py:169-176), GRSS2018DataSet.get_data_point
 * (loader/GRSS2018DataLoader.py:10-44) and the per-pixel Python loop of
 * InMemoryImporter._get_data_with_labels (importer/InMemoryImporter.py:27-38).
 * The scene stays UNPADDED and un-normalised in HBM; numpy's "symmetric" padding
 * (common/common_nn_ops.py:55-60) is applied as index reflection inside the kernel.
 * ------------------------------------------------------------------------------------- */
enum { HYP_DT_F32 = 0, HYP_DT_U16 = 1 };
enum { HYP_GATHER_SAME_RES = 0, HYP_GATHER_GRSS2018 = 1 };

/* per-band min and max(value - min) over an [H,W,C] cube -> min_out[C], max_out[C] (float).
 * max_out is the maximum of the SHIFTED data, exactly as the reference computes it. */
int hyp_scene_minmax(const void* cube, int dtype, int H, int W, int C,
                     float* min_out, float* max_out, void* stream);

/* out[n, py, px, 0:C_hsi] = (casi[ry, rx, :] - casi_min) / casi_max   (IEEE division)
 * out[n, py, px, C_hsi]   = (lidar[ly, lx] - lidar_min) / lidar_max   (if lidar != NULL)
 * casi_min/casi_max NULL -> no normalisation.  targets_xy: int32 [N,2] = (x=column, y=row)
 * in scene coordinates (LiDAR resolution in GRSS2018 mode).  out row stride = out_ld floats
 * per pixel (>= C_hsi + has_lidar), extra channels are written as 0. */
int hyp_gather_patches(const void* casi, int casi_dtype, int Hc, int Wc, int C_hsi,
                       const float* casi_min, const float* casi_max,
                       const float* lidar, int Hl, int Wl,
                       const float* lidar_minmax /* device float[2] or NULL */,
                       int neighborhood, int mode,
                       const int32_t* targets_xy, int64_t N,
                       float* out, int out_ld, void* stream);

/* ---------------------------------------------------------------------------------------
 * NNModel engine.  Replaces the TF graph built by HYPELCNNModel.create_tensor_graph
 * (nnmodel/HYPELCNNModel.py:34-99), get_loss_func (:101-112) and optimize_nn
 * (common/common_nn_ops.py:208-240).
 * ------------------------------------------------------------------------------------- */
enum {
  HYP_MODEL_HYPELCNN = 0, /* nnmodel/HYPELCNNModel.py */
  HYP_MODEL_DUALCNN = 1,  /* nnmodel/DUALCNNModel.py:11-104 (tensor-core engine only) */
  HYP_MODEL_CONCNN = 2    /* nnmodel/CONCNNModel.py:23-64 (tensor-core engine only; pass lrelu_alpha = 0 for ReLU) */
};
enum {
  HYP_PRECISION_FP32 = 0,    /* fp32 FFMA everywhere (parity mode) */
  HYP_PRECISION_3XTF32 = 1,  /* tcgen05 kind::tf32, hi/lo split, fp32-accurate */
  HYP_PRECISION_BF16 = 2,    /* tcgen05 kind::f16, one bf16 plane per operand, fp32 accumulate: the fast mode */
  HYP_PRECISION_3XF16 = 3    /* tcgen05 kind::f16, fp16 hi/lo split (3 MMAs per K step of 16), fp32-accurate at twice
                                the 3xTF32 rate; operands are power-of-two scaled into the fp16 range */
};

typedef struct hyp_model_desc {
  int32_t kind;             /* HYP_MODEL_* */
  int32_t patch;            /* P: window edge = 2*neighborhood+1 */
  int32_t channels;         /* C_in (HSI bands + LiDAR) */
  int32_t classes;
  int32_t filter_count;     /* alg_param "filter_count" */
  int32_t spectral_levels;  /* "spectral_hierarchy_level" */
  int32_t spatial_levels;   /* "spatial_hierarchy_level" */
  int32_t degradation;      /* "degradation_coeff" */
  int32_t use_residual;     /* "use_residual" */
  int32_t precision_mode;   /* HYP_PRECISION_* */
  int32_t max_batch;        /* largest B any call will pass */
  int32_t reserved;         /* DUALCNN: "hs_lidar_diff" (pixels cropped from each side of the HSI window); else 0 */
  float lrelu_alpha;        /* "lrelu_alpha" */
  float bn_decay;           /* "bn_decay" */
  float bn_eps;             /* slim default 0.001 */
  float drop_out_ratio;     /* "drop_out_ratio"; HYPELCNN: keep_prob = 1 - ratio (HYPELCNNModel.py:123);
                               DUALCNN: keep_prob = ratio (slim dropout's positional keep_prob, DUALCNNModel.py:49) */
} hyp_model_desc;

int hyp_model_create(const hyp_model_desc* desc, hyp_model** out);
void hyp_model_destroy(hyp_model* m);

/* element counts of the flat buffers the caller must allocate (fp32 elements / bytes) */
int hyp_model_sizes(const hyp_model* m, int64_t* n_params, int64_t* n_bn_state,
                    int64_t* workspace_bytes, int32_t* n_variables);

/* variable table; names are the reference's TF checkpoint names, e.g.
 * "nn_core/conv_enc_0/weights" [1,1,145,120], "nn_core/fc_0/BatchNorm/moving_variance".
 * kind: 0 weights, 1 beta (both in `params`/`grads`), 2 moving_mean, 3 moving_variance
 * (both in `bn_state`).  offset is in fp32 elements inside the respective buffer. */
int hyp_model_variable(const hyp_model* m, int idx, char name[128], int32_t* kind,
                       int64_t* offset, int32_t shape[4], int32_t* rank);

int hyp_model_bind(hyp_model* m, float* params, float* grads, float* bn_state,
                   void* workspace, size_t workspace_bytes);

/* forward.  is_training: BN uses batch statistics, dropout active (Philox keyed by
 * dropout_seed), decoder branch runs; update_moving: also update BN moving statistics.
 * logits [B,classes]; recon [B,P*P*C] (nullable; only produced when is_training). */
int hyp_model_forward(hyp_model* m, const float* x, int64_t B, int is_training,
                      int update_moving, uint64_t dropout_seed,
                      float* logits, float* recon, void* stream);

/* per-sample loss of get_loss_func: CE_i (+ scalar reconstruction MSE when recon != NULL).
 * labels: uint8 class ids [B] (the argmax of the reference's one-hot rows). */
int hyp_model_loss(hyp_model* m, const float* logits, const float* recon, const float* x,
                   const uint8_t* labels, int64_t B, float* per_sample_loss, void* stream);

/* loss = mean_B(per-sample loss) and d loss / d params into `grads`, for the batch of the
 * immediately preceding hyp_model_forward(is_training=1) call (same x, same B).
 * loss_out: device float[3] = {total, mean CE, reconstruction MSE}. */
int hyp_model_loss_backward(hyp_model* m, const float* x, const uint8_t* labels, int64_t B,
                            float* loss_out, void* stream);

/* Overlap of the gradient all-reduce with the rest of backward (multi-GPU).  Parameters are laid
 * out in layer order and backward walks the layers last to first, so the tail [param_offset, n_params) of the
 * flat gradient buffer is final as soon as the first layer at or above param_offset has been processed:
 * hyp_model_loss_backward then records `event` (a cudaEvent_t) on its stream.  Up to 8 events with different offsets
 * can be registered (one call each; calling again with a registered event moves its offset): the buffer is then reduced
 * in as many pieces, each as soon as it is final.  event == NULL forgets all registrations. */
int hyp_model_set_grad_notify(hyp_model* m, int64_t param_offset, void* event);

/* TF1 AdamOptimizer (ApplyAdam): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1);
 * v += (g*g-v)(1-b2); p -= lr_t*m/(sqrt(v)+eps).  g = grads*grad_scale (1/world_size). */
int hyp_adam_step(float* params, const float* grads, float* m, float* v, int64_t n,
                  float lr, float b1, float b2, float eps, int64_t t, float grad_scale,
                  void* stream);

/* tf MomentumOptimizer (optimize_nn's second branch, common/common_nn_ops.py:223-227, use_nesterov = False):
 * accum = momentum * accum + g;  p -= lr * accum.  g = grads*grad_scale. */
int hyp_momentum_step(float* params, const float* grads, float* accum, int64_t n, float lr, float momentum,
                      float grad_scale, void* stream);

/* Training-time augmentation of a batch [B,P,P,C] (common/common_nn_ops.py:397-440, pinned to /cpu:0 and run per
 * sample inside tf.data in the reference): rot90 by k in {0,1,2} (tf.image.rot90, counter-clockwise), random
 * left-right / up-down flips, per-channel spectral offset U(-spectral, 0) — one draw per sample, Philox keyed by
 * (seed, sample).  choices_out (nullable) uint8 [B,4] = (k, flip_lr, flip_ud, 0) and deltas_out (nullable) float
 * [B,C] receive the draw.  in and out must not alias. */
int hyp_augment_patches(const float* in, float* out, int64_t B, int patch, int channels, int do_rotation,
                        int do_reflection, float spectral, uint64_t seed, uint8_t* choices_out, float* deltas_out,
                        void* stream);

/* Shadow GAN generator, forward only (gan/shadow_data_models.py:43-90; the frozen augmenter of
 * gan/gan_utilities.py:18-43 and create_inference_for_matrix_input, gan/wrappers/gan_common.py:282-304).
 * rows spectra of `bands` channels at row strides ld_in / ld_out; `copy_extra` channels after the bands (the LiDAR
 * channel of a patch pixel) are copied through.  weights: net1 w[K1], b, net2 w[K2], b, ... (K = bands, /2, /4, /8,
 * /4, /2, bands; 4 layers when encoder_only).  clip_invalid_values: keep the input spectrum unless the generated mean
 * is lower (is_shadow_graph) / higher than the input mean. */
int hyp_gan_generator_forward(const float* in, int ld_in, float* out, int ld_out, int64_t rows, int bands,
                              int copy_extra, const float* weights, int encoder_only, int clip_invalid_values,
                              int is_shadow_graph, void* stream);

/* ---- CycleGAN training on [rows, bands] spectra (gan/wrappers/cycle_gan_wrapper.py:48-116,189-333 over the models
 * of gan/shadow_data_models.py:43-123).  The host (hypelcnn_b200/gan/wrappers/cycle_gan_wrapper.py) chains these. ----
 * generator training forward: nets [rows][8][bands] = net0 (the input) .. net7 (the output), kept for backward */
int hyp_gan_generator_train_forward(const float* x, int64_t rows, int bands, const float* weights, float* nets,
                                    void* stream);
/* generator backward: gout = dL/dnet7 [rows,bands] -> gin = dL/dnet0 (nullable), gweights += dL/d(weights, biases) */
int hyp_gan_generator_backward(const float* nets, const float* gout, int64_t rows, int bands, const float* weights,
                               float* gin, float* gweights, void* stream);
/* discriminator (FC C->C, C->C leaky_relu 0.1, C->C/2 linear; weights = W1 [C,C], b1, W2, b2, W3 [C,C/2], b3):
 * forward keeps the hidden activations [rows][2][bands]; backward gives dL/dx (gin, nullable) and/or gweights += */
int hyp_gan_discriminator_forward(const float* x, int64_t rows, int bands, const float* weights, float* hidden,
                                  float* out, void* stream);
int hyp_gan_discriminator_backward(const float* x, const float* hidden, const float* gout, int64_t rows, int bands,
                                   const float* weights, float* gin, float* gweights, void* stream);
/* loss terms + gradients.  mode 0 (tfgan least_squares_*): loss_acc += sum 0.5*scale*(a-target)^2, grad = scale*(a-target);
 * mode 1 (absolute_difference): loss_acc += sum scale*|a-b|, grad = scale*sign(a-b);  mode 2 (tfgan wasserstein_*):
 * loss_acc += sum scale*a, grad = scale (the sign rides on scale).  scale = weight / numel (means).
 * grad nullable; accumulate != 0 adds to it.  loss_acc: device double, nullable. */
int hyp_gan_loss_grad(int mode, const float* a, const float* b, float target, float scale, int64_t numel, float* grad,
                      int accumulate, double* loss_acc, void* stream);
/* slim l2_regularizer(scale) on a weight range: loss_acc += scale * sum(w^2)/2, grads += scale * w (grads nullable) */
int hyp_gan_l2_regularizer(const float* weights, float* grads, int64_t n, float scale, double* loss_acc, void* stream);

/* ---- fused CycleGAN train ops (gan/wrappers/cycle_gan_wrapper.py:189-333 over gan/shadow_data_models.py:43-123): ONE kernel per
 * train op instead of a chain of per-op launches (a train iteration at the reference's batch of 32 is launch latency).
 * Flat weight buffers as in the per-op entry points above; bands a multiple of 8 in 8..64.
 *
 * generator step: grad_g / grad_f += d L_G / d [G (x -> y) | F (y -> x)] with
 *   L_G = 0.5 mean (D_Y(G x) - 1)^2 + 0.5 mean (D_X(F y) - 1)^2                       -> loss_acc[1]
 *       + cycle_weight    * (mean |F(G x) - x| + mean |G(F y) - y|)                   -> loss_acc[2]
 *       + 2 identity_weight * (mean |G(x) - x| + mean |F(y) - y|)                     -> loss_acc[3]
 * (tfgan counts the cycle term once per partial model with weight / 2 each; the "identity" terms apply each generator
 * to its own domain's input, as the reference does).  The discriminators are frozen.  gen_y = G(x), gen_x = F(y),
 * rec_x = F(G x), rec_y = G(F y): nullable [rows][bands] outputs.  loss_acc: device double[4], += ; [0] gets the
 * total of both entry points' terms. */
int hyp_gan_cycle_generator_step(const float* x, const float* y, int64_t rows, int bands, const float* w_g, const float* w_f,
                                 const float* w_dy, const float* w_dx, float cycle_weight, float identity_weight,
                                 float* grad_g, float* grad_f, double* loss_acc, float* gen_y, float* gen_x, float* rec_x,
                                 float* rec_y, void* stream);
/* discriminator step: grad_dy / grad_dx += d L_D / d [D_Y | D_X] with
 *   L_D = sum over both domains of 0.5 mean (D(real) - 1)^2 + 0.5 mean D(fake)^2      -> loss_acc[1]
 *       + reg_scale * sum w^2 / 2 over the two hidden layers' weight matrices          -> loss_acc[2]
 * fake = G(x) / F(y) computed here with the current generators, passed through tfgan.features.tensor_pool kept on the
 * device: pool_y / pool_x [slots][rows][bands]; pool_mode 0 = no pool, 1 = store the fresh fakes into pool_slot and use
 * them (pool still filling), 2 = use what pool_slot holds and replace it by the fresh fakes.  The caller draws mode and
 * slot (host RNG, as the reference's pool does). */
int hyp_gan_cycle_discriminator_step(const float* x, const float* y, int64_t rows, int bands, const float* w_g,
                                     const float* w_f, const float* w_dy, const float* w_dx, float reg_scale, float* grad_dy,
                                     float* grad_dx, double* loss_acc, float* pool_y, int pool_mode_y, int pool_slot_y,
                                     float* pool_x, int pool_mode_x, int pool_slot_x, void* stream);

/* ---- CUT / DCLGAN / DCL-CycleGAN (gan/wrappers/cut_wrapper.py:256-420, dcl_gan_wrapper.py, dcl_cycle_gan_wrapper.py).
 * generator backward that also takes the gradient of the ENCODER output net4 (generator_fn(create_only_encoder=True),
 * gan/shadow_data_models.py:43-75): gout (dL/dnet7) nullable -> only net1..net4 are differentiated; gout_enc nullable */
int hyp_gan_generator_backward_enc(const float* nets, const float* gout, const float* gout_enc, int64_t rows, int bands,
                                   const float* weights, float* gin, float* gweights, void* stream);
/* patch feature discriminator (gan/shadow_data_models.py:126-149): slices of bands / patch_count bands (last one
 * ragged), per slice FC in->ps->ps/4->ps/2->E with leaky_relu(0.1) after each; weights per slice (padded to the full
 * slice size): W1 [in,ps] b1 W2 b2 W3 b3 W4 [ps/2,E] b4.  forward: z [rows][slices][E] un-normalised, sumsq [slices] =
 * sum over the batch of z^2 (tf.math.l2_normalize over the whole [rows,E] tensor is applied by the consumers).
 * backward: gf = dL/d(normalised embedding), dot [slices] = sum gf*z (from hyp_gan_patchnce) -> gin [rows,bands]
 * (nullable) and gweights += (nullable). */
int64_t hyp_gan_feature_discriminator_weight_count(int bands, int patch_count, int embedded_feature_size);
int hyp_gan_feature_discriminator_forward(const float* x, int64_t rows, int bands, int patch_count,
                                          int embedded_feature_size, const float* weights, float* z, float* sumsq,
                                          void* stream);
int hyp_gan_feature_discriminator_backward(const float* x, const float* z, const float* sumsq, const float* gf,
                                           const float* dot, int64_t rows, int bands, int patch_count,
                                           int embedded_feature_size, const float* weights, float* gin, float* gweights,
                                           void* stream);
/* PatchNCE (cut_wrapper.py:360-420): logits = f_gen f_real^T / tau per sample, labels eye(slices) flattened, one
 * softmax over slices^2; loss_acc += scale * sum_b loss_b (scale = weight / rows).  g_gen / g_real (nullable pair) =
 * dL/d f; fused_grad 1 = TensorFlow's fused xent gradient (softmax - labels), 0 = exact (slices*softmax - labels). */
int hyp_gan_patchnce(const float* z_gen, const float* z_real, const float* sumsq_gen, const float* sumsq_real,
                     int64_t rows, int slices, int embedded_feature_size, float tau, float scale, int fused_grad,
                     float* g_gen, float* g_real, float* dot_gen, float* dot_real, double* loss_acc, void* stream);

/* tf.argmax (lowest index on ties) + tf.math.confusion_matrix accumulation
 * (common/common_nn_ops.py:246-262, :318).  labels/confusion nullable.
 * confusion: int32 [classes,classes], rows = labels, += semantics. */
int hyp_argmax_confusion(const float* logits, const uint8_t* labels, int64_t B, int classes,
                         uint8_t* pred, int32_t* confusion, void* stream);

/* perform_prediction's scatter (common/common_nn_ops.py:320-322): map[y*W + x] = pred. */
int hyp_scatter_class_map(const uint8_t* pred, const int32_t* targets_xy, int64_t N,
                          int H, int W, uint8_t* class_map, void* stream);

/* ---- introspection used by the parity tests -------------------------------------------- */
/* device pointer + element count of an internal tensor of the last forward/backward.
 * what: 0 activation (post BN/act/residual) of tensor `name`, 1 pre-BN output of layer
 * `name`, 2 gradient w.r.t. activation tensor `name`, 3 BatchNorm batch mean [Cout] of layer
 * `name`.  Tensors come back dense, [B][P*P][C]. */
int hyp_model_debug_tensor(hyp_model* m, const char* name, int what, float** ptr, int64_t* numel);
/* the 0/1 keep mask hyp_model_forward applies on dropout layer `layer_scope` for `seed` */
int hyp_model_dropout_mask(hyp_model* m, const char* layer_scope, uint64_t seed, int64_t B,
                           uint8_t* mask_out, void* stream);
/* Host only: CRC-32C (Castagnoli) update, *crc_inout = crc32c(previous *crc_inout, data[0..len)); start from 0.
 * The checksum of the TFRecord framing TensorFlow writes for importer/TFRecordImporter.py:16-72 and
 * utilities/tfrecord_writer.py:45-81 (masked: ((crc >> 15) | (crc << 17)) + 0xa282ead8). */
int hyp_crc32c(const void* data, uint64_t len, uint32_t* crc_inout);
/* Host only: decode one TIFF LZW strip / tile (compression 5) into out[0..out_capacity); *out_len = bytes written.
 * Replaces tifffile's decoder behind the reference's scene reads (loader/GRSS2013DataLoader.py, GRSS2018DataLoader.py:53-67,
 * GULFPORTDataLoader.py:22-43: `from tifffile import imread`). */
int hyp_tiff_lzw_decode(const void* data, uint64_t len, void* out, uint64_t out_capacity, uint64_t* out_len);

/* Debug / test hook, host only (no device needed): the static tile schedule of the persistent GEMM kernel
 * (hyp_tc_engine.cuh schedule_tiles) applied to a plain cost vector.  group_of_unit[u] = CTA group that runs unit u,
 * rank_in_group[u] = its position in that group's execution order.  windowed != 0: locality windows of 2*groups units
 * with load-aware dealing (K-major and level-wgrad launches); 0: plain longest-processing-time-first. */
int hyp_debug_schedule(const double* costs, int units, int groups, int windowed, int32_t* group_of_unit,
                       int32_t* rank_in_group);

/* Debug / test hook, host only: the tap groups of a level's wgrad launch (hyp_tc_engine.cuh level_tap_groups; level =
 * R kernels 1x1 .. (2R-1)x(2R-1) of nt x fpad filter columns each on P x P patches) and, per group and source
 * (activation) position q the launch visits, the gz position each of the group's taps reads.  Rows of out (6 int32):
 * group, set, dy, dx, q, output position (255 = outside the patch: that set reads zeros).  rows_out = rows produced
 * (rows beyond cap_rows are counted, not written). */
int hyp_debug_level_tap_groups(int P, int R, int nt, int fpad, int max_sets, int32_t* out, int cap_rows, int* rows_out);

/* probe of the tcgen05/TMA segment-GEMM building block used by the tensor-core precision
 * modes.  mn bit 0 = 0: A[M,K], B[N,K] -> D = A*B^T (K-major); 1: A[K,M], B[K,N]
 * -> D = A^T*B (MN-major).  mn bit 1: run as CTA pairs (tcgen05 cta_group::2).  mn bits 2..3: operand format
 * (0: two fp32 planes, 3 kind::tf32 MMAs; 1: two fp16 planes, 3 kind::f16 MMAs; 2: one bf16 plane).
 * stats (nullable): [ceil(M/128)][2][N] per-tile column sums / sums of squares. */
int hyp_debug_tc_gemm(int mn, const float* A, const float* B, int M, int N, int K, float* D,
                      float* stats, int raw_hi, int bn, int ksplit, int chunk_kb, void* stream);
/* number of kernels launched by this library on this thread since the last reset */
int64_t hyp_launch_count(int reset);
/* per-kernel-class timing with CUDA events around every launch (off by default; bench.py
 * turns it on for the timed region).  enable(1) clears previous records.  get(idx) sums the
 * records of kernel tag idx (synchronises the device); returns HYP_E_INVALID past the end.
 * flops = useful algorithmic FLOPs (GEMM kernels), bytes = algorithmic bytes (HBM kernels). */
int hyp_profile_enable(int on);
int hyp_profile_get(int idx, char name[64], double* total_ms, int64_t* launches, double* flops,
                    double* bytes);

#ifdef __cplusplus
}
#endif
#endif /* HYPELCNN_B200_H */
