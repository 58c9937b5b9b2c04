/*
 * mvae_b200 - C ABI of the B200-native MVAE training-step library (libmvae_b200.so).
 *
 * The reference (wenxuanliu/multimodal-vae) has no FFI layer: its boundary is the Python module
 * surface of mnist/model.py and mnist/train.py.  Each entry below replaces the ATen work behind
 * one part of that surface; the reference lines it stands in for are cited per entry.
 *
 * Conventions
 *   - every function returns 0 on success; on failure mvae_last_error() (thread-local) explains;
 *   - all tensor pointers are DEVICE pointers, row-major, owned by the caller (the library never
 *     allocates or frees tensor memory); workspaces are passed in by the caller;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*) of the CURRENT device and
 *     is CUDA-graph capturable; nothing synchronises;
 *   - dtype codes: MVAE_DT_F32 = fp32 storage, tensor cores run kind::tf32;
 *                  MVAE_DT_BF16 = bf16 storage, tensor cores run kind::f16 (bf16), fp32 accumulate.
 */
#ifndef MVAE_B200_H_
#define MVAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVAE_ABI_VERSION 5

#define MVAE_DT_F32 0
#define MVAE_DT_BF16 1
/* fp32 storage, error-compensated 3xTF32 tensor-core GEMMs: every operand is split into hi = tf32(x) and
 * lo = x - hi and the product is hi*hi + lo*hi + hi*lo in fp32 TMEM accumulators (fp32-grade results; the parity
 * mode that meets "rtol 1e-3 on every gradient" against the fp32 reference).  Accepted by mvae_mnist_sizes /
 * mvae_mnist_step (dtype) and by mvae_gemm (dtype, with x3_scratch). */
#define MVAE_DT_F32X3 2

/* PoE arithmetic:
 *   REF       - bit-for-bit the reference: var=exp(logvar)+eps, mu=sum(mu*var)/sum(var)
 *               (variance-weighted!), var=1/sum(1/var), no prior expert  (mnist/model.py:180-185)
 *   PRECISION - the paper's precision-weighted product, optional N(0,1) prior expert          */
#define MVAE_POE_REF 0
#define MVAE_POE_PRECISION 1

const char* mvae_last_error(void);
int mvae_abi_version(void);
/* Fails unless `device` is an sm_100 part (there is no fallback path). */
int mvae_device_check(int device);

/* ------------------------------------------------------------------------------------------
 * Generic tensor-core GEMM  C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias[N])
 * replaces: every nn.Linear forward / dgrad / wgrad on the path
 *           (mnist/model.py:104-110, 124-130, 146, 162-165 and their autograd backward).
 * a_major / b_major: 0 = contraction index contiguous ([rows, K] row-major);
 *                    1 = row index contiguous (the operand is stored as [K, rows] row-major).
 * accumulate != 0  : C (fp32) += result, split-K allowed (weight gradients).
 * col_sum/col_sumsq: optional [groups, N] fp32 accumulators (+= with atomics) of the stored
 *                    values per column, group = row / rows_per_group (BatchNorm batch statistics).
 */
typedef struct mvae_gemm_args {
  int dtype;          /* MVAE_DT_* of A and B */
  int M, N, K;
  const void* A; int64_t lda; int a_major;
  const void* B; int64_t ldb; int b_major;
  void* C; int64_t ldc; int c_dtype;
  const float* bias;
  int accumulate;
  float* col_sum;
  float* col_sumsq;
  int rows_per_group; /* <= 0: one group */
  int block_n;        /* 0 = auto */
  int split_k;        /* 0 = auto */
  int stages;         /* 0 = auto */
  void* debug_times;  /* NULL, or device int64 [ctas][8] receiving %globaltimer stamps (bring-up only) */
  void* x3_scratch;   /* MVAE_DT_F32X3 only: device scratch for the split operands, >= 12 * (M + N + 8) * (K + 4) bytes */
  int64_t x3_scratch_bytes;
  /* Activation fused into the epilogue (ABI 5; the north-star's Linear+Swish stacks).  act = MVAE_ACT_SWISH with
   *   act_out != NULL: C = A B^T + bias (the pre-activation, kept for the backward; C may be NULL) and
   *                    act_out = swish(C), same dtype / leading dimension as C            (forward of Linear + Swish)
   *   act_pre != NULL: C = (A B^T) * swish'(act_pre) and col_sum[n] += sum_m C[m, n]      (input gradient arriving at the
   *                    previous Linear's pre-activation act_pre [M, N], dtype of A; col_sum = that Linear's bias gradient)
   * act = MVAE_ACT_NONE (0): plain GEMM as above. */
  int act;
  void* act_out;
  const void* act_pre; int64_t ld_act_pre;
  /* sigmoid + binary cross entropy fused into the epilogue (ABI 5; mnist/model.py:135 + mnist/train.py:70 - the logits never
   * reach memory).  bce_target != NULL ([bce_target_rows, N], dtype of A; row m compares with bce_target[m % bce_target_rows]):
   *   x = A B^T + bias;  bce_loss[g] += bce_scale[g] * sum over the rows of group g = m / rows_per_group (<= 4 groups) of
   *   softplus(x) - t x  (= BCE(sigmoid(x), t));
   *   C = bce_scale[g] * (sigmoid(x) - t)  (the gradient at the logits);  col_sum[n] += sum_m C[m, n]  (one group: the bias
   *   gradient);  bce_probs (optional, layout / dtype of C) = sigmoid(x). */
  const void* bce_target; int64_t ld_bce_target; int bce_target_rows;
  float bce_scale[4];
  float* bce_loss;
  void* bce_probs;
  const float* bce_row_weight; /* optional [M] fp32: row m's loss and gradient are scaled by bce_row_weight[m] (per-sample
                                  missing-modality masks: 0 switches the row off, B / present-rows renormalises the mean) */
} mvae_gemm_args;
int mvae_gemm(const mvae_gemm_args* args, void* stream);
/* Bring-up: GEMM launches whose epilogue kind (0 store, 1 atomic, 2 BCE, 3 dgrad-BN) equals epilogue_kind write
 * eight %globaltimer stamps per CTA into the device buffer (NULL switches it off). */
void mvae_debug_gemm_times(void* device_int64_buffer, int epilogue_kind);
/* Bring-up: the slab-persistent chain kernels of the MNIST step (csrc/chain.cu) write %globaltimer stamps of their
 * producer / MMA / epilogue phases into a device buffer of 4 kernels x 148 CTAs x 32 int64 (NULL switches it off). */
void mvae_debug_chain_times(void* device_int64_buffer);

/* ------------------------------------------------------------------------------------------
 * MNIST MVAE (mnist/model.py:14-170).  All parameters live in ONE flat fp32 buffer (so that the data-
 * parallel gradient all-reduce and Adam are single flat operations); float BatchNorm buffers live in a
 * second flat buffer; num_batches_tracked is an int64[6] array in state_dict order.  The table below is
 * queried by the host to build views named exactly like the reference's state_dict keys.
 */
typedef struct mvae_tensor_info {
  char name[64];   /* reference state_dict key, e.g. "image_encoder.net.0.weight" */
  int kind;        /* 0 parameter (offset into params/grads/adam buffers, in floats)
                      1 float buffer: running_mean / running_var (offset into the buffer array)
                      2 num_batches_tracked (index into the int64 array) */
  int ndim;
  int64_t shape[2];
  int64_t offset;
} mvae_tensor_info;
int mvae_mnist_num_tensors(void);
int mvae_mnist_tensor_info(int n_latents, int index, mvae_tensor_info* out);

typedef struct mvae_mnist_size_info {
  int64_t param_floats;     /* length of the flat params / grads / adam_m / adam_v (/ bf16 mirror) buffers */
  int64_t encoder_param_floats; /* [0, this) = encoder parameters, [this, param_floats) = decoder parameters */
  int64_t buffer_floats;    /* length of the flat running-statistics buffer */
  int64_t num_bn;           /* length of the num_batches_tracked array */
  int64_t workspace_bytes;  /* activation workspace for (batch, n_latents, dtype), up to 3 terms */
} mvae_mnist_size_info;
int mvae_mnist_sizes(int n_latents, int batch, int dtype, mvae_mnist_size_info* out);

#define MVAE_TERM_JOINT 0 /* vae(image, text)  mnist/train.py:136 */
#define MVAE_TERM_IMAGE 1 /* vae(image=image)  mnist/train.py:137 */
#define MVAE_TERM_TEXT 2  /* vae(text=text)    mnist/train.py:138 */

/* One training step = mnist/train.py:132-153: zero_grad, the (up to) three forwards, the three
 * loss_function calls (mnist/train.py:64-81) summed, backward, optimizer.step().
 *   loss_t = lambda_image[t] * BCE_mean(recon_image_t, image) + lambda_text[t] * NLL_mean(recon_text_t, text)
 *            + kl_weight[t] * -0.5 * sum(1 + logvar_t - mu_t^2 - exp(logvar_t))
 * (the reference's MNIST step is kl_weight = 3 / (784 * batch), all lambdas 1; weak supervision
 * mnist/modal_weak.py:76-97 drops terms).  `eps` injects the N(0,1) draws of reparametrize
 * (mnist/model.py:27) for parity; NULL draws them in-kernel (Philox keyed by seed and step). */
typedef struct mvae_mnist_step_args {
  int batch, n_latents, dtype;
  int n_terms;
  int term_type[3];
  float lambda_image[3], lambda_text[3], kl_weight[3];
  int poe_mode, prior_expert;
  float poe_eps;
  const void* image;        /* [batch, 784] in `dtype` storage, values in [0,1] */
  const int64_t* text;      /* [batch] labels 0..9 */
  const float* eps;         /* [n_terms, batch, n_latents] or NULL */
  uint64_t seed;
  float* params;            /* flat fp32 parameters */
  void* params_bf16;        /* bf16 mirror (same element offsets); required for MVAE_DT_BF16 */
  float* buffers;           /* flat running_mean / running_var */
  int64_t* num_batches_tracked; /* [6] or NULL */
  float* grads;             /* flat fp32 gradients (accumulated into) */
  int do_backward, zero_grad, do_adam;
  float* adam_m; float* adam_v;
  int* adam_step;           /* device int, Adam's 1-based bias-correction clock: incremented at the start of a call that
                               opens an optimizer step (do_adam or advance_adam_step), never by forward-only calls */
  float lr, beta1, beta2, adam_eps, grad_scale;
  void* workspace; int64_t workspace_bytes;
  float* out_losses;        /* device [n_terms][4]: total, image BCE, text NLL, KL (already weighted) or NULL */
  void* out_recon_image;    /* optional [n_terms*batch, 784] probabilities, `dtype` storage */
  float* out_recon_text;    /* optional [n_terms*batch, 10] log-probabilities */
  float* out_mu;            /* optional [n_terms, batch, n_latents] */
  float* out_logvar;
  /* --- module-surface extensions (all zero / NULL for the fused training step) --- */
  int eval_mode;            /* 1: vae.eval() - BatchNorm running statistics, z = mu (mnist/model.py:29-30), no updates */
  int phase;                /* 0: forward (+ backward if do_backward); 2: backward only from the upstream gradients
                               below, reusing the workspace of the preceding forward (autograd path of forward());
                               3: forward + decoder-side backward, 4: encoder-side backward (data-parallel overlap:
                               the decoder gradient bucket is all-reduced between the two) */
  const float* z_in;        /* optional [n_terms*batch, n_latents]: decode these latents (decode_image / decode_text,
                               mnist/model.py:35-42), encoders and PoE are skipped */
  const void* d_recon_image;/* phase 2: gradient w.r.t. the recon_image probabilities, `dtype` storage, or NULL */
  const float* d_recon_text;/* phase 2: gradient w.r.t. the recon_text log-probabilities, or NULL */
  const float* d_mu;        /* phase 2: gradient w.r.t. mu [n_terms, batch, n_latents], or NULL */
  const float* d_logvar;
  /* --- ABI v4 --- */
  int* noise_step;          /* device int: the Philox counter of the in-kernel N(0,1) draws; incremented at the start of
                               every forward-type call (training or eval).  NULL: adam_step doubles as the counter */
  int advance_adam_step;    /* 1: this call opens an optimizer step whose Adam update is issued separately (mvae_adam_step
                               after a gradient all-reduce or after several accumulating calls): ++*adam_step */
} mvae_mnist_step_args;
int mvae_mnist_step(const mvae_mnist_step_args* args, void* stream);

/* Measurement aids.  mvae_launch_count: kernels launched by this library so far (monotonic).
 * mvae_mnist_step_profile: runs one step on `stream` only (no side stream) with a CUDA-event pair around every
 * launch; after a stream synchronise returns up to max_entries (label, milliseconds) pairs. */
long long mvae_launch_count(void);
/* Byte offset of a named intermediate buffer inside the step workspace (-1 if unknown); names as in
 * csrc/mnist_step.cu::Plan (h1pre, g2pre, dlog, dy2, dz, denc, ...).  Tests and bring-up only. */
long long mvae_mnist_workspace_offset(const char* name, int batch, int n_latents, int dtype);
int mvae_mnist_step_profile(const mvae_mnist_step_args* args, void* stream, int max_entries, char* labels,
                            int label_stride, float* ms_out, int* n_out);

/* torch.optim.Adam defaults semantics (mnist/train.py:118,153) over a flat buffer; *step_counter is the
 * 1-based step (device memory).  grads are multiplied by grad_scale first (1/world_size after an all-reduce). */
int mvae_adam_step(float* params, float* grads, float* m, float* v, void* params_bf16, int64_t count, float lr,
                   float beta1, float beta2, float eps, const int* step_counter, float grad_scale, int zero_grad,
                   void* stream);
int mvae_cast_f32_to_bf16(const float* in, void* out, int64_t count, void* stream);
/* Data-parallel training: gradient all-reduce over NVLink peer memory fused with the Adam update of one bucket
 * [lo, hi) of the flat buffers, ONE launch per rank (csrc/dp.cu): barrier -> reduce-scatter (peer loads) -> all-gather
 * (peer stores) -> barrier -> Adam(grad_scale).  grads[r] / flags[r] are the peer pointers of rank r's flat fp32 gradient
 * buffer and of its zero-initialised uint32[32] flag array (symmetric memory); params / moments / bf16 mirror are local.
 * A collective: every rank must enqueue the same sequence of calls.  Replaces mnist/train.py:152-153 on one device. */
typedef struct mvae_dp_reduce_adam_args {
  int world, rank;
  void* grads[8];
  void* flags[8];
  float* params;
  float* adam_m;
  float* adam_v;
  void* params_bf16;      /* optional bf16 mirror of params */
  int64_t lo, hi;         /* bucket (elements, multiples of 4) */
  float lr, beta1, beta2, eps, grad_scale;
  const int* adam_step;   /* device: 1-based step */
  int blocks;             /* 0: default grid */
} mvae_dp_reduce_adam_args;
int mvae_dp_reduce_adam(const mvae_dp_reduce_adam_args* args, void* stream);

/* uint8 pixels -> [0,1] activations (image.view(-1,784) of ToTensor(), mnist/train.py:106,131) */
int mvae_u8_to_act(const uint8_t* in, float* out_f32, void* out_bf16, int64_t count, float scale, void* stream);

/* loss_function (mnist/train.py:64-81) on the module outputs, for callers of forward() + loss_function():
 *   lambda_image * BCE_mean(recon_image, image) + lambda_text * NLL_mean(recon_text, text) + kl_weight * KL
 * recon_image / image: [batch, n_pixels] probabilities / targets (image_dtype storage) or both NULL;
 * recon_text: [batch, n_classes] log-probabilities + text int64 [batch], or both NULL.
 * forward writes out4 = {total, image term, text term, KL term} (device); backward scales by *grad_out. */
typedef struct mvae_elbo_loss_args {
  int image_dtype;
  int64_t batch;
  int n_pixels, n_classes, n_latents;
  const void* recon_image; const void* image;
  const float* recon_text; const int64_t* text;
  const float* mu; const float* logvar;
  float lambda_image, lambda_text, kl_weight;
} mvae_elbo_loss_args;
int mvae_elbo_loss_forward(const mvae_elbo_loss_args* args, float* out4, void* stream);
int mvae_elbo_loss_backward(const mvae_elbo_loss_args* args, const float* grad_out, void* d_recon_image,
                            float* d_recon_text, float* d_mu, float* d_logvar, void* stream);

/* ProductOfExperts (mnist/model.py:173-185) for M experts with an optional per-sample presence mask
 * [M, batch] (1 = present).  mu/logvar: [M, batch, dim]; outputs [batch, dim]. */
int mvae_poe_forward(int mode, int prior_expert, float eps, int n_experts, int64_t batch, int dim, const float* mu,
                     const float* logvar, const float* mask, float* out_mu, float* out_logvar, void* stream);
int mvae_poe_backward(int mode, int prior_expert, float eps, int n_experts, int64_t batch, int dim, const float* mu,
                      const float* logvar, const float* mask, const float* d_out_mu, const float* d_out_logvar,
                      float* d_mu, float* d_logvar, void* stream);


/* ------------------------------------------------------------------------------------------
 * Operator-level entries for the convolutional MVAEs (CelebA celeba/model.py:91-200, MultiMNIST
 * multimnist/model.py:150-266).  The host (multimodal-vae_b200/celeba.py) composes them with mvae_gemm into the
 * reference's forward / loss / backward; every convolution is a tensor-core GEMM over an im2col matrix.
 *
 * Activations are NHWC matrices [batch*H*W, C] in fp32 or bf16.  The K axis of a patch matrix is ordered
 * (kh, kw, c): Conv2d weights are therefore held as [Cout, kh, kw, Cin] and ConvTranspose2d weights as
 * [Cin, kh, kw, Cout] inside the flat parameter buffer (the host permutes at the state_dict boundary).
 */
#define MVAE_ACT_NONE 0
#define MVAE_ACT_RELU 1
#define MVAE_ACT_SWISH 2 /* x * sigmoid(x), celeba/model.py:238-245 */

typedef struct mvae_conv_geometry {
  int batch, height, width, channels; /* the IMAGE side: conv input / conv-transpose output */
  int kernel, stride, pad;
  int64_t stride_n, stride_h, stride_w, stride_c; /* element strides of the image tensor (NHWC dense or e.g. NCHW) */
} mvae_conv_geometry;
int mvae_conv_out_size(int in, int kernel, int stride, int pad);
/* Implicit GEMM: mvae_gemm where ONE operand is the patch matrix of `geometry` read straight from the channels-last bf16
 * image (never materialised): patch_operand 1 -> A = im2col(image) [M = batch*Ho*Wo, K = k*k*C] (args->A = image, lda
 * ignored; Conv2d forward celeba/model.py:101-113, ConvTranspose2d input gradient); patch_operand 2 -> B = im2col(image)
 * with the pixel index as the contraction (args->B = image, b_major = 1, N = k*k*C, K = batch*Ho*Wo; weight gradients).
 * Needs C % 8 == 0 and bf16 storage; bit-identical to mvae_im2col followed by mvae_gemm. */
int mvae_conv_gemm(const mvae_gemm_args* args, const mvae_conv_geometry* geometry, int patch_operand, void* stream);

/* One output-parity class of a ConvTranspose2d forward / Conv2d input gradient as an implicit GEMM, i.e. without the
 * [pixels, k*k*C_out] patch matrix and without mvae_col2im (celeba/model.py:142-152, multimnist/model.py:198-210):
 *   out[n, s*u + a, s*v + b, :] = sum_{th, tw, ci} x[n, u - pad_h + th, v - pad_w + tw, ci] * W[ci, kh[th], kw[tw], :]
 * x: channels-last bf16 [batch, in_h, in_w, channels] (channels % 64 == 0); weight: bf16 [channels, kernel*kernel, out_channels]
 * (ld_tap elements between taps); out: channels-last image [batch, out_h, out_w, ldc] (bf16 / fp32, ldc % 8 == 0).
 * The host loops over the stride*stride classes; mvae_b200._ops.transposed_conv_classes() derives the fields.
 * Parity: test_transposed_conv_implicit_matches_torch (5 geometries).  The bf16 hosts call this entry (one launch per class)
 * for geometries whose classes differ in shape (k5 s2), and mvae_convt_gemm (all classes in one launch) otherwise. */
typedef struct mvae_convt_class {
  int batch, in_h, in_w, channels;
  int out_h, out_w, out_channels;
  int kernel, stride;
  int a, b;                         /* output parity (row, column) of this class */
  int count_h, count_w;             /* class grid: output rows s*u + a for u < count_h */
  int taps_h, taps_w, pad_h, pad_w; /* tap window of the class (stride-1 gather over x) */
  int kh[8], kw[8];                 /* kernel row / column used by window position t */
} mvae_convt_class;
int mvae_convt_class_gemm(const mvae_convt_class* c, int dtype, const void* x, const void* weight, int64_t ld_tap, void* out,
                          int64_t ldc, int out_dtype, void* stream);
/* The whole transposed convolution in ONE launch (the bf16 hosts' default; parity:
 * test_transposed_conv_merged_classes_matches_torch) - grid.z enumerates the stride*stride parity classes of
 * mvae_convt_class_gemm.  Only for geometries whose classes have equal shape (kernel % stride == 0 and an output size
 * divisible by the stride, e.g. k4 s2 p1, or stride 1); returns 4 without launching otherwise (callers then loop over
 * mvae_convt_class_gemm).  x [batch, in_h, in_w, channels] bf16, weight [channels, kernel*kernel, out_channels] bf16,
 * out [batch, out_h, out_w, ldc] with out_h = (in_h-1)*stride - 2*pad + kernel. */
/* Host helper (no GPU work): the output-parity classes of a transposed convolution along one axis, stride <= 4.  For class
 * a < stride: count[a] outputs stride*u + a, taps[a] window positions, pad_lo[a], kernel index kh[a*8 + t] of window position t
 * (out[stride*u + a] = sum_t x[u - pad_lo[a] + t] * w[kh[a*8 + t]]).  Returns the output size, or -1 for a bad geometry. */
int mvae_convt_axis_classes(int kernel, int stride, int pad, int size_in, int* count, int* taps, int* pad_lo, int* kh);
int mvae_convt_gemm(int dtype, int batch, int in_h, int in_w, int channels, int out_channels, int kernel, int stride, int pad,
                    const void* x, const void* weight, int64_t ld_tap, void* out, int64_t ldc, int out_dtype, void* stream);
/* col[m, (kh*k+kw)*C + c] = image[n, ho*s-p+kh, wo*s-p+kw, c] (0 outside), m = (n*Ho+ho)*Wo+wo.
 * replaces: the patch gather inside nn.Conv2d forward / ConvTranspose2d backward (celeba/model.py:101-113, 142-152). */
int mvae_im2col(const mvae_conv_geometry* g, int image_dtype, const void* image, int col_dtype, void* col, int64_t ldcol,
                void* stream);
/* adjoint of mvae_im2col (gather form, no atomics): image[n,h,w,c] = sum of the col entries that im2col read from it.
 * replaces: the scatter inside nn.ConvTranspose2d forward / Conv2d input-gradient. */
int mvae_col2im(const mvae_conv_geometry* g, int col_dtype, const void* col, int64_t ldcol, int image_dtype, void* image,
                void* stream);
/* sum[g, c] += sum_rows x, sumsq[g, c] += sum_rows x^2 (sumsq may be NULL), group = row / rows_per_group; only columns
 * < valid_channels are accumulated.  BatchNorm batch statistics and bias gradients. */
int mvae_col_stats(int dtype, const void* x, int64_t rows, int channels, int valid_channels, int64_t rows_per_group,
                   float* sum, float* sumsq, void* stream);

/* nn.BatchNorm1d / nn.BatchNorm2d (channel-last matrix) fused with the activation that follows it.
 * Each statistics group (rows_per_group consecutive rows) stands for one forward pass of the reference (one ELBO
 * term); running statistics are updated once per group, in order, `updates_per_group` times each. */
typedef struct mvae_bn_act_args {
  int dtype;
  int64_t rows;
  int channels;
  int64_t rows_per_group; /* <= 0: one group */
  int act;                /* MVAE_ACT_* */
  int training;           /* 0: normalise with the running statistics, update nothing */
  const void* x;          /* pre-BatchNorm activations [rows, channels] */
  void* y;                /* forward output */
  const float* gamma; const float* beta;
  float* sum; float* sumsq; /* [groups, channels] batch sums */
  int stats_ready;          /* 1: sum / sumsq already hold the batch sums (e.g. from a GEMM epilogue) */
  float* save_mean; float* save_rstd; /* [groups, channels], written by forward, read by backward */
  float* running_mean; float* running_var;
  int updates_per_group;
  float momentum, eps;
  const void* dy; void* dx; /* backward: gradient at y -> gradient at x */
  float* s0; float* s1;     /* [groups, channels] scratch of the backward */
  float* dgamma; float* dbeta; /* += (may be NULL) */
} mvae_bn_act_args;
int mvae_bn_act_forward(const mvae_bn_act_args* args, void* stream);
int mvae_bn_act_backward(const mvae_bn_act_args* args, void* stream);

/* y[rep*rows + r] = dropout_rep(act(x[r])) for rep < repeat: activation after a Linear, optional nn.Dropout(p)
 * (celeba/model.py:116-117) with an independent Philox keep-mask per replica (one replica per reference forward pass).
 * backward: dx[r] = act'(x[r]) * sum_rep mask_rep * dy[rep*rows + r] / (1-p);  dbias[c] += sum_r dx[r, c] if not NULL. */
int mvae_act_forward(int dtype, int act, const void* x, void* y, int64_t rows, int channels, int repeat, float dropout_p,
                     uint64_t seed, const int* step_counter, void* stream);
int mvae_act_backward(int dtype, int act, const void* x, const void* dy, void* dx, int64_t rows, int channels, int repeat,
                      float dropout_p, uint64_t seed, const int* step_counter, float* dbias, void* stream);

/* F.sigmoid + F.binary_cross_entropy(mean) (celeba/model.py:163,200; celeba/train.py:66-74), value and gradient:
 *   loss[g] += sum_{rows of group g} BCE(sigmoid(logit), target)      (unscaled sum)
 *   dlogits  = grad_scale[g] * (sigmoid(logit) - target)
 * Row m compares with target row m % target_rows.  With target == NULL and dprobs != NULL it is the plain sigmoid
 * backward dlogits = dprobs * p * (1 - p).  Columns [cols, ld_dlogits) of dlogits are zero-filled. */
typedef struct mvae_sigmoid_bce_args {
  int64_t rows; int cols; int64_t rows_per_group;
  int logit_dtype; const void* logits; int64_t ld_logits;
  int target_dtype; const void* target; int64_t ld_target; int64_t target_rows;
  float grad_scale[3];
  float* loss;
  int prob_dtype; void* probs; int64_t ld_probs;
  int grad_dtype; void* dlogits; int64_t ld_dlogits;
  const float* dprobs; int64_t ld_dprobs;
} mvae_sigmoid_bce_args;
int mvae_sigmoid_bce(const mvae_sigmoid_bce_args* args, void* stream);

/* The latent path for all ELBO terms in one launch per direction: ProductOfExperts over the experts present in each
 * term (celeba/model.py:38-52, 224-235), reparametrize (:27-34) and the KL term of loss_function (celeba/train.py:79-80).
 * Expert tensors are [rows, 2*n_latents] (mu | logvar).  Term g reads expert A rows [expert_a_row0[g], +batch) (two terms
 * may share rows: their gradients are summed) and expert B rows [0, batch). */
typedef struct mvae_latent_args {
  int64_t batch; int n_latents; int n_terms;
  int term_type[3];          /* MVAE_TERM_JOINT / _IMAGE (expert A only) / _TEXT (expert B only) */
  int poe_mode, prior_expert; float poe_eps;
  const float* expert_a; int64_t ld_a; int64_t expert_a_row0[3];
  const float* expert_b; int64_t ld_b;
  const float* eps;          /* [n_terms, batch, n_latents] injected N(0,1) draws, or NULL (Philox: seed, *step_counter) */
  uint64_t seed; const int* step_counter;
  int training;              /* 0: z = mu */
  float kl_weight[3];
  int z_dtype; void* z; int64_t ld_z;   /* [n_terms*batch, ld_z] */
  float* mu; float* logvar;  /* optional [n_terms, batch, n_latents] */
  float* kl;                 /* [n_terms] += kl_weight[g] * KL_g */
  /* backward */
  int dz_dtype; const void* dz; int64_t ld_dz;
  const float* d_mu; const float* d_logvar; /* optional upstream gradients [n_terms, batch, n_latents] */
  int d_dtype; void* d_expert_a; int64_t ld_da; void* d_expert_b; int64_t ld_db;
  const float* row_weight;   /* optional [n_terms * batch] fp32: scales kl_weight[g] for row (g, b) (per-sample masks) */
} mvae_latent_args;
int mvae_latent_forward(const mvae_latent_args* args, void* stream);
int mvae_latent_backward(const mvae_latent_args* args, void* stream);

/* Per-sample missing-modality masks -> per-(term, row) weights, on the device (no host sync, CUDA-graph capturable):
 *   term g is ACTIVE on row b iff the modalities it needs are present (joint: both, image: has_image, text: has_text);
 *   weight[g * batch + b] = active ? batch / count_g : 0, count_g = active rows of term g (weight 0 everywhere if count_g = 0),
 * so that a loss kernel scaling by lambda / batch * weight yields the mean over the term's own rows
 * (mnist/paired_weak.py:82-104 per row instead of per batch).  counts (optional) [n_terms] receives count_g as floats. */
int mvae_mask_weights(const uint8_t* has_image, const uint8_t* has_text, int64_t batch, int n_terms, const int* term_type,
                      float* weight, float* counts, void* stream);

/* dst[r, c] = src[r, c] for c < cols, 0 for cols <= c < ld_dst: GEMM-operand copies of weights whose row length does
 * not meet the 16-byte TMA stride rule (e.g. Linear(18, 64), Linear(100, 6400) in bf16). */
int mvae_cast_pad_2d(const float* src, int64_t rows, int64_t cols, int64_t ld_src, int dst_dtype, void* dst, int64_t ld_dst,
                     void* stream);
/* Start of a training step: ++*step_counter (Adam bias correction, Philox counter), zero `zero_floats` floats (loss and
 * statistics accumulators), counters[i] += increments[i] (num_batches_tracked). */
int mvae_step_begin(int* step_counter, float* zero_buf, int64_t zero_floats, int64_t* counters, const int64_t* increments,
                    int n_counters, void* stream);


/* ------------------------------------------------------------------------------------------
 * MultiMNIST text path (multimnist/model.py:220-307).  The projections of every GRU cell are mvae_gemm calls; these
 * entries are what sits between them.  Matrices carry explicit leading dimensions so that the reference's torch.cat
 * of (embedding, z) and (hidden, z) are column ranges of one buffer.
 */
/* out[m, j] = act(table[indices[m * index_stride], j]), j < width: nn.Embedding (+ Swish, multimnist/model.py:299). */
int mvae_embed_forward(const int64_t* indices, int64_t index_stride, const float* table, int vocab, int width, int act,
                       int out_dtype, void* out, int64_t ld_out, int64_t rows, void* stream);
/* dtable[indices[m], j] += dout[m, j] * act'(table[indices[m], j]) */
int mvae_embed_backward(const int64_t* indices, int64_t index_stride, const float* table, int vocab, int width, int act,
                        int dout_dtype, const void* dout, int64_t ld_dout, int64_t rows, float* dtable, void* stream);

/* One nn.GRU cell step given gi = x W_ih^T + b_ih and gh = h W_hh^T + b_hh (fp32, gate order r, z, n):
 *   r = sigmoid(gi_r + gh_r), z = sigmoid(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (1 - z) * n + z * h.
 * forward writes h' (+ addend) to h_out (and h_out2) and keeps (r, z, n, gh_n) in `saved` [rows, 4*hidden];
 * backward turns the gradient at h' (dh_a + dh_b) into dgi / dgh (GEMM operands, activation dtype) and the direct part
 * of the gradient at h (dh_prev = dh * z; the caller adds dgh W_hh with an accumulating GEMM). */
typedef struct mvae_gru_cell_args {
  int64_t rows; int hidden;
  const float* gi; int64_t ld_gi;
  const float* gh; int64_t ld_gh;
  int h_dtype; const void* h_prev; int64_t ld_h_prev;   /* NULL: zeros */
  const void* addend; int64_t ld_addend;                /* optional, h_dtype */
  void* h_out; int64_t ld_h_out;
  void* h_out2; int64_t ld_h_out2;                       /* optional second destination */
  float* saved;
  int dh_a_dtype; const void* dh_a; int64_t ld_dh_a;
  int dh_b_dtype; const void* dh_b; int64_t ld_dh_b;
  int dg_dtype; void* dgi; void* dgh; int64_t ld_dg;     /* columns [3*hidden, ld_dg) are zero-filled */
  float* dh_prev; int64_t ld_dh_prev;
} mvae_gru_cell_args;
int mvae_gru_cell_forward(const mvae_gru_cell_args* args, void* stream);
int mvae_gru_cell_backward(const mvae_gru_cell_args* args, void* stream);

/* F.log_softmax + F.nll_loss + torch.max of one decoding step (multimnist/model.py:285-287, 303-306; train.py:78):
 *   logp = log_softmax(logits); argmax[m] = first maximum; loss[g] += -logp[m, target]; dlogits = grad_scale[g]*(softmax - onehot).
 * Row m compares with target[(m % target_rows) * target_stride]. */
typedef struct mvae_logsoftmax_nll_args {
  int64_t rows; int classes; int64_t rows_per_group;
  const float* logits; int64_t ld_logits;
  const int64_t* target; int64_t target_stride; int64_t target_rows;
  float grad_scale[3];
  float* loss;
  float* logp; int64_t ld_logp;
  int64_t* argmax;
  int grad_dtype; void* dlogits; int64_t ld_dlogits;
  const float* row_weight;   /* optional [rows] fp32: scales row m's loss and gradient (per-sample masks) */
} mvae_logsoftmax_nll_args;
int mvae_logsoftmax_nll(const mvae_logsoftmax_nll_args* args, void* stream);

/* Backward of F.log_softmax for callers that differentiate the returned log-probabilities themselves (module path):
 * dlogits = dlogp - exp(logp) * sum_c dlogp; columns [classes, ld_dlogits) zero-filled. */
int mvae_logsoftmax_backward(const float* logp, int64_t ld_logp, const float* dlogp, int64_t ld_dlogp, int64_t rows, int classes,
                             int grad_dtype, void* dlogits, int64_t ld_dlogits, void* stream);

/* dst[r, c] (+)= src[r, c] (+ src2[r, c]) for c < cols, with independent dtypes and leading dimensions. */
int mvae_copy_2d(int src_dtype, const void* src, int64_t ld_src, int dst_dtype, void* dst, int64_t ld_dst, int64_t rows,
                 int64_t cols, int accumulate, int src2_dtype, const void* src2, int64_t ld_src2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVAE_B200_H_ */
