// Launchers of the non-GEMM kernels (internal).  Every launcher enqueues on `st`, returns 0 / error code.
#pragma once
#include "common.cuh"

namespace mvae {

constexpr int kMaxGroups = 3;  // ELBO terms evaluated together: joint, image-only, text-only

// ---- elementwise.cu
int launch_bn_forward(int dtype, const void* x, void* y, int rows, int F, int rows_per_group, const float* sum,
                      const float* sumsq, const float* gamma, const float* beta, float* save_mean, float* save_rstd,
                      float* running_mean, float* running_var, int updates_per_group, float momentum, float eps,
                      int relu, cudaStream_t st);
int launch_bn_backward(int dtype, const void* dyhat, const void* x, void* dx, int rows, int F, int rows_per_group,
                       const float* s0, const float* s1, const float* mean, const float* rstd, const float* gamma,
                       float* dgamma, float* dbeta, cudaStream_t st);
int launch_cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t st);
int launch_u8_to_act(const uint8_t* in, float* o32, void* o16, long long n, float scale, cudaStream_t st);
int launch_adam(float* p, float* g, float* m, float* v, void* p16, long long n, float lr, float b1, float b2, float eps,
                const int* step_ptr, float grad_scale, int zero_grad, cudaStream_t st);
int launch_step_prep(int* step_ptr, int* step_ptr2, float* zero_buf, long long zero_n, long long* nbt, const long long (&inc)[6],
                     cudaStream_t st);
int launch_loss_pack(const float* acc, float* out, int G, cudaStream_t st);
int launch_sigmoid_backward(int dtype, const void* dprob, const void* prob, void* dlogit, int rows, int F, float* dbias,
                            cudaStream_t st);

// ---- tail.cu
enum : int { TERM_JOINT = 0, TERM_IMAGE = 1, TERM_TEXT = 2 };

struct TailArgs {
  int B = 0, n = 0, G = 0;
  int group_type[kMaxGroups] = {0, 0, 0};
  int poe_mode = 0;      // MVAE_POE_REF / MVAE_POE_PRECISION
  int prior_expert = 0;  // PRECISION mode only
  float poe_eps = 1e-8f;
  int z_dtype = MVAE_F32;
  // experts
  const float* enc_img = nullptr;    // [B, 2n] (mu | logvar), null if no group uses the image
  const float* txt_table = nullptr;  // [10, 2n] per-label (mu | logvar), null if no group uses the text
  const long long* labels = nullptr; // [B]
  // noise: injected [G, B, n] or generated (Philox4x32-10 keyed by seed, counter from *step_ptr)
  const float* eps = nullptr;
  unsigned long long seed = 0;
  const int* step_ptr = nullptr;
  int training = 1;  // 0: z = mu (mnist/model.py:29-30)
  float* noise_buf = nullptr;  // [G, B, n] optional: the forward leaves the draws it used here, the backward reads them back
                               // instead of re-running Philox (and stays correct if the step counter has moved on since)
  const float* z_in = nullptr;  // [G*B, n]: latents given by the caller (decode only; experts are ignored)
  float kl_weight[kMaxGroups] = {0, 0, 0};
  // text decoder layer 1 (Linear n -> 10), fused because z is in registers here
  const float* wt1 = nullptr;  // [10, n]
  const float* bt1 = nullptr;  // [10]
  // ---- forward outputs
  void* z = nullptr;           // [G*B, n] z_dtype
  float* mu = nullptr;         // [G, B, n] optional
  float* logvar = nullptr;     // [G, B, n] optional
  float* kl = nullptr;         // [G] += kl_weight[g] * KL_g
  float* t1pre = nullptr;      // [G*B, 10]
  float* t1_sum = nullptr;     // [G, 10] +=
  float* t1_sumsq = nullptr;   // [G, 10] +=
  // ---- backward inputs
  const float* dz = nullptr;       // [G*B, n] from the image decoder's dgrad
  const float* dmu_up = nullptr;   // [G, B, n] optional upstream gradients (module path)
  const float* dlogvar_up = nullptr;
  const float* t1_dyhat = nullptr; // [G*B, 10] masked gradient at the text decoder's BN output
  const float* t1_s0 = nullptr;    // [G, 10] sum dyhat
  const float* t1_s1 = nullptr;    // [G, 10] sum dyhat*xhat
  const float* t1_gamma = nullptr; // [10]
  // ---- backward outputs
  void* d_enc = nullptr;           // [B, 2n] z_dtype (operand of the encoder dgrad/wgrad GEMMs)
  float* d_enc_bias = nullptr;     // [2n] += column sums of d_enc
  float* d_txt_table = nullptr;    // [10, 2n] +=
  float* d_wt1 = nullptr;          // [10, n] +=
  float* d_t1_gamma = nullptr;     // [10] +=
  float* d_t1_beta = nullptr;      // [10] +=
};
int launch_tail_forward(const TailArgs& a, cudaStream_t st);
int launch_tail_backward(const TailArgs& a, cudaStream_t st);

struct TextDecArgs {
  int B = 0, G = 0;
  const float* t1pre = nullptr;   // [G*B, 10]
  const float* t1_sum = nullptr;  // [G, 10]
  const float* t1_sumsq = nullptr;
  const float* gamma = nullptr;   // [10]
  const float* beta = nullptr;
  float* running_mean = nullptr;  // [10] optional
  float* running_var = nullptr;
  float momentum = 0.1f, bn_eps = 1e-5f;
  int training = 1;               // 0: normalise with the running statistics
  const float* w2 = nullptr;      // [10, 10]
  const float* b2 = nullptr;      // [10]
  const long long* labels = nullptr;  // [B]
  float ce_scale[kMaxGroups] = {0, 0, 0};  // lambda_yx / B per group
  const float* dlogp_up = nullptr;     // [G*B, 10] upstream gradient of the log-probs (module path), or null
  int fused_loss = 1;                  // 1: gradient of ce_scale * NLL, loss accumulated
  // outputs
  float* logp = nullptr;      // [G*B, 10] optional
  float* ce = nullptr;        // [G] += ce_scale[g] * sum NLL
  float* dyhat = nullptr;     // [G*B, 10]
  float* s0 = nullptr;        // [G, 10] +=
  float* s1 = nullptr;        // [G, 10] +=
  float* d_w2 = nullptr;      // [10, 10] +=
  float* d_b2 = nullptr;      // [10] +=
  int backward = 1;           // 0: forward only (no dyhat / gradients)
};
int launch_textdec(const TextDecArgs& a, cudaStream_t st);

struct TextEncArgs {
  int B = 0, n = 0;
  const long long* labels = nullptr;
  const float* emb = nullptr;    // [10, 50]
  const float* gamma = nullptr;  // [50]
  const float* beta = nullptr;
  const float* w = nullptr;      // [2n, 50]
  const float* b = nullptr;      // [2n]
  float* running_mean = nullptr; // [50]
  float* running_var = nullptr;
  int updates = 1;               // how many reference forward passes this stands for
  float momentum = 0.1f, bn_eps = 1e-5f;
  int training = 1;
  float* table = nullptr;        // [10, 2n] out
  float* save = nullptr;         // workspace [10 + 10*50*2 + 50*2]: counts, xhat, h, mean, rstd
  // backward
  const float* d_table = nullptr;  // [10, 2n]
  float* d_emb = nullptr;          // [10, 50] +=
  float* d_gamma = nullptr;
  float* d_beta = nullptr;
  float* d_w = nullptr;            // [2n, 50] +=
  float* d_b = nullptr;            // [2n] +=
};
constexpr int kTextEncSaveFloats = 10 + 10 * 50 * 2 + 50 * 2;
int launch_textenc_forward(const TextEncArgs& a, cudaStream_t st);
int launch_textenc_backward(const TextEncArgs& a, cudaStream_t st);

// Standalone product of experts with a per-sample presence mask (ProductOfExperts()(mu, logvar, mask)).
int launch_poe_forward(int mode, int prior, float eps, int M, long long B, int D, const float* mu, const float* logvar,
                       const float* mask, float* out_mu, float* out_logvar, cudaStream_t st);
int launch_poe_backward(int mode, int prior, float eps, int M, long long B, int D, const float* mu, const float* logvar,
                        const float* mask, const float* d_out_mu, const float* d_out_logvar, float* d_mu,
                        float* d_logvar, cudaStream_t st);

// ---- chain.cu: slab-persistent layer chains of the MNIST step (bf16).  All pointers are device pointers; weights are
// the bf16 mirror, biases / BatchNorm parameters fp32; statistics buffers are [2][G][F] (sum | sumsq resp. s0 | s1) inside
// the zeroed accumulator region of the step workspace; counters are zeroed grid-barrier arrival counters.
bool chain_supported(int B, int G, int n);
void set_chain_debug_times(void* ptr);
struct ChainEncFwd {
  int B = 0, n = 0, bn_updates = 1;
  const __nv_bfloat16* image = nullptr;
  const __nv_bfloat16 *w1 = nullptr, *w2 = nullptr, *w3 = nullptr;
  const float *b1 = nullptr, *b2 = nullptr, *b3 = nullptr;
  const float *gamma1 = nullptr, *beta1 = nullptr, *gamma2 = nullptr, *beta2 = nullptr;
  float *st1 = nullptr, *st2 = nullptr, *sv1 = nullptr, *sv2 = nullptr;
  float *rm1 = nullptr, *rv1 = nullptr, *rm2 = nullptr, *rv2 = nullptr;
  unsigned int* counters = nullptr;   // [2]
  __nv_bfloat16 *h1pre = nullptr, *h1 = nullptr, *h2pre = nullptr, *h2 = nullptr;
  float* enc = nullptr;               // [B, 2n]
  unsigned int* err = nullptr;
};
struct ChainDecFwd {
  int B = 0, n = 0, G = 0;
  const __nv_bfloat16* z = nullptr;   // [G*B, n]
  const __nv_bfloat16 *w1 = nullptr, *w2 = nullptr, *w3 = nullptr;
  const float *b1 = nullptr, *b2 = nullptr, *b3 = nullptr;
  const float *gamma1 = nullptr, *beta1 = nullptr, *gamma2 = nullptr, *beta2 = nullptr;
  float *st1 = nullptr, *st2 = nullptr, *sv1 = nullptr, *sv2 = nullptr;
  float *rm1 = nullptr, *rv1 = nullptr, *rm2 = nullptr, *rv2 = nullptr;
  unsigned int* counters = nullptr;   // [6]
  __nv_bfloat16 *g1pre = nullptr, *g1 = nullptr, *g2pre = nullptr, *g2 = nullptr, *dlog = nullptr, *probs = nullptr;
  const __nv_bfloat16* image = nullptr;  // [B, 784] BCE target
  float bce_scale[3] = {0, 0, 0};
  float* loss = nullptr;              // [G] +=
  float* dbias3 = nullptr;            // [784] += (null: no backward)
  unsigned int* err = nullptr;
};
struct ChainDecBwd {
  int B = 0, n = 0, G = 0;
  const __nv_bfloat16* dlog = nullptr;
  const __nv_bfloat16 *w1 = nullptr, *w2 = nullptr, *w3 = nullptr;
  const float *gamma1 = nullptr, *beta1 = nullptr, *gamma2 = nullptr, *beta2 = nullptr;
  float *sb1 = nullptr, *sb2 = nullptr;
  float *sv1 = nullptr, *sv2 = nullptr;
  unsigned int* counters = nullptr;   // [6]
  const __nv_bfloat16 *g1pre = nullptr, *g2pre = nullptr;
  __nv_bfloat16 *dy2 = nullptr, *dy1 = nullptr;
  float* dz = nullptr;                // [G*B, n]
  float *dgamma1 = nullptr, *dbeta1 = nullptr, *dgamma2 = nullptr, *dbeta2 = nullptr;
  unsigned int* err = nullptr;
};
struct ChainEncBwd {
  int B = 0, n = 0;
  const __nv_bfloat16* denc = nullptr;  // [B, 2n]
  const __nv_bfloat16 *w2 = nullptr, *w3 = nullptr;
  const float *gamma1 = nullptr, *beta1 = nullptr, *gamma2 = nullptr, *beta2 = nullptr;
  float *sb1 = nullptr, *sb2 = nullptr;
  float *sv1 = nullptr, *sv2 = nullptr;
  unsigned int* counters = nullptr;   // [2]
  const __nv_bfloat16 *h1pre = nullptr, *h2pre = nullptr;
  __nv_bfloat16 *dye2 = nullptr, *dye1 = nullptr;
  float *dgamma1 = nullptr, *dbeta1 = nullptr, *dgamma2 = nullptr, *dbeta2 = nullptr;
  unsigned int* err = nullptr;
  int max_parts = 4;   // column split of the last layer: CTAs per slab (fewer when another kernel needs SMs beside this one)
};
int launch_chain_enc_fwd(const ChainEncFwd& a, cudaStream_t st);
int launch_chain_dec_fwd(const ChainDecFwd& a, cudaStream_t st);
int launch_chain_dec_bwd(const ChainDecBwd& a, cudaStream_t st);
int launch_chain_enc_bwd(const ChainEncBwd& a, cudaStream_t st);

}  // namespace mvae
