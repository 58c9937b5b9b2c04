// MNIST MVAE training step: parameter layout, workspace plan and the launch sequence
// (forward of all ELBO terms at once, backward, fused Adam) - the C ABI's mvae_mnist_* entries.
//
// Reference: MultimodalVAE (mnist/model.py:14-170), loss_function (mnist/train.py:64-81) and the
// three-forward step mnist/train.py:132-153, restructured as SURVEY.md section 7 describes:
//   * each encoder runs ONCE (its output is identical in the joint and the unimodal term), its
//     gradient contributions are summed and its running statistics are advanced once per term;
//   * both decoders run ONCE over the stacked [terms*B, .] batch with per-term BatchNorm statistics;
//   * the BCE of the image decoder is fused into the last GEMM's epilogue (logits never reach HBM).
#include <string.h>

#include "../../include/mvae_b200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace mvae {

namespace {

struct TensorInfo {
  const char* name;
  int ndim;
  long long shape[2];
  int kind;  // 0 parameter, 1 float buffer, 2 num_batches_tracked
};

// state_dict order of mnist/model.py (MultimodalVAE.__init__: image_encoder, image_decoder, text_encoder, text_decoder)
int build_tensor_list(int n, TensorInfo* out) {
  int c = 0;
  auto lin = [&](const char* w, const char* b, long long o, long long i) {
    out[c++] = {w, 2, {o, i}, 0};
    out[c++] = {b, 1, {o, 0}, 0};
  };
  auto bn = [&](const char* w, const char* b, const char* rm, const char* rv, const char* nbt, long long f) {
    out[c++] = {w, 1, {f, 0}, 0};
    out[c++] = {b, 1, {f, 0}, 0};
    out[c++] = {rm, 1, {f, 0}, 1};
    out[c++] = {rv, 1, {f, 0}, 1};
    out[c++] = {nbt, 0, {0, 0}, 2};
  };
  lin("image_encoder.net.0.weight", "image_encoder.net.0.bias", 400, 784);
  bn("image_encoder.net.1.weight", "image_encoder.net.1.bias", "image_encoder.net.1.running_mean",
     "image_encoder.net.1.running_var", "image_encoder.net.1.num_batches_tracked", 400);
  lin("image_encoder.net.3.weight", "image_encoder.net.3.bias", 200, 400);
  bn("image_encoder.net.4.weight", "image_encoder.net.4.bias", "image_encoder.net.4.running_mean",
     "image_encoder.net.4.running_var", "image_encoder.net.4.num_batches_tracked", 200);
  lin("image_encoder.net.6.weight", "image_encoder.net.6.bias", 2 * n, 200);
  lin("image_decoder.net.0.weight", "image_decoder.net.0.bias", 200, n);
  bn("image_decoder.net.1.weight", "image_decoder.net.1.bias", "image_decoder.net.1.running_mean",
     "image_decoder.net.1.running_var", "image_decoder.net.1.num_batches_tracked", 200);
  lin("image_decoder.net.3.weight", "image_decoder.net.3.bias", 400, 200);
  bn("image_decoder.net.4.weight", "image_decoder.net.4.bias", "image_decoder.net.4.running_mean",
     "image_decoder.net.4.running_var", "image_decoder.net.4.num_batches_tracked", 400);
  lin("image_decoder.net.6.weight", "image_decoder.net.6.bias", 784, 400);
  out[c++] = {"text_encoder.net.0.weight", 2, {10, 50}, 0};
  bn("text_encoder.net.1.weight", "text_encoder.net.1.bias", "text_encoder.net.1.running_mean",
     "text_encoder.net.1.running_var", "text_encoder.net.1.num_batches_tracked", 50);
  lin("text_encoder.net.3.weight", "text_encoder.net.3.bias", 2 * n, 50);
  lin("text_decoder.net.0.weight", "text_decoder.net.0.bias", 10, n);
  bn("text_decoder.net.1.weight", "text_decoder.net.1.bias", "text_decoder.net.1.running_mean",
     "text_decoder.net.1.running_var", "text_decoder.net.1.num_batches_tracked", 10);
  lin("text_decoder.net.3.weight", "text_decoder.net.3.bias", 10, 10);
  return c;
}

constexpr int kMaxTensors = 64;
constexpr long long kAlignFloats = 64;  // 256-byte alignment of every tensor in the flat buffers

long long numel(const TensorInfo& t) {
  long long n = 1;
  for (int i = 0; i < t.ndim; ++i) n *= t.shape[i];
  return n;
}
long long round_up(long long v, long long a) { return (v + a - 1) / a * a; }

struct Layout {
  int count = 0;
  TensorInfo info[kMaxTensors];
  long long offset[kMaxTensors];  // into params (kind 0), buffers (kind 1) or nbt (kind 2)
  long long param_floats = 0, buffer_floats = 0, nbt_count = 0, enc_floats = 0;
  long long find(const char* name) const {
    for (int i = 0; i < count; ++i)
      if (strcmp(info[i].name, name) == 0) return offset[i];
    return -1;
  }
};

// Parameters are laid out in two contiguous buckets so that data-parallel training can all-reduce the decoder
// gradients (complete after the decoder + tail backward) while the encoder backward is still running:
//   bucket 0 = image_encoder.* + text_encoder.*   [0, enc_floats)
//   bucket 1 = image_decoder.* + text_decoder.*   [enc_floats, param_floats)
// The tensor TABLE keeps the reference's state_dict order; only the offsets follow the buckets.
bool is_decoder_tensor(const char* name) { return strstr(name, "_decoder.") != nullptr; }

Layout make_layout(int n) {
  Layout L;
  L.count = build_tensor_list(n, L.info);
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < L.count; ++i) {
      if (L.info[i].kind != 0 || is_decoder_tensor(L.info[i].name) != (pass == 1)) continue;
      L.offset[i] = L.param_floats;
      L.param_floats += round_up(numel(L.info[i]), kAlignFloats);
    }
    if (pass == 0) L.enc_floats = L.param_floats;
  }
  for (int i = 0; i < L.count; ++i) {
    const long long ne = numel(L.info[i]);
    if (L.info[i].kind == 0) {
      continue;
    } else if (L.info[i].kind == 1) {
      L.offset[i] = L.buffer_floats;
      L.buffer_floats += round_up(ne, kAlignFloats);
    } else {
      L.offset[i] = L.nbt_count++;
    }
  }
  return L;
}

// ---------------------------------------------------------------- workspace plan
struct Plan {
  long long bytes = 0;
  // zeroed accumulators (one contiguous float region)
  long long acc_off = 0, acc_floats = 0;
  long long st_e1, st_e2, st_d1, st_d2, st_t1;          // forward BN sums: [2][groups][F]
  long long sb_e1, sb_e2, sb_d1, sb_d2, sb_t1;          // backward BN sums
  long long losses;                                      // [3][kMaxGroups]: bce, ce, kl
  long long bars;                                        // grid-barrier arrival counters (uint): [0..15] the chain kernels' BatchNorm barriers, [31] their time-out flag
  long long d_txt_table;                                 // [10][2n]
  // saved statistics
  long long sv_e1, sv_e2, sv_d1, sv_d2;                  // [2][groups][F]: mean, rstd
  long long txt_table, txt_save;
  // activations
  long long noise;                                       // [G*B][n] fp32: the reparametrize draws of the forward, re-read by the tail backward
  long long h1pre, h1, h2pre, h2, enc, z, t1pre, g1pre, g1, g2pre, g2, dlog, dyt, dy2, dy1, dz, denc, dye2, dye1;
  long long x3_a = -1, x3_b = -1;                        // 3xTF32 split-operand scratch (MVAE_DT_F32X3 only)
};

Plan make_plan(int B, int n, int dtype_code) {
  const bool x3 = dtype_code == MVAE_DT_F32X3;
  const int dtype = x3 ? MVAE_F32 : dtype_code;
  Plan p;
  const long long G = kMaxGroups, R = G * B;
  const long long es = dtype == MVAE_F32 ? 4 : 2;
  long long off = 0;
  auto take = [&](long long bytes) {
    const long long o = off;
    off += round_up(bytes, 256);
    return o;
  };
  // accumulators
  p.acc_off = off;
  long long f = 0;
  auto acc = [&](long long floats) {
    const long long o = p.acc_off + f * 4;
    f += round_up(floats, 64);
    return o;
  };
  p.st_e1 = acc(2 * 400); p.st_e2 = acc(2 * 200);
  p.st_d1 = acc(2 * G * 200); p.st_d2 = acc(2 * G * 400); p.st_t1 = acc(2 * G * 10);
  p.sb_e1 = acc(2 * 400); p.sb_e2 = acc(2 * 200);
  p.sb_d1 = acc(2 * G * 200); p.sb_d2 = acc(2 * G * 400); p.sb_t1 = acc(2 * G * 10);
  p.losses = acc(3 * G);
  p.bars = acc(32);
  p.d_txt_table = acc(10 * 2 * n);
  p.acc_floats = f;
  off += f * 4;
  off = round_up(off, 256);
  p.sv_e1 = take(2 * 400 * 4); p.sv_e2 = take(2 * 200 * 4);
  p.sv_d1 = take(2 * G * 200 * 4); p.sv_d2 = take(2 * G * 400 * 4);
  p.txt_table = take(10 * 2 * n * 4);
  p.txt_save = take(kTextEncSaveFloats * 4);
  p.h1pre = take(B * 400 * es); p.h1 = take(B * 400 * es);
  p.h2pre = take(B * 200 * es); p.h2 = take(B * 200 * es);
  p.enc = take(B * 2 * n * 4);
  p.z = take(R * n * es);
  p.noise = take(R * n * 4);
  p.t1pre = take(R * 10 * 4);
  p.g1pre = take(R * 200 * es); p.g1 = take(R * 200 * es);
  p.g2pre = take(R * 400 * es); p.g2 = take(R * 400 * es);
  p.dlog = take(R * 784 * es);
  p.dyt = take(R * 10 * 4);
  p.dy2 = take(R * 400 * es); p.dy1 = take(R * 200 * es);
  p.dz = take(R * n * 4);
  p.denc = take(B * 2 * n * es);
  p.dye2 = take(B * 200 * es); p.dye1 = take(B * 400 * es);
  if (x3) {
    // largest split A operand: dlogits [R, 784] (dgrad / wgrad of the last Linear); largest split B operand: g2 [R, 400]
    const long long b_max = R * 400 > 784ll * B ? R * 400 : 784ll * B;
    p.x3_a = take(3 * (R + 4) * 784 * 4);
    p.x3_b = take(3 * ((b_max > 784 * 400 ? b_max : 784 * 400) + 4 * 784) * 4);
  }
  p.bytes = off;
  return p;
}

struct Ptrs {
  char* ws;
  template <typename T>
  T* at(long long off) const {
    return reinterpret_cast<T*>(ws + off);
  }
};

// 3xTF32 mode of the current mvae_mnist_step call (set at its start; the helpers below stamp it on every GEMM)
struct X3Scratch {
  void* a = nullptr;
  void* b = nullptr;
};
thread_local X3Scratch g_x3;
void stamp_x3(GemmDesc& g) {
  if (g_x3.a != nullptr && g.kind == MVAE_F32) {
    g.x3 = 1;
    g.x3_a = g_x3.a;
    g.x3_b = g_x3.b;
  }
}

int gemm_fwd(int dtype, int M, int N, int K, const void* A, const void* W, void* C, int c_dtype, const float* bias,
             float* st_sum, float* st_sumsq, int rows_per_group, cudaStream_t st) {
  GemmDesc g;
  g.kind = dtype; g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = K; g.a_mn = 0;
  g.B = W; g.ldb = K; g.b_mn = 0;
  g.epi.kind = EPI_STORE;
  g.epi.C = C; g.epi.ldc = N; g.epi.c_dtype = c_dtype;
  g.epi.bias = bias;
  g.epi.stat0 = st_sum; g.epi.stat1 = st_sumsq;
  g.epi.rows_per_group = rows_per_group;
  stamp_x3(g);
  return launch_gemm(g, st);
}

// dX = dY * W with the ReLU/BatchNorm-backward statistics epilogue, or a plain store.
int gemm_dgrad(int dtype, int M, int Nin, int Kout, const void* dY, const void* W, void* dX, int c_dtype,
               const void* hpre, const float* mean, const float* rstd, const float* gamma, const float* beta, float* s0,
               float* s1, int rows_per_group, cudaStream_t st) {
  GemmDesc g;
  g.kind = dtype; g.M = M; g.N = Nin; g.K = Kout;  // contraction over the layer's OUTPUT features
  g.A = dY; g.lda = Kout; g.a_mn = 0;
  g.B = W; g.ldb = Nin; g.b_mn = 1;                // W is [Kout, Nin]: row index (Nin) contiguous -> MN-major
  g.epi.C = dX; g.epi.ldc = Nin; g.epi.c_dtype = c_dtype;
  g.epi.rows_per_group = rows_per_group;
  if (hpre != nullptr) {
    g.epi.kind = EPI_DGRAD_BN;
    g.epi.hpre = hpre; g.epi.ldh = Nin;
    g.epi.bn_mean = mean; g.epi.bn_rstd = rstd; g.epi.bn_gamma = gamma; g.epi.bn_beta = beta;
    g.epi.stat0 = s0; g.epi.stat1 = s1;
  } else {
    g.epi.kind = EPI_STORE;
  }
  stamp_x3(g);
  return launch_gemm(g, st);
}

// dW[Nout, Kin] += dY^T[Nout, rows] * X[rows, Kin]  (both operands MN-major, split-K with vector reductions)
int gemm_wgrad(int dtype, int rows, int Nout, int Kin, const void* dY, const void* X, float* dW, cudaStream_t st) {
  GemmDesc g;
  g.kind = dtype; g.M = Nout; g.N = Kin; g.K = rows;
  g.A = dY; g.lda = Nout; g.a_mn = 1;
  g.B = X; g.ldb = Kin; g.b_mn = 1;
  g.epi.kind = EPI_ATOMIC;
  g.epi.C = dW; g.epi.ldc = Kin; g.epi.c_dtype = MVAE_F32;
  stamp_x3(g);
  return launch_gemm(g, st);
}

}  // namespace
}  // namespace mvae

using namespace mvae;

namespace {
// A library-owned side stream per device: work that is off the critical path of the step (the per-label text
// encoder, the text decoder, every weight-gradient GEMM) is forked onto it and joined with events, which also
// works under CUDA-graph capture (the side stream is pulled into the capture by the event dependencies).
struct SideStream {
  cudaStream_t s = nullptr;
  cudaStream_t s3 = nullptr;   // second side stream: the text encoder's backward, one of the two last weight gradients
  cudaEvent_t ev[16];
  int next = 0;
  bool ok = false;
};
SideStream* side_stream_for_current_device() {
  static SideStream table[32];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
  SideStream* ss = &table[dev];
  if (!ss->ok) {
    if (cudaStreamCreateWithFlags(&ss->s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&ss->s3, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int i = 0; i < 16; ++i)
      if (cudaEventCreateWithFlags(&ss->ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ss->ok = true;
  }
  return ss;
}
// `to` waits for everything enqueued so far on `from`.
int stream_dep(SideStream* ss, cudaStream_t from, cudaStream_t to) {
  cudaEvent_t e = ss->ev[ss->next];
  ss->next = (ss->next + 1) % 16;
  MVAE_CUDA(cudaEventRecord(e, from));
  MVAE_CUDA(cudaStreamWaitEvent(to, e, 0));
  return 0;
}
}  // namespace

// Per-launch instrumentation of the step:
//   * g_launches counts every kernel launch the step enqueues (bench.py reports it as gpu_launches);
//   * MVAE_DEBUG_SYNC=1 synchronises after every launch and reports it (bring-up aid);
//   * profile mode (mvae_mnist_step_profile) brackets every launch with CUDA events on the launching stream
//     and returns the per-launch durations - this is how bench.py measures the kernels' time live.
static long long g_launches = 0;
struct StepProfile {
  bool on = false;
  int n = 0;
  cudaEvent_t ev[2 * 96];
  const char* label[96];
};
static StepProfile g_prof;

#define MVAE_STEP(call, lab)                                                           \
  do {                                                                                 \
    if (g_prof.on && g_prof.n < 96) cudaEventRecord(g_prof.ev[2 * g_prof.n], st);      \
    const int rc_ = (call);                                                            \
    g_pdl_next = 0; /* a programmatic-launch mark never outlives the launch it was set for */ \
    if (rc_) return 1;                                                                 \
    ++g_launches;                                                                      \
    if (g_prof.on && g_prof.n < 96) {                                                  \
      cudaEventRecord(g_prof.ev[2 * g_prof.n + 1], st);                                \
      g_prof.label[g_prof.n++] = lab;                                                  \
    }                                                                                  \
    if (debug_sync) {                                                                  \
      cudaError_t _e = cudaStreamSynchronize(st);                                      \
      fprintf(stderr, "[mvae] %s -> %s\n", lab, cudaGetErrorString(_e));               \
      fflush(stderr);                                                                  \
      if (_e != cudaSuccess) return ::mvae::cuda_fail(_e, lab, __FILE__, __LINE__);    \
    }                                                                                  \
  } while (0)

extern "C" {

int mvae_mnist_num_tensors(void) {
  TensorInfo tmp[kMaxTensors];
  return build_tensor_list(64, tmp);
}

int mvae_mnist_tensor_info(int n_latents, int index, mvae_tensor_info* out) {
  MVAE_REQUIRE(n_latents > 0 && n_latents % 2 == 0, "n_latents=%d must be positive and even", n_latents);
  MVAE_REQUIRE(out != nullptr, "tensor_info: null output");
  const Layout L = make_layout(n_latents);
  MVAE_REQUIRE(index >= 0 && index < L.count, "tensor index %d out of range", index);
  memset(out, 0, sizeof(*out));
  strncpy(out->name, L.info[index].name, sizeof(out->name) - 1);
  out->kind = L.info[index].kind;
  out->ndim = L.info[index].ndim;
  out->shape[0] = L.info[index].shape[0];
  out->shape[1] = L.info[index].shape[1];
  out->offset = L.offset[index];
  return 0;
}

int mvae_mnist_sizes(int n_latents, int batch, int dtype, mvae_mnist_size_info* out) {
  MVAE_REQUIRE(n_latents > 0 && n_latents % 4 == 0, "n_latents=%d must be a positive multiple of 4", n_latents);
  MVAE_REQUIRE(batch > 1, "batch=%d must be > 1 (train-mode BatchNorm)", batch);
  MVAE_REQUIRE(dtype == MVAE_DT_F32 || dtype == MVAE_DT_BF16 || dtype == MVAE_DT_F32X3, "bad dtype %d", dtype);
  const Layout L = make_layout(n_latents);
  const Plan p = make_plan(batch, n_latents, dtype);
  out->param_floats = L.param_floats;
  out->encoder_param_floats = L.enc_floats;
  out->buffer_floats = L.buffer_floats;
  out->num_bn = L.nbt_count;
  out->workspace_bytes = p.bytes;
  return 0;
}

int mvae_mnist_step(const mvae_mnist_step_args* a, void* stream_v) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  static const int debug_sync = env_int("MVAE_DEBUG_SYNC", 0);
  static const int use_side = env_int("MVAE_SIDE_STREAM", 1);
  MVAE_REQUIRE(a != nullptr, "mnist_step: null args");
  const bool x3 = a->dtype == MVAE_DT_F32X3;  // fp32 storage, error-compensated 3xTF32 GEMMs (shared split scratch: one stream)
  SideStream* ss = (use_side && !debug_sync && !g_prof.on && !x3) ? side_stream_for_current_device() : nullptr;
  cudaStream_t s2 = ss != nullptr ? ss->s : st;  // side stream (or the main one when disabled)
  cudaStream_t s3 = ss != nullptr ? ss->s3 : st; // second side stream
  auto dep = [&](cudaStream_t from, cudaStream_t to) -> int { return (ss != nullptr && from != to) ? stream_dep(ss, from, to) : 0; };
  const int B = a->batch, n = a->n_latents, dt = x3 ? MVAE_DT_F32 : a->dtype, G = a->n_terms;
  MVAE_REQUIRE(n > 0 && n % 4 == 0, "n_latents=%d must be a positive multiple of 4", n);
  MVAE_REQUIRE(B > 1, "batch=%d must be > 1", B);
  MVAE_REQUIRE(a->dtype == MVAE_DT_F32 || a->dtype == MVAE_DT_BF16 || x3, "bad dtype %d", a->dtype);
  MVAE_REQUIRE(G >= 1 && G <= kMaxGroups, "n_terms=%d out of range", G);
  MVAE_REQUIRE(a->params && a->workspace && a->buffers, "mnist_step: params / buffers / workspace missing");
  MVAE_REQUIRE(dt == MVAE_DT_F32 || a->params_bf16 != nullptr, "mnist_step: bf16 path needs the bf16 parameter mirror");
  const Layout L = make_layout(n);
  const Plan P = make_plan(B, n, a->dtype);
  MVAE_REQUIRE(a->workspace_bytes >= P.bytes, "workspace too small: %lld < %lld", (long long)a->workspace_bytes, P.bytes);
  MVAE_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "workspace must be 256-byte aligned");
  int n_img = 0, n_txt = 0;
  for (int g = 0; g < G; ++g) {
    MVAE_REQUIRE(a->term_type[g] >= 0 && a->term_type[g] <= 2, "bad term type");
    if (a->term_type[g] != MVAE_TERM_TEXT) ++n_img;
    if (a->term_type[g] != MVAE_TERM_IMAGE) ++n_txt;
  }
  MVAE_REQUIRE(a->image != nullptr && a->text != nullptr, "mnist_step: image and text are both needed (targets)");
  const bool bwd = a->do_backward != 0;
  MVAE_REQUIRE(!bwd || a->grads != nullptr, "mnist_step: backward needs the gradient buffer");
  // phase 0: forward (+ backward when do_backward); phase 2: backward only, driven by upstream gradients of the
  // module outputs (the forward of the same workspace must have run before - autograd path of MVAE.forward)
  // phase 3: forward + the decoder-side backward (through the tail backward); phase 4: the encoder-side backward
  // only (text + image encoder) - data-parallel training all-reduces the decoder bucket between the two
  const bool fwd = a->phase != 2 && a->phase != 4;
  const bool module_bwd = a->phase == 2;
  const bool bwd_dec = a->phase != 4;
  const bool bwd_enc = a->phase != 3;
  const bool training = a->eval_mode == 0;
  MVAE_REQUIRE(fwd || bwd, "mnist_step: phase 2 needs do_backward");
  MVAE_REQUIRE(training || !bwd, "mnist_step: backward needs train mode");
  MVAE_REQUIRE(!module_bwd || a->d_recon_image == nullptr || a->out_recon_image != nullptr,
               "mnist_step: phase 2 needs the saved recon_image probabilities");
  const bool decode_only = a->z_in != nullptr;
  if (decode_only) n_img = n_txt = 0;  // latents are given: no encoders, no PoE
  // bf16 training steps whose slabs are whole and co-resident run the slab-persistent chain kernels (chain.cu): one launch
  // per encoder / decoder direction instead of a GEMM + BatchNorm launch per layer.  MVAE_CHAIN=0 restores the per-layer path.
  static const int chain_env = env_int("MVAE_CHAIN", 1);
  const bool use_chain = chain_env != 0 && a->dtype == MVAE_DT_BF16 && training && bwd && !module_bwd && !decode_only &&
                         chain_supported(B, G, n);

  // Programmatic dependent launch along the critical path of the chain step (enc_fwd -> tail -> dec_fwd -> dec_bwd -> tail ->
  // enc_bwd): each of these kernels starts with griddep_wait, so its CTAs may be scheduled - and run their prologue - while the
  // previous kernel of the stream drains.  Cross-stream joins stay ordinary (full) dependencies.
  static const int pdl_env = env_int("MVAE_PDL", 14);   // bits: 0 enc_fwd, 1 tail_fwd, 2 dec_fwd, 3 dec_bwd, 4 tail_bwd, 5 enc_bwd
  const bool use_pdl = pdl_env != 0 && use_chain && !g_prof.on && !debug_sync;
  auto pdl_mark = [&](int bit) { g_pdl_next = (use_pdl && ((pdl_env >> bit) & 1)) ? 1 : 0; };   // MVAE_PDL: one bit per launch
  // one call = the whole step with the optimizer: the decoder bucket's Adam runs beside the encoder backward
  const bool adam_split = bwd && bwd_dec && bwd_enc && a->do_adam && !module_bwd && ss != nullptr && L.enc_floats % 4 == 0;

  const Ptrs W{static_cast<char*>(a->workspace)};
  g_x3.a = x3 ? W.at<void>(P.x3_a) : nullptr;
  g_x3.b = x3 ? W.at<void>(P.x3_b) : nullptr;
  float* prm = a->params;
  // GEMM weight operands: fp32 master (tf32 path) or the bf16 mirror (same offsets)
  auto wop = [&](const char* name) -> const void* {
    const long long o = L.find(name);
    return dt == MVAE_DT_F32 ? static_cast<const void*>(prm + o)
                             : static_cast<const void*>(static_cast<const __nv_bfloat16*>(a->params_bf16) + o);
  };
  auto pf = [&](const char* name) -> float* { return prm + L.find(name); };
  auto gf = [&](const char* name) -> float* { return a->grads ? a->grads + L.find(name) : nullptr; };
  auto bf = [&](const char* name) -> float* { return a->buffers + L.find(name); };
  const int R = G * B;

  // ---- step start: device step counter, num_batches_tracked, accumulators, (optionally) gradients
  {
    // BN order in the layout: ie.1, ie.4, id.1, id.4, te.1, td.1
    const long long t_ = training ? 1 : 0;
    const long long inc[6] = {t_ * n_img, t_ * n_img, t_ * G, t_ * G, t_ * n_txt, t_ * G};
    // two device clocks: the noise counter ticks on every forward-type call (Philox stream), Adam's bias-correction
    // clock only on calls that open an optimizer step (do_adam, or advance_adam_step for split / accumulated steps)
    int* adam_tick = (a->do_adam || a->advance_adam_step) ? a->adam_step : nullptr;
    int* noise_tick = a->noise_step;
    if (a->noise_step == nullptr && adam_tick == nullptr) noise_tick = a->adam_step;  // legacy callers: one shared clock
    if (fwd) MVAE_STEP(launch_step_prep(adam_tick, noise_tick, W.at<float>(P.acc_off), P.acc_floats,
                                        reinterpret_cast<long long*>(a->num_batches_tracked), inc, st), "launch_step_prep");
  }
  // zero_grad: nothing adds into the gradient buffer before the decoders run, so the memset goes beside the encoder forward
  // (side stream, joined after the tail forward)
  // Launch ORDER matters beside the dependencies: a chain kernel needs whole SMs (227 KB of shared memory, cooperative launch), and
  // an SM that is running blocks of another kernel cannot take one until they drain.  So at every fork the event is recorded
  // first, then the chain kernel is enqueued, and only then its side-stream siblings (they fill the SMs the chain leaves free).
  bool grads_zeroing = false;
  const bool want_grad_zero = fwd && bwd && a->zero_grad;
  if (want_grad_zero && dep(st, s2)) return 1;
  auto zero_grads_aside = [&]() -> int {
    if (want_grad_zero && !grads_zeroing) {
      MVAE_CUDA(cudaMemsetAsync(a->grads, 0, static_cast<size_t>(L.param_floats) * 4, s2));
      grads_zeroing = s2 != st;
    }
    return 0;
  };

  float* st_e1 = W.at<float>(P.st_e1); float* st_e2 = W.at<float>(P.st_e2);
  float* st_d1 = W.at<float>(P.st_d1); float* st_d2 = W.at<float>(P.st_d2); float* st_t1 = W.at<float>(P.st_t1);
  float* sb_e1 = W.at<float>(P.sb_e1); float* sb_e2 = W.at<float>(P.sb_e2);
  float* sb_d1 = W.at<float>(P.sb_d1); float* sb_d2 = W.at<float>(P.sb_d2); float* sb_t1 = W.at<float>(P.sb_t1);
  float* sv_e1 = W.at<float>(P.sv_e1); float* sv_e2 = W.at<float>(P.sv_e2);
  float* sv_d1 = W.at<float>(P.sv_d1); float* sv_d2 = W.at<float>(P.sv_d2);
  float* losses = W.at<float>(P.losses);
  unsigned int* bars = W.at<unsigned int>(P.bars);
  const float mom = 0.1f, bn_eps = 1e-5f;

  // ================================================================ forward
  TextEncArgs te;
  te.B = B; te.n = n; te.labels = reinterpret_cast<const long long*>(a->text);
  te.emb = pf("text_encoder.net.0.weight");
  te.gamma = pf("text_encoder.net.1.weight"); te.beta = pf("text_encoder.net.1.bias");
  te.w = pf("text_encoder.net.3.weight"); te.b = pf("text_encoder.net.3.bias");
  te.running_mean = bf("text_encoder.net.1.running_mean"); te.running_var = bf("text_encoder.net.1.running_var");
  te.updates = n_txt; te.momentum = mom; te.bn_eps = bn_eps; te.training = training ? 1 : 0;
  te.table = W.at<float>(P.txt_table); te.save = W.at<float>(P.txt_save);
  // Streams: st carries the critical path; s3 the text networks (per-label encoder, decoder, their backward); s2 the
  // weight-gradient GEMMs.  A join waits for everything queued on the joined stream, so what the critical path joins on
  // (the text decoder before the tail backward) must not share a stream with work that is enqueued earlier but may run later.
  if (dep(st, s3)) return 1;  // fork: the per-label text encoder runs beside the image encoder
  auto text_encoder_aside = [&]() -> int {
    if (fwd && n_txt > 0) MVAE_STEP(launch_textenc_forward(te, s3), "launch_textenc_forward");
    return 0;
  };
  if (!(n_img > 0 && use_chain && fwd)) {
    if (text_encoder_aside()) return 1;
    if (zero_grads_aside()) return 1;
  }

  unsigned int* chain_err = bars + 31;
  auto wb = [&](const char* name) -> const __nv_bfloat16* { return static_cast<const __nv_bfloat16*>(a->params_bf16) + L.find(name); };
  if (n_img > 0 && use_chain && fwd) {
    ChainEncFwd ce;
    ce.B = B; ce.n = n; ce.bn_updates = n_img;
    ce.image = static_cast<const __nv_bfloat16*>(a->image);
    ce.w1 = wb("image_encoder.net.0.weight"); ce.w2 = wb("image_encoder.net.3.weight"); ce.w3 = wb("image_encoder.net.6.weight");
    ce.b1 = pf("image_encoder.net.0.bias"); ce.b2 = pf("image_encoder.net.3.bias"); ce.b3 = pf("image_encoder.net.6.bias");
    ce.gamma1 = pf("image_encoder.net.1.weight"); ce.beta1 = pf("image_encoder.net.1.bias");
    ce.gamma2 = pf("image_encoder.net.4.weight"); ce.beta2 = pf("image_encoder.net.4.bias");
    ce.st1 = st_e1; ce.st2 = st_e2; ce.sv1 = sv_e1; ce.sv2 = sv_e2;
    ce.rm1 = bf("image_encoder.net.1.running_mean"); ce.rv1 = bf("image_encoder.net.1.running_var");
    ce.rm2 = bf("image_encoder.net.4.running_mean"); ce.rv2 = bf("image_encoder.net.4.running_var");
    ce.counters = bars + 0;
    ce.h1pre = W.at<__nv_bfloat16>(P.h1pre); ce.h1 = W.at<__nv_bfloat16>(P.h1);
    ce.h2pre = W.at<__nv_bfloat16>(P.h2pre); ce.h2 = W.at<__nv_bfloat16>(P.h2);
    ce.enc = W.at<float>(P.enc);
    ce.err = chain_err;
    pdl_mark(0);
    MVAE_STEP(launch_chain_enc_fwd(ce, st), "chain_enc_fwd");
    if (text_encoder_aside()) return 1;
    if (zero_grads_aside()) return 1;
  }
  if (n_img > 0 && !use_chain) {
    // ImageEncoder (mnist/model.py:99-117), once for all terms that use it: every Linear publishes its column sums from
    // the GEMM epilogue, launch_bn_forward applies BatchNorm + ReLU
    if (fwd) MVAE_STEP(gemm_fwd(dt, B, 400, 784, a->image, wop("image_encoder.net.0.weight"), W.at<void>(P.h1pre), dt,
                 pf("image_encoder.net.0.bias"), training ? st_e1 : nullptr, training ? st_e1 + 400 : nullptr, 1 << 30, st), "gemm_fwd:image_encoder.net.0.weight#3");
    if (fwd) MVAE_STEP(launch_bn_forward(dt, W.at<void>(P.h1pre), W.at<void>(P.h1), B, 400, B, training ? st_e1 : nullptr, st_e1 + 400,
                          pf("image_encoder.net.1.weight"), pf("image_encoder.net.1.bias"), sv_e1, sv_e1 + 400,
                          bf("image_encoder.net.1.running_mean"), bf("image_encoder.net.1.running_var"), n_img, mom,
                          bn_eps, 1, st), "launch_bn_forward:image_encoder.net.1.weight#4");
    if (fwd) MVAE_STEP(gemm_fwd(dt, B, 200, 400, W.at<void>(P.h1), wop("image_encoder.net.3.weight"), W.at<void>(P.h2pre), dt,
                 pf("image_encoder.net.3.bias"), training ? st_e2 : nullptr, training ? st_e2 + 200 : nullptr, 1 << 30, st), "gemm_fwd:image_encoder.net.3.weight#5");
    if (fwd) MVAE_STEP(launch_bn_forward(dt, W.at<void>(P.h2pre), W.at<void>(P.h2), B, 200, B, training ? st_e2 : nullptr, st_e2 + 200,
                          pf("image_encoder.net.4.weight"), pf("image_encoder.net.4.bias"), sv_e2, sv_e2 + 200,
                          bf("image_encoder.net.4.running_mean"), bf("image_encoder.net.4.running_var"), n_img, mom,
                          bn_eps, 1, st), "launch_bn_forward:image_encoder.net.4.weight#6");
    if (fwd) MVAE_STEP(gemm_fwd(dt, B, 2 * n, 200, W.at<void>(P.h2), wop("image_encoder.net.6.weight"), W.at<void>(P.enc), MVAE_F32,
                 pf("image_encoder.net.6.bias"), nullptr, nullptr, 1 << 30, st), "gemm_fwd:image_encoder.net.6.weight#7");
  }
  if (dep(s3, st)) return 1;  // join: the tail needs both experts
  TailArgs ta;
  ta.B = B; ta.n = n; ta.G = G;
  for (int g = 0; g < G; ++g) {
    ta.group_type[g] = a->term_type[g];
    ta.kl_weight[g] = a->kl_weight[g];
  }
  ta.poe_mode = a->poe_mode; ta.prior_expert = a->prior_expert; ta.poe_eps = a->poe_eps;
  ta.z_dtype = dt;
  ta.enc_img = n_img > 0 ? W.at<float>(P.enc) : nullptr;
  ta.txt_table = n_txt > 0 ? W.at<float>(P.txt_table) : nullptr;
  ta.labels = reinterpret_cast<const long long*>(a->text);
  ta.eps = a->eps; ta.seed = a->seed; ta.step_ptr = a->noise_step != nullptr ? a->noise_step : a->adam_step; ta.training = training ? 1 : 0;
  ta.z_in = a->z_in;
  ta.noise_buf = W.at<float>(P.noise);
  ta.wt1 = pf("text_decoder.net.0.weight"); ta.bt1 = pf("text_decoder.net.0.bias");
  ta.z = W.at<void>(P.z); ta.mu = a->out_mu; ta.logvar = a->out_logvar;
  ta.kl = losses + 2 * kMaxGroups;
  ta.t1pre = W.at<float>(P.t1pre); ta.t1_sum = st_t1; ta.t1_sumsq = st_t1 + G * 10;
  if (fwd) {
    pdl_mark(1);
    MVAE_STEP(launch_tail_forward(ta, st), "launch_tail_forward#8");
  }
  if (grads_zeroing && dep(s2, st)) return 1;   // join: the gradient buffer is clear before the first kernel that adds into it

  TextDecArgs td;
  td.B = B; td.G = G;
  td.t1pre = W.at<float>(P.t1pre); td.t1_sum = st_t1; td.t1_sumsq = st_t1 + G * 10;
  td.gamma = pf("text_decoder.net.1.weight"); td.beta = pf("text_decoder.net.1.bias");
  td.running_mean = bf("text_decoder.net.1.running_mean"); td.running_var = bf("text_decoder.net.1.running_var");
  td.momentum = mom; td.bn_eps = bn_eps; td.training = training ? 1 : 0;
  td.w2 = pf("text_decoder.net.3.weight"); td.b2 = pf("text_decoder.net.3.bias");
  td.labels = reinterpret_cast<const long long*>(a->text);
  for (int t = 0; t < G; ++t) td.ce_scale[t] = a->lambda_text[t] / static_cast<float>(B);
  td.fused_loss = module_bwd ? 0 : 1; td.backward = (bwd && !module_bwd) ? 1 : 0;
  td.dlogp_up = a->d_recon_text;
  td.logp = a->out_recon_text; td.ce = losses + kMaxGroups;
  td.dyhat = W.at<float>(P.dyt); td.s0 = sb_t1; td.s1 = sb_t1 + G * 10;
  td.d_w2 = gf("text_decoder.net.3.weight"); td.d_b2 = gf("text_decoder.net.3.bias");
  if (dep(st, s3)) return 1;  // fork: text decoder beside the image decoder
  if (fwd && !use_chain) MVAE_STEP(launch_textdec(td, s3), "launch_textdec");

  if (use_chain && fwd) {
    ChainDecFwd cd;
    cd.B = B; cd.n = n; cd.G = G;
    cd.z = W.at<__nv_bfloat16>(P.z);
    cd.w1 = wb("image_decoder.net.0.weight"); cd.w2 = wb("image_decoder.net.3.weight"); cd.w3 = wb("image_decoder.net.6.weight");
    cd.b1 = pf("image_decoder.net.0.bias"); cd.b2 = pf("image_decoder.net.3.bias"); cd.b3 = pf("image_decoder.net.6.bias");
    cd.gamma1 = pf("image_decoder.net.1.weight"); cd.beta1 = pf("image_decoder.net.1.bias");
    cd.gamma2 = pf("image_decoder.net.4.weight"); cd.beta2 = pf("image_decoder.net.4.bias");
    cd.st1 = st_d1; cd.st2 = st_d2; cd.sv1 = sv_d1; cd.sv2 = sv_d2;
    cd.rm1 = bf("image_decoder.net.1.running_mean"); cd.rv1 = bf("image_decoder.net.1.running_var");
    cd.rm2 = bf("image_decoder.net.4.running_mean"); cd.rv2 = bf("image_decoder.net.4.running_var");
    cd.counters = bars + 2;
    cd.g1pre = W.at<__nv_bfloat16>(P.g1pre); cd.g1 = W.at<__nv_bfloat16>(P.g1);
    cd.g2pre = W.at<__nv_bfloat16>(P.g2pre); cd.g2 = W.at<__nv_bfloat16>(P.g2);
    cd.dlog = W.at<__nv_bfloat16>(P.dlog);
    cd.probs = static_cast<__nv_bfloat16*>(a->out_recon_image);
    cd.image = static_cast<const __nv_bfloat16*>(a->image);
    for (int t = 0; t < G; ++t) cd.bce_scale[t] = a->lambda_image[t] / (static_cast<float>(B) * 784.f);
    cd.loss = losses;
    cd.dbias3 = gf("image_decoder.net.6.bias");
    cd.err = chain_err;
    pdl_mark(2);
    MVAE_STEP(launch_chain_dec_fwd(cd, st), "chain_dec_fwd");
  }
  if (fwd && use_chain) MVAE_STEP(launch_textdec(td, s3), "launch_textdec");
  if (!use_chain) {
  // ImageDecoder (mnist/model.py:120-135) on the stacked [G*B, n] latents, per-term BN statistics
  if (fwd) MVAE_STEP(gemm_fwd(dt, R, 200, n, W.at<void>(P.z), wop("image_decoder.net.0.weight"), W.at<void>(P.g1pre), dt,
               pf("image_decoder.net.0.bias"), training ? st_d1 : nullptr, training ? st_d1 + G * 200 : nullptr, B, st), "gemm_fwd:image_decoder.net.0.weight#9");
  if (fwd) MVAE_STEP(launch_bn_forward(dt, W.at<void>(P.g1pre), W.at<void>(P.g1), R, 200, B, training ? st_d1 : nullptr, st_d1 + G * 200,
                        pf("image_decoder.net.1.weight"), pf("image_decoder.net.1.bias"), sv_d1, sv_d1 + G * 200,
                        bf("image_decoder.net.1.running_mean"), bf("image_decoder.net.1.running_var"), 1, mom, bn_eps,
                        1, st), "launch_bn_forward:image_decoder.net.1.weight#10");
  if (fwd) MVAE_STEP(gemm_fwd(dt, R, 400, 200, W.at<void>(P.g1), wop("image_decoder.net.3.weight"), W.at<void>(P.g2pre), dt,
               pf("image_decoder.net.3.bias"), training ? st_d2 : nullptr, training ? st_d2 + G * 400 : nullptr, B, st), "gemm_fwd:image_decoder.net.3.weight#11");
  if (fwd) MVAE_STEP(launch_bn_forward(dt, W.at<void>(P.g2pre), W.at<void>(P.g2), R, 400, B, training ? st_d2 : nullptr, st_d2 + G * 400,
                        pf("image_decoder.net.4.weight"), pf("image_decoder.net.4.bias"), sv_d2, sv_d2 + G * 400,
                        bf("image_decoder.net.4.running_mean"), bf("image_decoder.net.4.running_var"), 1, mom, bn_eps,
                        1, st), "launch_bn_forward:image_decoder.net.4.weight#12");
  {
    // last Linear + sigmoid + BCE (mnist/model.py:130,135 + mnist/train.py:70) in one kernel
    GemmDesc g;
    g.kind = dt; g.M = R; g.N = 784; g.K = 400;
    g.A = W.at<void>(P.g2); g.lda = 400; g.a_mn = 0;
    g.B = wop("image_decoder.net.6.weight"); g.ldb = 400; g.b_mn = 0;
    g.epi.kind = EPI_BCE;
    g.epi.C = W.at<void>(P.dlog); g.epi.ldc = 784; g.epi.c_dtype = dt;
    g.epi.bias = pf("image_decoder.net.6.bias");
    g.epi.stat0 = bwd ? gf("image_decoder.net.6.bias") : nullptr;
    g.epi.rows_per_group = B;
    g.epi.target = a->image; g.epi.ldt = 784; g.epi.target_rows = B;
    for (int t = 0; t < G; ++t) g.epi.bce_scale[t] = a->lambda_image[t] / (static_cast<float>(B) * 784.f);
    g.epi.loss = losses;
    g.epi.probs = a->out_recon_image;
    stamp_x3(g);
    if (fwd) MVAE_STEP(launch_gemm(g, st), "gemm_fwd_bce:image_decoder.net.6.weight");
  }
  }  // !use_chain
  if (dep(st, s2)) return 1;  // the side stream may start the weight gradients once dlogits exist
  // ---- losses out: [G][4] = total, bce, ce, kl (all three partial sums are final once the decoders' forward is done;
  //      on the side stream the tiny kernel is off the critical path)
  bool losses_packed = false;
  if (a->out_losses != nullptr && fwd && ss != nullptr) {
    if (dep(st, s3)) return 1;   // behind the text decoder (cross entropy) and the image decoder (BCE)
    MVAE_STEP(launch_loss_pack(losses, a->out_losses, G, s3), "launch_loss_pack#32");
    losses_packed = true;
  }
  // ================================================================ backward
  bool enc_chain_launched = false;
  auto launch_encoder_chain_backward = [&]() -> int {
    ChainEncBwd cb;
    cb.B = B; cb.n = n;
    cb.denc = W.at<__nv_bfloat16>(P.denc);
    cb.w2 = wb("image_encoder.net.3.weight"); cb.w3 = wb("image_encoder.net.6.weight");
    cb.gamma1 = pf("image_encoder.net.1.weight"); cb.beta1 = pf("image_encoder.net.1.bias");
    cb.gamma2 = pf("image_encoder.net.4.weight"); cb.beta2 = pf("image_encoder.net.4.bias");
    cb.sb1 = sb_e1; cb.sb2 = sb_e2; cb.sv1 = sv_e1; cb.sv2 = sv_e2;
    cb.counters = bars + 14;
    cb.h1pre = W.at<__nv_bfloat16>(P.h1pre); cb.h2pre = W.at<__nv_bfloat16>(P.h2pre);
    cb.dye2 = W.at<__nv_bfloat16>(P.dye2); cb.dye1 = W.at<__nv_bfloat16>(P.dye1);
    cb.dgamma1 = gf("image_encoder.net.1.weight"); cb.dbeta1 = gf("image_encoder.net.1.bias");
    cb.dgamma2 = gf("image_encoder.net.4.weight"); cb.dbeta2 = gf("image_encoder.net.4.bias");
    cb.err = chain_err;
    // phase 4 = the data-parallel trainer's encoder side: the decoder bucket's exchange kernel runs beside this launch and
    // needs SMs of its own
    static const int dp_parts = env_int("MVAE_CHAIN_SPLIT_DP", 2);
    cb.max_parts = a->phase == 4 ? dp_parts : 4;
    pdl_mark(5);
    MVAE_STEP(launch_chain_enc_bwd(cb, st), "chain_enc_bwd");
    enc_chain_launched = true;
    return 0;
  };
  if (bwd && bwd_dec) {
    if (module_bwd) {
      // autograd path: dlogits = d(recon_image) * p * (1 - p) from the probabilities the forward returned
      // (+ the last Linear's bias gradient), and the text decoder's backward from d(log-probs)
      if (a->d_recon_image != nullptr) {
        MVAE_STEP(launch_sigmoid_backward(dt, a->d_recon_image, a->out_recon_image, W.at<void>(P.dlog), R, 784,
                                          gf("image_decoder.net.6.bias"), st), "launch_sigmoid_backward");
      } else {
        MVAE_CUDA(cudaMemsetAsync(W.at<void>(P.dlog), 0, static_cast<size_t>(R) * 784 * (dt == MVAE_DT_F32 ? 4 : 2), st));
      }
      td.backward = 1;
      MVAE_STEP(launch_textdec(td, st), "launch_textdec(bwd)");
      if (dep(st, s2)) return 1;
    }
    if (use_chain) {
      // ---- image decoder: the whole dgrad / ReLU / BatchNorm-backward chain in one launch; weight gradients beside it
      ChainDecBwd cb;
      cb.B = B; cb.n = n; cb.G = G;
      cb.dlog = W.at<__nv_bfloat16>(P.dlog);
      cb.w1 = wb("image_decoder.net.0.weight"); cb.w2 = wb("image_decoder.net.3.weight"); cb.w3 = wb("image_decoder.net.6.weight");
      cb.gamma1 = pf("image_decoder.net.1.weight"); cb.beta1 = pf("image_decoder.net.1.bias");
      cb.gamma2 = pf("image_decoder.net.4.weight"); cb.beta2 = pf("image_decoder.net.4.bias");
      cb.sb1 = sb_d1; cb.sb2 = sb_d2; cb.sv1 = sv_d1; cb.sv2 = sv_d2;
      cb.counters = bars + 8;
      cb.g1pre = W.at<__nv_bfloat16>(P.g1pre); cb.g2pre = W.at<__nv_bfloat16>(P.g2pre);
      cb.dy2 = W.at<__nv_bfloat16>(P.dy2); cb.dy1 = W.at<__nv_bfloat16>(P.dy1);
      cb.dz = W.at<float>(P.dz);
      cb.dgamma1 = gf("image_decoder.net.1.weight"); cb.dbeta1 = gf("image_decoder.net.1.bias");
      cb.dgamma2 = gf("image_decoder.net.4.weight"); cb.dbeta2 = gf("image_decoder.net.4.bias");
      cb.err = chain_err;
      pdl_mark(3);
      MVAE_STEP(launch_chain_dec_bwd(cb, st), "chain_dec_bwd");
      MVAE_STEP(gemm_wgrad(dt, R, 784, 400, W.at<void>(P.dlog), W.at<void>(P.g2), gf("image_decoder.net.6.weight"), s2), "gemm_wgrad:image_decoder.net.6.weight#16");
      if (dep(st, s2)) return 1;
      MVAE_STEP(gemm_wgrad(dt, R, 400, 200, W.at<void>(P.dy2), W.at<void>(P.g1), gf("image_decoder.net.3.weight"), s2), "gemm_wgrad:image_decoder.net.3.weight#19");
      MVAE_STEP(gemm_wgrad(dt, R, 200, n, W.at<void>(P.dy1), W.at<void>(P.z), gf("image_decoder.net.0.weight"), s2), "gemm_wgrad:image_decoder.net.0.weight#22");
    } else {
    // ---- image decoder
    MVAE_STEP(gemm_dgrad(dt, R, 400, 784, W.at<void>(P.dlog), wop("image_decoder.net.6.weight"), W.at<void>(P.dy2), dt,
                   W.at<void>(P.g2pre), sv_d2, sv_d2 + G * 400, pf("image_decoder.net.4.weight"),
                   pf("image_decoder.net.4.bias"), sb_d2, sb_d2 + G * 400, B, st), "gemm_dgrad:image_decoder.net.6.weight#15");
    MVAE_STEP(gemm_wgrad(dt, R, 784, 400, W.at<void>(P.dlog), W.at<void>(P.g2), gf("image_decoder.net.6.weight"), s2), "gemm_wgrad:image_decoder.net.6.weight#16");
    MVAE_STEP(launch_bn_backward(dt, W.at<void>(P.dy2), W.at<void>(P.g2pre), W.at<void>(P.dy2), R, 400, B, sb_d2,
                           sb_d2 + G * 400, sv_d2, sv_d2 + G * 400, pf("image_decoder.net.4.weight"),
                           gf("image_decoder.net.4.weight"), gf("image_decoder.net.4.bias"), st), "launch_bn_backward:image_decoder.net.4.weight#17");
    MVAE_STEP(gemm_dgrad(dt, R, 200, 400, W.at<void>(P.dy2), wop("image_decoder.net.3.weight"), W.at<void>(P.dy1), dt,
                   W.at<void>(P.g1pre), sv_d1, sv_d1 + G * 200, pf("image_decoder.net.1.weight"),
                   pf("image_decoder.net.1.bias"), sb_d1, sb_d1 + G * 200, B, st), "gemm_dgrad:image_decoder.net.3.weight#18");
    if (dep(st, s2)) return 1;
    MVAE_STEP(gemm_wgrad(dt, R, 400, 200, W.at<void>(P.dy2), W.at<void>(P.g1), gf("image_decoder.net.3.weight"), s2), "gemm_wgrad:image_decoder.net.3.weight#19");
    MVAE_STEP(launch_bn_backward(dt, W.at<void>(P.dy1), W.at<void>(P.g1pre), W.at<void>(P.dy1), R, 200, B, sb_d1,
                           sb_d1 + G * 200, sv_d1, sv_d1 + G * 200, pf("image_decoder.net.1.weight"),
                           gf("image_decoder.net.1.weight"), gf("image_decoder.net.1.bias"), st), "launch_bn_backward:image_decoder.net.1.weight#20");
    MVAE_STEP(gemm_dgrad(dt, R, n, 200, W.at<void>(P.dy1), wop("image_decoder.net.0.weight"), W.at<void>(P.dz), MVAE_F32,
                   nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1 << 30, st), "gemm_dgrad:image_decoder.net.0.weight#21");
    if (dep(st, s2)) return 1;
    MVAE_STEP(gemm_wgrad(dt, R, 200, n, W.at<void>(P.dy1), W.at<void>(P.z), gf("image_decoder.net.0.weight"), s2), "gemm_wgrad:image_decoder.net.0.weight#22");
    }  // !use_chain

    // ---- tail backward (text decoder front + reparametrize + KL + PoE)
    ta.dz = W.at<float>(P.dz);
    ta.dmu_up = a->d_mu; ta.dlogvar_up = a->d_logvar;
    ta.t1_dyhat = W.at<float>(P.dyt); ta.t1_s0 = sb_t1; ta.t1_s1 = sb_t1 + G * 10;
    ta.t1_gamma = pf("text_decoder.net.1.weight");
    ta.d_enc = W.at<void>(P.denc);
    ta.d_enc_bias = gf("image_encoder.net.6.bias");
    ta.d_txt_table = W.at<float>(P.d_txt_table);
    ta.d_wt1 = gf("text_decoder.net.0.weight");
    ta.d_t1_gamma = gf("text_decoder.net.1.weight"); ta.d_t1_beta = gf("text_decoder.net.1.bias");
    if (dep(s3, st)) return 1;  // join: text decoder results
    pdl_mark(4);
    MVAE_STEP(launch_tail_backward(ta, st), "launch_tail_backward");
    if (dep(st, s2)) return 1;  // fork: encoder weight gradients
    if (dep(st, s3)) return 1;  // fork: text encoder backward
    // (the encoder chain goes in before its side-stream siblings: see "launch ORDER" above)
    if (bwd_enc && n_img > 0 && use_chain && launch_encoder_chain_backward()) return 1;
    // Every decoder-side gradient is final now (the decoder weight gradients sit ahead on the side stream, everything else
    // was joined into the tail backward): the decoder bucket is updated on the side stream while the encoder chain runs,
    // and only the encoder bucket is left for the end of the step.
    if (adam_split) {
      MVAE_REQUIRE(a->adam_m && a->adam_v && a->adam_step, "mnist_step: Adam state missing");
      const long long e0 = L.enc_floats, nd = L.param_floats - L.enc_floats;
      MVAE_STEP(launch_adam(prm + e0, a->grads + e0, a->adam_m + e0, a->adam_v + e0,
                            a->params_bf16 ? static_cast<__nv_bfloat16*>(a->params_bf16) + e0 : nullptr, nd, a->lr, a->beta1, a->beta2,
                            a->adam_eps, a->adam_step, a->grad_scale, 0, s2), "launch_adam(decoders)");
    }
  }
  long long e1w_floats = 0, adam_tail_floats = 0;   // > 0: the final Adam launch only covers [0, adam_tail_floats)
  if (bwd && bwd_enc) {
    if (!bwd_dec && dep(st, s2)) return 1;
    if (!bwd_dec && dep(st, s3)) return 1;
    // ---- text encoder
    if (n_txt > 0) {
      te.d_table = W.at<float>(P.d_txt_table);
      te.d_emb = gf("text_encoder.net.0.weight");
      te.d_gamma = gf("text_encoder.net.1.weight"); te.d_beta = gf("text_encoder.net.1.bias");
      te.d_w = gf("text_encoder.net.3.weight"); te.d_b = gf("text_encoder.net.3.bias");
      MVAE_STEP(launch_textenc_backward(te, s3), "launch_textenc_backward");
    }
    // ---- image encoder
    if (n_img > 0 && use_chain) {
      MVAE_STEP(gemm_wgrad(dt, B, 2 * n, 200, W.at<void>(P.denc), W.at<void>(P.h2), gf("image_encoder.net.6.weight"), s2), "gemm_wgrad:image_encoder.net.6.weight#26");
      if (!enc_chain_launched && launch_encoder_chain_backward()) return 1;
      if (dep(st, s2)) return 1;
      if (dep(st, s3)) return 1;
      // the last two weight gradients are what is left of the step: side by side on the two side streams
      if (adam_split && dep(s2, s3)) return 1;   // (the weight gradient of the last encoder Linear sits on s2)
      MVAE_STEP(gemm_wgrad(dt, B, 400, 784, W.at<void>(P.dye1), a->image, gf("image_encoder.net.0.weight"), s2), "gemm_wgrad:image_encoder.net.0.weight#31");
      MVAE_STEP(gemm_wgrad(dt, B, 200, 400, W.at<void>(P.dye2), W.at<void>(P.h1), gf("image_encoder.net.3.weight"), s3), "gemm_wgrad:image_encoder.net.3.weight#29");
      // every encoder gradient except the first Linear's weight is final once that GEMM is done: their Adam runs here, beside
      // the last weight gradient, and only the first Linear's weight (the head of the encoder bucket) is left for the end
      e1w_floats = L.find("image_encoder.net.0.weight") == 0 ? round_up(400ll * 784, kAlignFloats) : 0;
      if (adam_split && e1w_floats > 0) {
        const long long e0 = e1w_floats, ne = L.enc_floats - e1w_floats;
        MVAE_STEP(launch_adam(prm + e0, a->grads + e0, a->adam_m + e0, a->adam_v + e0,
                              a->params_bf16 ? static_cast<__nv_bfloat16*>(a->params_bf16) + e0 : nullptr, ne, a->lr, a->beta1, a->beta2,
                              a->adam_eps, a->adam_step, a->grad_scale, 0, s3), "launch_adam(encoders but the first weight)");
        adam_tail_floats = e1w_floats;
      }
    }
    if (n_img > 0 && !use_chain) {
      MVAE_STEP(gemm_dgrad(dt, B, 200, 2 * n, W.at<void>(P.denc), wop("image_encoder.net.6.weight"), W.at<void>(P.dye2), dt,
                     W.at<void>(P.h2pre), sv_e2, sv_e2 + 200, pf("image_encoder.net.4.weight"),
                     pf("image_encoder.net.4.bias"), sb_e2, sb_e2 + 200, 1 << 30, st), "gemm_dgrad:image_encoder.net.6.weight#25");
      MVAE_STEP(gemm_wgrad(dt, B, 2 * n, 200, W.at<void>(P.denc), W.at<void>(P.h2), gf("image_encoder.net.6.weight"), s2), "gemm_wgrad:image_encoder.net.6.weight#26");
      MVAE_STEP(launch_bn_backward(dt, W.at<void>(P.dye2), W.at<void>(P.h2pre), W.at<void>(P.dye2), B, 200, B, sb_e2,
                             sb_e2 + 200, sv_e2, sv_e2 + 200, pf("image_encoder.net.4.weight"),
                             gf("image_encoder.net.4.weight"), gf("image_encoder.net.4.bias"), st), "launch_bn_backward:image_encoder.net.4.weight#27");
      MVAE_STEP(gemm_dgrad(dt, B, 400, 200, W.at<void>(P.dye2), wop("image_encoder.net.3.weight"), W.at<void>(P.dye1), dt,
                     W.at<void>(P.h1pre), sv_e1, sv_e1 + 400, pf("image_encoder.net.1.weight"),
                     pf("image_encoder.net.1.bias"), sb_e1, sb_e1 + 400, 1 << 30, st), "gemm_dgrad:image_encoder.net.3.weight#28");
      if (dep(st, s2)) return 1;
      MVAE_STEP(gemm_wgrad(dt, B, 200, 400, W.at<void>(P.dye2), W.at<void>(P.h1), gf("image_encoder.net.3.weight"), s2), "gemm_wgrad:image_encoder.net.3.weight#29");
      MVAE_STEP(launch_bn_backward(dt, W.at<void>(P.dye1), W.at<void>(P.h1pre), W.at<void>(P.dye1), B, 400, B, sb_e1,
                             sb_e1 + 400, sv_e1, sv_e1 + 400, pf("image_encoder.net.1.weight"),
                             gf("image_encoder.net.1.weight"), gf("image_encoder.net.1.bias"), st), "launch_bn_backward:image_encoder.net.1.weight#30");
      if (dep(st, s2)) return 1;
      MVAE_STEP(gemm_wgrad(dt, B, 400, 784, W.at<void>(P.dye1), a->image, gf("image_encoder.net.0.weight"), s2), "gemm_wgrad:image_encoder.net.0.weight#31");
    }
  }

  if (dep(s2, st)) return 1;  // join everything before the loss read-out / optimizer
  if (dep(s3, st)) return 1;
  // ---- losses out: [G][4] = total, bce, ce, kl
  if (a->out_losses != nullptr && fwd && !losses_packed)
    MVAE_STEP(launch_loss_pack(losses, a->out_losses, G, st), "launch_loss_pack#32");

  if (bwd && bwd_enc && a->do_adam) {
    MVAE_REQUIRE(a->adam_m && a->adam_v && a->adam_step, "mnist_step: Adam state missing");
    MVAE_STEP(launch_adam(prm, a->grads, a->adam_m, a->adam_v, a->params_bf16, adam_tail_floats > 0 ? adam_tail_floats : (adam_split ? L.enc_floats : L.param_floats), a->lr,
                    a->beta1, a->beta2, a->adam_eps, a->adam_step, a->grad_scale, 0, st), "launch_adam#33");
  }
  return 0;
}

long long mvae_launch_count(void) { return g_launches + mvae::noted_launches(); }

// Byte offset of a named workspace buffer (tests / bring-up: lets the host inspect intermediate tensors).
long long mvae_mnist_workspace_offset(const char* name, int batch, int n_latents, int dtype) {
  const Plan p = make_plan(batch, n_latents, dtype);
#define MVAE_OFF(f) if (strcmp(name, #f) == 0) return p.f;
  MVAE_OFF(bars) MVAE_OFF(st_e1) MVAE_OFF(st_e2) MVAE_OFF(st_d1) MVAE_OFF(st_d2) MVAE_OFF(st_t1)
  MVAE_OFF(sb_e1) MVAE_OFF(sb_e2) MVAE_OFF(sb_d1) MVAE_OFF(sb_d2) MVAE_OFF(sb_t1)
  MVAE_OFF(losses) MVAE_OFF(d_txt_table) MVAE_OFF(sv_e1) MVAE_OFF(sv_e2) MVAE_OFF(sv_d1) MVAE_OFF(sv_d2)
  MVAE_OFF(txt_table) MVAE_OFF(h1pre) MVAE_OFF(h1) MVAE_OFF(h2pre) MVAE_OFF(h2) MVAE_OFF(enc) MVAE_OFF(z)
  MVAE_OFF(t1pre) MVAE_OFF(g1pre) MVAE_OFF(g1) MVAE_OFF(g2pre) MVAE_OFF(g2) MVAE_OFF(dlog) MVAE_OFF(dyt)
  MVAE_OFF(dy2) MVAE_OFF(dy1) MVAE_OFF(dz) MVAE_OFF(denc) MVAE_OFF(dye2) MVAE_OFF(dye1)
#undef MVAE_OFF
  return -1;
}

int mvae_mnist_step_profile(const mvae_mnist_step_args* a, void* stream_v, int max_entries, char* labels, int label_stride,
                            float* ms_out, int* n_out) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  MVAE_REQUIRE(labels && ms_out && n_out && label_stride >= 16, "step_profile: bad output buffers");
  static bool created = false;
  if (!created) {
    for (int i = 0; i < 2 * 96; ++i) MVAE_CUDA(cudaEventCreate(&g_prof.ev[i]));
    created = true;
  }
  g_prof.on = true;
  g_prof.n = 0;
  const int rc = mvae_mnist_step(a, stream_v);
  g_prof.on = false;
  if (rc) return rc;
  MVAE_CUDA(cudaStreamSynchronize(st));
  int n = g_prof.n < max_entries ? g_prof.n : max_entries;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    MVAE_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
    ms_out[i] = ms;
    strncpy(labels + static_cast<size_t>(i) * label_stride, g_prof.label[i], label_stride - 1);
    labels[static_cast<size_t>(i) * label_stride + label_stride - 1] = 0;
  }
  *n_out = n;
  return 0;
}

int mvae_adam_step(float* params, float* grads, float* m, float* v, void* params_bf16, int64_t count, float lr,
                   float beta1, float beta2, float eps, const int* step_counter, float grad_scale, int zero_grad,
                   void* stream) {
  ++g_launches;
  return launch_adam(params, grads, m, v, params_bf16, count, lr, beta1, beta2, eps, step_counter, grad_scale, zero_grad,
                     static_cast<cudaStream_t>(stream));
}

int mvae_cast_f32_to_bf16(const float* in, void* out, int64_t count, void* stream) {
  return launch_cast_f32_bf16(in, out, count, static_cast<cudaStream_t>(stream));
}

int mvae_u8_to_act(const uint8_t* in, float* out_f32, void* out_bf16, int64_t count, float scale, void* stream) {
  return launch_u8_to_act(in, out_f32, out_bf16, count, scale, static_cast<cudaStream_t>(stream));
}

int mvae_poe_forward(int mode, int prior_expert, float eps, int n_experts, int64_t batch, int dim, const float* mu,
                     const float* logvar, const float* mask, float* out_mu, float* out_logvar, void* stream) {
  return launch_poe_forward(mode, prior_expert, eps, n_experts, batch, dim, mu, logvar, mask, out_mu, out_logvar,
                            static_cast<cudaStream_t>(stream));
}

int mvae_poe_backward(int mode, int prior_expert, float eps, int n_experts, int64_t batch, int dim, const float* mu,
                      const float* logvar, const float* mask, const float* d_out_mu, const float* d_out_logvar,
                      float* d_mu, float* d_logvar, void* stream) {
  return launch_poe_backward(mode, prior_expert, eps, n_experts, batch, dim, mu, logvar, mask, d_out_mu, d_out_logvar,
                             d_mu, d_logvar, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
