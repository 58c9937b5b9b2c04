// Shared host/device declarations for the MVAE B200 library (internal; the public C ABI is
// include/mvae_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mvae {

// Error plumbing: every C-ABI entry returns 0 on success; the message of the last failure on
// the calling thread is kept for mvae_last_error().
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define MVAE_CUDA(expr)                                                  \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) return ::mvae::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define MVAE_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      ::mvae::set_error(__VA_ARGS__); \
      return 1;                      \
    }                                \
  } while (0)

// Storage type of activations / gradients: 0 = fp32 (TF32 tensor path), 1 = bf16.
enum : int { MVAE_F32 = 0, MVAE_BF16 = 1 };

// ---------------------------------------------------------------- GEMM (tcgen05)
enum : int { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_BCE = 2, EPI_DGRAD_BN = 3, EPI_STORE_ACT = 4, EPI_DGRAD_ACT = 5 };

struct GemmEpilogue {
  int kind = EPI_STORE;
  void* C = nullptr;        // [M, N] row-major, leading dimension ldc (elements)
  long long ldc = 0;
  int c_dtype = MVAE_F32;   // EPI_ATOMIC requires fp32
  const float* bias = nullptr;  // [N]
  // Column statistics, accumulated with atomics into [groups][N] fp32 buffers; group = row / rows_per_group.
  //   EPI_STORE    : stat0 += sum(y), stat1 += sum(y*y)            (BatchNorm batch statistics)
  //   EPI_BCE      : stat0 += sum(dlogit)  (stat0 is a single [N] bias gradient, no groups)
  //   EPI_DGRAD_BN : stat0 += sum(dyhat), stat1 += sum(dyhat * xhat)
  float* stat0 = nullptr;
  float* stat1 = nullptr;
  int rows_per_group = 1 << 30;
  // EPI_BCE: logits never leave the SM. target is [target_rows, N] (activation dtype), row m reads
  // target[m % target_rows]; group g = m / rows_per_group has weight bce_scale[g] (= lambda / (B*N));
  // loss[g] += sum softplus(x) - t*x  (unscaled; the caller scales), dlogit = bce_scale[g]*(sigmoid(x)-t).
  const void* target = nullptr;
  long long ldt = 0;
  int target_rows = 1;
  float bce_scale[4] = {0, 0, 0, 0};
  float* loss = nullptr;   // [groups]
  void* probs = nullptr;   // optional sigmoid(x) output, same layout/dtype as C
  const float* row_w = nullptr;  // EPI_BCE: optional [M] per-row weight of loss and gradient (per-sample masks)
  // EPI_STORE_ACT (Linear + bias + Swish in one pass): C = pre = acc + bias (kept for the backward, may be null),
  //   probs = pre * sigmoid(pre) - the next layer's operand (same layout/dtype as C)
  // EPI_DGRAD_ACT (input gradient through a Swish): C = acc * swish'(hpre), stat0[col] += sum_rows C (= the bias
  //   gradient of the Linear that produced hpre)
  // EPI_DGRAD_BN: C = acc * 1[gamma*xhat+beta > 0], xhat = (hpre - mean[g])*rstd[g]
  const void* hpre = nullptr;
  long long ldh = 0;
  const float* bn_mean = nullptr;  // [groups][N]
  const float* bn_rstd = nullptr;  // [groups][N]
  const float* bn_gamma = nullptr; // [N]
  const float* bn_beta = nullptr;  // [N]
};

// Implicit patch-matrix operand (implicit GEMM for the convolutions): instead of reading a materialised im2col
// matrix col[m, (kh*k+kw)*C + c] = X[n, ho*s-p+kh, wo*s-p+kw, c] through TMA, warps 4..7 of the GEMM gather its
// 16-byte chunks (8 bf16 channels of one tap) straight from the NHWC activation into the swizzled operand stage.
//   mode 1: the patch matrix is A [M = n*Ho*Wo, K = k*k*C], K-major  (conv forward, transposed-conv dgrad)
//   mode 2: the patch matrix is B, MN-major: reduction index = pixel, N = k*k*C  (weight gradients)
struct ConvGather {
  int mode = 0;
  const void* X = nullptr;
  int H = 0, W = 0, C = 0, ksize = 0, stride = 1, pad = 0, Ho = 0, Wo = 0;
  long long sn = 0, sh = 0, sw = 0;  // element strides of X (channel stride 1)
  long long extent = 0;              // elements spanned by X (32-bit offsets inside the kernel)
  unsigned long long magic_hw = 0, magic_w = 0;  // filled by launch_gemm: multiply-shift division by Ho*Wo and Wo
  // mode 3 (one output-parity class of a transposed convolution, see _ops.transposed_conv_classes): the patch matrix is A
  // with a ksize x ksize_w tap window, pads (pad, pad_w), stride 1 over the Ho x Wo class grid; B is the weight tensor
  // [C, kk*kk, N] read tap by tap through a rank-3 tensor map (tap of window position (th, tw) = kh_tab[th]*kk + kw_tab[tw]);
  // row (n, u, v) of the result goes to output pixel (n, sc_stride*u + sc_a, sc_stride*v + sc_b) of an sc_hout x sc_wout image.
  int ksize_w = 0, pad_w = 0, kk = 0;
  signed char kh_tab[8] = {0, 0, 0, 0, 0, 0, 0, 0}, kw_tab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int sc_hout = 0, sc_wout = 0, sc_stride = 1, sc_a = 0, sc_b = 0;
  // mode 4: ALL sc_stride^2 parity classes of mode 3 in one launch (blockIdx.z = class (a, b) = (z / stride, z % stride));
  // needs classes of equal shape (same grid, same tap counts), so only the per-axis pad and tap tables vary with the class
  signed char ax_pad[4] = {0, 0, 0, 0};
  signed char ax_k[4][8] = {{0}};
};

struct GemmDesc {
  int kind = MVAE_F32;  // operand storage: fp32 (kind::tf32) or bf16 (kind::f16)
  int M = 0, N = 0, K = 0;
  // A is logically [M, K], B is logically [N, K]; C = A * B^T.
  // major = 0: contraction index contiguous (row-major [rows, K]);
  // major = 1: row index contiguous (stored as [K, rows] row-major).
  const void* A = nullptr;
  long long lda = 0;
  int a_mn = 0;
  const void* B = nullptr;
  long long ldb = 0;
  int b_mn = 0;
  int block_n = 0;  // 0 = auto
  int split_k = 0;  // 0 = auto (only EPI_ATOMIC may split)
  int stages = 0;   // 0 = auto
  long long* dbg = nullptr;  // device buffer [ctas][8] of %globaltimer stamps (bring-up only)
  // Error-compensated 3xTF32 (fp32 storage only): each operand is split into hi = tf32(x) and lo = x - hi and the
  // product is evaluated as hi*hi + lo*hi + hi*lo by ONE GEMM over a 3x longer contraction ([hi|lo|hi] x [hi|hi|lo]).
  // The split copies go to caller-provided scratch: x3_a >= 3*M*round_up(K,4) floats, x3_b >= 3*N*round_up(K,4).
  int x3 = 0;
  void* x3_a = nullptr;
  void* x3_b = nullptr;
  GemmEpilogue epi;
  ConvGather gather;
};

int launch_gemm(const GemmDesc& g, cudaStream_t stream);
void set_gemm_debug_times(void* ptr, int epi_kind);

// Tuning knobs readable from the environment (debug / bench sweeps only).
int env_int(const char* name, int dflt);

// Kernel launches enqueued by the operator-level entries (added to mvae_launch_count()).
void note_launch(int n = 1);
long long noted_launches();

// Set by the step orchestration right before the launch of a kernel that begins with ptx::griddep_wait(): that ONE launch
// carries the programmatic-stream-serialization attribute (its CTAs may be scheduled while the previous kernel of the stream
// drains) and clears the flag.  Never set it for a kernel without the wait: it would run beside its predecessor.
extern thread_local int g_pdl_next;

// One launch helper for every kernel of the library (error-checked cudaLaunchKernelEx).
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
int launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute pdl_attr[1];
  if (g_pdl_next) {
    g_pdl_next = 0;
    pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = pdl_attr;
    cfg.numAttrs = 1;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx", __FILE__, __LINE__);
  return 0;
}
#endif

}  // namespace mvae
