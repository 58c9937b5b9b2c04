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

#define MVAE_ABI_VERSION 1

#define MVAE_DT_F32 0
#define MVAE_DT_BF16 1

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
} mvae_gemm_args;
int mvae_gemm(const mvae_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVAE_B200_H_ */
