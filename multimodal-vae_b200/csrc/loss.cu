// Standalone ELBO loss on the module outputs (probabilities / log-probabilities), forward and backward:
// the reference's loss_function (mnist/train.py:64-81) for callers that use MVAE.forward() + loss_function()
// instead of the fused training step.  Memory-bound streaming reductions: 16-byte loads, warp shuffle + block
// reduction, one atomic per block per term.
#include <algorithm>

#include "../../include/mvae_b200.h"
#include "common.cuh"

namespace mvae {
namespace {

constexpr int kLossThreads = 256;

__device__ __forceinline__ float to_float(float x) { return x; }
__device__ __forceinline__ float to_float(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ void from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < kLossThreads / 32 ? s_red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

// out[1] += lambda_image * mean BCE; out[2] += lambda_text * mean NLL; out[3] += kl_weight * KL; out[0] = their sum
template <typename T>
__global__ void __launch_bounds__(kLossThreads)
    elbo_fwd_kernel(const T* __restrict__ prob, const T* __restrict__ target, long long n_pix, const float* __restrict__ logp,
                    const long long* __restrict__ labels, long long B, int C, const float* __restrict__ mu,
                    const float* __restrict__ logvar, long long n_lat, float w_img, float w_txt, float w_kl,
                    float* __restrict__ out) {
  __shared__ float s_red[kLossThreads / 32];
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  float bce = 0.f, nll = 0.f, kl = 0.f;
  if (prob != nullptr)
    for (long long i = tid; i < n_pix; i += nth) {
      const float p = to_float(prob[i]), t = to_float(target[i]);
      // F.binary_cross_entropy clamps both logs at -100 (mnist/train.py:70)
      bce -= t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(logf(1.f - p), -100.f);
    }
  if (logp != nullptr)
    for (long long i = tid; i < B; i += nth) nll -= logp[i * C + labels[i]];  // F.nll_loss (mnist/train.py:73)
  for (long long i = tid; i < n_lat; i += nth) {
    const float m = mu[i], l = logvar[i];
    kl += 1.f + l - m * m - expf(l);  // mnist/train.py:79
  }
  const float sb = block_sum(bce, s_red);
  const float sn = block_sum(nll, s_red);
  const float sk = block_sum(kl, s_red);
  if (threadIdx.x == 0) {
    const float a = w_img * sb, b = w_txt * sn, c = -0.5f * w_kl * sk;
    if (a != 0.f) atomicAdd(out + 1, a);
    if (b != 0.f) atomicAdd(out + 2, b);
    if (c != 0.f) atomicAdd(out + 3, c);
    atomicAdd(out + 0, a + b + c);
  }
}

template <typename T>
__global__ void __launch_bounds__(kLossThreads)
    elbo_bwd_kernel(const T* __restrict__ prob, const T* __restrict__ target, long long n_pix, const float* __restrict__ logp,
                    const long long* __restrict__ labels, long long B, int C, const float* __restrict__ mu,
                    const float* __restrict__ logvar, long long n_lat, float w_img, float w_txt, float w_kl,
                    const float* __restrict__ grad_out, T* __restrict__ d_prob, float* __restrict__ d_logp,
                    float* __restrict__ d_mu, float* __restrict__ d_logvar) {
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  const float go = grad_out != nullptr ? *grad_out : 1.f;
  if (d_prob != nullptr)
    for (long long i = tid; i < n_pix; i += nth) {
      const float p = to_float(prob[i]), t = to_float(target[i]);
      // ATen's binary_cross_entropy_backward: (p - t) / max((1 - p) * p, 1e-12)
      from_float(d_prob + i, go * w_img * (p - t) / fmaxf((1.f - p) * p, 1e-12f));
    }
  if (d_logp != nullptr)
    for (long long i = tid; i < B * C; i += nth) {
      const long long b = i / C;
      d_logp[i] = (i - b * C) == labels[b] ? -go * w_txt : 0.f;
    }
  for (long long i = tid; i < n_lat; i += nth) {
    const float m = mu[i], l = logvar[i];
    if (d_mu != nullptr) d_mu[i] = go * w_kl * m;
    if (d_logvar != nullptr) d_logvar[i] = go * 0.5f * w_kl * (expf(l) - 1.f);
  }
}

int blocks_for(long long n) {
  long long b = (n + kLossThreads - 1) / kLossThreads;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(b, 148 * 8)));
}

}  // namespace
}  // namespace mvae

using namespace mvae;

extern "C" {

int mvae_elbo_loss_forward(const mvae_elbo_loss_args* a, float* out4, void* stream) {
  MVAE_REQUIRE(a != nullptr && out4 != nullptr, "elbo_loss_forward: null arguments");
  MVAE_REQUIRE(a->mu != nullptr && a->logvar != nullptr && a->batch > 0 && a->n_latents > 0, "elbo_loss: mu/logvar missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MVAE_CUDA(cudaMemsetAsync(out4, 0, 4 * sizeof(float), st));
  const long long n_pix = a->recon_image != nullptr ? a->batch * static_cast<long long>(a->n_pixels) : 0;
  const float w_img = a->recon_image != nullptr ? a->lambda_image / static_cast<float>(n_pix) : 0.f;
  const float w_txt = a->recon_text != nullptr ? a->lambda_text / static_cast<float>(a->batch) : 0.f;
  const int blocks = blocks_for(std::max<long long>(n_pix, a->batch * static_cast<long long>(a->n_latents)));
  if (a->image_dtype == MVAE_DT_F32)
    elbo_fwd_kernel<float><<<blocks, kLossThreads, 0, st>>>(
        static_cast<const float*>(a->recon_image), static_cast<const float*>(a->image), n_pix, a->recon_text,
        reinterpret_cast<const long long*>(a->text), a->batch, a->n_classes, a->mu, a->logvar,
        a->batch * static_cast<long long>(a->n_latents), w_img, w_txt, a->kl_weight, out4);
  else
    elbo_fwd_kernel<__nv_bfloat16><<<blocks, kLossThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(a->recon_image), static_cast<const __nv_bfloat16*>(a->image), n_pix,
        a->recon_text, reinterpret_cast<const long long*>(a->text), a->batch, a->n_classes, a->mu, a->logvar,
        a->batch * static_cast<long long>(a->n_latents), w_img, w_txt, a->kl_weight, out4);
  MVAE_CUDA(cudaGetLastError());
  return 0;
}

int mvae_elbo_loss_backward(const mvae_elbo_loss_args* a, const float* grad_out, void* d_recon_image, float* d_recon_text,
                            float* d_mu, float* d_logvar, void* stream) {
  MVAE_REQUIRE(a != nullptr, "elbo_loss_backward: null arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n_pix = a->recon_image != nullptr ? a->batch * static_cast<long long>(a->n_pixels) : 0;
  const float w_img = a->recon_image != nullptr ? a->lambda_image / static_cast<float>(n_pix) : 0.f;
  const float w_txt = a->recon_text != nullptr ? a->lambda_text / static_cast<float>(a->batch) : 0.f;
  const int blocks = blocks_for(std::max<long long>(n_pix, a->batch * static_cast<long long>(a->n_latents)));
  if (a->image_dtype == MVAE_DT_F32)
    elbo_bwd_kernel<float><<<blocks, kLossThreads, 0, st>>>(
        static_cast<const float*>(a->recon_image), static_cast<const float*>(a->image), n_pix, a->recon_text,
        reinterpret_cast<const long long*>(a->text), a->batch, a->n_classes, a->mu, a->logvar,
        a->batch * static_cast<long long>(a->n_latents), w_img, w_txt, a->kl_weight, grad_out,
        static_cast<float*>(d_recon_image), d_recon_text, d_mu, d_logvar);
  else
    elbo_bwd_kernel<__nv_bfloat16><<<blocks, kLossThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(a->recon_image), static_cast<const __nv_bfloat16*>(a->image), n_pix,
        a->recon_text, reinterpret_cast<const long long*>(a->text), a->batch, a->n_classes, a->mu, a->logvar,
        a->batch * static_cast<long long>(a->n_latents), w_img, w_txt, a->kl_weight, grad_out,
        static_cast<__nv_bfloat16*>(d_recon_image), d_recon_text, d_mu, d_logvar);
  MVAE_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
