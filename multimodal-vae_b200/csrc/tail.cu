// The memory-bound "tail" of the MVAE step, one fused kernel per direction:
//
//   forward : encoder outputs -> ProductOfExperts (all ELBO terms at once: joint / image-only /
//             text-only) -> reparametrize -> z for the decoders, KL partials, and the text
//             decoder's first Linear (its input z is already in registers).
//   backward: dz (from the image decoder dgrad) + the text decoder's BatchNorm/Linear backward ->
//             reparametrize / KL / PoE backward -> gradient at the encoder outputs, summed over terms.
//
// One warp owns one sample (row) for ALL terms, lanes own latent pairs (float2, coalesced 256 B per
// row for n = 64), warp shuffles reduce across the latent axis, shared-memory atomics reduce across
// the rows of a block, one global atomic per address per block finishes.
//
// Reference: ProductOfExperts mnist/model.py:173-185, reparametrize :24-30, forward :53-84,
// KL term of loss_function mnist/train.py:79-80, TextDecoder first Linear mnist/model.py:162.
// Also here: the small text-side kernels (TextEncoder mnist/model.py:138-153 evaluated per label,
// TextDecoder BN/ReLU/Linear/log_softmax :163-170 with the NLL of mnist/train.py:73) and the
// standalone masked ProductOfExperts op.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "poe.cuh"
#include "ptx.cuh"
#include "../../include/mvae_b200.h"

namespace mvae {

namespace {

constexpr int kTailThreads = 256;
constexpr int kTailWarps = kTailThreads / 32;
constexpr int kTD = 10;  // text decoder width / number of classes

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ void store_pair(T* p, float a, float b) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float2*>(p) = make_float2(a, b);
  } else {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
}

// ================================================================= tail forward
// One warp per (term, sample): lanes own latent pairs.
template <typename ZT>
__global__ void __launch_bounds__(kTailThreads, 3) tail_fwd_kernel(const TailArgs a) {
  ptx::griddep_launch();   // programmatic launch (common.cuh, g_pdl_next): the next kernel's prologue may overlap this one
  ptx::griddep_wait();     // ... and this one starts while its predecessor drains; no global access before this line
  __shared__ float s_stat[kMaxGroups][2][kTD];
  __shared__ float s_kl[kMaxGroups];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kMaxGroups * 2 * kTD; i += blockDim.x) (&s_stat[0][0][0])[i] = 0.f;
  if (threadIdx.x < kMaxGroups) s_kl[threadIdx.x] = 0.f;
  __syncthreads();

  const int n = a.n, two_n = 2 * a.n;
  const uint32_t step = a.step_ptr != nullptr ? static_cast<uint32_t>(*a.step_ptr) : 0u;
  const long long items = static_cast<long long>(a.G) * a.B;

  for (long long it = blockIdx.x * kTailWarps + warp; it < items; it += static_cast<long long>(gridDim.x) * kTailWarps) {
    const int g = static_cast<int>(it / a.B);
    const int b = static_cast<int>(it - static_cast<long long>(g) * a.B);
    const int ty = a.group_type[g];
    const bool present[2] = {ty != TERM_TEXT && a.enc_img != nullptr, ty != TERM_IMAGE && a.txt_table != nullptr};
    const int label = present[1] ? static_cast<int>(a.labels[b]) : 0;
    float t1[kTD];
#pragma unroll
    for (int j = 0; j < kTD; ++j) t1[j] = 0.f;
    float klacc = 0.f;
    for (int k = lane * 2; k < n; k += 64) {
      float mi[2] = {0.f, 0.f}, li[2] = {0.f, 0.f}, mt[2] = {0.f, 0.f}, lt[2] = {0.f, 0.f};
      if (present[0]) {
        const float2 m2 = *reinterpret_cast<const float2*>(a.enc_img + static_cast<long long>(b) * two_n + k);
        const float2 l2 = *reinterpret_cast<const float2*>(a.enc_img + static_cast<long long>(b) * two_n + n + k);
        mi[0] = m2.x; mi[1] = m2.y; li[0] = l2.x; li[1] = l2.y;
      }
      if (present[1]) {
        const float2 m2 = *reinterpret_cast<const float2*>(a.txt_table + label * two_n + k);
        const float2 l2 = *reinterpret_cast<const float2*>(a.txt_table + label * two_n + n + k);
        mt[0] = m2.x; mt[1] = m2.y; lt[0] = l2.x; lt[1] = l2.y;
      }
      float2 e2 = make_float2(0.f, 0.f);
      if (a.training) {
        if (a.eps != nullptr)
          e2 = *reinterpret_cast<const float2*>(a.eps + it * n + k);
        else
          e2 = normal_pair(a.seed, step, (it * n + k) >> 1);
      }
      const float ee[2] = {e2.x, e2.y};
      float zz[2] = {0.f, 0.f}, mm[2] = {0.f, 0.f}, ll[2] = {0.f, 0.f};
      if (a.z_in != nullptr) {  // decode_image / decode_text: the caller's latents, no experts
        const float2 zi = *reinterpret_cast<const float2*>(a.z_in + it * n + k);
        zz[0] = zi.x;
        zz[1] = zi.y;
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float m[2] = {mi[c], mt[c]};
          const float lv[2] = {li[c], lt[c]};
          const Poe r = poe_eval<true>(a.poe_mode, a.prior_expert, a.poe_eps, m, lv, present);
          mm[c] = r.mu;
          ll[c] = r.logvar;
          // reparametrize (mnist/model.py:25-28): std = exp(0.5*logvar) = sqrt(pd_var); z = eps*std + mu
          zz[c] = a.training ? ee[c] * sqrtf(r.pd_var) + r.mu : r.mu;
          // KL integrand of mnist/train.py:79 (exp(logvar) = pd_var)
          klacc += 1.f + r.logvar - r.mu * r.mu - r.pd_var;
        }
      }
      store_pair(reinterpret_cast<ZT*>(a.z) + it * n + k, zz[0], zz[1]);
      if (a.mu != nullptr) {
        *reinterpret_cast<float2*>(a.mu + it * n + k) = make_float2(mm[0], mm[1]);
        *reinterpret_cast<float2*>(a.logvar + it * n + k) = make_float2(ll[0], ll[1]);
      }
      if (a.wt1 != nullptr) {
#pragma unroll
        for (int j = 0; j < kTD; ++j) {
          const float2 w = *reinterpret_cast<const float2*>(a.wt1 + j * n + k);
          t1[j] = fmaf(zz[0], w.x, fmaf(zz[1], w.y, t1[j]));
        }
      }
    }
    if (a.wt1 != nullptr) {
      float mine = 0.f;
#pragma unroll
      for (int j = 0; j < kTD; ++j) {
        const float sum = warp_sum(t1[j]);
        if (lane == j) mine = sum;
      }
      if (lane < kTD) {
        const float v = mine + a.bt1[lane];
        a.t1pre[it * kTD + lane] = v;
        atomicAdd(&s_stat[g][0][lane], v);
        atomicAdd(&s_stat[g][1][lane], v * v);
      }
    }
    const float ks = warp_sum(klacc);
    if (lane == 0) atomicAdd(&s_kl[g], ks);
  }
  __syncthreads();
  if (a.wt1 != nullptr)
    for (int i = threadIdx.x; i < a.G * 2 * kTD; i += blockDim.x) {
      const int g = i / (2 * kTD), w = (i / kTD) % 2, j = i % kTD;
      if (s_stat[g][w][j] != 0.f) atomicAdd((w == 0 ? a.t1_sum : a.t1_sumsq) + g * kTD + j, s_stat[g][w][j]);
    }
  if (threadIdx.x < a.G && a.kl != nullptr && s_kl[threadIdx.x] != 0.f)
    atomicAdd(a.kl + threadIdx.x, -0.5f * a.kl_weight[threadIdx.x] * s_kl[threadIdx.x]);
}

// ================================================================= tail forward, one warp per SAMPLE (n <= 64)
// The warp carries its sample through all ELBO terms: both experts are loaded and exponentiated once, the text decoder's
// first Linear keeps its weights in registers and all (term, feature) dot products are reduced in ONE transposing butterfly
// (31 shuffles for up to 32 values instead of five per value), the KL partials stay in registers until the warp is done.
// Same arithmetic per element as tail_fwd_kernel (poe_combine == poe_eval bit for bit, same Philox stream).
template <typename ZT, int kFwd2Warps>
__global__ void __launch_bounds__(32 * kFwd2Warps, 28 / kFwd2Warps) tail_fwd2_kernel(const TailArgs a) {
  ptx::griddep_launch();   // programmatic launch (common.cuh, g_pdl_next): the next kernel's prologue may overlap this one
  ptx::griddep_wait();     // ... and this one starts while its predecessor drains; no global access before this line
  __shared__ float s_stat[kMaxGroups][2][kTD];
  __shared__ float s_kl[kMaxGroups];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kMaxGroups * 2 * kTD; i += blockDim.x) (&s_stat[0][0][0])[i] = 0.f;
  if (threadIdx.x < kMaxGroups) s_kl[threadIdx.x] = 0.f;
  __syncthreads();

  const int n = a.n, two_n = 2 * a.n, G = a.G;
  const uint32_t step = a.step_ptr != nullptr ? static_cast<uint32_t>(*a.step_ptr) : 0u;
  const int k = lane * 2;
  const bool act = k < n;
  bool any_img = false, any_txt = false;
  for (int g = 0; g < G; ++g) {
    any_img |= a.group_type[g] != TERM_TEXT;
    any_txt |= a.group_type[g] != TERM_IMAGE;
  }
  any_img = any_img && a.enc_img != nullptr && a.z_in == nullptr;
  any_txt = any_txt && a.txt_table != nullptr && a.z_in == nullptr;
  const bool text_dec = a.wt1 != nullptr;
  float2 w[kTD];
#pragma unroll
  for (int j = 0; j < kTD; ++j) w[j] = (text_dec && act) ? *reinterpret_cast<const float2*>(a.wt1 + j * n + k) : make_float2(0.f, 0.f);
  float klacc[kMaxGroups] = {0.f, 0.f, 0.f};

  for (int b = blockIdx.x * kFwd2Warps + warp; b < a.B; b += gridDim.x * kFwd2Warps) {
    float mi[2] = {0.f, 0.f}, mt[2] = {0.f, 0.f};
    PoePart pi[2] = {{1.f, 1.f}, {1.f, 1.f}}, pt[2] = {{1.f, 1.f}, {1.f, 1.f}};
    if (act && any_img) {
      const float2 m2 = *reinterpret_cast<const float2*>(a.enc_img + static_cast<long long>(b) * two_n + k);
      const float2 l2 = *reinterpret_cast<const float2*>(a.enc_img + static_cast<long long>(b) * two_n + n + k);
      mi[0] = m2.x; mi[1] = m2.y;
      pi[0] = poe_part(l2.x, a.poe_eps); pi[1] = poe_part(l2.y, a.poe_eps);
    }
    if (act && any_txt) {
      const int label = static_cast<int>(a.labels[b]);
      const float2 m2 = *reinterpret_cast<const float2*>(a.txt_table + label * two_n + k);
      const float2 l2 = *reinterpret_cast<const float2*>(a.txt_table + label * two_n + n + k);
      mt[0] = m2.x; mt[1] = m2.y;
      pt[0] = poe_part(l2.x, a.poe_eps); pt[1] = poe_part(l2.y, a.poe_eps);
    }
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      if (g >= G) break;
      const int ty = a.group_type[g];
      const bool present[2] = {ty != TERM_TEXT && any_img, ty != TERM_IMAGE && any_txt};
      const long long it = static_cast<long long>(g) * a.B + b;
      float zz[2] = {0.f, 0.f}, mm[2] = {0.f, 0.f}, ll[2] = {0.f, 0.f};
      if (act) {
        if (a.z_in != nullptr) {  // decode_image / decode_text: the caller's latents, no experts
          const float2 zi = *reinterpret_cast<const float2*>(a.z_in + it * n + k);
          zz[0] = zi.x;
          zz[1] = zi.y;
        } else {
          float2 e2 = make_float2(0.f, 0.f);
          if (a.training) {
            if (a.eps != nullptr) {
              e2 = *reinterpret_cast<const float2*>(a.eps + it * n + k);
            } else {
              e2 = normal_pair(a.seed, step, (it * n + k) >> 1);
              if (a.noise_buf != nullptr) *reinterpret_cast<float2*>(a.noise_buf + it * n + k) = e2;
            }
          }
          const float ee[2] = {e2.x, e2.y};
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float m[2] = {mi[c], mt[c]};
            const PoePart parts[2] = {pi[c], pt[c]};
            const Poe r = poe_combine<true>(a.poe_mode, a.prior_expert, m, parts, present);
            mm[c] = r.mu;
            ll[c] = r.logvar;
            zz[c] = a.training ? ee[c] * sqrtf(r.pd_var) + r.mu : r.mu;   // reparametrize (mnist/model.py:25-28)
            klacc[g] += 1.f + r.logvar - r.mu * r.mu - r.pd_var;          // KL integrand of mnist/train.py:79
          }
        }
        store_pair(reinterpret_cast<ZT*>(a.z) + it * n + k, zz[0], zz[1]);
        if (a.mu != nullptr) {
          *reinterpret_cast<float2*>(a.mu + it * n + k) = make_float2(mm[0], mm[1]);
          *reinterpret_cast<float2*>(a.logvar + it * n + k) = make_float2(ll[0], ll[1]);
        }
      }
      if (text_dec) {
#pragma unroll
        for (int j = 0; j < kTD; ++j) v[g * kTD + j] = fmaf(zz[0], w[j].x, zz[1] * w[j].y);
      }
    }
    if (text_dec) {
      // transposing butterfly: afterwards lane L holds the warp-wide sum of value L (= term L / 10, feature L % 10)
#pragma unroll
      for (int h = 16; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
          const float send = up ? v[i] : v[i + h];
          const float keep = up ? v[i + h] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
      }
      if (lane < G * kTD) {
        const int g = lane / kTD, j = lane - g * kTD;
        const float t = v[0] + a.bt1[j];
        a.t1pre[(static_cast<long long>(g) * a.B + b) * kTD + j] = t;
        atomicAdd(&s_stat[g][0][j], t);
        atomicAdd(&s_stat[g][1][j], t * t);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    const float ks = warp_sum(klacc[g]);
    if (lane == 0 && g < G) atomicAdd(&s_kl[g], ks);
  }
  __syncthreads();
  if (text_dec)
    for (int i = threadIdx.x; i < G * 2 * kTD; i += blockDim.x) {
      const int g = i / (2 * kTD), ww = (i / kTD) % 2, j = i % kTD;
      if (s_stat[g][ww][j] != 0.f) atomicAdd((ww == 0 ? a.t1_sum : a.t1_sumsq) + g * kTD + j, s_stat[g][ww][j]);
    }
  if (threadIdx.x < G && a.kl != nullptr && s_kl[threadIdx.x] != 0.f)
    atomicAdd(a.kl + threadIdx.x, -0.5f * a.kl_weight[threadIdx.x] * s_kl[threadIdx.x]);
}

// ================================================================= tail backward
// Block = kBwdRows samples x G terms, one warp per (sample, term); the terms' contributions to one sample's
// encoder-output gradient are combined through shared memory (no global atomics on activations).
constexpr int kBwdRows = 4;
template <typename ZT>
__global__ void __launch_bounds__(32 * kMaxGroups * kBwdRows, 2) tail_bwd_kernel(const TailArgs a, int smem_floats) {
  ptx::griddep_launch();   // programmatic launch (common.cuh, g_pdl_next): the next kernel's prologue may overlap this one
  ptx::griddep_wait();     // ... and this one starts while its predecessor drains; no global access before this line
  extern __shared__ float sm[];
  // layout: d_txt_table [10][2n] | d_wt1 [10][n] | d_enc_bias [2n] | t1 coefficients [G][4][10] |
  //         image-expert combine [rows][G][2n] | text-expert combine [rows][G][2n] | labels [rows]
  const int n = a.n, two_n = 2 * a.n;
  float* s_tab = sm;
  float* s_w1 = s_tab + kTD * two_n;
  float* s_eb = s_w1 + kTD * n;
  float* s_co = s_eb + two_n;  // per group: mean, rstd, c0 = S0/B, c1 = S1/B
  float* s_cb = s_co + kMaxGroups * 4 * kTD;
  float* s_ct = s_cb + kBwdRows * kMaxGroups * two_n;
  int* s_lab = reinterpret_cast<int*>(s_ct + kBwdRows * kMaxGroups * two_n);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = a.G;
  const int wrow = warp / G, g = warp - wrow * G;
  for (int i = threadIdx.x; i < smem_floats; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const bool text_dec = a.t1_dyhat != nullptr;
  if (text_dec && threadIdx.x < G * kTD) {
    const int gg = threadIdx.x / kTD, j = threadIdx.x % kTD;
    const float mean = a.t1_sum[gg * kTD + j] / a.B;
    const float var = fmaxf(a.t1_sumsq[gg * kTD + j] / a.B - mean * mean, 0.f);
    s_co[(gg * 4 + 0) * kTD + j] = mean;
    s_co[(gg * 4 + 1) * kTD + j] = rsqrtf(var + 1e-5f);
    s_co[(gg * 4 + 2) * kTD + j] = a.t1_s0[gg * kTD + j] / a.B;
    s_co[(gg * 4 + 3) * kTD + j] = a.t1_s1[gg * kTD + j] / a.B;
  }
  __syncthreads();
  const uint32_t step = a.step_ptr != nullptr ? static_cast<uint32_t>(*a.step_ptr) : 0u;
  const int ty = a.group_type[g];
  const bool present[2] = {ty != TERM_TEXT, ty != TERM_IMAGE};
  const float c_kl = a.kl_weight[g];

  // Per-lane register accumulators for the latent pair of the first 64-column pass (covers n <= 64 entirely);
  // later passes (n > 64) fall back to shared atomics.
  float r_w1[kTD][2], r_eb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < kTD; ++j) r_w1[j][0] = r_w1[j][1] = 0.f;

  const int row_blocks = (a.B + kBwdRows - 1) / kBwdRows;
  for (int rb = blockIdx.x; rb < row_blocks; rb += gridDim.x) {
    const int b = rb * kBwdRows + wrow;
    const bool active = b < a.B;
    const long long row = static_cast<long long>(g) * a.B + b;
    if (active) {
      const int label = present[1] ? static_cast<int>(a.labels[b]) : 0;
      // text decoder: gradient at its first Linear's output (BatchNorm backward apply): lane j computes feature j,
      // then the ten values are broadcast to every lane (full-warp shuffles, outside the latent loop)
      float dt_mine = 0.f;
      if (text_dec && lane < kTD) {
        const float x = a.t1pre[row * kTD + lane];
        const float xh = (x - s_co[(g * 4 + 0) * kTD + lane]) * s_co[(g * 4 + 1) * kTD + lane];
        const float dy = a.t1_dyhat[row * kTD + lane];
        dt_mine = a.t1_gamma[lane] * s_co[(g * 4 + 1) * kTD + lane] *
                  (dy - s_co[(g * 4 + 2) * kTD + lane] - xh * s_co[(g * 4 + 3) * kTD + lane]);
      }
      float dtb[kTD];
#pragma unroll
      for (int j = 0; j < kTD; ++j) dtb[j] = __shfl_sync(0xffffffffu, dt_mine, j);
      for (int k = lane * 2; k < n; k += 64) {
        const bool first_pass = k < 64;
        float mi[2] = {0.f, 0.f}, li[2] = {0.f, 0.f}, mt[2] = {0.f, 0.f}, lt[2] = {0.f, 0.f};
        if (present[0]) {
          const float2 m2 = *reinterpret_cast<const float2*>(a.enc_img + static_cast<long long>(b) * two_n + k);
          const float2 l2 = *reinterpret_cast<const float2*>(a.enc_img + static_cast<long long>(b) * two_n + n + k);
          mi[0] = m2.x; mi[1] = m2.y; li[0] = l2.x; li[1] = l2.y;
        }
        if (present[1]) {
          const float2 m2 = *reinterpret_cast<const float2*>(a.txt_table + label * two_n + k);
          const float2 l2 = *reinterpret_cast<const float2*>(a.txt_table + label * two_n + n + k);
          mt[0] = m2.x; mt[1] = m2.y; lt[0] = l2.x; lt[1] = l2.y;
        }
        float2 e2 = make_float2(0.f, 0.f);
        if (a.training) {
          if (a.eps != nullptr)
            e2 = *reinterpret_cast<const float2*>(a.eps + row * n + k);
          else
            e2 = normal_pair(a.seed, step, (row * n + k) >> 1);
        }
        const float ee[2] = {e2.x, e2.y};
        float2 dz2 = make_float2(0.f, 0.f);
        if (a.dz != nullptr) dz2 = *reinterpret_cast<const float2*>(a.dz + row * n + k);
        float dzz[2] = {dz2.x, dz2.y};
        Poe r[2];
        float zz[2], sd[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float m[2] = {mi[c], mt[c]};
          const float lv[2] = {li[c], lt[c]};
          r[c] = poe_eval<false>(a.poe_mode, a.prior_expert, a.poe_eps, m, lv, present);
          sd[c] = sqrtf(r[c].pd_var);
          zz[c] = a.training ? ee[c] * sd[c] + r[c].mu : r[c].mu;
        }
        // text decoder first Linear: dz += dT1 * Wt1, dWt1 += dT1^T z
        if (text_dec) {
#pragma unroll
          for (int j = 0; j < kTD; ++j) {
            const float dj = dtb[j];
            const float2 w = *reinterpret_cast<const float2*>(a.wt1 + j * n + k);
            dzz[0] = fmaf(dj, w.x, dzz[0]);
            dzz[1] = fmaf(dj, w.y, dzz[1]);
            if (first_pass) {
              r_w1[j][0] = fmaf(dj, zz[0], r_w1[j][0]);
              r_w1[j][1] = fmaf(dj, zz[1], r_w1[j][1]);
            } else {
              atomicAdd(&s_w1[j * n + k], dj * zz[0]);
              atomicAdd(&s_w1[j * n + k + 1], dj * zz[1]);
            }
          }
        }
        float2 dmu_up = make_float2(0.f, 0.f), dlv_up = make_float2(0.f, 0.f);
        if (a.dmu_up != nullptr) dmu_up = *reinterpret_cast<const float2*>(a.dmu_up + row * n + k);
        if (a.dlogvar_up != nullptr) dlv_up = *reinterpret_cast<const float2*>(a.dlogvar_up + row * n + k);
        const float dmu_u[2] = {dmu_up.x, dmu_up.y}, dlv_u[2] = {dlv_up.x, dlv_up.y};
        float d_mi[2] = {0.f, 0.f}, d_li[2] = {0.f, 0.f}, d_mt[2] = {0.f, 0.f}, d_lt[2] = {0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          // z = mu + eps*exp(logvar/2);  KL = c * -0.5 * sum(1 + logvar - mu^2 - exp(logvar))
          const float dmu = dzz[c] + dmu_u[c] + c_kl * r[c].mu;
          float dlv = dlv_u[c] + 0.5f * c_kl * (r[c].pd_var - 1.f);
          if (a.training) dlv += dzz[c] * 0.5f * ee[c] * sd[c];
          if (present[0]) poe_grad(a.poe_mode, a.poe_eps, r[c], 0, mi[c], dmu, dlv, d_mi[c], d_li[c]);
          if (present[1]) poe_grad(a.poe_mode, a.poe_eps, r[c], 1, mt[c], dmu, dlv, d_mt[c], d_lt[c]);
        }
        if (present[0]) {
          float* cb = s_cb + (wrow * kMaxGroups + g) * two_n;
          cb[k] = d_mi[0]; cb[k + 1] = d_mi[1]; cb[n + k] = d_li[0]; cb[n + k + 1] = d_li[1];
        }
        if (present[1]) {
          float* ct = s_ct + (wrow * kMaxGroups + g) * two_n;
          ct[k] = d_mt[0]; ct[k + 1] = d_mt[1]; ct[n + k] = d_lt[0]; ct[n + k + 1] = d_lt[1];
          if (lane == 0) s_lab[wrow] = label;
        }
      }
    }
    __syncthreads();
    // text expert: scatter-add by label without atomics - thread t owns column t of the [10][2n] table
    if (a.d_txt_table != nullptr && a.txt_table != nullptr) {
      for (int t = threadIdx.x; t < two_n; t += blockDim.x) {
        for (int r = 0; r < kBwdRows; ++r) {
          if (rb * kBwdRows + r >= a.B) break;
          const int lab = s_lab[r];
          float v = 0.f;
          for (int gg = 0; gg < G; ++gg)
            if (a.group_type[gg] != TERM_IMAGE) v += s_ct[(r * kMaxGroups + gg) * two_n + t];
          s_tab[lab * two_n + t] += v;
        }
      }
    }
    // combine the terms' contributions to this sample's image-expert gradient (warp g == 0 of each sample)
    if (active && g == 0 && a.enc_img != nullptr && a.d_enc != nullptr) {
      ZT* de = reinterpret_cast<ZT*>(a.d_enc) + static_cast<long long>(b) * two_n;
      for (int k = lane * 2; k < n; k += 64) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int gg = 0; gg < G; ++gg) {
          if (a.group_type[gg] == TERM_TEXT) continue;
          const float* cb = s_cb + (wrow * kMaxGroups + gg) * two_n;
          v[0] += cb[k]; v[1] += cb[k + 1]; v[2] += cb[n + k]; v[3] += cb[n + k + 1];
        }
        store_pair(de + k, v[0], v[1]);
        store_pair(de + n + k, v[2], v[3]);
        if (k < 64) {
          r_eb[0] += v[0]; r_eb[1] += v[1]; r_eb[2] += v[2]; r_eb[3] += v[3];
        } else {
          atomicAdd(&s_eb[k], v[0]);
          atomicAdd(&s_eb[k + 1], v[1]);
          atomicAdd(&s_eb[n + k], v[2]);
          atomicAdd(&s_eb[n + k + 1], v[3]);
        }
      }
    }
    __syncthreads();
  }
  {
    const int k = lane * 2;
    if (k < n) {
      if (text_dec) {
#pragma unroll
        for (int j = 0; j < kTD; ++j) {
          atomicAdd(&s_w1[j * n + k], r_w1[j][0]);
          atomicAdd(&s_w1[j * n + k + 1], r_w1[j][1]);
        }
      }
      if (g == 0) {
        atomicAdd(&s_eb[k], r_eb[0]);
        atomicAdd(&s_eb[k + 1], r_eb[1]);
        atomicAdd(&s_eb[n + k], r_eb[2]);
        atomicAdd(&s_eb[n + k + 1], r_eb[3]);
      }
    }
  }
  __syncthreads();
  if (a.d_txt_table != nullptr)
    for (int i = threadIdx.x; i < kTD * two_n; i += blockDim.x)
      if (s_tab[i] != 0.f) atomicAdd(a.d_txt_table + i, s_tab[i]);
  if (text_dec && a.d_wt1 != nullptr)
    for (int i = threadIdx.x; i < kTD * n; i += blockDim.x) atomicAdd(a.d_wt1 + i, s_w1[i]);
  if (a.d_enc_bias != nullptr && a.enc_img != nullptr)
    for (int i = threadIdx.x; i < two_n; i += blockDim.x) atomicAdd(a.d_enc_bias + i, s_eb[i]);
  // BatchNorm affine gradients of the text decoder: dgamma = sum_g S1, dbeta = sum_g S0 (block 0 only)
  if (text_dec && blockIdx.x == 0 && threadIdx.x < kTD && a.d_t1_gamma != nullptr) {
    float dg = 0.f, db = 0.f;
    for (int gg = 0; gg < G; ++gg) {
      dg += a.t1_s1[gg * kTD + threadIdx.x];
      db += a.t1_s0[gg * kTD + threadIdx.x];
    }
    a.d_t1_gamma[threadIdx.x] += dg;
    a.d_t1_beta[threadIdx.x] += db;
  }
}

// ================================================================= tail backward, one warp per SAMPLE (n <= 64)
// The warp carries its sample through all ELBO terms: both experts are loaded and exponentiated once (poe_part / poe_combine),
// the draws of the forward are read back from noise_buf (no second Philox pass), the terms' contributions to the image
// expert's gradient are summed in registers (no shared-memory combine, no block barrier inside the loop), the text expert's
// gradient goes into the block's [10][2n] table with shared-memory adds, and the text decoder's first Linear keeps its weights
// in shared memory and its weight gradient in per-lane registers.  Same arithmetic as tail_bwd_kernel.
template <typename ZT, int kBwd2Warps>
__global__ void __launch_bounds__(32 * kBwd2Warps, 28 / kBwd2Warps) tail_bwd2_kernel(const TailArgs a) {
  ptx::griddep_launch();   // programmatic launch (common.cuh, g_pdl_next): the next kernel's prologue may overlap this one
  ptx::griddep_wait();     // ... and this one starts while its predecessor drains; no global access before this line
  extern __shared__ float sm[];
  const int n = a.n, two_n = 2 * a.n, G = a.G;
  // layout: text-expert table [10][2n] | Wt1 [10][n] | t1 coefficients [G][4][10] | d_wt1 [10][n] | d_enc_bias [2n]
  float* s_tab = sm;
  float* s_w = s_tab + kTD * two_n;
  float* s_co = s_w + kTD * n;
  float* s_dw = s_co + kMaxGroups * 4 * kTD;
  float* s_eb = s_dw + kTD * n;
  const int total = kTD * two_n + kTD * n + kMaxGroups * 4 * kTD + kTD * n + two_n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < total; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const bool text_dec = a.t1_dyhat != nullptr;
  if (text_dec) {
    for (int i = threadIdx.x; i < kTD * n; i += blockDim.x) s_w[i] = a.wt1[i];
    if (threadIdx.x < G * kTD) {
      const int gg = threadIdx.x / kTD, j = threadIdx.x % kTD;
      const float mean = a.t1_sum[gg * kTD + j] / a.B;
      const float var = fmaxf(a.t1_sumsq[gg * kTD + j] / a.B - mean * mean, 0.f);
      s_co[(gg * 4 + 0) * kTD + j] = mean;
      s_co[(gg * 4 + 1) * kTD + j] = rsqrtf(var + 1e-5f);
      s_co[(gg * 4 + 2) * kTD + j] = a.t1_s0[gg * kTD + j] / a.B;
      s_co[(gg * 4 + 3) * kTD + j] = a.t1_s1[gg * kTD + j] / a.B;
    }
  }
  __syncthreads();
  const uint32_t step = a.step_ptr != nullptr ? static_cast<uint32_t>(*a.step_ptr) : 0u;
  bool any_img = false, any_txt = false;
  for (int g = 0; g < G; ++g) {
    any_img |= a.group_type[g] != TERM_TEXT;
    any_txt |= a.group_type[g] != TERM_IMAGE;
  }
  const int k = lane * 2;
  const bool act = k < n;
  float r_w1[kTD][2], r_eb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < kTD; ++j) r_w1[j][0] = r_w1[j][1] = 0.f;
  // this lane's (term, feature) of the text decoder's BatchNorm backward: loop constants
  float co_mean = 0.f, co_rstd = 0.f, co_s0 = 0.f, co_s1 = 0.f, co_gamma = 0.f;
  int my_g = 0, my_j = 0;
  if (text_dec && lane < G * kTD) {
    my_g = lane / kTD;
    my_j = lane - my_g * kTD;
    co_mean = s_co[(my_g * 4 + 0) * kTD + my_j];
    co_rstd = s_co[(my_g * 4 + 1) * kTD + my_j];
    co_s0 = s_co[(my_g * 4 + 2) * kTD + my_j];
    co_s1 = s_co[(my_g * 4 + 3) * kTD + my_j];
    co_gamma = a.t1_gamma[my_j];
  }

  for (int b = blockIdx.x * kBwd2Warps + warp; b < a.B; b += gridDim.x * kBwd2Warps) {
    const int label = any_txt ? static_cast<int>(a.labels[b]) : 0;
    // text decoder: gradient at its first Linear's output (BatchNorm backward apply) for every term: lane = term * 10 + feature
    float dt_mine = 0.f;
    if (text_dec && lane < G * kTD) {
      const long long row = static_cast<long long>(my_g) * a.B + b;
      const float x = a.t1pre[row * kTD + my_j];
      const float xh = (x - co_mean) * co_rstd;
      const float dy = a.t1_dyhat[row * kTD + my_j];
      dt_mine = co_gamma * co_rstd * (dy - co_s0 - xh * co_s1);
    }
    float mi[2] = {0.f, 0.f}, mt[2] = {0.f, 0.f};
    PoePart pi[2] = {{1.f, 1.f}, {1.f, 1.f}}, pt[2] = {{1.f, 1.f}, {1.f, 1.f}};
    if (act && any_img) {
      const float2 m2 = *reinterpret_cast<const float2*>(a.enc_img + static_cast<long long>(b) * two_n + k);
      const float2 l2 = *reinterpret_cast<const float2*>(a.enc_img + static_cast<long long>(b) * two_n + n + k);
      mi[0] = m2.x; mi[1] = m2.y;
      pi[0] = poe_part(l2.x, a.poe_eps); pi[1] = poe_part(l2.y, a.poe_eps);
    }
    if (act && any_txt) {
      const float2 m2 = *reinterpret_cast<const float2*>(a.txt_table + label * two_n + k);
      const float2 l2 = *reinterpret_cast<const float2*>(a.txt_table + label * two_n + n + k);
      mt[0] = m2.x; mt[1] = m2.y;
      pt[0] = poe_part(l2.x, a.poe_eps); pt[1] = poe_part(l2.y, a.poe_eps);
    }
    float a_mi[2] = {0.f, 0.f}, a_li[2] = {0.f, 0.f}, a_mt[2] = {0.f, 0.f}, a_lt[2] = {0.f, 0.f};
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      if (g >= G) break;
      const int ty = a.group_type[g];
      const bool present[2] = {ty != TERM_TEXT, ty != TERM_IMAGE};
      const float c_kl = a.kl_weight[g];
      const long long row = static_cast<long long>(g) * a.B + b;
      float2 e2 = make_float2(0.f, 0.f), dz2 = make_float2(0.f, 0.f), dmu_up = make_float2(0.f, 0.f), dlv_up = make_float2(0.f, 0.f);
      if (act) {
        if (a.training) {
          if (a.eps != nullptr)
            e2 = *reinterpret_cast<const float2*>(a.eps + row * n + k);
          else if (a.noise_buf != nullptr)
            e2 = *reinterpret_cast<const float2*>(a.noise_buf + row * n + k);
          else
            e2 = normal_pair(a.seed, step, (row * n + k) >> 1);
        }
        if (a.dz != nullptr) dz2 = *reinterpret_cast<const float2*>(a.dz + row * n + k);
        if (a.dmu_up != nullptr) dmu_up = *reinterpret_cast<const float2*>(a.dmu_up + row * n + k);
        if (a.dlogvar_up != nullptr) dlv_up = *reinterpret_cast<const float2*>(a.dlogvar_up + row * n + k);
      }
      const float ee[2] = {e2.x, e2.y};
      float dzz[2] = {dz2.x, dz2.y};
      const float dmu_u[2] = {dmu_up.x, dmu_up.y}, dlv_u[2] = {dlv_up.x, dlv_up.y};
      Poe r[2];
      float zz[2], sd[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float m[2] = {present[0] ? mi[c] : 0.f, present[1] ? mt[c] : 0.f};
        const PoePart parts[2] = {pi[c], pt[c]};
        r[c] = poe_combine<false>(a.poe_mode, a.prior_expert, m, parts, present);
        sd[c] = sqrtf(r[c].pd_var);
        zz[c] = a.training ? ee[c] * sd[c] + r[c].mu : r[c].mu;
      }
      // text decoder first Linear: dz += dT1 * Wt1, dWt1 += dT1^T z
      if (text_dec) {
#pragma unroll
        for (int j = 0; j < kTD; ++j) {
          const float dj = __shfl_sync(0xffffffffu, dt_mine, g * kTD + j);
          if (act) {
            const float2 w = *reinterpret_cast<const float2*>(s_w + j * n + k);
            dzz[0] = fmaf(dj, w.x, dzz[0]);
            dzz[1] = fmaf(dj, w.y, dzz[1]);
            r_w1[j][0] = fmaf(dj, zz[0], r_w1[j][0]);
            r_w1[j][1] = fmaf(dj, zz[1], r_w1[j][1]);
          }
        }
      }
      if (act) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          // z = mu + eps*exp(logvar/2);  KL = c * -0.5 * sum(1 + logvar - mu^2 - exp(logvar))
          const float dmu = dzz[c] + dmu_u[c] + c_kl * r[c].mu;
          float dlv = dlv_u[c] + 0.5f * c_kl * (r[c].pd_var - 1.f);
          if (a.training) dlv += dzz[c] * 0.5f * ee[c] * sd[c];
          float dm = 0.f, dl = 0.f;
          if (present[0]) {
            poe_grad(a.poe_mode, a.poe_eps, r[c], 0, mi[c], dmu, dlv, dm, dl);
            a_mi[c] += dm;
            a_li[c] += dl;
          }
          if (present[1]) {
            poe_grad(a.poe_mode, a.poe_eps, r[c], 1, mt[c], dmu, dlv, dm, dl);
            a_mt[c] += dm;
            a_lt[c] += dl;
          }
        }
      }
    }
    if (act) {
      if (any_img && a.enc_img != nullptr && a.d_enc != nullptr) {
        ZT* de = reinterpret_cast<ZT*>(a.d_enc) + static_cast<long long>(b) * two_n;
        store_pair(de + k, a_mi[0], a_mi[1]);
        store_pair(de + n + k, a_li[0], a_li[1]);
        r_eb[0] += a_mi[0]; r_eb[1] += a_mi[1]; r_eb[2] += a_li[0]; r_eb[3] += a_li[1];
      }
      if (any_txt) {
        float* tm = s_tab + label * two_n + k;
        atomicAdd(tm, a_mt[0]);
        atomicAdd(tm + 1, a_mt[1]);
        atomicAdd(tm + n, a_lt[0]);
        atomicAdd(tm + n + 1, a_lt[1]);
      }
    }
  }
  if (act) {
    if (text_dec) {
#pragma unroll
      for (int j = 0; j < kTD; ++j) {
        atomicAdd(&s_dw[j * n + k], r_w1[j][0]);
        atomicAdd(&s_dw[j * n + k + 1], r_w1[j][1]);
      }
    }
    atomicAdd(&s_eb[k], r_eb[0]);
    atomicAdd(&s_eb[k + 1], r_eb[1]);
    atomicAdd(&s_eb[n + k], r_eb[2]);
    atomicAdd(&s_eb[n + k + 1], r_eb[3]);
  }
  __syncthreads();
  if (a.d_txt_table != nullptr && a.txt_table != nullptr)
    for (int i = threadIdx.x; i < kTD * two_n; i += blockDim.x)
      if (s_tab[i] != 0.f) atomicAdd(a.d_txt_table + i, s_tab[i]);
  if (text_dec && a.d_wt1 != nullptr)
    for (int i = threadIdx.x; i < kTD * n; i += blockDim.x) atomicAdd(a.d_wt1 + i, s_dw[i]);
  if (a.d_enc_bias != nullptr && a.enc_img != nullptr)
    for (int i = threadIdx.x; i < two_n; i += blockDim.x) atomicAdd(a.d_enc_bias + i, s_eb[i]);
  // BatchNorm affine gradients of the text decoder: dgamma = sum_g S1, dbeta = sum_g S0 (block 0 only)
  if (text_dec && blockIdx.x == 0 && threadIdx.x < kTD && a.d_t1_gamma != nullptr) {
    float dg = 0.f, db = 0.f;
    for (int gg = 0; gg < G; ++gg) {
      dg += a.t1_s1[gg * kTD + threadIdx.x];
      db += a.t1_s0[gg * kTD + threadIdx.x];
    }
    a.d_t1_gamma[threadIdx.x] += dg;
    a.d_t1_beta[threadIdx.x] += db;
  }
}

// ================================================================= the hot configuration, specialised (n == 64, training)
// tail_fwd3 / tail_bwd3: what tail_fwd2 / tail_bwd2 do, with everything that is uniform in the benchmarked step resolved at
// compile time - n = 64 (every lane owns one float2 of latents; shared-memory and row offsets are immediates), train mode, the
// PoE arithmetic as a template parameter, the text decoder's first Linear present, both experts present in the workspace, no
// caller-provided latents / upstream gradients / mu-logvar outputs.  The generic kernels spend ~80 % of their ~2 200
// instructions per sample on address arithmetic, constant-bank loads and branches around those options (ncu source view).
// No shared-memory atomics (fp32 adds on shared memory are compare-and-swap loops on this architecture): per-lane registers
// and per-warp tables, reduced across the block's warps once at the end.
constexpr int kT3Warps = 14;   // 2 blocks x 14 warps per SM: the 4096 samples of the benchmarked batch are resident at once
constexpr int kT3N = 64;

template <typename ZT, int kMode>
__global__ void __launch_bounds__(32 * kT3Warps, 2) tail_fwd3_kernel(const TailArgs a) {
  ptx::griddep_launch();   // programmatic launch (common.cuh, g_pdl_next): the next kernel's prologue may overlap this one
  ptx::griddep_wait();     // ... and this one starts while its predecessor drains; no global access before this line
  constexpr int n = kT3N, two_n = 2 * kT3N;
  __shared__ float s_red[kT3Warps][64 + kMaxGroups];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = a.G, B = a.B, k = lane * 2;
  const uint32_t step = a.step_ptr != nullptr ? static_cast<uint32_t>(*a.step_ptr) : 0u;
  bool p_img[kMaxGroups], p_txt[kMaxGroups];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    p_img[g] = g < G && a.group_type[g] != TERM_TEXT;
    p_txt[g] = g < G && a.group_type[g] != TERM_IMAGE;
  }
  const float poe_eps = a.poe_eps;
  const int prior = a.prior_expert;
  const float* __restrict__ enc_img = a.enc_img;
  const float* __restrict__ txt_table = a.txt_table;
  const long long* __restrict__ labels = a.labels;
  const float* __restrict__ eps = a.eps;
  float* __restrict__ noise_buf = a.noise_buf;
  ZT* __restrict__ zout = reinterpret_cast<ZT*>(a.z);
  float* __restrict__ mu_out = a.mu;
  float* __restrict__ lv_out = a.logvar;
  float2 w[kTD];
#pragma unroll
  for (int j = 0; j < kTD; ++j) w[j] = *reinterpret_cast<const float2*>(a.wt1 + j * n + k);
  // lane L < 10 G owns (term L / 10, feature L % 10) of the text decoder's first Linear after the butterfly
  const bool owner = lane < G * kTD;
  const int my_g = owner ? lane / kTD : 0, my_j = owner ? lane - my_g * kTD : 0;
  const float my_bias = owner ? a.bt1[my_j] : 0.f;
  float* __restrict__ my_t1 = a.t1pre + static_cast<long long>(my_g) * B * kTD + my_j;
  float st_sum = 0.f, st_sq = 0.f;
  float klacc[kMaxGroups] = {0.f, 0.f, 0.f};

  for (int b = blockIdx.x * kT3Warps + warp; b < B; b += gridDim.x * kT3Warps) {
    const float2 im = *reinterpret_cast<const float2*>(enc_img + static_cast<long long>(b) * two_n + k);
    const float2 il = *reinterpret_cast<const float2*>(enc_img + static_cast<long long>(b) * two_n + n + k);
    const int label = static_cast<int>(labels[b]);
    const float2 tm = *reinterpret_cast<const float2*>(txt_table + label * two_n + k);
    const float2 tl = *reinterpret_cast<const float2*>(txt_table + label * two_n + n + k);
    const float mi[2] = {im.x, im.y}, mt[2] = {tm.x, tm.y};
    const PoePart pi[2] = {poe_part(il.x, poe_eps), poe_part(il.y, poe_eps)};
    const PoePart pt[2] = {poe_part(tl.x, poe_eps), poe_part(tl.y, poe_eps)};
    // the draws of all terms: injected, or two Philox calls per lane (four normals each: terms 0 and 1, term 2)
    float2 e_all[kMaxGroups];
    if (eps != nullptr) {
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g)
        e_all[g] = g < G ? *reinterpret_cast<const float2*>(eps + (static_cast<long long>(g) * B + b) * n + k) : make_float2(0.f, 0.f);
    } else {
      const unsigned long long pair = (static_cast<unsigned long long>(b) * n + k) >> 1;
      const float4 q = normal_quad(a.seed, step, pair, 0x6d766166u);
      e_all[0] = make_float2(q.x, q.y);
      e_all[1] = make_float2(q.z, q.w);
      e_all[2] = make_float2(0.f, 0.f);
      if (G > 2) {
        const float4 q2 = normal_quad(a.seed, step, pair, 0x6d766167u);
        e_all[2] = make_float2(q2.x, q2.y);
      }
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g)
        if (g < G) *reinterpret_cast<float2*>(noise_buf + (static_cast<long long>(g) * B + b) * n + k) = e_all[g];
    }
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      if (g < G) {
        const long long e_off = (static_cast<long long>(g) * B + b) * n + k;
        const float2 e2 = e_all[g];
        const float ee[2] = {e2.x, e2.y};
        const bool present[2] = {p_img[g], p_txt[g]};
        float zz[2], mm[2], ll[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float m[2] = {mi[c], mt[c]};
          const PoePart parts[2] = {pi[c], pt[c]};
          const Poe r = poe_combine<true>(kMode, prior, m, parts, present);
          mm[c] = r.mu;
          ll[c] = r.logvar;
          zz[c] = ee[c] * sqrtf(r.pd_var) + r.mu;                    // reparametrize (mnist/model.py:25-28)
          klacc[g] += 1.f + r.logvar - r.mu * r.mu - r.pd_var;       // KL integrand of mnist/train.py:79
        }
        store_pair(zout + e_off, zz[0], zz[1]);
        if (mu_out != nullptr) {
          *reinterpret_cast<float2*>(mu_out + e_off) = make_float2(mm[0], mm[1]);
          *reinterpret_cast<float2*>(lv_out + e_off) = make_float2(ll[0], ll[1]);
        }
#pragma unroll
        for (int j = 0; j < kTD; ++j) v[g * kTD + j] = fmaf(zz[0], w[j].x, zz[1] * w[j].y);
      }
    }
    // transposing butterfly: afterwards lane L holds the warp-wide sum of value L
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
      const bool up = (lane & h) != 0;
#pragma unroll
      for (int i = 0; i < h; ++i) {
        const float send = up ? v[i] : v[i + h];
        const float keep = up ? v[i + h] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
      }
    }
    if (owner) {
      const float t = v[0] + my_bias;
      my_t1[static_cast<long long>(b) * kTD] = t;
      st_sum += t;
      st_sq = fmaf(t, t, st_sq);
    }
  }
  // block reduction of the BatchNorm statistics of the text decoder and of the KL partials, one global atomic per address
  s_red[warp][lane] = st_sum;
  s_red[warp][32 + lane] = st_sq;
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    const float ks = warp_sum(klacc[g]);
    if (lane == 0) s_red[warp][64 + g] = ks;
  }
  __syncthreads();
  if (threadIdx.x < 64 + kMaxGroups) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < kT3Warps; ++ww) t += s_red[ww][threadIdx.x];
    const int i = threadIdx.x;
    if (i < 64) {
      const int l = i & 31;
      if (l < G * kTD && t != 0.f) atomicAdd((i < 32 ? a.t1_sum : a.t1_sumsq) + l, t);   // [G][10] == lane index
    } else if (i - 64 < G && a.kl != nullptr && t != 0.f) {
      atomicAdd(a.kl + (i - 64), -0.5f * a.kl_weight[i - 64] * t);
    }
  }
}

template <typename ZT, int kMode>
__global__ void __launch_bounds__(32 * kT3Warps, 2) tail_bwd3_kernel(const TailArgs a) {
  ptx::griddep_launch();   // programmatic launch (common.cuh, g_pdl_next): the next kernel's prologue may overlap this one
  ptx::griddep_wait();     // ... and this one starts while its predecessor drains; no global access before this line
  constexpr int n = kT3N, two_n = 2 * kT3N;
  constexpr int kTab = kTD * two_n;   // floats per warp table
  extern __shared__ float sm[];       // per-warp text-expert tables [warps][10][2n] | Wt1 [10][n]
  float* s_w = sm + kT3Warps * kTab;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = a.G, B = a.B, k = lane * 2;
  for (int i = threadIdx.x; i < kT3Warps * kTab; i += blockDim.x) sm[i] = 0.f;
  for (int i = threadIdx.x; i < kTD * n; i += blockDim.x) s_w[i] = a.wt1[i];
  __syncthreads();
  bool p_img[kMaxGroups], p_txt[kMaxGroups];
  float c_kl[kMaxGroups];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    p_img[g] = g < G && a.group_type[g] != TERM_TEXT;
    p_txt[g] = g < G && a.group_type[g] != TERM_IMAGE;
    c_kl[g] = g < G ? a.kl_weight[g] : 0.f;
  }
  const float poe_eps = a.poe_eps;
  const int prior = a.prior_expert;
  const float* __restrict__ enc_img = a.enc_img;
  const float* __restrict__ txt_table = a.txt_table;
  const long long* __restrict__ labels = a.labels;
  const float* __restrict__ noise = a.eps != nullptr ? a.eps : a.noise_buf;
  const float* __restrict__ dz = a.dz;
  ZT* __restrict__ d_enc = reinterpret_cast<ZT*>(a.d_enc);
  // this lane's (term, feature) of the text decoder's BatchNorm backward: dT1 = gamma rstd (dy - S0/B - xhat S1/B)
  const bool owner = lane < G * kTD;
  const int my_g = owner ? lane / kTD : 0, my_j = owner ? lane - my_g * kTD : 0;
  float co_mean = 0.f, co_rstd = 0.f, co_s0 = 0.f, co_s1 = 0.f, co_gr = 0.f;
  if (owner) {
    const float inv_b = 1.f / static_cast<float>(B);
    co_mean = a.t1_sum[lane] / B;
    const float var = fmaxf(a.t1_sumsq[lane] / B - co_mean * co_mean, 0.f);
    co_rstd = rsqrtf(var + 1e-5f);
    co_s0 = a.t1_s0[lane] / B;
    co_s1 = a.t1_s1[lane] / B;
    co_gr = a.t1_gamma[my_j] * co_rstd;
    (void)inv_b;
  }
  const float* __restrict__ my_t1 = a.t1pre + static_cast<long long>(my_g) * B * kTD + my_j;
  const float* __restrict__ my_dy = a.t1_dyhat + static_cast<long long>(my_g) * B * kTD + my_j;
  float r_w1[kTD][2], r_eb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < kTD; ++j) r_w1[j][0] = r_w1[j][1] = 0.f;
  float* my_tab = sm + warp * kTab + k;

  for (int b = blockIdx.x * kT3Warps + warp; b < B; b += gridDim.x * kT3Warps) {
    // every global load of the sample is issued before the first use (one sample per warp: nothing else hides the latency)
    float2 e_all[kMaxGroups], dz_all[kMaxGroups];
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      const long long e_off = (static_cast<long long>(g < G ? g : 0) * B + b) * n + k;
      e_all[g] = *reinterpret_cast<const float2*>(noise + e_off);
      dz_all[g] = *reinterpret_cast<const float2*>(dz + e_off);
    }
    float dt_mine = 0.f;
    if (owner) {
      const float x = my_t1[static_cast<long long>(b) * kTD];
      const float dy = my_dy[static_cast<long long>(b) * kTD];
      const float xh = (x - co_mean) * co_rstd;
      dt_mine = co_gr * (dy - co_s0 - xh * co_s1);
    }
    const float2 im = *reinterpret_cast<const float2*>(enc_img + static_cast<long long>(b) * two_n + k);
    const float2 il = *reinterpret_cast<const float2*>(enc_img + static_cast<long long>(b) * two_n + n + k);
    const int label = static_cast<int>(labels[b]);
    const float2 tm = *reinterpret_cast<const float2*>(txt_table + label * two_n + k);
    const float2 tl = *reinterpret_cast<const float2*>(txt_table + label * two_n + n + k);
    const float mi[2] = {im.x, im.y}, mt[2] = {tm.x, tm.y};
    const PoePart pi[2] = {poe_part(il.x, poe_eps), poe_part(il.y, poe_eps)};
    const PoePart pt[2] = {poe_part(tl.x, poe_eps), poe_part(tl.y, poe_eps)};
    float a_mi[2] = {0.f, 0.f}, a_li[2] = {0.f, 0.f}, a_mt[2] = {0.f, 0.f}, a_lt[2] = {0.f, 0.f};
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      if (g < G) {
        const float2 e2 = e_all[g];
        const float2 dz2 = dz_all[g];
        const float ee[2] = {e2.x, e2.y};
        float dzz[2] = {dz2.x, dz2.y};
        const bool present[2] = {p_img[g], p_txt[g]};
        Poe r[2];
        float zz[2], sd[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float m[2] = {present[0] ? mi[c] : 0.f, present[1] ? mt[c] : 0.f};
          const PoePart parts[2] = {pi[c], pt[c]};
          r[c] = poe_combine<false>(kMode, prior, m, parts, present);
          sd[c] = sqrtf(r[c].pd_var);
          zz[c] = ee[c] * sd[c] + r[c].mu;
        }
        // text decoder first Linear: dz += dT1 * Wt1, dWt1 += dT1^T z
#pragma unroll
        for (int j = 0; j < kTD; ++j) {
          const float dj = __shfl_sync(0xffffffffu, dt_mine, g * kTD + j);
          const float2 wj = *reinterpret_cast<const float2*>(s_w + j * n + k);
          dzz[0] = fmaf(dj, wj.x, dzz[0]);
          dzz[1] = fmaf(dj, wj.y, dzz[1]);
          r_w1[j][0] = fmaf(dj, zz[0], r_w1[j][0]);
          r_w1[j][1] = fmaf(dj, zz[1], r_w1[j][1]);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          // z = mu + eps*exp(logvar/2);  KL = c * -0.5 * sum(1 + logvar - mu^2 - exp(logvar))
          const float dmu = dzz[c] + 0.f + c_kl[g] * r[c].mu;
          float dlv = 0.f + 0.5f * c_kl[g] * (r[c].pd_var - 1.f);
          dlv += dzz[c] * 0.5f * ee[c] * sd[c];
          float dm = 0.f, dl = 0.f;
          if (present[0]) {
            poe_grad(kMode, poe_eps, r[c], 0, mi[c], dmu, dlv, dm, dl);
            a_mi[c] += dm;
            a_li[c] += dl;
          }
          if (present[1]) {
            poe_grad(kMode, poe_eps, r[c], 1, mt[c], dmu, dlv, dm, dl);
            a_mt[c] += dm;
            a_lt[c] += dl;
          }
        }
      }
    }
    ZT* de = d_enc + static_cast<long long>(b) * two_n + k;
    store_pair(de, a_mi[0], a_mi[1]);
    store_pair(de + n, a_li[0], a_li[1]);
    r_eb[0] += a_mi[0]; r_eb[1] += a_mi[1]; r_eb[2] += a_li[0]; r_eb[3] += a_li[1];
    float2* tmp_m = reinterpret_cast<float2*>(my_tab + label * two_n);
    float2* tmp_l = reinterpret_cast<float2*>(my_tab + label * two_n + n);
    float2 vm = *tmp_m, vl = *tmp_l;
    vm.x += a_mt[0]; vm.y += a_mt[1]; vl.x += a_lt[0]; vl.y += a_lt[1];
    *tmp_m = vm;
    *tmp_l = vl;
  }
  __syncthreads();
  // the block's text-expert table: sum of the warps' tables, one global reduction per touched entry
  for (int i = threadIdx.x; i < kTab; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < kT3Warps; ++ww) t += sm[ww * kTab + i];
    if (t != 0.f) atomicAdd(a.d_txt_table + i, t);
  }
  __syncthreads();
  // weight gradient of the text decoder's first Linear [10][n] and the bias gradient of the image encoder's last Linear [2n]:
  // per-lane registers -> this warp's (now free) table -> sum over the warps
  {
    float* mine = sm + warp * kTab;
#pragma unroll
    for (int j = 0; j < kTD; ++j) *reinterpret_cast<float2*>(mine + j * n + k) = make_float2(r_w1[j][0], r_w1[j][1]);
    *reinterpret_cast<float2*>(mine + kTD * n + k) = make_float2(r_eb[0], r_eb[1]);
    *reinterpret_cast<float2*>(mine + kTD * n + n + k) = make_float2(r_eb[2], r_eb[3]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kTD * n + two_n; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < kT3Warps; ++ww) t += sm[ww * kTab + i];
    if (i < kTD * n) atomicAdd(a.d_wt1 + i, t);
    else if (a.d_enc_bias != nullptr) atomicAdd(a.d_enc_bias + (i - kTD * n), t);
  }
  // BatchNorm affine gradients of the text decoder: dgamma = sum_g S1, dbeta = sum_g S0 (block 0 only)
  if (blockIdx.x == 0 && threadIdx.x < kTD && a.d_t1_gamma != nullptr) {
    float dg = 0.f, db = 0.f;
    for (int gg = 0; gg < G; ++gg) {
      dg += a.t1_s1[gg * kTD + threadIdx.x];
      db += a.t1_s0[gg * kTD + threadIdx.x];
    }
    a.d_t1_gamma[threadIdx.x] += dg;
    a.d_t1_beta[threadIdx.x] += db;
  }
}

// ================================================================= text decoder (BN -> ReLU -> Linear -> log_softmax -> NLL)
// One thread per decoder row.  Forward + (optionally) the fused NLL loss and the backward down to the
// BatchNorm output: dyhat, its two per-group column sums, and the gradients of the second Linear.
__global__ void __launch_bounds__(256) textdec_kernel(const TextDecArgs a) {
  __shared__ float s_mean[kMaxGroups][kTD], s_rstd[kMaxGroups][kTD];
  __shared__ float s_w2[kTD][kTD], s_b2[kTD], s_gamma[kTD], s_beta[kTD];
  __shared__ float s_dw2[kTD][kTD], s_db2[kTD], s_s0[kMaxGroups][kTD], s_s1[kMaxGroups][kTD], s_ce[kMaxGroups];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < a.G * kTD) {
    const int g = tid / kTD, j = tid % kTD;
    float mean, rstd;
    if (a.training) {
      mean = a.t1_sum[tid] / a.B;
      const float var = fmaxf(a.t1_sumsq[tid] / a.B - mean * mean, 0.f);
      rstd = rsqrtf(var + a.bn_eps);
    } else {
      mean = a.running_mean[j];
      rstd = rsqrtf(a.running_var[j] + a.bn_eps);
    }
    s_mean[g][j] = mean;
    s_rstd[g][j] = rstd;
  }
  if (tid < kTD * kTD) {
    s_w2[tid / kTD][tid % kTD] = a.w2[tid];
    s_dw2[tid / kTD][tid % kTD] = 0.f;
  }
  if (tid < kTD) {
    s_b2[tid] = a.b2[tid];
    s_gamma[tid] = a.gamma[tid];
    s_beta[tid] = a.beta[tid];
    s_db2[tid] = 0.f;
  }
  if (tid < kMaxGroups * kTD) {
    (&s_s0[0][0])[tid] = 0.f;
    (&s_s1[0][0])[tid] = 0.f;
  }
  if (tid < kMaxGroups) s_ce[tid] = 0.f;
  __syncthreads();
  // running statistics: one update per group, in group order (block 0)
  if (a.training && blockIdx.x == 0 && tid < kTD && a.running_mean != nullptr) {
    float rm = a.running_mean[tid], rv = a.running_var[tid];
    for (int g = 0; g < a.G; ++g) {
      const float mean = s_mean[g][tid];
      const float var = fmaxf(a.t1_sumsq[g * kTD + tid] / a.B - mean * mean, 0.f);
      const float unb = a.B > 1 ? var * (static_cast<float>(a.B) / (a.B - 1)) : var;
      rm = (1.f - a.momentum) * rm + a.momentum * mean;
      rv = (1.f - a.momentum) * rv + a.momentum * unb;
    }
    a.running_mean[tid] = rm;
    a.running_var[tid] = rv;
  }
  const long long rows = static_cast<long long>(a.G) * a.B;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x) + tid;
  const bool valid = row < rows;
  const int g = valid ? static_cast<int>(row / a.B) : 0;
  float xh[kTD], t1[kTD], lp[kTD], dl[kTD], dy[kTD];
  float ce = 0.f;
#pragma unroll
  for (int j = 0; j < kTD; ++j) xh[j] = t1[j] = lp[j] = dl[j] = dy[j] = 0.f;
  if (valid) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kTD; ++j) {
      xh[j] = (a.t1pre[row * kTD + j] - s_mean[g][j]) * s_rstd[g][j];
      t1[j] = fmaxf(fmaf(s_gamma[j], xh[j], s_beta[j]), 0.f);
    }
#pragma unroll
    for (int o = 0; o < kTD; ++o) {
      float acc = s_b2[o];
#pragma unroll
      for (int j = 0; j < kTD; ++j) acc = fmaf(s_w2[o][j], t1[j], acc);
      lp[o] = acc;
      mx = fmaxf(mx, acc);
    }
    float se = 0.f;
#pragma unroll
    for (int o = 0; o < kTD; ++o) se += expf(lp[o] - mx);
    const float lse = mx + logf(se);
#pragma unroll
    for (int o = 0; o < kTD; ++o) lp[o] -= lse;  // log_softmax (mnist/model.py:170)
    if (a.logp != nullptr) {
#pragma unroll
      for (int o = 0; o < kTD; ++o) a.logp[row * kTD + o] = lp[o];
    }
    if (!a.backward && a.fused_loss && a.labels != nullptr && a.ce != nullptr) {
      // forward-only evaluation (mnist/loglikelihood.py:49-52): the NLL term without any gradient work
      const int label = static_cast<int>(a.labels[row % a.B]);
#pragma unroll
      for (int o = 0; o < kTD; ++o)
        if (o == label) ce = -a.ce_scale[g] * lp[o];
    }
    if (a.backward) {
      if (a.fused_loss) {
        const int label = static_cast<int>(a.labels[row % a.B]);
        const float sc = a.ce_scale[g];
#pragma unroll
        for (int o = 0; o < kTD; ++o) {
          dl[o] = sc * (expf(lp[o]) - (o == label ? 1.f : 0.f));
          if (o == label) ce = -sc * lp[o];  // F.nll_loss mean (mnist/train.py:73)
        }
      } else {
        float su = 0.f, up[kTD];
#pragma unroll
        for (int o = 0; o < kTD; ++o) {
          up[o] = a.dlogp_up != nullptr ? a.dlogp_up[row * kTD + o] : 0.f;
          su += up[o];
        }
#pragma unroll
        for (int o = 0; o < kTD; ++o) dl[o] = up[o] - expf(lp[o]) * su;  // log_softmax backward
      }
#pragma unroll
      for (int j = 0; j < kTD; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int o = 0; o < kTD; ++o) acc = fmaf(dl[o], s_w2[o][j], acc);
        dy[j] = t1[j] > 0.f ? acc : 0.f;
        a.dyhat[row * kTD + j] = dy[j];
      }
    }
  }
  if (a.backward) {
    // block-level reductions: warp shuffle, then one shared atomic per warp per value
#pragma unroll
    for (int o = 0; o < kTD; ++o) {
#pragma unroll
      for (int j = 0; j < kTD; ++j) {
        const float s = warp_sum(dl[o] * t1[j]);
        if (lane == 0) atomicAdd(&s_dw2[o][j], s);
      }
      const float sb = warp_sum(dl[o]);
      if (lane == 0) atomicAdd(&s_db2[o], sb);
    }
    // a warp may straddle a group boundary: reduce per group with a mask on the group id
    for (int gg = 0; gg < a.G; ++gg) {
      const bool mine = valid && g == gg;
      if (__any_sync(0xffffffffu, mine)) {
#pragma unroll
        for (int j = 0; j < kTD; ++j) {
          const float s0 = warp_sum(mine ? dy[j] : 0.f);
          const float s1 = warp_sum(mine ? dy[j] * xh[j] : 0.f);
          if (lane == 0) {
            atomicAdd(&s_s0[gg][j], s0);
            atomicAdd(&s_s1[gg][j], s1);
          }
        }
        const float c = warp_sum(mine ? ce : 0.f);
        if (lane == 0) atomicAdd(&s_ce[gg], c);
      }
    }
    __syncthreads();
    if (tid < kTD * kTD && a.d_w2 != nullptr) atomicAdd(a.d_w2 + tid, s_dw2[tid / kTD][tid % kTD]);
    if (tid < kTD && a.d_b2 != nullptr) atomicAdd(a.d_b2 + tid, s_db2[tid]);
    if (tid < a.G * kTD) {
      atomicAdd(a.s0 + tid, (&s_s0[0][0])[tid]);
      atomicAdd(a.s1 + tid, (&s_s1[0][0])[tid]);
    }
    if (tid < a.G && a.ce != nullptr && a.fused_loss) atomicAdd(a.ce + tid, s_ce[tid]);
  } else if (a.fused_loss && a.labels != nullptr && a.ce != nullptr) {
    for (int gg = 0; gg < a.G; ++gg) {
      const bool mine = valid && g == gg;
      if (__any_sync(0xffffffffu, mine)) {
        const float c = warp_sum(mine ? ce : 0.f);
        if (lane == 0) atomicAdd(&s_ce[gg], c);
      }
    }
    __syncthreads();
    if (tid < a.G) atomicAdd(a.ce + tid, s_ce[tid]);
  }
}

// ================================================================= text encoder, evaluated per LABEL
// Embedding(10,50) -> BatchNorm1d(50) -> ReLU -> Linear(50, 2n) has only ten distinct inputs, so the batch
// statistics are count-weighted sums over the ten embedding rows and the output is a [10, 2n] table the tail
// kernel gathers from.  Single block.
constexpr int kEmb = 50;
__global__ void __launch_bounds__(256) textenc_fwd_kernel(const TextEncArgs a) {
  __shared__ float s_cnt[kTD];
  __shared__ float s_h[kTD][kEmb];
  float* sv_cnt = a.save;
  float* sv_xh = a.save + kTD;
  float* sv_h = sv_xh + kTD * kEmb;
  float* sv_mean = sv_h + kTD * kEmb;
  float* sv_rstd = sv_mean + kEmb;
  const int tid = threadIdx.x;
  if (tid < kTD) s_cnt[tid] = 0.f;
  __syncthreads();
  {
    float local[kTD];
#pragma unroll
    for (int l = 0; l < kTD; ++l) local[l] = 0.f;
    for (int i = tid; i < a.B; i += blockDim.x) {
      const int lab = static_cast<int>(a.labels[i]);
#pragma unroll
      for (int l = 0; l < kTD; ++l) local[l] += (lab == l) ? 1.f : 0.f;
    }
#pragma unroll
    for (int l = 0; l < kTD; ++l) {
      const float s = warp_sum(local[l]);
      if ((tid & 31) == 0) atomicAdd(&s_cnt[l], s);
    }
  }
  __syncthreads();
  if (tid < kTD) sv_cnt[tid] = s_cnt[tid];
  if (tid < kEmb) {
    const int f = tid;
    float mean, rstd;
    if (a.training) {
      float s = 0.f;
      for (int l = 0; l < kTD; ++l) s += s_cnt[l] * a.emb[l * kEmb + f];
      mean = s / a.B;
      float v = 0.f;
      for (int l = 0; l < kTD; ++l) {
        const float d = a.emb[l * kEmb + f] - mean;
        v += s_cnt[l] * d * d;
      }
      const float var = v / a.B;
      rstd = rsqrtf(var + a.bn_eps);
      if (a.running_mean != nullptr) {
        const float unb = a.B > 1 ? var * (static_cast<float>(a.B) / (a.B - 1)) : var;
        float rm = a.running_mean[f], rv = a.running_var[f];
        for (int u = 0; u < a.updates; ++u) {
          rm = (1.f - a.momentum) * rm + a.momentum * mean;
          rv = (1.f - a.momentum) * rv + a.momentum * unb;
        }
        a.running_mean[f] = rm;
        a.running_var[f] = rv;
      }
    } else {
      mean = a.running_mean[f];
      rstd = rsqrtf(a.running_var[f] + a.bn_eps);
    }
    sv_mean[f] = mean;
    sv_rstd[f] = rstd;
    for (int l = 0; l < kTD; ++l) {
      const float xh = (a.emb[l * kEmb + f] - mean) * rstd;
      const float h = fmaxf(fmaf(a.gamma[f], xh, a.beta[f]), 0.f);
      sv_xh[l * kEmb + f] = xh;
      sv_h[l * kEmb + f] = h;
      s_h[l][f] = h;
    }
  }
  __syncthreads();
  const int two_n = 2 * a.n;
  for (int i = tid; i < kTD * two_n; i += blockDim.x) {
    const int l = i / two_n, o = i % two_n;
    float acc = a.b[o];
    for (int f = 0; f < kEmb; ++f) acc = fmaf(s_h[l][f], a.w[o * kEmb + f], acc);
    a.table[i] = acc;
  }
}

// Backward of the per-label text encoder.  Everything it touches (d_table [10][2n], W [2n][50], the saved activations) is
// staged in shared memory with coalesced loads first: the contractions then run out of shared memory instead of
// chains of dependent global loads (42 -> a few us; single block, off the critical path on a side stream).
constexpr int kTextMaxOut = 128;   // shared-memory staging covers 2n <= 128; wider latents read global memory
__global__ void __launch_bounds__(256) textenc_bwd_kernel(const TextEncArgs a) {
  __shared__ float s_dh[kTD][kEmb];
  __shared__ float s_dt[kTD][kTextMaxOut];
  __shared__ float s_w[kTextMaxOut][kEmb + 1];
  __shared__ float s_h[kTD][kEmb];
  const float* sv_cnt = a.save;
  const float* sv_xh = a.save + kTD;
  const float* sv_h = sv_xh + kTD * kEmb;
  const float* sv_rstd = sv_h + kTD * kEmb + kEmb;
  const int tid = threadIdx.x;
  const int two_n = 2 * a.n;
  const bool staged = two_n <= kTextMaxOut;
  if (staged) {
    for (int i = tid; i < kTD * two_n; i += blockDim.x) s_dt[i / two_n][i % two_n] = a.d_table[i];
    for (int i = tid; i < two_n * kEmb; i += blockDim.x) s_w[i / kEmb][i % kEmb] = a.w[i];
    for (int i = tid; i < kTD * kEmb; i += blockDim.x) s_h[i / kEmb][i % kEmb] = sv_h[i];
    __syncthreads();
  }
  // Linear(50 -> 2n): dW[o][f] = sum_l dT[l][o] h[l][f]; db[o] = sum_l dT[l][o]; dh[l][f] = sum_o dT[l][o] W[o][f]
  for (int i = tid; i < two_n * kEmb; i += blockDim.x) {
    const int o = i / kEmb, f = i % kEmb;
    float acc = 0.f;
    if (staged) {
#pragma unroll
      for (int l = 0; l < kTD; ++l) acc = fmaf(s_dt[l][o], s_h[l][f], acc);
    } else {
      for (int l = 0; l < kTD; ++l) acc = fmaf(a.d_table[l * two_n + o], sv_h[l * kEmb + f], acc);
    }
    atomicAdd(a.d_w + i, acc);   // one owner per address: a reduction without a load to wait for
  }
  for (int o = tid; o < two_n; o += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < kTD; ++l) acc += staged ? s_dt[l][o] : a.d_table[l * two_n + o];
    atomicAdd(a.d_b + o, acc);
  }
  for (int i = tid; i < kTD * kEmb; i += blockDim.x) {
    const int l = i / kEmb, f = i % kEmb;
    float acc = 0.f;
    if (staged) {
      for (int o = 0; o < two_n; ++o) acc = fmaf(s_dt[l][o], s_w[o][f], acc);
    } else {
      for (int o = 0; o < two_n; ++o) acc = fmaf(a.d_table[l * two_n + o], a.w[o * kEmb + f], acc);
    }
    s_dh[l][f] = sv_h[i] > 0.f ? acc : 0.f;  // ReLU mask; this is dyhat aggregated over the label's samples
  }
  __syncthreads();
  if (tid < kEmb) {
    const int f = tid;
    float s0 = 0.f, s1 = 0.f;
    for (int l = 0; l < kTD; ++l) {
      s0 += s_dh[l][f];
      s1 += s_dh[l][f] * sv_xh[l * kEmb + f];
    }
    atomicAdd(a.d_gamma + f, s1);
    atomicAdd(a.d_beta + f, s0);
    const float gr = a.gamma[f] * sv_rstd[f];
    for (int l = 0; l < kTD; ++l) {
      const float c = sv_cnt[l] / a.B;
      atomicAdd(a.d_emb + l * kEmb + f, gr * (s_dh[l][f] - c * s0 - c * sv_xh[l * kEmb + f] * s1));
    }
  }
}

// ================================================================= standalone masked PoE
__global__ void __launch_bounds__(256)
    poe_fwd_kernel(int mode, int prior, float eps, int M, long long B, int D, const float* __restrict__ mu,
                   const float* __restrict__ logvar, const float* __restrict__ mask, float* __restrict__ out_mu,
                   float* __restrict__ out_logvar) {
  const long long total = B * D;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / D;
    float num = 0.f, S = (mode == MVAE_POE_PRECISION && prior) ? 1.f : 0.f, P = 0.f;
    for (int e = 0; e < M; ++e) {
      const float w = mask != nullptr ? mask[e * B + b] : 1.f;
      if (w == 0.f) continue;
      const float var = expf(logvar[e * total + i]) + eps;
      if (mode == MVAE_POE_REF) {
        num += mu[e * total + i] * var;
        S += var;
        P += 1.f / var;
      } else {
        const float t = w / var;
        num += mu[e * total + i] * t;
        S += t;
      }
    }
    if (mode == MVAE_POE_REF) {
      out_mu[i] = num / S;
      const float pd_var = 1.f / P;
      out_logvar[i] = logf(pd_var);
    } else {
      out_mu[i] = num / S;
      out_logvar[i] = logf(1.f / S);
    }
  }
}
__global__ void __launch_bounds__(256)
    poe_bwd_kernel(int mode, int prior, float eps, int M, long long B, int D, const float* __restrict__ mu,
                   const float* __restrict__ logvar, const float* __restrict__ mask, const float* __restrict__ dom,
                   const float* __restrict__ dol, float* __restrict__ dmu, float* __restrict__ dlv) {
  const long long total = B * D;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / D;
    float num = 0.f, S = (mode == MVAE_POE_PRECISION && prior) ? 1.f : 0.f, P = 0.f;
    for (int e = 0; e < M; ++e) {
      const float w = mask != nullptr ? mask[e * B + b] : 1.f;
      if (w == 0.f) continue;
      const float var = expf(logvar[e * total + i]) + eps;
      if (mode == MVAE_POE_REF) {
        num += mu[e * total + i] * var;
        S += var;
        P += 1.f / var;
      } else {
        const float t = w / var;
        num += mu[e * total + i] * t;
        S += t;
      }
    }
    const float pm = num / S;
    const float g_mu = dom != nullptr ? dom[i] : 0.f;
    const float g_lv = dol != nullptr ? dol[i] : 0.f;
    for (int e = 0; e < M; ++e) {
      const float w = mask != nullptr ? mask[e * B + b] : 1.f;
      float o_m = 0.f, o_l = 0.f;
      if (w != 0.f) {
        const float ex = expf(logvar[e * total + i]);
        const float var = ex + eps;
        if (mode == MVAE_POE_REF) {
          o_m = g_mu * var / S;
          const float dvar = g_mu * (mu[e * total + i] - pm) / S + g_lv / (P * var * var);
          o_l = dvar * ex;
        } else {
          const float t = w / var;
          o_m = g_mu * t / S;
          const float dT = g_mu * (mu[e * total + i] - pm) / S - g_lv / S;
          o_l = -dT * (w / (var * var)) * ex;
        }
      }
      dmu[e * total + i] = o_m;
      dlv[e * total + i] = o_l;
    }
  }
}

int check_tail(const TailArgs& a) {
  MVAE_REQUIRE(a.B > 0 && a.n > 0 && a.n % 2 == 0, "tail: B=%d n=%d (n must be even)", a.B, a.n);
  MVAE_REQUIRE(a.G >= 1 && a.G <= kMaxGroups, "tail: G=%d out of range", a.G);
  for (int g = 0; g < a.G; ++g) {
    const int t = a.group_type[g];
    MVAE_REQUIRE(t >= 0 && t <= 2, "tail: bad term type %d", t);
    if (a.z_in != nullptr) continue;
    MVAE_REQUIRE(t == TERM_TEXT || a.enc_img != nullptr, "tail: term %d needs the image expert", g);
    MVAE_REQUIRE(t == TERM_IMAGE || (a.txt_table != nullptr && a.labels != nullptr), "tail: term %d needs the text expert", g);
  }
  return 0;
}

}  // namespace

int launch_tail_forward(const TailArgs& a, cudaStream_t st) {
  if (check_tail(a)) return 1;
  MVAE_REQUIRE(a.z != nullptr, "tail_forward: z output missing");
  MVAE_REQUIRE(a.wt1 == nullptr || (a.bt1 && a.t1pre && a.t1_sum && a.t1_sumsq), "tail_forward: text decoder buffers missing");
  static const int use_fast = env_int("MVAE_TAIL_FAST", 1);
  if (use_fast && a.n == kT3N && a.training && a.z_in == nullptr && a.wt1 != nullptr && a.enc_img != nullptr &&
      a.txt_table != nullptr && a.labels != nullptr && (a.eps != nullptr || a.noise_buf != nullptr)) {
    // (without injected draws this kernel uses its own Philox stream - normal_quad - and always leaves the draws in noise_buf:
    //  every backward kernel reads them from there)
    const int blocks3 = std::max(1, std::min((a.B + kT3Warps - 1) / kT3Warps, 148 * 2));
    const bool ref = a.poe_mode == MVAE_POE_REF;
    if (a.z_dtype == MVAE_F32)
      return ref ? launch_kernel(tail_fwd3_kernel<float, MVAE_POE_REF>, dim3(blocks3), dim3(32 * kT3Warps), 0, st, a)
                 : launch_kernel(tail_fwd3_kernel<float, MVAE_POE_PRECISION>, dim3(blocks3), dim3(32 * kT3Warps), 0, st, a);
    return ref ? launch_kernel(tail_fwd3_kernel<__nv_bfloat16, MVAE_POE_REF>, dim3(blocks3), dim3(32 * kT3Warps), 0, st, a)
               : launch_kernel(tail_fwd3_kernel<__nv_bfloat16, MVAE_POE_PRECISION>, dim3(blocks3), dim3(32 * kT3Warps), 0, st, a);
  }
  if (a.n <= 64) {
    // one warp per sample (all terms): every warp resident at once for the batch sizes of the MNIST configurations
    static const int wpb = env_int("MVAE_TAIL_WARPS", 14);
    if (wpb == 14) {
      const int blocks2 = std::max(1, std::min((a.B + 13) / 14, 148 * 2));
      if (a.z_dtype == MVAE_F32) return launch_kernel(tail_fwd2_kernel<float, 14>, dim3(blocks2), dim3(32 * 14), 0, st, a);
      return launch_kernel(tail_fwd2_kernel<__nv_bfloat16, 14>, dim3(blocks2), dim3(32 * 14), 0, st, a);
    }
    const int blocks2 = std::max(1, std::min((a.B + 3) / 4, 148 * 7));
    if (a.z_dtype == MVAE_F32) return launch_kernel(tail_fwd2_kernel<float, 4>, dim3(blocks2), dim3(32 * 4), 0, st, a);
    return launch_kernel(tail_fwd2_kernel<__nv_bfloat16, 4>, dim3(blocks2), dim3(32 * 4), 0, st, a);
  }
  const long long items = static_cast<long long>(a.G) * a.B;
  int blocks = static_cast<int>(std::min<long long>((items + kTailWarps - 1) / kTailWarps, 148 * 8));
  if (blocks < 1) blocks = 1;
  if (a.z_dtype == MVAE_F32) return launch_kernel(tail_fwd_kernel<float>, dim3(blocks), dim3(kTailThreads), 0, st, a);
  return launch_kernel(tail_fwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(kTailThreads), 0, st, a);
}

int launch_tail_backward(const TailArgs& a, cudaStream_t st) {
  if (check_tail(a)) return 1;
  static const int use_fast = env_int("MVAE_TAIL_FAST", 1);
  if (use_fast && a.n == kT3N && a.training && a.z_in == nullptr && a.dz != nullptr && a.dmu_up == nullptr && a.dlogvar_up == nullptr &&
      a.t1_dyhat != nullptr && a.enc_img != nullptr && a.txt_table != nullptr && a.labels != nullptr && a.d_enc != nullptr &&
      a.d_txt_table != nullptr && a.d_wt1 != nullptr && (a.eps != nullptr || a.noise_buf != nullptr)) {
    const int smem3 = (kT3Warps * kTD * 2 * kT3N + kTD * kT3N) * 4;
    static bool attr3 = false;
    if (!attr3) {
      MVAE_CUDA(cudaFuncSetAttribute(tail_bwd3_kernel<float, MVAE_POE_REF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
      MVAE_CUDA(cudaFuncSetAttribute(tail_bwd3_kernel<float, MVAE_POE_PRECISION>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
      MVAE_CUDA(cudaFuncSetAttribute(tail_bwd3_kernel<__nv_bfloat16, MVAE_POE_REF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
      MVAE_CUDA(cudaFuncSetAttribute(tail_bwd3_kernel<__nv_bfloat16, MVAE_POE_PRECISION>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
      attr3 = true;
    }
    const int blocks3 = std::max(1, std::min((a.B + kT3Warps - 1) / kT3Warps, 148 * 2));
    const bool ref = a.poe_mode == MVAE_POE_REF;
    if (a.z_dtype == MVAE_F32)
      return ref ? launch_kernel(tail_bwd3_kernel<float, MVAE_POE_REF>, dim3(blocks3), dim3(32 * kT3Warps), smem3, st, a)
                 : launch_kernel(tail_bwd3_kernel<float, MVAE_POE_PRECISION>, dim3(blocks3), dim3(32 * kT3Warps), smem3, st, a);
    return ref ? launch_kernel(tail_bwd3_kernel<__nv_bfloat16, MVAE_POE_REF>, dim3(blocks3), dim3(32 * kT3Warps), smem3, st, a)
               : launch_kernel(tail_bwd3_kernel<__nv_bfloat16, MVAE_POE_PRECISION>, dim3(blocks3), dim3(32 * kT3Warps), smem3, st, a);
  }
  if (a.n <= 64 && a.n % 2 == 0) {
    // one warp per sample (all terms): the fast path for the latent sizes the MNIST configurations use
    const int smem2 = (kTD * 2 * a.n + kTD * a.n + kMaxGroups * 4 * kTD + kTD * a.n + 2 * a.n) * 4;
    static const int wpb = env_int("MVAE_TAIL_WARPS", 14);
    if (wpb == 14) {
      const int blocks2 = std::max(1, std::min((a.B + 13) / 14, 148 * 2));
      if (a.z_dtype == MVAE_F32) return launch_kernel(tail_bwd2_kernel<float, 14>, dim3(blocks2), dim3(32 * 14), smem2, st, a);
      return launch_kernel(tail_bwd2_kernel<__nv_bfloat16, 14>, dim3(blocks2), dim3(32 * 14), smem2, st, a);
    }
    const int blocks2 = std::max(1, std::min((a.B + 3) / 4, 148 * 7));
    if (a.z_dtype == MVAE_F32) return launch_kernel(tail_bwd2_kernel<float, 4>, dim3(blocks2), dim3(32 * 4), smem2, st, a);
    return launch_kernel(tail_bwd2_kernel<__nv_bfloat16, 4>, dim3(blocks2), dim3(32 * 4), smem2, st, a);
  }
  const int smem_floats = kTD * 2 * a.n + kTD * a.n + 2 * a.n + kMaxGroups * 4 * kTD +
                          2 * kBwdRows * kMaxGroups * 2 * a.n + kBwdRows;
  const int smem = smem_floats * 4;
  MVAE_REQUIRE(smem <= 48 * 1024, "tail_backward: n=%d too large for the shared accumulators", a.n);
  const int row_blocks = (a.B + kBwdRows - 1) / kBwdRows;
  int blocks = std::min(row_blocks, 148 * 2);
  const int threads = 32 * a.G * kBwdRows;
  if (a.z_dtype == MVAE_F32)
    return launch_kernel(tail_bwd_kernel<float>, dim3(blocks), dim3(threads), smem, st, a, smem_floats);
  return launch_kernel(tail_bwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(threads), smem, st, a, smem_floats);
}

int launch_textdec(const TextDecArgs& a, cudaStream_t st) {
  MVAE_REQUIRE(a.B > 0 && a.G >= 1 && a.G <= kMaxGroups, "textdec: B=%d G=%d", a.B, a.G);
  MVAE_REQUIRE(a.t1pre && a.gamma && a.beta && a.w2 && a.b2, "textdec: missing inputs");
  MVAE_REQUIRE(!a.training || (a.t1_sum && a.t1_sumsq), "textdec: batch statistics missing");
  MVAE_REQUIRE(a.training || (a.running_mean && a.running_var), "textdec: running statistics missing");
  if (a.backward) {
    MVAE_REQUIRE(a.dyhat && a.s0 && a.s1, "textdec: backward outputs missing");
    MVAE_REQUIRE(!a.fused_loss || a.labels != nullptr, "textdec: labels missing");
  }
  const long long rows = static_cast<long long>(a.G) * a.B;
  const int blocks = static_cast<int>((rows + 255) / 256);
  return launch_kernel(textdec_kernel, dim3(blocks), dim3(256), 0, st, a);
}

int launch_textenc_forward(const TextEncArgs& a, cudaStream_t st) {
  MVAE_REQUIRE(a.B > 0 && a.n > 0 && a.labels && a.emb && a.gamma && a.beta && a.w && a.b && a.table && a.save,
               "textenc_forward: missing arguments");
  return launch_kernel(textenc_fwd_kernel, dim3(1), dim3(256), 0, st, a);
}

int launch_textenc_backward(const TextEncArgs& a, cudaStream_t st) {
  MVAE_REQUIRE(a.d_table && a.d_emb && a.d_gamma && a.d_beta && a.d_w && a.d_b && a.save && a.w && a.gamma,
               "textenc_backward: missing arguments");
  return launch_kernel(textenc_bwd_kernel, dim3(1), dim3(256), 0, st, a);
}

int launch_poe_forward(int mode, int prior, float eps, int M, long long B, int D, const float* mu, const float* logvar,
                       const float* mask, float* out_mu, float* out_logvar, cudaStream_t st) {
  MVAE_REQUIRE(M >= 1 && B > 0 && D > 0 && mu && logvar && out_mu && out_logvar, "poe_forward: bad arguments");
  const long long total = B * D;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  poe_fwd_kernel<<<blocks, 256, 0, st>>>(mode, prior, eps, M, B, D, mu, logvar, mask, out_mu, out_logvar);
  MVAE_CUDA(cudaGetLastError());
  return 0;
}

int launch_poe_backward(int mode, int prior, float eps, int M, long long B, int D, const float* mu, const float* logvar,
                        const float* mask, const float* d_out_mu, const float* d_out_logvar, float* d_mu,
                        float* d_logvar, cudaStream_t st) {
  MVAE_REQUIRE(M >= 1 && B > 0 && D > 0 && mu && logvar && d_mu && d_logvar, "poe_backward: bad arguments");
  const long long total = B * D;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  poe_bwd_kernel<<<blocks, 256, 0, st>>>(mode, prior, eps, M, B, D, mu, logvar, mask, d_out_mu, d_out_logvar, d_mu,
                                         d_logvar);
  MVAE_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mvae
