// Memory-bound elementwise kernels of the MVAE step: BatchNorm(+ReLU) apply forward / backward
// (the batch statistics themselves come out of the GEMM epilogues), dtype casts, fused Adam.
// All are vectorised (16-byte accesses), coalesced, grid-stride with the grid sized to the SM count.
//
// Reference semantics: nn.BatchNorm1d in train mode + nn.ReLU (mnist/model.py:105-106 etc.),
// torch.optim.Adam defaults (mnist/train.py:118,153).
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace mvae {

namespace {

constexpr int kEwThreads = 256;

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T>
__device__ __forceinline__ void ld4(const T* p, float (&o)[4]) {
  if constexpr (sizeof(T) == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  } else {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
  }
}
template <typename T>
__device__ __forceinline__ void st4(T* p, const float (&v)[4]) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    const __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<const uint32_t*>(&a);
    t.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
}

// 16-byte vectors: 4 fp32 or 8 bf16 features per thread.
template <typename T>
struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
};
template <typename T>
__device__ __forceinline__ void ldv(const T* p, float (&o)[Vec16<T>::N]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  if constexpr (sizeof(T) == 4) {
    o[0] = __uint_as_float(t.x); o[1] = __uint_as_float(t.y); o[2] = __uint_as_float(t.z); o[3] = __uint_as_float(t.w);
  } else {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      o[2 * i] = __low2float(h);
      o[2 * i + 1] = __high2float(h);
    }
  }
}
template <typename T>
__device__ __forceinline__ void stv(T* p, const float (&v)[Vec16<T>::N]) {
  uint4 t;
  if constexpr (sizeof(T) == 4) {
    t = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    t = make_uint4(w[0], w[1], w[2], w[3]);
  }
  *reinterpret_cast<uint4*>(p) = t;
}

// N consecutive fp32 coefficients with 16-byte loads (f is a multiple of N, buffers are 16-byte aligned)
template <int N>
__device__ __forceinline__ void ldc(const float* p, float (&o)[N]) {
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + i);
    o[i] = t.x; o[i + 1] = t.y; o[i + 2] = t.z; o[i + 3] = t.w;
  }
}

constexpr int kBnRowsPerBlock = kEwThreads / 32;  // one warp = 32 consecutive 16-byte vectors of one row
constexpr int kBnUnroll = 4;                      // independent 16-byte loads in flight per thread

// ---------------------------------------------------------------- BatchNorm forward apply
// y = relu(gamma * (x - mean_g) * rstd_g + beta), mean/var from the per-group column sums that the producing
// GEMM's epilogue accumulated.  A warp covers 512 contiguous bytes of a row; each thread keeps kBnUnroll rows
// in flight.  Block row 0 also finalises: saves mean/rstd per group and folds each group's statistics into the
// running buffers in group order (the reference runs one forward per group).
template <typename T>
__global__ void __launch_bounds__(kEwThreads)
    bn_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int rows, int F, int rows_per_group, int groups,
                  const float* __restrict__ sum, const float* __restrict__ sumsq, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ save_mean, float* __restrict__ save_rstd,
                  float* __restrict__ running_mean, float* __restrict__ running_var, int updates_per_group,
                  float momentum, float eps, int relu) {
  constexpr int N = Vec16<T>::N;
  const int lane = threadIdx.x & 31, tr = threadIdx.x >> 5;
  const int f = (blockIdx.x * 32 + lane) * N;
  if (f >= F) return;
  float ga[N], be[N];
  ldc<N>(gamma + f, ga);
  ldc<N>(beta + f, be);
  const int slab = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * slab;
  const int r1 = min(rows, r0 + slab);
  for (int g = 0; g < groups; ++g) {
    const int gr0 = g * rows_per_group;
    const int gr1 = min(rows, gr0 + rows_per_group);
    const int a = max(r0, gr0), b = min(r1, gr1);
    const bool finalise = blockIdx.y == 0 && tr == 0;
    if (a >= b && !finalise) continue;
    const int cnt = gr1 - gr0;
    const float inv_cnt = 1.f / cnt;
    float mean[N], rstd[N], s_[N], ss_[N];
    if (sum == nullptr) {
      // eval mode (vae.eval()): normalise with the running statistics, update nothing
      ldc<N>(running_mean + f, mean);
      ldc<N>(running_var + f, ss_);
#pragma unroll
      for (int i = 0; i < N; ++i) rstd[i] = rsqrtf(ss_[i] + eps);
    } else {
      ldc<N>(sum + g * F + f, s_);
      ldc<N>(sumsq + g * F + f, ss_);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (sum == nullptr) break;
      const float s = s_[i], ss = ss_[i];
      mean[i] = s * inv_cnt;
      const float var = fmaxf(ss * inv_cnt - mean[i] * mean[i], 0.f);
      rstd[i] = rsqrtf(var + eps);
      if (finalise) {
        if (save_mean != nullptr) {
          save_mean[g * F + f + i] = mean[i];
          save_rstd[g * F + f + i] = rstd[i];
        }
        if (running_mean != nullptr) {
          const float unb = cnt > 1 ? var * (static_cast<float>(cnt) / (cnt - 1)) : var;
          float rm = running_mean[f + i], rv = running_var[f + i];
          for (int u = 0; u < updates_per_group; ++u) {
            rm = (1.f - momentum) * rm + momentum * mean[i];
            rv = (1.f - momentum) * rv + momentum * unb;
          }
          running_mean[f + i] = rm;
          running_var[f + i] = rv;
        }
      }
    }
    for (int r = a + tr; r < b; r += kBnRowsPerBlock * kBnUnroll) {
      float v[kBnUnroll][N];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int rr = r + u * kBnRowsPerBlock;
        if (rr < b) ldv(x + static_cast<long long>(rr) * F + f, v[u]);
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int rr = r + u * kBnRowsPerBlock;
        if (rr < b) {
          float o[N];
#pragma unroll
          for (int i = 0; i < N; ++i) {
            // folded form a*x + b, a = gamma*rstd, b = beta - mean*a: the SAME expression as the GEMM A-transform
            // and the dgrad epilogue's mask recomputation, so all three agree bit for bit
            const float a_ = ga[i] * rstd[i];
            const float t = fmaf(a_, v[u][i], fmaf(-mean[i], a_, be[i]));
            o[i] = relu ? fmaxf(t, 0.f) : t;
          }
          stv(y + static_cast<long long>(rr) * F + f, o);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- BatchNorm backward apply
// dx = gamma * rstd * (dyhat - S0/cnt - xhat * S1/cnt), S0 = sum dyhat, S1 = sum dyhat*xhat per group
// (both produced by the dgrad GEMM epilogue).  Block row 0 also emits dgamma += sum_g S1, dbeta += sum_g S0.
template <typename T>
__global__ void __launch_bounds__(kEwThreads)
    bn_bwd_kernel(const T* __restrict__ dyhat, const T* __restrict__ x, T* __restrict__ dx, int rows, int F,
                  int rows_per_group, int groups, const float* __restrict__ s0, const float* __restrict__ s1,
                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                  float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int N = Vec16<T>::N;
  const int lane = threadIdx.x & 31, tr = threadIdx.x >> 5;
  const int f = (blockIdx.x * 32 + lane) * N;
  if (f >= F) return;
  const int slab = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * slab;
  const int r1 = min(rows, r0 + slab);
  const bool finalise = blockIdx.y == 0 && tr == 0 && dgamma != nullptr;
  float dg[N], db[N];
#pragma unroll
  for (int i = 0; i < N; ++i) dg[i] = db[i] = 0.f;
  for (int g = 0; g < groups; ++g) {
    const int gr0 = g * rows_per_group;
    const int gr1 = min(rows, gr0 + rows_per_group);
    const int a = max(r0, gr0), b = min(r1, gr1);
    if (a >= b && !finalise) continue;
    const float inv = 1.f / (gr1 - gr0);
    float mu[N], sc[N], c0[N], c1[N], rs_[N], a0_[N], a1_[N], ga_[N];
    ldc<N>(mean + g * F + f, mu);
    ldc<N>(rstd + g * F + f, rs_);
    ldc<N>(s0 + g * F + f, a0_);
    ldc<N>(s1 + g * F + f, a1_);
    ldc<N>(gamma + f, ga_);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float rs = rs_[i];
      const float a0 = a0_[i], a1 = a1_[i];
      c0[i] = a0 * inv;
      c1[i] = a1 * inv * rs;           // multiplies (x - mean): xhat * S1/cnt = (x-mean) * rs * S1/cnt
      sc[i] = ga_[i] * rs;
      dg[i] += a1;
      db[i] += a0;
    }
    for (int r = a + tr; r < b; r += kBnRowsPerBlock * kBnUnroll) {
      float d[kBnUnroll][N], v[kBnUnroll][N];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int rr = r + u * kBnRowsPerBlock;
        if (rr < b) {
          ldv(dyhat + static_cast<long long>(rr) * F + f, d[u]);
          ldv(x + static_cast<long long>(rr) * F + f, v[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int rr = r + u * kBnRowsPerBlock;
        if (rr < b) {
          float o[N];
#pragma unroll
          for (int i = 0; i < N; ++i) o[i] = sc[i] * (d[u][i] - c0[i] - (v[u][i] - mu[i]) * c1[i]);
          stv(dx + static_cast<long long>(rr) * F + f, o);
        }
      }
    }
  }
  if (finalise) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      dgamma[f + i] += dg[i];
      dbeta[f + i] += db[i];
    }
  }
}

// ---------------------------------------------------------------- casts
__global__ void __launch_bounds__(kEwThreads) cast_f32_bf16_kernel(const float* __restrict__ in,
                                                                    __nv_bfloat16* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v[4];
    ld4(in + 4 * i, v);
    st4(out + 4 * i, v);
  }
}
__global__ void __launch_bounds__(kEwThreads) u8_to_act_kernel(const uint8_t* __restrict__ in, float* __restrict__ o32,
                                                                __nv_bfloat16* __restrict__ o16, long long n4,
                                                                float scale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uchar4 u = reinterpret_cast<const uchar4*>(in)[i];
    const float v[4] = {u.x * scale, u.y * scale, u.z * scale, u.w * scale};
    if (o32 != nullptr) st4(o32 + 4 * i, v);
    if (o16 != nullptr) st4(o16 + 4 * i, v);
  }
}

// ---------------------------------------------------------------- fused Adam over the flat parameter buffer
// p, g, m, v are the whole model in one fp32 buffer each (n multiple of 4).  Also refreshes the bf16 mirror of
// the parameters used by the bf16 tensor-core path and (optionally) zeroes the gradient for the next step.
__global__ void __launch_bounds__(kEwThreads)
    adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                __nv_bfloat16* __restrict__ p16, long long n4, float lr, float b1, float b2, float eps,
                const int* __restrict__ step_ptr, float grad_scale, int zero_grad) {
  // The step count lives on the device so that a captured CUDA graph stays valid from step to step.
  const float step = static_cast<float>(*step_ptr);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float pp[4], gg[4], mm[4], vv[4];
    ld4(p + 4 * i, pp);
    ld4(g + 4 * i, gg);
    ld4(m + 4 * i, mm);
    ld4(v + 4 * i, vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = gg[k] * grad_scale;
      mm[k] = b1 * mm[k] + (1.f - b1) * gr;
      vv[k] = b2 * vv[k] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
      pp[k] -= (lr / bc1) * (mm[k] / denom);
    }
    st4(p + 4 * i, pp);
    st4(m + 4 * i, mm);
    st4(v + 4 * i, vv);
    if (p16 != nullptr) st4(p16 + 4 * i, pp);
    if (zero_grad) {
      const float z[4] = {0.f, 0.f, 0.f, 0.f};
      st4(g + 4 * i, z);
    }
  }
}

dim3 bn_grid(int rows, int F, int vec) {
  const int gx = (F / vec + 31) / 32;
  int gy = (4 * sm_count() + gx - 1) / gx;  // ~4 blocks per SM
  const int max_gy = (rows + kBnRowsPerBlock - 1) / kBnRowsPerBlock;
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  return dim3(gx, gy, 1);
}

}  // namespace

int launch_bn_forward(int dtype, const void* x, void* y, int rows, int F, int rows_per_group, const float* sum,
                      const float* sumsq, const float* gamma, const float* beta, float* save_mean, float* save_rstd,
                      float* running_mean, float* running_var, int updates_per_group, float momentum, float eps,
                      int relu, cudaStream_t st) {
  const int vec = dtype == MVAE_F32 ? 4 : 8;
  MVAE_REQUIRE(F % vec == 0, "bn_forward: feature count %d must be a multiple of %d", F, vec);
  MVAE_REQUIRE(rows > 0 && rows_per_group > 0, "bn_forward: empty input");
  const int groups = (rows + rows_per_group - 1) / rows_per_group;
  const dim3 grid = bn_grid(rows, F, vec);
  if (dtype == MVAE_F32)
    return launch_kernel(bn_fwd_kernel<float>, grid, dim3(kEwThreads), 0, st, static_cast<const float*>(x),
                      static_cast<float*>(y), rows, F, rows_per_group, groups, sum, sumsq, gamma, beta, save_mean,
                      save_rstd, running_mean, running_var, updates_per_group, momentum, eps, relu);
  return launch_kernel(bn_fwd_kernel<__nv_bfloat16>, grid, dim3(kEwThreads), 0, st, static_cast<const __nv_bfloat16*>(x),
                    static_cast<__nv_bfloat16*>(y), rows, F, rows_per_group, groups, sum, sumsq, gamma, beta, save_mean,
                    save_rstd, running_mean, running_var, updates_per_group, momentum, eps, relu);
}

int launch_bn_backward(int dtype, const void* dyhat, const void* x, void* dx, int rows, int F, int rows_per_group,
                       const float* s0, const float* s1, const float* mean, const float* rstd, const float* gamma,
                       float* dgamma, float* dbeta, cudaStream_t st) {
  const int vec = dtype == MVAE_F32 ? 4 : 8;
  MVAE_REQUIRE(F % vec == 0, "bn_backward: feature count %d must be a multiple of %d", F, vec);
  MVAE_REQUIRE(rows > 0 && rows_per_group > 0, "bn_backward: empty input");
  const int groups = (rows + rows_per_group - 1) / rows_per_group;
  const dim3 grid = bn_grid(rows, F, vec);
  if (dtype == MVAE_F32)
    return launch_kernel(bn_bwd_kernel<float>, grid, dim3(kEwThreads), 0, st, static_cast<const float*>(dyhat),
                      static_cast<const float*>(x), static_cast<float*>(dx), rows, F, rows_per_group, groups, s0, s1, mean,
                      rstd, gamma, dgamma, dbeta);
  return launch_kernel(bn_bwd_kernel<__nv_bfloat16>, grid, dim3(kEwThreads), 0, st,
                    static_cast<const __nv_bfloat16*>(dyhat), static_cast<const __nv_bfloat16*>(x),
                    static_cast<__nv_bfloat16*>(dx), rows, F, rows_per_group, groups, s0, s1, mean, rstd, gamma, dgamma,
                    dbeta);
}

int launch_cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t st) {
  MVAE_REQUIRE(n % 4 == 0, "cast: element count %lld must be a multiple of 4", n);
  if (n == 0) return 0;
  const long long n4 = n / 4;
  const int blocks = static_cast<int>(std::min<long long>((n4 + kEwThreads - 1) / kEwThreads, 4ll * sm_count()));
  cast_f32_bf16_kernel<<<blocks, kEwThreads, 0, st>>>(in, static_cast<__nv_bfloat16*>(out), n4);
  MVAE_CUDA(cudaGetLastError());
  return 0;
}

int launch_u8_to_act(const uint8_t* in, float* o32, void* o16, long long n, float scale, cudaStream_t st) {
  MVAE_REQUIRE(n % 4 == 0, "u8_to_act: element count %lld must be a multiple of 4", n);
  if (n == 0) return 0;
  const long long n4 = n / 4;
  const int blocks = static_cast<int>(std::min<long long>((n4 + kEwThreads - 1) / kEwThreads, 4ll * sm_count()));
  u8_to_act_kernel<<<blocks, kEwThreads, 0, st>>>(in, o32, static_cast<__nv_bfloat16*>(o16), n4, scale);
  MVAE_CUDA(cudaGetLastError());
  return 0;
}

int launch_adam(float* p, float* g, float* m, float* v, void* p16, long long n, float lr, float b1, float b2, float eps,
                const int* step_ptr, float grad_scale, int zero_grad, cudaStream_t st) {
  MVAE_REQUIRE(n % 4 == 0 && step_ptr != nullptr, "adam: n=%lld must be a multiple of 4 and step_ptr non-null", n);
  const long long n4 = n / 4;
  const int blocks = static_cast<int>(std::min<long long>((n4 + kEwThreads - 1) / kEwThreads, 4ll * sm_count()));
  return launch_kernel(adam_kernel, dim3(blocks), dim3(kEwThreads), 0, st, p, g, m, v, static_cast<__nv_bfloat16*>(p16), n4,
                    lr, b1, b2, eps, step_ptr, grad_scale, zero_grad);
}

// Start of a step: bump the device-side step counter and clear the accumulators (statistics, loss partials,
// and - when the caller asks - the flat gradient buffer) that the step's kernels add into with atomics.
struct NbtInc {
  long long v[6];
};
__global__ void __launch_bounds__(kEwThreads)
    step_prep_kernel(int* step_ptr, int* step_ptr2, float4* zero_buf, long long n4, long long* nbt, NbtInc inc) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && step_ptr != nullptr) *step_ptr += 1;
  if (blockIdx.x == 0 && threadIdx.x == 0 && step_ptr2 != nullptr && step_ptr2 != step_ptr) *step_ptr2 += 1;
  if (blockIdx.x == 0 && threadIdx.x < 6 && nbt != nullptr) nbt[threadIdx.x] += inc.v[threadIdx.x];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    zero_buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

int launch_step_prep(int* step_ptr, int* step_ptr2, float* zero_buf, long long zero_n, long long* nbt, const long long (&inc)[6],
                     cudaStream_t st) {
  MVAE_REQUIRE(zero_n % 4 == 0, "step_prep: zero_n=%lld must be a multiple of 4", zero_n);
  const long long n4 = zero_n / 4;
  int blocks = static_cast<int>(std::min<long long>((n4 + kEwThreads - 1) / kEwThreads, 2ll * sm_count()));
  if (blocks < 1) blocks = 1;
  NbtInc i;
  for (int k = 0; k < 6; ++k) i.v[k] = inc[k];
  return launch_kernel(step_prep_kernel, dim3(blocks), dim3(kEwThreads), 0, st, step_ptr, step_ptr2, reinterpret_cast<float4*>(zero_buf),
                    n4, nbt, i);
}

// losses [3][kMaxGroups] (bce, ce, kl) -> out [G][4] (total, bce, ce, kl)
__global__ void loss_pack_kernel(const float* __restrict__ acc, float* __restrict__ out, int G) {
  const int g = threadIdx.x;
  if (g < G) {
    const float b = acc[g], c = acc[kMaxGroups + g], k = acc[2 * kMaxGroups + g];
    out[g * 4 + 0] = b + c + k;
    out[g * 4 + 1] = b;
    out[g * 4 + 2] = c;
    out[g * 4 + 3] = k;
  }
}
int launch_loss_pack(const float* acc, float* out, int G, cudaStream_t st) {
  return launch_kernel(loss_pack_kernel, dim3(1), dim3(32), 0, st, acc, out, G);
}

// ---------------------------------------------------------------- sigmoid backward (module / autograd path)
// dlogit = d(prob) * p * (1 - p) for recon_image = sigmoid(logits) (mnist/model.py:135); also the last Linear's
// bias gradient (column sums).  One thread owns 4 columns and a slab of rows.
template <typename T>
__global__ void __launch_bounds__(kEwThreads)
    sigmoid_bwd_kernel(const T* __restrict__ dp, const T* __restrict__ p, T* __restrict__ dl, int rows, int F,
                       float* __restrict__ dbias) {
  const int q = blockIdx.x * 64 + (threadIdx.x & 63);
  const int tr = threadIdx.x >> 6;
  if (q * 4 >= F) return;
  const int f = q * 4;
  const int slab = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * slab, r1 = min(rows, r0 + slab);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = r0 + tr; r < r1; r += kEwThreads / 64) {
    float a[4], b[4], o[4];
    ld4(dp + static_cast<long long>(r) * F + f, a);
    ld4(p + static_cast<long long>(r) * F + f, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i] = a[i] * b[i] * (1.f - b[i]);
      acc[i] += o[i];
    }
    st4(dl + static_cast<long long>(r) * F + f, o);
  }
  if (dbias != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(dbias + f + i, acc[i]);
  }
}

int launch_sigmoid_backward(int dtype, const void* dprob, const void* prob, void* dlogit, int rows, int F, float* dbias,
                            cudaStream_t st) {
  MVAE_REQUIRE(F % 4 == 0 && rows > 0, "sigmoid_backward: rows=%d F=%d", rows, F);
  const int gx = (F / 4 + 63) / 64;
  int gy = (2 * sm_count() + gx - 1) / gx;
  if (gy > (rows + 3) / 4) gy = (rows + 3) / 4;
  if (gy < 1) gy = 1;
  const dim3 grid(gx, gy, 1);
  if (dtype == MVAE_F32)
    sigmoid_bwd_kernel<float><<<grid, kEwThreads, 0, st>>>(static_cast<const float*>(dprob), static_cast<const float*>(prob),
                                                           static_cast<float*>(dlogit), rows, F, dbias);
  else
    sigmoid_bwd_kernel<__nv_bfloat16><<<grid, kEwThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dprob), static_cast<const __nv_bfloat16*>(prob),
        static_cast<__nv_bfloat16*>(dlogit), rows, F, dbias);
  MVAE_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mvae
