// Operator-level kernels of the convolutional MVAEs (CelebA celeba/model.py:91-200, MultiMNIST
// multimnist/model.py:150-266).  Convolutions run as GEMMs on the tcgen05 path (gemm.cu); this file holds what
// surrounds them:
//
//   im2col / col2im      patch gather and its adjoint over NHWC activations, K ordered (kh, kw, c) so that both
//                        sides move 16-byte channel vectors (conv weights are kept in [Cout, kh, kw, Cin] /
//                        [Cin, kh, kw, Cout] order inside the library; the host permutes at the state_dict boundary);
//                        a scalar strided variant reads / writes the user's NCHW fp32 images directly
//   column reductions    BatchNorm batch statistics, BatchNorm-backward sums, bias gradients
//   BatchNorm + act      nn.BatchNorm1d/2d (train or eval) fused with Swish / ReLU, forward and backward
//   act (+ dropout)      Swish / ReLU after a Linear, optional Dropout(p) with a Philox keep-mask, row replication
//   sigmoid + BCE        F.sigmoid + F.binary_cross_entropy (mean) forward and gradient in one pass over the logits
//   latent               ProductOfExperts -> reparametrize -> KL for all ELBO terms at once, forward and backward
//
// Activations are matrices [rows, C] (rows = batch * H * W, NHWC), stored fp32 or bf16.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "poe.cuh"
#include "../../include/mvae_b200.h"

namespace mvae {

namespace {

constexpr int kThreads = 256;

int sm_count_() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
inline int grid_for(long long items, int per_sm = 8) {
  long long b = (items + kThreads - 1) / kThreads;
  const long long cap = static_cast<long long>(per_sm) * sm_count_();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

template <typename T>
struct V16 {
  static constexpr int N = 16 / sizeof(T);
};
template <typename T>
__device__ __forceinline__ void ldv(const T* p, float (&o)[V16<T>::N]) {
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
__device__ __forceinline__ void stv(T* p, const float (&v)[V16<T>::N]) {
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
template <typename T>
__device__ __forceinline__ float ld1(const T* p) {
  if constexpr (sizeof(T) == 4) return *p; else return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void st1(T* p, float v) {
  if constexpr (sizeof(T) == 4) *p = v; else *p = __float2bfloat16_rn(v);
}

// activation codes of the public header
__device__ __forceinline__ float act_fwd(int act, float u) {
  if (act == MVAE_ACT_RELU) return fmaxf(u, 0.f);
  if (act == MVAE_ACT_SWISH) return u * __fdividef(1.f, 1.f + __expf(-u));   // x * sigmoid(x), celeba/model.py:240-241
  return u;
}
__device__ __forceinline__ float act_grad(int act, float u) {
  if (act == MVAE_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (act == MVAE_ACT_SWISH) {
    const float s = __fdividef(1.f, 1.f + __expf(-u));
    return s * (1.f + u * (1.f - s));
  }
  return 1.f;
}

struct Geom {
  int B, H, W, C, k, stride, pad, Ho, Wo;
};

// ================================================================= im2col / col2im, vector path (NHWC, dense)
// col[m, (kh*k + kw)*C + c] = x[n, ho*stride - pad + kh, wo*stride - pad + kw, c],  m = (n*Ho + ho)*Wo + wo
template <typename T>
__global__ void __launch_bounds__(kThreads) im2col_vec_kernel(const T* __restrict__ x, T* __restrict__ col,
                                                              long long ldcol, Geom g) {
  // blockIdx.y strides over output rows (n, ho); inside a row the thread index is (wo, tap, channel vector): a row of
  // taps (kh fixed) is one contiguous run of k*C input elements and k*C col elements - a strided memcpy
  constexpr int N = V16<T>::N;
  const int CV = g.C / N, kk = g.k * g.k;
  const int per_row = g.Wo * kk * CV;
  const long long n_rows = static_cast<long long>(g.B) * g.Ho;
  for (long long row = blockIdx.y; row < n_rows; row += gridDim.y) {
    const int ho = static_cast<int>(row % g.Ho);
    const long long n = row / g.Ho;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
      const int cv = i % CV;
      const int r = i / CV;
      const int tap = r % kk;
      const int wo = r / kk;
      const int kh = tap / g.k, kw = tap - kh * g.k;
      const int hi = ho * g.stride - g.pad + kh, wi = wo * g.stride - g.pad + kw;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W)
        v = *reinterpret_cast<const uint4*>(x + ((n * g.H + hi) * g.W + wi) * g.C + cv * N);
      const long long m = row * g.Wo + wo;
      *reinterpret_cast<uint4*>(col + m * ldcol + static_cast<long long>(tap) * g.C + cv * N) = v;
    }
  }
}
// y[n, h, w, c] = sum over taps with h = ho*stride - pad + kh (same for w) of col[(n,ho,wo), tap*C + c]
template <typename T>
__global__ void __launch_bounds__(kThreads) col2im_vec_kernel(const T* __restrict__ col, long long ldcol,
                                                              T* __restrict__ y, Geom g) {
  constexpr int N = V16<T>::N;
  const int CV = g.C / N;
  const int per_row = g.W * CV;
  const long long n_rows = static_cast<long long>(g.B) * g.H;
  for (long long row = blockIdx.y; row < n_rows; row += gridDim.y) {
    const int h = static_cast<int>(row % g.H);
    const long long n = row / g.H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
      const int cv = i % CV;
      const int w = i / CV;
      float acc[N];
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = 0.f;
      // only taps with (h + pad - kh) divisible by the stride contribute: kh = (h + pad) % stride, + stride, ...
      for (int kh = (h + g.pad) % g.stride; kh < g.k; kh += g.stride) {
        const int th = h + g.pad - kh;
        if (th < 0) break;
        const int ho = th / g.stride;
        if (ho >= g.Ho) continue;
        for (int kw = (w + g.pad) % g.stride; kw < g.k; kw += g.stride) {
          const int tw = w + g.pad - kw;
          if (tw < 0) break;
          const int wo = tw / g.stride;
          if (wo >= g.Wo) continue;
          float v[N];
          ldv(col + ((n * g.Ho + ho) * g.Wo + wo) * ldcol + static_cast<long long>(kh * g.k + kw) * g.C + cv * N, v);
#pragma unroll
          for (int j = 0; j < N; ++j) acc[j] += v[j];
        }
      }
      stv(y + (row * g.W + w) * g.C + cv * N, acc);
    }
  }
}

// ================================================================= scalar strided variants (NCHW user tensors, C = 3)
template <typename TI, typename TO, int kK = 0, int kC = 0>   // kK, kC > 0: compile-time kernel size / channel count
__global__ void __launch_bounds__(kThreads) im2col_any_kernel(const TI* __restrict__ x, long long sn, long long sh,
                                                              long long sw, long long sc, TO* __restrict__ col,
                                                              long long ldcol, Geom g) {
  // one thread per (m, kh): the k taps of that kernel row x all channels are k*C consecutive col entries.  Neighbouring
  // threads read neighbouring pixels of one channel plane (NCHW sources); the run is written with 8-byte packs when aligned.
  constexpr int kMaxRun = kK > 0 ? kK * kC : 1;
  const int K = kK > 0 ? kK : g.k, Cn = kC > 0 ? kC : g.C;
  const int run = K * Cn;
  const long long total = static_cast<long long>(g.B) * g.Ho * g.Wo * K;
  const bool packed = kK > 0 && sizeof(TO) == 2 && run % 4 == 0 && ldcol % 4 == 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kh = static_cast<int>(i % K);
    const long long m = i / K;
    const int wo = static_cast<int>(m % g.Wo);
    const long long t = m / g.Wo;
    const int ho = static_cast<int>(t % g.Ho);
    const long long n = t / g.Ho;
    const int hi = ho * g.stride - g.pad + kh;
    const int wi0 = wo * g.stride - g.pad;
    const bool row_in = hi >= 0 && hi < g.H;
    const TI* src = x + n * sn + hi * sh;
    TO* dst = col + m * ldcol + static_cast<long long>(kh) * run;
    if (packed) {
      float vals[kMaxRun];
#pragma unroll
      for (int kw = 0; kw < (kK > 0 ? kK : 1); ++kw) {
        const int wi = wi0 + kw;
        const bool in = row_in && wi >= 0 && wi < g.W;
#pragma unroll
        for (int c = 0; c < (kC > 0 ? kC : 1); ++c) vals[kw * (kC > 0 ? kC : 1) + c] = in ? ld1(src + wi * sw + c * sc) : 0.f;
      }
#pragma unroll
      for (int j = 0; j + 3 < kMaxRun; j += 4) {
        {
          const __nv_bfloat162 a = __floats2bfloat162_rn(vals[j], vals[j + 1]);
          const __nv_bfloat162 b = __floats2bfloat162_rn(vals[j + 2], vals[j + 3]);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&a);
          pk.y = *reinterpret_cast<const uint32_t*>(&b);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst) + j) = pk;
        }
      }
    } else {
      for (int kw = 0; kw < K; ++kw) {
        const int wi = wi0 + kw;
        const bool in = row_in && wi >= 0 && wi < g.W;
        for (int c = 0; c < Cn; ++c) st1(dst + kw * Cn + c, in ? ld1(src + wi * sw + c * sc) : 0.f);
      }
    }
  }
}
// one thread per output pixel, all channels (w fastest): a warp walks 32 neighbouring pixels, whose taps sit in a
// handful of contiguous col rows
template <typename TI, typename TO>
__global__ void __launch_bounds__(kThreads) col2im_any_kernel(const TI* __restrict__ col, long long ldcol,
                                                              TO* __restrict__ y, long long sn, long long sh,
                                                              long long sw, long long sc, Geom g) {
  constexpr int kMaxC = 8;
  const long long total = static_cast<long long>(g.B) * g.H * g.W;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int w = static_cast<int>(i % g.W);
    long long r = i / g.W;
    const int h = static_cast<int>(r % g.H);
    const long long n = r / g.H;
    for (int c0 = 0; c0 < g.C; c0 += kMaxC) {
      const int cn = min(kMaxC, g.C - c0);
      float acc[kMaxC];
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) acc[c] = 0.f;
      for (int kh = (h + g.pad) % g.stride; kh < g.k; kh += g.stride) {
        const int th = h + g.pad - kh;
        if (th < 0) break;
        const int ho = th / g.stride;
        if (ho >= g.Ho) continue;
        for (int kw = (w + g.pad) % g.stride; kw < g.k; kw += g.stride) {
          const int tw = w + g.pad - kw;
          if (tw < 0) break;
          const int wo = tw / g.stride;
          if (wo >= g.Wo) continue;
          const TI* src = col + ((n * g.Ho + ho) * g.Wo + wo) * ldcol + static_cast<long long>(kh * g.k + kw) * g.C + c0;
#pragma unroll
          for (int c = 0; c < kMaxC; ++c)
            if (c < cn) acc[c] += ld1(src + c);
        }
      }
#pragma unroll
      for (int c = 0; c < kMaxC; ++c)
        if (c < cn) st1(y + n * sn + h * sh + w * sw + (c0 + c) * sc, acc[c]);
    }
  }
}

// ================================================================= column reductions over [rows, C]
// One block owns a slab of rows inside ONE statistics group and TX column vectors; threads (tx, ty) stride the rows,
// partials meet in shared memory, one atomic per column per block lands in the [groups, C] accumulators.
enum : int { RED_STATS = 0, RED_BN_BWD = 1, RED_ACT_BWD = 2 };

struct RedArgs {
  const void* x;      // STATS: values; BN_BWD: pre-BatchNorm x; ACT_BWD: pre-activation x [rows / repeat, C]
  const void* dy;     // BN_BWD / ACT_BWD: upstream gradient [rows, C]
  void* dx;           // ACT_BWD: gradient at the pre-activation [rows / repeat, C]
  long long rows;     // rows of the reduction (ACT_BWD: rows of x)
  int C, c_valid;     // columns >= c_valid are skipped when accumulating
  long long rows_per_group;
  int groups, chunks; // grid.x = groups * chunks
  int act;
  float* a0;          // [groups, C] += : STATS sum(x) ; BN_BWD sum(dyhat) ; ACT_BWD sum(dx)
  float* a1;          // [groups, C] += : STATS sum(x^2) ; BN_BWD sum(dyhat * xhat) ; may be null
  const float* mean;  // BN_BWD [groups, C]
  const float* rstd;
  const float* gamma; // [C]
  const float* beta;
  // ACT_BWD dropout / replication
  int repeat;
  float keep_scale;   // 1 / (1 - p), 0 when no dropout
  uint32_t keep_thresh;  // keep iff 16-bit uniform < keep_thresh
  unsigned long long seed;
  const int* step_ptr;
};

// 16-bit uniforms for the N elements of vector `vec_index` (Dropout keep-mask, shared by forward and backward)
template <int N>
__device__ __forceinline__ void dropout_bits(unsigned long long seed, uint32_t step, unsigned long long vec_index,
                                             uint32_t (&u)[N]) {
  uint32_t r[4];
  philox4x32(static_cast<uint32_t>(vec_index), static_cast<uint32_t>(vec_index >> 32), step, 0x64726f70u,
             static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  if constexpr (N == 4) {
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = r[i] & 0xFFFFu;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      u[2 * i] = r[i] & 0xFFFFu;
      u[2 * i + 1] = r[i] >> 16;
    }
  }
}

template <typename T, int kMode>
__global__ void __launch_bounds__(kThreads) col_reduce_kernel(const RedArgs a) {
  constexpr int N = V16<T>::N;
  extern __shared__ float s_part[];  // [TY][TX * N * 2]
  const int CV = a.C / N;
  const int TX = min(CV, kThreads);
  const int TY = kThreads / TX;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int cv = blockIdx.y * TX + tx;
  const bool active = ty < TY && cv < CV;
  const int g = blockIdx.x / a.chunks, chunk = blockIdx.x - g * a.chunks;
  const long long g0 = g * a.rows_per_group, g1 = min(a.rows, g0 + a.rows_per_group);
  const long long slab = (g1 - g0 + a.chunks - 1) / a.chunks;
  const long long r0 = g0 + chunk * slab, r1 = min(g1, r0 + slab);
  const T* x = static_cast<const T*>(a.x);
  const T* dy = static_cast<const T*>(a.dy);
  T* dx = static_cast<T*>(a.dx);
  float acc0[N], acc1[N];
#pragma unroll
  for (int j = 0; j < N; ++j) acc0[j] = acc1[j] = 0.f;
  if (active) {
    const int c = cv * N;
    float sc[N], sh[N], mu[N], rs[N];
    if (kMode == RED_BN_BWD) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        mu[j] = a.mean[g * a.C + c + j];
        rs[j] = a.rstd[g * a.C + c + j];
        sc[j] = a.gamma[c + j] * rs[j];
        sh[j] = fmaf(-mu[j], sc[j], a.beta[c + j]);
      }
    }
    const uint32_t step = (kMode == RED_ACT_BWD && a.step_ptr != nullptr) ? static_cast<uint32_t>(*a.step_ptr) : 0u;
    long long r = r0 + ty;
    if (kMode != RED_ACT_BWD) {
      // four rows in flight per thread (the loads of a batch are issued before any of its math)
      constexpr int U = 4;
      for (; r + static_cast<long long>(U - 1) * TY < r1; r += static_cast<long long>(U) * TY) {
        float v[U][N], d[U][N];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          ldv(x + (r + static_cast<long long>(u) * TY) * a.C + c, v[u]);
          if (kMode == RED_BN_BWD) ldv(dy + (r + static_cast<long long>(u) * TY) * a.C + c, d[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int j = 0; j < N; ++j) {
            if (kMode == RED_STATS) {
              acc0[j] += v[u][j];
              acc1[j] += v[u][j] * v[u][j];
            } else {
              const float uu = fmaf(sc[j], v[u][j], sh[j]);
              const float dh = d[u][j] * act_grad(a.act, uu);
              acc0[j] += dh;
              acc1[j] += dh * (v[u][j] - mu[j]) * rs[j];
            }
          }
        }
      }
    }
    for (; r < r1; r += TY) {
      float v[N];
      ldv(x + r * a.C + c, v);
      if (kMode == RED_STATS) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          acc0[j] += v[j];
          acc1[j] += v[j] * v[j];
        }
      } else if (kMode == RED_BN_BWD) {
        float d[N];
        ldv(dy + r * a.C + c, d);
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float u = fmaf(sc[j], v[j], sh[j]);
          const float dh = d[j] * act_grad(a.act, u);
          acc0[j] += dh;
          acc1[j] += dh * (v[j] - mu[j]) * rs[j];
        }
      } else {
        float o[N];
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = 0.f;
        for (int rep = 0; rep < a.repeat; ++rep) {
          const long long ro = rep * a.rows + r;
          float d[N];
          ldv(dy + ro * a.C + c, d);
          if (a.keep_scale > 0.f) {
            uint32_t u[N];
            dropout_bits<N>(a.seed, step, static_cast<unsigned long long>(ro) * CV + cv, u);
#pragma unroll
            for (int j = 0; j < N; ++j) d[j] = u[j] < a.keep_thresh ? d[j] * a.keep_scale : 0.f;
          }
#pragma unroll
          for (int j = 0; j < N; ++j) o[j] += d[j];
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
          o[j] *= act_grad(a.act, v[j]);
          acc0[j] += o[j];
        }
        stv(dx + r * a.C + c, o);
      }
    }
  }
  if (a.a0 == nullptr) return;
  float* mine = s_part + (static_cast<size_t>(ty) * TX + tx) * (2 * N);
  if (active) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      mine[j] = acc0[j];
      mine[N + j] = acc1[j];
    }
  }
  __syncthreads();
  if (active && ty == 0) {
    for (int t = 1; t < TY; ++t) {
      const float* o = s_part + (static_cast<size_t>(t) * TX + tx) * (2 * N);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        acc0[j] += o[j];
        acc1[j] += o[N + j];
      }
    }
    const int c = cv * N;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (c + j < a.c_valid) {
        atomicAdd(a.a0 + g * a.C + c + j, acc0[j]);
        if (a.a1 != nullptr) atomicAdd(a.a1 + g * a.C + c + j, acc1[j]);
      }
    }
  }
}

template <typename T, int kMode>
int launch_col_reduce_t(RedArgs a, cudaStream_t st) {
  constexpr int N = V16<T>::N;
  MVAE_REQUIRE(a.C % N == 0, "column reduction: %d columns must be a multiple of %d", a.C, N);
  MVAE_REQUIRE(a.rows > 0 && a.rows_per_group > 0, "column reduction: empty input");
  const int CV = a.C / N;
  const int TX = std::min(CV, kThreads);
  const int gy = (CV + TX - 1) / TX;
  a.groups = static_cast<int>((a.rows + a.rows_per_group - 1) / a.rows_per_group);
  const long long rows_g = std::min(a.rows, a.rows_per_group);
  const int TY = kThreads / TX;
  long long chunks = (4ll * sm_count_()) / (static_cast<long long>(a.groups) * gy);
  const long long max_chunks = (rows_g + TY * 4 - 1) / (TY * 4);  // at least 4 row iterations per thread
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  a.chunks = static_cast<int>(chunks);
  const size_t smem = static_cast<size_t>(TY) * TX * 2 * N * sizeof(float);
  col_reduce_kernel<T, kMode><<<dim3(a.groups * a.chunks, gy), kThreads, smem, st>>>(a);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}
template <int kMode>
int launch_col_reduce(int dtype, const RedArgs& a, cudaStream_t st) {
  if (dtype == MVAE_F32) return launch_col_reduce_t<float, kMode>(a, st);
  return launch_col_reduce_t<__nv_bfloat16, kMode>(a, st);
}

// ================================================================= BatchNorm (+ activation) apply kernels
// Every block first builds the per-(group, channel) coefficient table in shared memory from the batch sums
// (block 0 also publishes mean / rstd, the running statistics and - backward - dgamma / dbeta), then streams.
struct BnArgs {
  const void* x; void* y;          // forward: x -> y ; backward: x (pre-BN), dy -> dx
  const void* dy; void* dx;
  long long rows; int C; long long rows_per_group; int groups;
  int act, training;
  const float* sum; const float* sumsq;      // forward batch sums [groups, C]
  const float* gamma; const float* beta;
  float* save_mean; float* save_rstd;        // [groups, C]
  float* running_mean; float* running_var; int updates; float momentum, eps;
  const float* s0; const float* s1;          // backward sums [groups, C]
  float* dgamma; float* dbeta;
};

template <typename T>
__global__ void __launch_bounds__(kThreads) bn_act_fwd_kernel(const BnArgs a) {
  constexpr int N = V16<T>::N;
  extern __shared__ float s_tab[];  // [groups*C] scale | [groups*C] shift
  const int GC = a.groups * a.C;
  float* s_scale = s_tab;
  float* s_shift = s_tab + GC;
  for (int i = threadIdx.x; i < GC; i += kThreads) {
    const int g = i / a.C, c = i - g * a.C;
    float mean, rstd;
    if (a.training) {
      const long long g0 = g * a.rows_per_group;
      const long long cnt = min(a.rows, g0 + a.rows_per_group) - g0;
      const float inv = 1.f / static_cast<float>(cnt);
      mean = a.sum[i] * inv;
      const float var = fmaxf(a.sumsq[i] * inv - mean * mean, 0.f);
      rstd = rsqrtf(var + a.eps);
      if (blockIdx.x == 0) {
        if (a.save_mean != nullptr) {
          a.save_mean[i] = mean;
          a.save_rstd[i] = rstd;
        }
        if (a.running_mean != nullptr) {
          // groups are separate forward passes of the reference (one per ELBO term), applied in order
          const float unb = cnt > 1 ? var * (static_cast<float>(cnt) / static_cast<float>(cnt - 1)) : var;
          if (g == 0) {
            float rm = a.running_mean[c], rv = a.running_var[c];
            for (int gg = 0; gg < a.groups; ++gg) {
              float m2 = mean, u2 = unb;
              if (gg > 0) {
                const long long h0 = gg * a.rows_per_group;
                const long long cn = min(a.rows, h0 + a.rows_per_group) - h0;
                const float iv = 1.f / static_cast<float>(cn);
                m2 = a.sum[gg * a.C + c] * iv;
                const float v2 = fmaxf(a.sumsq[gg * a.C + c] * iv - m2 * m2, 0.f);
                u2 = cn > 1 ? v2 * (static_cast<float>(cn) / static_cast<float>(cn - 1)) : v2;
              }
              for (int u = 0; u < a.updates; ++u) {
                rm = (1.f - a.momentum) * rm + a.momentum * m2;
                rv = (1.f - a.momentum) * rv + a.momentum * u2;
              }
            }
            a.running_mean[c] = rm;
            a.running_var[c] = rv;
          }
        }
      }
    } else {
      mean = a.running_mean[c];
      rstd = rsqrtf(a.running_var[c] + a.eps);
    }
    const float sc = a.gamma[c] * rstd;
    s_scale[i] = sc;
    s_shift[i] = fmaf(-mean, sc, a.beta[c]);
  }
  __syncthreads();
  const int CV = a.C / N;
  const long long total = a.rows * CV;
  const T* x = static_cast<const T*>(a.x);
  T* y = static_cast<T*>(a.y);
  const long long stride = static_cast<long long>(gridDim.x) * kThreads;
  long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x;
  constexpr int U = 4;   // vectors in flight per thread
  for (; i + (U - 1) * stride < total; i += U * stride) {
    float v[U][N];
#pragma unroll
    for (int u = 0; u < U; ++u) ldv(x + (i + u * stride) * N, v[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ii = i + u * stride;
      const int cv = static_cast<int>(ii % CV);
      const int g = static_cast<int>((ii / CV) / a.rows_per_group);
      const int t = g * a.C + cv * N;
#pragma unroll
      for (int j = 0; j < N; ++j) v[u][j] = act_fwd(a.act, fmaf(s_scale[t + j], v[u][j], s_shift[t + j]));
      stv(y + ii * N, v[u]);
    }
  }
  for (; i < total; i += stride) {
    const int cv = static_cast<int>(i % CV);
    const long long r = i / CV;
    const int g = static_cast<int>(r / a.rows_per_group);
    const int t = g * a.C + cv * N;
    float v[N];
    ldv(x + i * N, v);
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = act_fwd(a.act, fmaf(s_scale[t + j], v[j], s_shift[t + j]));
    stv(y + i * N, v);
  }
}

// dx = gamma*rstd * (dyhat - S0/cnt - xhat*S1/cnt), dyhat = dy * act'(gamma*xhat + beta)
template <typename T>
__global__ void __launch_bounds__(kThreads) bn_act_bwd_kernel(const BnArgs a) {
  constexpr int N = V16<T>::N;
  extern __shared__ float s_tab[];  // scale | shift | c0 | c1 (multiplies x - mean) | mean
  const int GC = a.groups * a.C;
  float* s_scale = s_tab;
  float* s_shift = s_tab + GC;
  float* s_c0 = s_tab + 2 * GC;
  float* s_c1 = s_tab + 3 * GC;
  float* s_mean = s_tab + 4 * GC;
  for (int i = threadIdx.x; i < GC; i += kThreads) {
    const int g = i / a.C, c = i - g * a.C;
    const long long g0 = g * a.rows_per_group;
    const long long cnt = min(a.rows, g0 + a.rows_per_group) - g0;
    const float inv = 1.f / static_cast<float>(cnt);
    const float mean = a.save_mean[i], rstd = a.save_rstd[i];
    const float sc = a.gamma[c] * rstd;
    s_scale[i] = sc;
    s_shift[i] = fmaf(-mean, sc, a.beta[c]);
    s_c0[i] = a.s0[i] * inv;
    s_c1[i] = a.s1[i] * inv * rstd;   // multiplies (x - mean): xhat * S1/cnt
    s_mean[i] = mean;
    if (blockIdx.x == 0 && g == 0 && a.dgamma != nullptr) {
      float dg = 0.f, db = 0.f;
      for (int gg = 0; gg < a.groups; ++gg) {
        dg += a.s1[gg * a.C + c];
        db += a.s0[gg * a.C + c];
      }
      a.dgamma[c] += dg;
      a.dbeta[c] += db;
    }
  }
  __syncthreads();
  const int CV = a.C / N;
  const long long total = a.rows * CV;
  const T* x = static_cast<const T*>(a.x);
  const T* dy = static_cast<const T*>(a.dy);
  T* dx = static_cast<T*>(a.dx);
  const long long stride = static_cast<long long>(gridDim.x) * kThreads;
  long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x;
  constexpr int U = 2;   // (x, dy) vector pairs in flight per thread
  for (; i + (U - 1) * stride < total; i += U * stride) {
    float v[U][N], d[U][N];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ldv(x + (i + u * stride) * N, v[u]);
      ldv(dy + (i + u * stride) * N, d[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ii = i + u * stride;
      const int cv = static_cast<int>(ii % CV);
      const int g = static_cast<int>((ii / CV) / a.rows_per_group);
      const int t = g * a.C + cv * N;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float uu = fmaf(s_scale[t + j], v[u][j], s_shift[t + j]);
        const float dh = d[u][j] * act_grad(a.act, uu);
        d[u][j] = s_scale[t + j] * (dh - s_c0[t + j] - (v[u][j] - s_mean[t + j]) * s_c1[t + j]);
      }
      stv(dx + ii * N, d[u]);
    }
  }
  for (; i < total; i += stride) {
    const int cv = static_cast<int>(i % CV);
    const long long r = i / CV;
    const int g = static_cast<int>(r / a.rows_per_group);
    const int t = g * a.C + cv * N;
    float v[N], d[N];
    ldv(x + i * N, v);
    ldv(dy + i * N, d);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float u = fmaf(s_scale[t + j], v[j], s_shift[t + j]);
      const float dh = d[j] * act_grad(a.act, u);
      d[j] = s_scale[t + j] * (dh - s_c0[t + j] - (v[j] - s_mean[t + j]) * s_c1[t + j]);
    }
    stv(dx + i * N, d);
  }
}

// ================================================================= activation (+ dropout, row replication) forward
template <typename T>
__global__ void __launch_bounds__(kThreads) act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long rows,
                                                           int C, int act, int repeat, float keep_scale,
                                                           uint32_t keep_thresh, unsigned long long seed,
                                                           const int* __restrict__ step_ptr) {
  constexpr int N = V16<T>::N;
  const int CV = C / N;
  const long long total = rows * CV;
  const uint32_t step = step_ptr != nullptr ? static_cast<uint32_t>(*step_ptr) : 0u;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    float v[N];
    ldv(x + i * N, v);
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = act_fwd(act, v[j]);
    for (int rep = 0; rep < repeat; ++rep) {
      const long long o = rep * total + i;  // == (rep*rows + r) * CV + cv
      float w[N];
#pragma unroll
      for (int j = 0; j < N; ++j) w[j] = v[j];
      if (keep_scale > 0.f) {
        uint32_t u[N];
        dropout_bits<N>(seed, step, static_cast<unsigned long long>(o), u);
#pragma unroll
        for (int j = 0; j < N; ++j) w[j] = u[j] < keep_thresh ? w[j] * keep_scale : 0.f;
      }
      stv(y + o * N, w);
    }
  }
}

// ================================================================= sigmoid + binary cross entropy, value and gradient
// loss[g] += sum over the group's elements of BCE(sigmoid(x), t)   (log terms clamped at -100 like ATen)
// dlogit = scale[g] * (sigmoid(x) - t);  element (m, c) reads target[(m % target_rows), c].
struct BceArgs {
  const void* logits; long long ldl; int logit_dtype;
  const void* target; long long ldt; int target_dtype; long long target_rows;
  long long rows; int cols; long long rows_per_group;
  float scale[kMaxGroups];
  float* loss;
  void* probs; long long ldp; int prob_dtype;
  void* dlogits; long long ldd; int grad_dtype; int ld_pad_zero;  // zero the [cols, ldd) tail of each gradient row
  const void* dprobs; long long lddp;   // optional upstream gradient w.r.t. the probabilities (module path), fp32
};
__device__ __forceinline__ float ld_any(const void* p, int dtype, long long i) {
  return dtype == MVAE_F32 ? static_cast<const float*>(p)[i] : __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dtype, long long i, float v) {
  if (dtype == MVAE_F32) static_cast<float*>(p)[i] = v; else static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}
// fp32, 4 elements per thread, the training configuration of the image decoders: logits, targets and gradients fp32 and
// vector-aligned, loss + dlogits (+ probs).  Same arithmetic as the generic kernel below.
__global__ void __launch_bounds__(kThreads) sigmoid_bce_f32x4_kernel(const BceArgs a) {
  __shared__ float s_loss[kMaxGroups];
  if (threadIdx.x < kMaxGroups) s_loss[threadIdx.x] = 0.f;
  __syncthreads();
  const int c4 = a.cols / 4;
  const long long total = a.rows * c4;
  const float* logits = static_cast<const float*>(a.logits);
  const float* target = static_cast<const float*>(a.target);
  float* dl = static_cast<float*>(a.dlogits);
  float* pr = static_cast<float*>(a.probs);
  float acc[kMaxGroups] = {0.f, 0.f, 0.f};
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int c = static_cast<int>(i % c4) * 4;
    const long long m = i / c4;
    const int g = static_cast<int>(m / a.rows_per_group);
    const float4 x4 = *reinterpret_cast<const float4*>(logits + m * a.ldl + c);
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(target + (m % a.target_rows) * a.ldt + c));
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ts[4] = {t4.x, t4.y, t4.z, t4.w};
    float ps[4], ds[4], l = 0.f;
    const float sc = a.scale[g];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float x = xs[k], t = ts[k];
      const float e = __expf(-fabsf(x));
      const float r = __fdividef(1.f, 1.f + e);
      const float p = x >= 0.f ? r : e * r;
      const float sp = __logf(1.f + e);
      const float log_p = fmaxf(-(fmaxf(-x, 0.f) + sp), -100.f);
      const float log_q = fmaxf(-(fmaxf(x, 0.f) + sp), -100.f);
      l -= t * log_p + (1.f - t) * log_q;
      ps[k] = p;
      ds[k] = sc * (p - t);
    }
    if (g == 0) acc[0] += l; else if (g == 1) acc[1] += l; else acc[2] += l;
    if (pr != nullptr) *reinterpret_cast<float4*>(pr + m * a.ldp + c) = make_float4(ps[0], ps[1], ps[2], ps[3]);
    if (dl != nullptr) *reinterpret_cast<float4*>(dl + m * a.ldd + c) = make_float4(ds[0], ds[1], ds[2], ds[3]);
  }
  if (a.loss != nullptr) {
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      float v = acc[g];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&s_loss[g], v);
    }
    __syncthreads();
    if (threadIdx.x < kMaxGroups && s_loss[threadIdx.x] != 0.f) atomicAdd(a.loss + threadIdx.x, s_loss[threadIdx.x]);
  }
}

__global__ void __launch_bounds__(kThreads) sigmoid_bce_kernel(const BceArgs a) {
  __shared__ float s_loss[kMaxGroups];
  if (threadIdx.x < kMaxGroups) s_loss[threadIdx.x] = 0.f;
  __syncthreads();
  const int wcols = a.ld_pad_zero ? static_cast<int>(a.ldd) : a.cols;
  const long long total = a.rows * wcols;
  float acc[kMaxGroups] = {0.f, 0.f, 0.f};
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int c = static_cast<int>(i % wcols);
    const long long m = i / wcols;
    if (c >= a.cols) {
      st_any(a.dlogits, a.grad_dtype, m * a.ldd + c, 0.f);
      continue;
    }
    const int g = static_cast<int>(m / a.rows_per_group);
    const float x = ld_any(a.logits, a.logit_dtype, m * a.ldl + c);
    const float e = __expf(-fabsf(x));
    const float r = __fdividef(1.f, 1.f + e);
    const float p = x >= 0.f ? r : e * r;
    if (a.probs != nullptr) st_any(a.probs, a.prob_dtype, m * a.ldp + c, p);
    if (a.target != nullptr) {
      const float t = ld_any(a.target, a.target_dtype, (m % a.target_rows) * a.ldt + c);
      const float sp = __logf(1.f + e);                       // log(1 + exp(-|x|))
      const float log_p = fmaxf(-(fmaxf(-x, 0.f) + sp), -100.f);   // log sigmoid(x)
      const float log_q = fmaxf(-(fmaxf(x, 0.f) + sp), -100.f);    // log (1 - sigmoid(x))
      const float l = -(t * log_p + (1.f - t) * log_q);
      if (g == 0) acc[0] += l; else if (g == 1) acc[1] += l; else acc[2] += l;
      if (a.dlogits != nullptr) st_any(a.dlogits, a.grad_dtype, m * a.ldd + c, a.scale[g] * (p - t));
    } else if (a.dlogits != nullptr && a.dprobs != nullptr) {
      const float dp = static_cast<const float*>(a.dprobs)[m * a.lddp + c];
      st_any(a.dlogits, a.grad_dtype, m * a.ldd + c, dp * p * (1.f - p));
    }
  }
  if (a.loss != nullptr) {
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      float v = acc[g];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&s_loss[g], v);
    }
    __syncthreads();
    if (threadIdx.x < kMaxGroups && s_loss[threadIdx.x] != 0.f) atomicAdd(a.loss + threadIdx.x, s_loss[threadIdx.x]);
  }
}

// ================================================================= latent path: PoE -> reparametrize -> KL, all terms
// One thread owns one (sample, latent) for every ELBO term, so the gradients of an expert shared by two terms are
// summed in registers and written once (no atomics on the expert gradients).
struct LatentArgs {
  long long B; int n, G;
  int term_type[kMaxGroups];
  int poe_mode, prior; float poe_eps;
  const float* enc_a; long long ld_a; long long a_row0[kMaxGroups];  // expert A (image) rows of term g start at a_row0[g]
  const float* enc_b; long long ld_b;                                // expert B (text / attrs) [B, 2n]
  const float* eps; unsigned long long seed; const int* step_ptr; int training;
  float kl_weight[kMaxGroups];
  void* z; long long ldz; int z_dtype;      // [G*B, ldz]
  float* mu; float* logvar;                 // [G, B, n] optional
  float* kl;                                // [G] += kl_weight[g] * KL_g
  const void* dz; long long lddz; int dz_dtype;      // backward: [G*B, lddz]
  const float* dmu_up; const float* dlogvar_up;      // optional [G, B, n]
  void* d_enc_a; long long ld_da; int d_dtype;       // rows like enc_a
  void* d_enc_b; long long ld_db;
  const float* row_w;                                // optional [G * B]: scales kl_weight[g] per row
};

template <bool kBackward>
__global__ void __launch_bounds__(kThreads) latent_kernel(const LatentArgs a) {
  __shared__ float s_kl[kMaxGroups];
  if (threadIdx.x < kMaxGroups) s_kl[threadIdx.x] = 0.f;
  __syncthreads();
  const long long total = a.B * a.n;
  const uint32_t step = a.step_ptr != nullptr ? static_cast<uint32_t>(*a.step_ptr) : 0u;
  float klacc[kMaxGroups] = {0.f, 0.f, 0.f};
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int j = static_cast<int>(i % a.n);
    const long long b = i / a.n;
    float mb = 0.f, lb = 0.f;
    if (a.enc_b != nullptr) {
      mb = a.enc_b[b * a.ld_b + j];
      lb = a.enc_b[b * a.ld_b + a.n + j];
    }
    float dA_m[kMaxGroups] = {0.f, 0.f, 0.f}, dA_l[kMaxGroups] = {0.f, 0.f, 0.f};
    float dB_m = 0.f, dB_l = 0.f;
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      if (g >= a.G) break;
      const int ty = a.term_type[g];
      const bool present[2] = {ty != MVAE_TERM_TEXT && a.enc_a != nullptr, ty != MVAE_TERM_IMAGE && a.enc_b != nullptr};
      float m[2] = {0.f, mb}, lv[2] = {0.f, lb};
      if (present[0]) {
        m[0] = a.enc_a[(a.a_row0[g] + b) * a.ld_a + j];
        lv[0] = a.enc_a[(a.a_row0[g] + b) * a.ld_a + a.n + j];
      }
      const Poe r = poe_eval<true>(a.poe_mode, a.prior, a.poe_eps, m, lv, present);
      const long long e = (static_cast<long long>(g) * a.B + b) * a.n + j;
      const float klw = a.row_w != nullptr ? a.kl_weight[g] * a.row_w[static_cast<long long>(g) * a.B + b] : a.kl_weight[g];
      float noise = 0.f;
      if (a.training) {
        if (a.eps != nullptr) {
          noise = a.eps[e];
        } else {
          const float2 p2 = normal_pair(a.seed, step, static_cast<unsigned long long>(e >> 1));
          noise = (e & 1) ? p2.y : p2.x;
        }
      }
      const float sd = sqrtf(r.pd_var);  // exp(logvar / 2)
      if (!kBackward) {
        const long long row = static_cast<long long>(g) * a.B + b;
        const float zv = a.training ? fmaf(noise, sd, r.mu) : r.mu;
        if (a.z_dtype == MVAE_F32) static_cast<float*>(a.z)[row * a.ldz + j] = zv;
        else static_cast<__nv_bfloat16*>(a.z)[row * a.ldz + j] = __float2bfloat16_rn(zv);
        if (a.mu != nullptr) {
          a.mu[e] = r.mu;
          a.logvar[e] = r.logvar;
        }
        // -0.5 * (1 + logvar - mu^2 - exp(logvar)), celeba/train.py:79
        klacc[g] += klw * (-0.5f) * (1.f + r.logvar - r.mu * r.mu - r.pd_var);
      } else {
        const long long row = static_cast<long long>(g) * a.B + b;
        float dzv = 0.f;
        if (a.dz != nullptr) dzv = ld_any(a.dz, a.dz_dtype, row * a.lddz + j);
        float dmu = dzv + klw * r.mu;
        float dlv = (a.training ? dzv * 0.5f * noise * sd : 0.f) + klw * 0.5f * (r.pd_var - 1.f);
        if (a.dmu_up != nullptr) {
          dmu += a.dmu_up[e];
          dlv += a.dlogvar_up[e];
        }
        if (present[0]) poe_grad(a.poe_mode, a.poe_eps, r, 0, m[0], dmu, dlv, dA_m[g], dA_l[g]);
        if (present[1]) {
          float gm, gl;
          poe_grad(a.poe_mode, a.poe_eps, r, 1, m[1], dmu, dlv, gm, gl);
          dB_m += gm;
          dB_l += gl;
        }
      }
    }
    if (kBackward) {
      if (a.d_enc_a != nullptr) {
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g) {
          if (g >= a.G || a.term_type[g] == MVAE_TERM_TEXT) continue;
          bool dup = false;
          float sm = dA_m[g], sl = dA_l[g];
#pragma unroll
          for (int h = 0; h < kMaxGroups; ++h) {
            if (h >= a.G || h == g || a.term_type[h] == MVAE_TERM_TEXT || a.a_row0[h] != a.a_row0[g]) continue;
            if (h < g) dup = true;
            sm += dA_m[h];
            sl += dA_l[h];
          }
          if (dup) continue;
          st_any(a.d_enc_a, a.d_dtype, (a.a_row0[g] + b) * a.ld_da + j, sm);
          st_any(a.d_enc_a, a.d_dtype, (a.a_row0[g] + b) * a.ld_da + a.n + j, sl);
        }
      }
      if (a.d_enc_b != nullptr) {
        st_any(a.d_enc_b, a.d_dtype, b * a.ld_db + j, dB_m);
        st_any(a.d_enc_b, a.d_dtype, b * a.ld_db + a.n + j, dB_l);
      }
    }
  }
  if (!kBackward && a.kl != nullptr) {
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      float v = klacc[g];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&s_kl[g], v);
    }
    __syncthreads();
    if (threadIdx.x < kMaxGroups && s_kl[threadIdx.x] != 0.f) atomicAdd(a.kl + threadIdx.x, s_kl[threadIdx.x]);
  }
}

// ================================================================= small helpers
template <typename TO>
__global__ void __launch_bounds__(kThreads) cast_pad_kernel(const float* __restrict__ src, long long rows, long long cols,
                                                            long long ld_src, TO* __restrict__ dst, long long ld_dst) {
  const long long total = rows * ld_dst;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const long long c = i % ld_dst, r = i / ld_dst;
    st1(dst + i, c < cols ? src[r * ld_src + c] : 0.f);
  }
}

__global__ void __launch_bounds__(kThreads) step_begin_kernel(int* step_ptr, float4* zero_buf, long long n4,
                                                              long long* counters, const long long* inc, int n_counters) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && step_ptr != nullptr) *step_ptr += 1;
  if (blockIdx.x == 0 && counters != nullptr && threadIdx.x < n_counters) counters[threadIdx.x] += inc[threadIdx.x];
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * kThreads)
    zero_buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

Geom make_geom(const mvae_conv_geometry* q) {
  Geom g;
  g.B = q->batch; g.H = q->height; g.W = q->width; g.C = q->channels;
  g.k = q->kernel; g.stride = q->stride; g.pad = q->pad;
  g.Ho = (q->height + 2 * q->pad - q->kernel) / q->stride + 1;
  g.Wo = (q->width + 2 * q->pad - q->kernel) / q->stride + 1;
  return g;
}
int check_geom(const mvae_conv_geometry* q) {
  MVAE_REQUIRE(q != nullptr, "conv geometry: null");
  MVAE_REQUIRE(q->batch > 0 && q->height > 0 && q->width > 0 && q->channels > 0, "conv geometry: empty tensor");
  MVAE_REQUIRE(q->kernel > 0 && q->stride > 0 && q->pad >= 0 && q->kernel <= 8, "conv geometry: bad kernel/stride/pad");
  MVAE_REQUIRE(q->height + 2 * q->pad >= q->kernel && q->width + 2 * q->pad >= q->kernel, "conv geometry: kernel larger than input");
  return 0;
}
bool dense_nhwc(const mvae_conv_geometry* q) {
  return q->stride_c == 1 && q->stride_w == q->channels && q->stride_h == static_cast<long long>(q->width) * q->channels &&
         q->stride_n == static_cast<long long>(q->height) * q->width * q->channels;
}

}  // namespace

}  // namespace mvae

using namespace mvae;

extern "C" {

int mvae_conv_out_size(int in, int kernel, int stride, int pad) { return (in + 2 * pad - kernel) / stride + 1; }

int mvae_im2col(const mvae_conv_geometry* q, int image_dtype, const void* image, int col_dtype, void* col, int64_t ldcol,
                void* stream) {
  if (int rc = check_geom(q)) return rc;
  MVAE_REQUIRE(image != nullptr && col != nullptr, "im2col: null tensor");
  const Geom g = make_geom(q);
  const long long K = static_cast<long long>(g.k) * g.k * g.C;
  MVAE_REQUIRE(ldcol >= K, "im2col: ldcol %lld < K %lld", static_cast<long long>(ldcol), K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vec = col_dtype == MVAE_F32 ? 4 : 8;
  const long long items = static_cast<long long>(g.B) * g.Ho * g.Wo * K;
  if (image_dtype == col_dtype && dense_nhwc(q) && g.C % vec == 0 && ldcol % vec == 0 &&
      reinterpret_cast<uintptr_t>(image) % 16 == 0 && reinterpret_cast<uintptr_t>(col) % 16 == 0) {
    const int per_row = g.Wo * g.k * g.k * (g.C / vec);
    const int threads = std::max(64, std::min(kThreads, (per_row + 31) / 32 * 32));
    const dim3 grid(std::max(1, std::min((per_row + threads - 1) / threads, 8)),
                    static_cast<unsigned>(std::min<long long>(static_cast<long long>(g.B) * g.Ho, 16 * sm_count_())));
    if (col_dtype == MVAE_F32)
      im2col_vec_kernel<float><<<grid, threads, 0, st>>>(static_cast<const float*>(image), static_cast<float*>(col), ldcol, g);
    else
      im2col_vec_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(static_cast<const __nv_bfloat16*>(image),
                                                                   static_cast<__nv_bfloat16*>(col), ldcol, g);
  } else {
    const int blocks = grid_for(items / (static_cast<long long>(g.C) * g.k));
#define MVAE_I2C(TI, TO)                                                                                              \
  im2col_any_kernel<TI, TO><<<blocks, kThreads, 0, st>>>(static_cast<const TI*>(image), q->stride_n, q->stride_h,     \
                                                         q->stride_w, q->stride_c, static_cast<TO*>(col), ldcol, g)
    if (image_dtype == MVAE_F32 && col_dtype == MVAE_F32) MVAE_I2C(float, float);
    else if (image_dtype == MVAE_F32 && g.k == 4 && g.C == 3)
      im2col_any_kernel<float, __nv_bfloat16, 4, 3><<<blocks, kThreads, 0, st>>>(
          static_cast<const float*>(image), q->stride_n, q->stride_h, q->stride_w, q->stride_c, static_cast<__nv_bfloat16*>(col), ldcol, g);
    else if (image_dtype == MVAE_F32 && g.k == 4 && g.C == 1)
      im2col_any_kernel<float, __nv_bfloat16, 4, 1><<<blocks, kThreads, 0, st>>>(
          static_cast<const float*>(image), q->stride_n, q->stride_h, q->stride_w, q->stride_c, static_cast<__nv_bfloat16*>(col), ldcol, g);
    else if (image_dtype == MVAE_F32) MVAE_I2C(float, __nv_bfloat16);
    else if (col_dtype == MVAE_F32) MVAE_I2C(__nv_bfloat16, float);
    else MVAE_I2C(__nv_bfloat16, __nv_bfloat16);
#undef MVAE_I2C
  }
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_col2im(const mvae_conv_geometry* q, int col_dtype, const void* col, int64_t ldcol, int image_dtype, void* image,
                void* stream) {
  if (int rc = check_geom(q)) return rc;
  MVAE_REQUIRE(image != nullptr && col != nullptr, "col2im: null tensor");
  const Geom g = make_geom(q);
  const long long K = static_cast<long long>(g.k) * g.k * g.C;
  MVAE_REQUIRE(ldcol >= K, "col2im: ldcol %lld < K %lld", static_cast<long long>(ldcol), K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vec = col_dtype == MVAE_F32 ? 4 : 8;
  const long long items = static_cast<long long>(g.B) * g.H * g.W * g.C;
  if (image_dtype == col_dtype && dense_nhwc(q) && g.C % vec == 0 && ldcol % vec == 0 &&
      reinterpret_cast<uintptr_t>(image) % 16 == 0 && reinterpret_cast<uintptr_t>(col) % 16 == 0) {
    const int per_row = g.W * (g.C / vec);
    const int threads = std::max(64, std::min(kThreads, (per_row + 31) / 32 * 32));
    const dim3 grid(std::max(1, std::min((per_row + threads - 1) / threads, 8)),
                    static_cast<unsigned>(std::min<long long>(static_cast<long long>(g.B) * g.H, 16 * sm_count_())));
    if (col_dtype == MVAE_F32)
      col2im_vec_kernel<float><<<grid, threads, 0, st>>>(static_cast<const float*>(col), ldcol, static_cast<float*>(image), g);
    else
      col2im_vec_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(static_cast<const __nv_bfloat16*>(col), ldcol,
                                                                   static_cast<__nv_bfloat16*>(image), g);
  } else {
    const int blocks = grid_for(items / g.C);
#define MVAE_C2I(TI, TO)                                                                                          \
  col2im_any_kernel<TI, TO><<<blocks, kThreads, 0, st>>>(static_cast<const TI*>(col), ldcol, static_cast<TO*>(image), \
                                                         q->stride_n, q->stride_h, q->stride_w, q->stride_c, g)
    if (col_dtype == MVAE_F32 && image_dtype == MVAE_F32) MVAE_C2I(float, float);
    else if (col_dtype == MVAE_F32) MVAE_C2I(float, __nv_bfloat16);
    else if (image_dtype == MVAE_F32) MVAE_C2I(__nv_bfloat16, float);
    else MVAE_C2I(__nv_bfloat16, __nv_bfloat16);
#undef MVAE_C2I
  }
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_col_stats(int dtype, const void* x, int64_t rows, int channels, int valid_channels, int64_t rows_per_group,
                   float* sum, float* sumsq, void* stream) {
  MVAE_REQUIRE(x != nullptr && sum != nullptr, "col_stats: null tensor");
  RedArgs a = {};
  a.x = x; a.rows = rows; a.C = channels; a.c_valid = valid_channels > 0 ? valid_channels : channels;
  a.rows_per_group = rows_per_group > 0 ? rows_per_group : rows;
  a.a0 = sum; a.a1 = sumsq;
  return launch_col_reduce<RED_STATS>(dtype, a, static_cast<cudaStream_t>(stream));
}

static int bn_common(const mvae_bn_act_args* p, BnArgs& b) {
  MVAE_REQUIRE(p != nullptr, "bn_act: null args");
  const int vec = p->dtype == MVAE_F32 ? 4 : 8;
  MVAE_REQUIRE(p->channels > 0 && p->channels % vec == 0, "bn_act: %d channels must be a multiple of %d", p->channels, vec);
  MVAE_REQUIRE(p->rows > 0, "bn_act: empty input");
  b = BnArgs{};
  b.rows = p->rows; b.C = p->channels;
  b.rows_per_group = p->rows_per_group > 0 ? p->rows_per_group : p->rows;
  b.groups = static_cast<int>((b.rows + b.rows_per_group - 1) / b.rows_per_group);
  MVAE_REQUIRE(b.groups <= kMaxGroups, "bn_act: %d statistics groups (max %d)", b.groups, kMaxGroups);
  b.act = p->act; b.training = p->training;
  b.gamma = p->gamma; b.beta = p->beta;
  b.save_mean = p->save_mean; b.save_rstd = p->save_rstd;
  b.running_mean = p->running_mean; b.running_var = p->running_var;
  b.updates = p->updates_per_group; b.momentum = p->momentum; b.eps = p->eps;
  MVAE_REQUIRE(b.gamma != nullptr && b.beta != nullptr, "bn_act: gamma/beta required");
  return 0;
}

int mvae_bn_act_forward(const mvae_bn_act_args* p, void* stream) {
  BnArgs b;
  if (int rc = bn_common(p, b)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MVAE_REQUIRE(p->x != nullptr && p->y != nullptr, "bn_act_forward: x / y required");
  b.x = p->x; b.y = p->y;
  const size_t gc = static_cast<size_t>(b.groups) * b.C;
  if (p->training) {
    MVAE_REQUIRE(p->sum != nullptr && p->sumsq != nullptr, "bn_act_forward: sum / sumsq buffers required in training mode");
    if (!p->stats_ready) {
      MVAE_CUDA(cudaMemsetAsync(p->sum, 0, gc * sizeof(float), st));
      MVAE_CUDA(cudaMemsetAsync(p->sumsq, 0, gc * sizeof(float), st));
      if (int rc = mvae_col_stats(p->dtype, p->x, p->rows, p->channels, p->channels, b.rows_per_group, p->sum, p->sumsq, stream))
        return rc;
    }
    b.sum = p->sum; b.sumsq = p->sumsq;
  } else {
    MVAE_REQUIRE(p->running_mean != nullptr && p->running_var != nullptr, "bn_act_forward: eval mode needs running statistics");
  }
  const int vec = p->dtype == MVAE_F32 ? 4 : 8;
  const int blocks = grid_for(b.rows * (b.C / vec), 4);
  const size_t smem = 2 * gc * sizeof(float);
  MVAE_REQUIRE(smem <= 200 * 1024, "bn_act_forward: groups * channels = %zu too large for the coefficient table", gc);
  if (smem > 48 * 1024) {
    MVAE_CUDA(cudaFuncSetAttribute(bn_act_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MVAE_CUDA(cudaFuncSetAttribute(bn_act_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (p->dtype == MVAE_F32) bn_act_fwd_kernel<float><<<blocks, kThreads, smem, st>>>(b);
  else bn_act_fwd_kernel<__nv_bfloat16><<<blocks, kThreads, smem, st>>>(b);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_bn_act_backward(const mvae_bn_act_args* p, void* stream) {
  BnArgs b;
  if (int rc = bn_common(p, b)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MVAE_REQUIRE(p->x != nullptr && p->dy != nullptr && p->dx != nullptr, "bn_act_backward: x / dy / dx required");
  MVAE_REQUIRE(p->s0 != nullptr && p->s1 != nullptr && p->save_mean != nullptr && p->save_rstd != nullptr,
               "bn_act_backward: s0 / s1 scratch and the saved statistics are required");
  const size_t gc = static_cast<size_t>(b.groups) * b.C;
  MVAE_CUDA(cudaMemsetAsync(p->s0, 0, gc * sizeof(float), st));
  MVAE_CUDA(cudaMemsetAsync(p->s1, 0, gc * sizeof(float), st));
  RedArgs r = {};
  r.x = p->x; r.dy = p->dy; r.rows = p->rows; r.C = p->channels; r.c_valid = p->channels;
  r.rows_per_group = b.rows_per_group; r.act = p->act;
  r.a0 = p->s0; r.a1 = p->s1; r.mean = p->save_mean; r.rstd = p->save_rstd; r.gamma = p->gamma; r.beta = p->beta;
  if (int rc = launch_col_reduce<RED_BN_BWD>(p->dtype, r, st)) return rc;
  b.x = p->x; b.dy = p->dy; b.dx = p->dx; b.s0 = p->s0; b.s1 = p->s1;
  b.dgamma = p->dgamma; b.dbeta = p->dbeta;
  const int vec = p->dtype == MVAE_F32 ? 4 : 8;
  const int blocks = grid_for(b.rows * (b.C / vec), 4);
  const size_t smem = 5 * gc * sizeof(float);
  MVAE_REQUIRE(smem <= 200 * 1024, "bn_act_backward: groups * channels = %zu too large for the coefficient table", gc);
  if (smem > 48 * 1024) {
    MVAE_CUDA(cudaFuncSetAttribute(bn_act_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MVAE_CUDA(cudaFuncSetAttribute(bn_act_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (p->dtype == MVAE_F32) bn_act_bwd_kernel<float><<<blocks, kThreads, smem, st>>>(b);
  else bn_act_bwd_kernel<__nv_bfloat16><<<blocks, kThreads, smem, st>>>(b);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

static int dropout_params(float p, float* keep_scale, uint32_t* thresh) {
  MVAE_REQUIRE(p >= 0.f && p < 1.f, "dropout probability %f outside [0, 1)", p);
  if (p == 0.f) {
    *keep_scale = 0.f;
    *thresh = 65536u;
  } else {
    *thresh = static_cast<uint32_t>((1.0 - static_cast<double>(p)) * 65536.0 + 0.5);
    *keep_scale = 65536.0f / static_cast<float>(*thresh);   // exactly unbiased for the realised keep probability
  }
  return 0;
}

int mvae_act_forward(int dtype, int act, const void* x, void* y, int64_t rows, int channels, int repeat, float dropout_p,
                     uint64_t seed, const int* step_counter, void* stream) {
  const int vec = dtype == MVAE_F32 ? 4 : 8;
  MVAE_REQUIRE(x != nullptr && y != nullptr && rows > 0, "act_forward: null / empty tensor");
  MVAE_REQUIRE(channels % vec == 0, "act_forward: %d channels must be a multiple of %d", channels, vec);
  MVAE_REQUIRE(repeat >= 1, "act_forward: repeat >= 1");
  float ks; uint32_t th;
  if (int rc = dropout_params(dropout_p, &ks, &th)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = grid_for(rows * (channels / vec), 4);
  if (dtype == MVAE_F32)
    act_fwd_kernel<float><<<blocks, kThreads, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(y), rows, channels,
                                                       act, repeat, ks, th, seed, step_counter);
  else
    act_fwd_kernel<__nv_bfloat16><<<blocks, kThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                               static_cast<__nv_bfloat16*>(y), rows, channels, act, repeat,
                                                               ks, th, seed, step_counter);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_act_backward(int dtype, int act, const void* x, const void* dy, void* dx, int64_t rows, int channels, int repeat,
                      float dropout_p, uint64_t seed, const int* step_counter, float* dbias, void* stream) {
  MVAE_REQUIRE(x != nullptr && dy != nullptr && dx != nullptr && rows > 0, "act_backward: null / empty tensor");
  MVAE_REQUIRE(repeat >= 1, "act_backward: repeat >= 1");
  RedArgs r = {};
  if (int rc = dropout_params(dropout_p, &r.keep_scale, &r.keep_thresh)) return rc;
  r.x = x; r.dy = dy; r.dx = dx; r.rows = rows; r.C = channels; r.c_valid = channels; r.rows_per_group = rows;
  r.act = act; r.a0 = dbias; r.a1 = nullptr; r.repeat = repeat; r.seed = seed; r.step_ptr = step_counter;
  return launch_col_reduce<RED_ACT_BWD>(dtype, r, static_cast<cudaStream_t>(stream));
}

int mvae_sigmoid_bce(const mvae_sigmoid_bce_args* p, void* stream) {
  MVAE_REQUIRE(p != nullptr && p->logits != nullptr, "sigmoid_bce: logits required");
  MVAE_REQUIRE(p->rows > 0 && p->cols > 0, "sigmoid_bce: empty input");
  BceArgs a = {};
  a.logits = p->logits; a.ldl = p->ld_logits; a.logit_dtype = p->logit_dtype;
  a.target = p->target; a.ldt = p->ld_target; a.target_dtype = p->target_dtype;
  a.target_rows = p->target_rows > 0 ? p->target_rows : p->rows;
  a.rows = p->rows; a.cols = p->cols;
  a.rows_per_group = p->rows_per_group > 0 ? p->rows_per_group : p->rows;
  MVAE_REQUIRE((a.rows + a.rows_per_group - 1) / a.rows_per_group <= kMaxGroups, "sigmoid_bce: more than %d groups", kMaxGroups);
  for (int g = 0; g < kMaxGroups; ++g) a.scale[g] = p->grad_scale[g];
  a.loss = p->loss;
  a.probs = p->probs; a.ldp = p->ld_probs; a.prob_dtype = p->prob_dtype;
  a.dlogits = p->dlogits; a.ldd = p->ld_dlogits; a.grad_dtype = p->grad_dtype;
  a.ld_pad_zero = (p->dlogits != nullptr && p->ld_dlogits > p->cols) ? 1 : 0;
  a.dprobs = p->dprobs; a.lddp = p->ld_dprobs;
  MVAE_REQUIRE(a.ldl >= a.cols, "sigmoid_bce: ld_logits < cols");
  const long long items = a.rows * (a.ld_pad_zero ? a.ldd : a.cols);
  auto al16 = [](const void* q) { return q == nullptr || reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  const bool f32x4 = a.target != nullptr && a.logit_dtype == MVAE_F32 && a.target_dtype == MVAE_F32 && a.cols % 4 == 0 &&
                     a.ldl % 4 == 0 && a.ldt % 4 == 0 && !a.ld_pad_zero && al16(a.logits) && al16(a.target) &&
                     (a.dlogits == nullptr || (a.grad_dtype == MVAE_F32 && a.ldd % 4 == 0 && al16(a.dlogits))) &&
                     (a.probs == nullptr || (a.prob_dtype == MVAE_F32 && a.ldp % 4 == 0 && al16(a.probs)));
  if (f32x4)
    sigmoid_bce_f32x4_kernel<<<grid_for(items / 4, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  else
    sigmoid_bce_kernel<<<grid_for(items, 4), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

static int latent_fill(const mvae_latent_args* p, LatentArgs& a) {
  MVAE_REQUIRE(p != nullptr, "latent: null args");
  MVAE_REQUIRE(p->batch > 0 && p->n_latents > 0, "latent: empty input");
  MVAE_REQUIRE(p->n_terms >= 1 && p->n_terms <= kMaxGroups, "latent: n_terms %d outside [1, %d]", p->n_terms, kMaxGroups);
  a = LatentArgs{};
  a.B = p->batch; a.n = p->n_latents; a.G = p->n_terms;
  for (int g = 0; g < p->n_terms; ++g) {
    a.term_type[g] = p->term_type[g];
    a.a_row0[g] = p->expert_a_row0[g];
    a.kl_weight[g] = p->kl_weight[g];
    MVAE_REQUIRE(p->term_type[g] >= 0 && p->term_type[g] <= 2, "latent: bad term type %d", p->term_type[g]);
    MVAE_REQUIRE(p->term_type[g] == MVAE_TERM_TEXT || p->expert_a != nullptr , "latent: term %d needs expert A", g);
    MVAE_REQUIRE(p->term_type[g] == MVAE_TERM_IMAGE || p->expert_b != nullptr, "latent: term %d needs expert B", g);
  }
  a.poe_mode = p->poe_mode; a.prior = p->prior_expert; a.poe_eps = p->poe_eps;
  a.enc_a = p->expert_a; a.ld_a = p->ld_a; a.enc_b = p->expert_b; a.ld_b = p->ld_b;
  a.eps = p->eps; a.seed = p->seed; a.step_ptr = p->step_counter; a.training = p->training;
  a.z = p->z; a.ldz = p->ld_z; a.z_dtype = p->z_dtype; a.mu = p->mu; a.logvar = p->logvar; a.kl = p->kl;
  a.dz = p->dz; a.lddz = p->ld_dz; a.dz_dtype = p->dz_dtype; a.dmu_up = p->d_mu; a.dlogvar_up = p->d_logvar;
  a.d_enc_a = p->d_expert_a; a.ld_da = p->ld_da; a.d_enc_b = p->d_expert_b; a.ld_db = p->ld_db; a.d_dtype = p->d_dtype;
  a.row_w = p->row_weight;
  return 0;
}

int mvae_latent_forward(const mvae_latent_args* p, void* stream) {
  LatentArgs a;
  if (int rc = latent_fill(p, a)) return rc;
  MVAE_REQUIRE(a.z != nullptr && a.ldz >= a.n, "latent_forward: z required with ld_z >= n_latents");
  latent_kernel<false><<<grid_for(a.B * a.n, 4), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}
int mvae_latent_backward(const mvae_latent_args* p, void* stream) {
  LatentArgs a;
  if (int rc = latent_fill(p, a)) return rc;
  MVAE_REQUIRE(a.d_enc_a != nullptr || a.d_enc_b != nullptr, "latent_backward: no gradient output");
  latent_kernel<true><<<grid_for(a.B * a.n, 4), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

// Per-sample presence masks -> per-(term, row) weights batch / count (one block; see include/mvae_b200.h).
__global__ void __launch_bounds__(1024) mask_weights_kernel(const uint8_t* __restrict__ hi, const uint8_t* __restrict__ ht, long long B,
                                                            int G, int t0, int t1, int t2, float* __restrict__ w,
                                                            float* __restrict__ counts) {
  __shared__ int s_cnt[kMaxGroups];
  if (threadIdx.x < kMaxGroups) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int ty[kMaxGroups] = {t0, t1, t2};
  int cnt[kMaxGroups] = {0, 0, 0};
  for (long long b = threadIdx.x; b < B; b += blockDim.x) {
    const bool i = hi[b] != 0, t = ht[b] != 0;
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g)
      if (g < G) cnt[g] += (ty[g] == MVAE_TERM_JOINT ? (i && t) : ty[g] == MVAE_TERM_IMAGE ? i : t) ? 1 : 0;
  }
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    int v = cnt[g];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0) atomicAdd(&s_cnt[g], v);
  }
  __syncthreads();
  for (long long b = threadIdx.x; b < B; b += blockDim.x) {
    const bool i = hi[b] != 0, t = ht[b] != 0;
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g)
      if (g < G) {
        const bool on = ty[g] == MVAE_TERM_JOINT ? (i && t) : ty[g] == MVAE_TERM_IMAGE ? i : t;
        w[static_cast<long long>(g) * B + b] = (on && s_cnt[g] > 0) ? static_cast<float>(B) / static_cast<float>(s_cnt[g]) : 0.f;
      }
  }
  if (counts != nullptr && threadIdx.x < G) counts[threadIdx.x] = static_cast<float>(s_cnt[threadIdx.x]);
}

int mvae_mask_weights(const uint8_t* has_image, const uint8_t* has_text, int64_t batch, int n_terms, const int* term_type,
                      float* weight, float* counts, void* stream) {
  MVAE_REQUIRE(has_image != nullptr && has_text != nullptr && weight != nullptr && term_type != nullptr, "mask_weights: null argument");
  MVAE_REQUIRE(batch > 0 && n_terms >= 1 && n_terms <= kMaxGroups, "mask_weights: batch / n_terms out of range");
  int t[kMaxGroups] = {0, 0, 0};
  for (int g = 0; g < n_terms; ++g) {
    MVAE_REQUIRE(term_type[g] >= 0 && term_type[g] <= 2, "mask_weights: bad term type %d", term_type[g]);
    t[g] = term_type[g];
  }
  mask_weights_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(has_image, has_text, batch, n_terms, t[0], t[1], t[2], weight, counts);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_cast_pad_2d(const float* src, int64_t rows, int64_t cols, int64_t ld_src, int dst_dtype, void* dst, int64_t ld_dst,
                     void* stream) {
  MVAE_REQUIRE(src != nullptr && dst != nullptr && rows > 0 && cols > 0, "cast_pad_2d: null / empty tensor");
  MVAE_REQUIRE(ld_src >= cols && ld_dst >= cols, "cast_pad_2d: leading dimension smaller than the row length");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = grid_for(rows * ld_dst, 4);
  if (dst_dtype == MVAE_F32) cast_pad_kernel<float><<<blocks, kThreads, 0, st>>>(src, rows, cols, ld_src, static_cast<float*>(dst), ld_dst);
  else cast_pad_kernel<__nv_bfloat16><<<blocks, kThreads, 0, st>>>(src, rows, cols, ld_src, static_cast<__nv_bfloat16*>(dst), ld_dst);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_step_begin(int* step_counter, float* zero_buf, int64_t zero_floats, int64_t* counters, const int64_t* increments,
                    int n_counters, void* stream) {
  MVAE_REQUIRE(zero_floats % 4 == 0, "step_begin: zero_floats must be a multiple of 4");
  MVAE_REQUIRE(n_counters >= 0 && n_counters <= kThreads, "step_begin: at most %d counters", kThreads);
  const long long n4 = zero_floats / 4;
  step_begin_kernel<<<grid_for(std::max<long long>(n4, 1), 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      step_counter, reinterpret_cast<float4*>(zero_buf), zero_buf != nullptr ? n4 : 0,
      reinterpret_cast<long long*>(counters), reinterpret_cast<const long long*>(increments), n_counters);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

}  // extern "C"
