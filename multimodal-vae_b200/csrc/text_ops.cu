// Operator-level kernels of the MultiMNIST text path (multimnist/model.py:220-307): character embedding lookups,
// GRU cell gate math (the two projections of every cell are tcgen05 GEMMs, gemm.cu), log_softmax + NLL + greedy
// argmax of the autoregressive text decoder, and two small 2-D helpers.  All tensors are row-major matrices with an
// explicit leading dimension so that concatenations ([embedding | z], [hidden | z]) are just column ranges.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "../../include/mvae_b200.h"

namespace mvae {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxVocab = 32;

inline int grid_for(long long items, int per_sm = 4) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  long long b = (items + kThreads - 1) / kThreads;
  const long long cap = static_cast<long long>(per_sm) * sms;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

__device__ __forceinline__ float ld_any(const void* p, int dtype, long long i) {
  return dtype == MVAE_F32 ? static_cast<const float*>(p)[i] : __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dtype, long long i, float v) {
  if (dtype == MVAE_F32) static_cast<float*>(p)[i] = v; else static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float act_f(int act, float u) {
  if (act == MVAE_ACT_RELU) return fmaxf(u, 0.f);
  if (act == MVAE_ACT_SWISH) return u * sigmoidf(u);
  return u;
}
__device__ __forceinline__ float act_g(int act, float u) {
  if (act == MVAE_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (act == MVAE_ACT_SWISH) {
    const float s = sigmoidf(u);
    return s * (1.f + u * (1.f - s));
  }
  return 1.f;
}

// ---------------------------------------------------------------- embedding
// out[m, j] = act(table[idx[m * idx_stride], j]),  j < width   (nn.Embedding, optionally followed by Swish)
__global__ void __launch_bounds__(kThreads) embed_fwd_kernel(const long long* __restrict__ idx, long long idx_stride,
                                                             const float* __restrict__ table, int vocab, int width, int act,
                                                             void* out, int out_dtype, long long ldo, long long rows) {
  const long long total = rows * width;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int j = static_cast<int>(i % width);
    const long long m = i / width;
    long long c = idx[m * idx_stride];
    c = c < 0 ? 0 : (c >= vocab ? vocab - 1 : c);
    st_any(out, out_dtype, m * ldo + j, act_f(act, table[c * width + j]));
  }
}
// dtable[idx[m], j] += dout[m, j] * act'(table[idx[m], j]); a block first sums its rows per character in shared memory.
// blockIdx.y selects a tile of up to kThreads columns (tables wider than a block).
__global__ void __launch_bounds__(kThreads) embed_bwd_kernel(const long long* __restrict__ idx, long long idx_stride,
                                                             const float* __restrict__ table, int vocab, int full_width, int act,
                                                             const void* dout, int dout_dtype, long long ldd, long long rows,
                                                             float* __restrict__ dtable) {
  extern __shared__ float s_acc[];  // [lanes][vocab, width]
  const int col0 = blockIdx.y * kThreads;
  const int width = min(kThreads, full_width - col0);
  const int lanes = max(1, min(kThreads / width, 4));   // row lanes per column
  const int vw = vocab * width;
  for (int i = threadIdx.x; i < lanes * vw; i += kThreads) s_acc[i] = 0.f;
  __syncthreads();
  const long long slab = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * slab, r1 = min(rows, r0 + slab);
  const int lane = threadIdx.x / width, j = threadIdx.x - lane * width;
  if (lane < lanes) {
    float* mine = s_acc + lane * vw;     // (lane, column) pairs own disjoint addresses: no atomics needed
    for (long long m = r0 + lane; m < r1; m += lanes) {
      long long c = idx[m * idx_stride];
      c = c < 0 ? 0 : (c >= vocab ? vocab - 1 : c);
      mine[c * width + j] += ld_any(dout, dout_dtype, m * ldd + col0 + j) * act_g(act, table[c * full_width + col0 + j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < vw; i += kThreads) {
    float v = 0.f;
    for (int l = 0; l < lanes; ++l) v += s_acc[l * vw + i];
    const int c = i / width, jj = i - c * width;
    if (v != 0.f) atomicAdd(dtable + static_cast<long long>(c) * full_width + col0 + jj, v);
  }
}

// ---------------------------------------------------------------- GRU cell (gate order r, z, n like nn.GRU)
struct GruArgs {
  long long rows; int H;
  const float* gi; long long ldgi;    // [rows, 3H] = x W_ih^T + b_ih
  const float* gh; long long ldgh;    // [rows, 3H] = h W_hh^T + b_hh
  const void* h_prev; int h_dtype; long long ldh;   // [rows, H] or null (zeros)
  const void* addend; long long ld_add;             // optional (activation dtype): out1 = h_new + addend
  void* out1; long long ld1;          // h_new (activation dtype)
  void* out2; long long ld2;          // optional second copy (e.g. the [hidden | z] concatenation)
  float* saved;                       // [rows, 4H]: r, z, n, hn (= gh_n) for the backward
  // backward
  const void* dh_a; int dha_dtype; long long ld_dha;   // gradient at h_new, summed from up to two sources
  const void* dh_b; int dhb_dtype; long long ld_dhb;
  void* dgi; void* dgh; long long ldg; int g_dtype;   // [rows, ldg] gradients at gi / gh (columns [3H, ldg) zeroed)
  float* dh_prev; long long ld_dhp;                   // direct part dh * z (stored, not accumulated)
};

__global__ void __launch_bounds__(kThreads) gru_fwd_kernel(const GruArgs a) {
  const int H = a.H;
  const long long total = a.rows * H;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int j = static_cast<int>(i % H);
    const long long m = i / H;
    const float* gi = a.gi + m * a.ldgi;
    const float* gh = a.gh + m * a.ldgh;
    const float r = sigmoidf(gi[j] + gh[j]);
    const float z = sigmoidf(gi[H + j] + gh[H + j]);
    const float hn = gh[2 * H + j];
    const float n = tanhf(gi[2 * H + j] + r * hn);
    const float hp = a.h_prev != nullptr ? ld_any(a.h_prev, a.h_dtype, m * a.ldh + j) : 0.f;
    float h = (1.f - z) * n + z * hp;
    if (a.saved != nullptr) {
      float* s = a.saved + m * 4 * H;
      s[j] = r; s[H + j] = z; s[2 * H + j] = n; s[3 * H + j] = hn;
    }
    if (a.addend != nullptr) h += ld_any(a.addend, a.h_dtype, m * a.ld_add + j);
    st_any(a.out1, a.h_dtype, m * a.ld1 + j, h);
    if (a.out2 != nullptr) st_any(a.out2, a.h_dtype, m * a.ld2 + j, h);
  }
}

__global__ void __launch_bounds__(kThreads) gru_bwd_kernel(const GruArgs a) {
  const int H = a.H;
  const long long ldg = a.ldg;
  const long long total = a.rows * ldg;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int j = static_cast<int>(i % ldg);
    const long long m = i / ldg;
    if (j >= H) {
      if (j >= 3 * H) {   // padding columns of the GEMM operands
        st_any(a.dgi, a.g_dtype, m * ldg + j, 0.f);
        st_any(a.dgh, a.g_dtype, m * ldg + j, 0.f);
      }
      continue;
    }
    float dh = 0.f;
    if (a.dh_a != nullptr) dh += ld_any(a.dh_a, a.dha_dtype, m * a.ld_dha + j);
    if (a.dh_b != nullptr) dh += ld_any(a.dh_b, a.dhb_dtype, m * a.ld_dhb + j);
    const float* s = a.saved + m * 4 * H;
    const float r = s[j], z = s[H + j], n = s[2 * H + j], hn = s[3 * H + j];
    const float hp = a.h_prev != nullptr ? ld_any(a.h_prev, a.h_dtype, m * a.ldh + j) : 0.f;
    const float dn_pre = dh * (1.f - z) * (1.f - n * n);
    const float dz_pre = dh * (hp - n) * z * (1.f - z);
    const float dr_pre = dn_pre * hn * r * (1.f - r);
    st_any(a.dgi, a.g_dtype, m * ldg + j, dr_pre);
    st_any(a.dgi, a.g_dtype, m * ldg + H + j, dz_pre);
    st_any(a.dgi, a.g_dtype, m * ldg + 2 * H + j, dn_pre);
    st_any(a.dgh, a.g_dtype, m * ldg + j, dr_pre);
    st_any(a.dgh, a.g_dtype, m * ldg + H + j, dz_pre);
    st_any(a.dgh, a.g_dtype, m * ldg + 2 * H + j, dn_pre * r);
    if (a.dh_prev != nullptr) a.dh_prev[m * a.ld_dhp + j] = dh * z;
  }
}

// ---------------------------------------------------------------- log_softmax + NLL + argmax (one thread per row)
struct LsmArgs {
  long long rows; int classes; long long rows_per_group;
  const float* logits; long long ldl;
  const long long* target; long long target_stride; long long target_rows;   // target of row m: target[(m % target_rows) * stride]
  float scale[kMaxGroups];
  float* loss;                       // [groups] += -logp[target]  (unscaled)
  float* logp; long long ldp;        // optional output
  long long* argmax;                 // optional [rows]
  void* dlogits; int g_dtype; long long ldd;   // scale[g] * (softmax - onehot), columns [classes, ldd) zeroed
  const float* row_w;                // optional [rows]: weight of row m's loss and gradient
};
__global__ void __launch_bounds__(kThreads) logsoftmax_nll_kernel(const LsmArgs a) {
  __shared__ float s_loss[kMaxGroups];
  if (threadIdx.x < kMaxGroups) s_loss[threadIdx.x] = 0.f;
  __syncthreads();
  float acc[kMaxGroups] = {0.f, 0.f, 0.f};
  for (long long m = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; m < a.rows;
       m += static_cast<long long>(gridDim.x) * kThreads) {
    const float* x = a.logits + m * a.ldl;
    float mx = x[0];
    int am = 0;
    for (int c = 1; c < a.classes; ++c)
      if (x[c] > mx) { mx = x[c]; am = c; }       // first maximum, like torch.max
    float se = 0.f;
    for (int c = 0; c < a.classes; ++c) se += expf(x[c] - mx);
    const float lse = mx + logf(se);
    if (a.argmax != nullptr) a.argmax[m] = am;
    if (a.logp != nullptr)
      for (int c = 0; c < a.classes; ++c) a.logp[m * a.ldp + c] = x[c] - lse;
    if (a.target != nullptr) {
      const int g = static_cast<int>(m / a.rows_per_group);
      long long t = a.target[(m % a.target_rows) * a.target_stride];
      t = t < 0 ? 0 : (t >= a.classes ? a.classes - 1 : t);
      const float rw = a.row_w != nullptr ? a.row_w[m] : 1.f;
      const float l = rw * (lse - x[t]);
      if (g == 0) acc[0] += l; else if (g == 1) acc[1] += l; else acc[2] += l;
      if (a.dlogits != nullptr) {
        const float sc = a.scale[g] * rw;
        for (int c = 0; c < a.classes; ++c)
          st_any(a.dlogits, a.g_dtype, m * a.ldd + c, sc * (expf(x[c] - lse) - (c == t ? 1.f : 0.f)));
        for (long long c = a.classes; c < a.ldd; ++c) st_any(a.dlogits, a.g_dtype, m * a.ldd + c, 0.f);
      }
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

// dlogits = dlogp - softmax * sum(dlogp): backward of F.log_softmax given an upstream gradient of the log-probabilities
__global__ void __launch_bounds__(kThreads) logsoftmax_bwd_kernel(const float* __restrict__ logp, long long ldp,
                                                                  const float* __restrict__ dlogp, long long ldg, long long rows,
                                                                  int classes, void* dlogits, int g_dtype, long long ldd) {
  for (long long m = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; m < rows;
       m += static_cast<long long>(gridDim.x) * kThreads) {
    float s = 0.f;
    for (int c = 0; c < classes; ++c) s += dlogp[m * ldg + c];
    for (int c = 0; c < classes; ++c) st_any(dlogits, g_dtype, m * ldd + c, dlogp[m * ldg + c] - expf(logp[m * ldp + c]) * s);
    for (long long c = classes; c < ldd; ++c) st_any(dlogits, g_dtype, m * ldd + c, 0.f);
  }
}

// ---------------------------------------------------------------- 2-D copy / add with leading dimensions
__global__ void __launch_bounds__(kThreads) copy2d_kernel(const void* src, int src_dtype, long long ld_src, void* dst,
                                                          int dst_dtype, long long ld_dst, long long rows, long long cols,
                                                          int accumulate, const void* src2, int src2_dtype, long long ld_src2) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const long long c = i % cols, r = i / cols;
    float v = ld_any(src, src_dtype, r * ld_src + c);
    if (src2 != nullptr) v += ld_any(src2, src2_dtype, r * ld_src2 + c);
    if (accumulate) v += ld_any(dst, dst_dtype, r * ld_dst + c);
    st_any(dst, dst_dtype, r * ld_dst + c, v);
  }
}

}  // namespace

}  // namespace mvae

using namespace mvae;

extern "C" {

int mvae_embed_forward(const int64_t* indices, int64_t index_stride, const float* table, int vocab, int width, int act,
                       int out_dtype, void* out, int64_t ld_out, int64_t rows, void* stream) {
  MVAE_REQUIRE(indices != nullptr && table != nullptr && out != nullptr, "embed_forward: null tensor");
  MVAE_REQUIRE(rows > 0 && width > 0 && vocab > 0 && ld_out >= width, "embed_forward: bad shape");
  embed_fwd_kernel<<<grid_for(rows * width), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(indices), index_stride, table, vocab, width, act, out, out_dtype, ld_out, rows);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_embed_backward(const int64_t* indices, int64_t index_stride, const float* table, int vocab, int width, int act,
                        int dout_dtype, const void* dout, int64_t ld_dout, int64_t rows, float* dtable, void* stream) {
  MVAE_REQUIRE(indices != nullptr && table != nullptr && dout != nullptr && dtable != nullptr, "embed_backward: null tensor");
  MVAE_REQUIRE(rows > 0 && width > 0 && vocab > 0 && vocab <= kMaxVocab, "embed_backward: vocabulary %d outside [1, %d]", vocab, kMaxVocab);
  const int tile = std::min(width, kThreads);           // columns per block (blockIdx.y walks the tiles of wider tables)
  const int lanes = std::max(1, std::min(kThreads / tile, 4));
  const size_t smem = static_cast<size_t>(lanes) * vocab * tile * sizeof(float);
  MVAE_REQUIRE(smem <= 48 * 1024, "embed_backward: table too large for shared memory");
  const int blocks = static_cast<int>(std::min<long long>((rows + 15) / 16, 592));
  embed_bwd_kernel<<<dim3(blocks, (width + kThreads - 1) / kThreads), kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(indices), index_stride, table, vocab, width, act, dout, dout_dtype, ld_dout, rows, dtable);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

static int gru_fill(const mvae_gru_cell_args* p, GruArgs& a) {
  MVAE_REQUIRE(p != nullptr, "gru_cell: null args");
  MVAE_REQUIRE(p->rows > 0 && p->hidden > 0, "gru_cell: empty input");
  a = GruArgs{};
  a.rows = p->rows; a.H = p->hidden;
  a.gi = p->gi; a.ldgi = p->ld_gi; a.gh = p->gh; a.ldgh = p->ld_gh;
  a.h_prev = p->h_prev; a.h_dtype = p->h_dtype; a.ldh = p->ld_h_prev;
  a.addend = p->addend; a.ld_add = p->ld_addend;
  a.out1 = p->h_out; a.ld1 = p->ld_h_out; a.out2 = p->h_out2; a.ld2 = p->ld_h_out2;
  a.saved = p->saved;
  a.dh_a = p->dh_a; a.dha_dtype = p->dh_a_dtype; a.ld_dha = p->ld_dh_a;
  a.dh_b = p->dh_b; a.dhb_dtype = p->dh_b_dtype; a.ld_dhb = p->ld_dh_b;
  a.dgi = p->dgi; a.dgh = p->dgh; a.ldg = p->ld_dg; a.g_dtype = p->dg_dtype;
  a.dh_prev = p->dh_prev; a.ld_dhp = p->ld_dh_prev;
  return 0;
}

int mvae_gru_cell_forward(const mvae_gru_cell_args* p, void* stream) {
  GruArgs a;
  if (int rc = gru_fill(p, a)) return rc;
  MVAE_REQUIRE(a.gi != nullptr && a.gh != nullptr && a.out1 != nullptr, "gru_cell_forward: gi / gh / h_out required");
  MVAE_REQUIRE(a.ldgi >= 3 * a.H && a.ldgh >= 3 * a.H && a.ld1 >= a.H, "gru_cell_forward: leading dimension too small");
  gru_fwd_kernel<<<grid_for(a.rows * a.H), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_gru_cell_backward(const mvae_gru_cell_args* p, void* stream) {
  GruArgs a;
  if (int rc = gru_fill(p, a)) return rc;
  MVAE_REQUIRE(a.saved != nullptr && a.dgi != nullptr && a.dgh != nullptr, "gru_cell_backward: saved gates / dgi / dgh required");
  MVAE_REQUIRE(a.dh_a != nullptr || a.dh_b != nullptr, "gru_cell_backward: no upstream gradient");
  MVAE_REQUIRE(a.ldg >= 3 * a.H, "gru_cell_backward: ld_dg < 3 * hidden");
  gru_bwd_kernel<<<grid_for(a.rows * a.ldg), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_logsoftmax_nll(const mvae_logsoftmax_nll_args* p, void* stream) {
  MVAE_REQUIRE(p != nullptr && p->logits != nullptr, "logsoftmax_nll: logits required");
  MVAE_REQUIRE(p->rows > 0 && p->classes > 0 && p->classes <= 64 && p->ld_logits >= p->classes, "logsoftmax_nll: bad shape");
  LsmArgs a = {};
  a.rows = p->rows; a.classes = p->classes;
  a.rows_per_group = p->rows_per_group > 0 ? p->rows_per_group : p->rows;
  MVAE_REQUIRE((a.rows + a.rows_per_group - 1) / a.rows_per_group <= kMaxGroups, "logsoftmax_nll: more than %d groups", kMaxGroups);
  a.logits = p->logits; a.ldl = p->ld_logits;
  a.target = reinterpret_cast<const long long*>(p->target); a.target_stride = p->target_stride > 0 ? p->target_stride : 1;
  a.target_rows = p->target_rows > 0 ? p->target_rows : p->rows;
  for (int g = 0; g < kMaxGroups; ++g) a.scale[g] = p->grad_scale[g];
  a.loss = p->loss; a.logp = p->logp; a.ldp = p->ld_logp;
  a.argmax = reinterpret_cast<long long*>(p->argmax);
  a.dlogits = p->dlogits; a.g_dtype = p->grad_dtype; a.ldd = p->ld_dlogits;
  a.row_w = p->row_weight;
  MVAE_REQUIRE(a.dlogits == nullptr || a.ldd >= a.classes, "logsoftmax_nll: ld_dlogits < classes");
  logsoftmax_nll_kernel<<<grid_for(a.rows), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_logsoftmax_backward(const float* logp, int64_t ld_logp, const float* dlogp, int64_t ld_dlogp, int64_t rows, int classes,
                            int grad_dtype, void* dlogits, int64_t ld_dlogits, void* stream) {
  MVAE_REQUIRE(logp != nullptr && dlogp != nullptr && dlogits != nullptr, "logsoftmax_backward: null tensor");
  MVAE_REQUIRE(rows > 0 && classes > 0 && ld_logp >= classes && ld_dlogp >= classes && ld_dlogits >= classes,
               "logsoftmax_backward: bad shape");
  logsoftmax_bwd_kernel<<<grid_for(rows), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(logp, ld_logp, dlogp, ld_dlogp, rows,
                                                                                          classes, dlogits, grad_dtype, ld_dlogits);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int mvae_copy_2d(int src_dtype, const void* src, int64_t ld_src, int dst_dtype, void* dst, int64_t ld_dst, int64_t rows,
                 int64_t cols, int accumulate, int src2_dtype, const void* src2, int64_t ld_src2, void* stream) {
  MVAE_REQUIRE(src != nullptr && dst != nullptr && rows > 0 && cols > 0, "copy_2d: null / empty tensor");
  MVAE_REQUIRE(ld_src >= cols && ld_dst >= cols && (src2 == nullptr || ld_src2 >= cols), "copy_2d: leading dimension too small");
  copy2d_kernel<<<grid_for(rows * cols), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      src, src_dtype, ld_src, dst, dst_dtype, ld_dst, rows, cols, accumulate, src2, src2_dtype, ld_src2);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

}  // extern "C"
