// tcgen05 / TMEM / TMA GEMM for the MVAE Linear layers (forward, dgrad, wgrad) with the
// BatchNorm-statistics, BCE-with-logits and ReLU/BN-backward epilogues fused in.
//
//   C[M,N] = A[M,K] * B[N,K]^T        (fp32 accumulate in TMEM)
//
// One CTA = one 128 x block_n output tile (UMMA M=128, N=block_n, cta_group::1), 8 warps:
//   warp 0   : TMA producer (one elected lane) - A/B tiles into a SWIZZLE_128B smem ring
//   warp 1   : MMA issuer   (one elected lane) - tcgen05.mma kind::tf32 / kind::f16(bf16)
//   warp 2   : TMEM allocation
//   all 8    : epilogue - tcgen05.ld TMEM -> registers -> padded smem tile -> row pass with
//              coalesced global I/O and per-column statistics.
// Both operands may be K-major (contraction contiguous: forward) or MN-major (row index
// contiguous: dgrad's weight, wgrad's activations/gradients), so no transposed copies of
// activations or weights are ever materialised.  Several CTAs co-reside per SM (smem permitting)
// which is what hides the prologue/epilogue of these small, latency-bound problems.
//
// Reference semantics: nn.Linear (mnist/model.py:104-110,124-130), the BCE of
// mnist/train.py:70 on the sigmoid of mnist/model.py:135, BatchNorm1d+ReLU backward.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace mvae {

namespace {

constexpr int kBlockM = 128;
constexpr int kAStageBytes = kBlockM * 128;  // 16 KB: 128 rows x 128 B (either major)
constexpr int kGemmThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kStagePad = 4;  // staging row stride = block_n + 4 words -> conflict-free float4 rows

struct GemmKParams {
  int M, N, K;
  int block_n;
  int a_mn, b_mn;
  int stages;
  int kb_total;
  int kb_per_split;
  int b_stage_bytes;  // smem bytes reserved per stage for B (multiple of 1024)
  int b_tx_bytes;     // bytes TMA actually writes per stage for B
  int vec_ok;         // all epilogue tensors allow 4-element vector access
  int direct_store;   // plain STORE epilogue without statistics: TMEM -> registers -> global, no staging pass
  int aux_off;        // byte offset (dynamic smem) of the prefetched auxiliary tile, or -1
  int red_off;        // byte offset (dynamic smem) of the [2][8][block_n] column-statistics scratch, or -1
  int stat_group_stride;
  long long* dbg;  // optional per-CTA timestamps (MVAE_GEMM_DEBUG_TIMES), 8 slots per CTA
  GemmEpilogue epi;
  ConvGather gather;  // implicit patch-matrix operand (mode 0: none)
};

template <int kKind>
struct ActT {
  using type = float;
};
template <>
struct ActT<MVAE_BF16> {
  using type = __nv_bfloat16;
};

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }

// 4 contiguous elements starting at p (n_valid of them in range).
template <typename T>
__device__ __forceinline__ void load4(const T* p, bool vec, int n_valid, float (&o)[4]) {
  if (vec && n_valid >= 4) {
    // read-only path (ld.global.nc): lets the compiler hoist these loads above the epilogue's global stores
    if constexpr (sizeof(T) == 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p));
      o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
      uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
      __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
      __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
      o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = (i < n_valid) ? to_f(p[i]) : 0.f;
  }
}
__device__ __forceinline__ void store4(float* p, bool vec, int n_valid, const float (&v)[4]) {
  if (vec && n_valid >= 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < n_valid) p[i] = v[i];
  }
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, bool vec, int n_valid, const float (&v)[4]) {
  if (vec && n_valid >= 4) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < n_valid) p[i] = __float2bfloat16_rn(v[i]);
  }
}
// One thread's share of the epilogue row pass: `rows` consecutive rows starting at global row m, 4 consecutive
// columns starting at global column cn (nv of them valid), values read from the fp32 staging tile at saddr.
// kFast = all 4 columns valid and every tensor is vector-aligned: the loop body is branch-free so that the
// compiler can software-pipeline the shared loads over the unrolled rows.
template <int kKind, int kEpi, typename CT, bool kFast>
__device__ __forceinline__ void epilogue_rows(const GemmKParams& p, uint32_t saddr, int ldst, int m, int rows, int cn,
                                              int nv, int lane, float (&loss_acc)[4], uint32_t aux_saddr,
                                              int aux_ld_bytes, float* s_red_col) {
  using act_t = typename ActT<kKind>::type;
  const GemmEpilogue& e = p.epi;
  float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
  float bias[4], g_mean[4], g_rstd[4], g_gamma[4], g_beta[4], g_a[4] = {0.f, 0.f, 0.f, 0.f}, g_b[4] = {0.f, 0.f, 0.f, 0.f};
  float lsum = 0.f, g_scale = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bias[i] = (e.bias != nullptr && i < nv) ? e.bias[cn + i] : 0.f;
    g_mean[i] = g_rstd[i] = g_gamma[i] = g_beta[i] = 0.f;
    if (kEpi == EPI_DGRAD_BN && i < nv) {
      g_gamma[i] = e.bn_gamma[cn + i];
      g_beta[i] = e.bn_beta[cn + i];
    }
  }
  int g = m / e.rows_per_group;
  int seg_left = static_cast<int>(min(static_cast<long long>(g + 1) * e.rows_per_group - m, 1ll << 20));
  int trow = (kEpi == EPI_BCE) ? (m % e.target_rows) : 0;
  CT* cptr = reinterpret_cast<CT*>(e.C) + static_cast<long long>(m) * e.ldc + cn;
  CT* pptr = ((kEpi == EPI_BCE || kEpi == EPI_STORE_ACT) && e.probs != nullptr)
                 ? reinterpret_cast<CT*>(e.probs) + static_cast<long long>(m) * e.ldc + cn
                 : nullptr;
  const act_t* hptr = (kEpi == EPI_DGRAD_BN || kEpi == EPI_DGRAD_ACT)
                          ? reinterpret_cast<const act_t*>(e.hpre) + static_cast<long long>(m) * e.ldh + cn
                          : nullptr;
  int done = 0;
  while (done < rows) {
    const int seg = min(rows - done, seg_left);
    if (kEpi == EPI_BCE) g_scale = e.bce_scale[g & 3];
    if (kEpi == EPI_DGRAD_BN) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < nv) {
          g_mean[i] = e.bn_mean[static_cast<long long>(g) * p.N + cn + i];
          g_rstd[i] = e.bn_rstd[static_cast<long long>(g) * p.N + cn + i];
          g_a[i] = g_gamma[i] * g_rstd[i];
          g_b[i] = fmaf(-g_mean[i], g_a[i], g_beta[i]);
        }
    }
    // Rows in batches of kRB: issue every load of the batch (staging tile + auxiliary tensor) first, then the
    // math and the stores - with 8 warps per CTA the latency has to be hidden by ILP, not by occupancy.
    constexpr int kRB = 4;
    for (int it = 0; it < seg; it += kRB) {
      float v[kRB][4], aux[kRB][4];
#pragma unroll
      for (int u = 0; u < kRB; ++u) {
        if (it + u < seg) {
          ptx::lds128(saddr + u * ldst * 4, v[u]);
          if constexpr (kEpi == EPI_BCE || kEpi == EPI_DGRAD_BN || kEpi == EPI_DGRAD_ACT) {
            if (kFast && aux_saddr != 0) {
              ptx::lds_act4<act_t>(aux_saddr + (it + u) * aux_ld_bytes, aux[u]);  // prefetched during the main loop
            } else if constexpr (kEpi == EPI_BCE) {
              int tr_u = trow + u;
              if (tr_u >= e.target_rows) tr_u -= e.target_rows;
              load4(reinterpret_cast<const act_t*>(e.target) + static_cast<long long>(tr_u) * e.ldt + cn, kFast, nv, aux[u]);
            } else {
              load4(hptr + static_cast<long long>(u) * e.ldh, kFast, nv, aux[u]);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kRB; ++u) {
        if (it + u >= seg) break;
        CT* crow = cptr + static_cast<long long>(u) * e.ldc;
        if constexpr (kEpi == EPI_STORE) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[u][i] += bias[i];
            acc0[i] += v[u][i];
            acc1[i] = fmaf(v[u][i], v[u][i], acc1[i]);
          }
          store4(crow, kFast, nv, v[u]);
        } else if constexpr (kEpi == EPI_ATOMIC) {
          if constexpr (kFast && sizeof(CT) == 4) {
            ptx::red_add_v4(reinterpret_cast<float*>(crow), v[u]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < nv) atomicAdd(reinterpret_cast<float*>(crow) + i, v[u][i]);
          }
        } else if constexpr (kEpi == EPI_BCE) {
          // BCE on logits: loss = softplus(x) - t*x, d/dx = sigmoid(x) - t  (reference: sigmoid then
          // F.binary_cross_entropy, mnist/model.py:135 + mnist/train.py:70; identical for |x| < ~17).
          float d[4], pr[4];
          const float rw = e.row_w != nullptr ? __ldg(e.row_w + m + done + it + u) : 1.f;
          const float gs = g_scale * rw;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x = v[u][i] + bias[i];
            const float tg = aux[u][i];
            const float ex = ptx::ex2_approx(-1.4426950408889634f * fabsf(x));  // exp(-|x|) in (0, 1]
            // sigmoid(|x|) = 1/(1+ex) in [0.5, 1) on the FMA pipe (the SFU is the bottleneck of this pass):
            // linear seed on d in [1,2] (max rel. error 1/17) + three Newton steps r <- r*(2 - d*r)  -> 1e-9
            const float dd = 1.f + ex;
            float inv = fmaf(-0.47058823529f, dd, 1.41176470588f);
            inv = inv * fmaf(-dd, inv, 2.f);
            inv = inv * fmaf(-dd, inv, 2.f);
            inv = inv * fmaf(-dd, inv, 2.f);
            const float pz = x >= 0.f ? inv : ex * inv;
            pr[i] = pz;
            d[i] = gs * (pz - tg);
            if (kFast || i < nv) {
              // softplus(x) - t*x = max(x,0) - t*x + log(1+exp(-|x|)),  log(1+exp(-|x|)) = -ln(inv)
              const float sp = fmaf(-0.6931471805599453f, ptx::lg2_approx(inv), fmaf(-tg, x, fmaxf(x, 0.f)));
              lsum = fmaf(rw, sp, lsum);
              acc0[i] += d[i];
            }
          }
          store4(crow, kFast, nv, d);
          if (pptr != nullptr) store4(pptr + static_cast<long long>(u) * e.ldc, kFast, nv, pr);
        } else if constexpr (kEpi == EPI_DGRAD_BN) {
          // dyhat = dh * 1[relu input > 0]; the relu input is recomputed with the forward's own expression.
          float d[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float xh = (aux[u][i] - g_mean[i]) * g_rstd[i];
            const float y = fmaf(g_a[i], aux[u][i], g_b[i]);  // same folded form as the forward (bitwise same mask)
            d[i] = y > 0.f ? v[u][i] : 0.f;
            acc0[i] += d[i];
            acc1[i] = fmaf(d[i], xh, acc1[i]);
          }
          store4(crow, kFast, nv, d);
        } else if constexpr (kEpi == EPI_STORE_ACT) {
          // Linear + bias + Swish (x * sigmoid(x), multimnist/model.py:379-381): the pre-activation is kept for the
          // backward, the activation is the next layer's operand
          float y[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[u][i] += bias[i];
            y[i] = v[u][i] * __fdividef(1.f, 1.f + __expf(-v[u][i]));
          }
          if (e.C != nullptr) store4(crow, kFast, nv, v[u]);
          store4(pptr + static_cast<long long>(u) * e.ldc, kFast, nv, y);
        } else if constexpr (kEpi == EPI_DGRAD_ACT) {
          // d pre = d act * swish'(pre), swish'(x) = s + x s (1 - s); its column sums are the producing Linear's bias gradient
          float d[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x = aux[u][i];
            const float sg = __fdividef(1.f, 1.f + __expf(-x));
            d[i] = v[u][i] * fmaf(x * sg, 1.f - sg, sg);
            acc0[i] += d[i];
          }
          store4(crow, kFast, nv, d);
        }
      }
      const int adv = min(kRB, seg - it);  // a segment may end inside a batch (statistics-group boundary)
      saddr += adv * ldst * 4;
      cptr += static_cast<long long>(adv) * e.ldc;
      if constexpr (kEpi == EPI_BCE) {
        trow += adv;
        if (trow >= e.target_rows) trow -= e.target_rows;
        if (pptr != nullptr) pptr += static_cast<long long>(adv) * e.ldc;
      }
      if constexpr (kEpi == EPI_STORE_ACT) pptr += static_cast<long long>(adv) * e.ldc;
      if constexpr (kEpi == EPI_DGRAD_BN || kEpi == EPI_DGRAD_ACT) hptr += static_cast<long long>(adv) * e.ldh;
    }
    if constexpr (kEpi == EPI_BCE) lsum *= g_scale;
    // ---- statistics-group boundary (or end of this warp's rows): publish the column partials
    // (s_red_col != nullptr: the whole tile belongs to one statistics group - partials go to shared memory
    //  and the CTA issues ONE global reduction per column afterwards instead of one per warp)
    if (e.stat0 != nullptr && s_red_col == nullptr) {
      const long long goff = static_cast<long long>(g) * p.stat_group_stride;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < nv) {
          atomicAdd(e.stat0 + goff + cn + i, acc0[i]);
          if (e.stat1 != nullptr) atomicAdd(e.stat1 + goff + cn + i, acc1[i]);
        }
        acc0[i] = 0.f;
        acc1[i] = 0.f;
      }
    }
    if (kEpi == EPI_BCE) {
      loss_acc[g & 3] += lsum;  // per-thread partial per term; reduced per warp / per CTA by the caller
      lsum = 0.f;
    }
    done += seg;
    if (aux_saddr != 0) aux_saddr += seg * aux_ld_bytes;
    ++g;
    seg_left = e.rows_per_group;
  }
  if (e.stat0 != nullptr && s_red_col != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s_red_col[i] = acc0[i];
      s_red_col[8 * p.block_n + i] = acc1[i];
    }
  }
}

// kGather (separate instantiations): 1 = one operand is an implicit im2col patch matrix; 2 = output-parity class of a
// transposed convolution (A gathered through a tap window, B = weights tap by tap via a rank-3 map, rows scattered)
template <int kKind, int kEpi, int kGather = 0>
__global__ void __launch_bounds__(kGemmThreads)
    gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const GemmKParams p) {
  using act_t = typename ActT<kKind>::type;
  constexpr int ESZ = (kKind == MVAE_F32) ? 4 : 2;
  constexpr int BK = 128 / ESZ;   // contraction elements per stage (one 128-B swizzle span)
  constexpr int UK = 32 / ESZ;    // contraction elements per tcgen05.mma
  constexpr int ATOM = 128 / ESZ; // MN elements per 128-B span (MN-major operands)
  constexpr int FMT = (kKind == MVAE_F32) ? 2 : 1;
  constexpr uint32_t MN_SBO = (kKind == MVAE_F32) ? 512 : 1024;
  constexpr uint32_t MN_LAYOUT = (kKind == MVAE_F32) ? 1 : 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t ready_bar[kMaxStages];  // gathered operand: the stage's gathered half is in place, MMA may read it
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_loss[4];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * p.block_n;
  const int m0 = blockIdx.y * kBlockM;
  const int kb0 = (kGather == 3 ? 0 : static_cast<int>(blockIdx.z)) * p.kb_per_split;  // kGather 3: blockIdx.z is the parity class
  const int nkb = min(p.kb_per_split, p.kb_total - kb0);
  const int stage_bytes = kAStageBytes + p.b_stage_bytes;
  const int S = p.stages;

  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(p.block_n)) tmem_cols <<= 1;

  long long* dbg = p.dbg ? p.dbg + 8ll * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
  auto stamp = [&](int slot) {
    if (dbg != nullptr) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[slot] = t;
    }
  };
  if (threadIdx.x == 0) stamp(0);

  if (threadIdx.x < 4) s_loss[threadIdx.x] = 0.f;
  if (p.red_off >= 0) {
    float* z = reinterpret_cast<float*>(smem + p.red_off);
    for (int i = threadIdx.x; i < 16 * p.block_n; i += kGemmThreads) z[i] = 0.f;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
      ptx::mbar_init(&ready_bar[s], 64);  // gather: one 64-thread group per ring slot
    }
    ptx::mbar_init(&accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&tmem_slot, tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) stamp(1);

  // Auxiliary epilogue operand (BCE: the target image tile; dgrad: the pre-BatchNorm activations): every thread
  // copies exactly the elements its own row pass will consume into shared memory with cp.async NOW, so the
  // L2 latency is hidden behind the main loop (the epilogue threads are otherwise idle until the accumulator is done).
  if constexpr (kEpi == EPI_BCE || kEpi == EPI_DGRAD_BN || kEpi == EPI_DGRAD_ACT) {
    if (p.aux_off >= 0) {
      const GemmEpilogue& ee = p.epi;
      const act_t* src = reinterpret_cast<const act_t*>(kEpi == EPI_BCE ? ee.target : ee.hpre);
      const long long ld = kEpi == EPI_BCE ? ee.ldt : ee.ldh;
      const int aux_ld = p.block_n * ESZ;
      const uint32_t aux_base = ptx::smem_u32(smem) + p.aux_off;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int col = ch * 128 + lane * 4;
        const int cn = n0 + col;
        if (col < p.block_n && cn + 3 < p.N) {
          int m = m0 + warp * 16;
          int srow = kEpi == EPI_BCE ? m % ee.target_rows : m;
          for (int rr = 0; rr < 16 && m < p.M; ++rr, ++m) {
            ptx::cp_async<4 * ESZ>(aux_base + (warp * 16 + rr) * aux_ld + col * ESZ, src + static_cast<long long>(srow) * ld + cn);
            if (kEpi == EPI_BCE) {
              if (++srow == ee.target_rows) srow = 0;
            } else {
              ++srow;
            }
          }
        }
      }
      ptx::cp_async_commit();
    }
  }

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int a_boxes = kBlockM / ATOM;
      const int b_boxes = (p.block_n + ATOM - 1) / ATOM;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        const bool tma_a = kGather == 0 || p.gather.mode != 1, tma_b = kGather == 0 || p.gather.mode != 2;  // the gathered operand comes from warps 4..7
        ptx::mbar_expect_tx(&full_bar[s], (tma_a ? kAStageBytes : 0) + (tma_b ? p.b_tx_bytes : 0));
        uint8_t* sa = smem + s * stage_bytes;
        uint8_t* sb = sa + kAStageBytes;
        const int kc = (kb0 + i) * BK;
        if (!tma_a) {
        } else if (!p.a_mn) {
          ptx::tma_load_2d(sa, &tmA, &full_bar[s], kc, m0);
        } else {
          for (int j = 0; j < a_boxes; ++j) ptx::tma_load_2d(sa + j * (BK * 128), &tmA, &full_bar[s], m0 + j * ATOM, kc);
        }
        if (!tma_b) {
        } else if constexpr (kGather == 3) {
          // as kGather 2, with the tap tables of this CTA's parity class (a, b) = (z / stride, z % stride)
          const ConvGather& cg = p.gather;
          const int za = static_cast<int>(blockIdx.z) / cg.sc_stride, zb = static_cast<int>(blockIdx.z) - za * cg.sc_stride;
          const int w_pos = kc / cg.C, ci0 = kc - w_pos * cg.C;
          const int th = w_pos / cg.ksize_w, tw = w_pos - th * cg.ksize_w;
          const int tap = cg.ax_k[za][th] * cg.kk + cg.ax_k[zb][tw];
          for (int j = 0; j < b_boxes; ++j) ptx::tma_load_3d(sb + j * (BK * 128), &tmB, &full_bar[s], n0 + j * ATOM, tap, ci0);
        } else if constexpr (kGather == 2) {
          // weights of the tap this K block belongs to: window position (th, tw) -> tap kh_tab[th]*kk + kw_tab[tw]
          const ConvGather& cg = p.gather;
          const int w_pos = kc / cg.C, ci0 = kc - w_pos * cg.C;
          const int th = w_pos / cg.ksize_w, tw = w_pos - th * cg.ksize_w;
          const int tap = cg.kh_tab[th] * cg.kk + cg.kw_tab[tw];
          for (int j = 0; j < b_boxes; ++j) ptx::tma_load_3d(sb + j * (BK * 128), &tmB, &full_bar[s], n0 + j * ATOM, tap, ci0);
        } else if (!p.b_mn) {
          ptx::tma_load_2d(sb, &tmB, &full_bar[s], kc, n0);
        } else {
          for (int j = 0; j < b_boxes; ++j) ptx::tma_load_2d(sb + j * (BK * 128), &tmB, &full_bar[s], n0 + j * ATOM, kc);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc(FMT, p.a_mn, p.b_mn, kBlockM, p.block_n);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        if (kGather != 0) {  // TMA half landed AND the gathered half is in place
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::mbar_wait(&ready_bar[s], ph);
        } else {
          ptx::mbar_wait(&full_bar[s], ph);
        }
        ptx::tc_fence_after();
        if (i == 0) stamp(2);
        const uint32_t a_base = ptx::smem_u32(smem + s * stage_bytes);
        const uint32_t b_base = a_base + kAStageBytes;
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {
          // MN-major: 128-B column blocks BK*128 B apart (LBO); k-row groups of 8 (16-bit, SW128) or
          // 4 (tf32, SW128 with 32-B atoms) rows, dense -> SBO 1024 / 512.  K-major: +32 B per k-step
          // inside the 128-B swizzle span, 8-row groups 1024 B apart.
          const uint64_t adesc = p.a_mn ? ptx::make_smem_desc(a_base + k * (UK * 128), BK * 128, MN_SBO, MN_LAYOUT)
                                        : ptx::make_smem_desc(a_base + k * 32, 16, 1024);
          const uint64_t bdesc = p.b_mn ? ptx::make_smem_desc(b_base + k * (UK * 128), BK * 128, MN_SBO, MN_LAYOUT)
                                        : ptx::make_smem_desc(b_base + k * 32, 16, 1024);
          ptx::umma<kKind>(tmem_base, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      ptx::umma_commit(&accum_bar);  // accumulator complete
      stamp(3);
    }
  } else if (kGather == 3 && warp >= 4) {
    // ------------------------------------------------------------ all parity classes in one launch (warps 4..7)
    // Same producer as the kGather 1/2 patch-matrix-as-A path below (two 64-thread groups own the even / odd ring slots),
    // with the tap window's pads taken from the tables of this CTA's class.
    if constexpr (kGather == 3) {
      const ConvGather& cg = p.gather;
      const int za = static_cast<int>(blockIdx.z) / cg.sc_stride, zb = static_cast<int>(blockIdx.z) - za * cg.sc_stride;
      const int pad_h = cg.ax_pad[za], pad_w = cg.ax_pad[zb];
      const int grp = (warp - 4) >> 1;
      const int t64 = threadIdx.x & 63;
      const int c = t64 & 7, rbase = t64 >> 3;
      const uint32_t chunk_off = static_cast<uint32_t>((c ^ rbase) << 4);
      const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(cg.X);
      const uint32_t hw = static_cast<uint32_t>(cg.Ho * cg.Wo);
      int off[16], hw0[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int m = m0 + rbase + 8 * j;
        const bool valid = m < p.M;
        const uint32_t n = static_cast<uint32_t>((static_cast<unsigned long long>(m) * cg.magic_hw) >> 40);
        const uint32_t rem = static_cast<uint32_t>(m) - n * hw;
        const uint32_t u = static_cast<uint32_t>((static_cast<unsigned long long>(rem) * cg.magic_w) >> 40);
        const uint32_t v = rem - u * static_cast<uint32_t>(cg.Wo);
        const int hi0 = valid ? static_cast<int>(u) - pad_h : -20000;
        const int wi0 = static_cast<int>(v) - pad_w;
        off[j] = valid ? static_cast<int>(n) * static_cast<int>(cg.sn) + hi0 * static_cast<int>(cg.sh) + wi0 * static_cast<int>(cg.sw) : 0;
        hw0[j] = (hi0 << 16) | (wi0 & 0xffff);
      }
      int k0 = c * 8;
      int tap = k0 / cg.C, ch = k0 - tap * cg.C;
      int kh = tap / cg.ksize_w, kw = tap - kh * cg.ksize_w;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        if ((s & 1) == grp) {
          const uint32_t ph = (i / S) & 1;
          const int koff = kh * static_cast<int>(cg.sh) + kw * static_cast<int>(cg.sw) + ch;
          const bool kvalid = k0 < p.K;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t dst = ptx::smem_u32(smem + s * stage_bytes) + rbase * 128 + chunk_off;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int hi = (hw0[j] >> 16) + kh, wi = ((hw0[j] << 16) >> 16) + kw;
            const bool ok = kvalid && static_cast<unsigned>(hi) < static_cast<unsigned>(cg.H) &&
                            static_cast<unsigned>(wi) < static_cast<unsigned>(cg.W);
            ptx::cp_async16_zfill(dst + j * (8 * 128), ok ? X + (off[j] + koff) : X, ok ? 16u : 0u);
          }
          ptx::cp_async_commit();
          ptx::cp_async_wait_pending(0);
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(&ready_bar[s]);
        }
        k0 += BK;
        ch += BK;
        while (ch >= cg.C) {
          ch -= cg.C;
          if (++kw == cg.ksize_w) {
            kw = 0;
            ++kh;
          }
        }
      }
    }
  } else if ((kGather == 1 || kGather == 2) && warp >= 4) {
    // ------------------------------------------------------------ implicit patch-matrix operand (warps 4..7)
    // The convolution's im2col matrix is never materialised: these threads copy 16-byte chunks (8 bf16 channels of one
    // filter tap) from the NHWC activation into the SWIZZLE_128B stage the tensor core reads (logical chunk c of row r
    // lives at physical chunk c ^ (r & 7)); padding and out-of-range rows are zero-filled by cp.async.  Two groups of
    // 64 threads own the even / odd ring SLOTS (warps 4,5 / 6,7), so two stages are in flight and each is published
    // through ready_bar (proxy fence first) the moment it lands - the same decoupling TMA gives the other operand.
    // (Ownership is by slot, not by stage index: a group then sees every phase of its slots' barriers in order.)
    if constexpr (kGather == 1 || kGather == 2) {
      const ConvGather& cg = p.gather;
      const int grp = (warp - 4) >> 1;
      const int t64 = threadIdx.x & 63;
      const int c = t64 & 7, rbase = t64 >> 3;   // chunk column; rows rbase + 8 j  ((row & 7) == rbase)
      const uint32_t chunk_off = static_cast<uint32_t>((c ^ rbase) << 4);
      const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(cg.X);
      const uint32_t hw = static_cast<uint32_t>(cg.Ho * cg.Wo);
      const int win_w = kGather == 2 ? cg.ksize_w : cg.ksize;   // taps per window row
      const int pad_w = kGather == 2 ? cg.pad_w : cg.pad;
      // pixel index -> element offset of the patch origin and (hi0 << 16 | wi0 & 0xffff); exact magic-number division
      auto locate = [&](uint32_t m, bool valid, int& off, int& hw0) {
        const uint32_t n = static_cast<uint32_t>((static_cast<unsigned long long>(m) * cg.magic_hw) >> 40);
        const uint32_t rem = m - n * hw;
        const uint32_t ho = static_cast<uint32_t>((static_cast<unsigned long long>(rem) * cg.magic_w) >> 40);
        const uint32_t wo = rem - ho * static_cast<uint32_t>(cg.Wo);
        const int hi0 = valid ? static_cast<int>(ho) * cg.stride - cg.pad : -20000;
        const int wi0 = static_cast<int>(wo) * cg.stride - pad_w;
        off = valid ? static_cast<int>(n) * static_cast<int>(cg.sn) + hi0 * static_cast<int>(cg.sh) + wi0 * static_cast<int>(cg.sw) : 0;
        hw0 = (hi0 << 16) | (wi0 & 0xffff);
      };
      if (cg.mode == 1) {
        // A[m, k]: fixed pixels (16 rows per thread), k advances with the stage
        int off[16], hw0[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = m0 + rbase + 8 * j;
          locate(static_cast<uint32_t>(m), m < p.M, off[j], hw0[j]);
        }
        int k0 = kb0 * BK + c * 8;
        int tap = k0 / cg.C, ch = k0 - tap * cg.C;
        int kh = tap / win_w, kw = tap - kh * win_w;
        for (int i = 0; i < nkb; ++i) {
          const int s = i % S;
          if ((s & 1) == grp) {
            const uint32_t ph = (i / S) & 1;
            const int koff = kh * static_cast<int>(cg.sh) + kw * static_cast<int>(cg.sw) + ch;
            const bool kvalid = k0 < p.K;
            ptx::mbar_wait(&empty_bar[s], ph ^ 1);
            const uint32_t dst = ptx::smem_u32(smem + s * stage_bytes) + rbase * 128 + chunk_off;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int hi = (hw0[j] >> 16) + kh, wi = ((hw0[j] << 16) >> 16) + kw;
              const bool ok = kvalid && static_cast<unsigned>(hi) < static_cast<unsigned>(cg.H) &&
                              static_cast<unsigned>(wi) < static_cast<unsigned>(cg.W);
              ptx::cp_async16_zfill(dst + j * (8 * 128), ok ? X + (off[j] + koff) : X, ok ? 16u : 0u);
            }
            ptx::cp_async_commit();
            ptx::cp_async_wait_pending(0);
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(&ready_bar[s]);
          }
          // the next stage: BK further along (kh, kw, c)
          k0 += BK;
          ch += BK;
          while (ch >= cg.C) {
            ch -= cg.C;
            if (++kw == win_w) {
              kw = 0;
              ++kh;
            }
          }
        }
      } else {
        // B, MN-major: stage row kr = reduction index (pixel (kb0+i)*BK + kr), 128-byte column boxes of 64 patch
        // entries; this thread owns chunk c of every box (fixed taps / channels) for rows rbase + 8 jr
        const uint32_t b_boxes = static_cast<uint32_t>((p.block_n + ATOM - 1) / ATOM);
        int koff[4], khw[4];
        bool kvalid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k0 = n0 + j * ATOM + c * 8;
          const int tap = k0 / cg.C, ch = k0 - tap * cg.C;
          const int kh = tap / cg.ksize, kw = tap - kh * cg.ksize;
          khw[j] = (kh << 16) | kw;
          koff[j] = kh * static_cast<int>(cg.sh) + kw * static_cast<int>(cg.sw) + ch;
          kvalid[j] = static_cast<uint32_t>(j) < b_boxes && k0 < p.N;
        }
        for (int i = 0; i < nkb; ++i) {
          const int s = i % S;
          if ((s & 1) != grp) continue;
          const uint32_t ph = (i / S) & 1;
          int off[8], hw0[8];
#pragma unroll
          for (int jr = 0; jr < 8; ++jr) {
            const int m = (kb0 + i) * BK + rbase + 8 * jr;
            locate(static_cast<uint32_t>(m), m < p.K, off[jr], hw0[jr]);
          }
          ptx::mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t dst = ptx::smem_u32(smem + s * stage_bytes) + kAStageBytes + rbase * 128 + chunk_off;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (static_cast<uint32_t>(j) < b_boxes) {
#pragma unroll
              for (int jr = 0; jr < 8; ++jr) {
                const int hi = (hw0[jr] >> 16) + (khw[j] >> 16), wi = ((hw0[jr] << 16) >> 16) + (khw[j] & 0xffff);
                const bool ok = kvalid[j] && static_cast<unsigned>(hi) < static_cast<unsigned>(cg.H) &&
                                static_cast<unsigned>(wi) < static_cast<unsigned>(cg.W);
                ptx::cp_async16_zfill(dst + j * (BK * 128) + jr * (8 * 128), ok ? X + (off[jr] + koff[j]) : X, ok ? 16u : 0u);
              }
            }
          }
          ptx::cp_async_commit();
          ptx::cp_async_wait_pending(0);
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(&ready_bar[s]);
        }
      }
    }
  }
  __syncwarp();  // producer / MMA warps reconverge before joining the epilogue

  // -------------------------------------------------------------- epilogue (all 8 warps)
  // (1) TMEM -> registers -> padded fp32 smem tile.  Warps w and w+4 share TMEM lane quarter w&3 and take
  //     alternate 16-column chunks.  The tile aliases the operand ring, which is fully drained by now.
  const GemmEpilogue& e = p.epi;
  const uint32_t stage_addr = ptx::smem_u32(smem);
  const int ldst = p.block_n + kStagePad;  // words
  ptx::mbar_wait(&accum_bar, 0);
  ptx::tc_fence_after();
  if (threadIdx.x == 64) stamp(4);
  bool stored_directly = false;
  if constexpr (kEpi == EPI_STORE) {
    if (p.direct_store) {
      // Plain store (no column statistics): every thread owns one accumulator row (TMEM lane) and writes 16 consecutive
      // columns per tcgen05.ld - 32-byte (bf16) / 64-byte (fp32) sector-aligned runs, no shared-memory staging pass.
      stored_directly = true;
      const int q = warp & 3;
      const int row = m0 + q * 32 + lane;
      long long row_off = static_cast<long long>(row) * e.ldc;
      if constexpr (kGather == 3) {
        const ConvGather& cg = p.gather;
        const int za = static_cast<int>(blockIdx.z) / cg.sc_stride, zb = static_cast<int>(blockIdx.z) - za * cg.sc_stride;
        const int hw = cg.Ho * cg.Wo;
        const int n = row / hw, rem = row - n * hw;
        const int u = rem / cg.Wo, v2 = rem - u * cg.Wo;
        row_off = ((static_cast<long long>(n) * cg.sc_hout + cg.sc_stride * u + za) * cg.sc_wout + cg.sc_stride * v2 + zb) * e.ldc;
      }
      if constexpr (kGather == 2) {
        // row (n, u, v) of this parity class -> pixel (n, stride*u + a, stride*v + b) of the interleaved output image
        const ConvGather& cg = p.gather;
        const int hw = cg.Ho * cg.Wo;
        const int n = row / hw, rem = row - n * hw;
        const int u = rem / cg.Wo, v2 = rem - u * cg.Wo;
        row_off = ((static_cast<long long>(n) * cg.sc_hout + cg.sc_stride * u + cg.sc_a) * cg.sc_wout + cg.sc_stride * v2 + cg.sc_b) * e.ldc;
      }
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int c = (warp >> 2) * 16; c < p.block_n; c += 32) {
        uint32_t v[16];
        ptx::tmem_ld16(t_base + c, v);
        ptx::tmem_ld_wait();
        const int cn = n0 + c;
        if (row < p.M && cn < p.N) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (e.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (cn + j < p.N) f[j] += __ldg(e.bias + cn + j);
          }
          if (e.c_dtype == MVAE_F32) {
            float* dst = reinterpret_cast<float*>(e.C) + row_off + cn;
            if (cn + 16 <= p.N) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cn + j < p.N) dst[j] = f[j];
            }
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.C) + row_off + cn;
            if (cn + 16 <= p.N) {
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                w[j] = *reinterpret_cast<const uint32_t*>(&h);
              }
              *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(dst + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cn + j < p.N) dst[j] = __float2bfloat16_rn(f[j]);
            }
          }
        }
      }
    }
  }
  if (!stored_directly) {
  {
    const int q = warp & 3;
    const uint32_t row_addr = stage_addr + static_cast<uint32_t>((q * 32 + lane) * ldst) * 4u;
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int c = (warp >> 2) * 16; c < p.block_n; c += 32) {
      uint32_t v[16];
      ptx::tmem_ld16(t_base + c, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) ptx::sts128(row_addr + (c + 4 * j) * 4, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 64) stamp(5);

  // (2) Row pass: warp w owns rows [16w, 16w+16); lane owns 4 consecutive columns per 128-column chunk, so
  //     every global access is a full 512-B (fp32) / 256-B (bf16) line per warp and the per-column
  //     statistics accumulate in registers.  A flush (atomics) happens at each statistics-group boundary.
  {
    const int r_begin = warp * 16;
    const int rows_here = min(16, p.M - m0 - r_begin);
    float loss_acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.aux_off >= 0) ptx::cp_async_wait_all();  // this thread's own prefetched auxiliary elements
    // one statistics group for the whole tile (always true for the un-grouped BCE bias gradient)?
    const int g_first = m0 / e.rows_per_group;
    const int g_last = (min(m0 + kBlockM, p.M) - 1) / e.rows_per_group;
    const bool cta_reduce = p.red_off >= 0 && e.stat0 != nullptr && (g_first == g_last || p.stat_group_stride == 0);
    float* s_red = reinterpret_cast<float*>(smem + (p.red_off >= 0 ? p.red_off : 0));
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int col = ch * 128 + lane * 4;
      const int cn = n0 + col;
      int nv = 0;
      if (col < p.block_n) nv = min(4, p.N - cn);
      if (nv <= 0 || rows_here <= 0) continue;
      const uint32_t saddr = stage_addr + static_cast<uint32_t>(r_begin * ldst + col) * 4u;
      const bool fast = (nv == 4) && (p.vec_ok != 0);
      const int aux_ld = p.block_n * ESZ;
      const uint32_t aux_sa =
          (p.aux_off >= 0) ? stage_addr + static_cast<uint32_t>(p.aux_off + r_begin * aux_ld + col * ESZ) : 0u;
      float* red_col = cta_reduce ? s_red + warp * p.block_n + col : nullptr;
      if (e.c_dtype == MVAE_F32) {
        if (fast)
          epilogue_rows<kKind, kEpi, float, true>(p, saddr, ldst, m0 + r_begin, rows_here, cn, 4, lane, loss_acc, aux_sa, aux_ld, red_col);
        else
          epilogue_rows<kKind, kEpi, float, false>(p, saddr, ldst, m0 + r_begin, rows_here, cn, nv, lane, loss_acc, 0u, 0, red_col);
      } else {
        if (fast)
          epilogue_rows<kKind, kEpi, __nv_bfloat16, true>(p, saddr, ldst, m0 + r_begin, rows_here, cn, 4, lane, loss_acc, aux_sa, aux_ld, red_col);
        else
          epilogue_rows<kKind, kEpi, __nv_bfloat16, false>(p, saddr, ldst, m0 + r_begin, rows_here, cn, nv, lane, loss_acc, 0u, 0, red_col);
      }
    }
    if (kEpi == EPI_BCE) {
      // all lanes are converged here: warp-reduce the per-term loss partials, one shared atomic per warp per term
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float v = loss_acc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v != 0.f) atomicAdd(&s_loss[t], v);
      }
    }
    if (cta_reduce) {
      __syncthreads();
      const long long goff = static_cast<long long>(g_first) * p.stat_group_stride;
      for (int c = threadIdx.x; c < p.block_n; c += kGemmThreads) {
        const int n = n0 + c;
        if (n < p.N) {
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            a0 += s_red[w * p.block_n + c];
            a1 += s_red[(8 + w) * p.block_n + c];
          }
          atomicAdd(e.stat0 + goff + n, a0);
          if (e.stat1 != nullptr) atomicAdd(e.stat1 + goff + n, a1);
        }
      }
    }
    if (threadIdx.x == 64) stamp(6);
  }
  }  // !stored_directly

  ptx::tc_fence_before();
  __syncthreads();
  if (kEpi == EPI_BCE && threadIdx.x < 4 && p.epi.loss != nullptr && s_loss[threadIdx.x] != 0.f)
    atomicAdd(p.epi.loss + threadIdx.x, s_loss[threadIdx.x]);
  if (warp == 2) ptx::tmem_dealloc(tmem_base, tmem_cols);
  if (threadIdx.x == 64) stamp(7);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// rows x cols (cols contiguous) 2-D tensor, box = box_cols x box_rows, SWIZZLE_128B (box_cols * esz == 128).
int make_tmap(CUtensorMap* out, int kind, const void* base, long long rows, long long cols, long long ld,
              int box_cols, int box_rows, bool mn_major) {
  EncodeTiledFn enc = get_encode_fn();
  MVAE_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  const int esz = kind == MVAE_F32 ? 4 : 2;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  // TFLOAT32 maps make TMA round fp32 -> tf32 to nearest on the way into smem (measured: GEMM error
  // 2.9e-4 of max|C| vs 8.2e-4 with plain FLOAT32, where the tensor core truncates the mantissa).
  static const int tf32_map = env_int("MVAE_TMA_TF32", 1);
  CUtensorMapDataType dt = kind == MVAE_F32 ? (tf32_map ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32)
                                            : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  // MN-major 32-bit operands must use the 32-B-atom flavour of the 128-B swizzle (see ptx::make_smem_desc).
  const CUtensorMapSwizzle sw =
      (mn_major && kind == MVAE_F32) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVAE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r,
               rows, cols, ld, box_cols, box_rows);
  return 0;
}

// weights [c rows, taps, n] (n contiguous, ld_tap elements between taps), bf16, box = box_n x 1 tap x box_c rows, SWIZZLE_128B
int make_tmap_w3(CUtensorMap* out, const void* base, long long n, long long taps, long long c, long long ld_tap, int box_n,
                 int box_c) {
  EncodeTiledFn enc = get_encode_fn();
  MVAE_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(n), static_cast<cuuint64_t>(taps), static_cast<cuuint64_t>(c)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld_tap) * 2, static_cast<cuuint64_t>(taps) * static_cast<cuuint64_t>(ld_tap) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_n), 1u, static_cast<cuuint32_t>(box_c)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVAE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (rank 3) failed (%d): n=%lld taps=%lld c=%lld ld=%lld", (int)r, n, taps, c,
               ld_tap);
  return 0;
}

template <int kKind, int kEpi, int kGather = 0>
int ensure_smem(int dyn_smem) {
  static int smem_set = 0;  // per instantiation; monotone, benign race
  if (dyn_smem > smem_set) {
    if (smem_set == 0)  // always the largest shared-memory carve-out: several CTAs per SM is the operating point
      MVAE_CUDA(cudaFuncSetAttribute(gemm_kernel<kKind, kEpi, kGather>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
    MVAE_CUDA(cudaFuncSetAttribute(gemm_kernel<kKind, kEpi, kGather>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem));
    smem_set = dyn_smem;
  }
  return 0;
}
template <int kKind, int kEpi, int kGather = 0>
int launch_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKParams& kp, dim3 grid, int dyn_smem,
                cudaStream_t stream) {
  if (int rc = ensure_smem<kKind, kEpi, kGather>(dyn_smem)) return rc;
  return launch_kernel(gemm_kernel<kKind, kEpi, kGather>, grid, dim3(kGemmThreads), static_cast<size_t>(dyn_smem), stream, ta, tb, kp);
}

int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace

// Bring-up aid: timestamps of the next launches with epilogue kind `g_dbg_epi` go to `g_dbg_times`.
static long long* g_dbg_times = nullptr;
static int g_dbg_epi = -1;
void set_gemm_debug_times(void* ptr, int epi_kind) {
  g_dbg_times = static_cast<long long*>(ptr);
  g_dbg_epi = epi_kind;
}

static int launch_gemm_impl(const GemmDesc& g, cudaStream_t stream);

// ---- 3xTF32: operand split.  dst holds three copies of the [R, Cc] source - parts (p0, p1, p2), each the tf32-rounded
// value (hi, round to nearest even on the 10-bit mantissa) or the exact remainder x - hi (lo) - side by side along the
// contraction axis: along the columns for a K-major operand (part p at column offset p*cpad) or along the rows for an
// MN-major one (part p at row offset p*R).  Padding columns [Cc, cpad) are zero-filled.
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ src, long long ld_src, long long R, int Cc, int cpad,
                                                     float* __restrict__ dst, long long ld_dst, long long part_stride, int lo_mask) {
  const long long total = R * cpad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cpad;
    const int c = static_cast<int>(i - r * cpad);
    float hi = 0.f, lo = 0.f;
    if (c < Cc) {
      const float x = src[r * ld_src + c];
      const uint32_t b = __float_as_uint(x);
      hi = __uint_as_float((b + 0x0FFFu + ((b >> 13) & 1u)) & 0xFFFFE000u);
      lo = x - hi;
    }
    float* d = dst + r * ld_dst + c;
#pragma unroll
    for (int p = 0; p < 3; ++p) d[p * part_stride] = ((lo_mask >> p) & 1) ? lo : hi;
  }
}

static int launch_split3(const void* src, long long ld, int major, long long rows, int K, void* dst, int lo_mask, long long* ld_out,
                         cudaStream_t st) {
  // major 0: src is [rows, K] -> dst [rows, 3*Kp]; major 1: src is [K, rows] -> dst [3*K, rows_p]
  const int Kp = (K + 3) / 4 * 4;
  const long long rows_p = (rows + 3) / 4 * 4;
  long long R, ld_dst, part;
  int Cc, cpad;
  if (major == 0) { R = rows; Cc = K; cpad = Kp; ld_dst = 3ll * Kp; part = Kp; }
  else { R = K; Cc = static_cast<int>(rows); cpad = static_cast<int>(rows_p); ld_dst = rows_p; part = static_cast<long long>(K) * rows_p; }
  *ld_out = ld_dst;
  const long long total = R * cpad;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  split3_kernel<<<blocks, 256, 0, st>>>(static_cast<const float*>(src), ld, R, Cc, cpad, static_cast<float*>(dst), ld_dst, part, lo_mask);
  MVAE_CUDA(cudaGetLastError());
  note_launch(1);
  return 0;
}

int launch_gemm(const GemmDesc& g, cudaStream_t stream) {
  if (!g.x3) return launch_gemm_impl(g, stream);
  MVAE_REQUIRE(g.kind == MVAE_F32 && g.gather.mode == 0, "gemm: 3xTF32 needs fp32 storage and a plain (non-gather) GEMM");
  MVAE_REQUIRE(g.x3_a != nullptr && g.x3_b != nullptr, "gemm: 3xTF32 needs the two split-operand scratch buffers");
  GemmDesc h = g;
  h.x3 = 0;
  long long lda3 = 0, ldb3 = 0;
  if (launch_split3(g.A, g.lda, g.a_mn, g.M, g.K, g.x3_a, /*lo parts*/ 0b010, &lda3, stream)) return 1;   // [hi | lo | hi]
  if (launch_split3(g.B, g.ldb, g.b_mn, g.N, g.K, g.x3_b, /*lo parts*/ 0b100, &ldb3, stream)) return 1;   // [hi | hi | lo]
  h.A = g.x3_a; h.lda = lda3;
  h.B = g.x3_b; h.ldb = ldb3;
  // contraction length: K-major parts are padded to Kp columns, MN-major parts are exactly K rows; mixed majors must agree
  const int Kp = (g.K + 3) / 4 * 4;
  MVAE_REQUIRE((g.a_mn && g.b_mn) || (!g.a_mn && !g.b_mn) || Kp == g.K, "gemm: 3xTF32 with mixed operand majors needs K %% 4 == 0");
  h.K = 3 * ((g.a_mn && g.b_mn) ? g.K : Kp);
  // The tensor core adds each instruction's products into the fp32 TMEM accumulator with truncation; over a long
  // contraction that bias (~2^-24 per instruction, all of one sign) is amplified by the cancellation in weight-gradient
  // sums (zero-mean BatchNorm gradients x positive-mean inputs): measured 2.8e-3 on dW of the first Linear at B = 4096.
  // Split-K parts of <= 384 contraction elements are accumulated on chip, the parts are combined by round-to-nearest fp32
  // reductions in L2.
  if (h.epi.kind == EPI_ATOMIC && h.split_k <= 0) {
    static const int part = env_int("MVAE_X3_SPLIT_ELEMS", 384);
    h.split_k = std::max(1, (h.K + part - 1) / part);
  }
  return launch_gemm_impl(h, stream);
}

static int launch_gemm_impl(const GemmDesc& g, cudaStream_t stream) {
  const int esz = g.kind == MVAE_F32 ? 4 : 2;
  const int BK = 128 / esz;
  MVAE_REQUIRE(g.kind == MVAE_F32 || g.kind == MVAE_BF16, "gemm: bad kind %d", g.kind);
  MVAE_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem %dx%dx%d", g.M, g.N, g.K);
  MVAE_REQUIRE((g.gather.mode == 1 || g.gather.mode >= 3 || (g.lda * esz) % 16 == 0) && (g.gather.mode == 2 || (g.ldb * esz) % 16 == 0),
               "gemm: lda/ldb (%lld,%lld) must be 16-byte multiples", g.lda, g.ldb);
  MVAE_REQUIRE((g.gather.mode == 1 || g.gather.mode >= 3 || (reinterpret_cast<uintptr_t>(g.A) & 15) == 0) &&
                   (g.gather.mode == 2 || (reinterpret_cast<uintptr_t>(g.B) & 15) == 0),
               "gemm: A/B must be 16-byte aligned");
  const GemmEpilogue& e = g.epi;
  const ConvGather& cg = g.gather;
  if (cg.mode != 0) {
    MVAE_REQUIRE(cg.mode >= 1 && cg.mode <= 4, "gemm: bad gather mode %d", cg.mode);
    MVAE_REQUIRE(g.kind == MVAE_BF16, "gemm: the implicit patch-matrix operand is implemented for bf16 storage only");
    MVAE_REQUIRE(cg.X != nullptr && (reinterpret_cast<uintptr_t>(cg.X) & 15) == 0, "gemm: gather source must be 16-byte aligned");
    MVAE_REQUIRE(cg.C > 0 && cg.C % 8 == 0 && cg.sn % 8 == 0 && cg.sh % 8 == 0 && cg.sw % 8 == 0,
                 "gemm: gather needs channels and strides that are multiples of 8 (16-byte chunks never straddle a tap)");
    MVAE_REQUIRE(cg.ksize > 0 && cg.stride > 0 && cg.Ho > 0 && cg.Wo > 0, "gemm: bad gather geometry");
    MVAE_REQUIRE(e.kind == EPI_STORE || e.kind == EPI_ATOMIC, "gemm: gather supports the store / accumulate epilogues");
    MVAE_REQUIRE(static_cast<long long>(cg.Ho) * cg.Wo < 65536 && cg.H < 16384 && cg.W < 16384 && cg.pad < 4096,
                 "gemm: gather geometry too large for the packed coordinates");
    MVAE_REQUIRE(cg.extent < (1ll << 31) && (cg.mode == 2 ? g.K : g.M) < (1 << 24), "gemm: gather source too large for 32-bit offsets");
    if (cg.mode == 1) MVAE_REQUIRE(!g.a_mn && g.K == cg.ksize * cg.ksize * cg.C, "gemm: gather A needs K = k*k*C, K-major");
    if (cg.mode == 2) MVAE_REQUIRE(g.b_mn && g.N == cg.ksize * cg.ksize * cg.C, "gemm: gather B needs N = k*k*C, MN-major");
    if (cg.mode == 3 || cg.mode == 4) {
      MVAE_REQUIRE(!g.a_mn && g.b_mn && e.kind == EPI_STORE, "gemm: transposed-conv class needs K-major A, MN-major weights, store epilogue");
      MVAE_REQUIRE(cg.ksize_w > 0 && cg.ksize <= 8 && cg.ksize_w <= 8 && g.K == cg.ksize * cg.ksize_w * cg.C,
                   "gemm: transposed-conv class needs K = taps_h*taps_w*C with at most 8 taps per axis");
      MVAE_REQUIRE(cg.C % BK == 0, "gemm: transposed-conv class needs C %% %d == 0 (a K block never straddles a tap)", BK);
      MVAE_REQUIRE(cg.kk > 0 && cg.sc_stride > 0 && cg.sc_a < cg.sc_stride && cg.sc_b < cg.sc_stride &&
                       cg.sc_stride * (cg.Ho - 1) + cg.sc_a < cg.sc_hout && cg.sc_stride * (cg.Wo - 1) + cg.sc_b < cg.sc_wout,
                   "gemm: transposed-conv class does not fit the output image");
      if (cg.mode == 4)  // every class (a, b) < stride must fit: the last one reaches furthest
        MVAE_REQUIRE(cg.sc_stride <= 4 && cg.sc_stride * cg.Ho <= cg.sc_hout && cg.sc_stride * cg.Wo <= cg.sc_wout,
                     "gemm: merged transposed-conv classes need stride <= 4 and equal class grids inside the output image");
    }
  }
  MVAE_REQUIRE(e.C != nullptr || (e.kind == EPI_STORE_ACT && e.probs != nullptr), "gemm: null output");
  if (e.kind == EPI_STORE_ACT) MVAE_REQUIRE(e.probs != nullptr, "gemm: the activation epilogue needs its output");
  if (e.kind == EPI_DGRAD_ACT) MVAE_REQUIRE(e.hpre != nullptr, "gemm: the activation-backward epilogue needs the pre-activations");
  MVAE_REQUIRE(e.kind != EPI_ATOMIC || e.c_dtype == MVAE_F32, "gemm: atomic epilogue needs fp32 output");

  const int tiles_m = ceil_div(g.M, kBlockM);
  const int kb_total = ceil_div(g.K, BK);

  // ---- tile width / split-K / pipeline depth from a small cost model fitted to measured sweeps on B200
  // (tools/gemm_sweep.py): these problems are bound by L2->SM operand streaming (~77 GB/s per SM, ~10 TB/s
  // chip-wide), not by the tensor pipe, and the epilogue of one CTA only overlaps with the main loop of
  // ANOTHER CTA on the same SM - so shallow rings (2 stages) with 2-3 co-resident CTAs beat deep rings.
  const bool atomic = e.kind == EPI_ATOMIC;
  const int sms = 148;
  static const int use_aux = env_int("MVAE_GEMM_AUX", 1);
  auto al0 = [](const void* p, int a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
  const bool dgrad_aux = e.kind == EPI_DGRAD_BN || e.kind == EPI_DGRAD_ACT;
  bool aux_ok = use_aux != 0 && (e.kind == EPI_BCE || dgrad_aux) && (e.ldc % 4 == 0);
  if (e.kind == EPI_BCE) aux_ok = aux_ok && (e.ldt % 4 == 0) && al0(e.target, 4 * esz);
  if (dgrad_aux) aux_ok = aux_ok && (e.ldh % 4 == 0) && al0(e.hpre, 4 * esz);
  auto aux_bytes = [&](int bn) -> int { return aux_ok ? kBlockM * bn * esz : 0; };
  static const int use_red = env_int("MVAE_GEMM_CTA_REDUCE", 1);
  const bool red_ok = use_red != 0 && e.stat0 != nullptr && e.kind != EPI_ATOMIC;
  auto red_bytes = [&](int bn) -> int { return red_ok ? 16 * bn * 4 : 0; };
  auto plan_for = [&](int bn, int& split_o, int& stages_o, int& dyn_o, int& bstage_o, int& btx_o) -> double {
    const int tn = ceil_div(g.N, bn);
    const long long tiles = static_cast<long long>(tiles_m) * tn;
    int split = g.split_k;
    if (!atomic) split = 1;
    if (split <= 0) split = ceil_div(env_int("MVAE_GEMM_SPLIT_TARGET", 148), tiles);
    if (split > kb_total) split = kb_total;
    if (split < 1) split = 1;
    const int kbps = ceil_div(kb_total, split);
    split = ceil_div(kb_total, kbps);
    const long long ctas = tiles * split;
    const int b_boxes = ceil_div(bn, BK);
    const int b_tx = g.b_mn ? b_boxes * BK * 128 : bn * 128;
    const int b_stage = (b_tx + 1023) / 1024 * 1024;
    const int stage_bytes = kAStageBytes + b_stage;
    const int staging = kBlockM * (bn + kStagePad) * 4;
    int stages = g.stages > 0 ? g.stages
                 : g.gather.mode != 0 ? env_int("MVAE_GATHER_STAGES", 2)   // measured: 2 > 3 > 4 > 6 (co-residency beats depth)
                                      : env_int("MVAE_GEMM_STAGES", ctas > sms ? 2 : 4);
    if (stages > kbps) stages = kbps;
    if (stages > kMaxStages) stages = kMaxStages;
    const int max_dyn = 227 * 1024 - 2048;
    while (stages > 1 && stages * stage_bytes + 1024 > max_dyn) --stages;
    int dyn = stages * stage_bytes;
    if (dyn < staging) dyn = staging;
    dyn = (dyn + 15) / 16 * 16 + aux_bytes(bn) + red_bytes(bn);
    dyn += 1024;
    split_o = split; stages_o = stages; dyn_o = dyn; bstage_o = b_stage; btx_o = b_tx;
    if (dyn > max_dyn + 1024) return 1e30;
    int tmem_cols = 32;
    while (tmem_cols < bn) tmem_cols <<= 1;
    int occ = (227 * 1024) / (dyn + 1024);
    if (occ > 512 / tmem_cols) occ = 512 / tmem_cols;
    if (occ > 4) occ = 4;
    if (occ < 1) occ = 1;
    const double slots = static_cast<double>(sms) * occ;
    const double waves = static_cast<double>((ctas + static_cast<long long>(slots) - 1) / static_cast<long long>(slots));
    // operand bytes streamed from L2 by all CTAs
    const double bytes = (static_cast<double>(tn) * g.M + static_cast<double>(tiles_m) * tn * bn) * g.K * esz;
    const double active_sms = ctas < sms ? static_cast<double>(ctas) : static_cast<double>(sms);
    const double t_load = bytes / (77e9 * active_sms) * 1e6;            // us
    // epilogue time per CTA (us), measured: the plain store pass is ~10 ns per column; the BCE pass is bound
    // by SFU issue (3 MUFU per element, 16 rows x chunks per warp, independent of how many lanes are active:
    // ~5.5 us up to 128 columns, ~9 us beyond); the dgrad pass reads one more tensor
    double t_epi = (0.35 + 0.0105 * bn) * (atomic ? 1.3 : 1.0);
    if (e.kind == EPI_BCE) t_epi = bn <= 128 ? 6.5 : 12.0;
    if (dgrad_aux || e.kind == EPI_STORE_ACT) t_epi = 0.5 + 0.016 * bn;
    const double per_sm_ctas = static_cast<double>(ctas) / active_sms;
    // with co-residency the epilogues hide behind other CTAs' loads; one epilogue is always exposed
    const double t_epi_total = occ > 1 ? t_epi * (1.0 + 0.35 * (per_sm_ctas - 1.0)) : t_epi * per_sm_ctas;
    const double quant = waves * slots / static_cast<double>(ctas);     // tail-wave inefficiency
    return (t_load * (ctas > sms ? (0.5 + 0.5 * quant) : 1.0)) + t_epi_total + 1.2;
  };
  int block_n = g.block_n, split = 1, stages = 2, dyn = 0, b_stage = 0, b_tx = 0;
  if (block_n <= 0) {
    // sweep aid: MVAE_GEMM_BN_EPI2=112 forces the tile width of every BCE-epilogue GEMM, etc.
    char name[32];
    snprintf(name, sizeof(name), "MVAE_GEMM_BN_EPI%d", e.kind);
    block_n = env_int(name, 0);
  }
  if (block_n <= 0) {
    double best = 1e30;
    const int n_cap = 16 * ceil_div(g.N, 16);
    // heavy epilogues (BCE, dgrad+BatchNorm statistics) carry an auxiliary tile in shared memory and are bound by
    // SFU / issue rate per row, not per column: measured best at <= 128 columns (2 CTAs per SM, full lanes)
    const int bn_max = (e.kind == EPI_BCE || dgrad_aux) ? 128 : 256;
    for (int bn = 32; bn <= bn_max; bn += 16) {
      if (bn > n_cap && bn != 32) break;
      int sp, st, dy, bs, bt;
      const double t = plan_for(bn, sp, st, dy, bs, bt);
      if (t < best) {
        best = t;
        block_n = bn;
      }
    }
  }
  MVAE_REQUIRE(block_n >= 16 && block_n <= 256 && block_n % 16 == 0, "gemm: block_n %d invalid", block_n);
  MVAE_REQUIRE(plan_for(block_n, split, stages, dyn, b_stage, b_tx) < 1e29, "gemm: tile %d does not fit in shared memory", block_n);
  const int tiles_n = ceil_div(g.N, block_n);
  const int kb_per_split = ceil_div(kb_total, split);
  if (env_int("MVAE_GEMM_VERBOSE", 0))
    fprintf(stderr, "[mvae gemm] %dx%dx%d kind=%d epi=%d a_mn=%d b_mn=%d -> block_n=%d split=%d stages=%d smem=%d grid=%dx%dx%d\n",
            g.M, g.N, g.K, g.kind, e.kind, g.a_mn, g.b_mn, block_n, split, stages, dyn, tiles_n, tiles_m, split);

  CUtensorMap ta, tb;
  if (cg.mode == 1 || cg.mode >= 3) {
  } else if (!g.a_mn) {
    if (make_tmap(&ta, g.kind, g.A, g.M, g.K, g.lda, BK, kBlockM, false)) return 1;
  } else {
    if (make_tmap(&ta, g.kind, g.A, g.K, g.M, g.lda, BK, BK, true)) return 1;
  }
  if (cg.mode == 2) {
  } else if (cg.mode >= 3) {
    // weights [C, kk*kk taps, N] (N contiguous, ldb elements between taps): one 64 x 1 x BK box per tap and column block
    if (make_tmap_w3(&tb, g.B, g.N, static_cast<long long>(cg.kk) * cg.kk, cg.C, g.ldb, BK, BK)) return 1;
  } else if (!g.b_mn) {
    if (make_tmap(&tb, g.kind, g.B, g.N, g.K, g.ldb, BK, block_n, false)) return 1;
  } else {
    if (make_tmap(&tb, g.kind, g.B, g.K, g.N, g.ldb, BK, BK, true)) return 1;
  }
  if (cg.mode == 1 || cg.mode >= 3) ta = tb;  // the gathered operand has no tensor map; the kernel never touches this copy
  if (cg.mode == 2) tb = ta;

  GemmKParams kp;
  kp.M = g.M; kp.N = g.N; kp.K = g.K;
  kp.block_n = block_n;
  kp.a_mn = g.a_mn; kp.b_mn = g.b_mn;
  kp.stages = stages;
  kp.kb_total = kb_total;
  kp.kb_per_split = kb_per_split;
  kp.b_stage_bytes = b_stage;
  kp.b_tx_bytes = b_tx;
  kp.epi = e;
  kp.stat_group_stride = (e.kind == EPI_BCE) ? 0 : g.N;
  kp.dbg = g.dbg != nullptr ? g.dbg : (g_dbg_epi == e.kind ? g_dbg_times : nullptr);
  auto al = [](const void* p, int a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
  const int ea = 16 / 4 * (e.c_dtype == MVAE_F32 ? 4 : 2);  // bytes for a 4-element vector of C
  bool vec = (e.ldc % 4 == 0) && al(e.C, ea) && al(e.probs, ea);
  const int aa = 4 * esz;
  if (e.kind == EPI_BCE) vec = vec && (e.ldt % 4 == 0) && al(e.target, aa);
  if (dgrad_aux) vec = vec && (e.ldh % 4 == 0) && al(e.hpre, aa);
  kp.vec_ok = vec ? 1 : 0;
  {
    static const int use_direct = env_int("MVAE_GEMM_DIRECT_STORE", 1);
    const int cvec = e.c_dtype == MVAE_F32 ? 4 : 8;   // elements per 16-byte store
    kp.direct_store = (use_direct != 0 && e.kind == EPI_STORE && e.stat0 == nullptr &&
                       e.ldc % cvec == 0 && al(e.C, 16)) ? 1 : 0;
  }
  const int tail0 = dyn - 1024 - aux_bytes(block_n) - red_bytes(block_n);
  kp.aux_off = (aux_ok && vec) ? tail0 : -1;
  kp.red_off = red_ok ? tail0 + aux_bytes(block_n) : -1;
  kp.gather = cg;
  if (cg.mode >= 3) {
    kp.gather.mode = 1;  // inside the kernel: A is the gathered operand (instantiation kGather = 2 / 3)
    MVAE_REQUIRE(kp.direct_store, "gemm: transposed-conv class needs the direct store epilogue (ldc %% 8 == 0, aligned C, no statistics)");
  }
  if (cg.mode != 0) {  // exact division by multiply-shift: q = (m * magic) >> 40 for m * d < 2^40
    kp.gather.magic_hw = (1ull << 40) / static_cast<unsigned long long>(cg.Ho * cg.Wo) + 1;
    kp.gather.magic_w = (1ull << 40) / static_cast<unsigned long long>(cg.Wo) + 1;
  }
  if (e.kind == EPI_BCE) MVAE_REQUIRE(e.target != nullptr && e.target_rows > 0, "gemm: BCE epilogue needs a target");
  if (e.kind == EPI_DGRAD_BN)
    MVAE_REQUIRE(e.hpre && e.bn_mean && e.bn_rstd && e.bn_gamma && e.bn_beta, "gemm: dgrad-BN epilogue needs BN state");
  MVAE_REQUIRE(e.rows_per_group > 0, "gemm: rows_per_group must be positive");

  dim3 grid(tiles_n, tiles_m, cg.mode == 4 ? cg.sc_stride * cg.sc_stride : split);  // mode 4: z = parity class (split is 1)
  if (cg.mode == 3) return launch_inst<MVAE_BF16, EPI_STORE, 2>(ta, tb, kp, grid, dyn, stream);
  if (cg.mode == 4) return launch_inst<MVAE_BF16, EPI_STORE, 3>(ta, tb, kp, grid, dyn, stream);
  if (cg.mode != 0) {  // bf16 store / accumulate only (validated above)
    if (e.kind == EPI_STORE) return launch_inst<MVAE_BF16, EPI_STORE, 1>(ta, tb, kp, grid, dyn, stream);
    return launch_inst<MVAE_BF16, EPI_ATOMIC, 1>(ta, tb, kp, grid, dyn, stream);
  }
#define MVAE_GEMM_CASE(KIND, EPI)                                   \
  if (g.kind == KIND && e.kind == EPI) return launch_inst<KIND, EPI>(ta, tb, kp, grid, dyn, stream);
  MVAE_GEMM_CASE(MVAE_F32, EPI_STORE)
  MVAE_GEMM_CASE(MVAE_F32, EPI_ATOMIC)
  MVAE_GEMM_CASE(MVAE_F32, EPI_BCE)
  MVAE_GEMM_CASE(MVAE_F32, EPI_DGRAD_BN)
  MVAE_GEMM_CASE(MVAE_BF16, EPI_STORE)
  MVAE_GEMM_CASE(MVAE_BF16, EPI_ATOMIC)
  MVAE_GEMM_CASE(MVAE_BF16, EPI_BCE)
  MVAE_GEMM_CASE(MVAE_BF16, EPI_DGRAD_BN)
  MVAE_GEMM_CASE(MVAE_F32, EPI_STORE_ACT)
  MVAE_GEMM_CASE(MVAE_F32, EPI_DGRAD_ACT)
  MVAE_GEMM_CASE(MVAE_BF16, EPI_STORE_ACT)
  MVAE_GEMM_CASE(MVAE_BF16, EPI_DGRAD_ACT)
#undef MVAE_GEMM_CASE
  set_error("gemm: unsupported kind/epilogue %d/%d", g.kind, e.kind);
  return 1;
}

}  // namespace mvae
