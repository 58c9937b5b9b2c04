// tcgen05 / TMEM / TMA GEMM for the MVAE Linear layers (forward, dgrad, wgrad) with the
// BatchNorm-statistics, BCE-with-logits and ReLU/BN-backward epilogues fused in.
//
//   C[M,N] = A[M,K] * B[N,K]^T        (fp32 accumulate in TMEM)
//
// One CTA = one 128 x block_n output tile (UMMA M=128, N=block_n, cta_group::1), 6 warps:
//   warp 0   : TMA producer (one elected lane) - A/B tiles into a SWIZZLE_128B smem ring
//   warp 1   : MMA issuer   (one elected lane) - tcgen05.mma kind::tf32 / kind::f16(bf16)
//   warps 2-5: epilogue     - tcgen05.ld TMEM -> registers -> padded smem tile -> row pass with
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

#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace mvae {

namespace {

constexpr int kBlockM = 128;
constexpr int kAStageBytes = kBlockM * 128;  // 16 KB: 128 rows x 128 B (either major)
constexpr int kGemmThreads = 192;
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
  int stat_group_stride;
  GemmEpilogue epi;
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
    if constexpr (sizeof(T) == 4) {
      float4 t = *reinterpret_cast<const float4*>(p);
      o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
      uint2 t = *reinterpret_cast<const uint2*>(p);
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
__device__ __forceinline__ void store4_dyn(void* base, long long off, int dtype, bool vec, int n_valid,
                                           const float (&v)[4]) {
  if (dtype == MVAE_F32)
    store4(reinterpret_cast<float*>(base) + off, vec, n_valid, v);
  else
    store4(reinterpret_cast<__nv_bfloat16*>(base) + off, vec, n_valid, v);
}

template <int kKind, int kEpi>
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
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * p.block_n;
  const int m0 = blockIdx.y * kBlockM;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int nkb = min(p.kb_per_split, p.kb_total - kb0);
  const int stage_bytes = kAStageBytes + p.b_stage_bytes;
  const int S = p.stages;

  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(p.block_n)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
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

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int a_boxes = kBlockM / ATOM;
      const int b_boxes = (p.block_n + ATOM - 1) / ATOM;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        ptx::mbar_expect_tx(&full_bar[s], kAStageBytes + p.b_tx_bytes);
        uint8_t* sa = smem + s * stage_bytes;
        uint8_t* sb = sa + kAStageBytes;
        const int kc = (kb0 + i) * BK;
        if (!p.a_mn) {
          ptx::tma_load_2d(sa, &tmA, &full_bar[s], kc, m0);
        } else {
          for (int j = 0; j < a_boxes; ++j) ptx::tma_load_2d(sa + j * (BK * 128), &tmA, &full_bar[s], m0 + j * ATOM, kc);
        }
        if (!p.b_mn) {
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
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
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
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const GemmEpilogue& e = p.epi;
    float* stage = reinterpret_cast<float*>(smem);  // aliases the (drained) operand ring
    const int ldst = p.block_n + kStagePad;

    ptx::mbar_wait(&accum_bar, 0);
    ptx::tc_fence_after();
    {
      const int row = q * 32 + lane;
      float* dst_row = stage + row * ldst;
      for (int c = 0; c < p.block_n; c += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
        ptx::tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(dst_row + c);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
      }
    }
    ptx::tc_fence_before();
    ptx::named_bar_sync(1, 128);

    // Row pass: this warp owns 32 consecutive rows, each lane 4 consecutive columns per 128-col chunk.
    const int ew = warp - 2;
    const bool vec = p.vec_ok != 0;
    const int nch = (p.block_n + 127) / 128;
    float acc0[2][4], acc1[2][4], bias[2][4];
    float g_mean[2][4], g_rstd[2][4], g_gamma[2][4], g_beta[2][4];
    float lsum = 0.f, g_scale = 0.f;
    int nval[2], coln[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int col = ch * 128 + lane * 4;
      coln[ch] = n0 + col;
      int nv = 0;
      if (ch < nch && col < p.block_n) nv = min(4, p.N - coln[ch]);
      nval[ch] = nv < 0 ? 0 : nv;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc0[ch][i] = 0.f;
        acc1[ch][i] = 0.f;
        bias[ch][i] = (e.bias != nullptr && i < nval[ch]) ? e.bias[coln[ch] + i] : 0.f;
        g_mean[ch][i] = g_rstd[ch][i] = g_gamma[ch][i] = g_beta[ch][i] = 0.f;
        if (kEpi == EPI_DGRAD_BN && i < nval[ch]) {
          g_gamma[ch][i] = e.bn_gamma[coln[ch] + i];
          g_beta[ch][i] = e.bn_beta[coln[ch] + i];
        }
      }
    }

    auto flush = [&](int g) {
      if (e.stat0 != nullptr) {
        const long long goff = static_cast<long long>(g) * p.stat_group_stride;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < nval[ch]) {
              atomicAdd(e.stat0 + goff + coln[ch] + i, acc0[ch][i]);
              if (e.stat1 != nullptr) atomicAdd(e.stat1 + goff + coln[ch] + i, acc1[ch][i]);
            }
            acc0[ch][i] = 0.f;
            acc1[ch][i] = 0.f;
          }
        }
      }
      if (kEpi == EPI_BCE) {
        float s = lsum;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && e.loss != nullptr) atomicAdd(e.loss + g, s);
        lsum = 0.f;
      }
    };

    int cur_g = -1;
    for (int rr = 0; rr < 32; ++rr) {
      const int r = ew * 32 + rr;
      const int m = m0 + r;
      if (m >= p.M) break;
      const int g = m / e.rows_per_group;
      if (g != cur_g) {
        if (cur_g >= 0) flush(cur_g);
        cur_g = g;
        if (kEpi == EPI_BCE) g_scale = e.bce_scale[g & 3];
        if (kEpi == EPI_DGRAD_BN) {
#pragma unroll
          for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < nval[ch]) {
                g_mean[ch][i] = e.bn_mean[static_cast<long long>(g) * p.N + coln[ch] + i];
                g_rstd[ch][i] = e.bn_rstd[static_cast<long long>(g) * p.N + coln[ch] + i];
              }
        }
      }
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        if (nval[ch] == 0) continue;
        const int col = ch * 128 + lane * 4;
        const float4 t = *reinterpret_cast<const float4*>(stage + r * ldst + col);
        float v[4] = {t.x, t.y, t.z, t.w};
        const long long coff = static_cast<long long>(m) * e.ldc + coln[ch];
        if constexpr (kEpi == EPI_STORE) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[i] += bias[ch][i];
            acc0[ch][i] += v[i];
            acc1[ch][i] += v[i] * v[i];
          }
          store4_dyn(e.C, coff, e.c_dtype, vec, nval[ch], v);
        } else if constexpr (kEpi == EPI_ATOMIC) {
          float* c = reinterpret_cast<float*>(e.C) + coff;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nval[ch]) atomicAdd(c + i, v[i]);
        } else if constexpr (kEpi == EPI_BCE) {
          float tg[4], d[4], pr[4];
          const act_t* tp =
              reinterpret_cast<const act_t*>(e.target) + static_cast<long long>(m % e.target_rows) * e.ldt + coln[ch];
          load4(tp, vec, nval[ch], tg);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x = v[i] + bias[ch][i];
            const float ex = __expf(-fabsf(x));
            const float inv = 1.f / (1.f + ex);
            const float pz = x >= 0.f ? inv : ex * inv;
            pr[i] = pz;
            d[i] = g_scale * (pz - tg[i]);
            if (i < nval[ch]) {
              lsum += g_scale * (fmaxf(x, 0.f) - tg[i] * x + log1pf(ex));
              acc0[ch][i] += d[i];
            }
          }
          store4_dyn(e.C, coff, e.c_dtype, vec, nval[ch], d);
          if (e.probs != nullptr) store4_dyn(e.probs, coff, e.c_dtype, vec, nval[ch], pr);
        } else if constexpr (kEpi == EPI_DGRAD_BN) {
          float h[4], d[4];
          const act_t* hp = reinterpret_cast<const act_t*>(e.hpre) + static_cast<long long>(m) * e.ldh + coln[ch];
          load4(hp, vec, nval[ch], h);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float xh = (h[i] - g_mean[ch][i]) * g_rstd[ch][i];
            const float y = fmaf(g_gamma[ch][i], xh, g_beta[ch][i]);
            d[i] = y > 0.f ? v[i] : 0.f;
            acc0[ch][i] += d[i];
            acc1[ch][i] += d[i] * xh;
          }
          store4_dyn(e.C, coff, e.c_dtype, vec, nval[ch], d);
        }
      }
    }
    if (cur_g >= 0) flush(cur_g);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, tmem_cols);
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

template <int kKind, int kEpi>
int launch_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKParams& kp, dim3 grid, int dyn_smem,
                cudaStream_t stream) {
  static int smem_set = 0;  // per instantiation; monotone, benign race
  if (dyn_smem > smem_set) {
    MVAE_CUDA(cudaFuncSetAttribute(gemm_kernel<kKind, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem));
    smem_set = dyn_smem;
  }
  gemm_kernel<kKind, kEpi><<<grid, kGemmThreads, dyn_smem, stream>>>(ta, tb, kp);
  MVAE_CUDA(cudaGetLastError());
  return 0;
}

int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace

int launch_gemm(const GemmDesc& g, cudaStream_t stream) {
  const int esz = g.kind == MVAE_F32 ? 4 : 2;
  const int BK = 128 / esz;
  MVAE_REQUIRE(g.kind == MVAE_F32 || g.kind == MVAE_BF16, "gemm: bad kind %d", g.kind);
  MVAE_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem %dx%dx%d", g.M, g.N, g.K);
  MVAE_REQUIRE((g.lda * esz) % 16 == 0 && (g.ldb * esz) % 16 == 0, "gemm: lda/ldb (%lld,%lld) must be 16-byte multiples",
               g.lda, g.ldb);
  MVAE_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0,
               "gemm: A/B must be 16-byte aligned");
  const GemmEpilogue& e = g.epi;
  MVAE_REQUIRE(e.C != nullptr, "gemm: null output");
  MVAE_REQUIRE(e.kind != EPI_ATOMIC || e.c_dtype == MVAE_F32, "gemm: atomic epilogue needs fp32 output");

  const int tiles_m = ceil_div(g.M, kBlockM);
  const int kb_total = ceil_div(g.K, BK);

  // ---- tile width: fewest N tiles that still give the machine enough CTAs
  int block_n = g.block_n;
  if (block_n <= 0) {
    const int target_ctas = env_int("MVAE_GEMM_TARGET_CTAS", 120);
    int nt = ceil_div(g.N, 256);
    for (;; ++nt) {
      block_n = 16 * ceil_div(ceil_div(g.N, nt), 16);
      int splits_possible = (e.kind == EPI_ATOMIC) ? kb_total : 1;
      if (static_cast<long long>(tiles_m) * nt * splits_possible >= target_ctas) break;
      if (block_n <= 48) break;
    }
  }
  MVAE_REQUIRE(block_n >= 16 && block_n <= 256 && block_n % 16 == 0, "gemm: block_n %d invalid", block_n);
  const int tiles_n = ceil_div(g.N, block_n);

  // ---- split-K (wgrad): enough CTAs to cover the chip about once
  int split = g.split_k;
  if (e.kind != EPI_ATOMIC) split = 1;
  if (split <= 0) {
    const int target = env_int("MVAE_GEMM_SPLIT_TARGET", 148);
    split = ceil_div(target, static_cast<long long>(tiles_m) * tiles_n);
  }
  if (split > kb_total) split = kb_total;
  if (split < 1) split = 1;
  const int kb_per_split = ceil_div(kb_total, split);
  split = ceil_div(kb_total, kb_per_split);

  // ---- smem ring
  const int b_boxes = ceil_div(block_n, BK);
  const int b_tx = g.b_mn ? b_boxes * BK * 128 : block_n * 128;
  const int b_stage = (b_tx + 1023) / 1024 * 1024;
  const int stage_bytes = kAStageBytes + b_stage;
  const int staging = kBlockM * (block_n + kStagePad) * 4;
  int stages = g.stages > 0 ? g.stages : env_int("MVAE_GEMM_STAGES", 4);
  if (stages > kb_per_split) stages = kb_per_split;
  if (stages > kMaxStages) stages = kMaxStages;
  const int max_dyn = 227 * 1024 - 2048;
  while (stages > 1 && stages * stage_bytes + 1024 > max_dyn) --stages;
  int dyn = stages * stage_bytes;
  if (dyn < staging) dyn = staging;
  dyn += 1024;
  MVAE_REQUIRE(dyn <= max_dyn + 1024, "gemm: smem %d too large", dyn);

  CUtensorMap ta, tb;
  if (!g.a_mn) {
    if (make_tmap(&ta, g.kind, g.A, g.M, g.K, g.lda, BK, kBlockM, false)) return 1;
  } else {
    if (make_tmap(&ta, g.kind, g.A, g.K, g.M, g.lda, BK, BK, true)) return 1;
  }
  if (!g.b_mn) {
    if (make_tmap(&tb, g.kind, g.B, g.N, g.K, g.ldb, BK, block_n, false)) return 1;
  } else {
    if (make_tmap(&tb, g.kind, g.B, g.K, g.N, g.ldb, BK, BK, true)) return 1;
  }

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
  auto al = [](const void* p, int a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
  const int ea = 16 / 4 * (e.c_dtype == MVAE_F32 ? 4 : 2);  // bytes for a 4-element vector of C
  bool vec = (e.ldc % 4 == 0) && al(e.C, ea) && al(e.probs, ea);
  const int aa = 4 * esz;
  if (e.kind == EPI_BCE) vec = vec && (e.ldt % 4 == 0) && al(e.target, aa);
  if (e.kind == EPI_DGRAD_BN) vec = vec && (e.ldh % 4 == 0) && al(e.hpre, aa);
  kp.vec_ok = vec ? 1 : 0;
  if (e.kind == EPI_BCE) MVAE_REQUIRE(e.target != nullptr && e.target_rows > 0, "gemm: BCE epilogue needs a target");
  if (e.kind == EPI_DGRAD_BN)
    MVAE_REQUIRE(e.hpre && e.bn_mean && e.bn_rstd && e.bn_gamma && e.bn_beta, "gemm: dgrad-BN epilogue needs BN state");
  MVAE_REQUIRE(e.rows_per_group > 0, "gemm: rows_per_group must be positive");

  dim3 grid(tiles_n, tiles_m, split);
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
#undef MVAE_GEMM_CASE
  set_error("gemm: unsupported kind/epilogue %d/%d", g.kind, e.kind);
  return 1;
}

}  // namespace mvae
