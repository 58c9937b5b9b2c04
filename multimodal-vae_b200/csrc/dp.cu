// Data-parallel gradient exchange fused with the optimizer, over NVLink peer memory (no NCCL on the data path):
// ONE launch per gradient bucket does
//     barrier A (every rank's bucket is complete)
//  -> reduce-scatter: rank r sums slice r of the bucket over all ranks' gradient buffers (peer loads)
//  -> all-gather: and writes the sum back into every rank's buffer (peer stores)
//  -> barrier B (every slice has landed everywhere)
//  -> Adam on the whole bucket from the local, now identical, summed gradients (grad_scale = 1 / world).
// The gradient buffers and a small flag array per rank live in symmetric memory (torch.distributed._symmetric_memory
// gives every rank the peer pointers); everything else (parameters, moments) is local.  Every rank computes the same
// update from bit-identical sums, so the replicas stay bit-identical.  The mean over equal shards of per-shard means is
// the global mean, which is what mnist/train.py:127-153 computes on one device.
//
// Flags (uint32[32] per rank): [0..7] barrier A arrivals by source rank, [8..15] barrier B arrivals, [16] epoch (calls
// completed), [17] / [18] local "go" words of the two barriers, [19] / [20] local block counters, [21] error.
// All spins are bounded (~2 s): a lost rank flags an error instead of hanging the GPU.
#include "../../include/mvae_b200.h"
#include "common.cuh"

namespace mvae {
namespace {

constexpr int kDpThreads = 512;
constexpr int kDpMaxWorld = 8;

struct DpParams {
  int world, rank;
  float* grads[kDpMaxWorld];
  unsigned int* flags[kDpMaxWorld];
  float* params;
  float* adam_m;
  float* adam_v;
  __nv_bfloat16* params_bf16;
  long long lo4, hi4;   // bucket as float4 indices
  float lr, b1, b2, eps, grad_scale;
  const int* step_ptr;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long dp_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait until *p >= want (system scope when `sys`); false after ~2 s
__device__ __forceinline__ bool spin_ge(const unsigned int* p, unsigned int want, bool sys) {
  const unsigned long long t0 = dp_timer();
  unsigned int spins = 0;
  while (true) {
    const unsigned int v = sys ? ld_acquire_sys(p) : ld_acquire_gpu(p);
    if (static_cast<int>(v - want) >= 0) return true;
    if ((++spins & 63u) == 0 && dp_timer() - t0 > 2000000000ull) return false;
    __nanosleep(20);
  }
}

__global__ void __launch_bounds__(kDpThreads) dp_reduce_adam_kernel(const DpParams p) {
  unsigned int* F = p.flags[p.rank];
  __shared__ unsigned int s_epoch;
  __shared__ int s_last;
  const int tid = threadIdx.x;
  if (tid == 0) s_epoch = ld_acquire_gpu(F + 16) + 1u;
  __syncthreads();
  const unsigned int e = s_epoch;

  // ---- barrier A: every rank has entered the call, i.e. its gradients of the bucket are final
  if (blockIdx.x == 0) {
    if (tid < p.world) {
      st_release_sys(p.flags[tid] + p.rank, e);
      if (!spin_ge(F + tid, e, true)) atomicExch(F + 21, 0xA000u | tid);
    }
    __syncthreads();
    if (tid == 0) st_release_gpu(F + 17, e);
  } else {
    if (tid == 0 && !spin_ge(F + 17, e, false)) atomicExch(F + 21, 0xA100u);
    __syncthreads();
  }

  // ---- reduce-scatter + all-gather of this rank's slice
  const long long n4 = p.hi4 - p.lo4;
  const long long slice = (n4 + p.world - 1) / p.world;
  const long long s0 = p.lo4 + slice * p.rank;
  const long long s1 = (s0 + slice < p.hi4) ? s0 + slice : p.hi4;
  for (long long i = s0 + blockIdx.x * static_cast<long long>(kDpThreads) + tid; i < s1; i += static_cast<long long>(gridDim.x) * kDpThreads) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kDpMaxWorld; ++r) {
      if (r < p.world) {   // fixed rank order: every owner adds in the same order
        const float4 v = __ldcg(reinterpret_cast<const float4*>(p.grads[r]) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
#pragma unroll
    for (int r = 0; r < kDpMaxWorld; ++r)
      if (r < p.world) __stcg(reinterpret_cast<float4*>(p.grads[r]) + i, acc);
  }
  __threadfence_system();
  __syncthreads();

  // ---- barrier B: all slices have landed in every rank's buffer
  if (tid == 0) s_last = (atomicAdd(F + 19, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (s_last) {
    if (tid == 0) F[19] = 0u;
    if (tid < p.world) {
      st_release_sys(p.flags[tid] + 8 + p.rank, e);
      if (!spin_ge(F + 8 + tid, e, true)) atomicExch(F + 21, 0xB000u | tid);
    }
    __syncthreads();
    if (tid == 0) st_release_gpu(F + 18, e);
  } else {
    if (tid == 0 && !spin_ge(F + 18, e, false)) atomicExch(F + 21, 0xB100u);
    __syncthreads();
  }

  // ---- Adam on the bucket (torch.optim.Adam defaults semantics, as adam_kernel in elementwise.cu)
  const float step = static_cast<float>(*p.step_ptr);
  const float bc1 = 1.f - powf(p.b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(p.b2, step));
  const float* g = p.grads[p.rank];
  for (long long i = p.lo4 + blockIdx.x * static_cast<long long>(kDpThreads) + tid; i < p.hi4; i += static_cast<long long>(gridDim.x) * kDpThreads) {
    const float4 gv = __ldcg(reinterpret_cast<const float4*>(g) + i);
    float4 pv = reinterpret_cast<const float4*>(p.params)[i];
    float4 mv = reinterpret_cast<const float4*>(p.adam_m)[i];
    float4 vv = reinterpret_cast<const float4*>(p.adam_v)[i];
    float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w}, v2[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = gg[k] * p.grad_scale;
      mm[k] = p.b1 * mm[k] + (1.f - p.b1) * gr;
      v2[k] = p.b2 * v2[k] + (1.f - p.b2) * gr * gr;
      const float denom = sqrtf(v2[k]) / bc2_sqrt + p.eps;
      pp[k] -= (p.lr / bc1) * (mm[k] / denom);
    }
    reinterpret_cast<float4*>(p.params)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(p.adam_m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(p.adam_v)[i] = make_float4(v2[0], v2[1], v2[2], v2[3]);
    if (p.params_bf16 != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp[0], pp[1]), hi = __floats2bfloat162_rn(pp[2], pp[3]);
      uint2 w;
      w.x = *reinterpret_cast<uint32_t*>(&lo);
      w.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(p.params_bf16)[i] = w;
    }
  }
  // ---- the call is complete once every block is through: the last one publishes the epoch
  __syncthreads();
  if (tid == 0 && atomicAdd(F + 20, 1u) == gridDim.x - 1) {
    F[20] = 0u;
    st_release_gpu(F + 16, e);
  }
}

}  // namespace
}  // namespace mvae

using namespace mvae;

extern "C" int mvae_dp_reduce_adam(const mvae_dp_reduce_adam_args* a, void* stream) {
  MVAE_REQUIRE(a != nullptr, "dp_reduce_adam: null args");
  MVAE_REQUIRE(a->world >= 2 && a->world <= kDpMaxWorld && a->rank >= 0 && a->rank < a->world, "dp_reduce_adam: world=%d rank=%d", a->world,
               a->rank);
  MVAE_REQUIRE(a->lo % 4 == 0 && a->hi % 4 == 0 && a->lo >= 0 && a->hi > a->lo, "dp_reduce_adam: bucket [%lld, %lld) must be float4-aligned",
               (long long)a->lo, (long long)a->hi);
  MVAE_REQUIRE(a->params && a->adam_m && a->adam_v && a->adam_step, "dp_reduce_adam: optimizer state missing");
  DpParams p;
  p.world = a->world; p.rank = a->rank;
  for (int r = 0; r < kDpMaxWorld; ++r) {
    p.grads[r] = r < a->world ? static_cast<float*>(a->grads[r]) : nullptr;
    p.flags[r] = r < a->world ? static_cast<unsigned int*>(a->flags[r]) : nullptr;
    MVAE_REQUIRE(r >= a->world || (p.grads[r] != nullptr && p.flags[r] != nullptr), "dp_reduce_adam: peer pointer %d missing", r);
  }
  p.params = a->params; p.adam_m = a->adam_m; p.adam_v = a->adam_v; p.params_bf16 = static_cast<__nv_bfloat16*>(a->params_bf16);
  p.lo4 = a->lo / 4; p.hi4 = a->hi / 4;
  p.lr = a->lr; p.b1 = a->beta1; p.b2 = a->beta2; p.eps = a->eps; p.grad_scale = a->grad_scale;
  p.step_ptr = a->adam_step;
  // cooperative launch: the grid starts only when all its blocks can be resident, which the in-kernel barriers need
  // (the decoder bucket's call runs beside the encoder-side backward)
  int blocks = a->blocks > 0 ? a->blocks : 96;
  if (blocks > 128) blocks = 128;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(kDpThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MVAE_CUDA(cudaLaunchKernelEx(&cfg, dp_reduce_adam_kernel, p));
  note_launch(1);
  return 0;
}
