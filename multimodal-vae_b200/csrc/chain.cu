// Slab-persistent layer chains of the MNIST MVAE step (bf16): ONE CTA owns a 128-row slab of the batch and carries it
// through a whole chain of Linear layers - the activation never leaves the SM between layers.
//
//   enc_fwd : image -> Linear 784->400 -> BN+ReLU -> Linear 400->200 -> BN+ReLU -> Linear 200->2n      (mnist/model.py:99-117)
//   dec_fwd : z -> Linear n->200 -> BN+ReLU -> Linear 200->400 -> BN+ReLU -> Linear 400->784 + sigmoid/BCE/dlogits
//                                                                            (mnist/model.py:120-135, mnist/train.py:70)
//   dec_bwd : dlogits -> dgrad 784->400 -> ReLU/BN backward -> dgrad 400->200 -> ReLU/BN backward -> dgrad 200->n
//   enc_bwd : d(enc) -> dgrad 2n->200 -> ReLU/BN backward -> dgrad 200->400 -> ReLU/BN backward
//
// Roles (576 threads): warp 0 = TMA producer (weights - and the streamed A operand of the two K = 784 layers - into a
// SWIZZLE_128B shared-memory ring, running ahead across layer boundaries), warp 1 = tcgen05.mma issuer (fp32 accumulators
// in TMEM, a whole layer of up to 400 columns, or a ring of four 112-column chunks for the 784-wide layer), warps 2..17 =
// epilogue: four warps per TMEM lane quarter, each taking every fourth 32-column unit of its 32 rows.  The epilogue of a
// layer writes the NEXT layer's A operand straight into shared memory in the UMMA K-major swizzled layout (the "arena":
// 7 panels of 128 rows x 64 columns), so `h = relu(bn(x W^T + b))` feeds the next tcgen05.mma without touching HBM; what
// the weight-gradient GEMMs and the backward need (pre-/post-BatchNorm activations, gradients) is written to global
// memory once, on the side.
//
// tcgen05.ld hands a lane one accumulator ROW; everything the epilogue does is per COLUMN (coefficients, statistics, 64-
// byte global segments per row), so each unit is transposed through a 2 KB per-warp tile (16 rows x 32 fp32, 16-byte
// chunks XOR-swizzled: conflict-free both ways): afterwards lane (rsel, cq) holds columns 4cq..4cq+3 of rows 4rsel..4rsel+3.
//
// BatchNorm needs whole-batch statistics per ELBO term: the four warps of a column unit leave their slab sums in a
// shared table, one global atomic per column and CTA publishes them, all CTAs of a term meet at a grid barrier (arrival
// counter in the step workspace, cooperative launch guarantees co-residency), read the sums back and normalise their own
// slab: the forward from the bf16 pre-activations it stashed in the arena (in place), the backward from TMEM.  One
// barrier per BatchNorm layer and direction.  One extra CTA (the "keeper") owns no slab: it waits for every group's
// barrier and does what needs all groups (running statistics in group order, dgamma / dbeta) beside the others.
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace mvae {

namespace {

constexpr int kEpiWarps = 16;
constexpr int kCEpi = kEpiWarps * 32;           // 512 epilogue threads
constexpr int kCThreads = 64 + kCEpi;           // 576
constexpr int kPanel = 16384;                   // 128 rows x 128 B: one K-major SWIZZLE_128B panel (64 bf16 columns)
constexpr int kArenaPanels = 7;                 // 448 columns >= the widest resident activation (400)
constexpr int kArena = kArenaPanels * kPanel;   // 114688
constexpr int kSmemTotal = 232448;              // 227 KB, no static shared memory
constexpr int kBarOff = kSmemTotal - 192;       // mbarriers + the TMEM base address
constexpr int kRedStride = 800;                 // floats per TMEM lane quarter in the reduction table
constexpr int kRedOff = kBarOff - 4 * kRedStride * 4;
constexpr int kTabFloats = 1600;                // coefficient tables: 4 x 400 floats (or one of 800)
constexpr int kTabOff = kRedOff - kTabFloats * 4;
constexpr int kTileBytes = 2048;                // per epilogue warp: 16 rows x 32 fp32
constexpr int kTileOff = kTabOff - kEpiWarps * kTileBytes;
constexpr int kRing1Stage = 32768;              // resident-A layers: one weight chunk (<= 224 rows K-major / four 64x64 boxes MN-major)
constexpr int kRing1Stages = 2;
static_assert(kArena + kRing1Stages * kRing1Stage <= kTileOff, "ring 1 overlaps the epilogue tiles");
constexpr int kMaxPass = 16;
constexpr int kMaxLayer = 3;

enum : int { CE_FWD_BN = 0, CE_FWD_STORE = 1, CE_BCE = 2, CE_DGRAD_BN = 3, CE_DGRAD_STORE = 4 };

struct CRing {
  int base, stage_bytes, stages, a_bytes;
};

// One pass of the producer / MMA pipelines: a K loop over 64-wide panels for one or two accumulator chunks.
struct CPass {
  int cfg;                // ring configuration
  int k_panels;           // 64-wide contraction panels
  int last_ksteps;        // 16-wide k-steps in the last panel (1..4)
  int a_stream;           // tensor map of the streamed A operand (one [128 x 64] tile per stage), or -1: resident in the arena
  int a_wait;             // 0: nothing, 1: wait for the TMA-loaded resident A, 2: wait for the A written by the epilogue
  int b_tm;               // tensor map of B
  int b_mn;               // 0: B tile rows are output columns (K-major); 1: MN-major 64 x 64 boxes
  int n_chunks;
  int c_n0[2];            // first output column of the chunk
  int c_mma_n[2];         // UMMA N
  int c_boxes[2];         // MN-major: 64-column boxes
  int c_bytes[2];         // bytes TMA writes per stage for the chunk
  int c_boff[2];          // byte offset inside the stage's B slot
  int c_tmem[2];          // TMEM column base
  int c_buf[2];           // accumulator barrier pair
  int c_half[2];          // CTA pairs: output columns of the chunk each CTA supplies as B operand (c_mma_n / 2), else 0
};

struct CLayer {
  int kind, N, n_chunks;   // accumulator chunk ci uses barrier pair ci (BatchNorm / store layers: TMEM column = layer column)
  int u1, u2, u3;          // BCE layer: first 32-column unit of chunks 1, 2, 3 (chunk ci sits in TMEM buffer ci & 1)
  int buf_cols;            // BCE layer: TMEM columns per accumulator buffer
  const float* bias;
  const float* gamma;
  const float* beta;
  float* stat0;            // [G][N] forward: sum y / backward: sum dyhat
  float* stat1;            // [G][N] forward: sum y^2 / backward: sum dyhat * xhat
  float* save_mean;        // [G][N] forward: written; backward: read
  float* save_rstd;
  float* running_mean;     // forward, keeper CTA
  float* running_var;
  int bn_updates;
  unsigned int* counter;   // [G] grid-barrier arrivals
  __nv_bfloat16* out_pre;  // forward BN: pre-BatchNorm activations [R][N]
  __nv_bfloat16* out_post; // forward BN: post BN+ReLU [R][N]; backward BN: gradient at the layer's pre-BatchNorm output
  int write_arena;         // the result is the next layer's A operand
  const __nv_bfloat16* hpre;  // backward BN: pre-BatchNorm activations of the forward
  float* dgamma;
  float* dbeta;
  float* out_f32;          // store kinds
  int ld_out;
  // BCE
  const __nv_bfloat16* target;
  int target_rows;
  float bce_scale[3];
  float* loss;             // [G]
  float* dbias;            // [N] +=
  __nv_bfloat16* dlog;
  __nv_bfloat16* probs;
};

struct alignas(64) CParams {
  CUtensorMap tm[5];
  CUtensorMap tmo[4];     // outputs written by TMA from the arena: [2 * layer] pre-BatchNorm / [2 * layer + 1] post (forward),
                          // [2 * layer + 1] gradient at the pre-BatchNorm output (backward); boxes of 64 columns x 128 rows
  CRing ring[2];
  CPass pass[kMaxPass];
  CLayer layer[kMaxLayer];
  int n_pass, n_layers;
  int n_slabs;            // slabs; CTA n_slabs * parts is the keeper
  // Column split of ONE layer (enc_bwd's last: no later layer needs its output on chip): `parts` CTAs share a slab, CTA
  // `n_slabs * part + slab` takes columns [part_n0, part_n0 + part_n) of the split layer.  The layers before it are
  // computed by every part on its own (redundantly, on SMs that would idle otherwise) - only part 0 publishes their
  // statistics and results - so no activation ever moves between CTAs.
  int parts, split_layer, split_pass;
  int part_n0[4], part_n[4];
  int rows_per_group;     // rows of one statistics group (= batch)
  int init_tm, init_panels;  // resident A loaded by TMA at kernel start (z / d_enc), -1: none
  float momentum, eps;
  unsigned int* err;      // device flag: non-zero = a wait timed out (bring-up)
  long long* dbg;         // bring-up: [ctas][32] %globaltimer stamps (mvae_debug_chain_times), or null
};

// time-outs of the waits below count SM cycles (%clock64 is SM-local and cheap; %globaltimer is a slow read that would
// sit in every blocking wait): ~1.9 GHz, so 8e9 cycles ~ 4 s
__device__ __forceinline__ unsigned long long cyc() { return static_cast<unsigned long long>(clock64()); }
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// mbarrier wait that cannot hang the GPU: after ~4 s the kernel records where it was stuck and traps.
template <bool kPair = false>
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  if constexpr (kPair) return ptx::mbar_try_wait_cluster(bar, parity);   // arrivals come from both CTAs of the pair
  else return ptx::mbar_try_wait(bar, parity);
}
template <bool kPair = false>
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity, unsigned int* err, unsigned int code) {
  if (mbar_try<kPair>(bar, parity)) return;
  const unsigned long long t0 = cyc();
  unsigned int spins = 0;
  while (!mbar_try<kPair>(bar, parity)) {
    if ((++spins & 1023u) == 0 && cyc() - t0 > 8000000000ull) {
      if (err != nullptr) atomicExch(err, code);
      __threadfence();
      asm volatile("trap;");
    }
  }
}

// the same for the 16 epilogue warps: they back off between polls, so that their spinning does not take issue slots from
// the producer / MMA threads that share their schedulers
__device__ __forceinline__ void mbar_wait_epi(uint64_t* bar, uint32_t parity, unsigned int* err, unsigned int code) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = cyc();
  unsigned int spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if ((++spins & 255u) == 0 && cyc() - t0 > 8000000000ull) {
      if (err != nullptr) atomicExch(err, code);
      __threadfence();
      asm volatile("trap;");
    }
  }
}

// tcgen05.mma (bf16 x bf16 -> fp32) from shared-memory descriptors given as (low word, shared high word): SWIZZLE_128B,
// 8-row groups 1024 bytes apart, descriptor version 1.  The low word is (address >> 4) | (leading byte offset >> 4) << 16.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// the same for a CTA pair (cta_group::2, M = 256): issued by the leader, A = each CTA's own 128 rows at the same shared-memory
// offset, B = N / 2 output columns from each CTA, D = 128 lanes x N columns of TMEM in each CTA
__device__ __forceinline__ void umma_bf16_lohi_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void bar_epi() { ptx::named_bar_sync(1, kCEpi); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// packed fp32 pairs (sm_100 add / mul / fma .f32x2: one issue slot for two lanes of arithmetic)
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float4 tab4(const float* t) { return *reinterpret_cast<const float4*>(t); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float v[4];
  ptx::lds128(addr, v);
  return make_float4(v[0], v[1], v[2], v[3]);
}

// Grid barrier of one statistics group: arrive, then wait until all `expected` CTAs of the group have arrived.
// Called by all epilogue threads after their global atomics.  A lost CTA cannot hang the GPU: after ~2 s the wait
// gives up and flags the error (results are then garbage, which the flag reports).
__device__ __forceinline__ void group_barrier(unsigned int* counter, unsigned int expected, unsigned int* err, int et, bool arrive = true) {
  __threadfence();
  bar_epi();
  if (et == 0) {
    if (arrive) atomicAdd(counter, 1u);
    unsigned int seen = 0;
    const unsigned long long t0 = cyc();
    unsigned int spins = 0;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen >= expected) break;
      if ((++spins & 255u) == 0 && cyc() - t0 > 4000000000ull) {
        if (err != nullptr) atomicExch(err, 0xBA00u | (seen & 0xffu));
        break;
      }
      __nanosleep(32);
    }
    __threadfence();
  }
  bar_epi();
}
// wait (without arriving) until a group's counter is complete - the keeper CTA
__device__ __forceinline__ void group_wait(const unsigned int* counter, unsigned int expected, unsigned int* err) {
  unsigned int seen = 0;
  const unsigned long long t0 = cyc();
  unsigned int spins = 0;
  while (true) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    if (seen >= expected) break;
    if ((++spins & 255u) == 0 && cyc() - t0 > 4000000000ull) {
      if (err != nullptr) atomicExch(err, 0xBB00u | (seen & 0xffu));
      break;
    }
    __nanosleep(100);
  }
  __threadfence();
}

// ---------------------------------------------------------------- the epilogue's transposing tile and the arena
// tile: 16 rows x 128 B, 16-byte chunk c of row r at chunk c ^ (r & 7)
__device__ __forceinline__ uint32_t tile_addr(uint32_t tile, int r, int c) { return tile + r * 128 + (((c ^ r) & 7) << 4); }
// arena: K-major SWIZZLE_128B panels of 128 rows x 128 bytes; 16-byte chunk c of row r lives at chunk c ^ (r & 7)
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t w0, uint32_t w1) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(w0), "r"(w1) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 r;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr) : "memory");
  return r;
}

// The layer kinds are template parameters and the layer sequence is unrolled at compile time: `p.layer[il].x` is then a
// direct constant-bank operand (an indexed constant load behind an asm statement costs hundreds of cycles on the long
// scoreboard), and every instantiation carries only the epilogues it runs.
//
// kPair: two slab CTAs on neighbouring SMs (a cluster of two) run every Linear as ONE cta_group::2 tcgen05.mma of M = 256:
// each CTA keeps its own 128 rows (A operand, accumulator, epilogue - all as before) but loads only HALF of every weight
// tile, the tensor core reads the other half from the partner's shared memory.  The weight ring then holds twice the
// stages in the same bytes: the rings are latency-bound (a stage's turn-around is ~1.7 us whatever its size), so the
// weight streams run twice as fast, and the L2 -> SM traffic of the weights halves.  The leader (cluster rank 0) issues
// the MMAs and owns the barriers the MMA thread waits on; commits arrive on both CTAs' barriers (multicast).
template <int kKind0, int kKind1, int kKind2, bool kPair>
__global__ void __launch_bounds__(kCThreads, 1) chain_kernel(const __grid_constant__ CParams p) {
  // No static shared memory: the dynamic window then starts 1024-byte aligned (checked below), which the SWIZZLE_128B
  // tiles need, and every byte of the 227 KB is planned: [arena | weight ring | tiles | tables | reduction table | barriers].
  extern __shared__ __align__(1024) uint8_t smem[];
  float* tab = reinterpret_cast<float*>(smem + kTabOff);
  float* red = reinterpret_cast<float*>(smem + kRedOff);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* acc_full = full_bar + 8;
  uint64_t* acc_empty = full_bar + 12;
  uint64_t* a_tma_bar = full_bar + 16;
  uint64_t* a_epi_bar = full_bar + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_slab_ctas = p.n_slabs * p.parts;
  const bool slab = static_cast<int>(blockIdx.x) < n_slab_ctas;
  const bool keeper = static_cast<int>(blockIdx.x) == n_slab_ctas;   // (pairs: one more CTA fills the keeper's cluster and idles)
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  const int part = slab ? static_cast<int>(blockIdx.x) / p.n_slabs : 0;
  const int m0 = (static_cast<int>(blockIdx.x) % p.n_slabs) * 128;
  const int grp = slab ? m0 / p.rows_per_group : 0;
  const unsigned int slabs_per_group = static_cast<unsigned int>(p.rows_per_group / 128);
  const bool group_leader = slab && part == 0 && (m0 % p.rows_per_group) == 0;
  const uint32_t smem_u = ptx::smem_u32(smem);
  long long* dbg = p.dbg != nullptr ? p.dbg + 32ll * blockIdx.x : nullptr;
  auto stamp = [&](int slot) {
    if (dbg != nullptr) dbg[slot] = static_cast<long long>(gtimer());
  };

  if (threadIdx.x == 0) {
    stamp(0);
    if ((smem_u & 1023u) != 0u) {
      if (p.err != nullptr) atomicExch(p.err, 0xA11Au);
      asm volatile("trap;");
    }
    for (int i = 0; i < 5; ++i) ptx::prefetch_tmap(&p.tm[i]);
    for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.tmo[i]);
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], kPair ? 2 * kEpiWarps : kCEpi);   // pairs: one arrival per epilogue warp of both CTAs
    }
    ptx::mbar_init(a_tma_bar, 1);
    ptx::mbar_init(a_epi_bar, kPair ? 2 * kEpiWarps : kCEpi);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (kPair) {
      ptx::tmem_alloc_pair(tmem_slot, 512);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (kPair) ptx::cluster_sync();   // the partner's barriers are initialised before anything arrives on them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above touched shared memory / TMEM / the kernel parameters only.  Under a programmatic launch the previous
  // kernel of the stream may still be running: nothing of global memory is read or written before it has completed.
  ptx::griddep_wait();

  // All per-slot / per-buffer phase bookkeeping below lives in BIT MASKS, never in indexed arrays: with the whole
  // shared-memory carve-out taken there is almost no L1, and a local-memory array costs an L2 round trip per access.
  if (warp == 0) {
    // =================================================================== TMA producer
    if (lane == 0 && slab) {
      // pairs: both CTAs load (their own A rows, their half of B); the bytes of both are counted on the LEADER's barrier
      const uint32_t full_l = kPair ? ptx::mapa(ptx::smem_u32(full_bar), 0) : 0u;   // + 8 * slot
      auto load2d = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, uint32_t bar_l, int c0, int c1) {
        if constexpr (kPair) ptx::tma_load_2d_pair(dst, tm, bar_l, c0, c1);
        else ptx::tma_load_2d(dst, tm, bar, c0, c1);
      };
      if (p.init_tm >= 0) {
        if (rank == 0) ptx::mbar_expect_tx(a_tma_bar, static_cast<uint32_t>(p.init_panels * kPanel) * (kPair ? 2u : 1u));
        const uint32_t a_tma_l = kPair ? ptx::mapa(ptx::smem_u32(a_tma_bar), 0) : 0u;
        for (int pn = 0; pn < p.init_panels; ++pn) load2d(smem + pn * kPanel, &p.tm[p.init_tm], a_tma_bar, a_tma_l, pn * 64, m0);
      }
      uint32_t fill_par = 0, fill_any = 0;   // bit s: parity of the number of fills of slot s / slot ever filled
      int cur_cfg = -1, s = 0;
      for (int ip = 0; ip < p.n_pass; ++ip) {
        const CPass& ps = p.pass[ip];
        const CRing rg = p.ring[ps.cfg];
        if (ps.cfg != cur_cfg) {
          // the two ring geometries overlap in shared memory: everything in flight must have been consumed
          for (int t = 0; t < 4; ++t)
            if ((fill_any >> t) & 1u) mbar_wait_b(&empty_bar[t], ((fill_par >> t) & 1u) ^ 1u, p.err, 0x100u + t);
          cur_cfg = ps.cfg;
          s = 0;
        }
        // everything the k loop needs lives in registers: a parameter read behind an asm statement is an indexed constant load
        const int k_panels = ps.k_panels, a_stream = ps.a_stream, b_mn = ps.b_mn, nch = ps.n_chunks;
        const int n0_0 = (ip == p.split_pass ? p.part_n0[part] : ps.c_n0[0]) + static_cast<int>(rank) * ps.c_half[0];
        const int n0_1 = ps.c_n0[1] + static_cast<int>(rank) * ps.c_half[1];
        const int boff_0 = ps.c_boff[0], boff_1 = ps.c_boff[1];
        int boxes_0 = ps.c_boxes[0];
        const int boxes_1 = ps.c_boxes[1];
        const bool split_ps = ip == p.split_pass;
        if (split_ps) boxes_0 = (p.part_n[part] + 63) >> 6;
        const CUtensorMap* tm_a = &p.tm[a_stream >= 0 ? a_stream : 0];
        const CUtensorMap* tm_b = &p.tm[ps.b_tm];
        const int ring_base = rg.base, stage_bytes = rg.stage_bytes, a_bytes = rg.a_bytes;
        const uint32_t bytes = ((a_stream >= 0 ? static_cast<uint32_t>(kPanel) : 0u) +
                                static_cast<uint32_t>(split_ps && b_mn ? boxes_0 * 8192 : ps.c_bytes[0]) +
                                (nch > 1 ? static_cast<uint32_t>(ps.c_bytes[1]) : 0u)) * (kPair ? 2u : 1u);
        for (int kp = 0; kp < k_panels; ++kp) {
          mbar_wait_b(&empty_bar[s], ((fill_par >> s) & 1u) ^ 1u, p.err, 0x110u + s);
          if (rank == 0) ptx::mbar_expect_tx(&full_bar[s], bytes);
          uint64_t* fb = &full_bar[s];
          const uint32_t fl = full_l + 8u * s;
          uint8_t* stage = smem + ring_base + s * stage_bytes;
          if (a_stream >= 0) load2d(stage, tm_a, fb, fl, kp * 64, m0);
          uint8_t* bslot = stage + a_bytes;
          if (!b_mn) {
            load2d(bslot + boff_0, tm_b, fb, fl, kp * 64, n0_0);
            if (nch > 1) load2d(bslot + boff_1, tm_b, fb, fl, kp * 64, n0_1);
          } else {
            for (int jb = 0; jb < boxes_0; ++jb) load2d(bslot + boff_0 + jb * 8192, tm_b, fb, fl, n0_0 + jb * 64, kp * 64);
            if (nch > 1)
              for (int jb = 0; jb < boxes_1; ++jb) load2d(bslot + boff_1 + jb * 8192, tm_b, fb, fl, n0_1 + jb * 64, kp * 64);
          }
          fill_par ^= 1u << s;
          fill_any |= 1u << s;
          if (++s == rg.stages) s = 0;
        }
        if (ip < 8) stamp(1 + ip);
      }
    }
    // every load of this CTA is issued (keeper / idle CTA: nothing to issue): the next kernel of the stream may be scheduled
    // onto SMs as they drain (it waits in its own griddep_wait for this grid to complete)
    if (lane == 0) ptx::griddep_launch();
  } else if (warp == 1) {
    // =================================================================== MMA issuer
    if (lane == 0 && slab && rank == 0) {
      uint32_t use_par = 0, acc_par = 0;   // bit s: parity of uses of ring slot s; bit b: parity of uses of accumulator b
      uint32_t n_tma_waits = 0, n_epi_waits = 0;
      int cur_cfg = -1, s = 0;
      auto commit = [&](uint64_t* bar) {
        if constexpr (kPair) ptx::umma_commit_pair(bar);   // arrives in both CTAs
        else ptx::umma_commit(bar);
      };
      for (int ip = 0; ip < p.n_pass; ++ip) {
        const CPass& ps = p.pass[ip];
        const CRing rg = p.ring[ps.cfg];
        if (ps.cfg != cur_cfg) {
          cur_cfg = ps.cfg;
          s = 0;
        }
        if (ps.a_wait == 1) {
          mbar_wait_b<kPair>(a_tma_bar, n_tma_waits & 1, p.err, 0x200u);
          ++n_tma_waits;
        } else if (ps.a_wait == 2) {
          mbar_wait_b<kPair>(a_epi_bar, n_epi_waits & 1, p.err, 0x201u);
          ++n_epi_waits;
        }
        const int nc = ps.n_chunks;
        const int buf0 = ps.c_buf[0], buf1 = ps.c_buf[nc - 1];
        mbar_wait_b<kPair>(&acc_empty[buf0], ((acc_par >> buf0) & 1u) ^ 1u, p.err, 0x210u + buf0);
        if (nc > 1) mbar_wait_b<kPair>(&acc_empty[buf1], ((acc_par >> buf1) & 1u) ^ 1u, p.err, 0x210u + buf1);
        ptx::tc_fence_after();
        const uint32_t idesc0 = ptx::make_idesc(1, 0, ps.b_mn, kPair ? 256 : 128, ip == p.split_pass ? p.part_n[part] : ps.c_mma_n[0]);
        const uint32_t idesc1 = ptx::make_idesc(1, 0, ps.b_mn, kPair ? 256 : 128, ps.c_mma_n[nc - 1]);
        const uint32_t boff0 = ps.c_boff[0], boff1 = ps.c_boff[nc - 1];
        const uint32_t tm0 = tmem_base + ps.c_tmem[0], tm1 = tmem_base + ps.c_tmem[nc - 1];
        const int b_mn = ps.b_mn, k_panels = ps.k_panels, last_ksteps = ps.last_ksteps, a_stream = ps.a_stream;
        const int ring_base = rg.base, stage_bytes = rg.stage_bytes, a_bytes = rg.a_bytes;
        const uint32_t b_lbo = b_mn ? ((8192u >> 4) << 16) : (1u << 16);   // MN-major: 64-column boxes 8 KB apart
        const uint32_t b_kstep = b_mn ? (2048u >> 4) : 2u;                 // 16 k-rows of an MN-major box / 32 bytes of a K-major row
        for (int kp = 0; kp < k_panels; ++kp) {
          mbar_wait_b<kPair>(&full_bar[s], (use_par >> s) & 1u, p.err, 0x220u + s);
          use_par ^= 1u << s;
          ptx::tc_fence_after();
          const uint32_t stage = smem_u + ring_base + s * stage_bytes;
          const uint32_t a_base = a_stream >= 0 ? stage : smem_u + kp * kPanel;
          const uint32_t bslot = stage + a_bytes;
          const int nks = (kp == k_panels - 1) ? last_ksteps : 4;
          // descriptors are built once per stage; a k-step only advances their 16-byte-unit address fields
          const uint32_t a_lo = ((a_base & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t b0_lo = (((bslot + boff0) & 0x3FFFFu) >> 4) | b_lbo;
          const uint32_t b1_lo = (((bslot + boff1) & 0x3FFFFu) >> 4) | b_lbo;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < nks) {
              const uint32_t acc = (kp | ks) != 0 ? 1u : 0u;
              if constexpr (kPair) {
                umma_bf16_lohi_pair(tm0, a_lo + 2 * ks, b0_lo + b_kstep * ks, kDescHi, idesc0, acc);
                if (nc > 1) umma_bf16_lohi_pair(tm1, a_lo + 2 * ks, b1_lo + b_kstep * ks, kDescHi, idesc1, acc);
              } else {
                umma_bf16_lohi(tm0, a_lo + 2 * ks, b0_lo + b_kstep * ks, kDescHi, idesc0, acc);
                if (nc > 1) umma_bf16_lohi(tm1, a_lo + 2 * ks, b1_lo + b_kstep * ks, kDescHi, idesc1, acc);
              }
            }
          }
          commit(&empty_bar[s]);
          if (++s == rg.stages) s = 0;
        }
        commit(&acc_full[buf0]);
        acc_par ^= 1u << buf0;
        if (nc > 1) {
          commit(&acc_full[buf1]);
          acc_par ^= 1u << buf1;
        }
        if (ip < 8) stamp(9 + ip);
      }
    }
  } else if (slab) {
    // =================================================================== epilogue (16 warps)
    const int et = threadIdx.x - 64;
    const int q = warp & 3;               // TMEM lane quarter this warp may read
    const int jw = (warp - 2) >> 2;       // the four warps of a quarter take every fourth 32-column unit
    const long long wrow0 = static_cast<long long>(m0) + q * 32;   // first global row of this warp
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t tile = smem_u + kTileOff + (warp - 2) * kTileBytes;
    const int rsel = lane >> 3, cq = lane & 7;   // column phase: rows 4*rsel + i of a 16-row half, columns 4*cq .. 4*cq+3 of the unit
    uint32_t epi_par = 0;                 // bit b: parity of the accumulator-full phases consumed so far
    const float inv_cnt = 1.f / static_cast<float>(p.rows_per_group);
    // "this warp is done with ..." for the MMA thread: per thread on the CTA's own barrier, or - pairs - one arrival per warp
    // on the leader's barrier (the MMA thread there waits for the epilogues of both CTAs)
    auto epi_arrive = [&](uint64_t* bar) {
      if constexpr (kPair) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(bar), 0));
      } else {
        ptx::mbar_arrive(bar);
      }
    };

    // this lane's accumulator row of a 32-column unit, raw from TMEM (the upper 16 columns zero when `wide` is false)
    // This lane's accumulator row of a 32-column unit, raw from TMEM.  Columns beyond the layer hold whatever TMEM holds:
    // no live lane consumes them (the allocation is 512 columns, the read never leaves it).
    auto load_unit = [&](uint32_t taddr, uint32_t (&v)[32]) {
      uint32_t lo[16], hi[16];
      ptx::tmem_ld16(taddr, lo);
      ptx::tmem_ld16(taddr + 16, hi);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        v[k] = lo[k];
        v[16 + k] = hi[k];
      }
    };
    // Addresses that depend on the lane only are computed once: the transposed reads of the tile (rows 4 rsel + i, chunk cq)
    // and the per-row part of an arena address (row q*32 + 4 rsel + i of a panel; a 16-row half adds 2048 bytes).
    const uint32_t trd0 = tile_addr(tile, 4 * rsel + 0, cq), trd1 = tile_addr(tile, 4 * rsel + 1, cq);
    const uint32_t trd2 = tile_addr(tile, 4 * rsel + 2, cq), trd3 = tile_addr(tile, 4 * rsel + 3, cq);
    const uint32_t arow = smem_u + (q * 32 + 4 * rsel) * 128;   // rows (q*32 + 4 rsel + i): + i * 128; row & 7 = 4 (rsel & 1) + i
    // arena address of (row q*32 + half*16 + 4 rsel + i, columns col0..col0+3) = ua[i] + half * 2048
    auto unit_arena = [&](int col0, uint32_t (&ua)[4]) {
      const uint32_t base = arow + (col0 >> 6) * kPanel + ((col0 & 4) << 1);
      const uint32_t cx = static_cast<uint32_t>((((col0 & 63) >> 3) ^ (4 * (rsel & 1))) << 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) ua[i] = base + i * 128 + (cx ^ (i << 4));
    };
    // rows [16*half, 16*half+16) of the warp go into the tile
    auto deposit = [&](int half, const uint32_t (&v)[32]) {
      if ((lane >> 4) == half) {
        const int tr = lane & 15;
#pragma unroll
        for (int c = 0; c < 8; ++c) ptx::sts128(tile_addr(tile, tr, c), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
      __syncwarp();
    };

    auto run_layer = [&](auto kind_c, auto il_c) {
      constexpr int kKind = decltype(kind_c)::value;
      constexpr int il = decltype(il_c)::value;
      const CLayer& L = p.layer[il];
      const int N = L.N;
      const int Npad = (N + 15) & ~15;
      const int n_units = (Npad + 31) >> 5;
      // column split (see CParams): this CTA's units [u_lo, u_hi) / columns [n_lo, n_hi) of the layer; TMEM column = column - n_lo.
      // `shadow`: a part > 0 in a layer before the split one - it computes, but publishes nothing.
      const bool split = il == p.split_layer;
      const bool shadow = part > 0 && !split;
      const int n_lo = split ? p.part_n0[part] : 0;
      const int n_hi = split ? min(N, n_lo + p.part_n[part]) : N;
      const int u_lo = n_lo >> 5;
      const int u_hi = split ? (((n_hi + 15) & ~15) + 31) >> 5 : n_units;
      const unsigned int arrivals = slabs_per_group * (split ? static_cast<unsigned int>(p.parts) : 1u);
      if (et == 0) stamp(17 + il * 4);
      auto wait_all_chunks = [&]() {
        for (int ci = 0; ci < L.n_chunks; ++ci) mbar_wait_epi(&acc_full[ci], (epi_par >> ci) & 1u, p.err, 0x300u + il * 16 + ci);
        ptx::tc_fence_after();
      };
      auto release_all_chunks = [&]() {
        ptx::tc_fence_before();
        for (int ci = 0; ci < L.n_chunks; ++ci) {
          epi_arrive(&acc_empty[ci]);
          epi_par ^= 1u << ci;
        }
      };
      // the arena's panels of this layer -> global memory through an output tensor map (one thread; asynchronous: the
      // epilogue warps issue no global stores for the activations, and the copy runs under whatever comes next)
      auto store_arena = [&](const CUtensorMap* om) {
        if (shadow) return;
        const int pn_hi = (n_hi + 63) >> 6;
        for (int pn = n_lo >> 6; pn < pn_hi; ++pn) ptx::tma_store_2d(om, smem + pn * kPanel, pn * 64, m0);
        ptx::bulk_commit_group();
      };
      // publish the slab's column sums: red[quarter][stat][column] -> one global atomic per column and statistic
      auto publish_stats = [&]() {
        bar_epi();
        if (shadow) return;
        for (int c = n_lo + et; c < n_hi; c += kCEpi) {
          const float a0 = (red[c] + red[kRedStride + c]) + (red[2 * kRedStride + c] + red[3 * kRedStride + c]);
          const float a1 = (red[400 + c] + red[kRedStride + 400 + c]) + (red[2 * kRedStride + 400 + c] + red[3 * kRedStride + 400 + c]);
          atomicAdd(L.stat0 + grp * N + c, a0);
          atomicAdd(L.stat1 + grp * N + c, a1);
        }
      };

      if constexpr (kKind == CE_FWD_BN) {
        float* s_bias = tab;
        float* s_ca = tab + 400;
        float* s_cb = tab + 800;
        for (int c = et; c < 400; c += kCEpi) s_bias[c] = c < N ? L.bias[c] : 0.f;
        if (et == 0) ptx::bulk_wait_read_all();   // the previous layer's output copy has left the arena
        bar_epi();
        wait_all_chunks();
        // ---- pass 1: slab statistics; the bf16 pre-activations are parked in the arena (the layer's own A operand is spent)
        for (int u = u_lo + jw; u < u_hi; u += 4) {
          uint32_t v[32];
          load_unit(t_row + 32 * (u - u_lo), v);
          const int col0 = 32 * u + 4 * cq;
          const bool live = col0 < N, padded = col0 < Npad;
          const float4 bs = padded ? tab4(s_bias + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
          const f2 bs01 = pk2(bs.x, bs.y), bs23 = pk2(bs.z, bs.w);
          f2 s0a = pk2(0.f, 0.f), s0b = s0a, s1a = s0a, s1b = s0a;   // packed pairs: columns (0, 1) and (2, 3)
          uint32_t ua[4];
          unit_arena(col0, ua);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            deposit(half, v);
            if (padded) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 a = lds_f4(i == 0 ? trd0 : (i == 1 ? trd1 : (i == 2 ? trd2 : trd3)));
                const f2 x01 = add2(pk2(a.x, a.y), bs01), x23 = add2(pk2(a.z, a.w), bs23);
                s0a = add2(s0a, x01); s0b = add2(s0b, x23);
                s1a = fma2(x01, x01, s1a); s1b = fma2(x23, x23, s1b);
                float x0, x1, x2, x3;
                upk2(x01, x0, x1); upk2(x23, x2, x3);
                sts64(ua[i] + half * 2048, pack_bf16(x0, x1), pack_bf16(x2, x3));
              }
            }
            __syncwarp();
          }
          float s0[4], s1[4];
          upk2(s0a, s0[0], s0[1]); upk2(s0b, s0[2], s0[3]);
          upk2(s1a, s1[0], s1[1]); upk2(s1b, s1[2], s1[3]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            s0[k] += __shfl_xor_sync(0xffffffffu, s0[k], 8);
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 8);
            s0[k] += __shfl_xor_sync(0xffffffffu, s0[k], 16);
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 16);
          }
          if (live && rsel == 0) {
            *reinterpret_cast<float4*>(red + q * kRedStride + col0) = make_float4(s0[0], s0[1], s0[2], s0[3]);
            *reinterpret_cast<float4*>(red + q * kRedStride + 400 + col0) = make_float4(s1[0], s1[1], s1[2], s1[3]);
          }
        }
        release_all_chunks();
        ptx::fence_proxy_async_smem();             // the parked pre-activations are read by the TMA store below
        publish_stats();                           // (starts with a barrier of the epilogue warps)
        if (et == 0) store_arena(&p.tmo[2 * il]);  // pre-BatchNorm activations -> global (the backward reads them)
        // column split: the other parts of this slab fetch these columns from there after the barrier - the copy has to be
        // complete (not just read out of the arena) before this CTA arrives
        const bool exchange = split && p.parts > 1;
        if (exchange && et == 0) ptx::bulk_wait_all();
        if (et == 0) stamp(18 + il * 4);
        group_barrier(L.counter + grp, arrivals, p.err, et, !shadow);
        if (et == 0) stamp(19 + il * 4);
        if (exchange && et == 0) {
          // the partners' pre-activations: their panels of the same rows, global (L2) -> this CTA's arena; every part then
          // normalises the whole slab itself (pass 2 below), so the next layer finds its full A operand on chip
          ptx::fence_proxy_async_all();
          const int pn_lo = n_lo >> 6, pn_hi = (n_hi + 63) >> 6, pn_all = (N + 63) >> 6;
          ptx::mbar_expect_tx(a_tma_bar, static_cast<uint32_t>((pn_all - (pn_hi - pn_lo)) * kPanel));
          for (int pn = 0; pn < pn_all; ++pn)
            if (pn < pn_lo || pn >= pn_hi) ptx::tma_load_2d(smem + pn * kPanel, &p.tmo[2 * il], a_tma_bar, pn * 64, m0);
        }
        for (int c = et; c < 400; c += kCEpi) {
          float a_ = 0.f, b_ = 0.f;
          if (c < N) {
            const float mean = __ldcg(L.stat0 + grp * N + c) * inv_cnt;
            const float var = fmaxf(__ldcg(L.stat1 + grp * N + c) * inv_cnt - mean * mean, 0.f);
            const float rstd = rsqrtf(var + p.eps);
            a_ = L.gamma[c] * rstd;
            b_ = fmaf(-mean, a_, L.beta[c]);
            if (group_leader) {
              L.save_mean[grp * N + c] = mean;
              L.save_rstd[grp * N + c] = rstd;
            }
          }
          s_ca[c] = a_;
          s_cb[c] = b_;
        }
        if (et == 0) ptx::bulk_wait_read_all();    // pass 2 overwrites the arena in place
        if (exchange) mbar_wait_epi(a_tma_bar, 0, p.err, 0x3F0u + il);   // (one split layer per kernel: phase 0)
        bar_epi();
        // ---- pass 2: BatchNorm + ReLU in place in the arena -> next layer's A operand; both copies the backward and the
        //      weight gradients need go to global memory from here (after the barrier: nothing for its fence to wait on)
        for (int u = jw; u < n_units; u += 4) {
          const int col0 = 32 * u + 4 * cq;
          if (!(col0 < Npad)) continue;
          const float4 ca = tab4(s_ca + col0);
          const float4 cb = tab4(s_cb + col0);
          const f2 ca01 = pk2(ca.x, ca.y), ca23 = pk2(ca.z, ca.w), cb01 = pk2(cb.x, cb.y), cb23 = pk2(cb.z, cb.w);
          uint32_t ua[4];
          unit_arena(col0, ua);
#pragma unroll
          for (int hi = 0; hi < 8; ++hi) {
            const uint32_t addr = ua[hi & 3] + (hi >> 2) * 2048;
            const uint2 xw = lds64(addr);
            float y0, y1, y2, y3;
            upk2(fma2(ca01, pk2(bf_lo(xw.x), bf_hi(xw.x)), cb01), y0, y1);
            upk2(fma2(ca23, pk2(bf_lo(xw.y), bf_hi(xw.y)), cb23), y2, y3);
            sts64(addr, pack_bf16(fmaxf(y0, 0.f), fmaxf(y1, 0.f)), pack_bf16(fmaxf(y2, 0.f), fmaxf(y3, 0.f)));
          }
        }
        ptx::fence_proxy_async_smem();
        if (L.write_arena) epi_arrive(a_epi_bar);
        bar_epi();
        if (et == 0) store_arena(&p.tmo[2 * il + 1]);   // post-BatchNorm activations -> global (the next weight gradient's operand)
      } else if constexpr (kKind == CE_DGRAD_BN) {
        float* s_a = tab;
        float* s_b = tab + 400;
        float* s_rs = tab + 800;    // pass 1: rstd            pass 2: k1 = a * mean(dyhat * xhat) * rstd
        float* s_mr = tab + 1200;   // pass 1: -mean * rstd    pass 2: k2 = k1 * mean - a * mean(dyhat)
        for (int c = et; c < 400; c += kCEpi) {
          float a_ = 0.f, b_ = 0.f, rs = 0.f, mr = 0.f;
          if (c < N) {
            const float mean = L.save_mean[grp * N + c];
            rs = L.save_rstd[grp * N + c];
            a_ = L.gamma[c] * rs;
            b_ = fmaf(-mean, a_, L.beta[c]);   // the forward's own expression: bit-identical ReLU mask
            mr = -mean * rs;
          }
          s_a[c] = a_; s_b[c] = b_; s_rs[c] = rs; s_mr[c] = mr;
        }
        if (et == 0) ptx::bulk_wait_read_all();   // the previous layer's output copy has left the arena
        bar_epi();
        // this lane's pre-activations of a unit: [half * 4 + i] -> 4 bf16 (pass 2 requests them one unit ahead: an L2 round
        // trip is ~1 us; pass 1 has no registers left for that)
        auto load_hpre = [&](int u, uint2 (&h)[8]) {
          const int c0 = 32 * u + 4 * cq;
          const bool lv = u < u_hi && c0 < N;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            h[k] = lv ? __ldg(reinterpret_cast<const uint2*>(L.hpre + (wrow0 + (k >> 2) * 16 + 4 * rsel + (k & 3)) * N + c0))
                      : make_uint2(0u, 0u);
        };
        wait_all_chunks();
        // ---- pass 1: ReLU mask, slab sums of dyhat and dyhat * xhat; the masked gradient dyhat is parked in the arena (bf16):
        //      pass 2 then needs neither TMEM nor the transposing tile nor the mask again.
        //      sum dyhat * xhat = rstd * sum(dyhat * x) - mean * rstd * sum(dyhat): the loop accumulates sum(dyhat * x)
        for (int u = u_lo + jw; u < u_hi; u += 4) {
          const int col0 = 32 * u + 4 * cq;
          const bool live = col0 < N;
          uint2 hx[8];
          load_hpre(u, hx);   // in flight during the TMEM load and the first deposit
          uint32_t v[32];
          load_unit(t_row + 32 * (u - u_lo), v);
          const float4 ta = live ? tab4(s_a + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 tb = live ? tab4(s_b + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
          float s0[4] = {0.f, 0.f, 0.f, 0.f}, sx[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t ua[4];
          unit_arena(col0, ua);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            deposit(half, v);
            if (live) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 a = lds_f4(i == 0 ? trd0 : (i == 1 ? trd1 : (i == 2 ? trd2 : trd3)));
                const uint2 hw = hx[half * 4 + i];
                const float x0 = bf_lo(hw.x), x1 = bf_hi(hw.x), x2 = bf_lo(hw.y), x3 = bf_hi(hw.y);
                const float d0 = fmaf(ta.x, x0, tb.x) > 0.f ? a.x : 0.f;
                const float d1 = fmaf(ta.y, x1, tb.y) > 0.f ? a.y : 0.f;
                const float d2 = fmaf(ta.z, x2, tb.z) > 0.f ? a.z : 0.f;
                const float d3 = fmaf(ta.w, x3, tb.w) > 0.f ? a.w : 0.f;
                sts64(ua[i] + half * 2048, pack_bf16(d0, d1), pack_bf16(d2, d3));
                s0[0] += d0; s0[1] += d1; s0[2] += d2; s0[3] += d3;
                sx[0] = fmaf(d0, x0, sx[0]); sx[1] = fmaf(d1, x1, sx[1]); sx[2] = fmaf(d2, x2, sx[2]); sx[3] = fmaf(d3, x3, sx[3]);
              }
            }
            __syncwarp();
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            s0[k] += __shfl_xor_sync(0xffffffffu, s0[k], 8);
            sx[k] += __shfl_xor_sync(0xffffffffu, sx[k], 8);
            s0[k] += __shfl_xor_sync(0xffffffffu, s0[k], 16);
            sx[k] += __shfl_xor_sync(0xffffffffu, sx[k], 16);
          }
          if (live && rsel == 0) {
            const float4 trs = tab4(s_rs + col0), tmr = tab4(s_mr + col0);
            *reinterpret_cast<float4*>(red + q * kRedStride + col0) = make_float4(s0[0], s0[1], s0[2], s0[3]);
            *reinterpret_cast<float4*>(red + q * kRedStride + 400 + col0) =
                make_float4(fmaf(trs.x, sx[0], tmr.x * s0[0]), fmaf(trs.y, sx[1], tmr.y * s0[1]), fmaf(trs.z, sx[2], tmr.z * s0[2]),
                            fmaf(trs.w, sx[3], tmr.w * s0[3]));
          }
        }
        release_all_chunks();
        publish_stats();
        uint2 hx2[8];
        load_hpre(u_lo + jw, hx2);
        if (et == 0) stamp(18 + il * 4);
        group_barrier(L.counter + grp, arrivals, p.err, et, !shadow);
        if (et == 0) stamp(19 + il * 4);
        for (int c = et; c < 400; c += kCEpi) {
          if (c < N) {
            const float c0 = __ldcg(L.stat0 + grp * N + c) * inv_cnt;
            const float c1 = __ldcg(L.stat1 + grp * N + c) * inv_cnt;
            const float a_ = s_a[c];
            const float mean = L.save_mean[grp * N + c];
            const float k1 = a_ * c1 * s_rs[c];
            s_rs[c] = k1;
            s_mr[c] = fmaf(k1, mean, -a_ * c0);
          }
        }
        bar_epi();
        // ---- pass 2: dx = gamma * rstd * (dyhat - mean(dyhat) - xhat * mean(dyhat * xhat)) = a * dyhat - k1 * x + k2,
        //      in place in the arena (the next dgrad's A operand) and to global memory (the weight gradient's operand)
        for (int u = u_lo + jw; u < u_hi; u += 4) {
          const int col0 = 32 * u + 4 * cq;
          const bool live = col0 < N, padded = col0 < Npad;
          uint2 hn[8];
          load_hpre(u + 4, hn);
          if (padded) {
            const float4 ta = live ? tab4(s_a + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 k1 = live ? tab4(s_rs + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 k2 = live ? tab4(s_mr + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
            const f2 ta01 = pk2(ta.x, ta.y), ta23 = pk2(ta.z, ta.w), nk01 = pk2(-k1.x, -k1.y), nk23 = pk2(-k1.z, -k1.w);
            const f2 k201 = pk2(k2.x, k2.y), k223 = pk2(k2.z, k2.w);
            uint32_t ua[4];
            unit_arena(col0, ua);
#pragma unroll
            for (int hi = 0; hi < 8; ++hi) {
              const uint32_t addr = ua[hi & 3] + (hi >> 2) * 2048;
              const uint2 dw = live ? lds64(addr) : make_uint2(0u, 0u);
              const uint2 hw = hx2[hi];
              float y0, y1, y2, y3;   // a * dyhat - k1 * x + k2 on packed pairs
              upk2(fma2(ta01, pk2(bf_lo(dw.x), bf_hi(dw.x)), fma2(nk01, pk2(bf_lo(hw.x), bf_hi(hw.x)), k201)), y0, y1);
              upk2(fma2(ta23, pk2(bf_lo(dw.y), bf_hi(dw.y)), fma2(nk23, pk2(bf_lo(hw.y), bf_hi(hw.y)), k223)), y2, y3);
              sts64(addr, pack_bf16(y0, y1), pack_bf16(y2, y3));
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) hx2[k] = hn[k];
        }
        ptx::fence_proxy_async_smem();
        if (L.write_arena) epi_arrive(a_epi_bar);
        bar_epi();
        if (et == 0) store_arena(&p.tmo[2 * il + 1]);   // gradient at the pre-BatchNorm output -> global (weight gradient operand)
      } else if constexpr (kKind == CE_FWD_STORE || kKind == CE_DGRAD_STORE) {
        float* s_bias = tab;
        for (int c = et; c < 400; c += kCEpi) s_bias[c] = (L.bias != nullptr && c < N) ? L.bias[c] : 0.f;
        bar_epi();
        wait_all_chunks();
        for (int u = jw; u < n_units; u += 4) {
          uint32_t v[32];
          load_unit(t_row + 32 * u, v);
          const int col0 = 32 * u + 4 * cq;
          const bool live = col0 < N;
          const float4 bs = live ? tab4(s_bias + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            deposit(half, v);
            if (live && !shadow) {
              float* dst = L.out_f32 + (wrow0 + half * 16 + 4 * rsel) * L.ld_out + col0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 a = lds_f4(i == 0 ? trd0 : (i == 1 ? trd1 : (i == 2 ? trd2 : trd3)));
                *reinterpret_cast<float4*>(dst + static_cast<long long>(i) * L.ld_out) = make_float4(a.x + bs.x, a.y + bs.y, a.z + bs.z, a.w + bs.w);
              }
            }
            __syncwarp();
          }
        }
        release_all_chunks();
      } else if constexpr (kKind == CE_BCE) {
        // last decoder Linear + sigmoid + BCE (mnist/model.py:130,135 + mnist/train.py:70): logits never leave the SM.
        // loss = softplus(x) - t x, dlogit = scale * (sigmoid(x) - t); also the bias gradient (column sums of dlogit).
        float* s_bias = tab;
        for (int c = et; c < 800; c += kCEpi) s_bias[c] = c < N ? L.bias[c] : 0.f;
        bar_epi();
        const float scale = grp == 0 ? L.bce_scale[0] : (grp == 1 ? L.bce_scale[1] : L.bce_scale[2]);
        const __nv_bfloat16* tbase = L.target + (static_cast<long long>(m0 % L.target_rows) + q * 32) * N;
        // sums of max(x, 0), of t x (pairs of lanes: packed arithmetic) and of log2 sigmoid(|x|)
        f2 macc = pk2(0.f, 0.f), txacc = pk2(0.f, 0.f);
        float llog = 0.f;
        const f2 one2 = pk2(1.f, 1.f), sc2 = pk2(scale, scale), nsc2 = pk2(-scale, -scale);
        // The accumulator arrives in chunks (TMEM buffer = chunk & 1) while the MMA warp works on the next one; the 32-column
        // units of the whole layer go round-robin over the four warps of a quarter.  Every warp waits for and releases every
        // chunk in order (an arrival on a buffer's "empty" barrier must follow this warp's wait on its "full" phase).
        const int u1 = L.u1, u2 = L.u2, u3 = L.u3, n_chunks = L.n_chunks, buf_cols = L.buf_cols;
        int ci_waited = -1;
        for (int u = jw; u < n_units; u += 4) {
          const int ci = (u >= u1 ? 1 : 0) + (u >= u2 ? 1 : 0) + (u >= u3 ? 1 : 0);
          const int ustart = ci == 0 ? 0 : (ci == 1 ? u1 : (ci == 2 ? u2 : u3));
          const int col0 = 32 * u + 4 * cq;
          const bool live = col0 < N;
          uint2 tx[8];   // [half][i] -> 4 bf16 targets, in flight during the barrier wait and the TMEM load
#pragma unroll
          for (int k = 0; k < 8; ++k)
            tx[k] = live ? __ldg(reinterpret_cast<const uint2*>(tbase + ((k >> 2) * 16 + 4 * rsel + (k & 3)) * static_cast<long long>(N) + col0))
                         : make_uint2(0u, 0u);
          while (ci_waited < ci) {
            if (ci_waited >= 0) {   // done with chunk ci_waited
              ptx::tc_fence_before();
              epi_arrive(&acc_empty[ci_waited & 1]);
            }
            ++ci_waited;
            const int b = ci_waited & 1;
            mbar_wait_epi(&acc_full[b], (epi_par >> b) & 1u, p.err, 0x300u + il * 16 + ci_waited);
            epi_par ^= 1u << b;
            ptx::tc_fence_after();
          }
          uint32_t v[32];
          load_unit(t_row + (ci & 1) * buf_cols + 32 * (u - ustart), v);
          const float4 bs = live ? tab4(s_bias + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
          const f2 bs01 = pk2(bs.x, bs.y), bs23 = pk2(bs.z, bs.w);
          f2 sd01 = pk2(0.f, 0.f), sd23 = pk2(0.f, 0.f);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            deposit(half, v);
            if (live) {
              const long long orow = (wrow0 + half * 16 + 4 * rsel) * N + col0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 a = lds_f4(i == 0 ? trd0 : (i == 1 ? trd1 : (i == 2 ? trd2 : trd3)));
                const uint2 tw = tx[half * 4 + i];
                uint32_t dw[2], pw[2];
                f2 prod2;
                // softplus(x) - t x = max(x, 0) - t x - ln sigmoid(|x|); two elements per packed operation
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                  const uint32_t tword = h2 == 0 ? tw.x : tw.y;
                  const f2 t2 = pk2(bf_lo(tword), bf_hi(tword));
                  const f2 x2 = add2(h2 == 0 ? pk2(a.x, a.y) : pk2(a.z, a.w), h2 == 0 ? bs01 : bs23);
                  float x0, x1;
                  upk2(x2, x0, x1);
                  const float e0 = ptx::ex2_approx(-1.4426950408889634f * fabsf(x0));   // exp(-|x|) in (0, 1]
                  const float e1 = ptx::ex2_approx(-1.4426950408889634f * fabsf(x1));
                  const f2 ex2v = pk2(e0, e1);
                  float den0, den1;
                  upk2(add2(ex2v, one2), den0, den1);
                  const float i0 = ptx::rcp_approx(den0), i1 = ptx::rcp_approx(den1);    // sigmoid(|x|) in [0.5, 1)
                  const f2 inv2 = pk2(i0, i1);
                  float q0, q1;
                  upk2(mul2(ex2v, inv2), q0, q1);
                  const f2 p2 = pk2(x0 >= 0.f ? i0 : q0, x1 >= 0.f ? i1 : q1);
                  const f2 d2 = fma2(t2, nsc2, mul2(p2, sc2));                           // scale * (p - t)
                  if (h2 == 0) sd01 = add2(sd01, d2);
                  else sd23 = add2(sd23, d2);
                  prod2 = h2 == 0 ? inv2 : mul2(prod2, inv2);                            // >= 1/16 per lane: one log per four elements
                  macc = add2(macc, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
                  txacc = fma2(t2, x2, txacc);
                  float d0, d1, p0, p1;
                  upk2(d2, d0, d1);
                  upk2(p2, p0, p1);
                  dw[h2] = pack_bf16(d0, d1);
                  pw[h2] = pack_bf16(p0, p1);
                }
                float pl, ph;
                upk2(prod2, pl, ph);
                llog += ptx::lg2_approx(pl * ph);
                *reinterpret_cast<uint2*>(L.dlog + orow + static_cast<long long>(i) * N) = make_uint2(dw[0], dw[1]);
                if (L.probs != nullptr) *reinterpret_cast<uint2*>(L.probs + orow + static_cast<long long>(i) * N) = make_uint2(pw[0], pw[1]);
              }
            }
            __syncwarp();
          }
          if (L.dbias != nullptr) {
            float sd[4];
            upk2(sd01, sd[0], sd[1]);
            upk2(sd23, sd[2], sd[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              sd[e] += __shfl_xor_sync(0xffffffffu, sd[e], 8);
              sd[e] += __shfl_xor_sync(0xffffffffu, sd[e], 16);
            }
            if (live && rsel == 0) *reinterpret_cast<float4*>(red + q * kRedStride + col0) = make_float4(sd[0], sd[1], sd[2], sd[3]);
          }
        }
        // the remaining chunks (this warp has no more units in them)
        while (ci_waited < n_chunks - 1) {
          if (ci_waited >= 0) {
            ptx::tc_fence_before();
            epi_arrive(&acc_empty[ci_waited & 1]);
          }
          ++ci_waited;
          const int b = ci_waited & 1;
          mbar_wait_epi(&acc_full[b], (epi_par >> b) & 1u, p.err, 0x300u + il * 16 + ci_waited);
          epi_par ^= 1u << b;
        }
        ptx::tc_fence_before();
        epi_arrive(&acc_empty[ci_waited & 1]);
        float m_lo, m_hi, tx_lo, tx_hi;
        upk2(macc, m_lo, m_hi);
        upk2(txacc, tx_lo, tx_hi);
        float lsum = fmaf(-0.6931471805599453f, llog, (m_lo + m_hi) - (tx_lo + tx_hi));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        if (lane == 0 && L.loss != nullptr) atomicAdd(L.loss + grp, scale * lsum);
        if (L.dbias != nullptr) {
          bar_epi();
          for (int c = et; c < N; c += kCEpi)
            atomicAdd(L.dbias + c, (red[c] + red[kRedStride + c]) + (red[2 * kRedStride + c] + red[3 * kRedStride + c]));
        }
      }
      bar_epi();   // the tables are reused by the next layer
      if (et == 0) stamp(20 + il * 4);
    };
    run_layer(std::integral_constant<int, kKind0>{}, std::integral_constant<int, 0>{});
    run_layer(std::integral_constant<int, kKind1>{}, std::integral_constant<int, 1>{});
    if constexpr (kKind2 >= 0) run_layer(std::integral_constant<int, kKind2>{}, std::integral_constant<int, 2>{});
    if (et == 0) ptx::bulk_wait_all();   // the arena's output copies are complete before the CTA gives up its shared memory
  } else if (keeper) {
    // =================================================================== keeper CTA: what needs every group's statistics,
    // in group order (one reference forward pass per group), while the slab CTAs carry on
    const int et = threadIdx.x - 64;
    const int groups = p.n_slabs / static_cast<int>(slabs_per_group);
    for (int il = 0; il < p.n_layers; ++il) {
      const CLayer& L = p.layer[il];
      if (L.kind != CE_FWD_BN && L.kind != CE_DGRAD_BN) continue;
      if (et == 0)
        for (int g = 0; g < groups; ++g)
          group_wait(L.counter + g, slabs_per_group * (il == p.split_layer ? static_cast<unsigned int>(p.parts) : 1u), p.err);
      bar_epi();
      const int N = L.N;
      const float cnt = static_cast<float>(p.rows_per_group);
      for (int c = et; c < N; c += kCEpi) {
        if (L.kind == CE_FWD_BN) {
          if (L.running_mean == nullptr) continue;
          float rm = L.running_mean[c], rv = L.running_var[c];
          for (int g = 0; g < groups; ++g) {
            const float mean = __ldcg(L.stat0 + g * N + c) / cnt;
            const float var = fmaxf(__ldcg(L.stat1 + g * N + c) / cnt - mean * mean, 0.f);
            const float unb = cnt > 1.f ? var * (cnt / (cnt - 1.f)) : var;
            for (int u = 0; u < L.bn_updates; ++u) {
              rm = (1.f - p.momentum) * rm + p.momentum * mean;
              rv = (1.f - p.momentum) * rv + p.momentum * unb;
            }
          }
          L.running_mean[c] = rm;
          L.running_var[c] = rv;
        } else {
          float dg = 0.f, db = 0.f;
          for (int g = 0; g < groups; ++g) {
            db += __ldcg(L.stat0 + g * N + c);
            dg += __ldcg(L.stat1 + g * N + c);
          }
          if (L.dgamma != nullptr) {
            L.dgamma[c] += dg;
            L.dbeta[c] += db;
          }
        }
      }
    }
  }
  __syncwarp();
  ptx::tc_fence_before();
  if constexpr (kPair) {
    ptx::cluster_sync();   // the partner's MMAs have read this CTA's shared memory, its epilogue has arrived on our barriers
    if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, 512);
  } else {
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) stamp(31);
}

// ---------------------------------------------------------------- host side
int chain_sms();
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn chain_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// bf16 matrix [rows, cols] (cols contiguous, leading dimension ld), box = box_cols x box_rows, SWIZZLE_128B, zero OOB fill
int chain_tmap(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
  EncodeTiledFn enc = chain_encode_fn();
  MVAE_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVAE_REQUIRE(r == CUDA_SUCCESS, "chain: cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r, rows,
               cols, ld, box_cols, box_rows);
  return 0;
}

// `boxes` / `bytes`: what ONE CTA loads per stage for the chunk (pairs: its half of the chunk's output columns)
void set_chunk(CPass& ps, int c, int n0, int mma_n, int boxes, int bytes, int boff, int tmem, int buf, bool pair = false) {
  ps.c_n0[c] = n0; ps.c_mma_n[c] = mma_n; ps.c_boxes[c] = boxes; ps.c_bytes[c] = bytes; ps.c_boff[c] = boff;
  ps.c_tmem[c] = tmem; ps.c_buf[c] = buf; ps.c_half[c] = pair ? mma_n / 2 : 0;
}
int ksteps_of(int K) { return ((K - 1) % 64) / 16 + 1; }   // 16-wide k-steps in the last 64-wide panel
int panels_of(int K) { return (K + 63) / 64; }

constexpr int kFwdChunk = 224;   // K-major weight chunk of a resident-A layer: 224 rows x 128 B = 28672 B per stage
constexpr int kBwdChunk = 256;   // MN-major: four 64 x 64 boxes = 32768 B per stage

// Passes of one resident-A layer: one accumulator chunk per pass (wide chunks: a tcgen05.commit costs the issuing thread
// ~0.25 us, so a stage has to carry a few hundred cycles of MMA work).  BatchNorm / store layers: TMEM column = layer
// column, barrier pair = chunk index.  BCE layer (784 columns): chunk widths are multiples of 32 (whole epilogue units),
// two TMEM buffers of 224 columns alternate.  Also fills the layer's chunk description.  Returns the next free pass index.
int add_resident_layer(CParams& p, int ip, CLayer& L, int N, int K, int b_tm, bool b_mn, int first_a_wait, bool bce, bool pair) {
  const int Npad = (N + 15) & ~15;
  int widths[4] = {0, 0, 0, 0};
  int nch = 0;
  if (bce) {
    int left = Npad;
    widths[nch++] = std::min(left, 224);
    left -= widths[0];
    while (left > 0 && nch < 4) {
      widths[nch] = std::min(left, 192);
      left -= widths[nch++];
    }
  } else {
    const int cw = b_mn ? kBwdChunk : kFwdChunk;
    int left = Npad;
    while (left > 0 && nch < 4) {
      widths[nch] = std::min(left, cw);
      left -= widths[nch++];
    }
  }
  L.n_chunks = nch;
  L.buf_cols = 224;
  L.u1 = L.u2 = L.u3 = 1 << 20;
  int n0 = 0;
  for (int c = 0; c < nch; ++c) {
    CPass& ps = p.pass[ip++];
    const int w = widths[c];
    ps.cfg = 1; ps.k_panels = panels_of(K); ps.last_ksteps = ksteps_of(K); ps.a_stream = -1;
    ps.a_wait = c == 0 ? first_a_wait : 0; ps.b_tm = b_tm; ps.b_mn = b_mn ? 1 : 0; ps.n_chunks = 1;
    const int tmem = bce ? (c & 1) * 224 : n0;
    const int buf = bce ? (c & 1) : c;
    if (b_mn) {
      const int boxes = ((pair ? w / 2 : w) + 63) / 64;
      set_chunk(ps, 0, n0, w, boxes, boxes * 8192, 0, tmem, buf, pair);
    } else {
      // the box is always kFwdChunk rows - half of that per CTA of a pair (rows beyond the matrix: zero fill)
      set_chunk(ps, 0, n0, w, 0, kFwdChunk * 128 / (pair ? 2 : 1), 0, tmem, buf, pair);
    }
    if (c == 1) L.u1 = n0 / 32;
    if (c == 2) L.u2 = n0 / 32;
    if (c == 3) L.u3 = n0 / 32;
    n0 += w;
  }
  return ip;
}

template <int kKind0, int kKind1, int kKind2, bool kPair>
int launch_chain_v(const CParams& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    MVAE_CUDA(cudaFuncSetAttribute(chain_kernel<kKind0, kKind1, kKind2, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_set = true;
  }
  static const int coop = env_int("MVAE_CHAIN_COOP", 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_slabs * p.parts + (kPair ? 2 : 1));   // slab CTAs + keeper (+ one idle CTA that completes the keeper's cluster)
  cfg.blockDim = dim3(kCThreads);
  cfg.dynamicSmemBytes = kSmemTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  int na = 0;
  if (g_pdl_next) {   // the kernel's prologue (barriers, TMEM, tensor-map prefetch) may run while the previous kernel drains
    g_pdl_next = 0;
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (kPair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  // Cooperative launch (every CTA resident at once: the grid barriers cannot deadlock) - except for the clustered kernels:
  // a launch that is cooperative AND clustered cannot be replayed by ncu (LaunchFailed), and it is not needed: the grid is
  // smaller than the SM count, one CTA fills an SM, and whatever holds an SM beside this kernel (weight-gradient GEMMs, the
  // text kernels, Adam, the data-parallel exchange) never waits for it - a CTA that finds no SM at first gets one when they end.
  if (coop && !kPair) {
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, chain_kernel<kKind0, kKind1, kKind2, kPair>, p);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(chain_kernel)", __FILE__, __LINE__);
  return 0;
}
template <int kKind0, int kKind1, int kKind2>
int launch_chain(const CParams& p, bool pair, cudaStream_t st) {
  return pair ? launch_chain_v<kKind0, kKind1, kKind2, true>(p, st) : launch_chain_v<kKind0, kKind1, kKind2, false>(p, st);
}
// CTA pairs (cta_group::2): an even number of slabs, and room for the keeper's cluster
bool chain_pair(int n_slabs) {
  static const int on = env_int("MVAE_CHAIN_PAIR", 1);
  return on != 0 && n_slabs % 2 == 0 && n_slabs + 2 <= chain_sms();
}
constexpr int kPairStage = 16384;   // resident-A layers of a pair: half a weight chunk per CTA, four stages in the same 64 KB
void pair_rings(CParams& p) { p.ring[1] = CRing{kArena, kPairStage, 4, 0}; }

long long* g_chain_dbg = nullptr;

void init_params(CParams& p, int rows_per_group, int n_slabs, unsigned int* err, int kind) {
  memset(&p, 0, sizeof(p));
  p.dbg = g_chain_dbg != nullptr ? g_chain_dbg + static_cast<long long>(kind) * 148 * 32 : nullptr;
  p.rows_per_group = rows_per_group;
  p.n_slabs = n_slabs;
  p.init_tm = -1;
  p.parts = 1;
  p.split_layer = -1;
  p.split_pass = -1;
  p.momentum = 0.1f;
  p.eps = 1e-5f;
  p.err = err;
  p.ring[0] = CRing{kArena, kRing1Stage, kRing1Stages, 0};
  p.ring[1] = p.ring[0];
}

}  // namespace

// bring-up: %globaltimer stamps of the chain kernels go to a device buffer of 4 x 148 x 32 int64 (null switches it off)
void set_chain_debug_times(void* ptr) { g_chain_dbg = static_cast<long long*>(ptr); }

// true if the chain kernels can run this step: all slabs whole, co-resident (plus the keeper CTA), dimensions inside the on-chip plan
namespace {
int chain_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}
}  // namespace
bool chain_supported(int B, int G, int n) {
  const int sms = chain_sms();
  return B >= 128 && B % 128 == 0 && G >= 1 && G <= 3 && (G * B) / 128 + 2 <= sms && n >= 16 && n <= 64 && n % 16 == 0;
}

int launch_chain_enc_fwd(const ChainEncFwd& a, cudaStream_t st) {
  CParams p;
  const int slabs = a.B / 128;
  init_params(p, a.B, slabs, a.err, 0);
  const int n2 = 2 * a.n;
  // The first layer (75 % of the kernel's FLOPs, tensor-bound on 32 SMs) split by columns over four CTAs per slab: each
  // part computes its columns of x W1^T, the parts swap pre-activations through L2 behind the BatchNorm barrier they
  // share anyway, and each normalises the whole slab; the two small layers after it are repeated by every part.
  static const int split_want = env_int("MVAE_CHAIN_SPLIT_FWD", 4);
  const int parts = split_want >= 4 && 4 * slabs + 1 <= chain_sms() ? 4 : (split_want >= 2 && 2 * slabs + 1 <= chain_sms() ? 2 : 1);
  const bool pair = parts == 1 && chain_pair(slabs);
  const int hv = pair ? 2 : 1;   // a CTA of a pair loads half of every weight tile
  if (chain_tmap(&p.tm[0], a.image, a.B, 784, 784, 64, 128)) return 1;
  if (chain_tmap(&p.tm[1], a.w1, 400, 784, 784, 64, parts == 4 ? 128 : (parts == 2 ? 256 : 208 / hv))) return 1;
  if (chain_tmap(&p.tm[2], a.w2, 200, 400, 400, 64, kFwdChunk / hv)) return 1;
  if (chain_tmap(&p.tm[3], a.w3, n2, 200, 200, 64, kFwdChunk / hv)) return 1;
  p.tm[4] = p.tm[3];
  // streamed image panel + all 400 rows of W1 per 64-wide k panel: 2 stages of 68 KB, or - pairs - 4 stages of 42 KB
  p.ring[0] = CRing{0, kPanel + 2 * 26624 / hv, pair ? 4 : 2, kPanel};
  if (parts == 4) p.ring[0] = CRing{0, 2 * kPanel, 4, kPanel};   // image panel + up to 128 rows of W1
  if (parts == 2) p.ring[0] = CRing{0, 3 * kPanel, 3, kPanel};   // image panel + up to 256 rows of W1
  if (pair) pair_rings(p);
  p.n_layers = 3;
  CLayer& l1 = p.layer[0];
  CLayer& l2 = p.layer[1];
  CLayer& l3 = p.layer[2];
  CPass& e1 = p.pass[0];
  e1.cfg = 0; e1.k_panels = panels_of(784); e1.last_ksteps = ksteps_of(784); e1.a_stream = 0; e1.a_wait = 0; e1.b_tm = 1; e1.b_mn = 0;
  e1.n_chunks = 2;
  set_chunk(e1, 0, 0, 208, 0, 26624 / hv, 0, 0, 0, pair);
  set_chunk(e1, 1, 208, 192, 0, 26624 / hv, 26624 / hv, 208, 1, pair);
  l1.kind = CE_FWD_BN; l1.N = 400; l1.n_chunks = 2;
  if (parts > 1) {
    p.parts = parts;
    p.split_layer = 0;
    p.split_pass = 0;
    const int n0_4[4] = {0, 128, 256, 320}, n_4[4] = {128, 128, 64, 80};   // whole 64-column panels (TMA stores / loads)
    const int n0_2[4] = {0, 256, 0, 0}, n_2[4] = {256, 144, 0, 0};
    for (int j = 0; j < 4; ++j) {
      p.part_n0[j] = parts == 4 ? n0_4[j] : n0_2[j];
      p.part_n[j] = parts == 4 ? n_4[j] : n_2[j];
    }
    e1.n_chunks = 1;
    set_chunk(e1, 0, 0, p.part_n[0], 0, p.part_n[0] * 128, 0, 0, 0);
    l1.n_chunks = 1;
  }
  int ip = 1;
  ip = add_resident_layer(p, ip, l2, 200, 400, 2, false, 2, false, pair);
  ip = add_resident_layer(p, ip, l3, n2, 200, 3, false, 2, false, pair);
  p.n_pass = ip;
  if (chain_tmap(&p.tmo[0], a.h1pre, a.B, 400, 400, 64, 128)) return 1;
  if (chain_tmap(&p.tmo[1], a.h1, a.B, 400, 400, 64, 128)) return 1;
  if (chain_tmap(&p.tmo[2], a.h2pre, a.B, 200, 200, 64, 128)) return 1;
  if (chain_tmap(&p.tmo[3], a.h2, a.B, 200, 200, 64, 128)) return 1;
  l1.bias = a.b1; l1.gamma = a.gamma1; l1.beta = a.beta1; l1.stat0 = a.st1; l1.stat1 = a.st1 + 400;
  l1.save_mean = a.sv1; l1.save_rstd = a.sv1 + 400; l1.running_mean = a.rm1; l1.running_var = a.rv1; l1.bn_updates = a.bn_updates;
  l1.counter = a.counters; l1.out_pre = a.h1pre; l1.out_post = a.h1; l1.write_arena = 1;
  l2.kind = CE_FWD_BN; l2.N = 200;
  l2.bias = a.b2; l2.gamma = a.gamma2; l2.beta = a.beta2; l2.stat0 = a.st2; l2.stat1 = a.st2 + 200;
  l2.save_mean = a.sv2; l2.save_rstd = a.sv2 + 200; l2.running_mean = a.rm2; l2.running_var = a.rv2; l2.bn_updates = a.bn_updates;
  l2.counter = a.counters + 1; l2.out_pre = a.h2pre; l2.out_post = a.h2; l2.write_arena = 1;
  l3.kind = CE_FWD_STORE; l3.N = n2;
  l3.bias = a.b3; l3.out_f32 = a.enc; l3.ld_out = n2;
  return launch_chain<CE_FWD_BN, CE_FWD_BN, CE_FWD_STORE>(p, pair, st);
}

int launch_chain_dec_fwd(const ChainDecFwd& a, cudaStream_t st) {
  CParams p;
  const int R = a.G * a.B, G = a.G;
  init_params(p, a.B, R / 128, a.err, 1);
  const bool pair = chain_pair(R / 128);
  const int hv = pair ? 2 : 1;
  if (pair) pair_rings(p);
  if (chain_tmap(&p.tm[0], a.z, R, a.n, a.n, 64, 128)) return 1;
  if (chain_tmap(&p.tm[1], a.w1, 200, a.n, a.n, 64, kFwdChunk / hv)) return 1;
  if (chain_tmap(&p.tm[2], a.w2, 400, 200, 200, 64, kFwdChunk / hv)) return 1;
  if (chain_tmap(&p.tm[3], a.w3, 784, 400, 400, 64, kFwdChunk / hv)) return 1;
  p.tm[4] = p.tm[3];
  p.init_tm = 0;
  p.init_panels = panels_of(a.n);
  p.n_layers = 3;
  CLayer& l1 = p.layer[0];
  CLayer& l2 = p.layer[1];
  CLayer& l3 = p.layer[2];
  int ip = 0;
  ip = add_resident_layer(p, ip, l1, 200, a.n, 1, false, 1, false, pair);
  ip = add_resident_layer(p, ip, l2, 400, 200, 2, false, 2, false, pair);
  ip = add_resident_layer(p, ip, l3, 784, 400, 3, false, 2, true, pair);
  p.n_pass = ip;
  if (chain_tmap(&p.tmo[0], a.g1pre, R, 200, 200, 64, 128)) return 1;
  if (chain_tmap(&p.tmo[1], a.g1, R, 200, 200, 64, 128)) return 1;
  if (chain_tmap(&p.tmo[2], a.g2pre, R, 400, 400, 64, 128)) return 1;
  if (chain_tmap(&p.tmo[3], a.g2, R, 400, 400, 64, 128)) return 1;
  l1.kind = CE_FWD_BN; l1.N = 200;
  l1.bias = a.b1; l1.gamma = a.gamma1; l1.beta = a.beta1; l1.stat0 = a.st1; l1.stat1 = a.st1 + G * 200;
  l1.save_mean = a.sv1; l1.save_rstd = a.sv1 + G * 200; l1.running_mean = a.rm1; l1.running_var = a.rv1; l1.bn_updates = 1;
  l1.counter = a.counters; l1.out_pre = a.g1pre; l1.out_post = a.g1; l1.write_arena = 1;
  l2.kind = CE_FWD_BN; l2.N = 400;
  l2.bias = a.b2; l2.gamma = a.gamma2; l2.beta = a.beta2; l2.stat0 = a.st2; l2.stat1 = a.st2 + G * 400;
  l2.save_mean = a.sv2; l2.save_rstd = a.sv2 + G * 400; l2.running_mean = a.rm2; l2.running_var = a.rv2; l2.bn_updates = 1;
  l2.counter = a.counters + 3; l2.out_pre = a.g2pre; l2.out_post = a.g2; l2.write_arena = 1;
  l3.kind = CE_BCE; l3.N = 784;
  l3.bias = a.b3; l3.target = a.image; l3.target_rows = a.B;
  for (int g = 0; g < 3; ++g) l3.bce_scale[g] = a.bce_scale[g];
  l3.loss = a.loss; l3.dbias = a.dbias3; l3.dlog = a.dlog; l3.probs = a.probs;
  return launch_chain<CE_FWD_BN, CE_FWD_BN, CE_BCE>(p, pair, st);
}

int launch_chain_dec_bwd(const ChainDecBwd& a, cudaStream_t st) {
  CParams p;
  const int R = a.G * a.B, G = a.G;
  init_params(p, a.B, R / 128, a.err, 2);
  if (chain_tmap(&p.tm[0], a.dlog, R, 784, 784, 64, 128)) return 1;
  if (chain_tmap(&p.tm[1], a.w3, 784, 400, 400, 64, 64)) return 1;   // dgrad: W[out, in] read as the MN-major B operand
  if (chain_tmap(&p.tm[2], a.w2, 400, 200, 200, 64, 64)) return 1;
  if (chain_tmap(&p.tm[3], a.w1, 200, a.n, a.n, 64, 64)) return 1;
  p.tm[4] = p.tm[3];
  // streamed dlogits panel + all 400 columns of W3 per 64-wide k panel: 2 stages of 72 KB, or - pairs - 3 stages of 48 KB
  const bool pair = chain_pair(R / 128);
  p.ring[0] = pair ? CRing{0, kPanel + 4 * 8192, 3, kPanel} : CRing{0, kPanel + 7 * 8192, 2, kPanel};
  if (pair) pair_rings(p);
  p.n_layers = 3;
  CLayer& l2 = p.layer[0];
  CLayer& l1 = p.layer[1];
  CLayer& l0 = p.layer[2];
  CPass& g3 = p.pass[0];
  g3.cfg = 0; g3.k_panels = panels_of(784); g3.last_ksteps = ksteps_of(784); g3.a_stream = 0; g3.a_wait = 0; g3.b_tm = 1; g3.b_mn = 1;
  g3.n_chunks = 2;
  if (pair) {   // per CTA: 128 + 72 columns = 2 + 2 boxes
    set_chunk(g3, 0, 0, 256, 2, 16384, 0, 0, 0, true);
    set_chunk(g3, 1, 256, 144, 2, 16384, 16384, 256, 1, true);
  } else {
    set_chunk(g3, 0, 0, 256, 4, 32768, 0, 0, 0);
    set_chunk(g3, 1, 256, 144, 3, 24576, 32768, 256, 1);
  }
  l2.kind = CE_DGRAD_BN; l2.N = 400; l2.n_chunks = 2;
  int ip = 1;
  ip = add_resident_layer(p, ip, l1, 200, 400, 2, true, 2, false, pair);
  ip = add_resident_layer(p, ip, l0, a.n, 200, 3, true, 2, false, pair);
  p.n_pass = ip;
  if (chain_tmap(&p.tmo[1], a.dy2, R, 400, 400, 64, 128)) return 1;
  if (chain_tmap(&p.tmo[3], a.dy1, R, 200, 200, 64, 128)) return 1;
  p.tmo[0] = p.tmo[1];
  p.tmo[2] = p.tmo[3];
  l2.gamma = a.gamma2; l2.beta = a.beta2; l2.stat0 = a.sb2; l2.stat1 = a.sb2 + G * 400; l2.save_mean = a.sv2; l2.save_rstd = a.sv2 + G * 400;
  l2.counter = a.counters; l2.hpre = a.g2pre; l2.out_post = a.dy2; l2.write_arena = 1; l2.dgamma = a.dgamma2; l2.dbeta = a.dbeta2;
  l1.kind = CE_DGRAD_BN; l1.N = 200;
  l1.gamma = a.gamma1; l1.beta = a.beta1; l1.stat0 = a.sb1; l1.stat1 = a.sb1 + G * 200; l1.save_mean = a.sv1; l1.save_rstd = a.sv1 + G * 200;
  l1.counter = a.counters + 3; l1.hpre = a.g1pre; l1.out_post = a.dy1; l1.write_arena = 1; l1.dgamma = a.dgamma1; l1.dbeta = a.dbeta1;
  l0.kind = CE_DGRAD_STORE; l0.N = a.n;
  l0.out_f32 = a.dz; l0.ld_out = a.n;
  return launch_chain<CE_DGRAD_BN, CE_DGRAD_BN, CE_DGRAD_STORE>(p, pair, st);
}

int launch_chain_enc_bwd(const ChainEncBwd& a, cudaStream_t st) {
  CParams p;
  const int slabs = a.B / 128;
  init_params(p, a.B, slabs, a.err, 3);
  const int n2 = 2 * a.n;
  if (chain_tmap(&p.tm[0], a.denc, a.B, n2, n2, 64, 128)) return 1;
  if (chain_tmap(&p.tm[1], a.w3, n2, 200, 200, 64, 64)) return 1;
  if (chain_tmap(&p.tm[2], a.w2, 200, 400, 400, 64, 64)) return 1;
  p.tm[3] = p.tm[2];
  p.tm[4] = p.tm[2];
  p.init_tm = 0;
  p.init_panels = panels_of(n2);
  p.n_layers = 2;
  CLayer& l2 = p.layer[0];
  CLayer& l1 = p.layer[1];
  int ip = 0;
  // 33 of 148 SMs would carry this kernel: the last layer's 400 columns are split over up to four CTAs per slab instead
  // (each repeats the small first layer on its own), which is worth more here than pairing the slabs
  static const int split_on = env_int("MVAE_CHAIN_SPLIT", 1);
  const int parts = !split_on ? 1 : (a.max_parts >= 4 && 4 * slabs + 1 <= chain_sms() ? 4 : (a.max_parts >= 2 && 2 * slabs + 1 <= chain_sms() ? 2 : 1));
  const bool pair = parts == 1 && chain_pair(slabs);
  if (pair) pair_rings(p);
  ip = add_resident_layer(p, ip, l2, 200, n2, 1, true, 1, false, pair);
  if (parts > 1) {
    p.parts = parts;
    p.split_layer = 1;
    p.split_pass = ip;
    const int n0_4[4] = {0, 128, 256, 320}, n_4[4] = {128, 128, 64, 80};
    const int n0_2[4] = {0, 256, 0, 0}, n_2[4] = {256, 144, 0, 0};
    for (int j = 0; j < 4; ++j) {
      p.part_n0[j] = parts == 4 ? n0_4[j] : n0_2[j];
      p.part_n[j] = parts == 4 ? n_4[j] : n_2[j];
    }
    ip = add_resident_layer(p, ip, l1, p.part_n[0], 200, 2, true, 2, false, false);   // one chunk per part, widths from part_n
  } else {
    ip = add_resident_layer(p, ip, l1, 400, 200, 2, true, 2, false, pair);
  }
  p.n_pass = ip;
  if (chain_tmap(&p.tmo[1], a.dye2, a.B, 200, 200, 64, 128)) return 1;
  if (chain_tmap(&p.tmo[3], a.dye1, a.B, 400, 400, 64, 128)) return 1;
  p.tmo[0] = p.tmo[1];
  p.tmo[2] = p.tmo[3];
  l2.kind = CE_DGRAD_BN; l2.N = 200;
  l2.gamma = a.gamma2; l2.beta = a.beta2; l2.stat0 = a.sb2; l2.stat1 = a.sb2 + 200; l2.save_mean = a.sv2; l2.save_rstd = a.sv2 + 200;
  l2.counter = a.counters; l2.hpre = a.h2pre; l2.out_post = a.dye2; l2.write_arena = 1; l2.dgamma = a.dgamma2; l2.dbeta = a.dbeta2;
  l1.kind = CE_DGRAD_BN; l1.N = 400;
  l1.gamma = a.gamma1; l1.beta = a.beta1; l1.stat0 = a.sb1; l1.stat1 = a.sb1 + 400; l1.save_mean = a.sv1; l1.save_rstd = a.sv1 + 400;
  l1.counter = a.counters + 1; l1.hpre = a.h1pre; l1.out_post = a.dye1; l1.write_arena = 0; l1.dgamma = a.dgamma1; l1.dbeta = a.dbeta1;
  return launch_chain<CE_DGRAD_BN, CE_DGRAD_BN, -1>(p, pair, st);
}

}  // namespace mvae
