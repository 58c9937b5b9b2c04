// Error state, environment knobs and the generic GEMM entry of the C ABI (include/mvae_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/mvae_b200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace mvae {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", static_cast<int>(e), cudaGetErrorString(e), file, line, what);
  return 2;
}

static std::atomic<long long> g_noted_launches{0};
void note_launch(int n) { g_noted_launches.fetch_add(n, std::memory_order_relaxed); }
long long noted_launches() { return g_noted_launches.load(std::memory_order_relaxed); }

thread_local int g_pdl_next = 0;

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  if (s == nullptr || *s == 0) return dflt;
  return atoi(s);
}

}  // namespace mvae

using namespace mvae;

extern "C" {

const char* mvae_last_error(void) { return g_err; }

int mvae_abi_version(void) { return MVAE_ABI_VERSION; }

int mvae_device_check(int device) {
  cudaDeviceProp prop;
  MVAE_CUDA(cudaGetDeviceProperties(&prop, device));
  MVAE_REQUIRE(prop.major == 10, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
               prop.minor);
  return 0;
}

void mvae_debug_chain_times(void* device_int64_buffer) { set_chain_debug_times(device_int64_buffer); }

void mvae_debug_gemm_times(void* device_int64_buffer, int epilogue_kind) {
  set_gemm_debug_times(device_int64_buffer, epilogue_kind);
}

static int gemm_entry(const mvae_gemm_args* a, const mvae_conv_geometry* cg, int patch_operand, void* stream);

int mvae_gemm(const mvae_gemm_args* a, void* stream) { return gemm_entry(a, nullptr, 0, stream); }

int mvae_conv_gemm(const mvae_gemm_args* a, const mvae_conv_geometry* geometry, int patch_operand, void* stream) {
  MVAE_REQUIRE(geometry != nullptr && (patch_operand == 1 || patch_operand == 2), "mvae_conv_gemm: geometry / operand");
  return gemm_entry(a, geometry, patch_operand, stream);
}

int mvae_convt_class_gemm(const mvae_convt_class* c, int dtype, const void* x, const void* weight, int64_t ld_tap, void* out,
                          int64_t ldc, int out_dtype, void* stream) {
  MVAE_REQUIRE(c != nullptr && x != nullptr && weight != nullptr && out != nullptr, "mvae_convt_class_gemm: null argument");
  MVAE_REQUIRE(c->taps_h >= 1 && c->taps_h <= 8 && c->taps_w >= 1 && c->taps_w <= 8, "mvae_convt_class_gemm: 1..8 taps per axis");
  MVAE_REQUIRE(c->count_h > 0 && c->count_w > 0, "mvae_convt_class_gemm: empty class");
  GemmDesc g;
  g.kind = dtype;
  g.M = c->batch * c->count_h * c->count_w;
  g.N = c->out_channels;
  g.K = c->taps_h * c->taps_w * c->channels;
  g.A = x; g.lda = 0; g.a_mn = 0;
  g.B = weight; g.ldb = ld_tap; g.b_mn = 1;
  g.epi.kind = EPI_STORE;
  g.epi.C = out; g.epi.ldc = ldc; g.epi.c_dtype = out_dtype;
  g.epi.rows_per_group = 1 << 30;
  ConvGather& cg = g.gather;
  cg.mode = 3;
  cg.X = x;
  cg.H = c->in_h; cg.W = c->in_w; cg.C = c->channels;
  cg.ksize = c->taps_h; cg.ksize_w = c->taps_w; cg.stride = 1; cg.pad = c->pad_h; cg.pad_w = c->pad_w;
  cg.Ho = c->count_h; cg.Wo = c->count_w;
  cg.sw = c->channels; cg.sh = static_cast<long long>(c->in_w) * c->channels; cg.sn = cg.sh * c->in_h;
  cg.extent = cg.sn * c->batch;
  cg.kk = c->kernel;
  for (int t = 0; t < 8; ++t) {
    cg.kh_tab[t] = static_cast<signed char>(t < c->taps_h ? c->kh[t] : 0);
    cg.kw_tab[t] = static_cast<signed char>(t < c->taps_w ? c->kw[t] : 0);
    MVAE_REQUIRE(cg.kh_tab[t] >= 0 && cg.kh_tab[t] < c->kernel && cg.kw_tab[t] >= 0 && cg.kw_tab[t] < c->kernel,
                 "mvae_convt_class_gemm: tap table out of range");
  }
  cg.sc_hout = c->out_h; cg.sc_wout = c->out_w; cg.sc_stride = c->stride; cg.sc_a = c->a; cg.sc_b = c->b;
  note_launch(1);
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

int mvae_convt_axis_classes(int kernel, int stride, int pad, int size_in, int* count, int* taps, int* pad_lo, int* kh) {
  if (kernel <= 0 || stride <= 0 || stride > 4 || pad < 0 || size_in <= 0) return -1;
  const int size_out = (size_in - 1) * stride - 2 * pad + kernel;
  for (int a = 0; a < stride; ++a) {
    const int r = (a + pad) % stride;
    const int n = kernel > r ? (kernel - r + stride - 1) / stride : 0;
    if (n > 8) return -1;
    taps[a] = n;
    count[a] = size_out > a ? (size_out - a + stride - 1) / stride : 0;
    pad_lo[a] = n - 1 - (a + pad) / stride;
    for (int t = 0; t < 8; ++t) kh[a * 8 + t] = t < n ? r + stride * (n - 1 - t) : 0;
  }
  return size_out;
}

int mvae_convt_gemm(int dtype, int batch, int in_h, int in_w, int channels, int out_channels, int kernel, int stride, int pad,
                    const void* x, const void* weight, int64_t ld_tap, void* out, int64_t ldc, int out_dtype, void* stream) {
  MVAE_REQUIRE(x != nullptr && weight != nullptr && out != nullptr, "mvae_convt_gemm: null argument");
  MVAE_REQUIRE(batch > 0 && in_h > 0 && in_w > 0 && kernel > 0 && stride > 0 && stride <= 4 && pad >= 0, "mvae_convt_gemm: bad geometry");
  const int out_h = (in_h - 1) * stride - 2 * pad + kernel, out_w = (in_w - 1) * stride - 2 * pad + kernel;
  // per-axis parity classes: all must have the same tap count and the same grid
  int cnt_h[4], cnt_w[4], taps[4], pad_lo[4], kh[32];
  MVAE_REQUIRE(mvae_convt_axis_classes(kernel, stride, pad, in_h, cnt_h, taps, pad_lo, kh) == out_h &&
                   mvae_convt_axis_classes(kernel, stride, pad, in_w, cnt_w, taps, pad_lo, kh) == out_w,
               "mvae_convt_gemm: bad geometry");
  const int taps0 = taps[0], cnt_h0 = cnt_h[0], cnt_w0 = cnt_w[0];
  GemmDesc g;
  ConvGather& cg = g.gather;
  for (int a = 0; a < stride; ++a) {
    if (taps[a] != taps0 || cnt_h[a] != cnt_h0 || cnt_w[a] != cnt_w0 || taps0 < 1 || taps0 > 8 || cnt_h0 < 1 || cnt_w0 < 1) {
      set_error("mvae_convt_gemm: parity classes of k=%d s=%d p=%d differ in shape; use mvae_convt_class_gemm per class", kernel,
                stride, pad);
      return 4;
    }
    cg.ax_pad[a] = static_cast<signed char>(pad_lo[a]);
    for (int t = 0; t < taps[a]; ++t) cg.ax_k[a][t] = static_cast<signed char>(kh[a * 8 + t]);
  }
  g.kind = dtype;
  g.M = batch * cnt_h0 * cnt_w0;
  g.N = out_channels;
  g.K = taps0 * taps0 * channels;
  g.A = x; g.lda = 0; g.a_mn = 0;
  g.B = weight; g.ldb = ld_tap; g.b_mn = 1;
  g.epi.kind = EPI_STORE;
  g.epi.C = out; g.epi.ldc = ldc; g.epi.c_dtype = out_dtype;
  g.epi.rows_per_group = 1 << 30;
  cg.mode = 4;
  cg.X = x;
  cg.H = in_h; cg.W = in_w; cg.C = channels;
  cg.ksize = taps0; cg.ksize_w = taps0; cg.stride = 1; cg.pad = cg.ax_pad[0]; cg.pad_w = cg.ax_pad[0];
  cg.Ho = cnt_h0; cg.Wo = cnt_w0;
  cg.sw = channels; cg.sh = static_cast<long long>(in_w) * channels; cg.sn = cg.sh * in_h;
  cg.extent = cg.sn * batch;
  cg.kk = kernel;
  cg.sc_hout = out_h; cg.sc_wout = out_w; cg.sc_stride = stride; cg.sc_a = 0; cg.sc_b = 0;
  note_launch(1);
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

static int gemm_entry(const mvae_gemm_args* a, const mvae_conv_geometry* cg, int patch_operand, void* stream) {
  MVAE_REQUIRE(a != nullptr, "mvae_gemm: null args");
  GemmDesc g;
  if (cg != nullptr) {
    MVAE_REQUIRE(cg->stride_c == 1, "mvae_conv_gemm: the image must be channels-last (stride_c == 1)");
    g.gather.mode = patch_operand;
    g.gather.X = patch_operand == 1 ? a->A : a->B;
    g.gather.H = cg->height; g.gather.W = cg->width; g.gather.C = cg->channels;
    g.gather.ksize = cg->kernel; g.gather.stride = cg->stride; g.gather.pad = cg->pad;
    g.gather.Ho = (cg->height + 2 * cg->pad - cg->kernel) / cg->stride + 1;
    g.gather.Wo = (cg->width + 2 * cg->pad - cg->kernel) / cg->stride + 1;
    g.gather.sn = cg->stride_n; g.gather.sh = cg->stride_h; g.gather.sw = cg->stride_w;
    g.gather.extent = (cg->batch - 1) * cg->stride_n + (cg->height - 1) * cg->stride_h + (cg->width - 1) * cg->stride_w + cg->channels;
    const long long pixels = static_cast<long long>(cg->batch) * g.gather.Ho * g.gather.Wo;
    MVAE_REQUIRE(pixels == (patch_operand == 1 ? a->M : a->K), "mvae_conv_gemm: %lld output pixels do not match the GEMM shape",
                 pixels);
  }
  g.kind = a->dtype == MVAE_DT_F32X3 ? MVAE_F32 : a->dtype;
  g.M = a->M; g.N = a->N; g.K = a->K;
  if (a->dtype == MVAE_DT_F32X3) {
    const long long kp = (a->K + 3) / 4 * 4 + 4;
    const long long need_a = 12ll * (a->M + 4) * kp, need_b = 12ll * (a->N + 4) * kp;
    MVAE_REQUIRE(a->x3_scratch != nullptr && a->x3_scratch_bytes >= need_a + need_b + 512,
                 "mvae_gemm: MVAE_DT_F32X3 needs x3_scratch of at least %lld bytes", need_a + need_b + 512);
    g.x3 = 1;
    g.x3_a = a->x3_scratch;
    g.x3_b = static_cast<char*>(a->x3_scratch) + (need_a + 255) / 256 * 256;
  }
  g.A = a->A; g.lda = a->lda; g.a_mn = a->a_major;
  g.B = a->B; g.ldb = a->ldb; g.b_mn = a->b_major;
  g.block_n = a->block_n; g.split_k = a->split_k; g.stages = a->stages;
  g.epi.kind = a->accumulate ? EPI_ATOMIC : EPI_STORE;
  g.epi.C = a->C; g.epi.ldc = a->ldc; g.epi.c_dtype = a->c_dtype;
  g.epi.bias = a->bias;
  g.epi.stat0 = a->col_sum; g.epi.stat1 = a->col_sumsq;
  g.epi.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : (1 << 30);
  g.dbg = reinterpret_cast<long long*>(a->debug_times);
  if (a->act != MVAE_ACT_NONE) {
    MVAE_REQUIRE(a->act == MVAE_ACT_SWISH && !a->accumulate && cg == nullptr, "mvae_gemm: fused activation = Swish on a plain, non-accumulating GEMM");
    MVAE_REQUIRE((a->act_out != nullptr) != (a->act_pre != nullptr), "mvae_gemm: fused activation needs exactly one of act_out / act_pre");
    if (a->act_out != nullptr) {
      MVAE_REQUIRE(a->col_sum == nullptr && a->col_sumsq == nullptr, "mvae_gemm: no column statistics with a fused forward activation");
      g.epi.kind = EPI_STORE_ACT;
      g.epi.probs = a->act_out;
    } else {
      MVAE_REQUIRE(a->bias == nullptr && a->col_sumsq == nullptr, "mvae_gemm: activation backward takes no bias / col_sumsq");
      g.epi.kind = EPI_DGRAD_ACT;
      g.epi.hpre = a->act_pre;
      g.epi.ldh = a->ld_act_pre;
      g.epi.rows_per_group = 1 << 30;
    }
  }
  if (a->bce_target != nullptr) {
    MVAE_REQUIRE(a->act == MVAE_ACT_NONE && !a->accumulate && cg == nullptr && a->col_sumsq == nullptr,
                 "mvae_gemm: the BCE epilogue needs a plain, non-accumulating GEMM without activation / col_sumsq");
    MVAE_REQUIRE(a->bce_target_rows > 0 && a->rows_per_group > 0 && (a->M + a->rows_per_group - 1) / a->rows_per_group <= 4,
                 "mvae_gemm: the BCE epilogue needs target rows and at most 4 row groups");
    g.epi.kind = EPI_BCE;
    g.epi.target = a->bce_target; g.epi.ldt = a->ld_bce_target; g.epi.target_rows = a->bce_target_rows;
    for (int t = 0; t < 4; ++t) g.epi.bce_scale[t] = a->bce_scale[t];
    g.epi.loss = a->bce_loss;
    g.epi.probs = a->bce_probs;
    g.epi.row_w = a->bce_row_weight;
  }
  note_launch(1);
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
