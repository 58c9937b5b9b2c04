// Device helpers shared by the latent-path kernels (tail.cu for MNIST, conv_ops.cu for the conv models):
// Philox4x32-10 + Box-Muller noise, and the ProductOfExperts arithmetic on one latent element with its gradient.
// Reference: ProductOfExperts mnist/model.py:173-185 == celeba/model.py:224-235, reparametrize mnist/model.py:24-30.
#pragma once
#include "common.cuh"
#include "../../include/mvae_b200.h"

namespace mvae {

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                           uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// Two standard normals for element pair `pair_index` of the step.
__device__ __forceinline__ float2 normal_pair(unsigned long long seed, uint32_t step, unsigned long long pair_index) {
  uint32_t r[4];
  philox4x32(static_cast<uint32_t>(pair_index), static_cast<uint32_t>(pair_index >> 32), step, 0x6d766165u,
             static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  const float u1 = (static_cast<float>(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (static_cast<float>(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(rad * c, rad * s);
}

// Four standard normals from ONE Philox call (all four output words): two Box-Muller pairs.  `tag` separates streams.
__device__ __forceinline__ float4 normal_quad(unsigned long long seed, uint32_t step, unsigned long long index, uint32_t tag) {
  uint32_t r[4];
  philox4x32(static_cast<uint32_t>(index), static_cast<uint32_t>(index >> 32), step, tag, static_cast<uint32_t>(seed),
             static_cast<uint32_t>(seed >> 32), r);
  float o[4];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = (static_cast<float>(r[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = (static_cast<float>(r[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    o[2 * h] = rad * cs;
    o[2 * h + 1] = rad * sn;
  }
  return make_float4(o[0], o[1], o[2], o[3]);
}

// Fast SFU forms (ex2/lg2/rcp.approx + one FMA): relative error ~1e-6 for the magnitudes that occur here
// (|logvar| < ~20), far inside the parity tolerances; the tail kernels are instruction-bound, these cut the
// transcendental cost 5x.
__device__ __forceinline__ float fexp(float x) { return __expf(x); }
__device__ __forceinline__ float flog(float x) { return __logf(x); }
__device__ __forceinline__ float frcp(float x) { return __fdividef(1.f, x); }

// ---------------------------------------------------------------- PoE on one latent element, M <= 2 experts
// Everything the backward needs is kept so that no transcendental is evaluated twice:
// exp(logvar) == pd_var and exp(logvar/2) == sqrt(pd_var) by construction.
struct Poe {
  float mu, logvar, pd_var;
  float var[2], inv[2];  // var_i = exp(logvar_i) + eps, 1/var_i
  float invS;            // REF: 1/sum(var_i); PRECISION: 1/sum(1/var_i) (+1 with the prior expert)
};
// present[i]: expert i takes part.  REF: mnist/model.py:180-185 (variance-weighted mu).
template <bool kNeedLogvar>
__device__ __forceinline__ Poe poe_eval(int mode, int prior, float eps, const float (&m)[2], const float (&lv)[2],
                                        const bool (&present)[2]) {
  Poe r;
  r.var[0] = r.var[1] = 1.f;
  r.inv[0] = r.inv[1] = 1.f;
  float num = 0.f, S = 0.f, P = (mode == MVAE_POE_PRECISION && prior) ? 1.f : 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if (present[i]) {
      const float var = fexp(lv[i]) + eps;
      const float inv = frcp(var);
      r.var[i] = var;
      r.inv[i] = inv;
      num += m[i] * (mode == MVAE_POE_REF ? var : inv);
      S += var;
      P += inv;
    }
  r.pd_var = frcp(P);
  if (mode == MVAE_POE_REF) {
    r.invS = frcp(S);
    r.mu = num * r.invS;
  } else {
    r.invS = r.pd_var;
    r.mu = num * r.pd_var;
  }
  r.logvar = kNeedLogvar ? flog(r.pd_var) : 0.f;
  return r;
}
// The same product from per-expert parts computed once (var_i = exp(logvar_i) + eps, 1/var_i): a kernel that evaluates several
// ELBO terms on one sample shares the transcendentals between the terms.  Same operations in the same order as poe_eval, so
// the results are bit-identical.
struct PoePart {
  float var, inv;
};
__device__ __forceinline__ PoePart poe_part(float lv, float eps) {
  PoePart p;
  p.var = fexp(lv) + eps;
  p.inv = frcp(p.var);
  return p;
}
template <bool kNeedLogvar>
__device__ __forceinline__ Poe poe_combine(int mode, int prior, const float (&m)[2], const PoePart (&pt)[2], const bool (&present)[2]) {
  Poe r;
  r.var[0] = r.var[1] = 1.f;
  r.inv[0] = r.inv[1] = 1.f;
  float num = 0.f, S = 0.f, P = (mode == MVAE_POE_PRECISION && prior) ? 1.f : 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if (present[i]) {
      r.var[i] = pt[i].var;
      r.inv[i] = pt[i].inv;
      num += m[i] * (mode == MVAE_POE_REF ? pt[i].var : pt[i].inv);
      S += pt[i].var;
      P += pt[i].inv;
    }
  r.pd_var = frcp(P);
  if (mode == MVAE_POE_REF) {
    r.invS = frcp(S);
    r.mu = num * r.invS;
  } else {
    r.invS = r.pd_var;
    r.mu = num * r.pd_var;
  }
  r.logvar = kNeedLogvar ? flog(r.pd_var) : 0.f;
  return r;
}

// Gradients w.r.t. expert i's (mu_i, logvar_i) given d(mu), d(logvar) of the product.
__device__ __forceinline__ void poe_grad(int mode, float eps, const Poe& r, int i, float m_i, float dmu, float dlv,
                                         float& dm_i, float& dlv_i) {
  const float e = r.var[i] - eps;  // d var_i / d logvar_i = exp(logvar_i) (the eps is additive)
  if (mode == MVAE_POE_REF) {
    dm_i = dmu * r.var[i] * r.invS;
    const float dvar = dmu * (m_i - r.mu) * r.invS + dlv * r.pd_var * r.inv[i] * r.inv[i];
    dlv_i = dvar * e;
  } else {
    dm_i = dmu * r.inv[i] * r.invS;
    const float dT = (dmu * (m_i - r.mu) - dlv) * r.invS;
    dlv_i = -dT * r.inv[i] * r.inv[i] * e;
  }
}

}  // namespace mvae
