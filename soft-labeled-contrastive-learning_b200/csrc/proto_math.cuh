// proto_math.cuh -- per-pixel arithmetic of the prototype loss shared by proto.cu (thread-per-pixel-quad kernels) and
// the fused target tile kernel of class_sums.cu.  Reference: MPCL.forward, utils/loss.py:529-571; SURVEY.md A.1.
#pragma once

#include "common.cuh"

#include <math.h>

namespace slcl {

struct MarginConst {
  float inv_t /* 1 / T */, scale /* T / T_b */, cos_m, sin_m, th, mm;
  int easy, normalize;
};

// One pixel of MPCL.forward (utils/loss.py:529-571) and its closed-form
// derivative.  cosv: cosines; M: positive weights (one-hot or
// soft row); selw: pixel_sel_loc value (1 when absent).  Returns the row loss;
// coef[k] = selw * dl/dcos_k * inv_n, coef[K] = selw * (sum_k dl/dcos_k cos_k) * inv_n^2.
template <int K>
__device__ __forceinline__ float margin_row(const float (&cosv)[K], const float (&M)[K], float selw, float inv_n,
                                            const MarginConst& mc, float (&coef)[K + 1]) {
  // Divisions by T are multiplications by a host-computed 1/T (<= 1 ulp from the reference's
  // torch.div, far inside the 1e-4 budget); cs/sine uses rsqrt of the clamped 1-cos^2.
  float plain[K], marg[K], dphi[K];
  float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float cs = cosv[k];
    plain[k] = cs * mc.inv_t;                                        // :530
    const float u = 1.0f - cs * cs;                                  // :534
    const float uc = fminf(fmaxf(u, 1e-4f), 1.0f);
    const float rs = rsqrtf(uc);
    const float sine = uc * rs;
    const float phi = cs * mc.cos_m - sine * mc.sin_m;               // :536
    const bool on = mc.easy ? (cs > 0.f) : (cs > mc.th);             // :538-541
    const float ph = on ? phi : (mc.easy ? cs : cs - mc.mm);
    marg[k] = ph * mc.inv_t;                                         // :543
    const bool unclamped = (u >= 1e-4f) && (u <= 1.0f);
    dphi[k] = on ? (unclamped ? fmaf(mc.sin_m * cs, rs, mc.cos_m) : mc.cos_m) : 1.0f;
    m1 = fmaxf(m1, plain[k]);                                        // :531 (detached)
    m2 = fmaxf(m2, marg[k]);                                         // :545 (detached)
  }
  float z[K], ez[K];
  float s = 0.f, sum_m = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    z[k] = (plain[k] - m1) * (1.0f - M[k]) + (marg[k] - m2) * M[k];  // :550-554
    ez[k] = expf(z[k]);
    s += ez[k];
    sum_m += M[k];
  }
  const float den = s + 1e-4f;                                       // :556
  const float lse = logf(den);
  const float inv_den = 1.0f / den;
  float row = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) row += M[k] * (z[k] - lse);            // :562 / :568
  float bsum = 0.f;
  const float w_row = selw * inv_n * mc.inv_t;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float dz = -mc.scale * (M[k] - sum_m * ez[k] * inv_den);
    const float e = dz * ((1.0f - M[k]) + M[k] * dphi[k]);           // times 1/T, folded into w_row
    coef[k] = w_row * e;
    bsum = fmaf(e, cosv[k], bsum);
  }
  coef[K] = mc.normalize ? w_row * bsum * inv_n : 0.f;
  return -mc.scale * row;
}

// The same pixel for the two inputs no reference caller differentiates: returns the row loss l and
//   dM[k] = dl / dM_k = -(T/T_b) [ (z_k - lse) + (marg'_k - plain'_k) (M_k - p_k sum_j M_j) ],   p_k = exp(z_k) / (s + 1e-4)
// (M enters through z_k = plain'_k (1 - M_k) + marg'_k M_k, :550-554, and through the weights of :562; plain' / marg' are
// the max-shifted logits, whose shifts are detached, :531 / :545).  Same expressions as margin_row up to `lse`.
template <int K>
__device__ __forceinline__ float margin_row_aux(const float (&cosv)[K], const float (&M)[K], const MarginConst& mc,
                                                float (&dM)[K]) {
  float plain[K], marg[K];
  float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float cs = cosv[k];
    plain[k] = cs * mc.inv_t;
    const float u = 1.0f - cs * cs;
    const float uc = fminf(fmaxf(u, 1e-4f), 1.0f);
    const float rs = rsqrtf(uc);
    const float sine = uc * rs;
    const float phi = cs * mc.cos_m - sine * mc.sin_m;
    const bool on = mc.easy ? (cs > 0.f) : (cs > mc.th);
    const float ph = on ? phi : (mc.easy ? cs : cs - mc.mm);
    marg[k] = ph * mc.inv_t;
    m1 = fmaxf(m1, plain[k]);
    m2 = fmaxf(m2, marg[k]);
  }
  float z[K], ez[K];
  float s = 0.f, sum_m = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    z[k] = (plain[k] - m1) * (1.0f - M[k]) + (marg[k] - m2) * M[k];
    ez[k] = expf(z[k]);
    s += ez[k];
    sum_m += M[k];
  }
  const float den = s + 1e-4f;
  const float lse = logf(den);
  const float inv_den = 1.0f / den;
  float row = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    row += M[k] * (z[k] - lse);
    const float delta = (marg[k] - m2) - (plain[k] - m1);
    dM[k] = -mc.scale * ((z[k] - lse) + delta * (M[k] - ez[k] * inv_den * sum_m));
  }
  return -mc.scale * row;
}


// host side: the constants of one MPCL instance (utils/loss.py:475-480)
inline MarginConst make_margin_const(const slcl_proto_params_t* p) {
  MarginConst mc;
  mc.inv_t = (float)(1.0 / (double)p->temperature);
  mc.scale = p->temperature / p->base_temperature;
  mc.cos_m = (float)cos((double)p->margin);
  mc.sin_m = (float)sin((double)p->margin);
  mc.th = (float)cos(M_PI - (double)p->margin);
  mc.mm = (float)(sin(M_PI - (double)p->margin) * (double)p->margin);
  mc.easy = p->easy_margin;
  mc.normalize = p->normalize;
  return mc;
}

}  // namespace slcl
