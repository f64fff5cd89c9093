// Per-node closed-form CRPS value + gradient (fp32), shared by the device kernel and the
// host-compiled unit check (tests/test_crps_math_host.py builds this header for the CPU).
//
// Follows models/loss.py (MixedLoss.crps :203-272 with helpers :81-200, MixedNormalCRPS.crps :12-68,
// NormalCRPS.crps :346-369) and the links of models/model_utils.py:89-113, in the common
// sub-expression form of SURVEY.md Appendix A.  The gradient is the hand-derived reverse sweep; the
// identities dA/dz_c = -P_c^2 and dA/dz_u = P_u^2 collapse the erf/exp terms exactly.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define RC_HD __host__ __device__ __forceinline__
#else
#define RC_HD inline
#endif

namespace rc {

constexpr float kLogC = -4.605170185988091f;      // np.log(0.01), models/loss.py:72
constexpr float kLinkEps = 1e-6f;                  // models/model_utils.py:5
constexpr float kUScale = 2.12f;                   // models/model_utils.py:104
constexpr float kInvSqrtPi = 0.5641895835477563f;
constexpr float kSqrt2 = 1.4142135623730951f;
constexpr float kInvSqrt2 = 0.7071067811865476f;
constexpr float kLogSqrt2Pi = 0.9189385332046727f;

RC_HD float norm_cdf(float z) { return 0.5f * (1.0f + erff(z * kInvSqrt2)); }
RC_HD float norm_pdf(float z) { return expf(-0.5f * z * z - kLogSqrt2Pi); }
RC_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
RC_HD float softplusf_(float x) { return x > 20.0f ? x : log1pf(expf(x)); }          // torch threshold 20
RC_HD float softplus_grad(float x) { return x > 20.0f ? 1.0f : sigmoidf_(x); }

// number of head columns per loss kind
RC_HD int loss_width(int kind) { return kind + 2; }

// links: raw -> post (in place on a 5-float register row), models/model_utils.py:89-113
RC_HD void apply_links(float* v, int kind) {
  v[1] = softplusf_(v[1]) + kLinkEps;
  if (kind >= 1) v[2] = sigmoidf_(v[2]);
  if (kind >= 2) v[3] = softplusf_(v[3]) + kLinkEps;
  if (kind >= 3) v[4] = sigmoidf_(v[4]) * kUScale;
}
// d post -> d raw given the RAW values r
RC_HD void links_backward(const float* r, float* g, int kind) {
  g[1] *= softplus_grad(r[1]);
  if (kind >= 1) { float s = sigmoidf_(r[2]); g[2] *= s * (1.0f - s); }
  if (kind >= 2) g[3] *= softplus_grad(r[3]);
  if (kind >= 3) { float s = sigmoidf_(r[4]); g[4] *= kUScale * s * (1.0f - s); }
}

// NormalCRPS: v = (mu, sigma).  Returns the node loss, fills g[0..1].
RC_HD float crps_normal(const float* v, float y, float* g) {
  const float s = v[1];
  const float z = (y - v[0]) / s;
  const float two_cdf_m1 = 2.0f * norm_cdf(z) - 1.0f;
  const float pdf = norm_pdf(z);
  g[0] = -two_cdf_m1;
  g[1] = 2.0f * pdf - kInvSqrtPi;
  return s * (z * two_cdf_m1 + 2.0f * pdf - kInvSqrtPi);
}

// MixedNormalCRPS: v = (mu, sigma, p).
RC_HD float crps_mixed_normal(const float* v, float y, float* g) {
  const float mu = v[0], s = v[1], p = v[2], q = 1.0f - p;
  const float zy = (y - mu) / s, zc = (kLogC - mu) / s;
  const float Fy = norm_cdf(zy), Fc = norm_cdf(zc), fy = norm_pdf(zy), fc = norm_pdf(zc);
  const float F2c = norm_cdf(kSqrt2 * zc);
  const float Pc = p + q * Fc;
  const float cy = 2.0f * (p + q * Fy) - 1.0f;
  const float A = -zc * Pc * Pc - 2.0f * q * fc * Pc - q * q * kInvSqrtPi * (1.0f - F2c);
  const float Bq = zy * cy + 2.0f * q * fy + A;
  const float dAdp = (-2.0f * zc * Pc - 2.0f * q * fc) * (1.0f - Fc) + 2.0f * fc * Pc
                     + 2.0f * q * kInvSqrtPi * (1.0f - F2c);
  const float dzy = s * cy, dzc = -s * Pc * Pc;
  g[0] = -(dzy + dzc) / s;
  g[1] = Bq - (dzy * zy + dzc * zc) / s;
  g[2] = s * (2.0f * zy * (1.0f - Fy) - 2.0f * fy + dAdp);
  return s * Bq;
}

// MixedLoss: v = (mu, sigma, p, sigma_u, u); learn_u selects the sigmoid blend (models/loss.py:266)
// against the hard switch (:268).  g[4] (d/du) is only meaningful when learn_u.
RC_HD float crps_mixed(const float* v, float y, float xi, float t, bool learn_u, float* g) {
  const float mu = v[0], s = v[1], p = v[2], su = v[3], u = v[4], q = 1.0f - p;
  const float inv_s = 1.0f / s, inv_su = 1.0f / su;     // one division per scale (the reference divides each time: <= 1 ulp apart)
  const float zc = (kLogC - mu) * inv_s, zu = (u - mu) * inv_s, zy = (y - mu) * inv_s;
  const float Fc = norm_cdf(zc), Fu = norm_cdf(zu), Fy = norm_cdf(zy);
  const float fc = norm_pdf(zc), fu = norm_pdf(zu), fy = norm_pdf(zy);
  const float dF2 = norm_cdf(kSqrt2 * zu) - norm_cdf(kSqrt2 * zc);
  const float Pc = p + q * Fc;
  const float Pu = q * (1.0f - Fu);
  const float Pm = 1.0f - (p + q * Fu);                     // 1 - m_u, models/loss.py:107,122
  const float A = -zc * Pc * Pc + zu * Pu * Pu - 2.0f * q * fc * Pc - 2.0f * q * fu * Pu
                  - q * q * kInvSqrtPi * dF2;
  const float cy = 2.0f * (p + q * Fy) - 1.0f;
  const float Bq = zy * cy + 2.0f * q * fy + A;                      // body / sigma
  const float Uq = zu + 2.0f * q * fu - 2.0f * zu * Pu + A;          // upper / sigma
  const float inv2mx = 1.0f / (2.0f - xi), inv1mx = 1.0f / (1.0f - xi);
  const float tail_u_q = Pm * Pm * inv2mx;                           // tail(u) / sigma_u
  const float x = (y - u) * inv_su;
  float Tq, S = 0.0f, one_m_T = 0.0f;
  if (x > 0.0f) {
    const float base = 1.0f + xi * x;
    if (xi == 0.5f) {                                                // every shipped params.json: S = base^-2, S^(1-xi) = base^-1
      const float rb = 1.0f / base;
      S = rb * rb;
      one_m_T = 1.0f - rb;
    } else {
      S = powf(base, -1.0f / xi);                                    // GPD survival, models/loss.py:90
      one_m_T = 1.0f - powf(base, -(1.0f - xi) / xi);                // 1 - S^(1-xi)
    }
    Tq = x - 2.0f * Pm * inv1mx * one_m_T + tail_u_q;
  } else {
    Tq = fabsf(x) + tail_u_q;
  }
  const float L1 = s * Bq + su * tail_u_q;
  const float L2 = su * Tq + s * Uq;
  float w, du;
  if (learn_u) {
    w = sigmoidf_((u - y) * t);
    du = (L1 - L2) * w * (1.0f - w) * t;
  } else {
    w = (y < u) ? 1.0f : 0.0f;
    du = 0.0f;
  }
  const float loss = w * (L1 - L2) + L2;
  // ---- reverse sweep (upstream gradient 1)
  const float dL1 = w, dL2 = 1.0f - w;
  float ds = dL1 * Bq + dL2 * Uq;
  float dsu = dL1 * tail_u_q + dL2 * Tq;
  const float dBq = dL1 * s, dUq = dL2 * s, dA = dBq + dUq;
  const float dPm = su * (2.0f * Pm * inv2mx - dL2 * 2.0f * inv1mx * one_m_T);
  const float dTq = dL2 * su;
  // d tail(y) / d x:  1 - 2*Pm*S for x > 0 (T' = -(1-xi) S),  sign(x) otherwise (|x|' = 0 at 0)
  const float dx = dTq * (x > 0.0f ? 1.0f - 2.0f * Pm * S : (x < 0.0f ? -1.0f : 0.0f));
  du += -dx * inv_su;
  dsu += -dx * x * inv_su;
  const float dzy = dBq * cy;
  const float dzc = -dA * Pc * Pc;
  const float dzu = dA * Pu * Pu + dUq * (1.0f - 2.0f * Pu) - dPm * q * fu;
  const float dAdp = (-2.0f * zc * Pc - 2.0f * q * fc) * (1.0f - Fc) - (2.0f * zu * Pu - 2.0f * q * fu) * (1.0f - Fu)
                     + 2.0f * fc * Pc + 2.0f * fu * Pu + 2.0f * q * kInvSqrtPi * dF2;
  g[2] = dBq * (2.0f * zy * (1.0f - Fy) - 2.0f * fy) + dA * dAdp
         + dUq * (-2.0f * fu + 2.0f * zu * (1.0f - Fu)) - dPm * (1.0f - Fu);
  g[0] = -(dzy + dzc + dzu) * inv_s;
  g[1] = ds - (dzy * zy + dzc * zc + dzu * zu) * inv_s;
  g[3] = dsu;
  g[4] = du + dzu * inv_s;
  return loss;
}

// One node: `row` holds post-processed (raw_input=0) or raw (raw_input=1) head outputs.
// Writes the gradient w.r.t. `row` into g[0..width) (un-normalised) and returns the node loss.
// In raw mode every link is evaluated once: value and derivative share the exponential.
template <int kind>
RC_HD float crps_node_k(const float* row, float y, int raw_input, float u_fixed, float xi, float t, float* g) {
  float v[5] = {0.f, 1.f, 0.f, 1.f, 0.f};
  float dlink[5] = {1.f, 1.f, 1.f, 1.f, 1.f};
  constexpr int width = kind + 2;
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < width; ++i) v[i] = row[i];
  if (raw_input) {
    {                                                     // sigma = softplus(r1) + 1e-6, d/dr = e/(1+e) (torch: threshold 20)
      const float e = expf(v[1]);
      dlink[1] = v[1] > 20.0f ? 1.0f : e / (1.0f + e);
      v[1] = (v[1] > 20.0f ? v[1] : log1pf(e)) + kLinkEps;
    }
    if (kind >= 1) { const float sg = sigmoidf_(v[2]); v[2] = sg; dlink[2] = sg * (1.0f - sg); }
    if (kind >= 2) {
      const float e = expf(v[3]);
      dlink[3] = v[3] > 20.0f ? 1.0f : e / (1.0f + e);
      v[3] = (v[3] > 20.0f ? v[3] : log1pf(e)) + kLinkEps;
    }
    if (kind >= 3) { const float sg = sigmoidf_(v[4]); v[4] = sg * kUScale; dlink[4] = kUScale * sg * (1.0f - sg); }
  }
  float loss;
  if (kind == RC_LOSS_NORMAL) {
    loss = crps_normal(v, y, g);
  } else if (kind == RC_LOSS_MIXED_NORMAL) {
    loss = crps_mixed_normal(v, y, g);
  } else {
    const bool learn_u = (kind == RC_LOSS_MIXED_U);
    if (!learn_u) v[4] = u_fixed;
    loss = crps_mixed(v, y, xi, t, learn_u, g);
  }
  if (raw_input) {
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int i = 1; i < width; ++i) g[i] *= dlink[i];
  }
  return loss;
}

RC_HD float crps_node(const float* row, float y, int kind, int raw_input, float u_fixed, float xi, float t, float* g) {
  switch (kind) {
    case RC_LOSS_NORMAL: return crps_node_k<RC_LOSS_NORMAL>(row, y, raw_input, u_fixed, xi, t, g);
    case RC_LOSS_MIXED_NORMAL: return crps_node_k<RC_LOSS_MIXED_NORMAL>(row, y, raw_input, u_fixed, xi, t, g);
    case RC_LOSS_MIXED: return crps_node_k<RC_LOSS_MIXED>(row, y, raw_input, u_fixed, xi, t, g);
    default: return crps_node_k<RC_LOSS_MIXED_U>(row, y, raw_input, u_fixed, xi, t, g);
  }
}

}  // namespace rc
