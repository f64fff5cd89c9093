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

// ---------------------------------------------------------------------------------------------------------------
// Fast transcendental building blocks.  The node formula needs, per node, the normal cdf at five points, the pdf at
// three, two softplus and three sigmoid values; with libm-grade erff / expf / log1pf / IEEE divisions that is ~1100
// instructions per node and the kernel is issue bound at 15 % of HBM bandwidth.  Here every value comes from ONE
// exponential per argument:
//   * upper tail  Q(|z|) = 0.5 erfc(|z|/sqrt 2) = exp(-z^2/2) * G(t),  t = 1 / (1 + p |z|),  G a degree-8 polynomial
//     (fit to erfc(x) exp(x^2), absolute error of erf <= 2.2e-7 in fp32 arithmetic - the size of erff's own 2 ulp);
//     the pdf is the same exponential, and Phi(sqrt 2 z) needs exp(-z^2) = its square;
//   * softplus(x) = max(x, 0) + log1p(u), u = exp(-|x|), log1p(u) = 2 atanh(u / (2 + u)) as a 7-term series
//     (relative error < 2e-8 for every u in (0, 1], so sigma = softplus + 1e-6 keeps its relative accuracy when x << 0),
//     and sigmoid(x) = 1/(1+u) or u/(1+u) from the same u; sigmoid'(x) = u / (1+u)^2 without cancellation;
//   * reciprocals and exponentials use the approximate hardware instructions on the device (MUFU.RCP / MUFU.EX2,
//     <= 2 ulp); the host build (tests/test_crps_math_host.py) runs the same formulas with libm.
// torch's softplus threshold (x > 20 -> x) needs no branch: there log1p(u) < 2.1e-9 vanishes against x in fp32.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float rc_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rc_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#define RC_EXP(x) rc_ex2((x) * 1.4426950408889634f)
#define RC_RCP(x) rc_rcp(x)
#else
#define RC_EXP(x) expf(x)
#define RC_RCP(x) (1.0f / (x))
#endif

constexpr float kErfP = 0.85f;                       // t = 1 / (1 + kErfP * x) for erfc(x)
constexpr float kInvSqrt2Pi = 0.3989422804014327f;

// erfc(x) * exp(x^2) for x >= 0
RC_HD float erfcx_poly(float x) {
  const float t = RC_RCP(fmaf(kErfP, x, 1.0f));
  float g = 0.0915326178073883f;
  g = fmaf(g, t, -0.4815599322319031f);
  g = fmaf(g, t, 0.9648935198783875f);
  g = fmaf(g, t, -0.7351084351539612f);
  g = fmaf(g, t, -0.223374143242836f);
  g = fmaf(g, t, 0.46228551864624023f);
  g = fmaf(g, t, 0.43765759468078613f);
  g = fmaf(g, t, 0.48367324471473694f);
  return g * t;
}

struct NormZ {
  float pdf;    // phi(z)
  float cdf;    // Phi(z)
  float sf;     // 1 - Phi(z), computed without cancellation for z > 0
  float e;      // exp(-z^2 / 2)
};
RC_HD NormZ norm_z(float z) {
  NormZ r;
  const float a = fabsf(z);
  r.e = RC_EXP(-0.5f * z * z);
  const float tail = 0.5f * r.e * erfcx_poly(a * kInvSqrt2);           // Q(|z|)
  r.pdf = r.e * kInvSqrt2Pi;
  r.cdf = z >= 0.0f ? 1.0f - tail : tail;
  r.sf = z >= 0.0f ? tail : 1.0f - tail;
  return r;
}
// Phi(sqrt(2) z) given e = exp(-z^2 / 2)
RC_HD float norm_cdf_sqrt2(float z, float e) {
  const float tail = 0.5f * e * e * erfcx_poly(fabsf(z));
  return z >= 0.0f ? 1.0f - tail : tail;
}
RC_HD float norm_cdf(float z) { return norm_z(z).cdf; }
RC_HD float norm_pdf(float z) { return RC_EXP(-0.5f * z * z) * kInvSqrt2Pi; }

// softplus(x) and its derivative sigmoid(x); sigmoid(x) and its derivative
RC_HD void softplus_and_grad(float x, float& sp, float& dsp) {
  const float u = RC_EXP(-fabsf(x));
  const float s = u * RC_RCP(2.0f + u), s2 = s * s;
  float q = 1.0f / 13.0f;
  q = fmaf(q, s2, 1.0f / 11.0f);
  q = fmaf(q, s2, 1.0f / 9.0f);
  q = fmaf(q, s2, 1.0f / 7.0f);
  q = fmaf(q, s2, 1.0f / 5.0f);
  q = fmaf(q, s2, 1.0f / 3.0f);
  q = fmaf(q, s2, 1.0f);
  sp = fmaxf(x, 0.0f) + 2.0f * s * q;
  const float r = RC_RCP(1.0f + u);
  dsp = x >= 0.0f ? r : u * r;
}
RC_HD void sigmoid_and_grad(float x, float& sg, float& dsg) {
  const float u = RC_EXP(-fabsf(x));
  const float r = RC_RCP(1.0f + u);
  sg = x >= 0.0f ? r : u * r;
  dsg = u * r * r;
}
RC_HD float sigmoidf_(float x) { float s, d; sigmoid_and_grad(x, s, d); return s; }
RC_HD float softplusf_(float x) { float s, d; softplus_and_grad(x, s, d); return s; }
RC_HD float softplus_grad(float x) { float s, d; softplus_and_grad(x, s, d); return d; }

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
  if (kind >= 1) { float s, d; sigmoid_and_grad(r[2], s, d); g[2] *= d; }
  if (kind >= 2) g[3] *= softplus_grad(r[3]);
  if (kind >= 3) { float s, d; sigmoid_and_grad(r[4], s, d); g[4] *= kUScale * d; }
}

// NormalCRPS: v = (mu, sigma).  Returns the node loss, fills g[0..1].
RC_HD float crps_normal(const float* v, float y, float* g) {
  const float s = v[1];
  const float z = (y - v[0]) * RC_RCP(s);
  const NormZ n = norm_z(z);
  const float two_cdf_m1 = n.cdf - n.sf;
  g[0] = -two_cdf_m1;
  g[1] = 2.0f * n.pdf - kInvSqrtPi;
  return s * (z * two_cdf_m1 + 2.0f * n.pdf - kInvSqrtPi);
}

// MixedNormalCRPS: v = (mu, sigma, p).
RC_HD float crps_mixed_normal(const float* v, float y, float* g) {
  const float mu = v[0], s = v[1], p = v[2], q = 1.0f - p;
  const float inv_s = RC_RCP(s);
  const float zy = (y - mu) * inv_s, zc = (kLogC - mu) * inv_s;
  const NormZ ny = norm_z(zy), nc = norm_z(zc);
  const float Fy = ny.cdf, Fc = nc.cdf, fy = ny.pdf, fc = nc.pdf;
  const float S2c = 1.0f - norm_cdf_sqrt2(zc, nc.e);
  const float Pc = p + q * Fc;
  const float cy = 2.0f * (p + q * Fy) - 1.0f;
  const float A = -zc * Pc * Pc - 2.0f * q * fc * Pc - q * q * kInvSqrtPi * S2c;
  const float Bq = zy * cy + 2.0f * q * fy + A;
  const float dAdp = (-2.0f * zc * Pc - 2.0f * q * fc) * nc.sf + 2.0f * fc * Pc + 2.0f * q * kInvSqrtPi * S2c;
  const float dzy = s * cy, dzc = -s * Pc * Pc;
  g[0] = -(dzy + dzc) * inv_s;
  g[1] = Bq - (dzy * zy + dzc * zc) * inv_s;
  g[2] = s * (2.0f * zy * ny.sf - 2.0f * fy + dAdp);
  return s * Bq;
}

// MixedLoss: v = (mu, sigma, p, sigma_u, u); learn_u selects the sigmoid blend (models/loss.py:266)
// against the hard switch (:268).  g[4] (d/du) is only meaningful when learn_u.
RC_HD float crps_mixed(const float* v, float y, float xi, float t, bool learn_u, float* g) {
  const float mu = v[0], s = v[1], p = v[2], su = v[3], u = v[4], q = 1.0f - p;
  const float inv_s = RC_RCP(s), inv_su = RC_RCP(su);   // one reciprocal per scale (the reference divides each time)
  const float zc = (kLogC - mu) * inv_s, zu = (u - mu) * inv_s, zy = (y - mu) * inv_s;
  const NormZ nc = norm_z(zc), nu = norm_z(zu), ny = norm_z(zy);
  const float Fc = nc.cdf, Fy = ny.cdf, Su = nu.sf;
  const float fc = nc.pdf, fu = nu.pdf, fy = ny.pdf;
  const float dF2 = norm_cdf_sqrt2(zu, nu.e) - norm_cdf_sqrt2(zc, nc.e);
  const float Pc = p + q * Fc;
  const float Pu = q * Su;
  const float Pm = Pu;                                      // 1 - m_u = 1 - (p + q Fu) = q (1 - Fu), models/loss.py:107,122
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
      const float rb = RC_RCP(base);
      S = rb * rb;
      one_m_T = xi * x * rb;                                         // 1 - 1/base without cancellation
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
    float dw;
    sigmoid_and_grad((u - y) * t, w, dw);
    du = (L1 - L2) * dw * t;
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
  const float dAdp = (-2.0f * zc * Pc - 2.0f * q * fc) * nc.sf - (2.0f * zu * Pu - 2.0f * q * fu) * Su
                     + 2.0f * fc * Pc + 2.0f * fu * Pu + 2.0f * q * kInvSqrtPi * dF2;
  g[2] = dBq * (2.0f * zy * ny.sf - 2.0f * fy) + dA * dAdp
         + dUq * (-2.0f * fu + 2.0f * zu * Su) - dPm * Su;
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
  if (raw_input) {                                        // every link once: value and derivative share the exponential
    { float sp; softplus_and_grad(v[1], sp, dlink[1]); v[1] = sp + kLinkEps; }            // sigma = softplus(r1) + 1e-6
    if (kind >= 1) { float sg; sigmoid_and_grad(v[2], sg, dlink[2]); v[2] = sg; }
    if (kind >= 2) { float sp; softplus_and_grad(v[3], sp, dlink[3]); v[3] = sp + kLinkEps; }
    if (kind >= 3) { float sg, d; sigmoid_and_grad(v[4], sg, d); v[4] = sg * kUScale; dlink[4] = kUScale * d; }
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
