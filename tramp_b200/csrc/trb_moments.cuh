// Elementwise exponential-family moments in FP64 (device functions).
//
// Each routine restates one reference formula (paths relative to
// /root/reference/tramp) with the same operand order and the same branch
// thresholds, so results agree with numpy/scipy to a few ulp.
#pragma once
#include "trb_common.cuh"

namespace trb {

constexpr double kSqrt2 = 1.4142135623730951;          // np.sqrt(2)
constexpr double kSqrt2OverPi = 0.7978845608028654;    // np.sqrt(2/np.pi)
constexpr double kTwoOverSqrtPi = 1.1283791670955126;  // 2/np.sqrt(np.pi)
constexpr double kSqrtPi = 1.7724538509055159;         // np.sqrt(np.pi)
constexpr double kLn2 = 0.6931471805599453;

// numpy logaddexp (npy_math: npy_logaddexp)
__device__ __forceinline__ double logaddexp(double x1, double x2) {
  if (x1 == x2) return x1 + kLn2;
  const double tmp = x1 - x2;
  if (tmp > 0) return x1 + log1p(exp(-tmp));
  if (tmp <= 0) return x2 + log1p(exp(tmp));
  return tmp;  // NaN
}

// beliefs/normal.py:3-4
__device__ __forceinline__ double normal_A(double a, double b) {
  return 0.5 * (b * b / a + log(kTwoPi / a));
}

// scipy.special.expit (double): 1 / (1 + exp(-x))
__device__ __forceinline__ double expit(double x) { return 1.0 / (1.0 + exp(-x)); }

// utils/truncated_normal.py:21-29
__device__ __forceinline__ double log_Phi(double x) {
  if (!(x < 30.0)) return 0.0;
  return log(0.5 * erfcx(-x / kSqrt2)) - 0.5 * x * x;
}

__device__ __forceinline__ double sign_of(double y) {
  return (y > 0.0) ? 1.0 : ((y < 0.0) ? -1.0 : 0.0);
}

// ---- utils/truncated_normal.py:14-200, finite-interval F0/F1/F2 -------------
struct F012 {
  double f0, f1, f2;
};

__device__ inline F012 trunc_F(double x, double y) {
  const double thresh = 1e-7;
  if (fabs(x) > fabs(y)) {  // `switch`, :14-18
    const double t = x;
    x = y;
    y = t;
  }
  F012 o;
  const double x2 = x * x;
  if (isinf(y)) {  // F*_inf :32-34, 92-94, 147-149
    const double s = sign_of(y);
    const double e = erfcx(s * x);
    o.f0 = log(e) - x2;
    o.f1 = s / e;
    o.f2 = s * x / e;
  } else if (fabs(x - y) <= thresh) {  // F*_close :37-45, 97-105, 152-161
    const double e = y - x;
    const double e2 = e * e, e3 = e2 * e, e4 = e2 * e2, x4 = x2 * x2;
    o.f0 = (-x * e + (1.0 / 6) * (x2 - 2) * e2 - (1.0 / 180) * (x4 + 2 * x2 - 8) +
            log(2 * e / kSqrtPi)) -
           x2;
    o.f1 = kSqrtPi * (x + (1.0 / 2) * e - (1.0 / 6) * e2 - (1.0 / 12) * e3 +
                      (1.0 / 90) * x * (x2 + 1.) * e4);
    o.f2 = kSqrtPi * (x2 - 1.0 / 2 + x * e - (1.0 / 3) * (x2 - 1) * e2 - (1.0 / 3) * x * e3 +
                      (1.0 / 90) * (2 * x4 + 3 * x2 - 8) * e4);
  } else if (x < 0 && y < 0) {  // F*_neg :48-53, 108-110, 164-166
    const double D = exp(x2 - y * y);
    const double den = D * erfcx(-y) - erfcx(-x);
    o.f0 = log(fabs(den)) - x2;
    o.f1 = (1 - D) / den;
    o.f2 = (x - D * y) / den;
  } else if (x > 0 && y > 0) {  // F*_pos :56-61, 113-115, 169-171
    const double D = exp(x2 - y * y);
    const double den = erfcx(x) - D * erfcx(y);
    o.f0 = log(fabs(den)) - x2;
    o.f1 = (1 - D) / den;
    o.f2 = (x - D * y) / den;
  } else {  // F*_other :64-65, 118-120, 174-176
    const double D = exp(x2 - y * y);
    const double den = erf(y) - erf(x);
    o.f0 = log(fabs(den));
    o.f1 = exp(-x2) * (1 - D) / den;
    o.f2 = exp(-x2) * (x - D * y) / den;
  }
  return o;
}

struct TruncMoments {
  double mean, var, logZ, proba;
};

// utils/truncated_normal.py:234-298 (mean, var, log_proba, proba, logZ)
__device__ inline TruncMoments truncated_normal(double r0, double v0, double zmin, double zmax) {
  const double s0 = sqrt(v0);
  const double ymin = (zmin - r0) / s0;
  const double ymax = (zmax - r0) / s0;
  double g0, g1, g2;
  if (zmax == INFINITY) {  // G*_inf(ymin, +1) :218-231
    const double u = ymin / kSqrt2;
    const double e = erfcx(u);
    g0 = log_Phi(-ymin);
    g1 = kSqrt2OverPi * (1.0 / e);
    g2 = kTwoOverSqrtPi * (u / e);
  } else if (zmin == -INFINITY) {  // G*_inf(ymax, -1)
    const double u = ymax / kSqrt2;
    const double e = erfcx(-u);
    g0 = log_Phi(ymax);
    g1 = kSqrt2OverPi * (-1.0 / e);
    g2 = kTwoOverSqrtPi * (-u / e);
  } else {  // G0/G1/G2 :203-215
    const F012 f = trunc_F(ymin / kSqrt2, ymax / kSqrt2);
    g0 = log(0.5) + f.f0;
    g1 = kSqrt2OverPi * f.f1;
    g2 = kTwoOverSqrtPi * f.f2;
  }
  TruncMoments t;
  t.mean = r0 + s0 * g1;
  t.var = v0 * (1. + g2 - g1 * g1);
  t.logZ = 0.5 * log(kTwoPi * v0) + 0.5 * r0 * r0 / v0 + g0;
  const double lo = (zmin == -INFINITY) ? -INFINITY : ymin;
  const double hi = (zmax == INFINITY) ? INFINITY : ymax;
  t.proba = 0.5 * (1 + erf(hi / kSqrt2)) - 0.5 * (1 + erf(lo / kSqrt2));  // utils/misc.py:50-52
  return t;
}

// ---- separable factors ------------------------------------------------------
struct RV {
  double r, v;
};

// beliefs/positive.py:12-17 through the half-infinite path of truncated_normal
__device__ __forceinline__ RV positive_rv(double a, double b) {
  const double r0 = b / a, v0 = 1 / a;
  const double s0 = sqrt(v0);
  const double ymin = (0.0 - r0) / s0;
  const double u = ymin / kSqrt2;
  const double e = erfcx(u);
  const double g1 = kSqrt2OverPi * (1.0 / e);
  const double g2 = kTwoOverSqrtPi * (u / e);
  RV o;
  o.r = r0 + s0 * g1;
  o.v = v0 * (1. + g2 - g1 * g1);
  return o;
}

// compute_forward_posterior / compute_backward_posterior, elementwise part
// (the `.mean()` over components is done by the caller).
__device__ __forceinline__ RV factor_moments(const trb_factor& f, double a, double b, double y) {
  RV o;
  switch (f.kind) {
    case TRB_GAUSS_BERNOULLI_PRIOR: {  // gauss_bernoulli_prior.py:70-74, beliefs/sparse.py:9-22
      const double aa = a + f.p0, bb = b + f.p1;
      const double s = expit(normal_A(aa, bb) - f.p2);
      const double ba = bb / aa;
      o.r = s * ba;
      o.v = s / aa + s * (1 - s) * (ba * ba);
      break;
    }
    case TRB_BINARY_PRIOR: {  // binary_prior.py:57-60, beliefs/binary.py:8-13
      const double t = tanh(b + f.p0);
      o.r = t;
      o.v = 1 - t * t;
      break;
    }
    case TRB_GAUSSIAN_PRIOR: {  // gaussian_prior.py:63-68
      const double aa = a + f.p0, bb = b + f.p1;
      o.r = bb / aa;
      o.v = 1 / aa;
      break;
    }
    case TRB_GAUSSIAN_LIKELIHOOD: {  // gaussian_likelihood.py:43-49
      const double ay = f.p0, by = f.p0 * y;
      const double aa = a + ay, bb = b + by;
      o.r = bb / aa;
      o.v = 1 / aa;
      break;
    }
    case TRB_SGN_LIKELIHOOD: {  // sgn_likelihood.py:32-34
      const RV p = positive_rv(a, b * y);
      o.r = y * p.r;
      o.v = p.v;
      break;
    }
    default: {  // TRB_ABS_LIKELIHOOD, abs_likelihood.py:31-33
      const double t = tanh(b * y);
      o.r = y * t;
      o.v = (y * y) * (1 - t * t);
      break;
    }
  }
  return o;
}

// beliefs/sparse.py:9-12: weight of the Gaussian component, expit(normal.A(a, b) - eta)
__device__ __forceinline__ double sparse_weight(const trb_factor& f, double a, double b) {
  const double aa = a + f.p0, bb = b + f.p1;
  return expit(normal_A(aa, bb) - f.p2);
}

// scalar_log_partition, elementwise (the reference's compute_log_partition is
// its mean over components)
__device__ __forceinline__ double factor_log_partition(const trb_factor& f, double a, double b,
                                                       double y) {
  switch (f.kind) {
    case TRB_GAUSS_BERNOULLI_PRIOR: {  // gauss_bernoulli_prior.py:79-82, sparse.py:5-6
      const double aa = a + f.p0, bb = b + f.p1;
      return logaddexp(f.p2, normal_A(aa, bb)) - f.p3;
    }
    case TRB_BINARY_PRIOR: {  // binary_prior.py:65-67
      const double bb = b + f.p0;
      return logaddexp(bb, -bb) - logaddexp(f.p0, -f.p0) - 0.5 * a;
    }
    case TRB_GAUSSIAN_PRIOR:  // gaussian_prior.py:70-73
      return normal_A(a + f.p0, b + f.p1) - normal_A(f.p0, f.p1);
    case TRB_GAUSSIAN_LIKELIHOOD: {  // gaussian_likelihood.py:51-55
      const double ay = f.p0, by = f.p0 * y;
      return normal_A(a + ay, b + by) - normal_A(ay, by);
    }
    case TRB_SGN_LIKELIHOOD: {  // sgn_likelihood.py:39-40 -> positive.A -> truncated_normal_logZ
      const double bb = b * y;
      const double r0 = bb / a, v0 = 1 / a;
      const double s0 = sqrt(v0);
      const double ymin = (0.0 - r0) / s0;
      return 0.5 * log(kTwoPi * v0) + 0.5 * r0 * r0 / v0 + log_Phi(-ymin);
    }
    default: {  // abs_likelihood.py:38-39
      const double bb = b * y;
      return -0.5 * a * (y * y) + logaddexp(bb, -bb);
    }
  }
}

__device__ __forceinline__ bool factor_is_constant_message(int kind) {
  return kind == TRB_GAUSSIAN_PRIOR || kind == TRB_GAUSSIAN_LIKELIHOOD;
}

}  // namespace trb
