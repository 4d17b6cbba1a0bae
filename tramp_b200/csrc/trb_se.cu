// State Evolution (SE) on the device: the scalar twin of the EP sweep.
//
// reference: tramp/algos/state_evolution.py:5-27 drives the schedule of
// algos/message_passing.py:249-269 on the precisions `a` alone; each factor's
// update averages its posterior variance over the law of its incoming beliefs
// (priors/base_prior.py:66-74, likelihoods/base_likelihood.py:73-81), a Gaussian
// integral the reference hands to scipy.integrate.quad / dblquad on [-10, 10]
// (utils/integration.py:13-46), i.e. hundreds of Python callbacks per factor
// update.  Here an integral is a composite Gauss-Legendre sum in a sinh-mapped
// variable (nodes crowd around the point where the integrand has its structure,
// see trb_quadrature), evaluated by the 256 threads of one CTA with the moment
// routines of trb_moments.cuh; one CTA owns
// one SE problem and runs ALL its iterations (forward pass, backward pass,
// variable update, damping, NaN check, EarlyStopping) without leaving the SM,
// so a whole grid of problems -- a phase diagram -- is one launch.
#include "trb_moments.cuh"

using namespace trb;

namespace {

constexpr int kSeThreads = 384;

constexpr double kLimit = 10.0;                  // utils/integration.py:27, 45
constexpr double kInvSqrt2Pi = 0.3989422804014327;  // utils/misc.py:46-47 (norm_pdf)

// One dimension of the rule: composite Gauss-Legendre in u, t = c + kappa sinh(u).
struct Rule {
  const double* __restrict__ x;
  const double* __restrict__ w;
  int Q, P;
  double kappa;
};
struct Quad {
  Rule r1, r2;
};

__host__ __device__ inline Quad make_quad(const trb_quadrature& q) {
  Quad o;
  o.r1.x = q.x;
  o.r1.w = q.w;
  o.r1.Q = q.Q;
  o.r1.P = q.P;
  o.r1.kappa = q.kappa;
  o.r2.x = q.x2;
  o.r2.w = q.w2;
  o.r2.Q = q.Q2;
  o.r2.P = q.P2;
  o.r2.kappa = q.kappa2;
  return o;
}

// The u-interval of one integral, for a centre c (clamped into [-10, 10]).
struct Map {
  double c, u_lo, du, kappa;
};
__device__ __forceinline__ Map make_map(const Rule& r, double c) {
  Map m;
  m.c = (c == c) ? fmin(fmax(c, -kLimit), kLimit) : 0.0;
  m.kappa = r.kappa;
  m.u_lo = asinh((-kLimit - m.c) / r.kappa);
  m.du = (asinh((kLimit - m.c) / r.kappa) - m.u_lo) / r.P;
  return m;
}
// node `idx` (= panel * Q + q) of the mapped rule: abscissa t, weight including
// the Jacobian kappa cosh(u) and the standard normal density at t
__device__ __forceinline__ void map_node(const Rule& r, const Map& m, int idx, double* t, double* wt) {
  const int p = idx / r.Q, q = idx - p * r.Q;
  const double u = m.u_lo + (p + 0.5 * (r.x[q] + 1.0)) * m.du;
  const double e = exp(u), ei = 1.0 / e;
  const double tt = m.c + m.kappa * (0.5 * (e - ei));
  *t = tt;
  *wt = (0.5 * m.du * r.w[q]) * (m.kappa * (0.5 * (e + ei))) * (kInvSqrt2Pi * exp(-0.5 * tt * tt));
}

// integral over [-10, 10] of N(t) f(mean + s t) dt: utils/integration.py:13-28
// (gaussian_measure) with the quad() call replaced by the mapped rule centred at
// t = c.  Block-wide; result in every thread.
template <class F>
__device__ __forceinline__ double measure_1d(double mean, double s, double c, F f, const Quad& q,
                                             double* sh) {
  const Map m = make_map(q.r1, c);
  const int K = q.r1.P * q.r1.Q;
  double acc = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double t, wt;
    map_node(q.r1, m, k, &t, &wt);
    acc += wt * f(mean + s * t);
  }
  return block_sum(acc, sh);
}

// integral of N(x1) N(x2) f(s1 x1, x2): utils/integration.py:31-46 (gaussian_measure_2d
// with m1 = m2 = 0, s2 = 1).  Outer variable centred at 0, inner at c2(x1).
template <class F, class C2>
__device__ __forceinline__ double measure_2d(double s1, F f, C2 c2, const Quad& q, double* sh) {
  const Map m1 = make_map(q.r2, 0.0);
  const int K = q.r2.P * q.r2.Q;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  double acc = 0.0;
  // one outer node per warp at a time, the inner rule spread over its lanes
  for (int i = warp; i < K; i += nwarp) {
    double t1, w1;
    map_node(q.r2, m1, i, &t1, &w1);
    const double z = s1 * t1;
    const Map m2 = make_map(q.r2, c2(z));
    double inner = 0.0;
    for (int j = lane; j < K; j += 32) {
      double t2, w2;
      map_node(q.r2, m2, j, &t2, &w2);
      inner += w2 * f(z, t2);
    }
    acc += w1 * inner;
  }
  return block_sum(acc, sh);
}

// scalar_forward_variance / scalar_log_partition of a prior, scalar_backward_variance /
// compute_log_partition of a likelihood
__device__ __forceinline__ double factor_scalar(const trb_factor& f, int what, double a, double b,
                                                double y) {
  return what == TRB_MEASURE_V ? factor_moments(f, a, b, y).v : factor_log_partition(f, a, b, y);
}

// beliefs/positive.py:24-26 -> utils/truncated_normal.py:284-291 with zmin = 0, zmax = inf
__device__ __forceinline__ double positive_p(double a, double b) {
  const double r0 = b / a, v0 = 1 / a;
  const double s0 = sqrt(v0);
  const double ymin = (0.0 - r0) / s0;
  return 0.5 * (1 + 1.0) - 0.5 * (1 + erf(ymin / kSqrt2));  // utils/misc.py:50-52 (norm_cdf)
}

// Prior.beliefs_measure(ax, f).  tau is used by the Gaussian prior's free energy only.
__device__ double prior_measure(const trb_factor& f, int what, double a, double tau, const Quad& q,
                                double* sh) {
  switch (f.kind) {
    case TRB_GAUSS_BERNOULLI_PRIOR: {  // gauss_bernoulli_prior.py:112-118
      const double var = 1 / f.p0, mean = f.p1 / f.p0;
      // rho from eta = normal.A(a0, b0) - log(rho / (1 - rho)) (gauss_bernoulli_prior.py:36)
      const double rho = expit(normal_A(f.p0, f.p1) - f.p2);
      auto g = [&](double b) { return factor_scalar(f, what, a, b, 0.0); };
      // the moments are functions of b + b0: centre the rule where b = -b0
      const double s0 = sqrt(a), m1 = a * mean, s1 = sqrt(a + (a * a) * var);
      const double mu_0 = measure_1d(0.0, s0, s0 > 0 ? -f.p1 / s0 : 0.0, g, q, sh);
      const double mu_1 = measure_1d(m1, s1, s1 > 0 ? (-f.p1 - m1) / s1 : 0.0, g, q, sh);
      return (1 - rho) * mu_0 + rho * mu_1;
    }
    case TRB_BINARY_PRIOR: {  // binary_prior.py:80-84
      const double p_pos = expit(2 * f.p0), p_neg = 1 - p_pos;
      auto g = [&](double b) { return factor_scalar(f, what, a, b, 0.0); };
      const double s0 = sqrt(a);  // tanh(b + b0): centre where b = -b0 (clamped into the domain)
      const double mu_pos = measure_1d(+a, s0, s0 > 0 ? (-f.p0 - a) / s0 : 0.0, g, q, sh);
      const double mu_neg = measure_1d(-a, s0, s0 > 0 ? (-f.p0 + a) / s0 : 0.0, g, q, sh);
      return p_pos * mu_pos + p_neg * mu_neg;
    }
    default: {  // TRB_GAUSSIAN_PRIOR: closed forms, gaussian_prior.py:92-95, 129-138
      const double aa = a + f.p0;
      if (what == TRB_MEASURE_V) return 1 / aa;
      const double var = 1 / f.p0;
      const double I = 0.5 * log(aa * var);
      return 0.5 * a * tau - I;
    }
  }
}

// Likelihood.beliefs_measure(az, tau_z, f).  *domain is set when mz_hat <= 0.
__device__ double lik_measure(const trb_factor& f, int what, double a, double tau, const Quad& q,
                              double* sh, int* domain) {
  if (f.kind == TRB_GAUSSIAN_LIKELIHOOD) {  // gaussian_likelihood.py:57-60, 129-132
    const double aa = a + f.p0, var = 1 / f.p0;
    if (what == TRB_MEASURE_V) return 1 / aa;
    return 0.5 * a * tau - 1 - 0.5 * log(aa * var);
  }
  const double mz_hat = a - 1 / tau;
  if (!(mz_hat > 0)) {  // sgn_likelihood.py:81, abs_likelihood.py:58
    *domain = 1;
    return NAN;
  }
  if (f.kind == TRB_SGN_LIKELIHOOD) {  // sgn_likelihood.py:79-92
    const double sz_eff = sqrt(mz_hat + (mz_hat * mz_hat) * tau);
    auto f_pos = [&](double bz) { return positive_p(a, +bz) * factor_scalar(f, what, a, bz, +1.0); };
    auto f_neg = [&](double bz) { return positive_p(a, -bz) * factor_scalar(f, what, a, bz, -1.0); };
    const double mu_pos = measure_1d(0.0, sz_eff, 0.0, f_pos, q, sh);
    const double mu_neg = measure_1d(0.0, sz_eff, 0.0, f_neg, q, sh);
    return mu_pos + mu_neg;
  }
  // TRB_ABS_LIKELIHOOD, abs_likelihood.py:56-65
  const double sq = sqrt(mz_hat);
  auto g = [&](double z, double xi_b) {
    const double bz = mz_hat * z + sq * xi_b;
    return factor_scalar(f, what, a, bz, fabs(z));
  };
  auto centre = [&](double z) { return -sq * z; };  // b_z = 0 on the line xi_b = -sqrt(mz_hat) z
  return measure_2d(sqrt(tau), g, centre, q, sh);
}

__global__ void __launch_bounds__(kSeThreads)
k_se_measure(const trb_factor* __restrict__ factors, int factor_stride, int what,
             const double* __restrict__ a, const double* __restrict__ tau, Quad q,
             double* __restrict__ out, int* __restrict__ flags) {
  __shared__ double sh[33];
  const int b = blockIdx.x;
  const trb_factor f = factors[(size_t)b * factor_stride];
  const double tb = tau ? tau[b] : 0.0;
  int domain = 0;
  const double mu = (f.kind <= TRB_GAUSSIAN_PRIOR) ? prior_measure(f, what, a[b], tb, q, sh)
                                                   : lik_measure(f, what, a[b], tb, q, sh, &domain);
  if (threadIdx.x == 0) {
    out[b] = mu;
    if (flags && domain) atomicOr(&flags[b], TRB_FLAG_SE_DOMAIN);
  }
}

// ---- the linear channel ------------------------------------------------------
struct ChannelSE {
  int kind;
  double alpha, mean_spectrum;            // Marchenko-Pastur
  const double* __restrict__ s2;          // empirical spectrum
  int R, Nz, Nx, rank;
};

// compute_n_eff: analytical_linear_channel.py:25-36 / linear_channel.py:58-67
__device__ double channel_n_eff(const ChannelSE& c, double az, double ax, double* sh) {
  if (ax == 0) return 0.;
  const double ratio = az / ax;
  if (c.kind == TRB_SE_MARCHENKO_PASTUR) {
    if (ratio == 0) return fmin(1.0, c.alpha);
    const double gamma = ax / az;
    // marchenko_pastur_ensemble.py:9-11, 40-46
    const double sa = sqrt(c.alpha);
    const double z_max = (1 + sa) * (1 + sa), z_min = (1 - sa) * (1 - sa);
    const double d = sqrt(gamma * z_max + 1) - sqrt(gamma * z_min + 1);
    const double F = d * d;
    const double eta = 1 - F / (4 * gamma);
    return 1 - eta;
  }
  if (ratio == 0) return (double)c.rank / c.Nz;
  double part = 0.0;
  for (int i = threadIdx.x; i < c.rank; i += blockDim.x) part += c.s2[i] / (ratio + c.s2[i]);
  return block_sum(part, sh) / c.Nz;
}

// compute_forward_error: analytical_linear_channel.py:46-51 / linear_channel.py:99-105, 123-125
__device__ double channel_forward_error(const ChannelSE& c, double az, double ax, double* sh) {
  if (c.kind == TRB_SE_MARCHENKO_PASTUR) {
    if (ax == 0) return c.mean_spectrum / (c.alpha * az);
    return channel_n_eff(c, az, ax, sh) / (c.alpha * ax);
  }
  if (ax == 0) {
    double part = 0.0;
    for (int i = threadIdx.x; i < c.rank; i += blockDim.x) part += c.s2[i];
    const double s_mean = block_sum(part, sh) / c.rank;
    return s_mean * c.rank / (c.Nx * az);
  }
  const double alpha = (double)c.Nx / c.Nz;
  return channel_n_eff(c, az, ax, sh) / (alpha * ax);
}

// compute_backward_error: analytical_linear_channel.py:38-44 / linear_channel.py:91-97, 119-121
__device__ double channel_backward_error(const ChannelSE& c, double az, double ax, double* sh) {
  const double az_v = (az != az) ? az : fmax(1e-11, az);
  const double n_eff = channel_n_eff(c, az_v, ax, sh);
  return (1 - n_eff) / az_v;
}

// ---- the whole recursion -----------------------------------------------------
__global__ void __launch_bounds__(kSeThreads) k_se_run(trb_se se, int it0, int n_iter) {
  __shared__ double sh[33];
  const int g = blockIdx.x;
  const int G = se.G;
  if (se.active && !se.active[g]) return;
  const trb_factor prior = se.prior[g];
  const trb_factor lik = se.lik[g];
  const Quad q = make_quad(se.quad);
  const double tau_z = se.tau_z[g];
  ChannelSE ch;
  ch.kind = se.channel;
  ch.alpha = se.alpha ? se.alpha[g] : 0.0;
  ch.mean_spectrum = se.mean_spectrum ? se.mean_spectrum[g] : 0.0;
  ch.s2 = se.s2 ? se.s2 + (size_t)g * se.stride_s2 : nullptr;
  ch.R = se.R;
  ch.Nz = se.Nz;
  ch.Nx = se.Nx;
  ch.rank = se.rank;

  double a[8], old_a[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = se.edge_a[(size_t)e * G + g];
  double vx = se.vx[g], vz = se.vz[g];
  double old_vx = vx, old_vz = vz;  // `old_message_dag`: the state after the previous iteration
  bool have_old_vs = false;         // EarlyStopping.old_vs (callbacks.py:207-208: reset at i == 0)
  int flag = 0, done = 0, stop = 0;

  for (int it = 0; it < n_iter && !stop; ++it) {
#pragma unroll
    for (int e = 0; e < 8; ++e) old_a[e] = a[e];
    old_vx = vx;
    old_vz = vz;
    int nan_a = 0, domain = 0;
    // ---- forward pass (message_passing.py:249-256): prior, x, linear, z
    {  // e1 = prior -> x : Prior.compute_forward_state_evolution (base_prior.py:66-69)
      double a_new;
      if (prior.kind == TRB_GAUSSIAN_PRIOR) {
        a_new = prior.p0;  // gaussian_prior.py:102-104
      } else {
        const double v = prior_measure(prior, TRB_MEASURE_V, a[7], 0.0, q, sh);
        a_new = clip_a_new(v, a[7], prior.amin, prior.amax);
      }
      nan_a |= (a_new != a_new);
      a[0] = damp(se.damp1, a[0], a_new);
    }
    a[1] = a[0];  // SISOVariable pass-through (sub_variables.py:33-37)
    if (!nan_a) {  // e3 = linear -> z : Channel.compute_forward_state_evolution (base_channel.py:19-22)
      const double v = channel_forward_error(ch, a[1], a[5], sh);
      const double a_new = clip_a_new(v, a[5], se.lin_amin, se.lin_amax);
      nan_a |= (a_new != a_new);
      a[2] = damp(se.damp3, a[2], a_new);
    }
    a[3] = a[2];
    // ---- backward pass (message_passing.py:258-265): likelihood, z, linear, x
    if (!nan_a) {  // e5 = likelihood -> z (base_likelihood.py:73-76)
      double a_new;
      if (lik.kind == TRB_GAUSSIAN_LIKELIHOOD) {
        a_new = lik.p0;  // gaussian_likelihood.py:66-68
      } else {
        const double v = lik_measure(lik, TRB_MEASURE_V, a[3], tau_z, q, sh, &domain);
        a_new = clip_a_new(v, a[3], lik.amin, lik.amax);
      }
      nan_a |= (a_new != a_new);
      a[4] = damp(se.damp5, a[4], a_new);
    }
    a[5] = a[4];
    if (!nan_a) {  // e7 = linear -> x (base_channel.py:24-27)
      const double v = channel_backward_error(ch, a[1], a[5], sh);
      const double a_new = clip_a_new(v, a[1], se.lin_amin, se.lin_amax);
      nan_a |= (a_new != a_new);
      a[6] = damp(se.damp7, a[6], a_new);
    }
    a[7] = a[6];
    if (nan_a || domain) {
      // check_message (message_passing.py:187-198): restore old_message_dag and raise
      flag |= domain ? TRB_FLAG_SE_DOMAIN : TRB_FLAG_NAN_A;
      flag |= TRB_FLAG_RESTORED;
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = old_a[e];
      stop = 1;
      break;
    }
    for (int e = 0; e < 8; e += 2)
      if (a[e] < 0) flag |= TRB_FLAG_NEG_A;
    // ---- update_variables (message_passing.py:267-269, base.py:167-170)
    vx = 1. / (a[0] + a[6]);
    vz = 1. / (a[2] + a[4]);
    ++done;
    const int row = it0 + it;
    if (threadIdx.x == 0 && se.rec_vx && row < se.max_records) {
      se.rec_vx[(size_t)row * G + g] = vx;
      se.rec_vz[(size_t)row * G + g] = vz;
    }
    // ---- EarlyStopping (callbacks.py:206-243), i = it
    if (se.es_tol >= 0) {
      const bool use_x = se.es_vars & 1, use_z = se.es_vars & 2;
      const bool below = (use_x && vx < se.es_min_variance) || (use_z && vz < se.es_min_variance);
      if (below) {
        stop = 1;
      } else if ((use_x && vx != vx) || (use_z && vz != vz)) {
        flag |= TRB_FLAG_RESTORED;
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = old_a[e];
        vx = old_vx;
        vz = old_vz;
        stop = 1;
      } else if (have_old_vs) {
        double tol = 0.0, increase = -INFINITY;
        if (use_x) {
          tol = fmax(tol, fabs(old_vx - vx));
          increase = fmax(increase, vx - old_vx);
        }
        if (use_z) {
          tol = fmax(tol, fabs(old_vz - vz));
          increase = fmax(increase, vz - old_vz);
        }
        if (tol < se.es_tol) {
          flag |= TRB_FLAG_CONVERGED;
          stop = 1;
        } else if (it > se.es_wait_increase && increase > se.es_max_increase) {
          flag |= TRB_FLAG_DIVERGED | TRB_FLAG_RESTORED;
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] = old_a[e];
          vx = old_vx;
          vz = old_vz;
          stop = 1;
        }
      }
      have_old_vs = true;
    }
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) se.edge_a[(size_t)e * G + g] = a[e];
    se.vx[g] = vx;
    se.vz[g] = vz;
    if (se.n_iter) se.n_iter[g] += done;
    if (se.flags && flag) atomicOr(&se.flags[g], flag);
    if (se.active && stop) se.active[g] = 0;
  }
}

int check_quad(const trb_quadrature* q) {
  if (!q || !q->x || !q->w || q->Q <= 0 || q->P <= 0 || !(q->kappa > 0)) return 0;
  if (!q->x2 || !q->w2 || q->Q2 <= 0 || q->P2 <= 0 || !(q->kappa2 > 0)) return 0;
  if ((long long)q->P * q->Q > (1 << 20) || (long long)q->P2 * q->Q2 > (1 << 14)) return 0;
  return 1;
}

}  // namespace

extern "C" size_t trb_sizeof_se(void) { return sizeof(trb_se); }

extern "C" int trb_se_measure(const trb_factor* factors, int factor_stride, int what, int B,
                              const double* a, const double* tau, const trb_quadrature* q,
                              double* out, int32_t* flags, void* stream) {
  TRB_CHECK_ARG(factors && a && out, "null pointer");
  TRB_CHECK_ARG(B > 0, "bad shape");
  TRB_CHECK_ARG(factor_stride == 0 || factor_stride == 1, "factor_stride must be 0 or 1");
  TRB_CHECK_ARG(what == TRB_MEASURE_V || what == TRB_MEASURE_A, "unknown measure");
  TRB_CHECK_ARG(check_quad(q), "bad quadrature rule");
  trb_launch_scope scope_(0, (cudaStream_t)stream);
  k_se_measure<<<B, kSeThreads, 0, (cudaStream_t)stream>>>(factors, factor_stride, what, a, tau,
                                                           make_quad(*q), out, flags);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_se_run(const trb_se* se, int it0, int n_iter, void* stream) {
  TRB_CHECK_ARG(se, "null pointer");
  TRB_CHECK_ARG(se->G > 0 && n_iter >= 0 && it0 >= 0, "bad shape");
  TRB_CHECK_ARG(se->prior && se->lik && se->tau_z && se->edge_a && se->vx && se->vz,
                "null pointer");
  TRB_CHECK_ARG(check_quad(&se->quad), "bad quadrature rule");
  if (se->channel == TRB_SE_MARCHENKO_PASTUR) {
    TRB_CHECK_ARG(se->alpha && se->mean_spectrum, "Marchenko-Pastur channel needs alpha, mean_spectrum");
  } else if (se->channel == TRB_SE_SPECTRUM) {
    TRB_CHECK_ARG(se->s2 && se->R > 0 && se->rank > 0 && se->rank <= se->R && se->Nz > 0 && se->Nx > 0,
                  "empirical channel needs its spectrum");
  } else {
    return trb_set_error(TRB_ERR_INVALID, "trb_se_run: unknown channel kind");
  }
  TRB_CHECK_ARG(!se->rec_vx || (se->rec_vz && it0 + n_iter <= se->max_records),
                "records too short");
  if (n_iter == 0) return TRB_OK;
  trb_launch_scope scope_(0, (cudaStream_t)stream);
  k_se_run<<<se->G, kSeThreads, 0, (cudaStream_t)stream>>>(*se, it0, n_iter);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}
