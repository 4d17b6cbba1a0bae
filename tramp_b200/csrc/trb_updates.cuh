// The z / x updates of the EP sweep as device functions: the per-element arithmetic and the
// per-instance scalars, shared by the kernels of trb_sweep.cu (one CTA / cluster per instance,
// chunked) and by the epilogue of the expanding GEMV (trb_linear.cu), where the CTA that completes
// an instance's expansion updates the instance on the spot.
//
// reference: algos/message_passing.py:249-269 (forward / backward pass, update_variables),
// :70-127 (constant damping), :187-209 (NaN check); base.py:152-161, 250-255;
// base_channel.py:9-17; gaussian_likelihood.py:68-71; algos/callbacks.py:206-286; metrics.py:5-14.
#pragma once
#include "trb_common.cuh"

namespace trb {

__device__ __forceinline__ int slots_of(int b, int R, int B, int G) {
  if (G <= 0) return 1;  // pre-reduced: the full sum sits in slot 0
  const int64_t T = (int64_t)B * R;
  const int kf = (int)part_owner((int64_t)b * R, T, G);
  const int kl = (int)part_owner((int64_t)b * R + R - 1, T, G);
  return kl - kf + 1;
}

// ---- per-instance scalars at the end of the z / x updates (one thread per instance) ----
__device__ __forceinline__ void z_tail(const trb_sweep& sw, int b, int light, double* __restrict__ stats,
                                       double a3, double a5, double a_hat, int all, double d2, double n2) {
  const int B = sw.B;
  double* ea = sw.edge_a;
  if (sw.snap_edge_a) {  // one-iteration-back state, see k_z_update
    for (int e = 2; e < 6; ++e) sw.snap_edge_a[e * B + b] = ea[e * B + b];
    sw.snap_vz[b] = sw.vz[b];
  }
  ea[2 * B + b] = a3;
  ea[3 * B + b] = a3;  // e4 = e3 (sub_variables.py:21-25)
  ea[4 * B + b] = a5;
  ea[5 * B + b] = a5;  // e6 = e5 (sub_variables.py:27-31)
  sw.vz[b] = 1. / a_hat;
  if (!light) {
    stats[b * 4 + 0] = d2;
    stats[b * 4 + 1] = n2;
  }
  if (all) atomicOr(&sw.flags[b], all);
}

__device__ __forceinline__ void x_tail(const trb_sweep& sw, int b, int it, const double* __restrict__ stats,
                                       double a7, double a_hat, int all, double d2, double n2,
                                       double e_pos, double e_neg) {
  const int B = sw.B, N = sw.N;
  double* ea = sw.edge_a;
  if (sw.snap_edge_a) {
    sw.snap_edge_a[6 * B + b] = ea[6 * B + b];
    sw.snap_edge_a[7 * B + b] = ea[7 * B + b];
    sw.snap_vx[b] = sw.vx[b];
  }
  ea[6 * B + b] = a7;
  ea[7 * B + b] = a7;  // e8 = e7
  const double vx = 1. / a_hat;
  sw.vx[b] = vx;
  if (all) atomicOr(&sw.flags[b], all);
  sw.n_iter[b] += 1;
  const bool rec = it < sw.max_records;
  if (rec && sw.rec_vx) sw.rec_vx[(size_t)it * B + b] = vx;
  if (rec && sw.rec_vz) sw.rec_vz[(size_t)it * B + b] = sw.vz[b];
  if (sw.x_true) {
    const double mse = e_pos / N, mse_neg = e_neg / N;
    if (rec && sw.rec_mse) sw.rec_mse[(size_t)it * B + b] = mse;
    if (rec && sw.rec_smse) sw.rec_smse[(size_t)it * B + b] = fmin(mse, mse_neg);
  }
  // EarlyStoppingEP, callbacks.py:258-286: tol = rms(new-old)/rms(new), max over the
  // tracked variables; needs a previous estimate, i.e. it > 0.
  double tol = nan("");
  if (sw.es_mode == 1 && sw.es_tol >= 0) {
    // EarlyStopping on the variances (callbacks.py:206-243); the previous values are the
    // one-iteration-back state's (snap_vx: saved above, snap_vz: saved by the z update)
    const int vars = sw.es_vars ? sw.es_vars : 3;
    const int stop = early_stopping_variance(vars, it, vx, sw.vz[b], sw.snap_vx[b], sw.snap_vz[b],
                                             sw.es_tol, sw.es_min_variance, sw.es_max_increase,
                                             sw.es_wait_increase, &tol);
    if (stop) {
      sw.active[b] = 0;
      atomicOr(&sw.flags[b], stop);
    }
  } else if (it > 0) {
    const double tol_x = sqrt(d2 / N) / sqrt(n2 / N);
    const double tol_z = sqrt(stats[b * 4 + 0] / sw.M) / sqrt(stats[b * 4 + 1] / sw.M);
    const int vars = sw.es_vars ? sw.es_vars : 3;
    tol = (vars & 1) ? tol_x : tol_z;
    if ((vars & 2) && tol_z > tol) tol = tol_z;
    if (sw.es_tol >= 0) {
      if (tol < sw.es_tol) {
        sw.active[b] = 0;
        atomicOr(&sw.flags[b], TRB_FLAG_CONVERGED);
      } else if (it > sw.es_wait_increase && tol > sw.es_max_increase) {
        sw.active[b] = 0;
        atomicOr(&sw.flags[b], TRB_FLAG_DIVERGED);
      }
    }
  }
  if (rec && sw.rec_tol) sw.rec_tol[(size_t)it * B + b] = tol;
  if (all & (TRB_FLAG_NAN_A | TRB_FLAG_NAN_B)) sw.active[b] = 0;
}

// ---- z update with a constant likelihood message (Gaussian likelihood) ----------------------
struct ZScalars {
  double a3n, a3, ainv3, a5n, a5, a_hat;
};

__device__ __forceinline__ ZScalars z_scalars(const trb_sweep& sw, int b) {
  const int B = sw.B;
  const double* ea = sw.edge_a;
  ZScalars z;
  const double a6 = ea[5 * B + b];
  z.a3n = clip_a_new(sw.vlin[b], a6, sw.lin_amin, sw.lin_amax);  // base_channel.py:9-12
  z.a3 = damp(sw.damp3, ea[2 * B + b], z.a3n);
  z.ainv3 = a6 + z.a3n;
  z.a5n = sw.lik.p0;  // gaussian_likelihood.py:68-71
  z.a5 = damp(sw.damp5, ea[4 * B + b], z.a5n);
  z.a_hat = z.a3 + z.a5;
  return z;
}

__device__ __forceinline__ int z_scalar_flags(const ZScalars& z) {
  int flag = 0;
  if (z.a3n != z.a3n || z.a5n != z.a5n) flag |= TRB_FLAG_NAN_A;
  if (z.a3n < 0 || z.a5n < 0) flag |= TRB_FLAG_NEG_A;
  return flag;
}

// Elements i = start + u * stride (u < E), every load issued before the first use: e3 (= e4),
// e5 (= e6), the posterior mean of z and its tolerance sums.  CG: the expansion slots were written
// by other CTAs of this launch -> read them from L2.
template <int E, bool CG>
__device__ __forceinline__ void z_elements(const trb_sweep& sw, int b, int ns, int first,
                                           const trb_peers* peers, const ZScalars& z, int start,
                                           int stride, double (&red)[2], int& flag) {
  const int M = sw.M, ld = sw.ldm;
  const size_t off = (size_t)b * ld;
  const double* part = sw.part + (size_t)b * sw.nslots * ld;
  const double* b6 = ((first && sw.b6_init) ? sw.b6_init : sw.b5) + off;
  double* b3 = sw.b3 + off;
  double* b5 = sw.b5 + off;
  double* rz = sw.rz + off;
  const double* y = sw.y + off;
  const bool snap = sw.snap_edge_a != nullptr;
  const bool use_peers = peers != nullptr && peers->n > 0;
  double rx[E], b6v[E], b3o[E], yv[E], b5o[E], ro[E];
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = start + u * stride;
    rx[u] = 0.0;
    if (i < M) {
      rx[u] = use_peers ? peers_sum(*peers, off + i) : (CG ? __ldcg(part + i) : part[i]);
      b6v[u] = b6[i];
      b3o[u] = b3[i];
      yv[u] = y[i];
      b5o[u] = b5[i];
      ro[u] = rz[i];
    }
  }
  if (!use_peers) {
    for (int sl = 1; sl < ns; ++sl) {
#pragma unroll
      for (int u = 0; u < E; ++u) {
        const int i = start + u * stride;
        if (i < M) rx[u] += CG ? __ldcg(part + (size_t)sl * ld + i) : part[(size_t)sl * ld + i];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = start + u * stride;
    if (i < M) {
      const double b3n = rx[u] * z.ainv3 - b6v[u];
      if (b3n != b3n) flag |= TRB_FLAG_NAN_B;
      const double b3v = damp(sw.damp3, b3o[u], b3n);
      const double b5n = yv[u] * sw.lik.p0;
      if (b5n != b5n) flag |= TRB_FLAG_NAN_B;
      const double b5v = damp(sw.damp5, b5o[u], b5n);
      const double rnew = (b3v + b5v) / z.a_hat;  // base.py:152-161
      if (snap) {  // one-iteration-back state (message_passing.py:356)
        sw.snap_b3[off + i] = b3o[u];
        sw.snap_b5[off + i] = b5o[u];
        sw.snap_rz[off + i] = ro[u];
      }
      b3[i] = b3v;
      b5[i] = b5v;
      rz[i] = rnew;
      red[0] += (rnew - ro[u]) * (rnew - ro[u]);
      red[1] += rnew * rnew;
    }
  }
}

// ---- x update ------------------------------------------------------------------------------
struct XScalars {
  double a1, a7n, a7, ainv7, a_hat;
};

__device__ __forceinline__ XScalars x_scalars(const trb_sweep& sw, int b) {
  const int B = sw.B;
  const double* ea = sw.edge_a;
  XScalars x;
  x.a1 = ea[1 * B + b];  // e2 (= e1)
  x.a7n = clip_a_new(sw.vlin[b], x.a1, sw.lin_amin, sw.lin_amax);  // base_channel.py:14-17
  x.a7 = damp(sw.damp7, ea[6 * B + b], x.a7n);
  x.ainv7 = x.a1 + x.a7n;
  x.a_hat = x.a1 + x.a7;
  return x;
}

__device__ __forceinline__ int x_scalar_flags(const XScalars& x) {
  int flag = 0;
  if (x.a7n != x.a7n) flag |= TRB_FLAG_NAN_A;
  if (x.a7n < 0) flag |= TRB_FLAG_NEG_A;
  return flag;
}

// red: sum dr^2, sum r^2, sum (r - x)^2, sum (r + x)^2
template <int E, bool CG>
__device__ __forceinline__ void x_elements(const trb_sweep& sw, int b, int ns, const trb_peers* peers,
                                           const XScalars& x, int start, int stride, double (&red)[4],
                                           int& flag) {
  const int N = sw.N, ld = sw.ldn;
  const size_t off = (size_t)b * ld;
  const double* part = sw.part + (size_t)b * sw.nslots * ld;
  const bool null_space = (sw.R_total > 0 ? sw.R_total : sw.R) < N;
  const double* b1 = sw.b1 + off;
  double* b7 = sw.b7 + off;
  double* rx = sw.rx + off;
  const double* xt = sw.x_true ? sw.x_true + off : nullptr;
  const bool snap = sw.snap_edge_a != nullptr;
  const bool use_peers = peers != nullptr && peers->n > 0;
  double rzv[E], b1v[E], b7o[E], ro[E], xv[E];
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = start + u * stride;
    rzv[u] = 0.0;
    xv[u] = 0.0;
    if (i < N) {
      rzv[u] = use_peers ? peers_sum(*peers, off + i) : (CG ? __ldcg(part + i) : part[i]);
      b1v[u] = b1[i];
      b7o[u] = b7[i];
      ro[u] = rx[i];
      if (xt) xv[u] = xt[i];
    }
  }
  if (!use_peers) {
    for (int sl = 1; sl < ns; ++sl) {
#pragma unroll
      for (int u = 0; u < E; ++u) {
        const int i = start + u * stride;
        if (i < N) rzv[u] += CG ? __ldcg(part + (size_t)sl * ld + i) : part[(size_t)sl * ld + i];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = start + u * stride;
    if (i < N) {
      double r = rzv[u];
      if (null_space) r = b1v[u] / x.a1 + r;
      const double b7n = r * x.ainv7 - b1v[u];
      if (b7n != b7n) flag |= TRB_FLAG_NAN_B;
      const double b7v = damp(sw.damp7, b7o[u], b7n);
      if (snap) {
        sw.snap_b7[off + i] = b7o[u];
        sw.snap_rx[off + i] = ro[u];
      }
      b7[i] = b7v;
      const double rnew = (b1v[u] + b7v) / x.a_hat;
      rx[i] = rnew;
      red[0] += (rnew - ro[u]) * (rnew - ro[u]);
      red[1] += rnew * rnew;
      red[2] += (rnew - xv[u]) * (rnew - xv[u]);  // metrics.py:5-6
      red[3] += (rnew + xv[u]) * (rnew + xv[u]);  // metrics.py:9-14
    }
  }
}

}  // namespace trb
