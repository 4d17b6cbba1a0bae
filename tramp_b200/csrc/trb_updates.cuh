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
  if (T * G <= (int64_t)0xffffffffu) {  // part_owner in 32-bit arithmetic: every thread of the update kernels runs this
    const unsigned int t = (unsigned int)T, g = (unsigned int)G, r0 = (unsigned int)b * (unsigned int)R;
    return (int)(((r0 + (unsigned int)R) * g - 1u) / t - ((r0 + 1u) * g - 1u) / t) + 1;
  }
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

// ---- element arithmetic, split into "issue every load" and "compute" -------------------------
// The chunked kernels first issue ALL their loads -- the instance's `active` flag, the scalars and
// the vector elements are independent addresses -- and only then look at any of them: one memory
// round trip instead of a chain of four (ncu, r02c: these kernels are bound by the latency of
// their dependent accesses, not by bandwidth).  The expansion arrives in per-CTA slots; slots 0 and
// 1 are loaded with the rest (an instance rarely spans more than two CTAs), further slots after.

// ---- z update with a constant likelihood message (Gaussian likelihood) ----------------------
struct ZRaw {  // scalars as loaded
  double a6, a3_old, a5_old, vlin;
};

__device__ __forceinline__ ZRaw z_raw(const trb_sweep& sw, int b) {
  const int B = sw.B;
  const double* ea = sw.edge_a;
  ZRaw r;
  r.a6 = ea[5 * B + b];
  r.a3_old = ea[2 * B + b];
  r.a5_old = ea[4 * B + b];
  r.vlin = sw.vlin[b];
  return r;
}

struct ZScalars {
  double a3n, a3, ainv3, a5n, a5, a_hat;
};

__device__ __forceinline__ ZScalars z_scalars(const trb_sweep& sw, const ZRaw& r) {
  ZScalars z;
  z.a3n = clip_a_new(r.vlin, r.a6, sw.lin_amin, sw.lin_amax);  // base_channel.py:9-12
  z.a3 = damp(sw.damp3, r.a3_old, z.a3n);
  z.ainv3 = r.a6 + z.a3n;
  z.a5n = sw.lik.p0;  // gaussian_likelihood.py:68-71
  z.a5 = damp(sw.damp5, r.a5_old, z.a5n);
  z.a_hat = z.a3 + z.a5;
  return z;
}

__device__ __forceinline__ int z_scalar_flags(const ZScalars& z) {
  int flag = 0;
  if (z.a3n != z.a3n || z.a5n != z.a5n) flag |= TRB_FLAG_NAN_A;
  if (z.a3n < 0 || z.a5n < 0) flag |= TRB_FLAG_NEG_A;
  return flag;
}

template <int E>
struct ZLoads {
  double rx[E], rx1[E], b6v[E], b3o[E], yv[E], b5o[E], ro[E];
};

// elements i = start + u * stride (u < E); ns = slots of this instance's expansion
template <int E>
__device__ __forceinline__ void z_load(const trb_sweep& sw, int b, int ns, int first,
                                       const trb_peers* peers, int start, int stride, ZLoads<E>& l) {
  const int M = sw.M, ld = sw.ldm;
  const size_t off = (size_t)b * ld;
  const double* part = sw.part + (size_t)b * sw.nslots * ld;
  const double* b6 = ((first && sw.b6_init) ? sw.b6_init : sw.b5) + off;
  const bool use_peers = peers != nullptr && peers->n > 0;
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = start + u * stride;
    l.rx[u] = 0.0;
    l.rx1[u] = 0.0;
    if (i < M) {
      if (!use_peers) {
        l.rx[u] = part[i];
        if (ns > 1) l.rx1[u] = part[(size_t)ld + i];
      }
      l.b6v[u] = b6[i];
      l.b3o[u] = sw.b3[off + i];
      l.yv[u] = sw.y[off + i];
      l.b5o[u] = sw.b5[off + i];
      l.ro[u] = sw.rz[off + i];
    }
  }
}

// e3 (= e4), e5 (= e6), the posterior mean of z and its tolerance sums
template <int E>
__device__ __forceinline__ void z_compute(const trb_sweep& sw, int b, int ns, const trb_peers* peers,
                                          const ZScalars& z, int start, int stride, ZLoads<E>& l,
                                          double (&red)[2], int& flag) {
  const int M = sw.M, ld = sw.ldm;
  const size_t off = (size_t)b * ld;
  const double* part = sw.part + (size_t)b * sw.nslots * ld;
  const bool snap = sw.snap_edge_a != nullptr;
  const bool use_peers = peers != nullptr && peers->n > 0;
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = start + u * stride;
    if (i < M) {
      if (use_peers) l.rx[u] = peers_sum(*peers, off + i);
      else if (ns > 1) l.rx[u] += l.rx1[u];
    }
  }
  if (!use_peers) {
    for (int sl = 2; sl < ns; ++sl) {
#pragma unroll
      for (int u = 0; u < E; ++u) {
        const int i = start + u * stride;
        if (i < M) l.rx[u] += part[(size_t)sl * ld + i];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = start + u * stride;
    if (i < M) {
      const double b3n = l.rx[u] * z.ainv3 - l.b6v[u];
      if (b3n != b3n) flag |= TRB_FLAG_NAN_B;
      const double b3v = damp(sw.damp3, l.b3o[u], b3n);
      const double b5n = l.yv[u] * sw.lik.p0;
      if (b5n != b5n) flag |= TRB_FLAG_NAN_B;
      const double b5v = damp(sw.damp5, l.b5o[u], b5n);
      const double rnew = (b3v + b5v) / z.a_hat;  // base.py:152-161
      if (snap) {  // one-iteration-back state (message_passing.py:356)
        sw.snap_b3[off + i] = l.b3o[u];
        sw.snap_b5[off + i] = l.b5o[u];
        sw.snap_rz[off + i] = l.ro[u];
      }
      sw.b3[off + i] = b3v;
      sw.b5[off + i] = b5v;
      sw.rz[off + i] = rnew;
      red[0] += (rnew - l.ro[u]) * (rnew - l.ro[u]);
      red[1] += rnew * rnew;
    }
  }
}

// ---- x update ------------------------------------------------------------------------------
struct XRaw {
  double a1, a7_old, vlin;
};

__device__ __forceinline__ XRaw x_raw(const trb_sweep& sw, int b) {
  const int B = sw.B;
  const double* ea = sw.edge_a;
  XRaw r;
  r.a1 = ea[1 * B + b];  // e2 (= e1)
  r.a7_old = ea[6 * B + b];
  r.vlin = sw.vlin[b];
  return r;
}

struct XScalars {
  double a1, a7n, a7, ainv7, a_hat;
};

__device__ __forceinline__ XScalars x_scalars(const trb_sweep& sw, const XRaw& r) {
  XScalars x;
  x.a1 = r.a1;
  x.a7n = clip_a_new(r.vlin, r.a1, sw.lin_amin, sw.lin_amax);  // base_channel.py:14-17
  x.a7 = damp(sw.damp7, r.a7_old, x.a7n);
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

template <int E>
struct XLoads {
  double rz[E], rz1[E], b1v[E], b7o[E], ro[E], xv[E];
};

template <int E>
__device__ __forceinline__ void x_load(const trb_sweep& sw, int b, int ns, const trb_peers* peers,
                                       int start, int stride, XLoads<E>& l) {
  const int N = sw.N, ld = sw.ldn;
  const size_t off = (size_t)b * ld;
  const double* part = sw.part + (size_t)b * sw.nslots * ld;
  const bool use_peers = peers != nullptr && peers->n > 0;
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = start + u * stride;
    l.rz[u] = 0.0;
    l.rz1[u] = 0.0;
    l.xv[u] = 0.0;
    if (i < N) {
      if (!use_peers) {
        l.rz[u] = part[i];
        if (ns > 1) l.rz1[u] = part[(size_t)ld + i];
      }
      l.b1v[u] = sw.b1[off + i];
      l.b7o[u] = sw.b7[off + i];
      l.ro[u] = sw.rx[off + i];
      if (sw.x_true) l.xv[u] = sw.x_true[off + i];
    }
  }
}

// red: sum dr^2, sum r^2, sum (r - x)^2, sum (r + x)^2
template <int E>
__device__ __forceinline__ void x_compute(const trb_sweep& sw, int b, int ns, const trb_peers* peers,
                                          const XScalars& x, int start, int stride, XLoads<E>& l,
                                          double (&red)[4], int& flag) {
  const int N = sw.N, ld = sw.ldn;
  const size_t off = (size_t)b * ld;
  const double* part = sw.part + (size_t)b * sw.nslots * ld;
  const bool null_space = (sw.R_total > 0 ? sw.R_total : sw.R) < N;
  const bool snap = sw.snap_edge_a != nullptr;
  const bool use_peers = peers != nullptr && peers->n > 0;
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = start + u * stride;
    if (i < N) {
      if (use_peers) l.rz[u] = peers_sum(*peers, off + i);
      else if (ns > 1) l.rz[u] += l.rz1[u];
    }
  }
  if (!use_peers) {
    for (int sl = 2; sl < ns; ++sl) {
#pragma unroll
      for (int u = 0; u < E; ++u) {
        const int i = start + u * stride;
        if (i < N) l.rz[u] += part[(size_t)sl * ld + i];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = start + u * stride;
    if (i < N) {
      double r = l.rz[u];
      if (null_space) r = l.b1v[u] / x.a1 + r;
      const double b7n = r * x.ainv7 - l.b1v[u];
      if (b7n != b7n) flag |= TRB_FLAG_NAN_B;
      const double b7v = damp(sw.damp7, l.b7o[u], b7n);
      if (snap) {
        sw.snap_b7[off + i] = l.b7o[u];
        sw.snap_rx[off + i] = l.ro[u];
      }
      sw.b7[off + i] = b7v;
      const double rnew = (l.b1v[u] + b7v) / x.a_hat;
      sw.rx[off + i] = rnew;
      red[0] += (rnew - l.ro[u]) * (rnew - l.ro[u]);
      red[1] += rnew * rnew;
      red[2] += (rnew - l.xv[u]) * (rnew - l.xv[u]);  // metrics.py:5-6
      red[3] += (rnew + l.xv[u]) * (rnew + l.xv[u]);  // metrics.py:9-14
    }
  }
}

// ---- CTA-level sums for the chunked kernels: one barrier -------------------------------------
// Every warp leaves its K sums and its OR of `flag` in shared memory; after the barrier thread 0
// (only) holds the CTA totals, added in warp order.  sh: K * 8 doubles, shi: 8 ints, <= 8 warps.
template <int K>
__device__ __forceinline__ void cta_sums_to_thread0(double (&v)[K], int& flag, double* sh, int* shi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  flag = __reduce_or_sync(0xffffffffu, flag);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) sh[k * 8 + warp] = v[k];
    shi[warp] = flag;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double t = 0.0;
      for (int w = 0; w < nwarp; ++w) t += sh[k * 8 + w];
      v[k] = t;
    }
    int all = 0;
    for (int w = 0; w < nwarp; ++w) all |= shi[w];
    flag = all;
  }
}

// Thread 0 of a chunk CTA has written the chunk's sums: count the chunk in with release / acquire
// semantics on the counter itself (no separate fence); true for the thread whose chunk completes
// the instance -- it may then read every chunk's sums.
__device__ __forceinline__ bool chunk_arrive_last(unsigned int* cnt, int nchunk) {
  unsigned int before;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(before) : "l"(cnt), "r"(1u) : "memory");
  const bool last = (before + 1u == (unsigned int)nchunk);
  if (last) *cnt = 0;  // nobody else touches the counter before the next launch
  return last;
}

}  // namespace trb
