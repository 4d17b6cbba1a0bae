// Persistent whole-sweep kernel for ONE launch-bound instance.
//
// reference: the same loop as trb_sweep.cu -- algos/message_passing.py:330-357
// (iterate), :249-269 (forward / backward pass, update_variables), :70-127
// (constant damping), :187-209 (NaN check), algos/callbacks.py:250-286
// (EarlyStoppingEP) -- on the chain prior -> x -> LinearChannel -> z -> likelihood.
//
// A single instance of the size the reference itself runs (N = 1000 ... a few
// thousand) streams 12 ... 200 MB of operators per iteration: 2 ... 30 us of HBM /
// L2 time, against a dozen dependent kernel launches per iteration in
// trb_sweep.cu (63 ... 170 us per iteration measured, even replayed as a CUDA
// graph).  Here ALL iterations of the sweep run inside one cooperative launch of
// one CTA per SM, with four grid-wide barriers per iteration:
//
//   [all]   F1  prior update over all N (every CTA redundantly: O(N) work, no barrier)
//   [rows]  P1  tz = V_R^T b1 for this CTA's singular indices
//   ---- barrier 1
//   [all]   S1  coefficients of all R indices (shared memory), forward variance
//   [cols]  P2  r_x = U_R coef for this CTA's columns of z (no partial sums to merge)
//   ---- barrier 2
//   [all]   Z   e3, likelihood moments, e5, posterior of z over all M
//   [rows]  P3  tx = U_R^T b5 for this CTA's singular indices
//   ---- barrier 3
//   [all]   S2  coefficients, backward variance
//   [cols]  P4  r_z = [b2/a2 +] V_R coef for this CTA's columns of x
//   ---- barrier 4
//   [all]   X   e7, posterior of x, MSE / tolerance records, early-stop decision
//
// The elementwise phases are recomputed by every CTA (identical instructions on
// identical data, hence identical bits), which is what removes the two barriers
// a distributed mean would need per factor; each CTA writes only its own slice
// of the state vectors.  The state is double-buffered between the live buffers
// and the one-iteration-back snapshot buffers of the descriptor (iteration k
// reads parity k, writes parity k+1), so the snapshot the reference keeps
// (`old_message_dag`) costs nothing and a NaN / divergence rolls back by
// choosing the other parity.  Each operator is still streamed exactly twice per
// iteration: 16 R (N + M) bytes.
#include <cooperative_groups.h>
#include "trb_moments.cuh"

namespace cg = cooperative_groups;
using namespace trb;

bool trb_profile_events_enabled();

namespace {

constexpr int kPsThreads = 512;

struct StateBuf {
  double* b1;
  double* b3;
  double* b5;
  double* b7;
  double* rx;
  double* rz;
  double* tx;
};

// t[i] = A[i, :] . vec for rows [r0, r1); vec (length n) sits in shared memory.
// A row belongs to a group of `wpr` warps (wpr = 1 when the CTA owns at least as
// many rows as it has warps), lanes stride over column pairs with 16-byte loads,
// kProjUnroll of them in flight per lane; no block-wide barrier unless wpr > 1.
constexpr int kProjUnroll = 8;
__device__ __forceinline__ void project_rows(const double* __restrict__ A, int ld, int n, int r0,
                                             int r1, const double* vec, double* __restrict__ t,
                                             double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nrows = r1 - r0;
  if (nrows <= 0) return;  // uniform over the CTA
  int wpr = 1;
  while (wpr * 2 * nrows <= nwarp) wpr *= 2;
  const int rpr = nwarp / wpr;  // rows in flight per round
  const int slot = warp / wpr, part = warp % wpr;
  const int npair = n >> 1;
  const double2* v2 = reinterpret_cast<const double2*>(vec);
  for (int base = r0; base < r1; base += rpr) {
    const int i = base + slot;
    double acc = 0.0;
    if (i < r1) {
      const double* row = A + (size_t)i * ld;
      const int stride = wpr * 32;
      // every batch is fully predicated (no serial tail): a load whose result is
      // not consumed for ~1 us is what this loop is bound by
      for (int q0 = part * 32 + lane; q0 < npair; q0 += kProjUnroll * stride) {
        double2 x[kProjUnroll];
#pragma unroll
        for (int u = 0; u < kProjUnroll; ++u) {
          const int q = q0 + u * stride;
          x[u] = (q < npair) ? ldg_stream(row + 2 * q) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < kProjUnroll; ++u) {
          const int q = q0 + u * stride;
          if (q < npair) {
            const double2 v = v2[q];
            acc += x[u].x * v.x;
            acc += x[u].y * v.y;
          }
        }
      }
      if ((n & 1) && part == 0 && lane == 0) acc += row[n - 1] * vec[n - 1];
      acc = warp_sum(acc);
      if (wpr == 1 && lane == 0) t[i] = acc;
    }
    if (wpr > 1) {  // uniform over the CTA
      __syncthreads();
      if (lane == 0) red[warp] = acc;
      __syncthreads();
      if ((int)threadIdx.x < rpr && base + (int)threadIdx.x < r1) {
        double sum = 0.0;
        for (int p = 0; p < wpr; ++p) sum += red[threadIdx.x * wpr + p];
        t[base + threadIdx.x] = sum;
      }
    }
  }
}

// out[j] = sum_i coef[i] A[i, j] for columns [c0, c1); coef (length R) in shared
// memory.  Threads form (row group, column) pairs, kExpUnroll loads in flight per
// thread; the row groups' partial sums are added in a fixed order.
constexpr int kExpUnroll = 16;
__device__ __forceinline__ void expand_cols(const double* __restrict__ A, int ld, int R, int c0,
                                            int c1, const double* coef, double* __restrict__ out,
                                            double* red) {
  const int w = c1 - c0;
  if (w <= 0) return;  // uniform over the CTA
  int W = 1;
  while (W < w) W <<= 1;
  const int T = blockDim.x;
  for (int cb = 0; cb < w; cb += T) {  // w > T only if there are fewer CTAs than n / T
    const int ww = min(w - cb, T);
    const int Wc = W < T ? W : T;
    const int tcol = threadIdx.x % Wc, trow = threadIdx.x / Wc, nrg = T / Wc;
    double acc = 0.0;
    if (tcol < ww) {
      const double* Ac = A + c0 + cb + tcol;
      for (int i0 = trow; i0 < R; i0 += kExpUnroll * nrg) {  // fully predicated batches, no serial tail
        double x[kExpUnroll];
#pragma unroll
        for (int u = 0; u < kExpUnroll; ++u) {
          const int i = i0 + u * nrg;
          x[u] = (i < R) ? ldg_stream1(Ac + (size_t)i * ld) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kExpUnroll; ++u) {
          const int i = i0 + u * nrg;
          if (i < R) acc += coef[i] * x[u];
        }
      }
    }
    __syncthreads();
    red[threadIdx.x] = acc;
    __syncthreads();
    if ((int)threadIdx.x < ww) {
      double s = 0.0;
      for (int g = 0; g < nrg; ++g) s += red[g * Wc + threadIdx.x];
      out[c0 + cb + threadIdx.x] = s;
    }
  }
}

// Forward / backward variance of the channel (linear_channel.py:58-67, 91-105) and
// the coefficients of all R singular indices into shared memory (:74; see
// k_lin_rescale in trb_linear.cu for the three forms).
__device__ __forceinline__ double rescale_all(int dir, int R, int Nz, int Nx, int rank,
                                              bool null_space, const double* __restrict__ s,
                                              const double* __restrict__ s2, double az, double ax,
                                              const double* __restrict__ tz,
                                              const double* __restrict__ tx, double* coef,
                                              double* sh) {
  double az_v = az;
  if (dir == 1) az_v = (az != az) ? az : fmax(1e-11, az);
  double v;
  if (dir == 0 && ax == 0) {
    double part = 0.0;
    for (int i = threadIdx.x; i < rank; i += blockDim.x) part += s2[i];
    const double s_mean = block_sum(part, sh) / rank;
    v = s_mean * rank / (Nx * az);
  } else {
    double n_eff;
    if (ax == 0) {
      n_eff = 0.;
    } else {
      const double ratio = az_v / ax;
      if (ratio == 0) {
        n_eff = (double)rank / Nz;
      } else {
        double part = 0.0;
        for (int i = threadIdx.x; i < rank; i += blockDim.x) part += s2[i] / (ratio + s2[i]);
        n_eff = block_sum(part, sh) / Nz;
      }
    }
    if (dir == 0) {
      const double alpha = (double)Nx / Nz;
      v = n_eff / (alpha * ax);
    } else {
      v = (1 - n_eff) / az_v;
    }
  }
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const double si = s[i], s2i = s2[i];
    const double res = 1 / (az + ax * s2i);
    const double tzi = __ldcg(tz + i), txi = __ldcg(tx + i);
    double c;
    if (dir == 0) {
      c = si * (res * (tzi + si * txi));
    } else if (!null_space) {
      c = res * (tzi + si * txi);
    } else {
      c = res * (si * txi - (ax * s2i / az) * tzi);
    }
    coef[i] = c;
  }
  __syncthreads();
  return v;
}

// kCluster = false: one CTA per SM, cooperative launch, cg grid barrier.
// kCluster = true : the whole grid is ONE thread-block cluster (16 CTAs) and the
//   barrier is the hardware cluster barrier: for the smallest instances (N <~ 600)
//   the four barriers, not the operator traffic, are the cost (18.5 vs 26 us per
//   iteration measured); beyond, 16 SMs cannot pull the operators fast enough.
template <bool kCluster>
__global__ void __launch_bounds__(kPsThreads, 1)
k_sweep_persistent(trb_sweep sw, int it0, int n_iter, int fresh, int ldmax) {
  auto sync_all = [&]() {
    if constexpr (kCluster) {
      __threadfence();  // global writes of this CTA before the release-arrive
      cluster_sync();   // barrier.cluster.arrive.release + wait.acquire
    } else {
      cg::this_grid().sync();
    }
  };
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;              // [ldmax] moments r, then the vector being projected (b1 / b5)
  double* sB = smem + ldmax;      // [ldmax] e8 (= b7) carried from X to the next F1; b3 inside Z
  double* sC = smem + 2 * ldmax;  // [R] coefficients
  __shared__ double sh[33 * 4];
  __shared__ double red[kPsThreads];
  if (!sw.active[0]) return;  // uniform over the grid
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int N = sw.N, M = sw.M, R = sw.R, ldn = sw.ldn, ldm = sw.ldm;
  const bool null_space = R < N;
  const int r0 = (int)part_begin(cta, R, G), r1 = (int)part_begin(cta + 1, R, G);
  const int cm0 = (int)part_begin(cta, M, G), cm1 = (int)part_begin(cta + 1, M, G);
  const int cn0 = (int)part_begin(cta, N, G), cn1 = (int)part_begin(cta + 1, N, G);
  StateBuf P[2];
  P[0] = {sw.b1, sw.b3, sw.b5, sw.b7, sw.rx, sw.rz, sw.tx};
  P[1] = {sw.snap_b1, sw.snap_b3, sw.snap_b5, sw.snap_b7, sw.snap_rx, sw.snap_rz, sw.snap_tx};
  double* tz = sw.tz;      // [R]   written by row owners, read by all after barrier 1
  double* rxl = sw.scr_m;  // [ldm] U_R coef, written by column owners
  double* rzl = sw.scr_n;  // [ldn] V_R coef
  const double* y = sw.y;
  const double* xt = sw.x_true;
  const bool const_prior = sw.prior.kind == TRB_GAUSSIAN_PRIOR;
  const bool const_lik = sw.lik.kind == TRB_GAUSSIAN_LIKELIHOOD;

  double a[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = sw.edge_a[e];
  double vx = sw.vx[0], vz = sw.vz[0];
  int flag = 0, par = 0, done = 0, stop_flags = 0;
  bool rolled_back = false;

  // e8 of the first iteration into sB; tx of the first iteration into P[0].tx
  {
    const double* b8 = (fresh && sw.b8_init) ? sw.b8_init : sw.b7;
    for (int i = tid; i < N; i += T) sB[i] = b8[i];
    if (fresh == 2) {
      for (int i = r0 + tid; i < r1; i += T) P[0].tx[i] = 0.0;
    } else if (fresh == 1) {
      const double* b6 = sw.b6_init ? sw.b6_init : sw.b5;
      for (int i = tid; i < M; i += T) sA[i] = b6[i];
      __syncthreads();
      project_rows(sw.Ut, ldm, M, r0, r1, sA, P[0].tx, red);
    }
    __syncthreads();
  }

  for (int k = 0; k < n_iter; ++k) {
    const int it = it0 + k;
    const bool first = fresh && k == 0;
    const StateBuf cur = P[par], nxt = P[par ^ 1];
    double a_old[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a_old[e] = a[e];
    const double vx_old = vx, vz_old = vz;
    int it_flag = 0;

    // ---- F1: prior, e8 -> e1 (= e2).  base_prior.py:13-16, gaussian_prior.py:86-89
    {
      const double a8 = a[7];
      double a1n;
      if (const_prior) {
        a1n = sw.prior.p0;
        for (int i = tid; i < N; i += T) {
          const double bn = sw.prior.p1;
          const double bd = damp(sw.damp1, __ldcg(cur.b1 + i), bn);
          sA[i] = bd;
          if (i >= cn0 && i < cn1) nxt.b1[i] = bd;
        }
      } else {
        double vsum = 0.0;
        for (int i = tid; i < N; i += T) {
          const RV m = factor_moments(sw.prior, a8, sB[i], 0.0);
          sA[i] = m.r;
          vsum += m.v;
        }
        const double v = block_sum(vsum, sh) / N;
        a1n = clip_a_new(v, a8, sw.prior.amin, sw.prior.amax);
        const double ainv = a8 + a1n;
        for (int i = tid; i < N; i += T) {
          const double bn = sA[i] * ainv - sB[i];
          if (bn != bn) it_flag |= TRB_FLAG_NAN_B;
          const double bd = damp(sw.damp1, __ldcg(cur.b1 + i), bn);
          sA[i] = bd;
          if (i >= cn0 && i < cn1) nxt.b1[i] = bd;
        }
      }
      if (a1n != a1n) it_flag |= TRB_FLAG_NAN_A;
      if (a1n < 0) it_flag |= TRB_FLAG_NEG_A;
      a[0] = damp(sw.damp1, a[0], a1n);
      a[1] = a[0];  // sub_variables.py:21-25
    }
    __syncthreads();
    // ---- P1: tz = V_R^T b2
    project_rows(sw.Vt, ldn, N, r0, r1, sA, tz, red);
    sync_all();  // barrier 1: tz (and, in the first iteration, tx) complete

    // ---- S1 + P2: forward variance, r_x = U_R coef
    const double vlin_f =
        rescale_all(0, R, N, M, sw.rank, null_space, sw.s, sw.s2, a[1], a[5], tz, cur.tx, sC, sh);
    expand_cols(sw.Ut, ldm, R, cm0, cm1, sC, rxl, red);
    sync_all();  // barrier 2: rxl complete

    // ---- Z: e3 (= e4), likelihood e5 (= e6), posterior of z
    double dz2 = 0.0, nz2 = 0.0;
    {
      const double a6 = a[5];
      const double a3n = clip_a_new(vlin_f, a6, sw.lin_amin, sw.lin_amax);  // base_channel.py:9-12
      const double a3 = damp(sw.damp3, a[2], a3n);
      const double ainv3 = a6 + a3n;
      const double* b6 = (first && sw.b6_init) ? sw.b6_init : cur.b5;
      double vsum = 0.0;
      for (int i = tid; i < M; i += T) {
        const double b3n = __ldcg(rxl + i) * ainv3 - __ldcg(b6 + i);
        if (b3n != b3n) it_flag |= TRB_FLAG_NAN_B;
        const double b3v = damp(sw.damp3, __ldcg(cur.b3 + i), b3n);
        sB[i] = b3v;
        if (i >= cm0 && i < cm1) nxt.b3[i] = b3v;
        if (!const_lik) {
          const RV m = factor_moments(sw.lik, a3, b3v, y[i]);
          sA[i] = m.r;
          vsum += m.v;
        }
      }
      double a5n;
      if (const_lik) {
        a5n = sw.lik.p0;  // gaussian_likelihood.py:68-71
      } else {
        const double v = block_sum(vsum, sh) / M;
        a5n = clip_a_new(v, a3, sw.lik.amin, sw.lik.amax);  // base_likelihood.py:25-28
      }
      const double a5 = damp(sw.damp5, a[4], a5n);
      const double ainv5 = a3 + a5n;
      const double a_hat = a3 + a5;
      double rd[2] = {0.0, 0.0};
      for (int i = tid; i < M; i += T) {
        const double b3v = sB[i];
        const double b5n = const_lik ? y[i] * sw.lik.p0 : sA[i] * ainv5 - b3v;
        if (b5n != b5n) it_flag |= TRB_FLAG_NAN_B;
        const double b5v = damp(sw.damp5, __ldcg(cur.b5 + i), b5n);
        const double rnew = (b3v + b5v) / a_hat;  // base.py:152-161
        const double ro = __ldcg(cur.rz + i);
        rd[0] += (rnew - ro) * (rnew - ro);
        rd[1] += rnew * rnew;
        if (i >= cm0 && i < cm1) {
          nxt.b5[i] = b5v;
          nxt.rz[i] = rnew;
        }
        sA[i] = b5v;  // same thread wrote and read sA[i]: no barrier needed in between
      }
      block_sum_n<2>(rd, sh);
      dz2 = rd[0];
      nz2 = rd[1];
      if (a3n != a3n || a5n != a5n) it_flag |= TRB_FLAG_NAN_A;
      if (a3n < 0 || a5n < 0) it_flag |= TRB_FLAG_NEG_A;
      a[2] = a3;
      a[3] = a3;
      a[4] = a5;
      a[5] = a5;
      vz = 1. / a_hat;
    }
    __syncthreads();
    // ---- P3: tx = U_R^T b6
    project_rows(sw.Ut, ldm, M, r0, r1, sA, nxt.tx, red);
    sync_all();  // barrier 3: tx complete

    // ---- S2 + P4: backward variance, r_z = V_R coef
    const double vlin_b =
        rescale_all(1, R, N, M, sw.rank, null_space, sw.s, sw.s2, a[1], a[5], tz, nxt.tx, sC, sh);
    expand_cols(sw.Vt, ldn, R, cn0, cn1, sC, rzl, red);
    sync_all();  // barrier 4: rzl complete

    // ---- X: e7 (= e8), posterior of x, records, early stopping
    {
      const double a1 = a[1];
      const double a7n = clip_a_new(vlin_b, a1, sw.lin_amin, sw.lin_amax);  // base_channel.py:14-17
      const double a7 = damp(sw.damp7, a[6], a7n);
      const double ainv7 = a1 + a7n;
      const double a_hat = a1 + a7;
      double rd[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = tid; i < N; i += T) {
        const double b1v = __ldcg(nxt.b1 + i);
        double r = __ldcg(rzl + i);
        if (null_space) r = b1v / a1 + r;
        const double b7n = r * ainv7 - b1v;
        if (b7n != b7n) it_flag |= TRB_FLAG_NAN_B;
        const double b7v = damp(sw.damp7, __ldcg(cur.b7 + i), b7n);
        const double rnew = (b1v + b7v) / a_hat;
        const double ro = __ldcg(cur.rx + i);
        const double xv = xt ? xt[i] : 0.0;
        rd[0] += (rnew - ro) * (rnew - ro);
        rd[1] += rnew * rnew;
        rd[2] += (rnew - xv) * (rnew - xv);  // metrics.py:5-6
        rd[3] += (rnew + xv) * (rnew + xv);  // metrics.py:9-14
        sB[i] = b7v;
        if (i >= cn0 && i < cn1) {
          nxt.b7[i] = b7v;
          nxt.rx[i] = rnew;
        }
      }
      block_sum_n<4>(rd, sh);
      if (a7n != a7n) it_flag |= TRB_FLAG_NAN_A;
      if (a7n < 0) it_flag |= TRB_FLAG_NEG_A;
      a[6] = a7;
      a[7] = a7;
      vx = 1. / a_hat;
      const int all = block_or(it_flag, (int*)&red[0]);
      flag |= all;
      ++done;
      par ^= 1;
      // EarlyStoppingEP, callbacks.py:258-286 / EarlyStopping on the variances, :206-243
      double tol = nan("");
      int stop = 0;
      if (sw.es_mode == 1 && sw.es_tol >= 0) {
        const int vars = sw.es_vars ? sw.es_vars : 3;
        stop = early_stopping_variance(vars, it, vx, vz, vx_old, vz_old, sw.es_tol,
                                       sw.es_min_variance, sw.es_max_increase, sw.es_wait_increase,
                                       &tol);
      } else if (it > 0) {
        const double tol_x = sqrt(rd[0] / N) / sqrt(rd[1] / N);
        const double tol_z = sqrt(dz2 / M) / sqrt(nz2 / M);
        const int vars = sw.es_vars ? sw.es_vars : 3;
        tol = (vars & 1) ? tol_x : tol_z;
        if ((vars & 2) && tol_z > tol) tol = tol_z;
        if (sw.es_tol >= 0) {
          if (tol < sw.es_tol) {
            stop = TRB_FLAG_CONVERGED;
          } else if (it > sw.es_wait_increase && tol > sw.es_max_increase) {
            stop = TRB_FLAG_DIVERGED;
          }
        }
      }
      if (cta == 0 && tid == 0) {
        const bool rec = it < sw.max_records;
        if (rec && sw.rec_vx) sw.rec_vx[it] = vx;
        if (rec && sw.rec_vz) sw.rec_vz[it] = vz;
        if (xt) {
          const double mse = rd[2] / N, mse_neg = rd[3] / N;
          if (rec && sw.rec_mse) sw.rec_mse[it] = mse;
          if (rec && sw.rec_smse) sw.rec_smse[it] = fmin(mse, mse_neg);
        }
        if (rec && sw.rec_tol) sw.rec_tol[it] = tol;
      }
      const bool bad = all & (TRB_FLAG_NAN_A | TRB_FLAG_NAN_B);
      if (stop || bad) {
        stop_flags = stop;
        if (bad || stop == TRB_FLAG_DIVERGED) {
          // reset_message_dag(old_message_dag): message_passing.py:196-197, callbacks.py:281-283
          par ^= 1;
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] = a_old[e];
          vx = vx_old;
          vz = vz_old;
          rolled_back = true;
        }
        __syncthreads();
        break;
      }
    }
    __syncthreads();
  }

  // ---- epilogue: the final state into the live buffers, the other parity into the snapshot
  sync_all();
  if (par == 1) {
    const size_t gt = (size_t)cta * T + tid, GT = (size_t)G * T;
    auto swap = [&](double* x, double* yv, int n) {
      for (size_t i = gt; i < (size_t)n; i += GT) {
        const double u = __ldcg(x + i), w = __ldcg(yv + i);
        x[i] = w;
        yv[i] = u;
      }
    };
    swap(P[0].b1, P[1].b1, N);
    swap(P[0].b7, P[1].b7, N);
    swap(P[0].rx, P[1].rx, N);
    swap(P[0].b3, P[1].b3, M);
    swap(P[0].b5, P[1].b5, M);
    swap(P[0].rz, P[1].rz, M);
    swap(P[0].tx, P[1].tx, R);
  }
  if (cta == 0 && tid == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) sw.edge_a[e] = a[e];
    sw.vx[0] = vx;
    sw.vz[0] = vz;
    sw.n_iter[0] += done;
    int f = flag | stop_flags;
    if (rolled_back) f |= TRB_FLAG_RESTORED;
    if (f) atomicOr(&sw.flags[0], f);
    if (stop_flags || (flag & (TRB_FLAG_NAN_A | TRB_FLAG_NAN_B))) sw.active[0] = 0;
  }
}

int g_persistent_mode = -2;  // -2 unset (environment TRB_PERSISTENT_SWEEP, else auto), -1 auto, 0 off,
                             // 1 on whenever the hard limits allow, 2 / 3 = as 1 but always the
                             // cooperative-grid / the single-cluster variant

constexpr size_t kPsMaxSmem = 200 * 1024;
// operator bytes per iteration up to which one cluster wins: measured crossover between 3.1 MB
// (N = 512: 19.9 vs 26.2 us) and 7.1 MB (N = 768: 27.9 vs 26.1 us), profiles/r01f_persistent_sizes.json
constexpr double kClusterMaxBytes = 5e6;

}  // namespace

extern "C" void trb_set_persistent_sweep(int mode) { g_persistent_mode = mode; }

// Runs the sweep in the persistent kernel if the instance qualifies.  Returns
// TRB_ERR_UNSUPPORTED (without setting an error) when the caller should fall
// back to the launch-per-stage path.
int trb_sweep_run_persistent(const trb_sweep* sw, int it0, int n_iter, int fresh, cudaStream_t st) {
  if (g_persistent_mode == -2) {
    const char* e = getenv("TRB_PERSISTENT_SWEEP");
    g_persistent_mode = e ? atoi(e) : -1;
    if (g_persistent_mode < -1 || g_persistent_mode > 3) g_persistent_mode = -1;
  }
  const int mode = g_persistent_mode;
  if (mode == 0 || trb_profile_events_enabled()) return TRB_ERR_UNSUPPORTED;
  if (sw->B != 1 || sw->comm || !sw->snap_edge_a) return TRB_ERR_UNSUPPORTED;
  if (!sw->Vt || !sw->Ut || !sw->s || !sw->s2) return TRB_ERR_UNSUPPORTED;
  if (sw->gemv_impl != 0 && sw->gemv_impl != 2) return TRB_ERR_UNSUPPORTED;
  if (sw->R_total > 0 && sw->R_total != sw->R) return TRB_ERR_UNSUPPORTED;
  if (n_iter <= 0) return TRB_OK;
  const int ldmax = sw->ldn > sw->ldm ? sw->ldn : sw->ldm;
  const size_t smem = sizeof(double) * (2 * (size_t)ldmax + sw->R);
  // launch-bound regime only: beyond ~250 us of streaming per iteration the
  // launches of trb_sweep.cu are hidden and its TMA-ring GEMVs are faster
  const double bytes = 16.0 * sw->R * ((double)sw->N + sw->M);
  if (mode < 0 && (bytes > 1.5e9 || n_iter < 2)) return TRB_ERR_UNSUPPORTED;
  if (smem > kPsMaxSmem) return TRB_ERR_UNSUPPORTED;
  static int coop = -1, sms = 0, cluster_ok = 0;
  if (coop < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaFuncSetAttribute(k_sweep_persistent<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kPsMaxSmem) != cudaSuccess) {
      cudaGetLastError();
      coop = 0;
    }
    cluster_ok = cudaFuncSetAttribute(k_sweep_persistent<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kPsMaxSmem) == cudaSuccess &&
                 cudaFuncSetAttribute(k_sweep_persistent<true>,
                                      cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    cudaGetLastError();
  }
  trb_sweep desc = *sw;
  int ld = ldmax;
  const bool want_cluster = mode == 3 || (mode != 2 && bytes <= kClusterMaxBytes);
  if (want_cluster && cluster_ok && sw->R >= 16) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16, 1, 1);
    cfg.blockDim = dim3(kPsThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    trb_launch_scope scope_(0, st);
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k_sweep_persistent<true>, desc, it0, n_iter, fresh, ld);
    if (e == cudaSuccess) return TRB_OK;
    cudaGetLastError();
    cluster_ok = 0;  // this device refuses the 16-CTA cluster: use the grid variant from now on
    if (mode == 3)
      return trb_set_error(TRB_ERR_CUDA, "k_sweep_persistent<cluster>: %s", cudaGetErrorString(e));
  } else if (mode == 3) {
    return TRB_ERR_UNSUPPORTED;
  }
  if (!coop || sms <= 0) return TRB_ERR_UNSUPPORTED;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_persistent<false>, kPsThreads,
                                                    smem) != cudaSuccess ||
      per_sm < 1) {
    cudaGetLastError();
    return TRB_ERR_UNSUPPORTED;
  }
  int grid = sms;
  if (grid > sw->R) grid = sw->R;  // at least one singular index per CTA
  void* args[] = {&desc, &it0, &n_iter, &fresh, &ld};
  trb_launch_scope scope_(0, st);
  const cudaError_t e = cudaLaunchCooperativeKernel((void*)k_sweep_persistent<false>, dim3(grid),
                                                    dim3(kPsThreads), args, smem, st);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (mode > 0) return trb_set_error(TRB_ERR_CUDA, "k_sweep_persistent: %s", cudaGetErrorString(e));
    return TRB_ERR_UNSUPPORTED;
  }
  return TRB_OK;
}
