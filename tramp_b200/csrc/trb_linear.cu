// LinearChannel in thin-SVD form: batched, HBM-bound FP64 GEMVs.
//
// reference: channels/linear/linear_channel.py:69-89 (compute_backward_mean /
// compute_forward_mean) streams U, V (full), the dense S and W -- nine GEMVs
// per iteration.  Here each instance's operators are stored as rows of
// singular vectors (Vt[R, ldn], Ut[R, ldm]) and each is streamed ONCE per use:
//   project:  t[i]  = <A[i, :], vec>            (U.T @ bx, V.T @ bz; :72-73)
//   expand :  out[j] = sum_i coef[i] * A[i, j]   (V @ rz_svd, U-side of W @ rz; :78, :88)
// Both walk the SAME partition of the B*R global row space: worker (CTA) k owns
// the contiguous rows [k*T/G, (k+1)*T/G), so every SM streams an equal,
// contiguous slab of HBM regardless of how instances fall on SMs.
//
// Two implementations, selected by `impl`:
//   1  LDG : plain 16-byte streaming loads (no shared-memory staging)
//   2  TMA : producer warp issues cp.async.bulk (1-D TMA) into a shared-memory
//            ring guarded by mbarriers; 8 consumer warps own column slices
//   0  default = TMA when the shape allows it, else LDG
#include "trb_common.cuh"

using namespace trb;

namespace {

constexpr int kMinRowsPerCta = 8;

// ------------------------------------------------------------------ geometry
struct Segment {
  int b, i0, i1;
};

// Iterate the instance segments of worker range [g, g1); returns false when done.
__device__ __forceinline__ bool next_segment(int64_t& g, int64_t g1, int R, Segment& s) {
  if (g >= g1) return false;
  s.b = (int)(g / R);
  s.i0 = (int)(g - (int64_t)s.b * R);
  const int64_t room = g1 - g;
  s.i1 = (room < (int64_t)(R - s.i0)) ? (int)(s.i0 + room) : R;
  g += (s.i1 - s.i0);
  return true;
}

// =============================================================== LDG kernels
constexpr int kLdgThreads = 512;

// project: one warp per row, vec staged in shared memory per instance.
__global__ void __launch_bounds__(kLdgThreads)
k_project_ldg(const double* __restrict__ A, int64_t strideA, int R, int n, int ld, int B,
              const double* __restrict__ vec, int ldvec, double* __restrict__ t,
              const int* __restrict__ active) {
  extern __shared__ __align__(16) double sh_vec[];  // ld doubles
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int64_t T = (int64_t)B * R, G = gridDim.x;
  int64_t g = part_begin(blockIdx.x, T, G);
  const int64_t g1 = part_begin(blockIdx.x + 1, T, G);
  Segment s;
  const int npair = ld >> 1;
  while (next_segment(g, g1, R, s)) {
    if (active && !active[s.b]) continue;
    __syncthreads();  // previous segment done with sh_vec
    for (int j = threadIdx.x; j < ld; j += blockDim.x)
      sh_vec[j] = (j < n) ? vec[(size_t)s.b * ldvec + j] : 0.0;
    __syncthreads();
    const double* Ab = A + (size_t)s.b * strideA;
    const double2* xv = reinterpret_cast<const double2*>(sh_vec);
    for (int i = s.i0 + 2 * warp; i < s.i1; i += 2 * nwarp) {
      const bool two = (i + 1 < s.i1);
      const double* r0 = Ab + (size_t)i * ld;
      const double* r1 = two ? r0 + ld : r0;
      double acc0 = 0.0, acc1 = 0.0;
      int p = lane;
      for (; p + 96 < npair; p += 128) {
        double2 a0[4], a1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a0[u] = ldg_stream(r0 + 2 * (p + 32 * u));
          a1[u] = ldg_stream(r1 + 2 * (p + 32 * u));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double2 x = xv[p + 32 * u];
          acc0 = fma(a0[u].x, x.x, acc0);
          acc0 = fma(a0[u].y, x.y, acc0);
          acc1 = fma(a1[u].x, x.x, acc1);
          acc1 = fma(a1[u].y, x.y, acc1);
        }
      }
      for (; p < npair; p += 32) {
        const double2 a0 = ldg_stream(r0 + 2 * p);
        const double2 a1 = ldg_stream(r1 + 2 * p);
        const double2 x = xv[p];
        acc0 = fma(a0.x, x.x, acc0);
        acc0 = fma(a0.y, x.y, acc0);
        acc1 = fma(a1.x, x.x, acc1);
        acc1 = fma(a1.y, x.y, acc1);
      }
      acc0 = warp_sum(acc0);
      acc1 = warp_sum(acc1);
      if (lane == 0) {
        t[(size_t)s.b * R + i] = acc0;
        if (two) t[(size_t)s.b * R + i + 1] = acc1;
      }
    }
  }
}

// expand: thread owns NB column pairs for the whole segment; no reduction
// until the segment ends, then one plain store per column into the slot.
template <int NB>
__global__ void __launch_bounds__(kLdgThreads)
k_expand_ldg(const double* __restrict__ A, int64_t strideA, int R, int ld, int B,
             const double* __restrict__ coef, double* __restrict__ part, int nslots,
             const int* __restrict__ active) {
  constexpr int UR = (NB >= 8) ? 2 : (16 / NB > 8 ? 8 : 16 / NB);  // rows in flight
  const int64_t T = (int64_t)B * R, G = gridDim.x;
  int64_t g = part_begin(blockIdx.x, T, G);
  const int64_t g1 = part_begin(blockIdx.x + 1, T, G);
  const int npair = ld >> 1;
  Segment s;
  while (next_segment(g, g1, R, s)) {
    if (active && !active[s.b]) continue;
    const double* Ab = A + (size_t)s.b * strideA;
    const double* cb = coef + (size_t)s.b * R;
    double2 acc[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) acc[k] = make_double2(0.0, 0.0);
    int i = s.i0;
    for (; i + UR <= s.i1; i += UR) {
      double2 a[UR][NB];
      double c[UR];
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        c[u] = __ldg(cb + i + u);
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const int p = threadIdx.x + k * kLdgThreads;
          a[u][k] = (p < npair) ? ldg_stream(Ab + (size_t)(i + u) * ld + 2 * p)
                                : make_double2(0.0, 0.0);
        }
      }
#pragma unroll
      for (int u = 0; u < UR; ++u)
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          acc[k].x = fma(c[u], a[u][k].x, acc[k].x);
          acc[k].y = fma(c[u], a[u][k].y, acc[k].y);
        }
    }
    for (; i < s.i1; ++i) {
      const double c = __ldg(cb + i);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const int p = threadIdx.x + k * kLdgThreads;
        if (p < npair) {
          const double2 a = ldg_stream(Ab + (size_t)i * ld + 2 * p);
          acc[k].x = fma(c, a.x, acc[k].x);
          acc[k].y = fma(c, a.y, acc[k].y);
        }
      }
    }
    const int slot = (int)(blockIdx.x - part_owner((int64_t)s.b * R, T, G));
    double* out = part + ((size_t)s.b * nslots + slot) * ld;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const int p = threadIdx.x + k * kLdgThreads;
      if (p < npair) *reinterpret_cast<double2*>(out + 2 * p) = acc[k];
    }
  }
}

// =============================================================== TMA kernels
// 8 consumer warps (256 threads) + 1 producer warp.  Consumer thread t owns the
// column pairs {t + 256*k, k < NB}.  A ring stage holds RC = max(1, 8/NB)
// whole rows (<= 32 KiB, 64 KiB for NB = 16).
constexpr int kConsumers = 256;
constexpr int kTmaThreads = kConsumers + 32;
constexpr int kMaxStages = 8;
constexpr int kGroupRows = 8;  // project: rows reduced per block-level reduction

template <int NB>
struct TmaCfg {
  static constexpr int RC = (NB >= 8) ? 1 : 8 / NB;
};

__device__ __forceinline__ void consumer_bar() {
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
}

// ---- rescale in the singular basis (linear_channel.py:58-67 compute_n_eff, :74 resolvent,
// :91-105 variances), shared by k_lin_rescale and by the epilogue of the fused projection
struct RescalePlan {
  double az_v, ratio;
  int sum_mode;  // 1 = sum of the spectrum (:100-102, dir 0 and ax == 0), 2 = sum s2 / (ratio + s2)
                 // (:66-67), 0 = none (:60-65: ax == 0 or ratio == 0)
};

__device__ __forceinline__ RescalePlan rescale_plan(int dir, double az, double ax) {
  RescalePlan p;
  p.az_v = az;
  if (dir == 1) p.az_v = (az != az) ? az : fmax(1e-11, az);  // :94 np.maximum(1e-11, az)
  p.ratio = p.az_v / ax;
  p.sum_mode = (dir == 0 && ax == 0) ? 1 : ((ax == 0 || p.ratio == 0) ? 0 : 2);
  return p;
}

// coefficient of one singular direction (linear_channel.py:74 resolvent, :69-89 means)
__device__ __forceinline__ double rescale_coef(int dir, bool null_space, double az, double ax, double si,
                                               double s2i, double tzi, double txi) {
  const double res = 1 / (az + ax * s2i);  // :74
  if (dir == 0) return si * (res * (tzi + si * txi));
  if (!null_space) return res * (tzi + si * txi);
  // res - 1/az = -(ax*s2/az)*res, applied to tz; the bz/az term is added by the consumer of the
  // expansion
  return res * (si * txi - (ax * s2i / az) * tzi);
}

// One thread's share (i = first, first + stride, ...) of one pass over the spectrum: the
// coefficients in the singular basis AND the sum the variance needs (they do not depend on each
// other).  U elements at a time, all their loads issued before the first store (the stores would
// otherwise serialise the loads: one memory round trip per element).  CG: tz / tx were written
// by other CTAs of this launch -> read them from L2.
template <bool CG, int U>
__device__ __forceinline__ double rescale_share(int dir, int R, int rank, bool null_space,
                                                const RescalePlan& pl, const double* __restrict__ sb,
                                                const double* __restrict__ s2b, double az, double ax,
                                                const double* tz, const double* tx, double* coef,
                                                double* snap_tx, int first, int stride) {
  double part = 0.0;
  for (int i0 = first; i0 < R; i0 += stride * U) {
    double s2v[U], sv[U], tzv[U], txv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * stride;
      if (i < R) {
        s2v[u] = s2b[i];
        if (coef) {
          sv[u] = sb[i];
          tzv[u] = CG ? __ldcg(tz + i) : tz[i];
          txv[u] = CG ? __ldcg(tx + i) : tx[i];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * stride;
      if (i < R) {
        const double s2i = s2v[u];
        if (i < rank) {
          if (pl.sum_mode == 1) part += s2i;
          else if (pl.sum_mode == 2) part += s2i / (pl.ratio + s2i);
        }
        if (coef) {
          if (snap_tx) snap_tx[i] = txv[u];
          coef[i] = rescale_coef(dir, null_space, az, ax, sv[u], s2i, tzv[u], txv[u]);
        }
      }
    }
  }
  return part;
}

// total = the instance-wide sum of rescale_share (ignored when sum_mode == 0)
__device__ __forceinline__ double rescale_variance(int dir, int Nz, int Nx, int rank,
                                                   const RescalePlan& pl, double az, double ax,
                                                   double total) {
  if (pl.sum_mode == 1) {  // :100-102
    const double s_mean = total / rank;
    return s_mean * rank / (Nx * az);
  }
  double n_eff;
  if (ax == 0) {  // :60-62
    n_eff = 0.;
  } else if (pl.ratio == 0) {  // :63-65
    n_eff = (double)rank / Nz;
  } else {  // :66-67
    n_eff = total / Nz;
  }
  if (dir == 0) {
    const double alpha = (double)Nx / Nz;
    return n_eff / (alpha * ax);  // :103-105
  }
  return (1 - n_eff) / pl.az_v;  // :95-97
}

// The rescale stage (S1 / S2 of the sweep) inside the projection that feeds it (P1 / P3): the
// coefficient of row i depends on row i only -- its own projection, the other projection, s_i and
// the instance's (az, ax) -- so the thread that finishes the block reduction of a row writes the
// coefficient next to the projection, from operands it loaded while the row was streaming; the
// variance needs the spectrum only and is computed by the CTA that owns the instance's first row
// from values it loaded a segment earlier.  No launch, no wait on other CTAs -- measured: a
// dependent access inside a kernel that saturates HBM queues behind ~30 MB of ring traffic
// (~4 us), which is what made the "last-arriving CTA rescales the instance" epilogue lose.
struct RescaleFused {
  int dir, Nz, Nx, rank, null_space;
  const double* s;
  const double* s2;
  int64_t stride_s;
  const double* az;       // [B]
  const double* ax;       // [B]
  const double* t_other;  // the projection this launch does not write (dir 0: tx, dir 1: tz), [B, R]
  double* coef;           // [B, R]
  double* v_out;          // [B]
  double* snap_tx;        // nullable, [B, R]: receives a copy of tx (dir 0 only)
};

constexpr int kVarRegs = 8;  // spectrum values per thread kept in flight for the next instance's variance

// MODE: 0 project, 1 expand, 2 project with the rescale inside (coefficients and variance)
template <int NB, int MODE>
__global__ void __launch_bounds__(kTmaThreads, 1)
k_gemv_tma(const double* __restrict__ A, int64_t strideA, int R, int n, int ld, int B,
           const double* __restrict__ vec, int ldvec,  // project: input vector; expand: coef [B,R]
           double* __restrict__ out, int nslots,       // project: t [B,R]; expand: part
           const int* __restrict__ active, int nstages, int stage_doubles,
           int row_stride,   // doubles between consecutive rows of A (>= ld)
           int out_ld,       // expand: leading dimension of `part`
           int npanels,      // > 1: A is npanels column panels of `ld` doubles (the last one may be
                             // narrower: full_ld, full_n); every CTA walks its rows once per panel
           int full_ld, int full_n, RescaleFused rf) {
  constexpr bool EXPAND = MODE == 1;
  constexpr int RC = TmaCfg<NB>::RC;
  extern __shared__ __align__(128) double ring[];  // nstages * stage_doubles
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ double red[2][kConsumers / 32][kGroupRows];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumers / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int64_t T = (int64_t)B * R, G = gridDim.x;
  const int64_t g0 = part_begin(blockIdx.x, T, G);
  const int64_t g1 = part_begin(blockIdx.x + 1, T, G);
  const int panel_ld = ld;
  Segment s;
  int stage = 0;
  uint32_t phase = 0;

  if (warp == kConsumers / 32) {
    // ------------------------------------------------------------ producer
    if (lane == 0) {
      for (int pq = 0; pq < npanels; ++pq) {
        const int c0 = pq * panel_ld;
        const int ldq = (npanels > 1 && full_ld - c0 < panel_ld) ? full_ld - c0 : panel_ld;
        int64_t g = g0;
        while (next_segment(g, g1, R, s)) {
          if (active && !active[s.b]) continue;
          const double* Ab = A + (size_t)s.b * strideA + c0;
          for (int i = s.i0; i < s.i1; i += RC) {
            const int rows = (s.i1 - i < RC) ? (s.i1 - i) : RC;
            const uint32_t bytes = (uint32_t)rows * (uint32_t)ldq * 8u;
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full_bar[stage], bytes);
            bulk_g2s(ring + (size_t)stage * stage_doubles, Ab + (size_t)i * row_stride, bytes,
                     &full_bar[stage]);
            if (++stage == nstages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  int red_buf = 0;
  // MODE 2 (see RescaleFused): the spectrum of the next instance, in flight while rows stream
  [[maybe_unused]] double nv_s2[kVarRegs], nv_az = 0.0, nv_ax = 0.0;
  [[maybe_unused]] int nv_b = -1;
  [[maybe_unused]] auto var_issue = [&](int bn) {
    nv_b = bn;
    nv_az = rf.az[bn];
    nv_ax = rf.ax[bn];
#pragma unroll
    for (int u = 0; u < kVarRegs; ++u) {
      const int i = tid + u * kConsumers;
      nv_s2[u] = (i < R) ? rf.s2[(size_t)bn * rf.stride_s + i] : 0.0;
    }
  };
  for (int pq = 0; pq < npanels; ++pq) {
  const int c0 = pq * panel_ld;
  if (npanels > 1) {  // this panel's extent
    ld = (full_ld - c0 < panel_ld) ? full_ld - c0 : panel_ld;
    n = (full_n - c0 < ld) ? full_n - c0 : ld;
    if (n < 0) n = 0;
  }
  const int npair = ld >> 1;
  const int accumulate = pq > 0;  // project: later panels add to t
  int64_t g = g0;
  while (next_segment(g, g1, R, s)) {
    if (active && !active[s.b]) continue;
    if constexpr (!EXPAND) {
      [[maybe_unused]] double seg_az = 0.0, seg_ax = 0.0;
      if constexpr (MODE == 2) {
        seg_az = rf.az[s.b];
        seg_ax = rf.ax[s.b];
        // after this segment the CTA goes on with row 0 of the next instance, if its range goes on
        const bool more = (s.i1 == R) && (g < g1);
        if (s.i0 == 0 && pq == 0) {
          // ---- this CTA owns the instance's first row: its variance (spectrum only), from the
          // values loaded during the previous segment (the CTA's first segment loads them here)
          if (nv_b != s.b) var_issue(s.b);
          const RescalePlan pl = rescale_plan(rf.dir, nv_az, nv_ax);
          double part = 0.0;
#pragma unroll
          for (int u = 0; u < kVarRegs; ++u) {
            const int i = tid + u * kConsumers;
            if (i < R && i < rf.rank) {
              if (pl.sum_mode == 1) part += nv_s2[u];
              else if (pl.sum_mode == 2) part += nv_s2[u] / (pl.ratio + nv_s2[u]);
            }
          }
          for (int i = tid + kVarRegs * kConsumers; i < R && i < rf.rank; i += kConsumers) {
            const double s2i = rf.s2[(size_t)s.b * rf.stride_s + i];
            if (pl.sum_mode == 1) part += s2i;
            else if (pl.sum_mode == 2) part += s2i / (pl.ratio + s2i);
          }
          part = warp_sum(part);
          consumer_bar();  // red is free: the previous segment's last group has been read
          if (lane == 0) red[red_buf][warp][0] = part;
          consumer_bar();
          if (tid == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < kConsumers / 32; ++w) tot += red[red_buf][w][0];
            rf.v_out[s.b] = rescale_variance(rf.dir, rf.Nz, rf.Nx, rf.rank, pl, nv_az, nv_ax, tot);
          }
          red_buf ^= 1;
        }
        if (more && pq == 0) var_issue(s.b + 1);
      }
      // ---- project: x slice in registers, 8-row groups, block reduction
      double2 x[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const int c = 2 * (tid + k * kConsumers);
        const double* vp = vec + (size_t)s.b * ldvec + c0;
        x[k].x = (c < n) ? vp[c] : 0.0;
        x[k].y = (c + 1 < n) ? vp[c + 1] : 0.0;
      }
      for (int ig = s.i0; ig < s.i1; ig += kGroupRows) {
        double acc[kGroupRows];
#pragma unroll
        for (int q = 0; q < kGroupRows; ++q) acc[q] = 0.0;
        // MODE 2: what the coefficient of this thread's row needs besides the projection, loaded
        // while the group's rows stream
        [[maybe_unused]] double pf_s = 0.0, pf_s2 = 0.0, pf_other = 0.0;
        if constexpr (MODE == 2) {
          if (pq == npanels - 1 && tid < kGroupRows && ig + tid < s.i1) {
            pf_s = rf.s[(size_t)s.b * rf.stride_s + ig + tid];
            pf_s2 = rf.s2[(size_t)s.b * rf.stride_s + ig + tid];
            pf_other = rf.t_other[(size_t)s.b * R + ig + tid];
          }
        }
#pragma unroll
        for (int c = 0; c < kGroupRows / RC; ++c) {
          const int i = ig + c * RC;
          if (i < s.i1) {
            const int rows = (s.i1 - i < RC) ? (s.i1 - i) : RC;
            mbar_wait(&full_bar[stage], phase);
            const double2* st =
                reinterpret_cast<const double2*>(ring + (size_t)stage * stage_doubles);
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
              if (rr < rows) {
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                  const int p = tid + k * kConsumers;
                  if (p < npair) {
                    const double2 a = st[rr * npair + p];
                    acc[c * RC + rr] = fma(a.x, x[k].x, acc[c * RC + rr]);
                    acc[c * RC + rr] = fma(a.y, x[k].y, acc[c * RC + rr]);
                  }
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == nstages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
#pragma unroll
        for (int q = 0; q < kGroupRows; ++q) acc[q] = warp_sum(acc[q]);
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < kGroupRows; ++q) red[red_buf][warp][q] = acc[q];
        }
        consumer_bar();
        if (tid < kGroupRows && ig + tid < s.i1) {
          double tot = 0.0;
#pragma unroll
          for (int w = 0; w < kConsumers / 32; ++w) tot += red[red_buf][w][tid];
          double* dst = out + (size_t)s.b * R + ig + tid;
          if (accumulate) tot = *dst + tot;
          *dst = tot;
          if constexpr (MODE == 2) {
            if (pq == npanels - 1) {  // the projection of this row is complete: its coefficient
              const size_t o = (size_t)s.b * R + ig + tid;
              if (rf.snap_tx) rf.snap_tx[o] = pf_other;  // dir 0: t_other is tx
              rf.coef[o] = rescale_coef(rf.dir, rf.null_space != 0, seg_az, seg_ax, pf_s, pf_s2,
                                        rf.dir == 0 ? tot : pf_other, rf.dir == 0 ? pf_other : tot);
            }
          }
        }
        red_buf ^= 1;
      }
    } else {
      // ---- expand: column accumulators in registers for the whole segment
      double2 acc[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) acc[k] = make_double2(0.0, 0.0);
      const double* cb = vec + (size_t)s.b * R;
      for (int i = s.i0; i < s.i1; i += RC) {
        const int rows = (s.i1 - i < RC) ? (s.i1 - i) : RC;
        double c[RC];
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) c[rr] = (rr < rows) ? __ldg(cb + i + rr) : 0.0;
        mbar_wait(&full_bar[stage], phase);
        const double2* st = reinterpret_cast<const double2*>(ring + (size_t)stage * stage_doubles);
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) {
          if (rr < rows) {
#pragma unroll
            for (int k = 0; k < NB; ++k) {
              const int p = tid + k * kConsumers;
              if (p < npair) {
                const double2 a = st[rr * npair + p];
                acc[k].x = fma(c[rr], a.x, acc[k].x);
                acc[k].y = fma(c[rr], a.y, acc[k].y);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      const int slot = (int)(blockIdx.x - part_owner((int64_t)s.b * R, T, G));
      double* o = out + ((size_t)s.b * nslots + slot) * out_ld + c0;
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const int p = tid + k * kConsumers;
        if (p < npair) *reinterpret_cast<double2*>(o + 2 * p) = acc[k];
      }
    }
  }
  }  // panels
}

// ---------------------------------------------------------------- small ops
__global__ void __launch_bounds__(256)
k_reduce_slots(int R, int n, int ld, int B, int G, int nslots, const double* part,
               const double* __restrict__ add, const double* __restrict__ add_div,
               double* out, size_t out_stride,  // out may alias slot 0 of part (in-place)
               trb_push push) {                 // push.n > 0: store to every rank's buffer instead
  const int b = blockIdx.y;  // grid (chunks, B)
  const int64_t T = (int64_t)B * R;
  const int kf = (int)part_owner((int64_t)b * R, T, G);
  const int kl = (int)part_owner((int64_t)b * R + R - 1, T, G);
  const int ns = kl - kf + 1;
  const double div = add ? add_div[b] : 1.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    double v = 0.0;
    const double* pj = part + (size_t)b * nslots * ld + j;
    int sl = 0;
    for (; sl + 8 <= ns; sl += 8) {  // eight loads in flight, added in slot order
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = pj[(size_t)(sl + u) * ld];
#pragma unroll
      for (int u = 0; u < 8; ++u) v += t[u];
    }
    for (; sl < ns; ++sl) v += pj[(size_t)sl * ld];
    if (add) v = add[(size_t)b * ld + j] / div + v;
    if (push.n > 0) {
      for (int r = 0; r < push.n; ++r) push.dst[r][(size_t)b * out_stride + j] = v;
    } else {
      out[(size_t)b * out_stride + j] = v;
    }
  }
  if (push.n > 0 && push.counter) {
    // the last CTA to finish publishes the exchange to every rank (trb_comm.cu):
    // CTA barrier, then one system-scope fence (cumulative over the pushes the barrier
    // ordered before it), then the arrival count
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned int total = gridDim.x * gridDim.y;
      if (atomicAdd(push.counter, 1u) == total - 1) {
        *push.counter = 0;
        __threadfence_system();
        for (int r = 0; r < push.n; ++r)
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(push.flags_of[r] + push.rank),
                       "l"(push.seq)
                       : "memory");
      }
    }
  }
}

// linear_channel.py:58-67 (compute_n_eff), :74 (resolvent), :91-105 (variances)
__global__ void __launch_bounds__(1024)
k_lin_rescale(int dir, int R, int Nz, int Nx, int rank, int null_space_flag,
              const double* __restrict__ s,
              const double* __restrict__ s2, int64_t stride_s, const double* __restrict__ az_arr,
              const double* __restrict__ ax_arr, const double* __restrict__ tz,
              const double* __restrict__ tx, double* __restrict__ coef, double* __restrict__ v_out,
              const int* __restrict__ active,
              double* __restrict__ snap_tx) {  // nullable: receives a copy of tx (one-iteration-back state)
  __shared__ double sh[33];
  const int b = blockIdx.y;  // grid (C, B): a cluster of C CTAs per instance
  if (active && !active[b]) return;
  const int T_ = blockDim.x * cluster_nctarank();
  const int gtid = cluster_ctarank() * blockDim.x + threadIdx.x;
  const double az = az_arr[b], ax = ax_arr[b];
  const double* sb = s + (size_t)b * stride_s;
  const double* s2b = s2 + (size_t)b * stride_s;
  const size_t off = (size_t)b * R;
  const RescalePlan pl = rescale_plan(dir, az, ax);
  const double part = rescale_share<false, 4>(dir, R, rank, null_space_flag != 0, pl, sb, s2b, az, ax,
                                           tz ? tz + off : nullptr, tx ? tx + off : nullptr,
                                           coef ? coef + off : nullptr, snap_tx ? snap_tx + off : nullptr,
                                           gtid, T_);
  if (!v_out) return;
  const double total = pl.sum_mode ? cluster_sum(part, sh) : 0.0;
  const double v = rescale_variance(dir, Nz, Nx, rank, pl, az, ax, total);
  if (gtid == 0) v_out[b] = v;
}

int pick_nb(int ld, int threads, int max_nb) {
  const int npair = ld >> 1;
  int nb = 1;
  while (nb * threads < npair) nb <<= 1;
  return (nb <= max_nb) ? nb : -1;
}

struct TmaPlan {
  int nb, stages, stage_doubles;
  size_t smem;
};

bool plan_tma(int ld, TmaPlan& p) {
  p.nb = pick_nb(ld, kConsumers, 16);
  if (p.nb < 0) return false;
  const int rc = (p.nb >= 8) ? 1 : 8 / p.nb;
  p.stage_doubles = ((rc * ld + 15) / 16) * 16;  // keep every stage 128-byte aligned
  const size_t budget = 200 * 1024;
  int st = (int)(budget / ((size_t)p.stage_doubles * 8));
  if (st > kMaxStages) st = kMaxStages;
  if (st < 2) return false;
  p.stages = st;
  p.smem = (size_t)st * p.stage_doubles * 8;
  return true;
}

template <int NB, int MODE>
int launch_tma(const TmaPlan& p, int G, const double* A, int64_t strideA, int R, int n, int ld,
               int B, const double* vec, int ldvec, double* out, int nslots, const int* active,
               cudaStream_t st, int row_stride, int out_ld, int npanels, int full_ld, int full_n,
               const RescaleFused& rf) {
  auto kern = k_gemv_tma<NB, MODE>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess)
      return trb_set_error(TRB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  kern<<<G, kTmaThreads, p.smem, st>>>(A, strideA, R, n, ld, B, vec, ldvec, out, nslots, active,
                                       p.stages, p.stage_doubles, row_stride, out_ld, npanels, full_ld,
                                       full_n, rf);
  return TRB_OK;
}

template <int MODE>
int dispatch_tma(const TmaPlan& p, int G, const double* A, int64_t strideA, int R, int n, int ld,
                 int B, const double* vec, int ldvec, double* out, int nslots, const int* active,
                 cudaStream_t st, int row_stride, int out_ld, int npanels, int full_ld, int full_n,
                 const RescaleFused& rf = RescaleFused()) {
#define TRB_TMA_CASE(NB_)                                                                        \
  case NB_:                                                                                      \
    return launch_tma<NB_, MODE>(p, G, A, strideA, R, n, ld, B, vec, ldvec, out, nslots, active, \
                                 st, row_stride, out_ld, npanels, full_ld, full_n, rf);
  switch (p.nb) {
    TRB_TMA_CASE(1)
    TRB_TMA_CASE(2)
    TRB_TMA_CASE(4)
    TRB_TMA_CASE(8)
    TRB_TMA_CASE(16)
  }
#undef TRB_TMA_CASE
  return trb_set_error(TRB_ERR_UNSUPPORTED, "no TMA GEMV instantiation for NB=%d", p.nb);
}

}  // namespace

// Operators wider than the TMA ring allows (ld > 8192 doubles) are processed as
// column panels of equal width in (4096, 8192]: every panel then has NB = 16 and
// one row per ring stage, so a stage is still one contiguous bulk copy.
struct Panels {
  int count, width;
};
static Panels plan_panels(int ld) {
  Panels p;
  const int kMax = 16 * kConsumers * 2;  // 8192
  p.count = (ld + kMax - 1) / kMax;
  p.width = ((ld + p.count - 1) / p.count + 1) & ~1;
  return p;
}

trb_expand_geom trb_expand_geometry(int B, int R) {
  trb_expand_geom g;
  const int64_t T = (int64_t)B * R;
  int sm = trb_sm_count_cached();
  if (sm <= 0) sm = 1;
  int64_t G = sm;
  const int64_t cap = T / kMinRowsPerCta;
  if (G > cap) G = cap;
  if (G < 1) G = 1;
  g.G = (int)G;
  int ns = 1;
  for (int b = 0; b < B; ++b) {
    const int kf = (int)part_owner((int64_t)b * R, T, G);
    const int kl = (int)part_owner((int64_t)b * R + R - 1, T, G);
    if (kl - kf + 1 > ns) ns = kl - kf + 1;
  }
  g.nslots = ns;
  return g;
}

extern "C" int trb_lin_expand_slots(int B, int R) {
  if (B <= 0 || R <= 0) return trb_set_error(TRB_ERR_INVALID, "trb_lin_expand_slots: bad shape");
  return trb_expand_geometry(B, R).nslots;
}

static int check_gemv_args(const double* A, int R, int n, int ld, int B, const void* x,
                           const void* y) {
  TRB_CHECK_ARG(A && x && y, "null pointer");
  TRB_CHECK_ARG(B > 0 && R > 0 && n > 0 && ld >= n, "bad shape");
  TRB_CHECK_ARG(ld % 2 == 0, "ld must be even (16-byte rows)");
  TRB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0, "operator must be 16-byte aligned");
  return TRB_OK;
}

extern "C" int trb_lin_project(const double* A, int64_t strideA, int R, int n, int ld, int B,
                               const double* vec, int ldvec, double* t, const int* active,
                               int impl, void* stream) {
  int rc = check_gemv_args(A, R, n, ld, B, vec, t);
  if (rc) return rc;
  TRB_CHECK_ARG(strideA % 2 == 0, "strideA must be even");
  TRB_CHECK_ARG(ldvec >= n, "ldvec < n");
  cudaStream_t st = (cudaStream_t)stream;
  const trb_expand_geom geo = trb_expand_geometry(B, R);
  TmaPlan p;
  const bool tma_ok = plan_tma(ld, p);
  if (impl == 0) impl = 2;
  trb_launch_scope scope_(1, st);
  if (impl == 2 && !tma_ok) {
    // rows wider than one ring stage: one launch, every CTA walks its rows once per column panel
    const Panels pan = plan_panels(ld);
    TmaPlan pq;
    if (!plan_tma(pan.width, pq) || pq.nb < 8)
      return trb_set_error(TRB_ERR_UNSUPPORTED, "trb_lin_project: cannot panel ld=%d", ld);
    rc = dispatch_tma<0>(pq, geo.G, A, strideA, R, n, pan.width, B, vec, ldvec, t, 0, active, st,
                             ld, 0, pan.count, ld, n);
    if (rc) return rc;
  } else if (impl == 2) {
    rc = dispatch_tma<0>(p, geo.G, A, strideA, R, n, ld, B, vec, ldvec, t, 0, active, st, ld, 0,
                             1, ld, n);
    if (rc) return rc;
  } else {
    const size_t smem = (size_t)ld * 8;
    if (smem > 200 * 1024)
      return trb_set_error(TRB_ERR_UNSUPPORTED, "trb_lin_project: ld=%d too large", ld);
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(k_project_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      configured = true;
    }
    k_project_ldg<<<geo.G, kLdgThreads, smem, st>>>(A, strideA, R, n, ld, B, vec, ldvec, t, active);
  }
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

// trb_lin_project followed by trb_lin_rescale in ONE launch (sweep stages P1+S1 and P3+S2, see
// RescaleFused).  dir 0: t_out is tz and t_other tx; dir 1: t_out is tx and t_other tz.  snap_tx
// (nullable, dir 0) receives a copy of tx.  TMA GEMV only: TRB_ERR_UNSUPPORTED (nothing launched)
// when the shape has no TMA plan.
int trb_lin_project_rescale(const double* A, int64_t strideA, int R, int n, int ld, int B,
                            const double* vec, int ldvec, double* t_out, const int* active, int dir,
                            int Nz, int Nx, int rank, int null_space, const double* s, const double* s2,
                            int64_t stride_s, const double* az, const double* ax, const double* t_other,
                            double* coef, double* v, double* snap_tx, void* stream) {
  int rc = check_gemv_args(A, R, n, ld, B, vec, t_out);
  if (rc) return rc;
  TRB_CHECK_ARG(strideA % 2 == 0, "strideA must be even");
  TRB_CHECK_ARG(ldvec >= n, "ldvec < n");
  TRB_CHECK_ARG(s && s2 && az && ax && t_other && coef && v, "null pointer");
  TRB_CHECK_ARG(dir == 0 || dir == 1, "dir must be 0 or 1");
  TRB_CHECK_ARG(dir == 0 || !snap_tx, "snap_tx is the copy of tx taken by the forward rescale");
  TRB_CHECK_ARG(R <= Nz && R <= Nx && rank >= 0 && rank <= R, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const trb_expand_geom geo = trb_expand_geometry(B, R);
  RescaleFused rf;
  rf.dir = dir;
  rf.Nz = Nz;
  rf.Nx = Nx;
  rf.rank = rank;
  rf.null_space = null_space;
  rf.s = s;
  rf.s2 = s2;
  rf.stride_s = stride_s;
  rf.az = az;
  rf.ax = ax;
  rf.t_other = t_other;
  rf.coef = coef;
  rf.v_out = v;
  rf.snap_tx = snap_tx;
  TmaPlan p, pq;
  const bool one_panel = plan_tma(ld, p);
  const Panels pan = plan_panels(ld);
  if (!one_panel && (!plan_tma(pan.width, pq) || pq.nb < 8)) return TRB_ERR_UNSUPPORTED;
  trb_launch_scope scope_(1, st);
  if (one_panel) {
    rc = dispatch_tma<2>(p, geo.G, A, strideA, R, n, ld, B, vec, ldvec, t_out, 0, active, st, ld, 0, 1, ld,
                         n, rf);
  } else {
    rc = dispatch_tma<2>(pq, geo.G, A, strideA, R, n, pan.width, B, vec, ldvec, t_out, 0, active, st, ld, 0,
                         pan.count, ld, n, rf);
  }
  if (rc) return rc;
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_lin_expand(const double* A, int64_t strideA, int R, int n, int ld, int B,
                              const double* coef, double* part, const int* active, int impl,
                              void* stream) {
  int rc = check_gemv_args(A, R, n, ld, B, coef, part);
  if (rc) return rc;
  TRB_CHECK_ARG(strideA % 2 == 0, "strideA must be even");
  cudaStream_t st = (cudaStream_t)stream;
  const trb_expand_geom geo = trb_expand_geometry(B, R);
  TmaPlan p;
  const bool tma_ok = plan_tma(ld, p);
  if (impl == 0) impl = 2;
  trb_launch_scope scope_(1, st);
  if (impl == 2 && !tma_ok) {
    const Panels pan = plan_panels(ld);
    TmaPlan pq;
    if (!plan_tma(pan.width, pq) || pq.nb < 8)
      return trb_set_error(TRB_ERR_UNSUPPORTED, "trb_lin_expand: cannot panel ld=%d", ld);
    rc = dispatch_tma<1>(pq, geo.G, A, strideA, R, n, pan.width, B, coef, 0, part, geo.nslots,
                            active, st, ld, ld, pan.count, ld, ld);
    if (rc) return rc;
  } else if (impl == 2) {
    rc = dispatch_tma<1>(p, geo.G, A, strideA, R, n, ld, B, coef, 0, part, geo.nslots, active, st,
                            ld, ld, 1, ld, n);
    if (rc) return rc;
  } else {
    const int nb = pick_nb(ld, kLdgThreads, 8);
    switch (nb) {
      case 1: k_expand_ldg<1><<<geo.G, kLdgThreads, 0, st>>>(A, strideA, R, ld, B, coef, part, geo.nslots, active); break;
      case 2: k_expand_ldg<2><<<geo.G, kLdgThreads, 0, st>>>(A, strideA, R, ld, B, coef, part, geo.nslots, active); break;
      case 4: k_expand_ldg<4><<<geo.G, kLdgThreads, 0, st>>>(A, strideA, R, ld, B, coef, part, geo.nslots, active); break;
      case 8: k_expand_ldg<8><<<geo.G, kLdgThreads, 0, st>>>(A, strideA, R, ld, B, coef, part, geo.nslots, active); break;
      default:
        return trb_set_error(TRB_ERR_UNSUPPORTED, "trb_lin_expand: ld=%d too large", ld);
    }
  }
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

static int reduce_slots_launch(int B, int R, int n, int ld, const double* part, const double* add,
                               const double* add_div, double* out, size_t out_stride,
                               cudaStream_t st, const trb_push* push = nullptr) {
  const trb_expand_geom geo = trb_expand_geometry(B, R);
  trb_launch_scope scope_(0, st);
  int chunks = (n + 255) / 256;  // one column per thread while that still fits ~8 CTAs per SM
  const int cap = (8 * trb_sm_count_cached() + B - 1) / B;
  if (chunks > cap) chunks = cap < 1 ? 1 : cap;
  trb_push none;
  none.n = 0;
  k_reduce_slots<<<dim3(chunks, B), 256, 0, st>>>(R, n, ld, B, geo.G, geo.nslots, part, add, add_div,
                                                 out, out_stride, push ? *push : none);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

// Sums the per-CTA slots of every instance into slot 0 (used by the sweep when an
// instance spans many CTAs, so that the update kernels read one vector, not nslots).
int trb_reduce_slots_inplace(int B, int R, int n, int ld, double* part, void* stream) {
  const trb_expand_geom geo = trb_expand_geometry(B, R);
  return reduce_slots_launch(B, R, n, ld, part, nullptr, nullptr, part, (size_t)geo.nslots * ld,
                             (cudaStream_t)stream);
}

// the sum of the slots of instance b goes to push->dst[r][b, :] (leading dimension ld) for every rank r
int trb_reduce_slots_push(int B, int R, int n, int ld, const double* part, const trb_push* push,
                          void* stream) {
  return reduce_slots_launch(B, R, n, ld, part, nullptr, nullptr, nullptr, (size_t)ld,
                             (cudaStream_t)stream, push);
}

extern "C" int trb_lin_reduce_slots(int B, int R, int n, int ld, const double* part,
                                    const double* add, const double* add_div, double* out,
                                    void* stream) {
  TRB_CHECK_ARG(part && out, "null pointer");
  TRB_CHECK_ARG(!add || add_div, "add needs add_div");
  TRB_CHECK_ARG(B > 0 && R > 0 && n > 0 && ld >= n, "bad shape");
  return reduce_slots_launch(B, R, n, ld, part, add, add_div, out, (size_t)ld, (cudaStream_t)stream);
}

int trb_lin_rescale_snap(int dir, int B, int R, int Nz, int Nx, int rank, int null_space,
                         const double* s, const double* s2, int64_t stride_s, const double* az,
                         const double* ax, const double* tz, const double* tx, double* coef, double* v,
                         const int* active, double* snap_tx, void* stream);

extern "C" int trb_lin_rescale(int dir, int B, int R, int Nz, int Nx, int rank, int null_space,
                               const double* s, const double* s2, int64_t stride_s,
                               const double* az, const double* ax, const double* tz,
                               const double* tx, double* coef, double* v, const int* active,
                               void* stream) {
  return trb_lin_rescale_snap(dir, B, R, Nz, Nx, rank, null_space, s, s2, stride_s, az, ax, tz, tx, coef, v,
                              active, nullptr, stream);
}

int trb_lin_rescale_snap(int dir, int B, int R, int Nz, int Nx, int rank, int null_space,
                         const double* s, const double* s2, int64_t stride_s, const double* az,
                         const double* ax, const double* tz, const double* tx, double* coef, double* v,
                         const int* active, double* snap_tx, void* stream) {
  TRB_CHECK_ARG(s && s2 && az && ax && (coef || v), "null pointer");
  TRB_CHECK_ARG(!coef || (tz && tx), "coef needs tz and tx");
  TRB_CHECK_ARG(dir == 0 || dir == 1, "dir must be 0 or 1");
  TRB_CHECK_ARG(B > 0 && R > 0 && R <= Nz && R <= Nx && rank >= 0 && rank <= R, "bad shape");
  trb_launch_scope scope_(0, (cudaStream_t)stream);
  // division-heavy FP64 per element: a large single spectrum gets the widest CTAs
  const int rs_threads = (R >= 8192 && B * 8 <= trb_sm_count_cached()) ? 1024 : 256;
  cudaError_t le = trb_launch_cluster(k_lin_rescale, trb_cluster_size(B, R), B, rs_threads,
                                      (cudaStream_t)stream, dir, R, Nz, Nx, rank, null_space, s, s2,
                                      stride_s, az, ax, tz, tx, coef, v, active, coef ? snap_tx : nullptr);
  if (le != cudaSuccess)
    return trb_set_error(TRB_ERR_CUDA, "trb_lin_rescale: %s", cudaGetErrorString(le));
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}
