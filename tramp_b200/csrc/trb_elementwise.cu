// Elementwise moment kernels: separable priors / likelihoods, truncated
// normal, Variable.posterior_rv.  One CTA per instance; FP64; coalesced
// double2 where alignment allows.  See include/tramp_b200.h for the contract.
#include <stdarg.h>
#include "trb_moments.cuh"

using namespace trb;

thread_local char trb_err_buf[512] = "";

int trb_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(trb_err_buf, sizeof(trb_err_buf), fmt, ap);
  va_end(ap);
  return code;
}

int trb_sm_count_cached() {
  static int sm = -1;
  if (sm < 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    sm = n;
  }
  return sm;
}

// ---- launch accounting -------------------------------------------------------
#include <vector>
namespace {
struct ProfileState {
  bool enabled = false;
  long long launches[3] = {0, 0, 0};  // update kernels, operator passes, set-up (trb_setup.cu)
  std::vector<cudaEvent_t> pool;         // recycled events
  std::vector<cudaEvent_t> begin[2], end[2];
  std::vector<int> order;                // kinds of the timed launches (0 / 1), in launch order
};
ProfileState g_prof;
cudaEvent_t prof_event() {
  if (!g_prof.pool.empty()) {
    cudaEvent_t e = g_prof.pool.back();
    g_prof.pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void trb_note_launch(int kind, cudaStream_t st, bool before) {
  if (before) g_prof.launches[kind] += 1;
  if (!g_prof.enabled || kind > 1) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, st);
  (before ? g_prof.begin[kind] : g_prof.end[kind]).push_back(e);
  if (before) g_prof.order.push_back(kind);
}

bool trb_profile_events_enabled() { return g_prof.enabled; }
void trb_profile_add_launches(long long n0, long long n1) {
  g_prof.launches[0] += n0;
  g_prof.launches[1] += n1;
}
long long trb_profile_launch_count(int kind) { return g_prof.launches[kind]; }

extern "C" void trb_profile_reset(int enable_events) {
  for (int k = 0; k < 2; ++k) {
    g_prof.launches[k] = 0;
    g_prof.launches[2] = 0;
    for (auto e : g_prof.begin[k]) g_prof.pool.push_back(e);
    for (auto e : g_prof.end[k]) g_prof.pool.push_back(e);
    g_prof.begin[k].clear();
    g_prof.end[k].clear();
  }
  g_prof.order.clear();
  g_prof.enabled = enable_events != 0;
}

extern "C" long long trb_profile_launches(int kind) {
  return (kind >= 0 && kind <= 2) ? g_prof.launches[kind] : g_prof.launches[0] + g_prof.launches[1];
}

// Sum of the CUDA-event durations of the GEMV launches recorded since the last
// reset; waits for the last one to finish.  Returns the number of timed launches.
extern "C" int trb_profile_gemv_ms(double* total_ms) {
  double tot = 0.0;
  const size_t n = g_prof.end[1].size();
  if (n) cudaEventSynchronize(g_prof.end[1][n - 1]);
  for (size_t i = 0; i < n && i < g_prof.begin[1].size(); ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof.begin[1][i], g_prof.end[1][i]) == cudaSuccess) tot += ms;
  }
  if (total_ms) *total_ms = tot;
  return (int)n;
}

// The event-timed launches since the last reset, in launch order: ms[i] = duration, kinds[i] = 0
// (update kernel) or 1 (operator pass).  Waits for the last one; returns how many there are (at
// most `cap` are written).
extern "C" int trb_profile_timeline(double* ms, int* kinds, int cap) {
  size_t next[2] = {0, 0};
  int n = 0;
  for (size_t i = 0; i < g_prof.order.size(); ++i) {
    const int k = g_prof.order[i];
    const size_t j = next[k]++;
    if (j >= g_prof.end[k].size()) break;
    if (n < cap) {
      float t = 0.f;
      cudaEventSynchronize(g_prof.end[k][j]);
      if (cudaEventElapsedTime(&t, g_prof.begin[k][j], g_prof.end[k][j]) != cudaSuccess) t = -1.f;
      if (ms) ms[n] = t;
      if (kinds) kinds[n] = k;
    }
    ++n;
  }
  return n;
}

extern "C" const char* trb_last_error(void) { return trb_err_buf; }
extern "C" int trb_version(void) { return TRB_VERSION; }
extern "C" size_t trb_sizeof_factor(void) { return sizeof(trb_factor); }
extern "C" size_t trb_sizeof_sweep(void) { return sizeof(trb_sweep); }
extern "C" int trb_device_sm_count(void) { return trb_sm_count_cached(); }

bool trb_allow_big_cluster(const void* kernel) {
  static const void* known[16];
  static bool ok[16];
  static int count = 0;
  for (int i = 0; i < count; ++i)
    if (known[i] == kernel) return ok[i];
  const bool allowed =
      cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  cudaGetLastError();
  if (count < 16) {
    known[count] = kernel;
    ok[count] = allowed;
    ++count;
  }
  return allowed;
}

int trb_cluster_size(int B, int n) {
  int sm = trb_sm_count_cached();
  if (sm <= 0) sm = 1;
  int c = 1;
  // double while the batch leaves SMs idle and every CTA keeps >= 2048 elements
  while (c < 16 && (long long)B * c * 2 <= sm && n / (c * 2) >= 2048) c *= 2;
  return c;
}

namespace {

constexpr int kEwThreads = 512;

// ---- posterior: r, v (mean or elementwise) ---------------------------------
__global__ void __launch_bounds__(kEwThreads)
k_factor_posterior(trb_factor f, int n, int ld, const double* __restrict__ a, int a_mode,
                   const double* __restrict__ b, const double* __restrict__ y,
                   double* __restrict__ r, double* __restrict__ v, int v_mode) {
  __shared__ double sh[33];
  const int inst = blockIdx.x;
  const size_t off = (size_t)inst * ld;
  const double a_s = a_mode ? 0.0 : a[inst];
  double vsum = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double ai = a_mode ? a[off + i] : a_s;
    const double yi = y ? y[off + i] : 0.0;
    const RV m = factor_moments(f, ai, b[off + i], yi);
    r[off + i] = m.r;
    if (v_mode == 2) v[off + i] = sparse_weight(f, ai, b[off + i]);
    else if (v_mode) v[off + i] = m.v;
    vsum += m.v;
  }
  if (!v_mode) {
    const double tot = block_sum(vsum, sh);
    if (threadIdx.x == 0) v[inst] = tot / n;
  }
}

__global__ void __launch_bounds__(kEwThreads)
k_factor_log_partition(trb_factor f, int n, int ld, const double* __restrict__ a, int a_mode,
                       const double* __restrict__ b, const double* __restrict__ y,
                       double* __restrict__ A, int A_mode) {
  __shared__ double sh[33];
  const int inst = blockIdx.x;
  const size_t off = (size_t)inst * ld;
  const double a_s = a_mode ? 0.0 : a[inst];
  double sum = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double ai = a_mode ? a[off + i] : a_s;
    const double yi = y ? y[off + i] : 0.0;
    const double Ai = factor_log_partition(f, ai, b[off + i], yi);
    if (A_mode) A[off + i] = Ai;
    sum += Ai;
  }
  if (!A_mode) {
    const double tot = block_sum(sum, sh);
    if (threadIdx.x == 0) A[inst] = tot / n;
  }
}

// ---- fused factor -> variable message --------------------------------------
// posterior -> mean(v) -> clip -> b_new -> damping.  One CTA per instance, or a
// thread-block cluster of C CTAs (grid (C, B)) when few large instances would
// leave the GPU idle; the mean runs over the cluster through distributed
// shared memory.  The moments are evaluated once; r is parked in `scratch`
// (same thread writes and re-reads it, so it comes back from L1/L2, not HBM).
__global__ void __launch_bounds__(kEwThreads)
k_factor_message(trb_factor f, int n, int ld, const double* __restrict__ a_in,
                 const double* __restrict__ b_in, const double* __restrict__ y, double* a_io,
                 double* b_io, double* a_copy, double damping, double* __restrict__ scratch,
                 int* flags, const int* __restrict__ active,
                 // one-iteration-back copies of what is overwritten (nullable): the sweep's
                 // `old_message_dag` (message_passing.py:356) costs one extra store, no copy kernel
                 double* __restrict__ snap_b, double* __restrict__ snap_a, double* __restrict__ snap_a_copy) {
  __shared__ double sh[33];
  __shared__ int sh_flag;
  const int inst = blockIdx.y;
  const int act = active ? active[inst] : 1;  // issued together with a: one round trip, not two
  const double a = a_in[inst];
  if (!act) return;  // uniform over the cluster
  const int T = blockDim.x * cluster_nctarank();
  const int gtid = cluster_ctarank() * blockDim.x + threadIdx.x;
  const size_t off = (size_t)inst * ld;
  double a_new;
  int flag = 0;
  if (factor_is_constant_message(f.kind)) {
    // gaussian_prior.py:86-89 / gaussian_likelihood.py:68-71: constants, no clip
    a_new = f.p0;
    for (int i = gtid; i < n; i += T) {
      const double bn = (f.kind == TRB_GAUSSIAN_PRIOR) ? f.p1 : y[off + i] * f.p0;
      const double bold = b_io[off + i];
      if (snap_b) snap_b[off + i] = bold;
      const double bd = damp(damping, bold, bn);
      b_io[off + i] = bd;
      if (bn != bn) flag |= TRB_FLAG_NAN_B;
    }
  } else {
    constexpr int U = 4;  // elements per thread whose loads are issued together
    const int step = T * U;
    double vsum = 0.0;
    for (int base = gtid; base < n; base += step) {
      double bv[U], yv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * T;
        if (i < n) {
          bv[u] = b_in[off + i];
          yv[u] = y ? y[off + i] : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * T;
        if (i < n) {
          const RV m = factor_moments(f, a, bv[u], yv[u]);
          scratch[off + i] = m.r;
          vsum += m.v;
        }
      }
    }
    const double v = cluster_sum(vsum, sh) / n;
    a_new = clip_a_new(v, a, f.amin, f.amax);
    const double ainv = a + a_new;
    for (int base = gtid; base < n; base += step) {
      double rv[U], bv[U], bo[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * T;
        if (i < n) {
          rv[u] = scratch[off + i];
          bv[u] = b_in[off + i];
          bo[u] = b_io[off + i];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * T;
        if (i < n) {
          const double bn = rv[u] * ainv - bv[u];
          if (snap_b) snap_b[off + i] = bo[u];
          b_io[off + i] = damp(damping, bo[u], bn);
          if (bn != bn) flag |= TRB_FLAG_NAN_B;
        }
      }
    }
  }
  if (a_new != a_new) flag |= TRB_FLAG_NAN_A;
  if (a_new < 0) flag |= TRB_FLAG_NEG_A;
  const int all = cluster_or(flag, &sh_flag);
  if (gtid == 0) {
    const double a_old = a_io[inst];
    if (snap_a) snap_a[inst] = a_old;
    if (snap_a_copy && a_copy) snap_a_copy[inst] = a_copy[inst];
    const double ad = damp(damping, a_old, a_new);
    a_io[inst] = ad;
    if (a_copy) a_copy[inst] = ad;
    if (flags && all) atomicOr(&flags[inst], all);
  }
}

__global__ void k_truncated_normal(int n, const double* __restrict__ r0,
                                   const double* __restrict__ v0, double zmin, double zmax,
                                   double* mean, double* var, double* logZ, double* proba) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const TruncMoments t = truncated_normal(r0[i], v0[i], zmin, zmax);
  if (mean) mean[i] = t.mean;
  if (var) var[i] = t.var;
  if (logZ) logZ[i] = t.logZ;
  if (proba) proba[i] = t.proba;
}

__global__ void __launch_bounds__(kEwThreads)
k_posterior_rv(int n, int ld, const double* __restrict__ a1, const double* __restrict__ b1,
               const double* __restrict__ a2, const double* __restrict__ b2,
               double* __restrict__ r, double* __restrict__ v) {
  const int inst = blockIdx.x;
  const size_t off = (size_t)inst * ld;
  const double a_hat = a1[inst] + a2[inst];
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    r[off + i] = (b1[off + i] + b2[off + i]) / a_hat;
  if (threadIdx.x == 0) v[inst] = 1. / a_hat;
}

bool needs_y(int kind) { return kind >= TRB_GAUSSIAN_LIKELIHOOD; }

}  // namespace

extern "C" int trb_factor_posterior(const trb_factor* f, int B, int n, int ld, const double* a,
                                    int a_mode, const double* b, const double* y, double* r,
                                    double* v, int v_mode, void* stream) {
  TRB_CHECK_ARG(f && a && b && r && v, "null pointer");
  TRB_CHECK_ARG(B > 0 && n > 0 && ld >= n, "bad shape");
  TRB_CHECK_ARG(f->kind >= 0 && f->kind <= TRB_ABS_LIKELIHOOD, "unknown factor kind");
  TRB_CHECK_ARG(!needs_y(f->kind) || y, "likelihood needs y");
  TRB_CHECK_ARG(v_mode >= 0 && v_mode <= 2, "v_mode must be 0, 1 or 2");
  TRB_CHECK_ARG(v_mode != 2 || f->kind == TRB_GAUSS_BERNOULLI_PRIOR, "v_mode 2 (mixture weight) is for the sparse belief");
  trb_launch_scope scope_(0, (cudaStream_t)stream);
  k_factor_posterior<<<B, kEwThreads, 0, (cudaStream_t)stream>>>(*f, n, ld, a, a_mode, b, y, r, v,
                                                                  v_mode);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_factor_log_partition(const trb_factor* f, int B, int n, int ld,
                                        const double* a, int a_mode, const double* b,
                                        const double* y, double* A, int A_mode, void* stream) {
  TRB_CHECK_ARG(f && a && b && A, "null pointer");
  TRB_CHECK_ARG(B > 0 && n > 0 && ld >= n, "bad shape");
  TRB_CHECK_ARG(f->kind >= 0 && f->kind <= TRB_ABS_LIKELIHOOD, "unknown factor kind");
  TRB_CHECK_ARG(!needs_y(f->kind) || y, "likelihood needs y");
  trb_launch_scope scope_(0, (cudaStream_t)stream);
  k_factor_log_partition<<<B, kEwThreads, 0, (cudaStream_t)stream>>>(*f, n, ld, a, a_mode, b, y, A,
                                                                      A_mode);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

// trb_factor_message that also keeps the overwritten message (the sweep's one-iteration-back state)
int trb_factor_message_snap(const trb_factor* f, int B, int n, int ld, const double* a_in,
                            const double* b_in, const double* y, double* a_io, double* b_io,
                            double* a_copy, double damping, double* scratch, int* flags,
                            const int* active, double* snap_b, double* snap_a, double* snap_a_copy,
                            void* stream);

extern "C" int trb_factor_message(const trb_factor* f, int B, int n, int ld, const double* a_in,
                                  const double* b_in, const double* y, double* a_io, double* b_io,
                                  double* a_copy, double damping, double* scratch, int* flags,
                                  const int* active, void* stream) {
  return trb_factor_message_snap(f, B, n, ld, a_in, b_in, y, a_io, b_io, a_copy, damping, scratch, flags,
                                 active, nullptr, nullptr, nullptr, stream);
}

int trb_factor_message_snap(const trb_factor* f, int B, int n, int ld, const double* a_in,
                            const double* b_in, const double* y, double* a_io, double* b_io,
                            double* a_copy, double damping, double* scratch, int* flags,
                            const int* active, double* snap_b, double* snap_a, double* snap_a_copy,
                            void* stream) {
  TRB_CHECK_ARG(f && a_in && b_in && a_io && b_io && scratch, "null pointer");
  TRB_CHECK_ARG(B > 0 && n > 0 && ld >= n, "bad shape");
  TRB_CHECK_ARG(f->kind >= 0 && f->kind <= TRB_ABS_LIKELIHOOD, "unknown factor kind");
  TRB_CHECK_ARG(!needs_y(f->kind) || y, "likelihood needs y");
  trb_launch_scope scope_(0, (cudaStream_t)stream);
  // 64 registers per thread: 2 CTAs of 512 threads or 4 of 256 per SM.  A batch that needs more
  // than one wave of 512-thread CTAs runs as half as many waves of 256-thread CTAs: the fixed part
  // of a CTA's life (launch, first loads, the reduction, the tail) is paid half as often.
  const int C = trb_cluster_size(B, n);
  const int threads = ((long long)B * C > 2LL * trb_sm_count_cached()) ? kEwThreads / 2 : kEwThreads;
  cudaError_t le = trb_launch_cluster(k_factor_message, C, B, threads,
                                      (cudaStream_t)stream, *f, n, ld, a_in, b_in, y, a_io, b_io,
                                      a_copy, damping, scratch, flags, active, snap_b, snap_a, snap_a_copy);
  if (le != cudaSuccess)
    return trb_set_error(TRB_ERR_CUDA, "trb_factor_message: %s", cudaGetErrorString(le));
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_truncated_normal(int n, const double* r0, const double* v0, double zmin,
                                    double zmax, double* mean, double* var, double* logZ,
                                    double* proba, void* stream) {
  TRB_CHECK_ARG(r0 && v0, "null pointer");
  TRB_CHECK_ARG(n > 0, "bad shape");
  TRB_CHECK_ARG(zmin < zmax, "zmin must be < zmax");  // truncated_normal.py:236
  trb_launch_scope scope_(0, (cudaStream_t)stream);
  k_truncated_normal<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, r0, v0, zmin, zmax, mean,
                                                                        var, logZ, proba);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_posterior_rv(int B, int n, int ld, const double* a1, const double* b1,
                                const double* a2, const double* b2, double* r, double* v,
                                void* stream) {
  TRB_CHECK_ARG(a1 && b1 && a2 && b2 && r && v, "null pointer");
  TRB_CHECK_ARG(B > 0 && n > 0 && ld >= n, "bad shape");
  trb_launch_scope scope_(0, (cudaStream_t)stream);
  k_posterior_rv<<<B, kEwThreads, 0, (cudaStream_t)stream>>>(n, ld, a1, b1, a2, b2, r, v);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}
