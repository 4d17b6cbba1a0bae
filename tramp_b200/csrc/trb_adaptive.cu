// Device-side building blocks of the factor-by-factor schedule: adaptive damping and dA
// (reference algos/message_passing.py:129-185), the EP objective (:306-328).
//
// The reference evaluates, for every single message, the local objective
//   A(target node) - A(variable of the edge)
// before and after the update and halves the step (up to 10 times) until it does not
// decrease.  On the device every message lives in (a[B], b[B, ld]) and the whole schedule is
// ENQUEUED from the host without reading anything back: the factor terms come from the moment
// kernels (trb_factor_log_partition) and the GEMV kernels (trb_lin_project), the rest from the
// kernels below; the per-instance accept / halve decision is a mask that stays on the device.
#include "trb_common.cuh"

using namespace trb;

namespace {

constexpr int kAdThreads = 256;

// (a_out, b_out) = old + beta * (new - old); beta per instance (beta_arr) or one scalar.
// message_passing.py:169-171.  A NULL b_old / b_new row source is not allowed.
__global__ void __launch_bounds__(kAdThreads)
k_message_trial(int n, int ld, const double* __restrict__ a_old, const double* __restrict__ b_old,
                const double* __restrict__ a_new, const double* __restrict__ b_new,
                const double* __restrict__ beta_arr, double beta, double* __restrict__ a_out,
                double* __restrict__ b_out) {
  const int b = blockIdx.y;
  const double bt = beta_arr ? beta_arr[b] : beta;
  const size_t off = (size_t)b * ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double o = b_old[off + i];
    b_out[off + i] = o + bt * (b_new[off + i] - o);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double o = a_old[b];
    a_out[b] = o + bt * (a_new[b] - o);
  }
}

// Variable.compute_log_partition of the two messages meeting on a variable (base.py:146-155):
// a = a1 + a2, b = b1 + b2, A = 0.5 * sum(b^2 / a + log(2 pi / a)), +inf if a <= 0.
__global__ void __launch_bounds__(kAdThreads)
k_variable_log_partition(int n, int ld, const double* __restrict__ a1, const double* __restrict__ b1,
                         const double* __restrict__ a2, const double* __restrict__ b2,
                         double* __restrict__ A) {
  __shared__ double sh[33];
  const int b = blockIdx.x;
  const size_t off = (size_t)b * ld;
  const double a = a1[b] + a2[b];
  const double lg = log(kTwoPi / a);
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double bb = b1[off + i] + b2[off + i];
    s += bb * bb / a + lg;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) A[b] = (a <= 0.0) ? INFINITY : 0.5 * s;
}

// LinearChannel.compute_log_partition (linear_channel.py:127-132) in the singular basis:
//   b = bz + W^T bx has the components tz_i + s_i tx_i along the R right singular vectors and
//   the part of bz orthogonal to them (squared norm bz2 - sum tz^2) in the null space, so
//   0.5 sum(b rz)      = 0.5 [ sum_i (tz_i + s_i tx_i)^2 / (az + ax s_i^2) + (bz2 - sum_i tz_i^2) / az ]
//   0.5 sum log(2pi/a) = 0.5 [ sum_i log(2 pi / (az + ax s_i^2)) + (Nz - R) log(2 pi / az) ]
// (the null-space terms only when R < Nz).
__global__ void __launch_bounds__(kAdThreads)
k_lin_log_partition(int R, int Nz, const double* __restrict__ s, const double* __restrict__ s2,
                    int64_t stride_s, const double* __restrict__ az_arr, const double* __restrict__ ax_arr,
                    const double* __restrict__ tz, const double* __restrict__ tx,
                    const double* __restrict__ bz2, double* __restrict__ A) {
  __shared__ double sh[33 * 3];
  const int b = blockIdx.x;
  const double az = az_arr[b], ax = ax_arr[b];
  const double* sb = s + (size_t)b * stride_s;
  const double* s2b = s2 + (size_t)b * stride_s;
  const size_t off = (size_t)b * R;
  double acc[3] = {0.0, 0.0, 0.0};  // quadratic term, log term, sum tz^2
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const double a = az + ax * s2b[i];
    const double t = tz[off + i] + sb[i] * tx[off + i];
    acc[0] += t * t / a;
    acc[1] += log(kTwoPi / a);
    acc[2] += tz[off + i] * tz[off + i];
  }
  block_sum_n<3>(acc, sh);
  if (threadIdx.x == 0) {
    double quad = acc[0], lg = acc[1];
    if (R < Nz) {
      quad += (bz2[b] - acc[2]) / az;
      lg += (double)(Nz - R) * log(kTwoPi / az);
    }
    A[b] = 0.5 * quad + 0.5 * lg;
  }
}

// out[b] = sum_i x[b, i] y[b, i]
__global__ void __launch_bounds__(kAdThreads)
k_row_dot(int n, int ld, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out) {
  __shared__ double sh[33];
  const int b = blockIdx.x;
  const size_t off = (size_t)b * ld;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s = fma(x[off + i], y[off + i], s);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[b] = s;
}

// Factor.compute_ab_new (base.py:250-255) for a channel: a_new = clip(1 / max(v, 1e-20) - a_in),
// b_new = r (a_in + a_new) - b_in.  add / add_div: r += add / add_div (the null-space term of the
// backward mean, linear_channel.py:69-83 in thin-SVD form).
__global__ void __launch_bounds__(kAdThreads)
k_message_from_posterior(int n, int ld, const double* __restrict__ r, const double* __restrict__ v,
                         const double* __restrict__ a_in, const double* __restrict__ b_in, double amin,
                         double amax, double* __restrict__ a_new, double* __restrict__ b_new) {
  const int b = blockIdx.y;
  const size_t off = (size_t)b * ld;
  const double a = a_in[b];
  const double an = clip_a_new(v[b], a, amin, amax);
  const double ainv = a + an;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    b_new[off + i] = r[off + i] * ainv - b_in[off + i];
  if (blockIdx.x == 0 && threadIdx.x == 0) a_new[b] = an;
}

// dst[b, :] = src[b, :] for the instances with mask[b] != 0 (and dst_a[b] = src_a[b])
__global__ void __launch_bounds__(kAdThreads)
k_rows_select(int n, int ld, const int* __restrict__ mask, const double* __restrict__ src_a,
              const double* __restrict__ src_b, double* __restrict__ dst_a, double* __restrict__ dst_b) {
  const int b = blockIdx.y;
  if (!mask[b]) return;
  const size_t off = (size_t)b * ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    dst_b[off + i] = src_b[off + i];
  if (blockIdx.x == 0 && threadIdx.x == 0 && src_a) dst_a[b] = src_a[b];
}

int chunks_for(int B, int n) {
  int c = (n + kAdThreads * 4 - 1) / (kAdThreads * 4);
  const int cap = (4 * trb_sm_count_cached() + B - 1) / B;
  if (c > cap) c = cap;
  return c < 1 ? 1 : c;
}

}  // namespace

#define TRB_VEC_ARGS(B, n, ld) TRB_CHECK_ARG((B) > 0 && (B) <= 65535 && (n) > 0 && (ld) >= (n), "bad shape")

extern "C" int trb_message_trial(int B, int n, int ld, const double* a_old, const double* b_old,
                                 const double* a_new, const double* b_new, const double* beta_arr,
                                 double beta, double* a_out, double* b_out, void* stream) {
  TRB_CHECK_ARG(a_old && b_old && a_new && b_new && a_out && b_out, "null pointer");
  TRB_VEC_ARGS(B, n, ld);
  cudaStream_t st = (cudaStream_t)stream;
  trb_launch_scope scope_(0, st);
  k_message_trial<<<dim3(chunks_for(B, n), B), kAdThreads, 0, st>>>(n, ld, a_old, b_old, a_new, b_new, beta_arr,
                                                                     beta, a_out, b_out);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_variable_log_partition(int B, int n, int ld, const double* a1, const double* b1,
                                          const double* a2, const double* b2, double* A, void* stream) {
  TRB_CHECK_ARG(a1 && b1 && a2 && b2 && A, "null pointer");
  TRB_VEC_ARGS(B, n, ld);
  cudaStream_t st = (cudaStream_t)stream;
  trb_launch_scope scope_(0, st);
  k_variable_log_partition<<<B, kAdThreads, 0, st>>>(n, ld, a1, b1, a2, b2, A);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_lin_log_partition(int B, int R, int Nz, const double* s, const double* s2, int64_t stride_s,
                                     const double* az, const double* ax, const double* tz, const double* tx,
                                     const double* bz2, double* A, void* stream) {
  TRB_CHECK_ARG(s && s2 && az && ax && tz && tx && A, "null pointer");
  TRB_CHECK_ARG(B > 0 && B <= 65535 && R > 0 && R <= Nz, "bad shape");
  TRB_CHECK_ARG(R == Nz || bz2, "bz2 = |bz|^2 is needed when R < Nz");
  cudaStream_t st = (cudaStream_t)stream;
  trb_launch_scope scope_(0, st);
  k_lin_log_partition<<<B, kAdThreads, 0, st>>>(R, Nz, s, s2, stride_s, az, ax, tz, tx, bz2, A);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_row_dot(int B, int n, int ld, const double* x, const double* y, double* out, void* stream) {
  TRB_CHECK_ARG(x && y && out, "null pointer");
  TRB_VEC_ARGS(B, n, ld);
  cudaStream_t st = (cudaStream_t)stream;
  trb_launch_scope scope_(0, st);
  k_row_dot<<<B, kAdThreads, 0, st>>>(n, ld, x, y, out);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_message_from_posterior(int B, int n, int ld, const double* r, const double* v,
                                          const double* a_in, const double* b_in, double amin, double amax,
                                          double* a_new, double* b_new, void* stream) {
  TRB_CHECK_ARG(r && v && a_in && b_in && a_new && b_new, "null pointer");
  TRB_VEC_ARGS(B, n, ld);
  cudaStream_t st = (cudaStream_t)stream;
  trb_launch_scope scope_(0, st);
  k_message_from_posterior<<<dim3(chunks_for(B, n), B), kAdThreads, 0, st>>>(n, ld, r, v, a_in, b_in, amin, amax,
                                                                              a_new, b_new);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_rows_select(int B, int n, int ld, const int* mask, const double* src_a, const double* src_b,
                               double* dst_a, double* dst_b, void* stream) {
  TRB_CHECK_ARG(mask && src_b && dst_b && (!src_a || dst_a), "null pointer");
  TRB_VEC_ARGS(B, n, ld);
  cudaStream_t st = (cudaStream_t)stream;
  trb_launch_scope scope_(0, st);
  k_rows_select<<<dim3(chunks_for(B, n), B), kAdThreads, 0, st>>>(n, ld, mask, src_a, src_b, dst_a, dst_b);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}
