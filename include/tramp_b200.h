/* tramp_b200 -- C ABI of the B200-native expectation-propagation (EP) sweep.
 *
 * The reference (sphinxteam/tramp) is pure Python/numpy and has NO FFI: its
 * plugin boundary is the Python factor protocol (compute_forward_posterior,
 * compute_backward_message, ... -- tramp/base.py:329-365, docs/implementation
 * .rst:41-147).  tramp_b200 keeps that protocol in Python (the tramp_b200
 * package) and adds this thin C layer underneath; each entry point below cites the reference
 * routine whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stubs
 * a tramp maintainer would add to call it.
 *
 * Conventions
 *  - every function returns 0 on success, a negative trb_status otherwise;
 *    trb_last_error() gives a thread-local message.  No C++ exception crosses.
 *  - all buffers are CALLER-OWNED DEVICE pointers (FP64, contiguous,
 *    batch-major); the library never frees or retains them past the call.
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous
 *    on that stream.  One host thread per plan/device.
 *  - a "batch" is B independent teacher-student instances: scalars are [B],
 *    vectors [B, ld] with leading dimension ld >= n (ld % 2 == 0; padding
 *    columns are ignored on input and left untouched on output).
 */
#ifndef TRAMP_B200_H
#define TRAMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRB_VERSION 100

typedef enum {
  TRB_OK = 0,
  TRB_ERR_INVALID = -1,  /* bad argument */
  TRB_ERR_CUDA = -2,     /* CUDA runtime error (message in trb_last_error) */
  TRB_ERR_UNSUPPORTED = -3
} trb_status;

/* Separable factor kinds (priors and likelihoods share one moment engine). */
typedef enum {
  TRB_GAUSS_BERNOULLI_PRIOR = 0, /* priors/gauss_bernoulli_prior.py:70-83 + beliefs/sparse.py */
  TRB_BINARY_PRIOR = 1,          /* priors/binary_prior.py:57-68 + beliefs/binary.py */
  TRB_GAUSSIAN_PRIOR = 2,        /* priors/gaussian_prior.py:63-89 */
  TRB_GAUSSIAN_LIKELIHOOD = 3,   /* likelihoods/gaussian_likelihood.py:43-71 */
  TRB_SGN_LIKELIHOOD = 4,        /* likelihoods/sgn_likelihood.py:32-41 + beliefs/positive.py */
  TRB_ABS_LIKELIHOOD = 5         /* likelihoods/abs_likelihood.py:31-40 */
} trb_factor_kind;

/* Natural parameters of one separable factor.
 *   GAUSS_BERNOULLI: p0 = 1/var, p1 = mean/var, p2 = eta (gauss_bernoulli_prior.py:33-36),
 *                    p3 = sparse.A(p0, p1, eta)  (the constant subtracted in :82)
 *   BINARY:          p0 = 0.5*log(p_pos/p_neg) (binary_prior.py:28)
 *   GAUSSIAN_PRIOR:  p0 = 1/var, p1 = mean/var
 *   GAUSSIAN_LIKELIHOOD: p0 = 1/var
 *   SGN, ABS: no parameters
 * amin/amax: Factor.AMIN/AMAX (base.py:238-243). */
typedef struct {
  int32_t kind;
  int32_t _pad;
  double p0, p1, p2, p3;
  double amin, amax;
} trb_factor;

/* flag bits written per instance by the message kernels
 * (algos/message_passing.py:187-209 check_message) */
#define TRB_FLAG_NAN_A 1
#define TRB_FLAG_NAN_B 2
#define TRB_FLAG_NEG_A 4
/* set by the sweep's EarlyStoppingEP logic (callbacks.py:266-283) */
#define TRB_FLAG_CONVERGED 8
#define TRB_FLAG_DIVERGED 16
/* the instance's messages were rolled back to the end of the previous iteration
 * (message_passing.py:196-197, callbacks.py:281-283 reset_message_dag) */
#define TRB_FLAG_RESTORED 32
/* a rank of a row-sharded operator waited > ~1 s for a peer's partial sums */
#define TRB_FLAG_COMM_TIMEOUT 64

const char* trb_last_error(void);
int trb_version(void);
size_t trb_sizeof_factor(void);
size_t trb_sizeof_sweep(void);
/* number of SMs of the current device (0 if no device) */
int trb_device_sm_count(void);

/* Launch accounting for benchmarks: trb_profile_reset(enable_events) zeroes the
 * counters (and, if enable_events, brackets every launch of the sweep with CUDA events
 * on its stream); trb_profile_launches(kind) = kernels launched since then
 * (kind 0 elementwise/update, 1 operator pass = GEMV or shared-operator GEMM,
 * 2 LinearChannel set-up kernels (trb_jacobi_sweep, ...), -1 = 0 and 1 together: the EP
 * sweep); trb_profile_gemv_ms sums the event-timed operator-pass durations and
 * returns how many launches were timed. */
void trb_profile_reset(int enable_events);
long long trb_profile_launches(int kind);
int trb_profile_gemv_ms(double* total_ms);
/* With events enabled every sweep launch is bracketed: the durations in launch order (ms[i];
 * kinds[i] = 0 update kernel, 1 operator pass).  Returns how many launches were timed since the
 * reset; at most `cap` entries are written. */
int trb_profile_timeline(double* ms, int* kinds, int cap);

/* ---- elementwise moment kernels ------------------------------------------
 * a_mode: 0 = one precision per instance a[B] (isotropic beliefs), 1 = one per
 * element a[B, ld] (isotropic=False in the reference's unit tests).
 * v_mode: 0 = v[B] is the per-instance MEAN of the elementwise variance
 * (`vx.mean()`), 1 = v[B, ld] elementwise, 2 (TRB_GAUSS_BERNOULLI_PRIOR only) =
 * v[B, ld] receives the weight of the Gaussian component instead,
 * expit(normal.A(a, b) - eta) (beliefs/sparse.py:9-12 `p`).  y is NULL for priors. */

/* compute_forward_posterior / compute_backward_posterior of the factor kinds */
int trb_factor_posterior(const trb_factor* f, int B, int n, int ld,
                         const double* a, int a_mode, const double* b,
                         const double* y, double* r, double* v, int v_mode,
                         void* stream);

/* compute_log_partition: A_mode 0 = per-instance MEAN A[B] (what the reference
 * returns), 1 = elementwise A[B, ld] (scalar_log_partition vectorised). */
int trb_factor_log_partition(const trb_factor* f, int B, int n, int ld,
                             const double* a, int a_mode, const double* b,
                             const double* y, double* A, int A_mode,
                             void* stream);

/* Whole factor->variable update in one pass over HBM:
 * posterior -> mean variance -> compute_ab_new clip (base.py:250-255) ->
 * constant damping against the stored message (message_passing.py:119-127).
 * (a_io, b_io) hold the stored (old) message on entry and the new one on exit.
 * Gaussian kinds emit their constant message (gaussian_prior.py:86-89,
 * gaussian_likelihood.py:68-71), still damped.  a_copy (nullable) receives a
 * copy of the new a (the SISOVariable pass-through, sub_variables.py:16-31).
 * scratch: [B, ld] doubles.  flags (nullable): int[B], OR-ed TRB_FLAG_*.
 * active (nullable): int[B], instances with 0 are skipped. */
int trb_factor_message(const trb_factor* f, int B, int n, int ld,
                       const double* a_in, const double* b_in, const double* y,
                       double* a_io, double* b_io, double* a_copy,
                       double damping, double* scratch, int* flags,
                       const int* active, void* stream);

/* Truncated-normal moments on [zmin, zmax] (utils/truncated_normal.py:234-298,
 * all five F0/F1/F2 branches and the half-infinite erfcx fast path).  Any of
 * the outputs may be NULL. r0, v0: [n]. */
int trb_truncated_normal(int n, const double* r0, const double* v0,
                         double zmin, double zmax, double* mean, double* var,
                         double* logZ, double* proba, void* stream);

/* Variable.posterior_rv (base.py:152-161): r = (b1+b2)/(a1+a2), v = 1/(a1+a2) */
int trb_posterior_rv(int B, int n, int ld, const double* a1, const double* b1,
                     const double* a2, const double* b2, double* r, double* v,
                     void* stream);

/* ---- LinearChannel in thin-SVD form --------------------------------------
 * W = U_R diag(s) V_R^T.  Operators are stored TRANSPOSED, one singular vector
 * per row: Vt[Bop, R, ldn] (rows = right singular vectors, length n = Nz) and
 * Ut[Bop, R, ldm].  strideA = elements between instances (0: all instances
 * share one operator).  impl: 0 = default, 1 = plain LDG kernels,
 * 2 = TMA(cp.async.bulk)+mbarrier ring. */

/* t[b, i] = sum_j A[b, i, j] * vec[b, j]      (U.T @ bx, V.T @ bz;
 * channels/linear/linear_channel.py:72-73).  t: [B, R]. */
int trb_lin_project(const double* A, int64_t strideA, int R, int n, int ld,
                    int B, const double* vec, int ldvec, double* t,
                    const int* active, int impl, void* stream);

/* number of partial-sum slots trb_lin_expand writes per instance (>= 1) */
int trb_lin_expand_slots(int B, int R);

/* part[b, slot, j] = sum_{i in rows of CTA `slot`} coef[b, i] * A[b, i, j]
 * (V @ rz_svd, linear_channel.py:78).  part: [B, nslots, ld]; slots beyond the
 * ones an instance uses are not written -- reduce with trb_lin_reduce_slots or
 * the fused sweep kernels. */
int trb_lin_expand(const double* A, int64_t strideA, int R, int n, int ld,
                   int B, const double* coef, double* part,
                   const int* active, int impl, void* stream);

/* out[b, j] = sum_slots part[b, slot, j] (+ add_scale[b] * add[b, j] if add) */
int trb_lin_reduce_slots(int B, int R, int n, int ld, const double* part,
                         const double* add, const double* add_scale_inv,
                         double* out, void* stream);

/* The same two passes for B >= 1 instances that SHARE one operator A[R, ld]
 * (2-D W in LinearChannel), as dense FP64 tensor-core GEMMs (DMMA m8n8k4):
 *   trb_lin_project_gemm: t[B, R]       = vec[B, n] . A^T     (linear_channel.py:72-73)
 *   trb_lin_expand_gemm : out[B, ldout] = coef[B, R] . A      (linear_channel.py:78, :88)
 * `out` is the complete sum (no slots); columns >= n of out are left untouched. */
int trb_lin_project_gemm(const double* A, int R, int n, int ld, int B,
                         const double* vec, int ldvec, double* t, void* stream);
int trb_lin_expand_gemm(const double* A, int R, int n, int ld, int B,
                        const double* coef, double* out, int ldout, void* stream);
/* kernel behind the two calls above: 0 (default) = TMA + mbarrier pipeline when
 * the operands are 16-byte aligned, else the cp.async one; 1 = always cp.async;
 * 2 / 4 = measurement probes that do NOT compute the product (MMA loop without
 * loads / loads without MMA), used by tools/bench_gemm.py only */
void trb_gemm_set_variant(int variant);

/* ---- factor-by-factor schedule: adaptive damping, dA, the EP objective (trb_adaptive.cu) ----
 * Device building blocks of algos/message_passing.py:129-185 (compute_dA, compute_adaptive_damping)
 * and :306-328 (update_objective).  A message is (a[B], b[B, ld]); nothing is read back. */
/* (a_out, b_out) = old + beta (new - old)  (:169-171); beta per instance (beta_arr[B]) or, if
 * beta_arr is NULL, the scalar `beta`.  Outputs may alias neither input. */
int trb_message_trial(int B, int n, int ld, const double* a_old, const double* b_old,
                      const double* a_new, const double* b_new, const double* beta_arr,
                      double beta, double* a_out, double* b_out, void* stream);
/* Variable.compute_log_partition of the two messages meeting on a variable (base.py:146-155):
 * A[b] = 0.5 sum((b1 + b2)^2 / (a1 + a2) + log(2 pi / (a1 + a2))), +inf if a1 + a2 <= 0 (a SUM) */
int trb_variable_log_partition(int B, int n, int ld, const double* a1, const double* b1,
                               const double* a2, const double* b2, double* A, void* stream);
/* LinearChannel.compute_log_partition (linear_channel.py:127-132) from the singular-basis vectors
 * tz = V_R^T bz, tx = U_R^T bx [B, R] and bz2[b] = |bz|^2 (needed when R < Nz: the part of bz in
 * the null space of W).  s, s2: [Bop, R], stride_s = 0 when shared.  A[B] (a SUM). */
int trb_lin_log_partition(int B, int R, int Nz, const double* s, const double* s2, int64_t stride_s,
                          const double* az, const double* ax, const double* tz, const double* tx,
                          const double* bz2, double* A, void* stream);
/* out[b] = sum_i x[b, i] y[b, i] */
int trb_row_dot(int B, int n, int ld, const double* x, const double* y, double* out, void* stream);
/* Factor.compute_ab_new (base.py:250-255), the message of a channel from its posterior:
 * a_new = clip(1 / max(v, 1e-20) - a_in, amin, amax), b_new = r (a_in + a_new) - b_in */
int trb_message_from_posterior(int B, int n, int ld, const double* r, const double* v,
                               const double* a_in, const double* b_in, double amin, double amax,
                               double* a_new, double* b_new, void* stream);
/* dst_b[b, :] = src_b[b, :] (and dst_a[b] = src_a[b] unless src_a is NULL) where mask[b] != 0:
 * the per-instance "accept this step size" of the adaptive damping */
int trb_rows_select(int B, int n, int ld, const int* mask, const double* src_a, const double* src_b,
                    double* dst_a, double* dst_b, void* stream);

/* ---- LinearChannel set-up: thin SVD by one-sided block Jacobi (trb_setup.cu) ----
 * Replaces np.linalg.svd / np.linalg.matrix_rank of channels/linear/linear_channel.py:8-15,
 * 36-46 (LAPACK gesdd on the host; counted in EP's total by examples/figures/benchmark.py:22).
 *
 * A[B, np, ld] (np % 32 == 0, ld % 64 == 0, 16-byte aligned, padding rows and columns
 * zero) holds np vectors per instance, one per row: either the rows of the Gram matrix of
 * the short side of W (then the rotated rows converge to lambda_i u_i^T) or the rows of the
 * short side of W itself (-> s_i v_i^T).  One call = one sweep: every pair of 16-row blocks
 * meets once (round-robin), A[b] <- Q^T A[b] in place with Q orthogonal, three launches per
 * round over the whole batch (pair Gram on DMMA, 32 x 32 eigenvectors by cyclic Jacobi in
 * shared memory, pair rotation on DMMA).
 *   Swork  [B, np/32, trb_jacobi_zsplit(B, np, ld), 1024]   partial pair Grams
 *   Jwork  [B, np/32, 1024]                                 pair rotations
 *   rot_flag [2, B, np/32] int                              [0]: 0 = pair already orthogonal, skipped;
 *                                                           [1]: scratch (arrival counters of a pair's CTAs)
 *   offmax [B]  out: largest |cos| between two rows of a pair BEFORE its rotation
 *   skip_tol   pairs whose largest |cos| is <= skip_tol are left untouched
 *   max_inner  sweeps of the inner 32 x 32 Jacobi (2 is enough: the outer sweeps converge
 *              quadratically either way) */
int trb_jacobi_zsplit(int B, int np, int ld);
/* tuning: the column range of a pair is split over CTAs until the grid holds about `waves`
 * waves of resident CTAs (default 4; fewer waves = longer CTAs, more tail) */
void trb_jacobi_set_waves(int waves);
/* kernel fusion (tuning / A-B tests), a bit mask, default 1: bit 0 = rows of at most 768 doubles and
 * one wave of pairs: Gram, eigenvectors and rotation of a pair as ONE kernel per round, the pair
 * resident in shared memory; bit 1 = Gram and eigenvectors in one kernel (the CTA that completes a
 * pair's Gram matrix goes on to its eigen-solve; measured no faster, off by default); 0 = three
 * launches per round */
void trb_jacobi_set_fused(int mask);
int trb_jacobi_sweep(double* A, int64_t strideA, int B, int np, int ld, double* Swork,
                     double* Jwork, int* rot_flag, double* offmax, double skip_tol,
                     int max_inner, void* stream);
/* norms[b, i] = || A[b, i, :n] ||_2;  A: [B, rows, ld] (ld even, 16-byte aligned) */
int trb_row_norms(const double* A, int64_t strideA, int B, int rows, int n, int ld,
                  double* norms, void* stream);
/* dst[b, i, :n] = scale[b, i] * src[b, perm[b, i], :n], i < R; columns n..ld_dst-1 of dst are
 * zeroed.  perm (int64) and scale may be NULL (identity / 1); dst may be src when perm is
 * NULL.  Used to sort, normalise and pad the singular vectors. */
int trb_rows_gather_scale(const double* src, int64_t stride_src, int ld_src,
                          const long long* perm, const double* scale, int B, int R, int n,
                          double* dst, int64_t stride_dst, int ld_dst, void* stream);

/* Spectrum rescale between the projections and the expansions, plus the
 * variances (linear_channel.py:58-67, 74, 91-105).
 *   dir = 0 (forward, x-side mean):  coef = s*res*(tz + s*tx),  v = forward variance
 *   dir = 1 (backward, z-side mean): coef = res*(tz + s*tx)               if !null_space
 *                                    coef = res*(s*tx - (ax*s2/az)*tz)   if  null_space
 *                                    (then rz = bz/az + V_R coef), v = backward variance
 * with res = 1/(az + ax*s2).  null_space = (min(Nx, Nz) < Nz), passed explicitly
 * because R may be a row shard of the spectrum.  s, s2: [Bop, R] (stride_s = 0
 * if shared); rank: singular = spectrum[:rank] (linear_channel.py:46).  Either
 * of coef / v may be NULL (a row-sharded operator computes the variance from
 * the full spectrum and the coefficients from its shard). */
int trb_lin_rescale(int dir, int B, int R, int Nz, int Nx, int rank, int null_space,
                    const double* s, const double* s2, int64_t stride_s,
                    const double* az, const double* ax, const double* tz,
                    const double* tx, double* coef, double* v,
                    const int* active, void* stream);

/* ---- peer-memory exchange for a row-sharded operator ----------------------
 * One process per GPU of one NVLink / NVSwitch node (<= TRB_MAX_RANKS).  Every
 * rank calls trb_comm_create (allocates its exchange buffer, returns a 64-byte
 * CUDA IPC handle), the caller all-gathers the handles by any out-of-band means
 * (torch.distributed in tramp_b200) and passes the nranks*64 bytes to
 * trb_comm_connect, which maps every peer's buffer.  Inside the sweep a rank
 * pushes its partial expansion into its slot of every rank's buffer (posted
 * stores over NVLink from the slot-reduction kernel), publishes a sequence number
 * into every rank's flag array, and the consumer kernel adds the ranks' vectors
 * from its own memory, in rank order (bit-identical on all ranks).
 * trb_comm_all_reduce exposes the same protocol as a stand-alone sum; every rank
 * must issue the same sequence of exchanges.  timeout_flag (nullable, device
 * int) is set to 1 if a peer did not publish within ~1 s. */
#define TRB_MAX_RANKS 8
typedef struct trb_comm trb_comm;
int trb_comm_create(int rank, int nranks, size_t vec_doubles, trb_comm** out,
                    unsigned char* handle64);
int trb_comm_connect(trb_comm* c, const unsigned char* handles);
int trb_comm_destroy(trb_comm* c);
int trb_comm_all_reduce(trb_comm* c, double* vec, size_t n, int* timeout_flag, void* stream);

/* ---- device-resident sweep ------------------------------------------------
 * algos/message_passing.py:330-357 (iterate) with :249-269 (forward_message,
 * backward_message, update_variables), :70-127 (constant damping), :187-209
 * (NaN check) and callbacks.py:250-286 (EarlyStoppingEP), for the chain
 * prior -> x -> LinearChannel -> z -> likelihood, B instances in lock step,
 * no host round trip inside. */
typedef struct {
  int32_t B, N, M, R, ldn, ldm, rank, nslots;
  trb_factor prior, lik;
  double lin_amin, lin_amax;
  /* operators */
  const double* Vt; int64_t strideV;
  const double* Ut; int64_t strideU;
  const double* s; const double* s2; int64_t stride_s;
  /* observations and (optional) ground truth for the MSE trajectory */
  const double* y;       /* [B, ldm] */
  const double* x_true;  /* [B, ldn] or NULL */
  /* messages: edge_a[8, B] = a of e1..e8 (SURVEY 3.3 numbering);
   * b1,b7: [B, ldn]; b3,b5: [B, ldm].  e2=e1, e4=e3, e6=e5, e8=e7 are exact
   * pass-throughs (sub_variables.py:16-31) and alias their source vector. */
  double* edge_a;
  double* b1; double* b3; double* b5; double* b7;
  /* iteration-0 inputs when the initializer gave e6/e8 a value different from
   * e5/e7 (NoisyInit); NULL = alias b5/b7 */
  const double* b6_init; const double* b8_init;
  /* damping of the factor->variable edges e1, e3, e5, e7 (0 = none) */
  double damp1, damp3, damp5, damp7;
  /* posteriors (update_variables): rx [B, ldn], rz [B, ldm], vx, vz [B] */
  double* rx; double* rz; double* vx; double* vz;
  /* scratch: tz, tx, coef: [B, R]; part: [B, nslots, max(ldn, ldm)];
   * scr_n: [B, ldn]; scr_m: [B, ldm]; vlin: [B]; stats: [B, 4] (columns 0, 1: tolerance sums of z;
   * columns 2, 3: reserved for the library) */
  double* tz; double* tx; double* coef; double* part;
  double* scr_n; double* scr_m; double* vlin; double* stats;
  /* per-instance control/status: active[B] (1 = iterate), flags[B],
   * n_iter[B] (iterations completed) */
  int32_t* active; int32_t* flags; int32_t* n_iter;
  /* per-iteration records, [max_records, B] each (NULL = do not record):
   * mse of x, sign-symmetric mse of x, v of x, v of z, EarlyStoppingEP tol */
  double* rec_mse; double* rec_smse; double* rec_vx; double* rec_vz; double* rec_tol;
  int32_t max_records;
  /* EarlyStoppingEP(ids="all"): enabled if es_tol >= 0 */
  double es_tol, es_max_increase; int32_t es_wait_increase;
  int32_t gemv_impl;   /* operator passes: 0 default (DMMA GEMM if the batch shares one operator
                        * and B >= 16, else TMA-ring GEMV), 1 LDG GEMV, 2 TMA-ring GEMV, 3 DMMA GEMM
                        * (needs strideV == strideU == 0) */
  int32_t es_vars;     /* variables the tolerance runs over: bit 0 = x, bit 1 = z (ids="all": 3) */
  /* one-iteration-back snapshot (`old_message_dag`, message_passing.py:356): same
   * shapes as edge_a, b1, b3, b5, b7, rx, rz, vx, vz, tx.  NULL snap_edge_a = no
   * snapshots (then a NaN / diverged instance is frozen but not rolled back). */
  double* snap_edge_a; double* snap_b1; double* snap_b3; double* snap_b5; double* snap_b7;
  double* snap_rx; double* snap_rz; double* snap_vx; double* snap_vz; double* snap_tx;
  /* rank of the whole thin SVD when R is only this GPU's row shard (0: = R) */
  int32_t R_total;
  /* Operator passes per iteration, used by trb_sweep_run:
   *  0  general: 4 passes, 16 R (N+M) bytes per instance-iteration, any likelihood;
   *  1  Gaussian likelihood, 3 passes, 8 R (2N+M) bytes: its message is the constant
   *     (1/var, y/var) (gaussian_likelihood.py:68-71), so with constant damping d5
   *     b5' = d5 b5 + (1-d5) y/var and U_R^T b5' = d5 tx + (1-d5) ty/var exactly:
   *     the pass U_R^T b6 becomes an R-vector recurrence on the cached ty = U_R^T y;
   *  2  as 1, and the z branch (U_R coef -> e3, posterior of z), which feeds nothing
   *     back when the likelihood message is constant, is evaluated only in the LAST
   *     iteration of each trb_sweep_run call: 2 passes, 16 R N bytes.  Needs damp3 = 0
   *     and no early stopping; the records of v_z stay exact, rec_tol covers x only.
   * 1 and 2 need lik.kind = TRB_GAUSSIAN_LIKELIHOOD, b6_init = NULL and `ty` filled by
   * trb_sweep_stage(TRB_STAGE_PROJECT_Y). */
  int32_t schedule;
  double* ty;  /* [B, R], U_R^T y (schedules 1, 2); else may be NULL */
  /* Row-sharded operator (one instance over several GPUs): R is this
   * rank's block of singular triplets, R_total the whole rank, s_full / s2_full
   * the whole spectrum [R_total] (replicated; the variances need all of it) and
   * comm the peer-memory exchange of the expansions (trb_comm_*).  NULL comm =
   * not sharded. */
  struct trb_comm* comm;
  const double* s_full; const double* s2_full;
  /* Which early-stopping test es_tol / es_max_increase / es_wait_increase drive:
   *  0  EarlyStoppingEP (callbacks.py:250-286): relative change of the posterior means;
   *  1  EarlyStopping (callbacks.py:195-243): absolute change of the posterior variances
   *     of the tracked variables, stop also when one drops below es_min_variance; a NaN
   *     variance or an increase above es_max_increase rolls back.  Needs the snapshot
   *     buffers (the previous variances are read from snap_vx / snap_vz). */
  int32_t es_mode; int32_t _pad_es;
  double es_min_variance;
} trb_sweep;

/* Stages of one iteration, in order (trb_sweep_run loops over them).  Back ends
 * that replace the GEMV stages (shared-W GEMM, multi-GPU row shards with an
 * all-reduce) call the other stages one by one through trb_sweep_stage. */
typedef enum {
  TRB_STAGE_PRIOR = 0,          /* F1 */
  TRB_STAGE_PROJECT_Z = 1,      /* P1: tz = V_R^T b2 */
  TRB_STAGE_PROJECT_X_INIT = 2, /* first iteration only: tx = U_R^T b6 */
  TRB_STAGE_RESCALE_FWD = 3,    /* S1 */
  TRB_STAGE_EXPAND_X = 4,       /* P2 */
  TRB_STAGE_Z_UPDATE = 5,       /* Z  */
  TRB_STAGE_PROJECT_X = 6,      /* P3 */
  TRB_STAGE_RESCALE_BWD = 7,    /* S2 */
  TRB_STAGE_EXPAND_Z = 8,       /* P4 */
  TRB_STAGE_X_UPDATE = 9,       /* X  */
  TRB_STAGE_SNAPSHOT = 10,
  TRB_STAGE_Z_UPDATE_LIGHT = 11, /* schedule 2: scalar part of Z and e5 only */
  TRB_STAGE_TX_RECUR = 12,       /* schedules 1, 2: tx = d5 tx + (1-d5) ty/var replaces P3 */
  TRB_STAGE_PROJECT_Y = 13       /* ty = U_R^T y, once per model */
} trb_stage;

/* pre_reduced = 1: `part` holds the complete expansion result in slot 0
 * ([B, 1, ld], nslots = 1) instead of per-CTA partial sums. */
int trb_sweep_stage(const trb_sweep* sw, int stage, int it, int first, int pre_reduced,
                    void* stream);

/* Run `n_iter` EP iterations.  `it0` is the index, within the current
 * iterate() call, of the first one (records go to row it0, it0+1, ...; the
 * early-stopping test needs it > 0).  fresh = 1: the message buffers hold the
 * initializer's values (b6_init / b8_init are honoured on the first iteration
 * and tx = U_R^T b6 is recomputed); fresh = 2: same, and the initial e6 has
 * b = 0, so tx is simply zeroed; fresh = 0: warm start / continuation. */
int trb_sweep_run(const trb_sweep* sw, int it0, int n_iter, int fresh, void* stream);

/* trb_sweep_run replays the identical middle iterations of a launch-bound sweep
 * (less than ~1 GB of operator traffic per iteration) as a CUDA graph; 0 turns
 * that off (default on; the environment variable TRB_CUDA_GRAPHS=0 does the same). */
void trb_set_cuda_graphs(int enabled);

/* trb_sweep_run runs the rescale stages S1 / S2 INSIDE the GEMV projections P1 / P3 that feed
 * them (the thread that finishes the block reduction of a row writes the row's coefficient next
 * to its projection, from operands loaded while the row streamed; the variance comes from the CTA
 * that owns the instance's first row): 7 launches per iteration instead of 9.  Applies when the
 * operator passes are TMA GEMVs and the sweep is not row-sharded; 0 turns it off (default on;
 * TRB_FUSE_RESCALE=0 does the same). */
void trb_set_fused_rescale(int enabled);

/* The x update (bit 0) and the z update with a Gaussian likelihood (bit 1) run as CHUNKED
 * kernels: 1024 elements per CTA, every load of a thread issued at once, the chunk sums added in
 * chunk order by the CTA that arrives last (arrival counters in columns 2, 3 of `stats`, the chunk
 * sums borrow scr_n / scr_m).  A cleared bit selects the one-CTA(-cluster)-per-instance kernel
 * instead; -1 = default (both; TRB_UPDATE_KERNELS=<mask> does the same).  Callers of
 * trb_sweep_stage zero `stats` once before the first stage; trb_sweep_run does it itself. */
void trb_set_update_kernels(int mask);

/* A single instance (B = 1) whose iteration is launch-bound runs ALL its
 * iterations inside one cooperative launch of one CTA per SM, four grid-wide
 * barriers per iteration (tramp_b200/csrc/trb_persist.cu); trb_sweep_run picks
 * it automatically.  The smallest instances (<= 5 MB of operator traffic per
 * iteration) use a single 16-CTA thread-block cluster and the hardware cluster
 * barrier instead of the whole grid.  mode: -1 = automatic (default), 0 = never,
 * 1 = whenever the hard limits allow (B = 1, snapshot buffers present,
 * 2 max(ldn, ldm) + R doubles of shared memory), 2 / 3 = as 1 but always the
 * grid / the cluster variant.  Environment variable TRB_PERSISTENT_SWEEP sets the
 * default. */
void trb_set_persistent_sweep(int mode);

/* ---- State Evolution (SE): the scalar twin of the sweep ---------------------
 * algos/state_evolution.py:5-27 runs the SAME schedule (message_passing.py:
 * 249-269, 330-357) on the precisions `a` alone; every factor replaces its
 * posterior by the AVERAGE of the posterior variance over the law of its
 * incoming beliefs ("beliefs_measure"), a Gaussian integral that the reference
 * evaluates with scipy.integrate.quad / dblquad over [-10, 10]
 * (utils/integration.py:13-46).  Here the integrals are fixed-node Gauss-
 * Legendre sums evaluated by one CTA per problem, and the whole SE recursion
 * of G independent problems (a grid of alpha / rho / noise values) runs in one
 * launch. */

/* Quadrature rule for  integral over [-10, 10] of N(t) f(m + s t) dt
 * (utils/integration.py:25-27).  The integrands of SE have their structure (the
 * sigmoid / tanh / erfcx transitions of the posterior moments) around one point
 * t = c, the image of b = -b0, at a scale that shrinks like 1/sqrt(a): scipy's
 * adaptive quad finds it by bisection; a fixed rule must be told.  The rule is a
 * composite Gauss-Legendre rule (P panels of Q nodes; template nodes x[Q] and
 * weights w[Q] on [-1, 1], device arrays) in the variable u of the map
 *     t = c + kappa sinh(u),   u in [asinh((-10 - c)/kappa), asinh((10 - c)/kappa)],
 * i.e. node spacing ~kappa at c growing in proportion to |t - c| away from it:
 * every feature whose width is not much smaller than its distance to c is
 * resolved, at any precision a.  The kernels compute c per integral.  The 2-D
 * measure of AbsLikelihood (abs_likelihood.py:56-65, gaussian_measure_2d) uses
 * the tensor rule (P2, Q2, kappa2): outer variable z centred at its kink z = 0,
 * inner variable centred on the line b_z = 0. */
typedef struct {
  const double* x; const double* w; int32_t Q; int32_t P; double kappa;
  const double* x2; const double* w2; int32_t Q2; int32_t P2; double kappa2;
} trb_quadrature;

#define TRB_MEASURE_V 0 /* average posterior variance: compute_forward_error / compute_backward_error
                         * (priors/base_prior.py:71-74, likelihoods/base_likelihood.py:78-81) */
#define TRB_MEASURE_A 1 /* average log-partition: compute_free_energy
                         * (base_prior.py:82-85, base_likelihood.py:88-92) */
/* the belief law of a likelihood needs az > 1/tau_z (sgn_likelihood.py:80-81,
 * abs_likelihood.py:57-58 `assert mz_hat > 0`); set instead of asserting */
#define TRB_FLAG_SE_DOMAIN 128

/* out[b] = factor.beliefs_measure(a[b] [, tau[b]], f) for f = scalar variance
 * (TRB_MEASURE_V) or scalar log-partition (TRB_MEASURE_A):
 *   GAUSS_BERNOULLI  priors/gauss_bernoulli_prior.py:112-118
 *   BINARY           priors/binary_prior.py:80-84
 *   SGN              likelihoods/sgn_likelihood.py:79-92
 *   ABS              likelihoods/abs_likelihood.py:56-65
 *   GAUSSIAN prior / likelihood: the closed forms the reference overrides with
 *     (gaussian_prior.py:92-95, 134-138; gaussian_likelihood.py:57-60, 129-132).
 * factors: DEVICE array, factors[b * factor_stride] (stride 0: one factor for all).
 * tau: [B] second moment of the factor's variable (likelihoods and the Gaussian
 * prior's free energy need it; NULL otherwise).  flags (nullable): int[B]. */
int trb_se_measure(const trb_factor* factors, int factor_stride, int what, int B,
                   const double* a, const double* tau, const trb_quadrature* q,
                   double* out, int32_t* flags, void* stream);

#define TRB_SE_MARCHENKO_PASTUR 0 /* channels/linear/analytical_linear_channel.py:25-51, 68-71 +
                                   * ensembles/marchenko_pastur_ensemble.py:40-46 */
#define TRB_SE_SPECTRUM 1         /* channels/linear/linear_channel.py:58-67, 91-105, 119-125 */

/* G independent SE problems on the chain prior -> x -> linear -> z -> likelihood,
 * one CTA each, all iterations inside one launch.  Edge numbering as trb_sweep. */
typedef struct {
  int32_t G;
  int32_t channel;             /* TRB_SE_MARCHENKO_PASTUR | TRB_SE_SPECTRUM */
  const trb_factor* prior;     /* device [G] */
  const trb_factor* lik;       /* device [G] */
  const double* tau_x;         /* [G] Prior.second_moment() */
  const double* tau_z;         /* [G] channel.second_moment(tau_x) (base_model.py:111-124) */
  /* Marchenko-Pastur channel */
  const double* alpha;         /* [G] Nx / Nz of the channel */
  const double* mean_spectrum; /* [G] marchenko_pastur_ensemble.py:13 */
  /* empirical spectrum of a LinearChannel: s2[G or 1, R] = diag(S^T S)[:R] */
  const double* s2; int64_t stride_s2; int32_t R, Nz, Nx, rank;
  double lin_amin, lin_amax;   /* AMIN / AMAX of the channel */
  double damp1, damp3, damp5, damp7;
  double* edge_a;              /* [8, G] in: initial values (initializer), out: final */
  double* vx; double* vz;      /* [G] Variable.posterior_v (base.py:167-170) */
  int32_t* active;             /* [G] 1 = iterate; cleared when a problem stops */
  int32_t* flags;              /* [G] TRB_FLAG_* */
  int32_t* n_iter;             /* [G] iterations completed (accumulates) */
  double* rec_vx; double* rec_vz; int32_t max_records;  /* [max_records, G] or NULL */
  /* EarlyStopping(ids, tol, min_variance, wait_increase, max_increase)
   * (callbacks.py:192-243); enabled if es_tol >= 0.  es_vars: bit 0 = x, bit 1 = z. */
  double es_tol, es_min_variance, es_max_increase;
  int32_t es_wait_increase, es_vars;
  trb_quadrature quad;
} trb_se;
size_t trb_sizeof_se(void);

/* Run `n_iter` SE iterations of every active problem; records go to rows it0,
 * it0+1, ...  A NaN message (message_passing.py:187-198) or a failed EarlyStopping
 * test rolls the problem back to the end of its previous iteration
 * (`old_message_dag`) and clears its active flag. */
int trb_se_run(const trb_se* se, int it0, int n_iter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRAMP_B200_H */
