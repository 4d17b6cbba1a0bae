// LinearChannel set-up: batched symmetric eigen-decomposition / thin SVD by one-sided
// BLOCK JACOBI (Hestenes) on FP64 tensor cores (DMMA, mma.sync m8n8k4 f64).
//
// reference: channels/linear/linear_channel.py:8-15 (`svd`: np.linalg.svd, LAPACK gesdd)
// and :36-46 (`matrix_rank`, `spectrum`, `singular`); examples/figures/benchmark.py:22
// counts this factorisation in EP's total time.
//
// What is factorised.  `A[B, np, ld]` holds np VECTORS of length <= ld per instance, one
// per row.  Two uses (tramp_b200/channels/linear_channel.py):
//   * A = G = W W^T (or W^T W, whichever is smaller; rows of a symmetric matrix): the
//     rotated rows converge to lambda_i u_i^T, i.e. row norms are the eigenvalues of G
//     (= squared singular values of W) and the normalised rows the singular vectors of
//     the short side; the other side follows by one DMMA GEMM (trb_gemm.cu).
//   * A = W (rows of the short side): the rotated rows converge to s_i v_i^T directly
//     (no squaring of the condition number).
// Both are the same iteration:  A <- Q^T A  with Q orthogonal, until the rows are
// mutually orthogonal.
//
// One ROUND of a sweep treats np/32 disjoint pairs of 16-row blocks (round-robin
// tournament, every pair of blocks meets once per sweep) with three launches over the
// WHOLE batch:
//   k_jacobi_gram   S = X X^T of the 32 rows X of a pair          (DMMA, 1-D TMA ring)
//   k_jacobi_eig    S = J diag J^T, cyclic Jacobi in shared memory (FP64 ALU, latency)
//   k_jacobi_rotate X <- J^T X in place                            (DMMA, 1-D TMA ring)
// Both GEMM-shaped kernels read the pair's rows through a ring of shared-memory stages
// filled by cp.async.bulk (mbarrier complete_tx), so the DMMA warps never wait on a
// global load; the rows of a pair cross HBM twice in and once out per round:
// arithmetic intensity 16/4 = 4 flop/B for the rotation (HBM-bound on B200 unless the
// instance stays in L2) and (10/16)*4 for the Gram (upper-triangular tiles only).
#include <stdlib.h>
#include "trb_common.cuh"

using namespace trb;

namespace {

constexpr int kBS = 16;                 // rows per block
constexpr int kPV = 2 * kBS;            // rows per pair
constexpr int kStagePos = 64;           // positions (columns) per ring stage
constexpr int kMmaWarps = 4;            // each owns 16 positions of a stage
constexpr int kSetupThreads = (kMmaWarps + 1) * 32;
constexpr int kRingStages = 4;
// Row strides of a stage in doubles.  The padding makes every LDS.128 of a fragment
// conflict-free: the Gram reads 2 adjacent rows x 4 adjacent 32-byte columns per quarter
// warp (row stride = 16 B mod 128), the rotation 4 adjacent rows x 2 adjacent 16-byte
// columns (row stride = 32 B mod 128).
constexpr int kGramStride = kStagePos + 2;
constexpr int kRotStride = kStagePos + 4;

// D(8x8) += A(8x4, row) * B(4x8, col).  lane = 4*g + t:
//   a = A[g][t], b = B[t][g], {c0, c1} = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// Round-robin tournament of nb (even) players: in round r (0 <= r < nb-1) pair k
// (0 <= k < nb/2) is (nb-1, r) for k = 0 and (r+k, r-k) mod (nb-1) otherwise.
__host__ __device__ __forceinline__ void rr_pair(int nb, int round, int k, int& p, int& q) {
  const int m = nb - 1;
  if (k == 0) {
    p = m;
    q = round % m;
  } else {
    p = (round + k) % m;
    q = (round - k + m) % m;
  }
}

// global row of local row v (0..31) of the pair (p, q)
__device__ __forceinline__ int pair_row(int p, int q, int v) {
  return (v < kBS) ? p * kBS + v : q * kBS + (v - kBS);
}

__device__ __forceinline__ void mma_bar() {
  asm volatile("bar.sync 1, %0;" ::"n"(kMmaWarps * 32) : "memory");
}

// Producer warp: streams positions [c0, c1) of the 32 rows of the pair into the ring,
// one 512-byte bulk copy per row and stage (lane r copies row r).
template <int STRIDE, int STAGES = kRingStages>
__device__ __forceinline__ void produce_pair(const double* Ab, int ld, int p, int q, int c0, int c1,
                                             double* ring, uint64_t* full_bar, uint64_t* empty_bar) {
  const int lane = threadIdx.x & 31;
  const double* src = Ab + (size_t)pair_row(p, q, lane) * ld;
  int stage = 0;
  uint32_t phase = 0;
  for (int c = c0; c < c1; c += kStagePos) {
    if (lane == 0) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      mbar_arrive_expect_tx(&full_bar[stage], kPV * kStagePos * 8u);
    }
    __syncwarp();
    bulk_g2s(ring + ((size_t)stage * kPV + lane) * STRIDE, src + c, kStagePos * 8u, &full_bar[stage]);
    if (++stage == STAGES) {
      stage = 0;
      phase ^= 1u;
    }
  }
}

template <int STAGES = kRingStages>
__device__ __forceinline__ void ring_init(uint64_t* full_bar, uint64_t* empty_bar) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kMmaWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
}

// ------------------------------------------------------------------ Gram of a pair
// grid (pairs, B, zsplit).  S[b, pair, z] (32 x 32, row-major, both triangles) = X X^T over
// the positions of chunk z.
__global__ void __launch_bounds__(kSetupThreads, 3)
k_jacobi_gram(const double* __restrict__ A, int64_t strideA, int ld, int nb, int round, int chunk,
              double* __restrict__ S) {
  extern __shared__ __align__(128) double ring[];  // kRingStages * 32 * kGramStride; reused for the reduction
  __shared__ __align__(8) uint64_t full_bar[kRingStages];
  __shared__ __align__(8) uint64_t empty_bar[kRingStages];
  ring_init(full_bar, empty_bar);

  int p, q;
  rr_pair(nb, round, blockIdx.x, p, q);
  const int b = blockIdx.y;
  const int c0 = blockIdx.z * chunk;
  const int c1 = min(ld, c0 + chunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Ab = A + (size_t)b * strideA;

  if (warp == kMmaWarps) {
    produce_pair<kGramStride>(Ab, ld, p, q, c0, c1, ring, full_bar, empty_bar);
    return;
  }
  const int g = lane >> 2, t = lane & 3;
  // upper-triangular 8x8 tiles (I <= J): 10 accumulator pairs
  double acc[10][2];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k][0] = acc[k][1] = 0.0;
  int stage = 0;
  uint32_t phase = 0;
  for (int c = c0; c < c1; c += kStagePos) {
    mbar_wait(&full_bar[stage], phase);
    const double* st = ring + (size_t)stage * kPV * kGramStride + 16 * warp + 4 * t;
    double x[4][4];
#pragma unroll
    for (int I = 0; I < 4; ++I) {
      const double2 lo = *reinterpret_cast<const double2*>(st + (8 * I + g) * kGramStride);
      const double2 hi = *reinterpret_cast<const double2*>(st + (8 * I + g) * kGramStride + 2);
      x[I][0] = lo.x, x[I][1] = lo.y, x[I][2] = hi.x, x[I][3] = hi.y;
    }
    // the same register is the A fragment of its row tile and the B fragment of its
    // column tile (X X^T); the 4 k-steps walk the thread's 4 consecutive positions
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      int k = 0;
#pragma unroll
      for (int I = 0; I < 4; ++I)
#pragma unroll
        for (int J = I; J < 4; ++J, ++k) dmma884(acc[k][0], acc[k][1], x[I][s], x[J][s]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    if (++stage == kRingStages) {
      stage = 0;
      phase ^= 1u;
    }
  }
  // ---- sum the 4 warps' partial tiles (the ring is drained: every stage was consumed)
  mma_bar();
  double* red = ring;  // [4][32][33]
  {
    int k = 0;
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
      for (int J = I; J < 4; ++J, ++k) {
        double* d = red + ((size_t)warp * kPV + 8 * I + g) * 33 + 8 * J + 2 * t;
        d[0] = acc[k][0];
        d[1] = acc[k][1];
      }
  }
  mma_bar();
  const int npairs = nb / 2;
  double* out = S + (((size_t)b * npairs + blockIdx.x) * gridDim.z + blockIdx.z) * (kPV * kPV);
  for (int e = tid; e < kPV * kPV; e += kMmaWarps * 32) {
    int i = e >> 5, j = e & 31;
    if ((i >> 3) > (j >> 3)) {  // lower tile: mirror
      const int tmp = i;
      i = j, j = tmp;
    }
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kMmaWarps; ++w) v += red[((size_t)w * kPV + i) * 33 + j];
    out[e] = v;
  }
}

// --------------------------------------------------- eigenvectors of a pair's Gram
// grid (pairs, B), 256 threads.  Cyclic two-sided Jacobi on the 32 x 32 matrix in shared
// memory with the round-robin parallel ordering: in a step the 16 disjoint index pairs
// are rotated at once, thread (k, m) updating the 2 x 2 block (rows of pair k, columns of
// pair m) of S and two rows of J.  Rotations are the small-angle ones (|tan| <= 1), so J
// stays close to the identity once S is nearly diagonal and the outer sweeps converge
// quadratically.  Also records the largest cosine between two rows of the pair BEFORE
// the rotation (the sweep's convergence measure) and a skip flag.
__device__ __forceinline__ unsigned long long as_ull(double v) {
  return (unsigned long long)__double_as_longlong(v);
}

// Shared-memory working set of the 32 x 32 eigen-problem of a pair
struct EigSmem {
  double S[kPV][kPV + 1];
  double J[kPV][kPV + 1];
  double c[kBS], s[kBS];
  double inv[kPV];
  double red[8];
  uchar2 pair[kPV - 1][kBS];  // the 31 x 16 index pairs of the inner round-robin, as (p, q) bytes
};

// Largest |cos| between two rows of the pair whose Gram matrix is in m.S (256 threads; ends with a
// barrier).  Rows whose norm is below 1e-13 of the pair's largest are rounding noise (padding
// rows, the null directions of a rank-deficient matrix): their direction means nothing and they
// are left out of the convergence measure.
__device__ __forceinline__ double pair_max_cosine(EigSmem& m, int tid) {
  if (tid < kPV) {
    const double d = m.S[tid][tid];
    double dmax = d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    m.inv[tid] = (d > 1e-26 * dmax && d > 0.0) ? rsqrt(d) : 0.0;
  }
  __syncthreads();
  double off = 0.0;
  for (int e = tid; e < kPV * kPV; e += 256) {
    const int i = e >> 5, j = e & 31;
    if (i < j) off = fmax(off, fabs(m.S[i][j]) * m.inv[i] * m.inv[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) off = fmax(off, __shfl_xor_sync(0xffffffffu, off, o));
  if ((tid & 31) == 0) m.red[tid >> 5] = off;
  __syncthreads();
  off = m.red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) off = fmax(off, m.red[w]);
  return off;
}

// m.S <- J^T m.S J (nearly) diagonal, m.J <- J, by `max_inner` sweeps of cyclic two-sided Jacobi
// (256 threads; m.J must hold the identity on entry; ends with a barrier).
__device__ __forceinline__ void pair_eigenvectors(EigSmem& m, int tid, int max_inner) {
  for (int e = tid; e < (kPV - 1) * kBS; e += 256) {
    int p, q;
    rr_pair(kPV, e >> 4, e & 15, p, q);
    m.pair[e >> 4][e & 15] = make_uchar2((unsigned char)p, (unsigned char)q);
  }
  __syncthreads();
  const int k = tid >> 4, mm = tid & 15;
  for (int sweep = 0; sweep < max_inner; ++sweep) {
    int rotated = 0;
    for (int step = 0; step < kPV - 1; ++step) {
      const uchar2 ik = m.pair[step][k], im = m.pair[step][mm];
      const int pk = ik.x, qk = ik.y, pm = im.x, qm = im.y;
      if (tid < kBS) {
        const int p = pm, q = qm;  // tid < 16: mm == tid
        const double app = m.S[p][p], aqq = m.S[q][q], apq = m.S[p][q];
        double c = 1.0, s = 0.0;
        if (apq * apq > 1.21e-32 * fabs(app * aqq) && apq != 0.0) {
          // small root of t^2 + 2 tau t - 1 = 0, tau = (aqq - app) / (2 apq), written without
          // forming tau: t = sgn(d h) |h| / (|d| + sqrt(d^2 + h^2)).  Only c has to be exact to
          // the last bit (c^2 + s^2 = c^2 (1 + t^2) = 1 keeps J orthogonal); an ulp or two in t
          // merely leaves |apq| * 1e-16 un-annihilated.  (An FP32 tangent was tried to shorten this
          // serial step: no gain on B200, and one more outer sweep.)
          const double d = aqq - app, h = 2.0 * apq;
          const double tt = copysign(fabs(h) / (fabs(d) + sqrt(fma(d, d, h * h))), (d >= 0.0) ? h : -h);
          c = rsqrt(fma(tt, tt, 1.0));
          s = tt * c;
          rotated = 1;
        }
        m.c[tid] = c;
        m.s[tid] = s;
      }
      __syncthreads();
      const double ck = m.c[k], sk = m.s[k], cm = m.c[mm], sm = m.s[mm];
      // S' = R_k^T S R_m on the 2x2 block, R = [[c, s], [-s, c]]
      const double a00 = m.S[pk][pm], a01 = m.S[pk][qm], a10 = m.S[qk][pm], a11 = m.S[qk][qm];
      const double t00 = ck * a00 - sk * a10, t01 = ck * a01 - sk * a11;
      const double t10 = sk * a00 + ck * a10, t11 = sk * a01 + ck * a11;
      double n00 = cm * t00 - sm * t01, n01 = sm * t00 + cm * t01;
      double n10 = cm * t10 - sm * t11, n11 = sm * t10 + cm * t11;
      if (k == mm && (ck != 1.0 || sk != 0.0)) n01 = n10 = 0.0;  // annihilated by construction
      // J' = J R_m on rows 2k, 2k+1
      const double j00 = m.J[2 * k][pm], j01 = m.J[2 * k][qm], j10 = m.J[2 * k + 1][pm], j11 = m.J[2 * k + 1][qm];
      m.S[pk][pm] = n00, m.S[pk][qm] = n01, m.S[qk][pm] = n10, m.S[qk][qm] = n11;
      m.J[2 * k][pm] = cm * j00 - sm * j01, m.J[2 * k][qm] = sm * j00 + cm * j01;
      m.J[2 * k + 1][pm] = cm * j10 - sm * j11, m.J[2 * k + 1][qm] = sm * j10 + cm * j11;
      __syncthreads();
    }
    if (sweep + 1 < max_inner && !__syncthreads_or(rotated)) break;
  }
}

__global__ void __launch_bounds__(256, 7)
k_jacobi_eig(const double* __restrict__ S, int zsplit, double* __restrict__ Jm, int* __restrict__ rot_flag,
             unsigned long long* __restrict__ offmax, double skip_tol, int max_inner) {
  __shared__ EigSmem m;
  const int tid = threadIdx.x;
  const int pair = blockIdx.x, b = blockIdx.y, npairs = gridDim.x;
  const double* in = S + ((size_t)b * npairs + pair) * zsplit * (kPV * kPV);
  for (int e = tid; e < kPV * kPV; e += 256) {
    double v = 0.0;
    for (int z = 0; z < zsplit; ++z) v += in[(size_t)z * (kPV * kPV) + e];
    m.S[e >> 5][e & 31] = v;
    m.J[e >> 5][e & 31] = ((e >> 5) == (e & 31)) ? 1.0 : 0.0;
  }
  __syncthreads();
  const double off = pair_max_cosine(m, tid);
  if (tid == 0) {
    atomicMax(offmax + b, as_ull(off));  // non-negative doubles order like their bit patterns
    rot_flag[(size_t)b * npairs + pair] = (off > skip_tol) ? 1 : 0;
  }
  if (!(off > skip_tol)) return;
  pair_eigenvectors(m, tid, max_inner);
  double* out = Jm + ((size_t)b * npairs + pair) * (kPV * kPV);
  for (int e = tid; e < kPV * kPV; e += 256) out[e] = m.J[e >> 5][e & 31];
}

// ------------------------------------------------- Gram + eigenvectors in one kernel
// grid (pairs, B, zsplit), 256 threads: warps 0-3 are the DMMA consumers of k_jacobi_gram, warp 4
// the TMA producer, warps 5-7 only join for the eigen-solve.  The CTA that finishes the pair's
// Gram matrix (the only one when zsplit = 1, else the last of the pair's zsplit CTAs to arrive:
// an arrival counter per pair, partial sums added in chunk order, so the result does not depend
// on which CTA is last) goes straight on to the 32 x 32 Jacobi.  That kernel is bound by
// instruction issue and FP64 latency, this one by HBM: with three CTAs per SM in different phases
// the eigen-solves were meant to run in the shadow of the other CTAs' streaming.  Measured: no gain
// (see g_jacobi_fused), so the variant is kept as an option (trb_jacobi_set_fused bit 1), off by
// default.
constexpr int kGeStages = 3;
__global__ void __launch_bounds__(256, 3)
k_jacobi_gram_eig(const double* __restrict__ A, int64_t strideA, int ld, int nb, int round, int chunk,
                  double* __restrict__ S, unsigned int* __restrict__ arrivals, double* __restrict__ Jm,
                  int* __restrict__ rot_flag, unsigned long long* __restrict__ offmax, double skip_tol,
                  int max_inner) {
  extern __shared__ __align__(128) double ring[];  // kGeStages * 32 * kGramStride; reused for the reduction
  __shared__ EigSmem m;
  __shared__ __align__(8) uint64_t full_bar[kGeStages];
  __shared__ __align__(8) uint64_t empty_bar[kGeStages];
  __shared__ int s_last;
  ring_init<kGeStages>(full_bar, empty_bar);

  int p, q;
  rr_pair(nb, round, blockIdx.x, p, q);
  const int b = blockIdx.y, npairs = nb / 2, zsplit = gridDim.z;
  const int c0 = blockIdx.z * chunk;
  const int c1 = min(ld, c0 + chunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Ab = A + (size_t)b * strideA;
  const int g = lane >> 2, t = lane & 3;
  double acc[10][2];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k][0] = acc[k][1] = 0.0;
  if (warp == kMmaWarps) {
    produce_pair<kGramStride, kGeStages>(Ab, ld, p, q, c0, c1, ring, full_bar, empty_bar);
  } else if (warp < kMmaWarps) {
    int stage = 0;
    uint32_t phase = 0;
    for (int c = c0; c < c1; c += kStagePos) {
      mbar_wait(&full_bar[stage], phase);
      const double* st = ring + (size_t)stage * kPV * kGramStride + 16 * warp + 4 * t;
      double x[4][4];
#pragma unroll
      for (int I = 0; I < 4; ++I) {
        const double2 lo = *reinterpret_cast<const double2*>(st + (8 * I + g) * kGramStride);
        const double2 hi = *reinterpret_cast<const double2*>(st + (8 * I + g) * kGramStride + 2);
        x[I][0] = lo.x, x[I][1] = lo.y, x[I][2] = hi.x, x[I][3] = hi.y;
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        int k = 0;
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
          for (int J = I; J < 4; ++J, ++k) dmma884(acc[k][0], acc[k][1], x[I][s], x[J][s]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == kGeStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  }
  __syncthreads();  // every stage was consumed: the ring is free for the reduction
  double* red = ring;  // [4][32][33]
  if (warp < kMmaWarps) {
    int k = 0;
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
      for (int J = I; J < 4; ++J, ++k) {
        double* d = red + ((size_t)warp * kPV + 8 * I + g) * 33 + 8 * J + 2 * t;
        d[0] = acc[k][0];
        d[1] = acc[k][1];
      }
  }
  __syncthreads();
  const size_t pair_id = (size_t)b * npairs + blockIdx.x;
  for (int e = tid; e < kPV * kPV; e += 256) {
    int i = e >> 5, j = e & 31;
    if ((i >> 3) > (j >> 3)) {  // lower tile: mirror
      const int tmp = i;
      i = j, j = tmp;
    }
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kMmaWarps; ++w) v += red[((size_t)w * kPV + i) * 33 + j];
    if (zsplit == 1) m.S[e >> 5][e & 31] = v;
    else S[(pair_id * zsplit + blockIdx.z) * (kPV * kPV) + e] = v;
  }
  if (zsplit > 1) {
    __threadfence();  // the partial Gram is visible before the arrival is
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(arrivals + pair_id, 1u) == (unsigned)(zsplit - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int e = tid; e < kPV * kPV; e += 256) {
      double v = 0.0;
      for (int z = 0; z < zsplit; ++z) v += __ldcg(S + (pair_id * zsplit + z) * (kPV * kPV) + e);
      m.S[e >> 5][e & 31] = v;
    }
    if (tid == 0) arrivals[pair_id] = 0;  // ready for the next round
  }
  for (int e = tid; e < kPV * kPV; e += 256) m.J[e >> 5][e & 31] = ((e >> 5) == (e & 31)) ? 1.0 : 0.0;
  __syncthreads();
  const double off = pair_max_cosine(m, tid);
  if (tid == 0) {
    atomicMax(offmax + b, as_ull(off));
    rot_flag[pair_id] = (off > skip_tol) ? 1 : 0;
  }
  if (!(off > skip_tol)) return;
  pair_eigenvectors(m, tid, max_inner);
  double* out = Jm + pair_id * (kPV * kPV);
  for (int e = tid; e < kPV * kPV; e += 256) out[e] = m.J[e >> 5][e & 31];
}

// ---------------------------------------------------------- a whole round in one kernel
// Short rows (ld <= kFusedMaxLd): the 32 rows of a pair fit in the shared memory of one CTA, so
// Gram, eigenvectors and rotation are ONE launch per round and the rows cross HBM once in, once
// out.  grid (pairs, B), 256 threads = 8 DMMA warps; warp w owns the 16-position units w, w + 8, ...
constexpr int kFusedMaxLd = 768;
__host__ __device__ constexpr int fused_stride(int ld) { return ld + 4; }  // = 32 B mod 128: see kRotStride

__global__ void __launch_bounds__(256, 1)
k_jacobi_round_fused(double* __restrict__ A, int64_t strideA, int ld, int nb, int round,
                     unsigned long long* __restrict__ offmax, double skip_tol, int max_inner) {
  extern __shared__ __align__(128) double X[];  // 32 rows x fused_stride(ld)
  __shared__ EigSmem m;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int XS = fused_stride(ld);
  int p, q;
  rr_pair(nb, round, blockIdx.x, p, q);
  double* Ab = A + (size_t)b * strideA;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int e = tid; e < kPV * kPV; e += 256) {
    m.S[e >> 5][e & 31] = 0.0;
    m.J[e >> 5][e & 31] = ((e >> 5) == (e & 31)) ? 1.0 : 0.0;
  }
  __syncthreads();
  if (warp == 0) {  // one bulk copy per row
    if (lane == 0) mbar_arrive_expect_tx(&bar, (uint32_t)(kPV * ld * 8));
    __syncwarp();
    bulk_g2s(X + (size_t)lane * XS, Ab + (size_t)pair_row(p, q, lane) * ld, (uint32_t)(ld * 8), &bar);
  }
  mbar_wait(&bar, 0);
  const int g = lane >> 2, t = lane & 3;
  const int units = ld / 16;
  {
    // ---- Gram: upper-triangular 8 x 8 tiles, as in k_jacobi_gram
    double acc[10][2];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k][0] = acc[k][1] = 0.0;
    for (int u = warp; u < units; u += 8) {
      const double* st = X + 16 * u + 4 * t;
      double x[4][4];
#pragma unroll
      for (int I = 0; I < 4; ++I) {
        const double2 lo = *reinterpret_cast<const double2*>(st + (size_t)(8 * I + g) * XS);
        const double2 hi = *reinterpret_cast<const double2*>(st + (size_t)(8 * I + g) * XS + 2);
        x[I][0] = lo.x, x[I][1] = lo.y, x[I][2] = hi.x, x[I][3] = hi.y;
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        int k = 0;
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
          for (int J = I; J < 4; ++J, ++k) dmma884(acc[k][0], acc[k][1], x[I][s], x[J][s]);
      }
    }
    // the warps add their partial tiles one after the other: a fixed order, bit-reproducible
    for (int w = 0; w < 8; ++w) {
      if (warp == w) {
        int k = 0;
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
          for (int J = I; J < 4; ++J, ++k) {
            m.S[8 * I + g][8 * J + 2 * t] += acc[k][0];
            m.S[8 * I + g][8 * J + 2 * t + 1] += acc[k][1];
          }
      }
      __syncthreads();
    }
  }
  for (int e = tid; e < kPV * kPV; e += 256) {  // mirror the lower tiles
    const int i = e >> 5, j = e & 31;
    if ((i >> 3) > (j >> 3)) m.S[i][j] = m.S[j][i];
  }
  __syncthreads();
  const double off = pair_max_cosine(m, tid);
  if (tid == 0) atomicMax(offmax + b, as_ull(off));
  if (!(off > skip_tol)) return;  // rows already orthogonal
  pair_eigenvectors(m, tid, max_inner);
  // ---- rotation X <- J^T X from shared memory straight to global memory (in place: the whole
  // pair has been read), as in k_jacobi_rotate
  double af[4][8];
#pragma unroll
  for (int I = 0; I < 4; ++I)
#pragma unroll
    for (int K = 0; K < 8; ++K) af[I][K] = m.J[4 * K + t][8 * I + g];
  for (int u = warp; u < units; u += 8) {
    const double* st = X + 16 * u + 2 * g;
    double2 bf[8];
#pragma unroll
    for (int K = 0; K < 8; ++K) bf[K] = *reinterpret_cast<const double2*>(st + (size_t)(4 * K + t) * XS);
    double acc[4][2][2];
#pragma unroll
    for (int I = 0; I < 4; ++I) acc[I][0][0] = acc[I][0][1] = acc[I][1][0] = acc[I][1][1] = 0.0;
#pragma unroll
    for (int K = 0; K < 8; ++K)
#pragma unroll
      for (int I = 0; I < 4; ++I) {
        dmma884(acc[I][0][0], acc[I][0][1], af[I][K], bf[K].x);
        dmma884(acc[I][1][0], acc[I][1][1], af[I][K], bf[K].y);
      }
#pragma unroll
    for (int I = 0; I < 4; ++I) {
      double2* o = reinterpret_cast<double2*>(Ab + (size_t)pair_row(p, q, 8 * I + g) * ld + 16 * u + 4 * t);
      o[0] = make_double2(acc[I][0][0], acc[I][1][0]);
      o[1] = make_double2(acc[I][0][1], acc[I][1][1]);
    }
  }
}

// --------------------------------------------------------------- rotation of a pair
// grid (pairs, B, zsplit).  X <- J^T X in place over the positions of chunk z.
__global__ void __launch_bounds__(kSetupThreads, 2)
k_jacobi_rotate(double* __restrict__ A, int64_t strideA, int ld, int nb, int round, int chunk,
                const double* __restrict__ Jm, const int* __restrict__ rot_flag) {
  extern __shared__ __align__(128) double ring[];  // kRingStages * 32 * kRotStride
  __shared__ __align__(8) uint64_t full_bar[kRingStages];
  __shared__ __align__(8) uint64_t empty_bar[kRingStages];
  const int npairs = nb / 2;
  const int b = blockIdx.y;
  if (!rot_flag[(size_t)b * npairs + blockIdx.x]) return;  // rows already orthogonal
  ring_init(full_bar, empty_bar);

  int p, q;
  rr_pair(nb, round, blockIdx.x, p, q);
  const int c0 = blockIdx.z * chunk;
  const int c1 = min(ld, c0 + chunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Ab = A + (size_t)b * strideA;

  if (warp == kMmaWarps) {
    produce_pair<kRotStride>(Ab, ld, p, q, c0, c1, ring, full_bar, empty_bar);
    return;
  }
  const int g = lane >> 2, t = lane & 3;
  // A fragments of J^T: tile (I, K) element [g][t] = J[4K + t][8I + g]
  const double* Jp = Jm + ((size_t)b * npairs + blockIdx.x) * (kPV * kPV);
  double af[4][8];
#pragma unroll
  for (int I = 0; I < 4; ++I)
#pragma unroll
    for (int K = 0; K < 8; ++K) af[I][K] = __ldg(Jp + (4 * K + t) * kPV + 8 * I + g);
  double* orow[4];
#pragma unroll
  for (int I = 0; I < 4; ++I) orow[I] = Ab + (size_t)pair_row(p, q, 8 * I + g) * ld + 16 * warp + 4 * t;

  int stage = 0;
  uint32_t phase = 0;
  for (int c = c0; c < c1; c += kStagePos) {
    mbar_wait(&full_bar[stage], phase);
    const double* st = ring + (size_t)stage * kPV * kRotStride + 16 * warp + 2 * g;
    double2 bf[8];
#pragma unroll
    for (int K = 0; K < 8; ++K) bf[K] = *reinterpret_cast<const double2*>(st + (4 * K + t) * kRotStride);
    double acc[4][2][2];
#pragma unroll
    for (int I = 0; I < 4; ++I) acc[I][0][0] = acc[I][0][1] = acc[I][1][0] = acc[I][1][1] = 0.0;
#pragma unroll
    for (int K = 0; K < 8; ++K)
#pragma unroll
      for (int I = 0; I < 4; ++I) {
        dmma884(acc[I][0][0], acc[I][0][1], af[I][K], bf[K].x);
        dmma884(acc[I][1][0], acc[I][1][1], af[I][K], bf[K].y);
      }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    // n-tile e holds positions 2n + e of the warp's 16: the thread's two tiles cover 4
    // consecutive positions {c0e0, c0e1, c1e0, c1e1}
#pragma unroll
    for (int I = 0; I < 4; ++I) {
      double2* o = reinterpret_cast<double2*>(orow[I] + c);
      o[0] = make_double2(acc[I][0][0], acc[I][1][0]);
      o[1] = make_double2(acc[I][0][1], acc[I][1][1]);
    }
    if (++stage == kRingStages) {
      stage = 0;
      phase ^= 1u;
    }
  }
}

// ------------------------------------------------------------------ finishing ops
// norms[b, i] = || A[b, i, :n] ||, one warp per row
__global__ void __launch_bounds__(256)
k_row_norms(const double* __restrict__ A, int64_t strideA, int rows, int n, int ld, double* __restrict__ norms) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), b = blockIdx.y, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const double* a = A + (size_t)b * strideA + (size_t)row * ld;
  double s = 0.0;
  for (int j = 2 * lane; j < n; j += 64) {
    const double2 v = *reinterpret_cast<const double2*>(a + j);
    s = fma(v.x, v.x, s);
    if (j + 1 < n) s = fma(v.y, v.y, s);
  }
  s = warp_sum(s);
  if (lane == 0) norms[(size_t)b * rows + row] = sqrt(s);
}

// dst[b, i, :n] = scale[b, i] * src[b, perm[b, i], :n]; columns n..ld_dst-1 are zeroed
__global__ void __launch_bounds__(256)
k_rows_gather_scale(const double* __restrict__ src, int64_t stride_src, int ld_src, const long long* __restrict__ perm,
                    const double* __restrict__ scale, int R, int n, double* __restrict__ dst, int64_t stride_dst,
                    int ld_dst) {
  const int i = blockIdx.x, b = blockIdx.y;
  const long long r = perm ? perm[(size_t)b * R + i] : i;
  const double sc = scale ? scale[(size_t)b * R + i] : 1.0;
  const double* s = src + (size_t)b * stride_src + (size_t)r * ld_src;
  double* d = dst + (size_t)b * stride_dst + (size_t)i * ld_dst;
  for (int j = threadIdx.x; j < ld_dst; j += blockDim.x) d[j] = (j < n) ? sc * s[j] : 0.0;
}

bool g_setup_attr_done = false;

int setup_attrs() {
  if (g_setup_attr_done) return TRB_OK;
  cudaError_t e = cudaFuncSetAttribute(k_jacobi_gram, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kRingStages * kPV * kGramStride * 8);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_jacobi_rotate, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kRingStages * kPV * kRotStride * 8);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_jacobi_gram_eig, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kGeStages * kPV * kGramStride * 8);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_jacobi_round_fused, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kPV * fused_stride(kFusedMaxLd) * 8);
  if (e != cudaSuccess) return trb_set_error(TRB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  g_setup_attr_done = true;
  return TRB_OK;
}

}  // namespace

// bit 0: one kernel per round for short rows and few pairs (default on); bit 1: Gram + eigenvectors in
// one kernel (default OFF: measured on B200 at 2048 x 2048, B = 64 / 16: 184.9 / 51.2 ms per sweep
// fused against 180.9 / 48.1 with the eigen-solves in a launch of their own -- at 3 CTAs per SM the
// 31 dependent steps of the inner Jacobi are latency-bound, ~25 us per pair, and leave the memory
// pipe idle about as long as the separate launch takes)
int g_jacobi_fused = 1;
extern "C" void trb_jacobi_set_fused(int mask) { g_jacobi_fused = mask & 3; }
int g_jacobi_waves = 4;  // measured on B200 (B = 16, 2048 x 2048): 4 waves 43.8 ms, 8 waves 47.0, 2 waves 44.8, 16 waves 52.9 per instance
extern "C" void trb_jacobi_set_waves(int waves) { g_jacobi_waves = waves > 0 ? waves : 4; }

extern "C" int trb_jacobi_zsplit(int B, int np, int ld) {
  // enough CTAs for ~g_jacobi_waves waves of 3 CTAs per SM, at least 4 stages per CTA
  const int npairs = np / kPV;
  const long long want = (long long)g_jacobi_waves * 3 * trb_sm_count_cached();
  long long z = (want + (long long)npairs * B - 1) / ((long long)npairs * B);
  const int zmax = (ld / kStagePos) / 4;
  if (z > zmax) z = zmax;
  if (z > 8) z = 8;
  if (z < 1) z = 1;
  return (int)z;
}

// all rounds of one sweep, enqueued on `st`
static int enqueue_jacobi_sweep(double* A, int64_t strideA, int B, int np, int ld, double* Swork, double* Jwork,
                                int* rot_flag, double* offmax, double skip_tol, int max_inner, cudaStream_t st) {
  const int nb = np / kBS, npairs = nb / 2;
  // Short rows AND few pairs (one wave of CTAs: the launch-bound regime): one launch per round with
  // the pair in shared memory.  With many pairs the three-kernel path is faster, because its
  // eigenvector kernel runs 8 CTAs per SM and hides the latency of the serial inner Jacobi, which
  // a CTA holding 100-200 KB of rows cannot (measured, 500 x 500: B = 1 / 8 / 64 instances 13.8 /
  // 1.96 / 1.66 ms each fused against 15.5 / 2.41 / 1.34 unfused).
  const int fused_per_sm = (int)(220000 / ((size_t)kPV * fused_stride(ld) * 8 + sizeof(EigSmem) + 64));
  if ((g_jacobi_fused & 1) && ld <= kFusedMaxLd &&
      (long long)npairs * B <= (long long)trb_sm_count_cached() * (fused_per_sm < 1 ? 1 : fused_per_sm)) {
    cudaMemsetAsync(offmax, 0, sizeof(double) * B, st);
    for (int round = 0; round < nb - 1; ++round) {
      trb_launch_scope scope_(2, st);
      k_jacobi_round_fused<<<dim3(npairs, B), 256, kPV * fused_stride(ld) * 8, st>>>(
          A, strideA, ld, nb, round, reinterpret_cast<unsigned long long*>(offmax), skip_tol, max_inner);
    }
    TRB_CHECK_LAUNCH();
    return TRB_OK;
  }
  // Swork holds up to trb_jacobi_zsplit() partial Grams per pair
  const int chunk = ((ld / kStagePos + trb_jacobi_zsplit(B, np, ld) - 1) / trb_jacobi_zsplit(B, np, ld)) * kStagePos;
  const int zsplit = (ld + chunk - 1) / chunk;
  const dim3 grid(npairs, B, zsplit);
  cudaMemsetAsync(offmax, 0, sizeof(double) * B, st);
  // rot_flag holds 2 * B * npairs ints: the rotate flags, then the per-pair arrival counters
  unsigned int* arrivals = reinterpret_cast<unsigned int*>(rot_flag + (size_t)B * npairs);
  if (g_jacobi_fused & 2) cudaMemsetAsync(arrivals, 0, sizeof(unsigned int) * (size_t)B * npairs, st);
  for (int round = 0; round < nb - 1; ++round) {
    if (g_jacobi_fused & 2) {  // Gram and eigenvectors in one launch
      trb_launch_scope scope_(2, st);
      k_jacobi_gram_eig<<<grid, 256, kGeStages * kPV * kGramStride * 8, st>>>(
          A, strideA, ld, nb, round, chunk, Swork, arrivals, Jwork, rot_flag,
          reinterpret_cast<unsigned long long*>(offmax), skip_tol, max_inner);
    } else {
      {
        trb_launch_scope scope_(2, st);
        k_jacobi_gram<<<grid, kSetupThreads, kRingStages * kPV * kGramStride * 8, st>>>(A, strideA, ld, nb, round,
                                                                                       chunk, Swork);
      }
      {
        trb_launch_scope scope_(2, st);
        k_jacobi_eig<<<dim3(npairs, B), 256, 0, st>>>(Swork, zsplit, Jwork, rot_flag,
                                                       reinterpret_cast<unsigned long long*>(offmax), skip_tol,
                                                       max_inner);
      }
    }
    {
      trb_launch_scope scope_(2, st);
      k_jacobi_rotate<<<grid, kSetupThreads, kRingStages * kPV * kRotStride * 8, st>>>(A, strideA, ld, nb, round, chunk,
                                                                                      Jwork, rot_flag);
    }
  }
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

// A sweep over a few small matrices is 3 (np/16 - 1) launches of a few microseconds each: launch
// latency, not HBM or the DMMA pipe, bounds it.  Such sweeps are captured once into a CUDA graph
// (on a private stream: torch's default stream is the legacy stream, which cannot be captured) and
// replayed -- every sweep of a factorisation has the same arguments.  The last graph is cached per
// host thread.
namespace {
struct JacobiGraphKey {
  double* A; int64_t strideA; int B, np, ld; double* S; double* J; int* flag; double* off; double skip_tol;
  int max_inner, waves, fused;
};
struct JacobiGraphCache {
  JacobiGraphKey key = {};
  cudaGraphExec_t exec = nullptr;
  cudaStream_t capture_stream = nullptr;
};
thread_local JacobiGraphCache g_jacobi_graph;
int g_jacobi_graphs = -1;
constexpr double kJacobiGraphMaxBytes = 48e6;  // rows of all instances: beyond it a round streams for > 20 us
}  // namespace

static int jacobi_sweep_graph(const JacobiGraphKey& key, cudaStream_t st) {
  JacobiGraphCache& gc = g_jacobi_graph;
  if (!gc.exec || memcmp(&gc.key, &key, sizeof(key)) != 0) {
    if (gc.exec) {
      cudaGraphExecDestroy(gc.exec);
      gc.exec = nullptr;
    }
    if (!gc.capture_stream && cudaStreamCreateWithFlags(&gc.capture_stream, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      return TRB_ERR_UNSUPPORTED;
    }
    if (cudaStreamBeginCapture(gc.capture_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return TRB_ERR_UNSUPPORTED;
    }
    const int rc = enqueue_jacobi_sweep(key.A, key.strideA, key.B, key.np, key.ld, key.S, key.J, key.flag, key.off,
                                        key.skip_tol, key.max_inner, gc.capture_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(gc.capture_stream, &graph);
    if (rc != TRB_OK || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return rc != TRB_OK ? rc : TRB_ERR_UNSUPPORTED;
    }
    const cudaError_t ie = cudaGraphInstantiate(&gc.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      gc.exec = nullptr;
      cudaGetLastError();
      return TRB_ERR_UNSUPPORTED;
    }
    gc.key = key;
  }
  const cudaError_t le = cudaGraphLaunch(gc.exec, st);
  if (le != cudaSuccess) return trb_set_error(TRB_ERR_CUDA, "trb_jacobi_sweep: graph launch: %s", cudaGetErrorString(le));
  return TRB_OK;
}

extern "C" int trb_jacobi_sweep(double* A, int64_t strideA, int B, int np, int ld, double* Swork, double* Jwork,
                                int* rot_flag, double* offmax, double skip_tol, int max_inner, void* stream) {
  TRB_CHECK_ARG(A && Swork && Jwork && rot_flag && offmax, "null pointer");
  TRB_CHECK_ARG(B > 0 && np >= kPV && np % kPV == 0, "np must be a positive multiple of 32");
  TRB_CHECK_ARG(ld >= kStagePos && ld % kStagePos == 0, "ld must be a positive multiple of 64");
  TRB_CHECK_ARG(strideA >= (int64_t)np * ld && (strideA % 2) == 0, "strideA too small");
  TRB_CHECK_ARG(((uintptr_t)A % 16) == 0, "A must be 16-byte aligned");
  TRB_CHECK_ARG(B <= 65535, "B > 65535");
  int rc = setup_attrs();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_jacobi_graphs < 0) {
    const char* e = getenv("TRB_CUDA_GRAPHS");
    g_jacobi_graphs = (e && e[0] == '0') ? 0 : 1;
  }
  if (g_jacobi_graphs && np > kPV && (double)B * np * ld * 8.0 <= kJacobiGraphMaxBytes) {
    JacobiGraphKey key = {};  // zero the padding: the key is compared bytewise
    key.A = A, key.strideA = strideA, key.B = B, key.np = np, key.ld = ld, key.S = Swork, key.J = Jwork;
    key.flag = rot_flag, key.off = offmax, key.skip_tol = skip_tol, key.max_inner = max_inner;
    key.waves = g_jacobi_waves;
    key.fused = g_jacobi_fused;
    rc = jacobi_sweep_graph(key, st);
    if (rc != TRB_ERR_UNSUPPORTED) return rc;  // else: plain launches
  }
  return enqueue_jacobi_sweep(A, strideA, B, np, ld, Swork, Jwork, rot_flag, offmax, skip_tol, max_inner, st);
}

extern "C" int trb_row_norms(const double* A, int64_t strideA, int B, int rows, int n, int ld, double* norms,
                             void* stream) {
  TRB_CHECK_ARG(A && norms, "null pointer");
  TRB_CHECK_ARG(B > 0 && B <= 65535 && rows > 0 && n > 0 && ld >= n && (ld % 2) == 0, "bad shape");
  TRB_CHECK_ARG(((uintptr_t)A % 16) == 0 && (strideA % 2) == 0, "A must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  trb_launch_scope scope_(2, st);
  k_row_norms<<<dim3((rows + 7) / 8, B), 256, 0, st>>>(A, strideA, rows, n, ld, norms);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_rows_gather_scale(const double* src, int64_t stride_src, int ld_src, const long long* perm,
                                     const double* scale, int B, int R, int n, double* dst, int64_t stride_dst,
                                     int ld_dst, void* stream) {
  TRB_CHECK_ARG(src && dst, "null pointer");
  TRB_CHECK_ARG(B > 0 && B <= 65535 && R > 0 && n > 0 && ld_src >= n && ld_dst >= n, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  trb_launch_scope scope_(2, st);
  k_rows_gather_scale<<<dim3(R, B), 256, 0, st>>>(src, stride_src, ld_src, perm, scale, R, n, dst, stride_dst, ld_dst);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}
