// Shared device/host helpers for the tramp_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/tramp_b200.h"

// ---------------------------------------------------------------- host side
extern thread_local char trb_err_buf[512];
int trb_set_error(int code, const char* fmt, ...);

#define TRB_CHECK_ARG(cond, what)                                              \
  do {                                                                         \
    if (!(cond)) return trb_set_error(TRB_ERR_INVALID, "%s: %s", __func__, what); \
  } while (0)

#define TRB_CHECK_LAUNCH()                                                     \
  do {                                                                         \
    cudaError_t e_ = cudaGetLastError();                                       \
    if (e_ != cudaSuccess)                                                     \
      return trb_set_error(TRB_ERR_CUDA, "%s: %s", __func__, cudaGetErrorString(e_)); \
  } while (0)

int trb_sm_count_cached();

// Launch accounting / optional per-kernel CUDA-event timing (trb_profile_*).
// kind: 0 = elementwise / update kernels, 1 = GEMV (project / expand).
void trb_note_launch(int kind, cudaStream_t st, bool before);
struct trb_launch_scope {
  int kind;
  cudaStream_t st;
  trb_launch_scope(int k, cudaStream_t s) : kind(k), st(s) { trb_note_launch(kind, st, true); }
  ~trb_launch_scope() { trb_note_launch(kind, st, false); }
};

// --------------------------------------------------------------- device side
namespace trb {

constexpr double kTwoPi = 6.283185307179586476925286766559;

__device__ __forceinline__ uint32_t smem_u32_(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block; result valid in every thread.  `sh` holds >= 33 doubles.
// Safe to call repeatedly with the same `sh` (leading barrier).
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = (lane < nwarp) ? sh[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

// Sums K values over the block with one barrier round; results valid in every
// thread.  `sh` holds >= 33*K doubles.
template <int K>
__device__ __forceinline__ void block_sum_n(double (&v)[K], double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) sh[k * 33 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double t = (lane < nwarp) ? sh[k * 33 + lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) sh[k * 33 + 32] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = sh[k * 33 + 32];
}

// ---- thread-block clusters: one instance may span C CTAs (grid = (C, B), cluster
// = (C, 1, 1)); a kernel launched without the cluster attribute sees C = 1.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the double at the address of local shared variable `p` in CTA `rank` of the cluster (DSMEM)
__device__ __forceinline__ double dsmem_ld(const double* p, uint32_t rank) {
  uint32_t remote;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32_(p)), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
  return v;
}
__device__ __forceinline__ int dsmem_ld_int(const int* p, uint32_t rank) {
  uint32_t remote;
  int v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32_(p)), "r"(rank));
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
  return v;
}

// Sums K values over all CTAs of the cluster (over the block when C = 1), in
// CTA-rank order, so every CTA gets the same bits.  `sh` holds >= 33*K doubles.
template <int K>
__device__ __forceinline__ void cluster_sum_n(double (&v)[K], double* sh) {
  block_sum_n<K>(v, sh);  // leaves the block totals in sh[k*33 + 32]
  const uint32_t C = cluster_nctarank();
  if (C == 1) return;
  cluster_sync();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double t = 0.0;
    for (uint32_t r = 0; r < C; ++r) t += dsmem_ld(&sh[k * 33 + 32], r);
    v[k] = t;
  }
  cluster_sync();  // nobody reuses sh while a peer is still reading it
}
__device__ __forceinline__ double cluster_sum(double v, double* sh) {
  double a[1] = {v};
  cluster_sum_n<1>(a, sh);
  return a[0];
}

__device__ __forceinline__ int block_or(int v, int* sh) {
  v = __reduce_or_sync(0xffffffffu, v);
  __syncthreads();
  if (threadIdx.x == 0) *sh = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && v) atomicOr(sh, v);
  __syncthreads();
  return *sh;
}
// OR over all CTAs of the cluster
__device__ __forceinline__ int cluster_or(int v, int* sh) {
  int all = block_or(v, sh);
  const uint32_t C = cluster_nctarank();
  if (C == 1) return all;
  cluster_sync();
  all = 0;
  for (uint32_t r = 0; r < C; ++r) all |= dsmem_ld_int(sh, r);
  cluster_sync();
  return all;
}

// streaming 16-byte load that does not pollute L1 (operators are read once)
__device__ __forceinline__ double2 ldg_stream(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
               : "=d"(r.x), "=d"(r.y)
               : "l"(p));
  return r;
}

// scalar version
__device__ __forceinline__ double ldg_stream1(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

// base.py:44-46 + 250-255.  np.maximum / np.clip propagate NaN whereas CUDA
// fmax/fmin drop it, so NaN is routed around them: a NaN variance must surface
// as a NaN message so that the check of message_passing.py:187-209 fires.
__device__ __forceinline__ double clip_a_new(double v, double a, double amin, double amax) {
  double vv = (v != v) ? v : fmax(v, 1e-20);
  double an = 1.0 / vv - a;
  if (an == an) an = fmin(fmax(an, amin), amax);
  return an;
}

// EarlyStopping (callbacks.py:206-243) on the posterior variances of the tracked
// variables (vars: bit 0 = x, bit 1 = z).  Returns 0 (go on), TRB_FLAG_CONVERGED
// (stop, keep the state) or TRB_FLAG_DIVERGED (stop, roll back); *tol_out receives
// max |dv| (NaN when there is no previous value yet).
__device__ __forceinline__ int early_stopping_variance(int vars, int it, double vx, double vz,
                                                       double vx_old, double vz_old, double es_tol,
                                                       double es_min_variance, double es_max_increase,
                                                       int es_wait_increase, double* tol_out) {
  const bool ux = vars & 1, uz = vars & 2;
  *tol_out = nan("");
  if ((ux && vx < es_min_variance) || (uz && vz < es_min_variance)) return TRB_FLAG_CONVERGED;
  if ((ux && vx != vx) || (uz && vz != vz)) return TRB_FLAG_DIVERGED;
  if (it == 0) return 0;  // old_vs is None in the first call
  double tol = 0.0, inc = -INFINITY;
  if (ux) {
    tol = fmax(tol, fabs(vx_old - vx));
    inc = fmax(inc, vx - vx_old);
  }
  if (uz) {
    tol = fmax(tol, fabs(vz_old - vz));
    inc = fmax(inc, vz - vz_old);
  }
  *tol_out = tol;
  if (tol < es_tol) return TRB_FLAG_CONVERGED;
  if (it > es_wait_increase && inc > es_max_increase) return TRB_FLAG_DIVERGED;
  return 0;
}

// message_passing.py:119-127; `if not damping: return data`
__device__ __forceinline__ double damp(double d, double old_v, double new_v) {
  return (d != 0.0) ? d * old_v + (1.0 - d) * new_v : new_v;
}

// ---- mbarrier + 1-D bulk copy (TMA) PTX wrappers --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion signalled on `bar` (complete_tx).
// dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace trb

// ---- peer-memory exchange of a row-sharded operator (trb_comm.cu) ------------
struct trb_peers {
  int n;                                          // ranks; 0 = not sharded
  unsigned long long seq;                         // exchange the consumer waits for
  const double* data[TRB_MAX_RANKS];              // every rank's vector of this exchange, in LOCAL memory
  unsigned long long* flags_of[TRB_MAX_RANKS];    // every rank's flag array (publish side)
  const unsigned long long* my_flags;             // this rank's flag array (wait side)
};
struct trb_push {
  int n;
  double* dst[TRB_MAX_RANKS];                     // this rank's slot in every rank's buffer
  // publish from the pushing kernel itself (its last CTA), saving a launch: the
  // sequence number goes to flags_of[r][rank]; counter is a zeroed device int
  unsigned long long* flags_of[TRB_MAX_RANKS];
  unsigned long long seq;
  int rank;
  unsigned int* counter;                          // NULL: the caller publishes separately
};
void trb_comm_push_targets(trb_comm* c, trb_push* push);
size_t trb_comm_capacity(const trb_comm* c);
int trb_comm_publish(trb_comm* c, trb_peers* peers, cudaStream_t st);
// like trb_comm_push_targets + trb_comm_publish, but the pushing kernel publishes (no signal launch)
void trb_comm_begin_exchange(trb_comm* c, trb_push* push);
const trb_peers* trb_comm_last(const trb_comm* c);

namespace trb {

// Block-wide wait until every rank has published exchange p.seq.  Returns false
// if a peer stayed silent for ~1 s (the caller flags the instance and goes on
// rather than hanging the GPU).
__device__ __forceinline__ bool peers_wait(const trb_peers& p) {
  __shared__ int peers_ok;
  if (threadIdx.x == 0) peers_ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < p.n) {
    const unsigned long long* f = p.my_flags + threadIdx.x;
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= p.seq) break;
      if (clock64() - t0 > 2000000000LL) {
        peers_ok = 0;
        break;
      }
    }
  }
  __syncthreads();
  return peers_ok != 0;
}
// sum over the ranks, in rank order, of element i of the exchanged vectors (the
// peers wrote them into this rank's memory; L2 is the point of coherence)
__device__ __forceinline__ double peers_sum(const trb_peers& p, size_t i) {
  double s = 0.0;
  for (int r = 0; r < p.n; ++r) s += __ldcg(p.data[r] + i);
  return s;
}

// Even split of T work items over G workers: worker k owns [part_begin(k), part_begin(k+1)).
__host__ __device__ __forceinline__ int64_t part_begin(int64_t k, int64_t T, int64_t G) {
  return (k * T) / G;
}
// worker that owns item g
__host__ __device__ __forceinline__ int64_t part_owner(int64_t g, int64_t T, int64_t G) {
  return ((g + 1) * G - 1) / T;
}

}  // namespace trb

// Number of CTAs (a thread-block cluster; 8 is the portable maximum, 16 is used for
// a very large single instance) that share
// one instance in the per-instance update kernels: 1 when the batch alone fills
// the GPU, more for a few large instances (BASELINE config 5: B = 1, N = 65536).
int trb_cluster_size(int B, int n);

// opt a kernel into non-portable (16-CTA) clusters, once; false if the device refuses
bool trb_allow_big_cluster(const void* kernel);

// kernel<<<(C, B), threads, 0, st>>> with cluster dimension (C, 1, 1).  C = 16 is
// the non-portable maximum: opted into per kernel, falling back to 8 if refused.
template <typename... KArgs, typename... Args>
cudaError_t trb_launch_cluster(void (*kernel)(KArgs...), int C, int B, int threads, cudaStream_t st,
                               Args... args) {
  if (C > 8 && !trb_allow_big_cluster(reinterpret_cast<const void*>(kernel))) C = 8;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, B, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  if (e != cudaSuccess && C > 8) {  // the device refused a 16-CTA cluster
    cudaGetLastError();
    cfg.gridDim = dim3(8, B, 1);
    attr[0].val.clusterDim.x = 8;
    e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  }
  return e;
}

// Launch geometry shared by trb_lin_expand and the kernels that reduce its slots.
struct trb_expand_geom {
  int G;       // number of CTAs (workers) over the B*R row space
  int nslots;  // max number of workers whose range touches one instance
};
trb_expand_geom trb_expand_geometry(int B, int R);
// part[b, 0, :] = sum of the slots of instance b (in place)
int trb_reduce_slots_inplace(int B, int R, int n, int ld, double* part, void* stream);
// the update kernels add up to this many slots themselves; beyond it (few instances
// spread over many CTAs) the sweep reduces the slots first
constexpr int kTrbDirectSlots = 4;
