// LinearChannel passes for a batch that SHARES one operator: dense FP64
// tensor-core GEMMs (DMMA, mma.sync m8n8k4 f64 -- the only FP64 tensor shape
// sm_100a executes; tcgen05 has no FP64 kind).
//
// reference: channels/linear/linear_channel.py:69-89.  With one W for B
// instances the four thin-SVD operator passes of an EP iteration are
//   project:  T[B, R]   = X[B, n]    . A[R, n]^T     (U.T @ bx, V.T @ bz; :72-73)
//   expand :  O[B, n]   = C[B, R]    . A[R, n]       (V @ rz_svd, U-side of W @ rz; :78, :88)
// i.e. arithmetic intensity B/4 flop per operator byte instead of 1/4: the
// passes leave the HBM roofline and are bound by the FP64 tensor pipe.
//
// Both kernels: CTA tile BM x 128 (BM = 128 or 64 instances), k-tile 16, 8 MMA
// warps (2 x 4), warp tile (BM/2) x 32 built from m8n8k4 DMMAs.  Tiles are
// rasterised instance-tile fastest, so the CTAs resident at one time read the
// same operator panel and the operator crosses HBM once per pass (the 126 MB
// L2 serves the re-reads).
//   k_dgemm_dmma_tma (default): warp-group specialised.  A producer lane issues
//     2-D TMA tile loads (cp.async.bulk.tensor, 128-byte swizzle, hardware zero
//     fill of the ragged edges) into a 6-stage ring guarded by full/empty
//     mbarriers; the MMA warps never meet at a CTA barrier; setmaxnreg moves
//     the producer group's registers to the MMA groups.  Rows and k indices are
//     permuted between shared memory and the DMMA fragments (any permutation of
//     the summation index / of the rows is legal as long as both operands and
//     the epilogue agree) so that every fragment load is one conflict-free
//     LDS.128 feeding two DMMAs.
//   k_dgemm_dmma (fallback when an operand is not 16-byte aligned, e.g. odd R):
//     4-stage cp.async (LDGSTS, zero-filling) ring with padded rows.
// Summation order differs from the GEMV path (k-tiles of 4), which is well
// inside the 1e-9 parity bar; results are run-to-run deterministic.
#include <cuda.h>
#include "trb_common.cuh"

using namespace trb;

namespace {

constexpr int kBN = 128;       // output columns per CTA
constexpr int kBK = 16;        // k per pipeline stage
constexpr int kStages = 4;
constexpr int kThreads = 256;  // 8 warps: 2 (rows) x 4 (columns)
constexpr int kLdK = kBK + 4;  // padded k-extent of a k-contiguous tile row (bank-conflict free)
constexpr int kLdN = kBN + 4;  // padded n-extent of an n-contiguous tile row

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col).  lane = 4*g + t:
//   a = A[g][t], b = B[t][g], {c0, c1} = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Per-thread description of the cp.async chunks it copies every stage.  The
// (row, column) of each chunk is fixed, so pointers and row validity are set up
// once; per stage only the k-tail byte count changes.
//
// k-contiguous tile (X always; the operator for project): ROWS rows of kBK
// doubles, chunk c = tid + kThreads*q -> row c / CH, k offset (c % CH) * VEC.
template <int ROWS, int VEC>
struct KTileLoader {
  static constexpr int CH = kBK / VEC;
  static constexpr int Q = ROWS * CH / kThreads;
  static constexpr int RSTEP = kThreads / CH;  // rows between consecutive q
  const double* p;   // chunk q = 0 of k-tile 0
  int64_t qstride;   // elements between q and q + 1
  uint32_t so;       // shared-memory offset (doubles) of chunk q = 0 within a stage
  int kc;            // k offset of this thread's chunks inside a k-tile
  uint32_t row_ok;   // bit q: row of chunk q exists
  __device__ __forceinline__ void init(const double* src, int64_t ld, int row0, int nrows) {
    const int r = threadIdx.x / CH;
    kc = (threadIdx.x % CH) * VEC;
    row_ok = 0;
#pragma unroll
    for (int q = 0; q < Q; ++q)
      if (row0 + r + q * RSTEP < nrows) row_ok |= 1u << q;
    // rows beyond nrows are never dereferenced (src-size 0), keep p in bounds for q = 0
    const int rr = (row0 + r < nrows) ? row0 + r : row0;
    p = src + (int64_t)rr * ld + kc;
    qstride = (int64_t)RSTEP * ld;
    so = r * kLdK + kc;
  }
  __device__ __forceinline__ void load(double* stage, int k0, int K) const {
    int kv = K - k0 - kc;
    kv = kv < 0 ? 0 : (kv > VEC ? VEC : kv);
    const double* src = p + k0;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int bytes = ((row_ok >> q) & 1u) ? kv * 8 : 0;
      const double* g = bytes ? src + q * qstride : p;
      if constexpr (VEC == 2) cp_async16(stage + so + q * RSTEP * kLdK, g, bytes);
      else cp_async8(stage + so + q * RSTEP * kLdK, g, bytes);
    }
  }
};

// n-contiguous tile (the operator for expand): kBK rows (k) of kBN doubles,
// chunk c = tid + kThreads*q -> k row c / 64, column (c % 64) * 2.
struct NTileLoader {
  static constexpr int CH = kBN / 2;
  static constexpr int Q = kBK * CH / kThreads;
  static constexpr int RSTEP = kThreads / CH;
  const double* p;
  int64_t ld;
  uint32_t so;
  int r, cbytes;  // first k row of this thread; valid bytes of its column pair
  __device__ __forceinline__ void init(const double* src, int64_t ld_, int n0, int ncols) {
    r = threadIdx.x / CH;
    const int nc = (threadIdx.x % CH) * 2;
    int cv = ncols - n0 - nc;
    cv = cv < 0 ? 0 : (cv > 2 ? 2 : cv);
    cbytes = cv * 8;
    ld = ld_;
    p = src + (cv ? n0 + nc : 0);
    so = r * kLdN + nc;
  }
  __device__ __forceinline__ void load(double* stage, int k0, int K) const {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int k = k0 + r + q * RSTEP;
      const int bytes = (k < K) ? cbytes : 0;
      const double* g = bytes ? p + (int64_t)k * ld : p;
      cp_async16(stage + so + q * RSTEP * kLdN, g, bytes);
    }
  }
};

// C[Mg, Ng] = X[Mg, K] . Bop, X row-major with leading dimension ldx.
//   EXPAND = false: Bop[k][n] = A[n*lda + k]   (project, A rows are k-contiguous)
//   EXPAND = true : Bop[k][n] = A[k*lda + n]   (expand,  A rows are n-contiguous)
template <int BM, bool EXPAND, int XVEC>
__global__ void __launch_bounds__(kThreads, 1)
k_dgemm_dmma(const double* __restrict__ X, int64_t ldx, const double* __restrict__ A, int64_t lda,
             double* __restrict__ C, int64_t ldc, int Mg, int Ng, int K, int m_tiles) {
  constexpr int WTM = BM / 2;    // warp tile rows
  constexpr int MI = WTM / 8;    // m8 blocks per warp
  constexpr int NI = 4;          // n8 blocks per warp (warp tile 32 columns)
  constexpr int XS = BM * kLdK;  // doubles per stage, X tile
  constexpr int BS = EXPAND ? kBK * kLdN : kBN * kLdK;
  constexpr int KSTEPS = kBK / 4;
  extern __shared__ __align__(16) double smem[];
  double* Xs = smem;
  double* Bs = smem + kStages * XS;

  const int tile = blockIdx.x;
  const int m0 = (tile % m_tiles) * BM;  // instance tile fastest: neighbours share the operator panel
  const int n0 = (tile / m_tiles) * kBN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp & 1) * WTM, wn = (warp >> 1) * 32;

  KTileLoader<BM, XVEC> xl;
  xl.init(X, ldx, m0, Mg);
  KTileLoader<kBN, 2> bl_k;
  NTileLoader bl_n;
  if constexpr (EXPAND) bl_n.init(A, lda, n0, Ng);
  else bl_k.init(A, lda, n0, Ng);

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = (K + kBK - 1) / kBK;
  auto load_stage = [&](int s, int kt) {
    const int k0 = kt * kBK;
    xl.load(Xs + s * XS, k0, K);
    if constexpr (EXPAND) bl_n.load(Bs + s * BS, k0, K);
    else bl_k.load(Bs + s * BS, k0, K);
  };
  // fragment addresses of this thread inside a stage
  const int xo = (wm + g) * kLdK + t;
  const int bo = EXPAND ? t * kLdN + wn + g : (wn + g) * kLdK + t;
  auto load_frags = [&](const double* xs, const double* bs, int ks, double (&a)[MI], double (&b)[NI]) {
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = xs[xo + i * 8 * kLdK + ks * 4];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      if constexpr (EXPAND) b[j] = bs[bo + ks * 4 * kLdN + j * 8];
      else b[j] = bs[bo + j * 8 * kLdK + ks * 4];
    }
  };

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }
  double a[2][MI], b[2][NI];  // fragments, double buffered across k-steps
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<kStages - 2>();
    __syncthreads();  // tile kt landed; everyone is done with the stage refilled below
    const double* xs = Xs + (kt % kStages) * XS;
    const double* bs = Bs + (kt % kStages) * BS;
    load_frags(xs, bs, 0, a[0], b[0]);
    {
      const int nk = kt + kStages - 1;
      if (nk < KT) load_stage(nk % kStages, nk);
      cp_async_commit();
    }
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      if (ks + 1 < KSTEPS) load_frags(xs, bs, ks + 1, a[(ks + 1) & 1], b[(ks + 1) & 1]);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
          dmma884(acc[i][j][0], acc[i][j][1], a[ks & 1][i], b[ks & 1][j]);
    }
  }
  cp_async_wait<0>();
  // epilogue: C[row][col], col pairs (2t, 2t+1); rows of C may be only 8-byte aligned
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = m0 + wm + i * 8 + g;
    if (row >= Mg) continue;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int col = n0 + wn + j * 8 + 2 * t;
      double* dst = C + (int64_t)row * ldc + col;
      if (col < Ng) dst[0] = acc[i][j][0];
      if (col + 1 < Ng) dst[1] = acc[i][j][1];
    }
  }
}

// ============================================================ TMA + mbarrier
constexpr int kTmaStages = 6;
constexpr int kMmaWarps = 8;
constexpr int kTmaThreads = (kMmaWarps + 4) * 32;  // + producer warp group
constexpr int kRowBytes = kBK * 8;                 // 128: one swizzle-128B row

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// Shared-memory tiles are [rows][16 doubles] with CU_TENSOR_MAP_SWIZZLE_128B:
// 16-byte chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4) (tile base
// 1024-byte aligned).
//
// Fragment <-> tile mapping (lane = 4g + t, pg = perm(g) = 4*(g&1) + (g>>1)):
//  * k-contiguous tile (X; operator for project): m8/n8 block i, lane row g is
//    tile row 8i + pg; the k8 half h is one LDS.128 of chunk 4h + t, whose .x /
//    .y feed the DMMA steps e = 0 / 1, i.e. the step (h, e) sums k = 8h + 2t + e.
//    A quarter warp (g in {2q, 2q+1}, t = 0..3) then touches rows q and q + 4
//    and all eight chunk positions: conflict free.
//  * n-contiguous tile (operator for expand): 16-column blocks [16 k][16 n'];
//    for step (h, e) lane (g, t) loads chunk g of row k = 8h + 2t + e: .x / .y
//    are the B fragments of two interleaved n8 blocks (columns 2g and 2g + 1).
//
// C[Mg, Ng] = X[Mg, K] . Bop as in k_dgemm_dmma.
template <int BM, bool EXPAND, int MODE = 0>
__global__ void __launch_bounds__(kTmaThreads, 1)
k_dgemm_dmma_tma(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapA,
                 double* __restrict__ C, int64_t ldc, int Mg, int Ng, int K, int m_tiles) {
  constexpr int WTM = BM / 2;
  constexpr int MI = WTM / 8;
  constexpr int NI = 4;
  constexpr uint32_t XBYTES = BM * kRowBytes;
  constexpr uint32_t BBYTES = kBN * kRowBytes;  // project: [128 n][16 k]; expand: 8 x [16 k][16 n']
  constexpr uint32_t STAGE = XBYTES + BBYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kTmaStages];
  __shared__ __align__(8) uint64_t empty_bar[kTmaStages];

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kTmaStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kMmaWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int tile = blockIdx.x;
  const int m0 = (tile % m_tiles) * BM;
  const int n0 = (tile / m_tiles) * kBN;
  const int KT = (K + kBK - 1) / kBK;

  // Warp-group specialisation: warps 0-7 (two warp groups) run the MMA loop, the
  // third warp group is the producer (one lane issues the TMA loads, running up
  // to kTmaStages tiles ahead of the slowest MMA warp).  The register file is
  // re-split with setmaxnreg -- 384 threads launch with 168 registers each; the
  // producer group drops to 40 and the MMA groups grow to 232, which holds the
  // 128 accumulator registers without spilling.
  if (warp >= kMmaWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == kMmaWarps && lane == 0 && MODE != 2) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = 0; kt < KT; ++kt) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full_bar[stage], STAGE);
        uint8_t* dst = base_ptr + (size_t)stage * STAGE;
        tma_load_2d(dst, &mapX, kt * kBK, m0, &full_bar[stage]);
        if constexpr (EXPAND) {
#pragma unroll
          for (int blk = 0; blk < kBN / 16; ++blk)
            tma_load_2d(dst + XBYTES + blk * (16 * kRowBytes), &mapA, n0 + blk * 16, kt * kBK,
                        &full_bar[stage]);
        } else {
          tma_load_2d(dst + XBYTES, &mapA, kt * kBK, n0, &full_bar[stage]);
        }
        if (++stage == kTmaStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");

  // ------------------------------------------------------------ MMA warps
  const int g = lane >> 2, t = lane & 3;
  const int pg = ((g & 1) << 2) | (g >> 1);
  const int wm = (warp & 1) * WTM, wn = (warp >> 1) * 32;
  // byte offsets inside a stage
  const uint32_t xrow = (uint32_t)(wm + pg) * kRowBytes;          // + i * 8 rows
  const uint32_t xch0 = (uint32_t)((t ^ pg) << 4);                // chunk 4h + t: (^4 for h = 1)
  uint32_t brow, bch0;
  if constexpr (EXPAND) {
    brow = XBYTES + (uint32_t)(wn / 16) * (16 * kRowBytes);       // + q * 2048 + k * 128
    bch0 = 0;
  } else {
    brow = XBYTES + (uint32_t)(wn + pg) * kRowBytes;              // + j * 8 rows
    bch0 = xch0;
  }

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  int stage = 0;
  uint32_t phase = 0;
  for (int kt = 0; kt < KT; ++kt) {
    if constexpr (MODE != 2)  // MODE 2: diagnostic, MMA loop on whatever shared memory holds
      mbar_wait(&full_bar[stage], phase);
    const uint32_t sb = base + (uint32_t)stage * STAGE;
#pragma unroll
    for (int h = 0; h < (MODE == 4 ? 0 : 2); ++h) {
      double2 a[MI];
#pragma unroll
      for (int i = 0; i < MI; ++i)
        a[i] = lds128(sb + xrow + i * (8 * kRowBytes) + (xch0 ^ (h << 6)));
      if constexpr (!EXPAND) {
        double2 b[NI];
#pragma unroll
        for (int j = 0; j < NI; ++j)
          b[j] = lds128(sb + brow + j * (8 * kRowBytes) + (bch0 ^ (h << 6)));
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 8 * h + 2 * t + e;  // k & 7 == 2t + e
          double2 b[2];
#pragma unroll
          for (int q = 0; q < 2; ++q)
            b[q] = lds128(sb + brow + q * (16 * kRowBytes) + k * kRowBytes +
                          ((g ^ (2 * t + e)) << 4));
#pragma unroll
          for (int i = 0; i < MI; ++i) {
            const double av = e ? a[i].y : a[i].x;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              dmma884(acc[i][2 * q][0], acc[i][2 * q][1], av, b[q].x);
              dmma884(acc[i][2 * q + 1][0], acc[i][2 * q + 1][1], av, b[q].y);
            }
          }
        }
      }
    }
    if constexpr (MODE != 2) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
    }
    if (++stage == kTmaStages) {
      stage = 0;
      phase ^= 1u;
    }
  }

  // ------------------------------------------------------------ epilogue
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = m0 + wm + i * 8 + pg;
    if (row >= Mg) continue;
    double* crow = C + (int64_t)row * ldc;
    if constexpr (!EXPAND) {
      // C-fragment columns 2t, 2t+1 are operator rows perm(2t) = t, perm(2t+1) = t + 4
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int col = n0 + wn + j * 8 + t;
        if (col < Ng) crow[col] = acc[i][j][0];
        if (col + 4 < Ng) crow[col + 4] = acc[i][j][1];
      }
    } else {
      // blocks (2q, 2q+1) interleave: this lane holds columns 16q + 4t .. 4t + 3
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int col = n0 + wn + 16 * q + 4 * t;
        if (col < Ng) crow[col] = acc[i][2 * q][0];
        if (col + 1 < Ng) crow[col + 1] = acc[i][2 * q + 1][0];
        if (col + 2 < Ng) crow[col + 2] = acc[i][2 * q][1];
        if (col + 3 < Ng) crow[col + 3] = acc[i][2 * q + 1][1];
      }
    }
  }
}

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
    cudaGetLastError();
  }
  return fn;
}

// 2-D FP64 tensor [rows][cols] with leading dimension ld, box [box_rows][16], 128-byte swizzle
bool make_map(CUtensorMap* map, const double* ptr, int64_t rows, int64_t cols, int64_t ld,
              int box_rows) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), dims, strides, box,
             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_gemm_variant = 0;  // 0 = TMA kernel when possible, 1 = always the cp.async kernel; 2, 4 = probes

// returns 1 if launched, 0 if the TMA path does not apply, < 0 on error
template <int BM, bool EXPAND, int MODE = 0>
int launch_gemm_tma(const double* X, int64_t ldx, const double* A, int64_t lda, int a_rows, int a_cols,
                    double* C, int64_t ldc, int Mg, int Ng, int K, cudaStream_t st) {
  alignas(64) CUtensorMap mapX, mapA;
  if (!make_map(&mapX, X, Mg, K, ldx, BM)) return 0;
  if (!make_map(&mapA, A, a_rows, a_cols, lda, EXPAND ? 16 : kBN)) return 0;
  auto kern = k_dgemm_dmma_tma<BM, EXPAND, MODE>;
  constexpr size_t smem = (size_t)kTmaStages * (BM + kBN) * kRowBytes + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
      return trb_set_error(TRB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int m_tiles = (Mg + BM - 1) / BM, n_tiles = (Ng + kBN - 1) / kBN;
  kern<<<m_tiles * n_tiles, kTmaThreads, smem, st>>>(mapX, mapA, C, ldc, Mg, Ng, K, m_tiles);
  return 1;
}

template <int BM, bool EXPAND>
constexpr size_t gemm_smem() {
  return sizeof(double) * kStages * (size_t)(BM * kLdK + (EXPAND ? kBK * kLdN : kBN * kLdK));
}

template <int BM, bool EXPAND, int XVEC>
int launch_gemm(const double* X, int64_t ldx, const double* A, int64_t lda, double* C, int64_t ldc,
                int Mg, int Ng, int K, cudaStream_t st) {
  auto kern = k_dgemm_dmma<BM, EXPAND, XVEC>;
  constexpr size_t smem = gemm_smem<BM, EXPAND>();
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
      return trb_set_error(TRB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int m_tiles = (Mg + BM - 1) / BM, n_tiles = (Ng + kBN - 1) / kBN;
  kern<<<m_tiles * n_tiles, kThreads, smem, st>>>(X, ldx, A, lda, C, ldc, Mg, Ng, K, m_tiles);
  return TRB_OK;
}

// a_rows x a_cols: extent of the operator A[R, n] (rows of singular vectors)
template <bool EXPAND>
int dispatch_gemm(const double* X, int64_t ldx, const double* A, int64_t lda, int a_rows, int a_cols,
                  double* C, int64_t ldc, int Mg, int Ng, int K, cudaStream_t st) {
  const bool x16 = (ldx % 2 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  const bool small = Mg <= 64;
  if ((g_gemm_variant == 2 || g_gemm_variant == 4) && x16 && !small) {
    // measurement probes (results are NOT the product): 2 = MMA loop without loads
    // (ceiling of the consumer side), 4 = TMA loads without MMA (ceiling of the feed)
    const int rc = g_gemm_variant == 2
        ? launch_gemm_tma<128, EXPAND, 2>(X, ldx, A, lda, a_rows, a_cols, C, ldc, Mg, Ng, K, st)
        : launch_gemm_tma<128, EXPAND, 4>(X, ldx, A, lda, a_rows, a_cols, C, ldc, Mg, Ng, K, st);
    if (rc != 0) return rc < 0 ? rc : TRB_OK;
  }
  if (g_gemm_variant == 0 && x16) {
    const int rc = small ? launch_gemm_tma<64, EXPAND>(X, ldx, A, lda, a_rows, a_cols, C, ldc, Mg, Ng, K, st)
                         : launch_gemm_tma<128, EXPAND>(X, ldx, A, lda, a_rows, a_cols, C, ldc, Mg, Ng, K, st);
    if (rc != 0) return rc < 0 ? rc : TRB_OK;
  }
  if (small) {
    return x16 ? launch_gemm<64, EXPAND, 2>(X, ldx, A, lda, C, ldc, Mg, Ng, K, st)
               : launch_gemm<64, EXPAND, 1>(X, ldx, A, lda, C, ldc, Mg, Ng, K, st);
  }
  return x16 ? launch_gemm<128, EXPAND, 2>(X, ldx, A, lda, C, ldc, Mg, Ng, K, st)
             : launch_gemm<128, EXPAND, 1>(X, ldx, A, lda, C, ldc, Mg, Ng, K, st);
}

int check_gemm_args(const double* A, int R, int n, int ld, int B, const void* x, const void* y) {
  TRB_CHECK_ARG(A && x && y, "null pointer");
  TRB_CHECK_ARG(B > 0 && R > 0 && n > 0 && ld >= n, "bad shape");
  TRB_CHECK_ARG(ld % 2 == 0, "ld must be even (16-byte rows)");
  TRB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0, "operator must be 16-byte aligned");
  return TRB_OK;
}

}  // namespace

extern "C" void trb_gemm_set_variant(int variant) { g_gemm_variant = variant; }

extern "C" int trb_lin_project_gemm(const double* A, int R, int n, int ld, int B, const double* vec,
                                    int ldvec, double* t, void* stream) {
  int rc = check_gemm_args(A, R, n, ld, B, vec, t);
  if (rc) return rc;
  TRB_CHECK_ARG(ldvec >= n, "ldvec < n");
  cudaStream_t st = (cudaStream_t)stream;
  trb_launch_scope scope_(1, st);
  rc = dispatch_gemm<false>(vec, ldvec, A, ld, R, n, t, R, B, R, n, st);
  if (rc) return rc;
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

extern "C" int trb_lin_expand_gemm(const double* A, int R, int n, int ld, int B, const double* coef,
                                   double* out, int ldout, void* stream) {
  int rc = check_gemm_args(A, R, n, ld, B, coef, out);
  if (rc) return rc;
  TRB_CHECK_ARG(ldout >= n, "ldout < n");
  cudaStream_t st = (cudaStream_t)stream;
  trb_launch_scope scope_(1, st);
  rc = dispatch_gemm<true>(coef, R, A, ld, R, n, out, ldout, B, n, R, st);
  if (rc) return rc;
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}
