// Peer-memory exchange for a ROW-SHARDED LinearChannel (BASELINE
// config 5): one process per GPU, every rank owns a block of singular triplets.
// The two expansions of an iteration are partial sums over the ranks.  Instead
// of a library all-reduce between two kernels, each rank
//   1. reduces its per-CTA slots and, in the same grid-wide kernel, PUSHES the
//      reduced vector into slot `rank` of every peer's exchange buffer (posted
//      stores over NVLink / NVSwitch on pointers mapped with CUDA IPC; a push
//      has no round trip, unlike a pull),
//   2. publishes a sequence number into every peer's flag array
//      (st.release.sys), and
//   3. the CONSUMER kernel (k_z_update / k_x_update, trb_sweep.cu) waits for the
//      flags (ld.acquire.sys) and adds the ranks' vectors itself, in rank order,
//      from its own memory -- every rank gets bit-identical sums, and there is no
//      separate all-reduce kernel nor a reduced copy written back to HBM.
// Slots are double buffered by the parity of the exchange counter: a rank
// rewrites parity p two exchanges later, by which time every peer has published
// (hence finished reading) the exchange in between.
#include "trb_common.cuh"

using namespace trb;

struct trb_comm {
  int rank, nranks;
  size_t vec_doubles;        // capacity of one exchange vector
  unsigned char* base;       // this rank's allocation
  unsigned char* peer_base[TRB_MAX_RANKS];
  bool opened[TRB_MAX_RANKS];
  unsigned long long seq;    // exchanges published so far
  trb_peers last;            // handles of the latest exchange, for its consumer
};

namespace {

constexpr size_t kFlagBytes = 256;  // TRB_MAX_RANKS x u64, padded

// layout of one rank's buffer: flags | data[parity 0..1][source rank 0..TRB_MAX_RANKS-1][vec]
size_t comm_bytes(size_t vec_doubles) {
  return kFlagBytes + 2 * TRB_MAX_RANKS * vec_doubles * sizeof(double);
}
double* slot_ptr(unsigned char* base, size_t vec, unsigned long long parity, int src_rank) {
  return reinterpret_cast<double*>(base + kFlagBytes) + (parity * TRB_MAX_RANKS + src_rank) * vec;
}

__global__ void k_comm_signal(trb_peers peers, int rank) {
  const int p = threadIdx.x;
  if (p >= peers.n) return;
  __threadfence_system();
  unsigned long long* dst = peers.flags_of[p] + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(peers.seq) : "memory");
}

}  // namespace

extern "C" int trb_comm_create(int rank, int nranks, size_t vec_doubles, trb_comm** out,
                               unsigned char* handle64) {
  TRB_CHECK_ARG(out && handle64, "null pointer");
  TRB_CHECK_ARG(nranks >= 1 && nranks <= TRB_MAX_RANKS && rank >= 0 && rank < nranks, "bad rank");
  TRB_CHECK_ARG(vec_doubles > 0, "empty exchange vector");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  trb_comm* c = new trb_comm();
  c->rank = rank;
  c->nranks = nranks;
  c->vec_doubles = (vec_doubles + 15) / 16 * 16;
  c->seq = 0;
  memset(&c->last, 0, sizeof(c->last));
  for (int r = 0; r < TRB_MAX_RANKS; ++r) {
    c->peer_base[r] = nullptr;
    c->opened[r] = false;
  }
  cudaError_t e = cudaMalloc(&c->base, comm_bytes(c->vec_doubles));
  if (e == cudaSuccess) e = cudaMemset(c->base, 0, comm_bytes(c->vec_doubles));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->base);
  if (e != cudaSuccess) {
    delete c;
    return trb_set_error(TRB_ERR_CUDA, "trb_comm_create: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  c->peer_base[rank] = c->base;
  *out = c;
  return TRB_OK;
}

extern "C" int trb_comm_connect(trb_comm* c, const unsigned char* handles) {
  TRB_CHECK_ARG(c && handles, "null pointer");
  for (int r = 0; r < c->nranks; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * r, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return trb_set_error(TRB_ERR_CUDA, "trb_comm_connect(rank %d): %s", r, cudaGetErrorString(e));
    c->peer_base[r] = (unsigned char*)p;
    c->opened[r] = true;
  }
  return TRB_OK;
}

extern "C" int trb_comm_destroy(trb_comm* c) {
  if (!c) return TRB_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->nranks; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  cudaFree(c->base);
  delete c;
  return TRB_OK;
}

// where this rank's vector of the NEXT exchange goes: its slot in every rank's buffer
void trb_comm_push_targets(trb_comm* c, trb_push* push) {
  push->n = c->nranks;
  push->counter = nullptr;
  push->seq = 0;
  push->rank = c->rank;
  for (int r = 0; r < c->nranks; ++r) {
    push->dst[r] = slot_ptr(c->peer_base[r], c->vec_doubles, c->seq & 1, c->rank);
    push->flags_of[r] = reinterpret_cast<unsigned long long*>(c->peer_base[r]);
  }
}

// One exchange whose pushing kernel also publishes: fills `push` (targets, flags,
// sequence number, arrival counter) and records the consumer's handles (trb_comm_last).
void trb_comm_begin_exchange(trb_comm* c, trb_push* push) {
  trb_comm_push_targets(c, push);
  const unsigned long long parity = c->seq & 1;
  c->seq += 1;
  push->seq = c->seq;
  push->counter = reinterpret_cast<unsigned int*>(c->base + kFlagBytes - 16);  // inside the flag block
  trb_peers& p = c->last;
  p.n = c->nranks;
  p.seq = c->seq;
  for (int r = 0; r < c->nranks; ++r) {
    p.data[r] = slot_ptr(c->base, c->vec_doubles, parity, r);
    p.flags_of[r] = reinterpret_cast<unsigned long long*>(c->peer_base[r]);
  }
  p.my_flags = reinterpret_cast<const unsigned long long*>(c->base);
}

size_t trb_comm_capacity(const trb_comm* c) { return c->vec_doubles; }

// Publish the vector just written: bump the exchange counter, tell every rank
// (including this one).  Returns the handles the consumer kernel needs.
int trb_comm_publish(trb_comm* c, trb_peers* peers, cudaStream_t st) {
  const unsigned long long parity = c->seq & 1;
  c->seq += 1;
  peers->n = c->nranks;
  peers->seq = c->seq;
  for (int r = 0; r < c->nranks; ++r) {
    peers->data[r] = slot_ptr(c->base, c->vec_doubles, parity, r);  // local: rank r pushed it here
    peers->flags_of[r] = reinterpret_cast<unsigned long long*>(c->peer_base[r]);
  }
  peers->my_flags = reinterpret_cast<const unsigned long long*>(c->base);
  c->last = *peers;
  trb_launch_scope scope_(0, st);
  k_comm_signal<<<1, 32, 0, st>>>(*peers, c->rank);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}

const trb_peers* trb_comm_last(const trb_comm* c) { return &c->last; }

// Stand-alone all-reduce(sum) of vec[0 .. n) over the ranks through the same
// protocol (used by the tests and by callers outside the sweep).
namespace {
__global__ void __launch_bounds__(256)
k_comm_push(const double* __restrict__ src, trb_push push, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = src[i];
    for (int r = 0; r < push.n; ++r) push.dst[r][i] = v;
  }
}
__global__ void __launch_bounds__(256)
k_comm_sum_out(trb_peers peers, double* __restrict__ out, size_t n, int* timeout_flag) {
  if (!trb::peers_wait(peers) && timeout_flag && threadIdx.x == 0) atomicOr(timeout_flag, 1);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = trb::peers_sum(peers, i);
}
}  // namespace

extern "C" int trb_comm_all_reduce(trb_comm* c, double* vec, size_t n, int* timeout_flag, void* stream) {
  TRB_CHECK_ARG(c && vec, "null pointer");
  TRB_CHECK_ARG(n > 0 && n <= c->vec_doubles, "vector longer than the exchange buffer");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 4 * trb_sm_count_cached()) blocks = 4 * trb_sm_count_cached();
  {
    trb_push push;
    trb_comm_push_targets(c, &push);
    trb_launch_scope scope_(0, st);
    k_comm_push<<<blocks, 256, 0, st>>>(vec, push, n);
  }
  trb_peers peers;
  int rc = trb_comm_publish(c, &peers, st);
  if (rc) return rc;
  trb_launch_scope scope_(0, st);
  k_comm_sum_out<<<blocks, 256, 0, st>>>(peers, vec, n, timeout_flag);
  TRB_CHECK_LAUNCH();
  return TRB_OK;
}
