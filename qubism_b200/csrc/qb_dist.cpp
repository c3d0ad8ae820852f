// Multi-GPU layer, first cut: NCCL send/recv global<->local qubit swaps.
//
// A non-diagonal gate whose target is a GLOBAL physical bit (one of the rank bits) cannot run
// inside a shard.  Instead of exchanging per gate, the k global bits that pending gates need
// are swapped with the TOP k local bits in one step: every rank keeps the 2^-k of its shard
// whose top-k local bits already equal its own value of those rank bits and trades each other
// contiguous 2^(L-k) block with exactly one peer (all peers are equidistant through NVSwitch).
// The logical->physical bit map is updated, so every later gate on those qubits is local.
// Volume per rank: (1 - 2^-k) of the shard each way -- one half-shard for k = 1.
#include "qb_dist.h"
#include "qb_kernels.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

namespace qb {

static thread_local std::string g_dist_err;
const char *dist_last_error() { return g_dist_err.c_str(); }

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    g_dist_err = std::string("cannot load libnccl.so.2: ") + dlerror();
    return nullptr;
  }
#define QB_SYM(field, name)                                             \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name)); \
  if (!api.field) {                                                     \
    g_dist_err = std::string("missing NCCL symbol ") + name;            \
    api.handle = nullptr;                                               \
    return nullptr;                                                     \
  }
  QB_SYM(GetUniqueId, "ncclGetUniqueId")
  QB_SYM(CommInitRank, "ncclCommInitRank")
  QB_SYM(CommDestroy, "ncclCommDestroy")
  QB_SYM(AllReduce, "ncclAllReduce")
  QB_SYM(AllGather, "ncclAllGather")
  QB_SYM(Send, "ncclSend")
  QB_SYM(Recv, "ncclRecv")
  QB_SYM(GroupStart, "ncclGroupStart")
  QB_SYM(GroupEnd, "ncclGroupEnd")
  QB_SYM(GetErrorString, "ncclGetErrorString")
#undef QB_SYM
  return &api;
}

}  // namespace

// In-process rank group (qb_init_group): the P ranks are P contexts of ONE process, each driven by
// its own host thread -- P GPUs without NCCL, or P "virtual ranks" on one GPU (tests: the sharded
// planner, the fused passes with rank bits and the peer swap kernel all run for P = 2 / 4 / 8 on a
// 1-GPU box).  Collectives are host-side: a generation barrier plus shared slots.  Shards are
// plain device pointers of the same process (peer access enabled between distinct devices).
struct DistGroup {
  int nranks = 1;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0, refs = 0;
  uint64_t generation = 0;
  bool broken = false;                 // a rank timed out: every later collective fails at once
  std::vector<double> slots;           // nranks x 64 doubles
  std::vector<void *> ptrs;            // nranks
  std::vector<qb_c64> gather;          // dist_read_logical staging
  int timeout_s = 300;
};

DistGroup *dist_group_create(int nranks) {
  DistGroup *g = new DistGroup();
  g->nranks = nranks;
  g->slots.assign(size_t(nranks) * 64, 0.0);
  g->ptrs.assign(nranks, nullptr);
  if (const char *e = getenv("QB_GROUP_TIMEOUT"))
    if (atoi(e) > 0) g->timeout_s = atoi(e);
  return g;
}

// generation barrier over the group's host threads; false = a rank never arrived
static bool group_barrier(DistGroup *g) {
  std::unique_lock<std::mutex> lk(g->mu);
  if (g->broken) return false;
  const uint64_t gen = g->generation;
  if (++g->arrived == g->nranks) {
    g->arrived = 0;
    ++g->generation;
    g->cv.notify_all();
    return true;
  }
  const bool ok = g->cv.wait_for(lk, std::chrono::seconds(g->timeout_s), [&] { return g->generation != gen || g->broken; });
  if (!ok || g->broken) {
    g->broken = true;
    g->cv.notify_all();
    return false;
  }
  return true;
}

struct DistState {
  int device = 0, rank = 0, nranks = 1, pbits = 0;
  DistGroup *grp = nullptr;          // in-process group (no NCCL) when set
  ncclComm_t comm = nullptr;
  double *scratch_dev = nullptr;   // 64 doubles
  unsigned char *ipc_dev = nullptr;  // nranks IPC handles
  bool ipc_ok = true;                // cleared for good after the first failed import
  double2 *bounce[2] = {nullptr, nullptr};
  size_t bounce_amps = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_recv[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
};

#define QB_NCCL(expr)                                                              \
  do {                                                                             \
    ncclResult_t r__ = (expr);                                                     \
    if (r__ != ncclSuccess) {                                                      \
      g_dist_err = std::string(#expr) + ": " + nccl()->GetErrorString(r__);        \
      return QB_ERR_NCCL;                                                          \
    }                                                                              \
  } while (0)

#define QB_DCUDA(expr)                                                             \
  do {                                                                             \
    cudaError_t e__ = (expr);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      g_dist_err = std::string(#expr) + ": " + cudaGetErrorString(e__);            \
      return e__ == cudaErrorMemoryAllocation ? QB_ERR_OOM : QB_ERR_CUDA;          \
    }                                                                              \
  } while (0)

int dist_unique_id(void *id128) {
  NcclApi *a = nccl();
  if (!a) return QB_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  QB_NCCL(a->GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return QB_OK;
}

int dist_create(DistState **out, int device, int rank, int nranks, const void *id128, cudaStream_t stream) {
  NcclApi *a = nccl();
  if (!a) return QB_ERR_NCCL;
  DistState *d = new DistState();
  d->device = device;
  d->rank = rank;
  d->nranks = nranks;
  d->pbits = __builtin_ctz((unsigned)nranks);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  QB_DCUDA(cudaSetDevice(device));
  QB_NCCL(a->CommInitRank(&d->comm, nranks, id, rank));
  QB_DCUDA(cudaMalloc(&d->scratch_dev, 64 * sizeof(double)));
  QB_DCUDA(cudaMalloc(&d->ipc_dev, size_t(nranks) * sizeof(cudaIpcMemHandle_t)));
  if (const char *e = getenv("QB_PEER_EXCHANGE"))
    if (*e == '0') d->ipc_ok = false;
  QB_DCUDA(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    QB_DCUDA(cudaEventCreateWithFlags(&d->ev_recv[i], cudaEventDisableTiming));
    QB_DCUDA(cudaEventCreateWithFlags(&d->ev_copy[i], cudaEventDisableTiming));
  }
  (void)stream;
  *out = d;
  return QB_OK;
}

int dist_create_group(DistState **out, int device, int rank, DistGroup *grp) {
  DistState *d = new DistState();
  d->device = device;
  d->rank = rank;
  d->nranks = grp->nranks;
  d->pbits = __builtin_ctz((unsigned)grp->nranks);
  d->grp = grp;
  QB_DCUDA(cudaSetDevice(device));
  QB_DCUDA(cudaMalloc(&d->scratch_dev, 64 * sizeof(double)));
  {
    std::lock_guard<std::mutex> lk(grp->mu);
    ++grp->refs;
  }
  *out = d;
  return QB_OK;
}

void dist_destroy(DistState *d) {
  if (!d) return;
  cudaSetDevice(d->device);
  if (d->grp) {
    bool last;
    {
      std::lock_guard<std::mutex> lk(d->grp->mu);
      last = --d->grp->refs == 0;
    }
    if (last) delete d->grp;
    cudaFree(d->scratch_dev);
    delete d;
    return;
  }
  if (d->comm && nccl()) nccl()->CommDestroy(d->comm);
  cudaFree(d->scratch_dev);
  cudaFree(d->ipc_dev);
  for (int i = 0; i < 2; ++i) {
    cudaFree(d->bounce[i]);
    if (d->ev_recv[i]) cudaEventDestroy(d->ev_recv[i]);
    if (d->ev_copy[i]) cudaEventDestroy(d->ev_copy[i]);
  }
  if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
  delete d;
}

#define QB_GROUP_BARRIER(d)                                                        \
  do {                                                                             \
    if (!group_barrier((d)->grp)) {                                                \
      g_dist_err = "rank group barrier timed out (a rank failed or never arrived)"; \
      return QB_ERR_NCCL;                                                          \
    }                                                                              \
  } while (0)

int dist_allreduce_sum(DistState *d, double *vals, int n, cudaStream_t stream) {
  if (n > 64) return QB_ERR_ARG;
  if (d->grp) {  // host-side: every rank's earlier stream work is complete, sums in rank order
    QB_DCUDA(cudaStreamSynchronize(stream));
    DistGroup *g = d->grp;
    memcpy(&g->slots[size_t(d->rank) * 64], vals, n * sizeof(double));
    QB_GROUP_BARRIER(d);
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int r = 0; r < d->nranks; ++r) s += g->slots[size_t(r) * 64 + i];
      vals[i] = s;
    }
    QB_GROUP_BARRIER(d);  // nobody overwrites a slot another rank is still reading
    return QB_OK;
  }
  NcclApi *a = nccl();
  QB_DCUDA(cudaMemcpyAsync(d->scratch_dev, vals, n * sizeof(double), cudaMemcpyHostToDevice, stream));
  QB_NCCL(a->AllReduce(d->scratch_dev, d->scratch_dev, n, ncclDouble, ncclSum, d->comm, stream));
  QB_DCUDA(cudaMemcpyAsync(vals, d->scratch_dev, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
  QB_DCUDA(cudaStreamSynchronize(stream));
  return QB_OK;
}

// stream-ordered barrier over all ranks: every rank's earlier work on its stream is complete
// before anything enqueued after it starts on any rank
int dist_stream_barrier(DistState *d, cudaStream_t stream);
static int stream_barrier(DistState *d, cudaStream_t stream) { return dist_stream_barrier(d, stream); }
int dist_stream_barrier(DistState *d, cudaStream_t stream) {
  if (d->grp) {
    QB_DCUDA(cudaStreamSynchronize(stream));
    QB_GROUP_BARRIER(d);
    return QB_OK;
  }
  NcclApi *a = nccl();
  QB_NCCL(a->AllReduce(d->scratch_dev + 32, d->scratch_dev + 32, 1, ncclDouble, ncclSum, d->comm, stream));
  return QB_OK;
}

int dist_register(DistState *d, double2 *ptr, std::vector<double2 *> &peers, cudaStream_t stream) {
  peers.clear();
  if (d->grp) {  // same process: the peers' shards are plain device pointers
    DistGroup *g = d->grp;
    g->ptrs[d->rank] = ptr;
    QB_GROUP_BARRIER(d);
    if (d->ipc_ok) {
      peers.resize(d->nranks);
      for (int r = 0; r < d->nranks; ++r) peers[r] = static_cast<double2 *>(g->ptrs[r]);
    }
    QB_GROUP_BARRIER(d);
    (void)stream;
    return QB_OK;
  }
  NcclApi *a = nccl();
  // every rank takes part in the handle exchange even if IPC is switched off, so that the
  // collective call sequence stays identical on all ranks
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  bool ok = d->ipc_ok;
  if (ok && cudaIpcGetMemHandle(&mine, ptr) != cudaSuccess) {
    cudaGetLastError();
    ok = false;
    memset(&mine, 0, sizeof mine);
  }
  std::vector<cudaIpcMemHandle_t> all(d->nranks);
  QB_DCUDA(cudaMemcpyAsync(d->ipc_dev + size_t(d->rank) * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice,
                           stream));
  QB_NCCL(a->AllGather(d->ipc_dev + size_t(d->rank) * sizeof mine, d->ipc_dev, sizeof mine, ncclChar, d->comm, stream));
  QB_DCUDA(cudaMemcpyAsync(all.data(), d->ipc_dev, all.size() * sizeof mine, cudaMemcpyDeviceToHost, stream));
  QB_DCUDA(cudaStreamSynchronize(stream));
  cudaIpcMemHandle_t zero;
  memset(&zero, 0, sizeof zero);
  for (int r = 0; r < d->nranks; ++r)
    if (memcmp(&all[r], &zero, sizeof zero) == 0) ok = false;  // some rank could not export
  std::vector<double2 *> out(d->nranks, nullptr);
  if (ok) {
    for (int r = 0; r < d->nranks && ok; ++r) {
      if (r == d->rank) {
        out[r] = ptr;
        continue;
      }
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
      } else {
        out[r] = static_cast<double2 *>(p);
      }
    }
  }
  // agree on the outcome: one failed import anywhere switches every rank to the NCCL path
  double flag = ok ? 0.0 : 1.0;
  int rc = dist_allreduce_sum(d, &flag, 1, stream);
  if (rc != QB_OK) return rc;
  if (flag != 0.0) {
    for (int r = 0; r < d->nranks; ++r)
      if (r != d->rank && out[r]) cudaIpcCloseMemHandle(out[r]);
    d->ipc_ok = false;
    return QB_OK;
  }
  peers.swap(out);
  return QB_OK;
}

int dist_unregister(DistState *d, std::vector<double2 *> &peers, cudaStream_t stream) {
  QB_DCUDA(cudaStreamSynchronize(stream));
  if (d->grp) {
    peers.clear();
    QB_GROUP_BARRIER(d);  // nobody frees a shard a peer's kernel may still touch
    return QB_OK;
  }
  for (int r = 0; r < (int)peers.size(); ++r)
    if (r != d->rank && peers[r]) cudaIpcCloseMemHandle(peers[r]);
  peers.clear();
  // nobody frees an exported shard while a peer may still have it mapped
  double z = 0.0;
  return dist_allreduce_sum(d, &z, 1, stream);
}

static int ensure_bounce(DistState *d, size_t amps) {
  if (d->bounce_amps >= amps) return QB_OK;
  for (int i = 0; i < 2; ++i) {
    if (d->bounce[i]) cudaFree(d->bounce[i]);
    d->bounce[i] = nullptr;
  }
  d->bounce_amps = 0;
  for (int i = 0; i < 2; ++i) QB_DCUDA(cudaMalloc(&d->bounce[i], amps * sizeof(double2)));
  d->bounce_amps = amps;
  return QB_OK;
}

int dist_make_local(DistState *d, double2 *amps, const std::vector<double2 *> &peers, int n, int L,
                    std::vector<int> &perm, const std::vector<const HostOp *> &pending, int sm_count,
                    cudaStream_t stream, qb_stats *stats, const std::vector<const HostOp *> *future) {
  const bool peer_path = (int)peers.size() == d->nranks;
  std::vector<SwapPair> sw = choose_swaps(n, L, perm, pending, peer_path, future);
  if (sw.empty()) {
    g_dist_err = "planner stuck but no global target pending";
    return QB_ERR_UNSUPPORTED;
  }
  return dist_swap_pairs(d, amps, peers, L, perm, sw, sm_count, stream, stats);
}

bool dist_has_peers(const DistState *d, const std::vector<double2 *> &peers) { return (int)peers.size() == d->nranks; }

// Execute one all-to-all swap step: the global physical bits sw[i].gbit trade places with the local
// bits sw[i].lbit (without peer mappings the local bits must be the TOP k local bits, descending).
int dist_swap_pairs(DistState *d, double2 *amps, const std::vector<double2 *> &peers, int L, std::vector<int> &perm,
                    const std::vector<SwapPair> &sw, int sm_count, cudaStream_t stream, qb_stats *stats) {
  NcclApi *a = d->grp ? nullptr : nccl();
  const bool peer_path = (int)peers.size() == d->nranks;
  const int k = (int)sw.size();
  if (k > L) {
    g_dist_err = "shard too small for the swap";
    return QB_ERR_UNSUPPORTED;
  }
  const uint64_t block = 1ull << (L - k);
  const std::vector<SwapStep> steps = swap_schedule(d->rank, L, sw);
  const uint64_t nsteps = steps.size();
  if (peer_path) {
    // NVLink-native path: barrier, one in-place swap kernel per partner (the lower rank of a pair
    // takes the first half of the index range, the higher rank the second), barrier.
    uint64_t swapped = 0;
    for (const SwapPair &sp : sw) swapped |= 1ull << sp.lbit;
    int rc = stream_barrier(d, stream);
    if (rc != QB_OK) return rc;
    for (const SwapStep &st : steps) {
      const uint64_t half = block / 2;
      const uint64_t tb = (d->rank < st.peer) ? 0 : half, te = (d->rank < st.peer) ? half : block;
      QB_DCUDA(launch_peer_swap(amps, peers[st.peer], tb, te, L, swapped, place_sel(st.my_sel, sw),
                                place_sel(st.peer_sel, sw), sm_count, stream));
      if (stats) stats->exchange_bytes += (te - tb) * sizeof(double2) * 2;  // read + written remotely
    }
    rc = stream_barrier(d, stream);
    if (rc != QB_OK) return rc;
    if (stats) stats->exchanges++;
    apply_swaps_to_perm(perm, sw);
    return QB_OK;
  }
  if (d->grp) {
    g_dist_err = "in-process rank groups exchange through peer pointers only (QB_PEER_EXCHANGE=0 needs NCCL)";
    return QB_ERR_UNSUPPORTED;
  }
  // NCCL path: every peer at once (one group = an all-to-all over the 2^k - 1 partners), in
  // pieces of <= 2^22 amplitudes (64 MiB) per peer, double-buffered: while the copy-back of
  // piece i drains on the copy stream, piece i+1 is already on the wire.
  const uint64_t lowest = (uint64_t)(L - k);  // the swapped bits are the top k local bits here
  const uint64_t piece = std::min<uint64_t>(block, 1ull << 22);
  int rc = ensure_bounce(d, piece * nsteps);
  if (rc != QB_OK) return rc;
  int slot = 0;
  bool used[2] = {false, false};
  for (uint64_t off = 0; off < block; off += piece) {
    if (used[slot]) QB_DCUDA(cudaStreamWaitEvent(stream, d->ev_copy[slot], 0));  // bounce free again
    QB_NCCL(a->GroupStart());
    for (uint64_t i = 0; i < nsteps; ++i) {
      double2 *src = amps + (place_sel(steps[i].my_sel, sw) >> lowest) * block + off;
      QB_NCCL(a->Send(src, piece * 2, ncclDouble, steps[i].peer, d->comm, stream));
      QB_NCCL(a->Recv(d->bounce[slot] + i * piece, piece * 2, ncclDouble, steps[i].peer, d->comm, stream));
    }
    QB_NCCL(a->GroupEnd());
    QB_DCUDA(cudaEventRecord(d->ev_recv[slot], stream));
    QB_DCUDA(cudaStreamWaitEvent(d->copy_stream, d->ev_recv[slot], 0));
    for (uint64_t i = 0; i < nsteps; ++i)
      QB_DCUDA(cudaMemcpyAsync(amps + (place_sel(steps[i].my_sel, sw) >> lowest) * block + off,
                               d->bounce[slot] + i * piece, piece * sizeof(double2), cudaMemcpyDeviceToDevice,
                               d->copy_stream));
    QB_DCUDA(cudaEventRecord(d->ev_copy[slot], d->copy_stream));
    used[slot] = true;
    slot ^= 1;
    if (stats) stats->exchange_bytes += nsteps * piece * sizeof(double2);
  }
  for (int s2 = 0; s2 < 2; ++s2)
    if (used[s2]) QB_DCUDA(cudaStreamWaitEvent(stream, d->ev_copy[s2], 0));
  if (stats) stats->exchanges++;
  apply_swaps_to_perm(perm, sw);
  return QB_OK;
}

// one piece (<= 2^22 amplitudes) of dist_read_logical
static int read_piece(DistState *d, const double2 *amps, int n, int L, const std::vector<int> &perm, uint64_t first,
                      uint64_t count, qb_c64 *out, cudaStream_t stream) {
  // Amplitudes this rank does not own start as -0.0: x + (-0.0) == x for EVERY x (including +0.0
  // and -0.0 themselves), so the sum over ranks returns the owner's bits unchanged.
  DistGroup *g = d->grp;
  std::vector<qb_c64> local;
  if (!g) local.assign(count, qb_c64{-0.0, -0.0});
  if (g && d->rank == 0) g->gather.assign(count, qb_c64{0.0, 0.0});
  if (g) QB_GROUP_BARRIER(d);
  qb_c64 *host = g ? g->gather.data() : local.data();  // group: ranks fill disjoint entries of one shared buffer
  QB_DCUDA(cudaStreamSynchronize(stream));
  // gather the elements this rank owns, merging contiguous runs into single copies
  uint64_t run_start = 0, run_len = 0, run_phys = 0;
  auto flush_run = [&]() -> int {
    if (run_len) QB_DCUDA(cudaMemcpyAsync(&host[run_start], amps + run_phys, run_len * sizeof(double2), cudaMemcpyDeviceToHost, stream));
    run_len = 0;
    return QB_OK;
  };
  for (uint64_t i = 0; i < count; ++i) {
    const uint64_t logical = first + i;
    uint64_t phys = 0;
    for (int q = 0; q < n; ++q)
      if ((logical >> q) & 1) phys |= 1ull << perm[q];
    const int owner = (int)(phys >> L);
    if (owner != d->rank) {
      int rc = flush_run();
      if (rc != QB_OK) return rc;
      continue;
    }
    const uint64_t local_idx = phys & ((1ull << L) - 1);
    if (run_len && local_idx == run_phys + run_len) {
      ++run_len;
    } else {
      int rc = flush_run();
      if (rc != QB_OK) return rc;
      run_start = i;
      run_phys = local_idx;
      run_len = 1;
    }
  }
  {
    int rc = flush_run();
    if (rc != QB_OK) return rc;
  }
  QB_DCUDA(cudaStreamSynchronize(stream));
  if (g) {
    QB_GROUP_BARRIER(d);
    memcpy(out, g->gather.data(), count * sizeof(qb_c64));
    QB_GROUP_BARRIER(d);
    return QB_OK;
  }
  NcclApi *a = nccl();
  double *tmp = nullptr;
  QB_DCUDA(cudaMalloc(&tmp, count * sizeof(double2)));
  cudaError_t e = cudaMemcpyAsync(tmp, host, count * sizeof(double2), cudaMemcpyHostToDevice, stream);
  ncclResult_t r = ncclSuccess;
  if (e == cudaSuccess) r = a->AllReduce(tmp, tmp, count * 2, ncclDouble, ncclSum, d->comm, stream);
  if (e == cudaSuccess && r == ncclSuccess)
    e = cudaMemcpyAsync(out, tmp, count * sizeof(double2), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(tmp);
  if (r != ncclSuccess) {
    g_dist_err = std::string("ncclAllReduce: ") + a->GetErrorString(r);
    return QB_ERR_NCCL;
  }
  if (e != cudaSuccess) {
    g_dist_err = std::string("distributed read: ") + cudaGetErrorString(e);
    return QB_ERR_CUDA;
  }
  return QB_OK;
}

int dist_read_logical(DistState *d, const double2 *amps, int n, int L, const std::vector<int> &perm, uint64_t first,
                      uint64_t count, qb_c64 *out, cudaStream_t stream) {
  // any range: pieces of 2^22 amplitudes (64 MiB of staging per piece)
  const uint64_t piece = 1ull << 22;
  for (uint64_t off = 0; off < count; off += piece) {
    const uint64_t c = std::min(piece, count - off);
    int rc = read_piece(d, amps, n, L, perm, first + off, c, out + off, stream);
    if (rc != QB_OK) return rc;
  }
  return QB_OK;
}

}  // namespace qb
