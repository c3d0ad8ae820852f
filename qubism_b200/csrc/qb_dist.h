// Multi-GPU layer: one process per GPU, rank r owns the amplitudes whose top log2(P)
// physical index bits equal r.  NCCL is loaded lazily (dlopen) so that a
// single-GPU process never needs it.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "qb_internal.h"

namespace qb {

struct DistState;
struct DistGroup;  // in-process rank group: P contexts of one process, one host thread each (no NCCL)
DistGroup *dist_group_create(int nranks);
int dist_create_group(DistState **out, int device, int rank, DistGroup *grp);

const char *dist_last_error();
int dist_unique_id(void *id128);
int dist_create(DistState **out, int device, int rank, int nranks, const void *id128, cudaStream_t stream);
void dist_destroy(DistState *d);

// in-place sum over ranks of n host doubles (n <= 64); synchronises the stream
int dist_allreduce_sum(DistState *d, double *vals, int n, cudaStream_t stream);

// Collective: export `ptr` (a cudaMalloc'ed shard) over CUDA IPC and map every peer's shard.
// peers[r] is rank r's shard as seen from this process (peers[rank] == ptr); empty if IPC is
// unavailable, in which case swaps fall back to NCCL send/recv.
int dist_register(DistState *d, double2 *ptr, std::vector<double2 *> &peers, cudaStream_t stream);
// Collective: unmap the peers' shards (call before freeing the exported buffer).
int dist_unregister(DistState *d, std::vector<double2 *> &peers, cudaStream_t stream);

// Exchange data so that the global targets of `pending` become local; updates perm.
// With peer mappings: one in-place swap kernel per partner over NVLink peer memory, evicting
// the local bits that are needed furthest in the future.  Without: NCCL send/recv of the top
// local bits through bounce buffers.
int dist_make_local(DistState *d, double2 *amps, const std::vector<double2 *> &peers, int n, int L,
                    std::vector<int> &perm, const std::vector<const HostOp *> &pending, int sm_count,
                    cudaStream_t stream, qb_stats *stats, const std::vector<const HostOp *> *future = nullptr);

// The same exchange for an explicit list of (global bit, local bit) pairs (layout changes that no
// gate asked for: operands of <.> / +: / tensor whose layouts diverged).
int dist_swap_pairs(DistState *d, double2 *amps, const std::vector<double2 *> &peers, int L, std::vector<int> &perm,
                    const std::vector<SwapPair> &sw, int sm_count, cudaStream_t stream, qb_stats *stats);
bool dist_has_peers(const DistState *d, const std::vector<double2 *> &peers);
// stream-ordered barrier over all ranks: every rank's earlier work on its stream is complete before
// anything enqueued after it starts on any rank
int dist_stream_barrier(DistState *d, cudaStream_t stream);

// Collective read of logical amplitudes [first, first + count) into `out` on every rank.
int dist_read_logical(DistState *d, const double2 *amps, int n, int L, const std::vector<int> &perm, uint64_t first,
                      uint64_t count, qb_c64 *out, cudaStream_t stream);

}  // namespace qb
