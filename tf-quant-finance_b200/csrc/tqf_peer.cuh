// All-reduce of a few hundred doubles over the GPUs of one box through peer
// memory (NVLink), callable from the tail of any kernel: used by the fused
// Longstaff-Schwartz pass (per-date normal equations) and by the payoff
// reduction of the fused path kernel (per-call payoff sums).  Replaces the
// ncclAllReduce of SURVEY 8e on the date-to-date / call-to-call critical path.
#pragma once

#include "tqf_common.cuh"

namespace tqf {

// Exchange buffer of one rank: flags uint64 [2][kLsmMaxPeers] (a 128-byte
// line per parity), then sums double [2][kLsmMaxPeers][kLsmPeerMaxSums].
constexpr int kLsmMaxPeers = 8;
constexpr int kLsmPeerMaxBatch = 16;
constexpr int kLsmPeerMaxSums = kLsmPeerMaxBatch * 27;
constexpr size_t kLsmPeerFlagBytes = 2 * 128;
constexpr size_t kLsmPeerBytes =
    kLsmPeerFlagBytes + 2ull * kLsmMaxPeers * kLsmPeerMaxSums * sizeof(double);


// What the tail needs of the peer exchange, passed BY VALUE (taking the address
// of the kernel parameter struct would move all of it to local memory, hot loop
// included: measured +14 us per pass).
struct PeerK {
  int peer_rank, peer_world;
  unsigned long long peer_epoch;
  unsigned char* peer_bufs[kLsmMaxPeers];
};

__device__ __forceinline__ unsigned long long* peer_flag(unsigned char* buf, int parity, int src) {
  return reinterpret_cast<unsigned long long*>(buf + parity * 128) + src;
}
// Two status words behind the eight flags of parity 0 (the flag line is 128 bytes).
__device__ __forceinline__ unsigned long long* peer_status_words(unsigned char* buf) {
  return reinterpret_cast<unsigned long long*>(buf) + kLsmMaxPeers;
}
__device__ __forceinline__ double* peer_sums(unsigned char* buf, int parity, int src) {
  return reinterpret_cast<double*>(buf + kLsmPeerFlagBytes) +
         (static_cast<size_t>(parity) * kLsmMaxPeers + src) * kLsmPeerMaxSums;
}

// All-reduce of the M local sums over the ranks of one box, executed by the
// tail CTA of every rank: each rank stores its sums into slot [parity][rank] of
// EVERY rank's buffer (peer stores over NVLink), fences, raises its flag there
// (st.release.sys) and waits for the flags of all ranks in its own buffer
// (ld.acquire.sys); then every rank adds the slots in rank order -- the same
// order everywhere, so all ranks solve from bit-identical sums.  Slots are
// double-buffered by the parity of the epoch: a rank cannot run two exchanges
// ahead of another one, because each exchange needs every rank's flag.
// Returns false when a peer did not arrive within ~10 s: the sums are then
// poisoned with NaN instead of hanging the GPU, and the time-out is counted in
// the status words of the rank's own buffer, which the host reads with
// tqf_peer_status to tell this cause of a NaN from any other.
// `nthreads` / `bar_id`: the threads of the CTA that call this together.  bar_id
// < 0: the whole CTA (__syncthreads); otherwise the first `nthreads` threads,
// synchronised on named barrier `bar_id` (warp-specialised kernels whose other
// warps do not take part).
__device__ __forceinline__ void peer_sync(int nthreads, int bar_id) {
  if (bar_id < 0)
    __syncthreads();
  else
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ bool peer_all_reduce(const PeerK& A, double* sums, int M,
                                                int nthreads = 0, int bar_id = -1) {
  __shared__ int s_timeout;
  if (bar_id < 0) nthreads = blockDim.x;
  const int parity = static_cast<int>(A.peer_epoch & 1ull);
  if (threadIdx.x == 0) s_timeout = 0;
  for (int i = threadIdx.x; i < A.peer_world * M; i += nthreads) {
    const int r = i / M, m = i - r * M;
    peer_sums(A.peer_bufs[r], parity, A.peer_rank)[m] = sums[m];
  }
  __threadfence_system();
  peer_sync(nthreads, bar_id);
  if (threadIdx.x < A.peer_world) {
    unsigned long long* remote = peer_flag(A.peer_bufs[threadIdx.x], parity, A.peer_rank);
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(remote), "l"(A.peer_epoch) : "memory");
    const unsigned long long* mine = peer_flag(A.peer_bufs[A.peer_rank], parity, threadIdx.x);
    const long long t0 = clock64();
    unsigned long long seen = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
      if (seen >= A.peer_epoch) break;
      if (clock64() - t0 > 20000000000ll) {
        s_timeout = 1;
        // status words of this rank's own buffer (tqf_peer_status): number of
        // exchanges that timed out, and the epoch of the last one
        unsigned long long* status = peer_status_words(A.peer_bufs[A.peer_rank]);
        atomicAdd(status, 1ull);
        atomicExch(status + 1, A.peer_epoch);
        break;
      }
    }
  }
  peer_sync(nthreads, bar_id);
  const bool ok = s_timeout == 0;
  for (int m = threadIdx.x; m < M; m += nthreads) {
    double v = 0.0;
    for (int r = 0; r < A.peer_world; ++r)
      v += *reinterpret_cast<volatile double*>(peer_sums(A.peer_bufs[A.peer_rank], parity, r) + m);
    sums[m] = ok ? v : __longlong_as_double(0x7ff8000000000000ll);
  }
  peer_sync(nthreads, bar_id);
  return ok;
}


// Host-side copy of the exchange set-up (tqf_*_set_peer_exchange).
struct PeerHost {
  int rank, world;
  unsigned long long epoch;
  unsigned char* bufs[kLsmMaxPeers];
};

}  // namespace tqf
