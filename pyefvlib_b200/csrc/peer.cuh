// peer.cuh -- layout of the per-rank peer-memory region (CUDA IPC over NVLink / NVSwitch) and the device-side
// pieces of the all-reduce that are FUSED into the Krylov kernels:
//   producer tail  (ar_publish): the last block of the kernel that produced the per-block partials folds them in fixed
//                  order and stores the result straight into every rank's slot, then raises the flags;
//   consumer head  (ar_wait / ar_value): every block of the next kernel waits for the flags of all ranks and sums the
//                  slots in rank order (bit-identical scalars on every rank).
// Region of every rank:
//   double slot[2][kMaxRanks][4]   all-reduce operands, double-buffered by the parity of the sequence number
//   uint32 flag[2][kMaxRanks]      flag[parity][r] = sequence number once rank r's operands have landed
//   int    error                   set when a spin wait times out
// Double buffering is safe: a rank publishes sequence k+2 only after consuming k+1, which every peer publishes from a
// kernel that is stream-ordered after its own consumer of k.
#pragma once
#include "common.cuh"

namespace efv {

constexpr int kMaxRanks = 16;
constexpr size_t kSlotBytes = 2 * kMaxRanks * 4 * sizeof(double);
constexpr size_t kFlagBytes = 2 * kMaxRanks * sizeof(uint32_t);
// small host-driven messages between ranks (set-up of per-level halos, counts): one 256-byte box per (parity, sender)
constexpr size_t kMailBytes = 256;
constexpr size_t kMailOffset = kSlotBytes + kFlagBytes + 64;
constexpr size_t kMailFlagOffset = kMailOffset + 2 * kMaxRanks * kMailBytes;
constexpr size_t kPeerBytes = kMailFlagOffset + 2 * kMaxRanks * sizeof(uint32_t) + 64;
// A rank that never arrives turns into an error, not a hang: spin waits give up after kPeerTimeoutNs of WALL time
// (%globaltimer), generous enough for ordinary rank skew (a peer writing output files, lazy module loads, uneven
// assembly, two ranks time-sliced on one GPU).  After a timeout the error word stays set until the host has reported it
// (comm_peer_error() clears it), and every later wait returns at once, so the solve ends instead of timing out again.
constexpr unsigned long long kPeerTimeoutNs = 60ULL * 1000000000ULL;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait until *f >= seq (wrap-safe); returns false on timeout / earlier error (err set to `code`)
__device__ __forceinline__ bool peer_wait_flag(volatile const uint32_t *f, uint32_t seq, int *err, int code) {
  if ((int32_t)(*f - seq) >= 0) return true;
  if (*reinterpret_cast<volatile int *>(err) != 0) return false;
  const unsigned long long t0 = global_ns();
  unsigned spins = 0;
  while ((int32_t)(*f - seq) < 0) {
    __nanosleep(100);
    if ((++spins & 1023u) == 0u) {
      if (*reinterpret_cast<volatile int *>(err) != 0) return false;
      if (global_ns() - t0 > kPeerTimeoutNs) { atomicExch(err, code); return false; }
    }
  }
  return true;
}

struct PeerPtrs { char *p[kMaxRanks]; };

__device__ __forceinline__ double *peer_slot(char *base, int parity, int r) {
  return reinterpret_cast<double *>(base) + ((size_t)parity * kMaxRanks + r) * 4;
}
__device__ __forceinline__ volatile uint32_t *peer_flag(char *base, int parity, int r) {
  return reinterpret_cast<volatile uint32_t *>(base + kSlotBytes) + (size_t)parity * kMaxRanks + r;
}
__device__ __forceinline__ int *peer_error(char *base) { return reinterpret_cast<int *>(base + kSlotBytes + kFlagBytes); }

// one fused all-reduce: which regions, who I am, the block counter of the producing kernel, the sequence number
struct PeerAR {
  PeerPtrs peers;
  int me = 0, nranks = 1;
  unsigned *counter = nullptr;
  uint32_t seq = 0;
  int enabled = 0;
};

// Called by EVERY thread of EVERY block of the producing kernel after the block stored its partial(s).
// partial: narr arrays of G per-block partials (entries of blocks that do not exist must already be zero).
__device__ __forceinline__ void ar_publish(const PeerAR &ar, const double *partial, int G, int narr, double *red) {
  __shared__ int s_last;
  __shared__ double s_mine[4];
  __threadfence();                                   // this block's partials before its ticket
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ar.counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int a = 0; a < narr; ++a) {
    double v = 0.0;
    for (int i = threadIdx.x; i < G; i += blockDim.x) v += __ldcg(partial + (size_t)a * G + i);
    v = block_sum(v, red);
    if (threadIdx.x == 0) s_mine[a] = v;
    __syncthreads();
  }
  const int parity = ar.seq & 1;
  if ((int)threadIdx.x < ar.nranks) {
    char *dst = ar.peers.p[threadIdx.x];
    double *slot = peer_slot(dst, parity, ar.me);
    for (int a = 0; a < narr; ++a) slot[a] = s_mine[a];
    __threadfence_system();
    *peer_flag(dst, parity, ar.me) = ar.seq;
  }
  if (threadIdx.x == 0) *ar.counter = 0u;            // next producer starts from zero (stream-ordered)
}
// Called by every thread of every block of the consuming kernel before it needs the scalars.
__device__ __forceinline__ void ar_wait(const PeerAR &ar) {
  char *self = ar.peers.p[ar.me];
  if ((int)threadIdx.x < ar.nranks) {
    peer_wait_flag(peer_flag(self, ar.seq & 1, threadIdx.x), ar.seq, peer_error(self), 1);
  }
  __syncthreads();
  __threadfence_system();
}
__device__ __forceinline__ double ar_value(const PeerAR &ar, int a) {
  char *self = ar.peers.p[ar.me];
  double acc = 0.0;
  for (int r = 0; r < ar.nranks; ++r) acc += *reinterpret_cast<volatile double *>(peer_slot(self, ar.seq & 1, r) + a);
  return acc;
}

PeerAR comm_peer_ar(unsigned *counter);   // next sequence number, enabled (host; comm.cu)

}  // namespace efv
