// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL communicator created from a unique id that the host
// layer distributes (torch.distributed), halo exchange of owned vertex values and all-reduce of Krylov scalars.
// NCCL is dlopen'ed (the path of the copy already loaded by torch is handed in) so that the process holds a single
// NCCL instance.  Everything runs on the library stream.
#include "views.cuh"
#include "peer.cuh"
#include <algorithm>
#include <dlfcn.h>
#include <cstring>

namespace efv {

constexpr int kMaxHaloNbr = kMaxRanks - 1;
constexpr size_t kHaloFlagBytes = 2 * kMaxRanks * sizeof(uint32_t) + 128;     // two parity sets of per-neighbour flags
struct HaloPush { double *dst[kMaxHaloNbr]; int64_t begin[kMaxHaloNbr + 1]; uint32_t *flag[kMaxHaloNbr]; int n; };

typedef struct { char internal[128]; } NcclUniqueId;
typedef void *NcclComm;
enum { kNcclSum = 0, kNcclMax = 2, kNcclFloat64 = 8 };

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static NcclComm g_comm = nullptr;
static int g_rank = 0, g_nranks = 1;

#define EFV_NCCL(call)                                                                                        \
  do {                                                                                                        \
    int r__ = (call);                                                                                         \
    if (r__ != 0) throw Error(std::string(#call) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?")); \
  } while (0)

static void load_nccl(const char *path) {
  if (g_nccl.handle) return;
  const char *cands[] = {path, "libnccl.so.2", "libnccl.so"};
  for (const char *c : cands) {
    if (!c || !*c) continue;
    g_nccl.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) throw Error("cannot dlopen NCCL");
#define SYM(field, name)                                                      \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.handle, name);                     \
  if (!g_nccl.field) throw Error(std::string("NCCL symbol missing: ") + name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
}

void comm_unique_id(const char *nccl_path, char *out128) {
  load_nccl(nccl_path);
  NcclUniqueId id;
  EFV_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, id.internal, 128);
}
static bool g_comm_on = false;
// id128 == nullptr: peer-memory-only communicator (no NCCL; every collective must then run over the CUDA-IPC peer
// regions).  This is what lets two ranks share ONE GPU (NCCL refuses duplicate devices), used by the 1-GPU parity test.
void comm_init(const char *nccl_path, int rank, int nranks, const char *id128) {
  EFV_REQUIRE(!g_comm_on, "communicator already initialised");
  (void)stream();
  if (id128) {
    load_nccl(nccl_path);
    NcclUniqueId id;
    memcpy(id.internal, id128, 128);
    EFV_NCCL(g_nccl.CommInitRank(&g_comm, nranks, id, rank));
  }
  g_rank = rank;
  g_nranks = nranks;
  g_comm_on = true;
}
void comm_peer_release();
void comm_destroy() {
  comm_peer_release();
  if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
  g_rank = 0; g_nranks = 1; g_comm_on = false;
}
bool comm_active() { return g_comm_on && g_nranks > 1; }
static void require_nccl(const char *what) {
  if (!g_comm) throw Error(std::string(what) + " needs NCCL, but the communicator was created in peer-memory-only mode "
                           "(or the peer regions could not be mapped): set EFV_COMM=nccl with one GPU per rank");
}
int comm_rank() { return g_rank; }
int comm_size() { return g_nranks; }

// ---- peer-memory collectives (NVLink / NVSwitch loads and stores, no NCCL on the critical path) ---------------------------
// Layout of the per-rank region and the device pieces fused into the Krylov kernels: peer.cuh.  k_allreduce_peer is
// the stand-alone form (one launch = fold partials + exchange) used outside the PCG iteration.
struct PeerRegion {
  bool active = false;
  char *local = nullptr;
  PeerPtrs all = {};
  uint32_t seq = 0;
};
static PeerRegion g_peer;

PeerAR comm_peer_ar(unsigned *counter) {
  EFV_REQUIRE(g_peer.active, "peer collectives are not initialised");
  PeerAR ar;
  ar.peers = g_peer.all;
  ar.me = g_rank;
  ar.nranks = g_nranks;
  ar.counter = counter;
  ar.seq = ++g_peer.seq;
  ar.enabled = 1;
  return ar;
}

// op: 0 sum, 1 max.  narr <= 4 arrays of G partials each.
// inplace != nullptr: additionally rewrite the partial arrays as (global value, 0, 0, ...) so that kernels which re-sum
// G partials read the all-reduced scalar without knowing about ranks (BiCGStab)
__global__ void k_allreduce_peer(const double *partial, int G, int narr, double *__restrict__ out, PeerPtrs peers,
                                 int me, int nranks, uint32_t seq, int op, double *inplace) {
  __shared__ double red[32];
  __shared__ double mine[4];
  for (int a = 0; a < narr; ++a) {
    double v = 0.0;
    if (op == 0) {
      for (int i = threadIdx.x; i < G; i += blockDim.x) v += partial[(size_t)a * G + i];
      v = block_sum(v, red);
    } else {
      for (int i = threadIdx.x; i < G; i += blockDim.x) v = fmax(v, partial[(size_t)a * G + i]);
      v = block_max(v, red);
    }
    if (threadIdx.x == 0) mine[a] = v;
    __syncthreads();
  }
  const int parity = seq & 1;
  if (threadIdx.x < nranks) {
    char *dst = peers.p[threadIdx.x];
    double *slot = peer_slot(dst, parity, me);
    for (int a = 0; a < narr; ++a) slot[a] = mine[a];
    __threadfence_system();
    *peer_flag(dst, parity, me) = seq;
  }
  char *self = peers.p[me];
  if (threadIdx.x < nranks) peer_wait_flag(peer_flag(self, parity, threadIdx.x), seq, peer_error(self), 1);
  __syncthreads();
  __threadfence_system();
  if (threadIdx.x < narr) {
    double acc = 0.0;
    for (int r = 0; r < nranks; ++r) {
      const double v = *reinterpret_cast<volatile double *>(peer_slot(self, parity, r) + threadIdx.x);
      acc = (op == 0) ? acc + v : fmax(acc, v);
    }
    if (out) out[threadIdx.x] = acc;
    if (inplace) inplace[(size_t)threadIdx.x * G] = acc;
  }
  if (inplace)
    for (int i = threadIdx.x; i < narr * G; i += blockDim.x)
      if (i % G != 0) inplace[i] = 0.0;
}

void comm_peer_export(char *out64) {
  EFV_REQUIRE(comm_active(), "initialise the communicator first");
  EFV_REQUIRE(g_nranks <= kMaxRanks, "peer collectives support at most 16 ranks");
  if (!g_peer.local) {
    EFV_CUDA(cudaMalloc((void **)&g_peer.local, kPeerBytes));
    EFV_CUDA(cudaMemset(g_peer.local, 0, kPeerBytes));
  }
  cudaIpcMemHandle_t h;
  EFV_CUDA(cudaIpcGetMemHandle(&h, g_peer.local));
  static_assert(sizeof(h) == 64, "IPC handle size");
  memcpy(out64, &h, 64);
}
void comm_peer_import(const char *handles) {
  EFV_REQUIRE(g_peer.local, "export the peer region first");
  for (int r = 0; r < g_nranks; ++r) g_peer.all.p[r] = nullptr;
  for (int r = 0; r < g_nranks; ++r) {
    if (r == g_rank) { g_peer.all.p[r] = g_peer.local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {                      // another node / no peer access: the caller falls back to NCCL on all ranks
      cudaGetLastError();
      comm_peer_release();
      throw Error(std::string("cudaIpcOpenMemHandle failed for rank ") + std::to_string(r) + ": " + cudaGetErrorString(e));
    }
    g_peer.all.p[r] = (char *)ptr;
  }
  g_peer.active = true;
  g_peer.seq = 0;
}
// unmap the peers' regions and stop using peer-memory collectives (the local region stays allocated: cheap, and other
// ranks may still have it mapped)
void comm_peer_release() {
  for (int r = 0; r < kMaxRanks; ++r) {
    if (r != g_rank && g_peer.all.p[r]) cudaIpcCloseMemHandle(g_peer.all.p[r]);
    g_peer.all.p[r] = nullptr;
  }
  cudaGetLastError();
  g_peer.active = false;
}
bool comm_peer_active() { return g_peer.active && comm_active(); }

// all-reduce `narr` arrays of G per-block partials into out[0..narr) (device), on every rank
void comm_allreduce_partials(const double *partial, int G, int narr, double *out, int op) {
  EFV_REQUIRE(comm_peer_active(), "peer collectives are not initialised");
  EFV_REQUIRE(narr >= 1 && narr <= 4, "at most 4 scalars per all-reduce");
  g_peer.seq++;
  EFV_LAUNCH(k_allreduce_peer, 1, 256, 0, partial, G, narr, out, g_peer.all, g_rank, g_nranks, g_peer.seq, op, (double *)nullptr);
}
__global__ void k_fold_partials(double *partial, int G, int narr, double *__restrict__ out, int scatter) {
  __shared__ double red[32];
  if (!scatter) {
    for (int a = 0; a < narr; ++a) {
      double v = 0.0;
      for (int i = threadIdx.x; i < G; i += blockDim.x) v += partial[(size_t)a * G + i];
      v = block_sum(v, red);
      if (threadIdx.x == 0) out[a] = v;
      __syncthreads();
    }
  } else {
    for (int i = threadIdx.x; i < narr * G; i += blockDim.x) partial[i] = (i % G == 0) ? out[i / G] : 0.0;
  }
}
// sum every one of the `narr` arrays of G partials over the ranks and leave (sum, 0, ..., 0) in its place
void comm_allreduce_partials_inplace(double *partial, int G, int narr) {
  if (!comm_active()) return;
  EFV_REQUIRE(narr >= 1 && narr <= 4, "at most 4 scalars per all-reduce");
  if (comm_peer_active()) {
    g_peer.seq++;
    EFV_LAUNCH(k_allreduce_peer, 1, 256, 0, partial, G, narr, (double *)nullptr, g_peer.all, g_rank, g_nranks, g_peer.seq, 0, partial);
    return;
  }
  require_nccl("all-reduce without peer regions");
  static double *tmp = nullptr;                      // process lifetime (no static destructor touching CUDA at exit)
  if (!tmp) EFV_CUDA(cudaMalloc((void **)&tmp, 4 * sizeof(double)));
  EFV_LAUNCH(k_fold_partials, 1, 256, 0, partial, G, narr, tmp, 0);
  EFV_NCCL(g_nccl.AllReduce(tmp, tmp, (size_t)narr, kNcclFloat64, kNcclSum, g_comm, stream()));
  EFV_LAUNCH(k_fold_partials, 1, 256, 0, partial, G, narr, tmp, 1);
}
// error word of the peer waits (1: all-reduce, 2: halo); reading it clears it, so one failed solve does not poison the next
int comm_peer_error() {
  if (!g_peer.local) return 0;
  int e = 0;
  EFV_CUDA(cudaMemcpyAsync(&e, g_peer.local + kSlotBytes + kFlagBytes, sizeof(int), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
  if (e) EFV_CUDA(cudaMemsetAsync(g_peer.local + kSlotBytes + kFlagBytes, 0, sizeof(int), stream()));
  return e;
}

void comm_allreduce_sum(double *dev, int count) {
  if (!comm_active()) return;
  if (comm_peer_active() && count <= 4) { comm_allreduce_partials(dev, 1, count, dev, 0); return; }
  require_nccl("all-reduce without peer regions");
  EFV_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, kNcclFloat64, kNcclSum, g_comm, stream()));
}
void comm_allreduce_max(double *dev, int count) {
  if (!comm_active()) return;
  if (comm_peer_active() && count <= 4) { comm_allreduce_partials(dev, 1, count, dev, 1); return; }
  require_nccl("all-reduce without peer regions");
  EFV_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, kNcclFloat64, kNcclMax, g_comm, stream()));
}

// ---- mailbox: small host-driven point-to-point messages over the peer regions --------------------------------------------
// Used at set-up time only (IPC handles and offsets of per-level halo buffers, row counts), where torch.distributed is
// out of reach of the C library and NCCL may be absent (ranks sharing a GPU).  Collective in COUNT: every rank must call
// comm_mail_exchange the same number of times (possibly with no peers), because the boxes are sequence-numbered.
struct MailMsg { unsigned long long w[kMailBytes / 8]; };
static uint32_t g_mail_seq = 0;
__global__ void k_mail_send(char *dst, int me, MailMsg msg, uint32_t seq) {
  unsigned long long *box = reinterpret_cast<unsigned long long *>(dst + kMailOffset + ((size_t)(seq & 1u) * kMaxRanks + me) * kMailBytes);
  if (threadIdx.x < kMailBytes / 8) box[threadIdx.x] = msg.w[threadIdx.x];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t *>(dst + kMailFlagOffset + ((size_t)(seq & 1u) * kMaxRanks + me) * sizeof(uint32_t)) = seq;
}
__global__ void k_mail_recv(char *self, int from, uint32_t seq, unsigned long long *__restrict__ out) {
  if (threadIdx.x == 0)
    peer_wait_flag(reinterpret_cast<volatile const uint32_t *>(self + kMailFlagOffset + ((size_t)(seq & 1u) * kMaxRanks + from) * sizeof(uint32_t)), seq,
                   peer_error(self), 3);
  __syncthreads();
  __threadfence_system();
  const volatile unsigned long long *box =
      reinterpret_cast<const volatile unsigned long long *>(self + kMailOffset + ((size_t)(seq & 1u) * kMaxRanks + from) * kMailBytes);
  if (threadIdx.x < kMailBytes / 8) out[threadIdx.x] = box[threadIdx.x];
}
// send one kMailBytes message to every rank of `peers` and receive one from each (recv: npeers * kMailBytes, peer order)
void comm_mail_exchange(int npeers, const int *peers, const char *send, char *recv) {
  EFV_REQUIRE(comm_peer_active(), "the mailbox needs the peer-memory regions");
  g_mail_seq++;
  DBuf<unsigned long long> tmp((size_t)std::max(npeers, 1) * (kMailBytes / 8));
  for (int i = 0; i < npeers; ++i) {
    MailMsg m;
    memcpy(m.w, send + (size_t)i * kMailBytes, kMailBytes);
    EFV_LAUNCH(k_mail_send, 1, 32, 0, g_peer.all.p[peers[i]], g_rank, m, g_mail_seq);
  }
  for (int i = 0; i < npeers; ++i) EFV_LAUNCH(k_mail_recv, 1, 32, 0, g_peer.local, peers[i], g_mail_seq, tmp.p + (size_t)i * (kMailBytes / 8));
  if (npeers) EFV_CUDA(cudaMemcpyAsync(recv, tmp.p, (size_t)npeers * kMailBytes, cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
  EFV_REQUIRE(comm_peer_error() == 0, "mailbox exchange timed out (a rank did not arrive)");
}
// one int64 per rank, gathered on every rank
void comm_allgather_i64(int64_t mine, int64_t *all) {
  std::vector<int> peers;
  for (int r = 0; r < g_nranks; ++r) if (r != g_rank) peers.push_back(r);
  std::vector<char> send(peers.size() * kMailBytes, 0), recv(peers.size() * kMailBytes, 0);
  for (size_t i = 0; i < peers.size(); ++i) memcpy(send.data() + i * kMailBytes, &mine, sizeof(int64_t));
  comm_mail_exchange((int)peers.size(), peers.data(), send.data(), recv.data());
  all[g_rank] = mine;
  for (size_t i = 0; i < peers.size(); ++i) memcpy(&all[peers[i]], recv.data() + i * kMailBytes, sizeof(int64_t));
}

// ---- gather stage: every rank's slice of a short vector on every rank (coarsest multigrid level) ---------------------------
// Buffer of every rank: [flags 2 x kMaxRanks | data parity 0 | data parity 1], mapped by all ranks through CUDA IPC.
constexpr size_t kGatherFlagBytes = 2 * kMaxRanks * sizeof(uint32_t) + 128;
struct GatherPtrs { char *p[kMaxRanks]; };
__global__ void k_gather_push(GatherPtrs all, int nranks, int me, size_t cap, int64_t offset, int64_t count, const double *__restrict__ src,
                              uint32_t seq, unsigned *__restrict__ ticket, const int *__restrict__ done) {
  // `done` is deliberately NOT honoured for the flags: peers that are still iterating would wait for them
  const size_t par = seq & 1u;
  const bool skip = done && *done;
  if (!skip)
    for (int r = 0; r < nranks; ++r) {
      double *dst = reinterpret_cast<double *>(all.p[r] + kGatherFlagBytes) + par * cap + offset;
      for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
    }
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if ((int)threadIdx.x < nranks) *reinterpret_cast<volatile uint32_t *>(all.p[threadIdx.x] + (par * kMaxRanks + me) * sizeof(uint32_t)) = seq;
  if (threadIdx.x == 0) *ticket = 0u;
}
__global__ void k_gather_wait_copy(const char *__restrict__ self, int nranks, size_t cap, int64_t count, uint32_t seq, double *__restrict__ dst,
                                   int *__restrict__ err) {
  const size_t par = seq & 1u;
  if ((int)threadIdx.x < nranks) peer_wait_flag(reinterpret_cast<volatile const uint32_t *>(self) + par * kMaxRanks + threadIdx.x, seq, err, 4);
  __syncthreads();
  __threadfence_system();
  const volatile double *src = reinterpret_cast<const volatile double *>(self + kGatherFlagBytes) + par * cap;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
struct GatherStage {
  char *local = nullptr;
  GatherPtrs all = {};
  size_t cap = 0;
  unsigned *ticket = nullptr;
  uint32_t seq = 0;
};
GatherStage *gather_stage_create(size_t cap_doubles) {
  EFV_REQUIRE(comm_peer_active(), "the gather stage needs the peer-memory regions");
  GatherStage *g = new GatherStage();
  g->cap = cap_doubles;
  const size_t bytes = kGatherFlagBytes + 2 * cap_doubles * sizeof(double) + 64;
  EFV_CUDA(cudaMalloc((void **)&g->local, bytes));
  EFV_CUDA(cudaMemset(g->local, 0, bytes));
  EFV_CUDA(cudaMalloc((void **)&g->ticket, sizeof(unsigned)));
  EFV_CUDA(cudaMemset(g->ticket, 0, sizeof(unsigned)));
  cudaIpcMemHandle_t mine;
  EFV_CUDA(cudaIpcGetMemHandle(&mine, g->local));
  std::vector<int> peers;
  for (int r = 0; r < g_nranks; ++r) if (r != g_rank) peers.push_back(r);
  std::vector<char> send(peers.size() * kMailBytes, 0), recv(peers.size() * kMailBytes, 0);
  for (size_t i = 0; i < peers.size(); ++i) memcpy(send.data() + i * kMailBytes, &mine, 64);
  comm_mail_exchange((int)peers.size(), peers.data(), send.data(), recv.data());
  g->all.p[g_rank] = g->local;
  for (size_t i = 0; i < peers.size(); ++i) {
    cudaIpcMemHandle_t h;
    memcpy(&h, recv.data() + i * kMailBytes, 64);
    void *ptr = nullptr;
    EFV_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    g->all.p[peers[i]] = (char *)ptr;
  }
  return g;
}
void gather_stage_destroy(GatherStage *g) {
  if (!g) return;
  for (int r = 0; r < kMaxRanks; ++r) if (r != g_rank && g->all.p[r]) cudaIpcCloseMemHandle(g->all.p[r]);
  if (g->local) cudaFree(g->local);
  if (g->ticket) cudaFree(g->ticket);
  cudaGetLastError();
  delete g;
}
// dst[0, total) <- concatenation over ranks of their (offset, count) slices of src; `done`: the push of a finished solve
// still raises its flags (see k_gather_push)
void gather_all(GatherStage *g, int64_t offset, int64_t count, int64_t total, const double *src, double *dst, const int *done) {
  EFV_REQUIRE((size_t)total <= g->cap, "gather larger than the stage");
  g->seq++;
  EFV_LAUNCH(k_gather_push, grid_for(std::max<int64_t>(count, 1), 256, 1), 256, 0, g->all, g_nranks, g_rank, g->cap, offset, count, src, g->seq,
             g->ticket, done);
  EFV_LAUNCH(k_gather_wait_copy, grid_for(total, 256, 1), 256, 0, g->local, g_nranks, g->cap, total, g->seq, dst,
             reinterpret_cast<int *>(g_peer.local + kSlotBytes + kFlagBytes));
}

// connect the peer staging buffers of a halo whose lists are already registered (C-side version of what the host layer
// does for the finest level with all_gather_object): neighbours swap (IPC handle, ghost count, offset, flag slot)
void halo_connect_peer(Halo &h) {
  EFV_REQUIRE(h.active, "register the halo first");
  const int n = (int)h.nbr.size();
  char mine[64];
  halo_stage_export(h, mine);
  std::vector<char> send((size_t)std::max(n, 1) * kMailBytes, 0), recv((size_t)std::max(n, 1) * kMailBytes, 0);
  for (int i = 0; i < n; ++i) {
    char *m = send.data() + (size_t)i * kMailBytes;
    memcpy(m, mine, 64);
    const int64_t info[3] = {h.n_ghost, h.recv_ptr[i], (int64_t)i};      // my ghost count; where neighbour i's values go; its flag slot
    memcpy(m + 64, info, sizeof(info));
  }
  comm_mail_exchange(n, h.nbr.data(), send.data(), recv.data());
  std::vector<char> handles((size_t)std::max(n, 1) * 64);
  std::vector<int64_t> off(n), ghosts(n);
  std::vector<int> slot(n);
  for (int i = 0; i < n; ++i) {
    const char *m = recv.data() + (size_t)i * kMailBytes;
    memcpy(handles.data() + (size_t)i * 64, m, 64);
    int64_t info[3];
    memcpy(info, m + 64, sizeof(info));
    ghosts[i] = info[0]; off[i] = info[1]; slot[i] = (int)info[2];
  }
  halo_stage_import(h, handles.data(), off.data(), slot.data(), ghosts.data());
}

// ---- halo exchange ----------------------------------------------------------------------------------------------------
__global__ void k_pack(const int32_t *__restrict__ idx, int64_t n, int ndof, const double *__restrict__ vec,
                       double *__restrict__ buf) {
  int64_t total = n * ndof;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = i / ndof;
    int f = (int)(i % ndof);
    buf[i] = vec[(int64_t)idx[k] * ndof + f];
  }
}

void halo_set(efv_system *s, int n_nbr, const int *nbr, const int64_t *send_ptr, const int32_t *send_idx,
              const int64_t *recv_ptr) {
  Halo &h = s->halo;
  EFV_REQUIRE(n_nbr <= kMaxHaloNbr, "a rank can have at most 15 halo neighbours");
  h.nbr.assign(nbr, nbr + n_nbr);
  h.send_ptr.assign(send_ptr, send_ptr + n_nbr + 1);
  h.recv_ptr.assign(recv_ptr, recv_ptr + n_nbr + 1);
  EFV_REQUIRE(h.recv_ptr[n_nbr] == s->nV - s->n_owned, "halo receive lists must cover the ghost vertices");
  for (int64_t k = 0; k < h.send_ptr[n_nbr]; ++k)
    EFV_REQUIRE(send_idx[k] >= 0 && send_idx[k] < s->n_owned, "halo send index is not an owned vertex");
  h.n_owned = s->n_owned;
  h.n_ghost = s->nV - s->n_owned;
  h.cap_ndof = s->ndof;
  h.send_idx.upload(send_idx, h.send_ptr[n_nbr]);
  h.send_idx_host.assign(send_idx, send_idx + h.send_ptr[n_nbr]);
  h.send_buf.alloc((size_t)h.send_ptr[n_nbr] * s->ndof);
  EFV_CUDA(cudaStreamSynchronize(stream()));
  h.active = true;
}

// peer-memory halo: every rank owns a staging buffer [flag block | ghost values (parity 0) | ghost values (parity 1)];
// neighbours store their interface values straight into it over NVLink and then raise the flag of their slot; the
// owner waits for the flags and copies the values behind its owned entries.  Buffers and flags are double-buffered
// by the parity of the exchange number: a neighbour can run at most one exchange ahead (its exchange k needs my push
// k), so its push k+1 lands in the other buffer while I may still be copying k, and k+2 cannot be pushed before my
// push k+1, which is stream-ordered after my copy of k.  Exchanges therefore need no collective between them.
__global__ void k_halo_push(HaloPush hp, const int32_t *__restrict__ idx, int ndof, const double *__restrict__ vec, uint32_t seq,
                            unsigned *__restrict__ ticket) {
  const int64_t total = hp.begin[hp.n] * ndof;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / ndof;
    const int f = (int)(i % ndof);
    int nb = 0;
    while (k >= hp.begin[nb + 1]) ++nb;
    hp.dst[nb][(k - hp.begin[nb]) * ndof + f] = vec[(int64_t)idx[k] * ndof + f];
  }
  // the last block to finish raises the flags (all stores of all blocks are then visible system-wide)
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < hp.n) *reinterpret_cast<volatile uint32_t *>(hp.flag[threadIdx.x]) = seq;
  if (threadIdx.x == 0) *ticket = 0u;
}
__global__ void k_halo_wait_copy(const char *__restrict__ stage, const double *__restrict__ src, int n_nbr, uint32_t seq, int64_t count,
                                 double *__restrict__ dst, int *__restrict__ err) {
  if (threadIdx.x < n_nbr)
    peer_wait_flag(reinterpret_cast<volatile const uint32_t *>(stage) + (seq & 1u) * kMaxRanks + threadIdx.x, seq, err, 2);
  __syncthreads();
  __threadfence_system();
  const volatile double *vsrc = src;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) dst[i] = vsrc[i];
}

static size_t halo_stage_bytes(const Halo &h) { return kHaloFlagBytes + 2 * (size_t)h.n_ghost * h.cap_ndof * sizeof(double) + 64; }

void halo_stage_export(Halo &h, char *out64) {
  EFV_REQUIRE(h.active, "register the halo first");
  if (!h.stage) {
    EFV_CUDA(cudaMalloc((void **)&h.stage, halo_stage_bytes(h)));
    EFV_CUDA(cudaMemset(h.stage, 0, halo_stage_bytes(h)));
    EFV_CUDA(cudaMalloc((void **)&h.ticket, sizeof(unsigned)));
    EFV_CUDA(cudaMemset(h.ticket, 0, sizeof(unsigned)));
  }
  cudaIpcMemHandle_t hd;
  EFV_CUDA(cudaIpcGetMemHandle(&hd, h.stage));
  memcpy(out64, &hd, 64);
}
// handles: one per neighbour (same order as the neighbour list); remote_off[i]: first ghost slot (in vertices) of
// neighbour i's staging area that belongs to this rank; remote_slot[i]: which flag of neighbour i this rank raises;
// remote_ghosts[i]: number of ghost vertices of neighbour i (= the stride between its two parity buffers)
void halo_stage_import(Halo &h, const char *handles, const int64_t *remote_off, const int *remote_slot, const int64_t *remote_ghosts) {
  EFV_REQUIRE(h.stage, "export the staging buffer first");
  const int n = (int)h.nbr.size();
  h.remote.assign(n, nullptr);
  h.remote_off.assign(remote_off, remote_off + n);
  h.remote_slot.assign(remote_slot, remote_slot + n);
  h.remote_ghosts.assign(remote_ghosts, remote_ghosts + n);
  for (int i = 0; i < n; ++i) {
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handles + 64 * (size_t)i, 64);
    void *ptr = nullptr;
    EFV_CUDA(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    h.remote[i] = (char *)ptr;
  }
  h.peer = true;
  h.seq = 0;
}
void halo_release(Halo &h) {
  for (char *r : h.remote) if (r) cudaIpcCloseMemHandle(r);
  h.remote.clear();
  if (h.stage) cudaFree(h.stage);
  if (h.ticket) cudaFree(h.ticket);
  h.stage = nullptr; h.ticket = nullptr; h.peer = false; h.active = false;
  cudaGetLastError();
}

// Peer path, first half: store my interface values into the neighbours' staging buffers and raise their flags.
// Returns false when the halo has no peer buffers (the caller then uses halo_exchange, i.e. NCCL).
bool halo_push(Halo &h, const double *vec, int ndof) {
  if (!h.active || !comm_active()) return false;
  if (!(h.peer && comm_peer_active())) return false;
  EFV_REQUIRE(ndof <= h.cap_ndof, "halo exchange wider than the registered capacity");
  const int n_nbr = (int)h.nbr.size();
  const int64_t n_send = h.send_ptr[n_nbr];
  h.seq++;
  const uint32_t par = h.seq & 1u;
  HaloPush hp;
  hp.n = n_nbr;
  for (int i = 0; i < n_nbr; ++i) {
    hp.dst[i] = reinterpret_cast<double *>(h.remote[i] + kHaloFlagBytes) + (size_t)par * h.remote_ghosts[i] * h.cap_ndof +
                h.remote_off[i] * ndof;
    hp.flag[i] = reinterpret_cast<uint32_t *>(h.remote[i]) + par * kMaxRanks + h.remote_slot[i];
    hp.begin[i] = h.send_ptr[i];
  }
  hp.begin[n_nbr] = n_send;
  EFV_LAUNCH(k_halo_push, grid_for(std::max<int64_t>(n_send * ndof, 1), 256, 2), 256, 0, hp, h.send_idx.p, ndof, vec, h.seq, h.ticket);
  return true;
}
// What a consumer kernel needs to read the ghosts of the exchange just pushed straight from the staging buffer: it waits
// for the neighbours' flags only when (and where) it first touches a ghost, so rows without ghosts overlap the transfer.
HaloIn halo_in(const Halo &h, int ndof) {
  HaloIn v;
  const uint32_t par = h.seq & 1u;
  v.ghost = reinterpret_cast<const double *>(h.stage + kHaloFlagBytes) + (size_t)par * h.n_ghost * h.cap_ndof;
  v.flags = reinterpret_cast<const uint32_t *>(h.stage) + par * kMaxRanks;
  v.n_nbr = (int)h.nbr.size();
  v.seq = h.seq;
  v.n_owned = h.n_owned;
  v.err = reinterpret_cast<int *>(g_peer.local + kSlotBytes + kFlagBytes);
  (void)ndof;
  return v;
}
// Peer path, second half for consumers that need the ghosts inside the vector: wait for the flags, copy behind the owned entries
void halo_wait_copy(Halo &h, double *vec, int ndof) {
  const int n_nbr = (int)h.nbr.size();
  const uint32_t par = h.seq & 1u;
  const int64_t count = h.n_ghost * ndof;
  const double *src = reinterpret_cast<const double *>(h.stage + kHaloFlagBytes) + (size_t)par * h.n_ghost * h.cap_ndof;
  EFV_LAUNCH(k_halo_wait_copy, grid_for(count, 256, 1), 256, 0, h.stage, src, n_nbr, h.seq, count, vec + h.n_owned * ndof,
             reinterpret_cast<int *>(g_peer.local + kSlotBytes + kFlagBytes));
}

// ghost entries of `vec` (ndof values per vertex, ndof <= the capacity the halo was registered with) <- owners' values
void halo_exchange(Halo &h, double *vec, int ndof) {
  if (!h.active || !comm_active()) return;
  EFV_REQUIRE(ndof <= h.cap_ndof, "halo exchange wider than the registered capacity");
  const int n_nbr = (int)h.nbr.size();
  const int64_t n_send = h.send_ptr[n_nbr];
  if (halo_push(h, vec, ndof)) {
    halo_wait_copy(h, vec, ndof);
    return;
  }
  require_nccl("halo exchange without peer staging buffers");
  if ((size_t)n_send * ndof > h.send_buf.n) h.send_buf.alloc((size_t)n_send * ndof);
  if (n_send) EFV_LAUNCH(k_pack, grid_for(n_send * ndof, 256), 256, 0, h.send_idx.p, n_send, ndof, vec, h.send_buf.p);
  EFV_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < n_nbr; ++i) {
    const int64_t ns = (h.send_ptr[i + 1] - h.send_ptr[i]) * ndof, nr = (h.recv_ptr[i + 1] - h.recv_ptr[i]) * ndof;
    if (ns) EFV_NCCL(g_nccl.Send(h.send_buf.p + h.send_ptr[i] * ndof, (size_t)ns, kNcclFloat64, h.nbr[i], g_comm, stream()));
    if (nr) EFV_NCCL(g_nccl.Recv(vec + (h.n_owned + h.recv_ptr[i]) * ndof, (size_t)nr, kNcclFloat64, h.nbr[i], g_comm, stream()));
  }
  EFV_NCCL(g_nccl.GroupEnd());
}
void halo_exchange(efv_system *s, double *vec) { halo_exchange(s->halo, vec, s->ndof); }

}  // namespace efv
