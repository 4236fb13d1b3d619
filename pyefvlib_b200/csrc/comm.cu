// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL communicator created from a unique id that the host
// layer distributes (torch.distributed), halo exchange of owned vertex values and all-reduce of Krylov scalars.
// NCCL is dlopen'ed (the path of the copy already loaded by torch is handed in) so that the process holds a single
// NCCL instance.  Everything runs on the library stream.
#include "views.cuh"
#include <dlfcn.h>
#include <cstring>

namespace efv {

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
void comm_init(const char *nccl_path, int rank, int nranks, const char *id128) {
  load_nccl(nccl_path);
  EFV_REQUIRE(!g_comm, "communicator already initialised");
  NcclUniqueId id;
  memcpy(id.internal, id128, 128);
  (void)stream();
  EFV_NCCL(g_nccl.CommInitRank(&g_comm, nranks, id, rank));
  g_rank = rank;
  g_nranks = nranks;
}
void comm_destroy() {
  if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
  g_rank = 0; g_nranks = 1;
}
bool comm_active() { return g_comm != nullptr && g_nranks > 1; }
int comm_rank() { return g_rank; }
int comm_size() { return g_nranks; }

void comm_allreduce_sum(double *dev, int count) {
  if (!comm_active()) return;
  EFV_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, kNcclFloat64, kNcclSum, g_comm, stream()));
}
void comm_allreduce_max(double *dev, int count) {
  if (!comm_active()) return;
  EFV_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, kNcclFloat64, kNcclMax, g_comm, stream()));
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
  h.nbr.assign(nbr, nbr + n_nbr);
  h.send_ptr.assign(send_ptr, send_ptr + n_nbr + 1);
  h.recv_ptr.assign(recv_ptr, recv_ptr + n_nbr + 1);
  EFV_REQUIRE(h.recv_ptr[n_nbr] == s->nV - s->n_owned, "halo receive lists must cover the ghost vertices");
  for (int64_t k = 0; k < h.send_ptr[n_nbr]; ++k)
    EFV_REQUIRE(send_idx[k] >= 0 && send_idx[k] < s->n_owned, "halo send index is not an owned vertex");
  h.send_idx.upload(send_idx, h.send_ptr[n_nbr]);
  h.send_buf.alloc((size_t)h.send_ptr[n_nbr] * s->ndof);
  EFV_CUDA(cudaStreamSynchronize(stream()));
  h.active = true;
}

// ghost entries of `vec` <- owner ranks' values
void halo_exchange(efv_system *s, double *vec) {
  Halo &h = s->halo;
  if (!h.active || !comm_active()) return;
  const int n_nbr = (int)h.nbr.size();
  const int64_t n_send = h.send_ptr[n_nbr];
  if (n_send) EFV_LAUNCH(k_pack, grid_for(n_send * s->ndof, 256), 256, 0, h.send_idx.p, n_send, s->ndof, vec, h.send_buf.p);
  EFV_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < n_nbr; ++i) {
    const int64_t ns = (h.send_ptr[i + 1] - h.send_ptr[i]) * s->ndof, nr = (h.recv_ptr[i + 1] - h.recv_ptr[i]) * s->ndof;
    if (ns) EFV_NCCL(g_nccl.Send(h.send_buf.p + h.send_ptr[i] * s->ndof, (size_t)ns, kNcclFloat64, h.nbr[i], g_comm, stream()));
    if (nr) EFV_NCCL(g_nccl.Recv(vec + (s->n_owned + h.recv_ptr[i]) * s->ndof, (size_t)nr, kNcclFloat64, h.nbr[i], g_comm, stream()));
  }
  EFV_NCCL(g_nccl.GroupEnd());
}

}  // namespace efv
