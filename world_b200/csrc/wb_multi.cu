// wb_multi.cu -- the one exchange step of the multi-GPU path: utterances are sharded over ranks
// with no data-path collective; the output arrays are reassembled on every rank by NCCL.  One process (or thread)
// per GPU; the caller moves the 128-byte NCCL id from rank 0 to the other ranks by whatever means it has (MPI,
// torch.distributed, a socket, a file) -- that is plumbing, the data path is here.
//
// NCCL is bound at RUN time (dlopen "libnccl.so.2"): the library keeps linking against cudart only, a process that
// already carries an NCCL (PyTorch's) shares it, and single-GPU users never load it.
#include "wb_internal.h"
#include <string>

#ifndef WB_EMU
#include <dlfcn.h>
#include <nccl.h>

namespace wb {

namespace {
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool nccl_load(std::string *err) {
  if (g_nccl.lib) return true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) { *err = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return false; }
#define WB_SYM(field, name)                                                        \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name));       \
  if (!g_nccl.field) { *err = std::string("NCCL symbol missing: ") + name; return false; }
  WB_SYM(GetUniqueId, "ncclGetUniqueId")
  WB_SYM(CommInitRank, "ncclCommInitRank")
  WB_SYM(CommDestroy, "ncclCommDestroy")
  WB_SYM(Broadcast, "ncclBroadcast")
  WB_SYM(AllGather, "ncclAllGather")
  WB_SYM(GroupStart, "ncclGroupStart")
  WB_SYM(GroupEnd, "ncclGroupEnd")
  WB_SYM(GetErrorString, "ncclGetErrorString")
#undef WB_SYM
  g_nccl.lib = lib;
  return true;
}
}  // namespace

struct Comm {
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0;
  cudaStream_t stream = nullptr;      // all collectives run here, ordered against the compute streams by events
  cudaEvent_t done = nullptr;
};

#define WB_NCCL(call, err)                                                                         \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess) { *(err) = std::string("NCCL: ") + g_nccl.GetErrorString(r_); return 1; } \
  } while (0)

int comm_unique_id(unsigned char *id128, std::string *err) {
  if (!nccl_load(err)) return 1;
  ncclUniqueId id;
  WB_NCCL(g_nccl.GetUniqueId(&id), err);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in every NCCL 2.x");
  memcpy(id128, &id, 128);
  return 0;
}

int comm_create(int n_ranks, int rank, const unsigned char *id128, Comm **out, std::string *err) {
  if (!nccl_load(err)) return 1;
  Comm *c = new Comm;
  c->n_ranks = n_ranks; c->rank = rank;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != ncclSuccess) { *err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); delete c; return 1; }
  // highest stream priority: when a compute kernel's CTA retires, the waiting blocks of the collective are placed
  // first -- at equal priority the blocks of the (earlier launched, 10^5-block) compute kernels keep filling the SMs
  // and the transfer only runs in their tails (profiles/r2j: 8 ranks at 0.84 with the transfer enqueued "under" the compute)
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming) != cudaSuccess) {
    *err = "cannot create the communication stream";
    g_nccl.CommDestroy(c->comm);
    delete c;
    return 1;
  }
  *out = c;
  return 0;
}

void comm_destroy(Comm *c) {
  if (!c) return;
  cudaStreamSynchronize(c->stream);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  cudaEventDestroy(c->done);
  cudaStreamDestroy(c->stream);
  delete c;
}

int comm_ranks(const Comm *c) { return c->n_ranks; }
int comm_rank(const Comm *c) { return c->rank; }

// Rows [row0, row0 + rows) of every rank's shard: `full` is [n_ranks][rows_per_rank][row_elems]; rank r owns block r
// and its rows are already in place.  One grouped set of broadcasts (root r sends its rows, everybody else receives
// them where they belong), enqueued on the communication stream after `after` (an event on a compute stream).
int comm_gather_rows(Comm *c, double *full, size_t row_elems, size_t rows_per_rank, size_t row0, size_t rows,
                     cudaEvent_t after, std::string *err) {
  if (rows == 0 || row_elems == 0) return 0;
  if (after && cudaStreamWaitEvent(c->stream, after, 0) != cudaSuccess) { *err = "cudaStreamWaitEvent failed"; return 1; }
  if (row0 == 0 && rows == rows_per_rank) {   // whole shards: the plain in-place all-gather
    WB_NCCL(g_nccl.AllGather(full + (size_t)c->rank * rows_per_rank * row_elems, full, rows_per_rank * row_elems, ncclDouble,
                             c->comm, c->stream), err);
    return 0;
  }
  WB_NCCL(g_nccl.GroupStart(), err);
  for (int r = 0; r < c->n_ranks; ++r) {
    double *at = full + ((size_t)r * rows_per_rank + row0) * row_elems;
    ncclResult_t res = g_nccl.Broadcast(at, at, rows * row_elems, ncclDouble, r, c->comm, c->stream);
    if (res != ncclSuccess) { g_nccl.GroupEnd(); *err = std::string("ncclBroadcast: ") + g_nccl.GetErrorString(res); return 1; }
  }
  WB_NCCL(g_nccl.GroupEnd(), err);
  return 0;
}

// The same for several arrays of one utterance slice in ONE NCCL group (one fused launch instead of one per array).
int comm_gather_rows_multi(Comm *c, int n_arrays, double *const *full, const size_t *row_elems, size_t rows_per_rank,
                           size_t row0, size_t rows, cudaEvent_t after, std::string *err) {
  if (rows == 0) return 0;
  if (after && cudaStreamWaitEvent(c->stream, after, 0) != cudaSuccess) { *err = "cudaStreamWaitEvent failed"; return 1; }
  WB_NCCL(g_nccl.GroupStart(), err);
  for (int a = 0; a < n_arrays; ++a) {
    if (!full[a] || row_elems[a] == 0) continue;
    for (int r = 0; r < c->n_ranks; ++r) {
      double *at = full[a] + ((size_t)r * rows_per_rank + row0) * row_elems[a];
      ncclResult_t res = g_nccl.Broadcast(at, at, rows * row_elems[a], ncclDouble, r, c->comm, c->stream);
      if (res != ncclSuccess) { g_nccl.GroupEnd(); *err = std::string("ncclBroadcast: ") + g_nccl.GetErrorString(res); return 1; }
    }
  }
  WB_NCCL(g_nccl.GroupEnd(), err);
  return 0;
}

// `s` waits for everything enqueued on the communication stream so far
int comm_join(Comm *c, cudaStream_t s, std::string *err) {
  if (cudaEventRecord(c->done, c->stream) != cudaSuccess || cudaStreamWaitEvent(s, c->done, 0) != cudaSuccess) {
    *err = "joining the communication stream failed";
    return 1;
  }
  return 0;
}

}  // namespace wb

#else   // host emulation: no devices, no NCCL -- a communicator of one rank, every call a no-op
namespace wb {
struct Comm { int n_ranks = 1, rank = 0; };
int comm_unique_id(unsigned char *id128, std::string *) { memset(id128, 0, 128); return 0; }
int comm_create(int n_ranks, int rank, const unsigned char *, Comm **out, std::string *err) {
  if (n_ranks != 1 || rank != 0) { *err = "the host emulation has no multi-GPU path"; return 1; }
  *out = new Comm;
  return 0;
}
void comm_destroy(Comm *c) { delete c; }
int comm_ranks(const Comm *c) { return c->n_ranks; }
int comm_rank(const Comm *c) { return c->rank; }
}  // namespace wb
#endif
