// comm.cu -- ghost exchange and scalar all-reduce over NCCL (NVLink 5 / NVSwitch).
//
// Replaces deal.II's Utilities::MPI::Partitioner point-to-point exchange inside
// MatrixFree::cell_loop and the MPI_Allreduce of Vector::operator* (SURVEY 2.3
// "Communication call sites").  NCCL is bound lazily with dlopen so that a
// single-GPU process needs no NCCL at all; inside a PyTorch process the already
// loaded libnccl.so.2 is reused.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace ab200 {

namespace {
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;                          // optional
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t *) = nullptr;  // optional
};
NcclApi g_nccl;

int load_nccl()
{
  if (g_nccl.handle) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) return fail(ADAFLO_B200_ERR_COMM, std::string("cannot load libnccl.so.2: ") + dlerror());
#define AB_SYM(field, name)                                                                  \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.handle, name);                                    \
  if (!g_nccl.field) return fail(ADAFLO_B200_ERR_COMM, std::string("missing NCCL symbol ") + name);
  AB_SYM(GetUniqueId, "ncclGetUniqueId")
  AB_SYM(CommInitRank, "ncclCommInitRank")
  AB_SYM(CommDestroy, "ncclCommDestroy")
  AB_SYM(GroupStart, "ncclGroupStart")
  AB_SYM(GroupEnd, "ncclGroupEnd")
  AB_SYM(Send, "ncclSend")
  AB_SYM(Recv, "ncclRecv")
  AB_SYM(AllReduce, "ncclAllReduce")
  AB_SYM(GetErrorString, "ncclGetErrorString")
#undef AB_SYM
  *(void **)(&g_nccl.CommAbort) = dlsym(g_nccl.handle, "ncclCommAbort");
  *(void **)(&g_nccl.CommGetAsyncError) = dlsym(g_nccl.handle, "ncclCommGetAsyncError");
  return 0;
}
}  // namespace

#define AB_NCCL(expr)                                                                                    \
  do {                                                                                                   \
    ncclResult_t r__ = (expr);                                                                           \
    if (r__ != ncclSuccess)                                                                              \
      return fail(ADAFLO_B200_ERR_COMM, std::string(#expr) + ": " + g_nccl.GetErrorString(r__));         \
  } while (0)

int comm_unique_id(unsigned char id[128])
{
  AB_CHECK(load_nccl());
  ncclUniqueId uid;
  AB_NCCL(g_nccl.GetUniqueId(&uid));
  static_assert(sizeof(uid) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(id, &uid, 128);
  return 0;
}

int comm_init(adaflo_b200_ctx_t ctx, const unsigned char id[128], int rank, int nranks)
{
  AB_CHECK(load_nccl());
  ncclUniqueId uid;
  std::memcpy(&uid, id, 128);
  ncclComm_t comm;
  AB_CUDA(cudaSetDevice(ctx->device));
  AB_NCCL(g_nccl.CommInitRank(&comm, nranks, uid, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (ctx->comm_timeout_s <= 0) ctx->comm_timeout_s = 120.0;  // default with a communicator; ADAFLO_B200_COMM_TIMEOUT overrides (ctx_create)
  return 0;
}

int stream_wait(adaflo_b200_ctx_t ctx, cudaStream_t s)
{
  if (ctx->comm_timeout_s <= 0) {
    AB_CUDA(cudaStreamSynchronize(s));
    return 0;
  }
  if (!ctx->ev_watch) AB_CUDA(cudaEventCreateWithFlags(&ctx->ev_watch, cudaEventDisableTiming));
  AB_CUDA(cudaEventRecord(ctx->ev_watch, s));
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spin = 0;; ++spin) {
    const cudaError_t e = cudaEventQuery(ctx->ev_watch);
    if (e == cudaSuccess) return 0;
    if (e != cudaErrorNotReady) return fail(ADAFLO_B200_ERR_CUDA, std::string("stream_wait: ") + cudaGetErrorString(e));
    if (spin < 2000) continue;  // short waits (a reduction of a Krylov step) never leave the spin
    const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (ctx->nccl_comm && g_nccl.CommGetAsyncError && (spin & 1023u) == 0) {
      ncclResult_t async = ncclSuccess;
      if (g_nccl.CommGetAsyncError((ncclComm_t)ctx->nccl_comm, &async) == ncclSuccess && async != ncclSuccess && async != ncclInProgress) {
        if (g_nccl.CommAbort) g_nccl.CommAbort((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
        return fail(ADAFLO_B200_ERR_COMM, std::string("asynchronous NCCL error: ") + g_nccl.GetErrorString(async));
      }
    }
    if (waited > ctx->comm_timeout_s) {
      // the stream does not drain: most likely a collective whose peers never arrived.  Abort the communicator so that the
      // stuck NCCL kernel ends and the process can report the error instead of spinning until an external timeout.
      if (ctx->nccl_comm && g_nccl.CommAbort) {
        g_nccl.CommAbort((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
      }
      char msg[256];
      std::snprintf(msg, sizeof(msg), "rank %d of %d: the stream did not drain within %.1f s (a collective whose peers never arrived? every rank "
                                      "must issue the same sequence of operator calls); communicator aborted", ctx->rank, ctx->nranks, ctx->comm_timeout_s);
      return fail(ADAFLO_B200_ERR_COMM, msg);
    }
    if (waited > 2e-3) std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
}

int comm_destroy(adaflo_b200_ctx_t ctx)
{
  if (ctx->nccl_comm) {
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  return 0;
}

int comm_allreduce(adaflo_b200_ctx_t ctx, double *dev, int k, int op_max)
{
  if (ctx->nranks <= 1) return 0;
  if (!ctx->nccl_comm) return fail(ADAFLO_B200_ERR_STATE, "communicator not initialised");
  AB_NCCL(g_nccl.AllReduce(dev, dev, (size_t)k, ncclDouble, op_max ? ncclMax : ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return 0;
}

// Several vectors of the SAME space may travel in one group (the three components of a normal field): vector i of the
// call uses slot "number of earlier vectors of its space" of every exchange's staging buffer, which grows on demand.
static int reserve_slots(adaflo_b200_mf_t mf, int n, const int *spaces, std::vector<int> &slot_of)
{
  slot_of.assign(n, 0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (spaces[j] == spaces[i]) ++slot_of[i];
  for (int i = 0; i < n; ++i) {
    Space &sp = mf->spaces[spaces[i]];
    for (Exchange &e : sp.exchange) {
      if (!e.n_send || slot_of[i] < e.send_slots) continue;
      // all earlier work on the stream may still use the old buffer
      AB_CUDA(cudaStreamSynchronize(mf->ctx->stream));
      if (e.send_buf) cudaFree(e.send_buf);
      e.send_slots = slot_of[i] + 1;
      AB_CUDA(cudaMalloc((void **)&e.send_buf, sizeof(double) * (size_t)e.n_send * sp.n_comp * e.send_slots));
    }
  }
  return 0;
}

// owner -> ghost import (LinearAlgebra::distributed::Vector::update_ghost_values) of several vectors in ONE
// NCCL group (one launch for the velocity and the pressure block of a vmult)
int comm_update_ghosts_n(adaflo_b200_mf_t mf, int n, const int *spaces, double *const *vecs)
{
  adaflo_b200_ctx_t ctx = mf->ctx;
  if (ctx->nranks <= 1) return 0;
  if (!ctx->nccl_comm) return fail(ADAFLO_B200_ERR_STATE, "communicator not initialised");
  std::vector<int> slot_of;
  AB_CHECK(reserve_slots(mf, n, spaces, slot_of));
  bool any = false;
  for (int i = 0; i < n; ++i) {
    Space &sp = mf->spaces[spaces[i]];
    for (Exchange &e : sp.exchange) {
      any = true;
      if (e.n_send)
        AB_CHECK(vec_pack(ctx->stream, e.send_buf + (size_t)slot_of[i] * e.n_send * sp.n_comp, vecs[i], e.send_dev, e.n_send, sp.n_comp));
    }
  }
  if (!any) return 0;
  AB_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < n; ++i) {
    Space &sp = mf->spaces[spaces[i]];
    for (Exchange &e : sp.exchange) {
      if (e.n_send)
        AB_NCCL(g_nccl.Send(e.send_buf + (size_t)slot_of[i] * e.n_send * sp.n_comp, (size_t)e.n_send * sp.n_comp, ncclDouble, e.peer,
                            (ncclComm_t)ctx->nccl_comm, ctx->stream));
      if (e.n_recv)
        AB_NCCL(g_nccl.Recv(vecs[i] + (size_t)e.recv_first * sp.n_comp, (size_t)e.n_recv * sp.n_comp, ncclDouble, e.peer,
                            (ncclComm_t)ctx->nccl_comm, ctx->stream));
    }
  }
  AB_NCCL(g_nccl.GroupEnd());
  return 0;
}

// ghost -> owner accumulation (Vector::compress(VectorOperation::add)); ghost entries are zeroed afterwards
int comm_compress_add_n(adaflo_b200_mf_t mf, int n, const int *spaces, double *const *vecs)
{
  adaflo_b200_ctx_t ctx = mf->ctx;
  if (ctx->nranks <= 1) return 0;
  if (!ctx->nccl_comm) return fail(ADAFLO_B200_ERR_STATE, "communicator not initialised");
  bool any = false;
  for (int i = 0; i < n; ++i) any |= !mf->spaces[spaces[i]].exchange.empty();
  if (!any) return 0;
  std::vector<int> slot_of;
  AB_CHECK(reserve_slots(mf, n, spaces, slot_of));
  AB_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < n; ++i) {
    Space &sp = mf->spaces[spaces[i]];
    for (Exchange &e : sp.exchange) {
      if (e.n_recv)
        AB_NCCL(g_nccl.Send(vecs[i] + (size_t)e.recv_first * sp.n_comp, (size_t)e.n_recv * sp.n_comp, ncclDouble, e.peer,
                            (ncclComm_t)ctx->nccl_comm, ctx->stream));
      if (e.n_send)
        AB_NCCL(g_nccl.Recv(e.send_buf + (size_t)slot_of[i] * e.n_send * sp.n_comp, (size_t)e.n_send * sp.n_comp, ncclDouble, e.peer,
                            (ncclComm_t)ctx->nccl_comm, ctx->stream));
    }
  }
  AB_NCCL(g_nccl.GroupEnd());
  for (int i = 0; i < n; ++i) {
    Space &sp = mf->spaces[spaces[i]];
    for (Exchange &e : sp.exchange)
      if (e.n_send)
        AB_CHECK(vec_unpack_add(ctx->stream, vecs[i], e.send_buf + (size_t)slot_of[i] * e.n_send * sp.n_comp, e.send_dev, e.n_send, sp.n_comp));
    if (sp.n_ghost_nodes)
      AB_CUDA(cudaMemsetAsync(vecs[i] + sp.n_owned_entries(), 0, sizeof(double) * (size_t)sp.n_ghost_nodes * sp.n_comp, ctx->stream));
  }
  return 0;
}

int comm_update_ghosts(adaflo_b200_mf_t mf, int space, double *vec) { return comm_update_ghosts_n(mf, 1, &space, &vec); }
int comm_compress_add(adaflo_b200_mf_t mf, int space, double *vec) { return comm_compress_add_n(mf, 1, &space, &vec); }

}  // namespace ab200
