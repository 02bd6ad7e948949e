// NCCL plumbing of the sharded zone loop (see comm.cuh). Only the handful of NCCL entry points used here are
// resolved, by name, from libnccl.so.2; their prototypes are the public ones of nccl.h (2.x ABI).
#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

namespace fb {

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi g_nccl;

template <typename F>
bool sym(void* h, const char* name, F* out) {
  *out = reinterpret_cast<F>(dlsym(h, name));
  return *out != nullptr;
}

int load_nccl(std::string* err) {
  if (g_nccl.ok) return 0;
  // the copy already in the process (torch's bundled NCCL) wins, so that one process never runs two NCCL builds
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    if (err) *err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
    return -6;
  }
  NcclApi a;
  a.handle = h;
  const bool all = sym(h, "ncclGetUniqueId", &a.GetUniqueId) && sym(h, "ncclCommInitRank", &a.CommInitRank) &&
                   sym(h, "ncclCommDestroy", &a.CommDestroy) && sym(h, "ncclAllReduce", &a.AllReduce) &&
                   sym(h, "ncclSend", &a.Send) && sym(h, "ncclRecv", &a.Recv) && sym(h, "ncclGroupStart", &a.GroupStart) &&
                   sym(h, "ncclGroupEnd", &a.GroupEnd) && sym(h, "ncclGetErrorString", &a.GetErrorString);
  if (!all) {
    if (err) *err = "libnccl.so.2 lacks a required entry point";
    return -6;
  }
  a.ok = true;
  g_nccl = a;
  return 0;
}

int nccl_fail(ncclResult_t r, const char* what, std::string* err) {
  if (err) *err = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
  return -7;
}

}  // namespace

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

static_assert(sizeof(ncclUniqueId) == kCommIdBytes, "fb_comm_unique_id hands out sizeof(ncclUniqueId) bytes");

int comm_unique_id(uint8_t* id128, std::string* err) {
  if (!id128) return -1;
  if (int rc = load_nccl(err)) return rc;
  ncclUniqueId id;
  const ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId", err);
  memcpy(id128, &id, sizeof id);
  return 0;
}

int comm_init(Comm** out, const uint8_t* id128, int rank, int world, std::string* err) {
  if (!out || !id128 || world < 1 || rank < 0 || rank >= world) return -1;
  *out = nullptr;
  if (int rc = load_nccl(err)) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  const ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return nccl_fail(r, "ncclCommInitRank", err);
  }
  *out = c;
  return 0;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
  delete c;
}

int comm_rank(const Comm* c) { return c ? c->rank : 0; }
int comm_world(const Comm* c) { return c ? c->world : 1; }

int comm_allreduce_i64(Comm* c, long long* buf_dev, size_t n, cudaStream_t stream, std::string* err) {
  if (!c || !buf_dev) return -1;
  if (n == 0) return 0;
  const ncclResult_t r = g_nccl.AllReduce(buf_dev, buf_dev, n, ncclInt64, ncclSum, c->comm, stream);
  return r == ncclSuccess ? 0 : nccl_fail(r, "ncclAllReduce", err);
}

int comm_gather_bytes(Comm* c, const void* send_dev, long long send_bytes, void* recv_dev, const long long* offsets,
                      const long long* counts, int root, cudaStream_t stream, std::string* err) {
  if (!c || root < 0 || root >= c->world || send_bytes < 0 || !offsets || !counts) return -1;
  if (counts[c->rank] != send_bytes) {
    if (err) *err = "gather: counts[rank] differs from send_bytes";
    return -1;
  }
  ncclResult_t r = g_nccl.GroupStart();
  if (r != ncclSuccess) return nccl_fail(r, "ncclGroupStart", err);
  if (send_bytes > 0) {
    r = g_nccl.Send(send_dev, static_cast<size_t>(send_bytes), ncclUint8, root, c->comm, stream);
    if (r != ncclSuccess) { g_nccl.GroupEnd(); return nccl_fail(r, "ncclSend", err); }
  }
  if (c->rank == root) {
    for (int p = 0; p < c->world; ++p) {
      if (counts[p] <= 0) continue;
      r = g_nccl.Recv(static_cast<uint8_t*>(recv_dev) + offsets[p], static_cast<size_t>(counts[p]), ncclUint8, p, c->comm, stream);
      if (r != ncclSuccess) { g_nccl.GroupEnd(); return nccl_fail(r, "ncclRecv", err); }
    }
  }
  r = g_nccl.GroupEnd();
  return r == ncclSuccess ? 0 : nccl_fail(r, "ncclGroupEnd", err);
}

}  // namespace fb
