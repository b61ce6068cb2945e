// Run-time binding of the few NCCL entry points the sharded prover uses (comm.h).  Types restated from NCCL's public
// nccl.h (stable since 2.x): ncclUniqueId is 128 opaque bytes passed by value, ncclUint32 = 3, ncclSum = 0.
#include "comm.h"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

namespace zkir {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[ZKIR_COMM_ID_BYTES]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclUint32 = 3, kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string load_error;
};

static NcclApi* api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, []() {
    const char* env = getenv("ZKIR_NCCL_LIB");
    void* h = nullptr;
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy the host process already uses (torch)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { const char* e = dlerror(); a.load_error = std::string("cannot load libnccl.so.2 (set ZKIR_NCCL_LIB): ") + (e ? e : "?"); return; }
    a.handle = h;
#define BIND(field, sym)                                                                \
  *(void**)(&a.field) = dlsym(h, sym);                                                   \
  if (!a.field) { a.load_error = std::string("libnccl lacks ") + sym; a.handle = nullptr; return; }
    BIND(GetUniqueId, "ncclGetUniqueId")
    BIND(CommInitRank, "ncclCommInitRank")
    BIND(CommDestroy, "ncclCommDestroy")
    BIND(AllGather, "ncclAllGather")
    BIND(AllReduce, "ncclAllReduce")
    BIND(Send, "ncclSend")
    BIND(Recv, "ncclRecv")
    BIND(GroupStart, "ncclGroupStart")
    BIND(GroupEnd, "ncclGroupEnd")
    BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
  });
  return &a;
}

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

static int fail(std::string* err, const char* what, ncclResult_t r) {
  if (err) *err = std::string(what) + ": " + (api()->GetErrorString ? api()->GetErrorString(r) : "NCCL error");
  return -3;
}
#define NC(call, what)                           \
  do {                                           \
    ncclResult_t r__ = (call);                   \
    if (r__ != 0) return fail(err, what, r__);   \
  } while (0)

int comm_unique_id(unsigned char id[ZKIR_COMM_ID_BYTES], std::string* err) {
  NcclApi* a = api();
  if (!a->handle) { if (err) *err = a->load_error; return -3; }
  ncclUniqueId u;
  NC(a->GetUniqueId(&u), "ncclGetUniqueId");
  memcpy(id, u.internal, ZKIR_COMM_ID_BYTES);
  return 0;
}

int comm_create(Comm** out, const unsigned char id[ZKIR_COMM_ID_BYTES], int rank, int world, std::string* err) {
  *out = nullptr;
  NcclApi* a = api();
  if (!a->handle) { if (err) *err = a->load_error; return -3; }
  ncclUniqueId u;
  memcpy(u.internal, id, ZKIR_COMM_ID_BYTES);
  Comm* c = new Comm();
  c->rank = rank; c->world = world;
  ncclResult_t r = a->CommInitRank(&c->comm, world, u, rank);
  if (r != 0) { delete c; return fail(err, "ncclCommInitRank", r); }
  *out = c;
  return 0;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->comm && api()->handle) api()->CommDestroy(c->comm);
  delete c;
}

int comm_all_gather_u32(Comm* c, unsigned* buf, size_t words_per_rank, cudaStream_t st, std::string* err) {
  NC(api()->AllGather(buf + (size_t)c->rank * words_per_rank, buf, words_per_rank, kNcclUint32, c->comm, st), "ncclAllGather");
  return 0;
}

int comm_all_reduce_sum_u32(Comm* c, unsigned* buf, size_t words, cudaStream_t st, std::string* err) {
  NC(api()->AllReduce(buf, buf, words, kNcclUint32, kNcclSum, c->comm, st), "ncclAllReduce");
  return 0;
}

int comm_all_gather_group_u32(Comm* c, unsigned* const* bufs, const size_t* words_per_rank, int n, cudaStream_t st, std::string* err) {
  NcclApi* a = api();
  NC(a->GroupStart(), "ncclGroupStart");
  for (int j = 0; j < n; j++) {
    ncclResult_t r = a->AllGather(bufs[j] + (size_t)c->rank * words_per_rank[j], bufs[j], words_per_rank[j], kNcclUint32, c->comm, st);
    if (r != 0) { a->GroupEnd(); return fail(err, "ncclAllGather", r); }
  }
  NC(a->GroupEnd(), "ncclGroupEnd");
  return 0;
}

int comm_exchange_u32(Comm* c, const P2POp* ops, size_t n, cudaStream_t st, std::string* err) {
  NcclApi* a = api();
  if (!n) return 0;
  NC(a->GroupStart(), "ncclGroupStart");
  for (size_t j = 0; j < n; j++) {
    const P2POp& o = ops[j];
    ncclResult_t r = o.is_send ? a->Send(o.ptr, o.words, kNcclUint32, o.peer, c->comm, st) : a->Recv(o.ptr, o.words, kNcclUint32, o.peer, c->comm, st);
    if (r != 0) { a->GroupEnd(); return fail(err, o.is_send ? "ncclSend" : "ncclRecv", r); }
  }
  NC(a->GroupEnd(), "ncclGroupEnd");
  return 0;
}
int comm_rank(const Comm* c) { return c->rank; }
int comm_world(const Comm* c) { return c->world; }

}  // namespace zkir
