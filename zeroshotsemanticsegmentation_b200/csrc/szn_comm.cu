// Gradient exchange of the image-sharded data-parallel path (the reference has no distributed code).
// The one collective of the path -- all-reduce(sum) of the parameter gradients -- goes through this C ABI on a communicator
// the host creates from a unique id (broadcast by the caller with whatever control plane it has; torch.distributed in
// ddp.py).  A "bucket" is a list of gradient buffers reduced IN PLACE inside one NCCL group, i.e. one fused NCCL launch:
// fc6 (411 MB) | fc7 (67 MB) | conv5 + conv4 | conv3 ... conv1 + heads + biases, issued as the backward pass finishes them.
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already has: the copy PyTorch ships), so libszn.so
// has no link-time dependency on it and single-GPU users never touch it.
#include "szn_internal.h"
#include <dlfcn.h>
#include <string.h>

namespace {

struct UniqueId {
  char internal[128];
};
typedef void* Comm;
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(Comm*, int, UniqueId, int);
typedef int (*CommDestroyFn)(Comm);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, Comm, cudaStream_t);
typedef int (*GroupFn)(void);
typedef const char* (*ErrStrFn)(int);

struct Nccl {
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllReduceFn all_reduce = nullptr;
  GroupFn group_start = nullptr, group_end = nullptr;
  ErrStrFn err = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n;
  static bool tried = false;
  if (tried) return n;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return n;
  n.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
  n.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
  n.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
  n.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
  n.group_start = (GroupFn)dlsym(h, "ncclGroupStart");
  n.group_end = (GroupFn)dlsym(h, "ncclGroupEnd");
  n.err = (ErrStrFn)dlsym(h, "ncclGetErrorString");
  n.ok = n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.all_reduce && n.group_start && n.group_end;
  return n;
}

int fail(const char* what, int rc) {
  char buf[256];
  snprintf(buf, sizeof buf, "%s: %s", what, (nccl().err && rc) ? nccl().err(rc) : "NCCL not available (libnccl.so.2 not found)");
  return szn::set_error(SZN_ERR_CUDA, buf);
}

// ncclDataType_t values (nccl.h): float32 = 7, float64 = 8, bfloat16 = 9; ncclSum = 0
int nccl_dtype(int szn_dtype) { return szn_dtype == SZN_BF16 ? 9 : szn_dtype == 3 ? 8 : 7; }

}  // namespace

extern "C" int szn_comm_available(void) { return nccl().ok ? 1 : 0; }

extern "C" int szn_comm_unique_id(char* id128) {
  if (!nccl().ok) return fail("szn_comm_unique_id", 0);
  UniqueId id;
  if (int rc = nccl().get_unique_id(&id)) return fail("ncclGetUniqueId", rc);
  memcpy(id128, id.internal, 128);
  return 0;
}

extern "C" int szn_comm_init(const char* id128, int rank, int world, void** comm_out) {
  if (!nccl().ok) return fail("szn_comm_init", 0);
  UniqueId id;
  memcpy(id.internal, id128, 128);
  Comm c = nullptr;
  if (int rc = nccl().comm_init_rank(&c, world, id, rank)) return fail("ncclCommInitRank", rc);
  *comm_out = c;
  return 0;
}

extern "C" int szn_comm_destroy(void* comm) {
  if (!nccl().ok || !comm) return 0;
  if (int rc = nccl().comm_destroy((Comm)comm)) return fail("ncclCommDestroy", rc);
  return 0;
}

extern "C" int szn_allreduce_bucket(void* comm, void* const* bufs, const long long* counts, int n, int dtype, void* stream) {
  if (!nccl().ok) return fail("szn_allreduce_bucket", 0);
  if (!comm || n < 0) return szn::set_error(SZN_ERR_ARG, "szn_allreduce_bucket: comm / n");
  const int dt = nccl_dtype(dtype);
  if (int rc = nccl().group_start()) return fail("ncclGroupStart", rc);
  int first_rc = 0;
  for (int i = 0; i < n; ++i) {
    const int rc = nccl().all_reduce(bufs[i], bufs[i], (size_t)counts[i], dt, /*ncclSum*/ 0, (Comm)comm, (cudaStream_t)stream);
    if (rc && !first_rc) first_rc = rc;
  }
  const int rc_end = nccl().group_end();
  if (first_rc) return fail("ncclAllReduce", first_rc);
  if (rc_end) return fail("ncclGroupEnd", rc_end);
  szn::count_launch();
  return 0;
}
