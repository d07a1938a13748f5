// NCCL helpers of the C-ABI: the one collective of the path is a sum all-reduce of the flat fp32 gradient
// bucket (replaces DataParallel's reduce-to-GPU0 + parameter re-broadcast, main.py:81-82).  NCCL is resolved at run time
// (dlopen of the libnccl.so.2 that the host process already carries, e.g. the one bundled with PyTorch), so the library has no
// link-time dependency on it and still loads on a box without NCCL - the comm entry points then fail loudly.
#include <dlfcn.h>
#include <string.h>
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
struct NcclId { char internal[128]; };             // ncclUniqueId
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(void**, int, NcclId, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
  void* h = nullptr;
  fn_get_id get_id = nullptr; fn_init_rank init_rank = nullptr; fn_allreduce allreduce = nullptr; fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // already mapped by the host process (PyTorch)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (h) {
      api.h = h;
      api.get_id = (fn_get_id)dlsym(h, "ncclGetUniqueId");
      api.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
      api.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
      api.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
      api.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
      if (!api.get_id || !api.init_rank || !api.allreduce || !api.destroy) api.h = nullptr;
    }
  }
  return api.h ? &api : nullptr;
}
static int nccl_fail(const char* what, int rc) {
  NcclApi* n = nccl_api();
  set_error("%s: NCCL error %d (%s)", what, rc, n && n->errstr ? n->errstr(rc) : "?");
  return rc > 0 ? rc : UMPR_ERR_ARG;
}
}  // namespace umpr

using namespace umpr;

extern "C" int umpr_comm_unique_id(void* id128) {
  NcclApi* n = nccl_api();
  if (!n) return fail_arg("comm: libnccl.so.2 is not available in this process");
  if (!id128) return fail_arg("comm_unique_id: NULL buffer");
  NcclId id;
  if (int rc = n->get_id(&id)) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int umpr_comm_init(int rank, int world, const void* id128, void** comm) {
  NcclApi* n = nccl_api();
  if (!n) return fail_arg("comm: libnccl.so.2 is not available in this process");
  if (!id128 || !comm || world < 1 || rank < 0 || rank >= world) return fail_arg("comm_init: rank=%d world=%d", rank, world);
  NcclId id;
  memcpy(&id, id128, sizeof(id));
  if (int rc = n->init_rank(comm, world, id, rank)) return nccl_fail("ncclCommInitRank", rc);
  return 0;
}

// in-place sum over all ranks of the flat fp32 gradient bucket, asynchronous on `stream`
extern "C" int umpr_allreduce(void* comm, float* flat, long n_floats, void* stream) {
  NcclApi* n = nccl_api();
  if (!n) return fail_arg("comm: libnccl.so.2 is not available in this process");
  if (!comm || !flat || n_floats < 0) return fail_arg("allreduce: bad arguments");
  if (n_floats == 0) return 0;
  if (int rc = n->allreduce(flat, flat, (size_t)n_floats, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, (cudaStream_t)stream)) return nccl_fail("ncclAllReduce", rc);
  return 0;
}

extern "C" int umpr_comm_destroy(void* comm) {
  NcclApi* n = nccl_api();
  if (!n) return fail_arg("comm: libnccl.so.2 is not available in this process");
  if (!comm) return 0;
  if (int rc = n->destroy(comm)) return nccl_fail("ncclCommDestroy", rc);
  return 0;
}

// Scratch bytes of the entry points that take a caller-owned workspace (PyTorch owns every buffer).
//   "coattn_fwd_tc": a = B, b = P            "cnet_conv_fwd_tc": a = cap (>= N*KC/8)          "cnet_conv_bwd_dx": a = kernel_count
//   "cnet_conv_bwd_dx_tc": (none)
extern "C" int umpr_workspace_bytes(const char* entry, long a, long b, long long* bytes) {
  if (!entry || !bytes) return fail_arg("workspace_bytes: NULL argument");
  if (!strcmp(entry, "coattn_fwd_tc")) {
    const long pv = b < 2560 ? b : 2560;        // images are sized for min(P, 2560) (valid) positions per sample: 20 tiles at most
    const long T = (pv + 127) / 128;
    *bytes = 2ll * a * T * 65536 + 4ll * a * b * 4 + 16ll * a + 4ll * a * b * 16 + 256;
  } else if (!strcmp(entry, "cnet_conv_fwd_tc")) {
    *bytes = 197632ll + 16ll * a;
  } else if (!strcmp(entry, "cnet_conv_bwd_dx")) {
    *bytes = 4ll * a * 3 * 128;
  } else if (!strcmp(entry, "cnet_conv_bwd_dx_tc")) {
    *bytes = 3ll * 65536;
  } else {
    return fail_arg("workspace_bytes: unknown entry point '%s'", entry);
  }
  return 0;
}
