// Lookahead-parallel exchange inside the library: one ncclAllGather of the fixed-size per-rank int32 record on the
// compute stream (capturable in the step's CUDA graph), replacing the reference's 3-4 pickled object collectives per
// step (lade/decoding.py:1023-1024, :1043-1058, :1088-1107: broadcast_object_list / all_gather_object of python lists).
//
// NCCL is not linked: the symbols are resolved at run time from the libnccl.so.2 that is already mapped into the
// process (PyTorch's), so the C-ABI library keeps building and loading on a box without NCCL; the entry points then
// return LADE_EUNSUPPORTED.
#include "state.cuh"

#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <string>

namespace lade {

struct NcclUniqueId { char internal[128]; };
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(void**, int, NcclUniqueId, int);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*CommDestroyFn)(void*);
typedef const char* (*GetErrorStringFn)(int);
constexpr int NCCL_INT32 = 2;     // ncclInt32 / ncclInt

struct NcclApi {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllGatherFn all_gather = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
  bool ok = false;
};

static NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {                       // the copy the process already uses (PyTorch's), if any
      api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if (api.handle) break;
    }
    for (int i = 0; !api.handle && i < 2; ++i) api.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) return;
    api.get_unique_id = (GetUniqueIdFn)dlsym(api.handle, "ncclGetUniqueId");
    api.comm_init_rank = (CommInitRankFn)dlsym(api.handle, "ncclCommInitRank");
    api.all_gather = (AllGatherFn)dlsym(api.handle, "ncclAllGather");
    api.comm_destroy = (CommDestroyFn)dlsym(api.handle, "ncclCommDestroy");
    api.error_string = (GetErrorStringFn)dlsym(api.handle, "ncclGetErrorString");
    api.ok = api.get_unique_id && api.comm_init_rank && api.all_gather && api.comm_destroy;
  });
  return api;
}

static int nccl_fail(int rc, const char* where) {
  NcclApi& a = nccl_api();
  std::string msg = std::string(where) + ": NCCL error " + std::to_string(rc);
  if (a.error_string) msg += std::string(" (") + a.error_string(rc) + ")";
  set_error_string(msg.c_str());
  return LADE_ECUDA;
}

}  // namespace lade

using namespace lade;

extern "C" {

int lade_nccl_available(void) { return nccl_api().ok ? 1 : 0; }

int lade_nccl_unique_id(void* id128) {
  if (!id128) return LADE_EINVAL;
  NcclApi& a = nccl_api();
  if (!a.ok) return LADE_EUNSUPPORTED;
  const int rc = a.get_unique_id(reinterpret_cast<NcclUniqueId*>(id128));
  return rc == 0 ? LADE_OK : nccl_fail(rc, "ncclGetUniqueId");
}

int lade_nccl_comm_create(const void* id128, int32_t world, int32_t rank, void** comm_out) {
  if (!id128 || !comm_out || world < 1 || rank < 0 || rank >= world) return LADE_EINVAL;
  NcclApi& a = nccl_api();
  if (!a.ok) return LADE_EUNSUPPORTED;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  void* comm = nullptr;
  const int rc = a.comm_init_rank(&comm, world, id, rank);        // collective: every rank of the group calls it
  if (rc != 0) return nccl_fail(rc, "ncclCommInitRank");
  *comm_out = comm;
  return LADE_OK;
}

int lade_nccl_comm_destroy(void* comm) {
  if (!comm) return LADE_EINVAL;
  NcclApi& a = nccl_api();
  if (!a.ok) return LADE_EUNSUPPORTED;
  const int rc = a.comm_destroy(comm);
  return rc == 0 ? LADE_OK : nccl_fail(rc, "ncclCommDestroy");
}

int lade_lp_exchange(LadeCtx* ctx, void* stream, void* nccl_comm, const int32_t* record_in, int32_t* records_all_out) {
  if (!ctx || !nccl_comm || !record_in || !records_all_out) return LADE_EINVAL;
  if (ctx->d.D < 2) return LADE_ESTATE;
  NcclApi& a = nccl_api();
  if (!a.ok) return LADE_EUNSUPPORTED;
  const size_t n = (size_t)(3 + ctx->d.GS + ctx->d.WCAP);           // == lade_lp_record_ints
  const int rc = a.all_gather(record_in, records_all_out, n, NCCL_INT32, nccl_comm, (cudaStream_t)stream);
  return rc == 0 ? LADE_OK : nccl_fail(rc, "ncclAllGather");
}

}  // extern "C"
