// Multi-GPU exchange steps of the hot path: an NCCL communicator owned by the engine handle (pp_comm_*).
//
// One process per GPU; the host bootstraps the 128-byte ncclUniqueId however it likes (the Python host broadcasts it
// through torch.distributed's store) and hands it to pp_comm_init.  NCCL is resolved at run time with dlopen/dlsym --
// first the copy already loaded in the process (PyTorch bundles one), then the system library -- so the shared
// library has no link-time NCCL dependency and never mixes two NCCL builds in one process.
//
// The only collective the path needs is an all-gather of row blocks of uneven size inside a contiguous rank range
// (RAFT pair shards, flow-completion feature shards, window predictions).  It is issued as one NCCL group of
// send/recv pairs over NVLink/NVSwitch (every member sends its block to every other member): in place, no staging
// copy, no padding to equal counts, enqueued on the caller's stream.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "engine.cuh"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the host process already uses
  if (lib == nullptr) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (lib == nullptr) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (lib == nullptr) return nullptr;
  bool ok = true;
  auto sym = [&](const char* name) {
    void* p = dlsym(lib, name);
    if (p == nullptr) ok = false;
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
  if (!ok) return nullptr;
  api.lib = lib;
  return &api;
}

#define PP_NCCL_CHECK(api, expr)                                                                    \
  do {                                                                                              \
    ncclResult_t _r = (expr);                                                                       \
    if (_r != ncclSuccess) {                                                                        \
      pp_set_error("%s:%d NCCL error %s: %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r)); \
      return PP_ERR_CUDA;                                                                           \
    }                                                                                               \
  } while (0)

}  // namespace

int pp_comm_unique_id_impl(void* out128) {
  NcclApi* api = nccl_api();
  PP_REQUIRE(api != nullptr, "pp_comm: libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  PP_NCCL_CHECK(api, api->GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, sizeof(id));
  return PP_OK;
}

int pp_comm_init_impl(PPEngine& e, const void* unique_id, int rank, int world) {
  NcclApi* api = nccl_api();
  PP_REQUIRE(api != nullptr, "pp_comm: libnccl.so.2 could not be loaded");
  PP_REQUIRE(world >= 1 && rank >= 0 && rank < world, "pp_comm_init: rank %d of %d", rank, world);
  PP_REQUIRE(e.comm == nullptr, "pp_comm_init: the engine already has a communicator");
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  PP_NCCL_CHECK(api, api->CommInitRank(&comm, world, id, rank));
  e.comm = comm;
  e.rank = rank;
  e.world = world;
  return PP_OK;
}

int pp_comm_destroy_impl(PPEngine& e) {
  if (e.comm != nullptr) {
    NcclApi* api = nccl_api();
    if (api != nullptr) api->CommDestroy(static_cast<ncclComm_t>(e.comm));
    e.comm = nullptr;
  }
  e.rank = 0;
  e.world = 1;
  return PP_OK;
}

// Member m (global rank first_rank + m) owns rows [row_offset[m], row_offset[m] + rows[m]) of `buf` (row_bytes each) and
// has filled them; afterwards every member holds every block.  Blocks may sit in any order and members may own nothing
// (pure receivers).  Ranks outside [first_rank, first_rank + n_members) return immediately.
int pp_comm_all_gather_blocks_impl(PPEngine& e, void* buf, const long long* row_offset, const long long* rows,
                                   size_t row_bytes, int first_rank, int n_members, cudaStream_t st) {
  if (n_members <= 1) return PP_OK;
  PP_REQUIRE(e.comm != nullptr, "pp_comm_all_gather_rows: pp_comm_init was not called");
  PP_REQUIRE(first_rank >= 0 && first_rank + n_members <= e.world, "pp_comm_all_gather_rows: ranks [%d, %d) of %d",
             first_rank, first_rank + n_members, e.world);
  const int me = e.rank - first_rank;
  if (me < 0 || me >= n_members) return PP_OK;
  NcclApi* api = nccl_api();
  ncclComm_t comm = static_cast<ncclComm_t>(e.comm);
  for (int m = 0; m < n_members; ++m)
    PP_REQUIRE(rows[m] >= 0 && row_offset[m] >= 0, "pp_comm_all_gather_rows: negative block");
  uint8_t* base = static_cast<uint8_t*>(buf);
  PP_NCCL_CHECK(api, api->GroupStart());
  for (int m = 0; m < n_members; ++m) {
    if (m == me) continue;
    if (rows[me] > 0)
      PP_NCCL_CHECK(api, api->Send(base + (size_t)row_offset[me] * row_bytes, (size_t)rows[me] * row_bytes, ncclUint8,
                                   first_rank + m, comm, st));
    if (rows[m] > 0)
      PP_NCCL_CHECK(api, api->Recv(base + (size_t)row_offset[m] * row_bytes, (size_t)rows[m] * row_bytes, ncclUint8,
                                   first_rank + m, comm, st));
  }
  PP_NCCL_CHECK(api, api->GroupEnd());
  e.launches++;
  return PP_OK;
}
