// comm.cu -- the exchange step between SNP shards, owned by the library.
//
// The path shards by SNP (rfit) / whole LD blocks (EigenSNP): one context per GPU, and after a sample-side sketch pass
// the N x l partial sums of the shards are added up (plus l x l Grams of row-sharded orthonormalisations).  The
// reference is a single process (nothing to replace); the communicator is NCCL over NVLink / NVSwitch:
//   gpca_comm_unique_id  -> ncclGetUniqueId   (one rank creates it, the host hands it to the others)
//   gpca_comm_init       -> ncclCommInitRank  (one context = one rank; contexts may live in threads of one process
//                                              or in one process each)
// and every collective is issued by the library on the context's own stream -- no host callback between passes.
// libnccl is bound at run time (dlopen), so the library loads on machines without it and a process that already
// carries an NCCL (e.g. PyTorch's) shares that copy.  The host-provided hook (gpca_set_allreduce) remains for hosts
// with their own transport and for single-GPU tests of the sharded code.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>

#include "driver_util.cuh"

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      const char* why = dlerror();      // (one call: dlerror() clears the message it returns)
      api.why = std::string("libnccl.so.2 not found: ") + (why ? why : "");
      return;
    }
    auto sym = [&](const char* s) { return dlsym(api.handle, s); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.ReduceScatter = (decltype(api.ReduceScatter))sym("ncclReduceScatter");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.Broadcast ||
        !api.GetErrorString) {
      api.why = "libnccl is missing a required symbol";
      api.handle = nullptr;
    }
  });
  return api.handle ? &api : nullptr;
}

int fail(gpca_ctx* c, int code, const std::string& msg) {
  c->set_error(msg);
  return code;
}
int nccl_fail(gpca_ctx* c, NcclApi* api, ncclResult_t r, const char* what) {
  return fail(c, GPCA_ERR_CUDA, std::string(what) + ": " + api->GetErrorString(r));
}
}  // namespace

static_assert(sizeof(ncclUniqueId) == GPCA_COMM_ID_BYTES, "gpca.h: GPCA_COMM_ID_BYTES must be sizeof(ncclUniqueId)");

extern "C" int gpca_comm_unique_id(uint8_t* id_out) {
  NcclApi* api = nccl_api();
  if (!api || !id_out) return GPCA_ERR_INVALID;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return GPCA_ERR_CUDA;
  memcpy(id_out, &id, sizeof(id));
  return GPCA_OK;
}

extern "C" int gpca_comm_init(gpca_ctx* c, const uint8_t* id_bytes, int rank, int world) {
  if (!c) return GPCA_ERR_INVALID;
  NcclApi* api = nccl_api();
  if (!api) return fail(c, GPCA_ERR_INVALID, "NCCL is not available (libnccl.so.2 could not be loaded)");
  if (!id_bytes || world < 1 || rank < 0 || rank >= world) return fail(c, GPCA_ERR_INVALID, "gpca_comm_init: bad rank / world");
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  gpca_comm_destroy(c);
  if (world == 1) {
    c->comm_rank = 0;
    c->comm_world = 1;
    return GPCA_OK;      // a single shard needs no communicator
  }
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm = nullptr;
  const ncclResult_t r = api->CommInitRank(&comm, world, id, rank);
  if (r != ncclSuccess) return nccl_fail(c, api, r, "ncclCommInitRank");
  c->nccl_comm = (void*)comm;
  c->comm_rank = rank;
  c->comm_world = world;
  return GPCA_OK;
}

void gpca_comm_destroy(gpca_ctx* c) {
  if (!c || !c->nccl_comm) return;
  NcclApi* api = nccl_api();
  if (api) {
    cudaStreamSynchronize(c->stream);
    api->CommDestroy((ncclComm_t)c->nccl_comm);
  }
  c->nccl_comm = nullptr;
  c->comm_rank = 0;
  c->comm_world = 1;
}

extern "C" int gpca_comm_finalize(gpca_ctx* c) {
  if (!c) return GPCA_ERR_INVALID;
  gpca_comm_destroy(c);
  return GPCA_OK;
}

extern "C" int gpca_comm_world(const gpca_ctx* c) { return c ? c->comm_world : 0; }

// sum of a device buffer over the shards, on the context's stream (dtype 0 = f32, 1 = f64); no-op for a single shard
int driver_allreduce(gpca_ctx* c, void* buf, uint64_t count, int dtype) {
  if (count == 0) return GPCA_OK;
  if (c->nccl_comm) {
    NcclApi* api = nccl_api();
    const ncclResult_t r = api->AllReduce(buf, buf, (size_t)count, dtype == 0 ? ncclFloat32 : ncclFloat64, ncclSum,
                                          (ncclComm_t)c->nccl_comm, c->stream);
    if (r != ncclSuccess) return nccl_fail(c, api, r, "ncclAllReduce");
    c->collectives++;
    return GPCA_OK;
  }
  if (!c->allreduce) return GPCA_OK;
  if (c->allreduce(buf, count, dtype, (void*)c->stream, c->allreduce_user) != 0)
    return fail(c, GPCA_ERR_CUDA, "allreduce hook failed");
  c->collectives++;
  return GPCA_OK;
}

// every shard gets shard 0's copy of a small device buffer (bytes); with the host hook there is no broadcast, so the
// caller must not depend on one (see gpca_rfit: an explicit seed is required there)
int driver_broadcast0(gpca_ctx* c, void* buf, uint64_t bytes) {
  if (!c->nccl_comm || bytes == 0) return GPCA_OK;
  NcclApi* api = nccl_api();
  const ncclResult_t r = api->Broadcast(buf, buf, (size_t)bytes, ncclUint8, 0, (ncclComm_t)c->nccl_comm, c->stream);
  if (r != ncclSuccess) return nccl_fail(c, api, r, "ncclBroadcast");
  c->collectives++;
  return GPCA_OK;
}
