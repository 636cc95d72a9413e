// comm.cu — multi-GPU exchange of the shard-local top-k lists (docs/SPEC.md §5): one process per
// GPU, ncclAllGather of [nq][k] 64-bit keys over NVLink/NVSwitch, then a G-way merge on device.
// The payload is tiny (k = 100: 800 B per query per rank), so the step is launch-latency bound;
// one collective carries a whole query batch and is enqueued on the caller's stream right behind
// the scan, with the merge kernel right behind it.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that the library has no link-time
// dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "handle.h"

struct OiNcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

struct OiComm {
  ncclComm_t comm = nullptr;
};

static OiNcclApi *nccl_api() {
  static OiNcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
      return;
    }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) {
      api.err = "libnccl.so.2 lacks a required symbol";
      api.lib = nullptr;
    }
  });
  return &api;
}

static_assert(sizeof(ncclUniqueId) == OI_UNIQUE_ID_BYTES, "ncclUniqueId size");

extern "C" oi_status oi_comm_unique_id(uint8_t out[OI_UNIQUE_ID_BYTES]) {
  if (!out) return OI_ERR_INVALID_ARG;
  OiNcclApi *api = nccl_api();
  if (!api->lib) return OI_ERR_COMM;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return OI_ERR_COMM;
  memcpy(out, &id, OI_UNIQUE_ID_BYTES);
  return OI_OK;
}

void oi_comm_destroy(oi_index *h) {
  if (!h->comm) return;
  OiNcclApi *api = nccl_api();
  if (api->lib && h->comm->comm) api->CommDestroy(h->comm->comm);
  delete h->comm;
  h->comm = nullptr;
  cudaFree(h->d_keys_local);
  h->d_keys_local = nullptr;
  cudaFree(h->d_gather);
  h->d_gather = nullptr;
  h->world = 1;
  h->rank = 0;
}

extern "C" oi_status oi_index_comm_init(oi_index *h, int32_t rank, int32_t world_size,
                                        const uint8_t unique_id[OI_UNIQUE_ID_BYTES]) {
  if (!h) return OI_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (world_size < 1 || rank < 0 || rank >= world_size) return h->fail(OI_ERR_INVALID_ARG, "rank %d / world %d", rank, world_size);
  cudaError_t ce = cudaSetDevice(h->desc.device);
  if (ce != cudaSuccess) return h->fail(OI_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(ce));
  oi_comm_destroy(h);
  if (world_size == 1) return OI_OK;
  if (!unique_id) return h->fail(OI_ERR_INVALID_ARG, "unique_id is NULL");
  OiNcclApi *api = nccl_api();
  if (!api->lib) return h->fail(OI_ERR_COMM, "NCCL not loadable: %s", api->err.c_str());
  ncclUniqueId id;
  memcpy(&id, unique_id, OI_UNIQUE_ID_BYTES);
  h->comm = new OiComm();
  ncclResult_t r = api->CommInitRank(&h->comm->comm, world_size, id, rank);
  if (r != ncclSuccess) {
    delete h->comm;
    h->comm = nullptr;
    return h->fail(OI_ERR_COMM, "ncclCommInitRank: %s", api->GetErrorString(r));
  }
  const size_t B = h->desc.max_batch, K = h->desc.max_k;
  // two lists per query (cosine | BM25): the hybrid call exchanges both in one collective
  if ((ce = cudaMalloc(&h->d_keys_local, 2 * B * K * sizeof(u64))) != cudaSuccess ||
      (ce = cudaMalloc(&h->d_gather, 2 * (size_t)world_size * B * K * sizeof(u64))) != cudaSuccess) {
    oi_comm_destroy(h);
    return h->fail(OI_ERR_OUT_OF_MEMORY, "cudaMalloc(gather): %s", cudaGetErrorString(ce));
  }
  h->rank = rank;
  h->world = world_size;
  return OI_OK;
}

oi_status oi_comm_gather_merge(oi_index *h, const u64 *d_local, uint32_t nq, uint32_t k, u64 *d_out, cudaStream_t st) {
  if (nq == 0) return OI_OK;
  OiNcclApi *api = nccl_api();
  if (!h->comm || !api->lib) return h->fail(OI_ERR_STATE, "oi_index_comm_init was not called");
  ncclResult_t r = h->comm_skip ? ncclSuccess : api->AllGather(d_local, h->d_gather, (size_t)nq * k, ncclUint64, h->comm->comm, st);
  if (r != ncclSuccess) return h->fail(OI_ERR_COMM, "ncclAllGather: %s", api->GetErrorString(r));
  cudaError_t e = oi_launch_merge_shards(h->d_gather, (uint32_t)h->world, nq, k, d_out, st, &h->launches);
  if (e != cudaSuccess) return h->fail(OI_ERR_CUDA, "merge_shards: %s", cudaGetErrorString(e));
  return OI_OK;
}

// The hybrid call's exchange: d_local2 = [cosine lists nq x k | BM25 lists nq x k] of this shard; ONE all-gather
// carries both modalities (one synchronisation point per batch instead of two: every collective waits for the
// slowest rank), then each modality is merged from its half of every rank's block.
oi_status oi_comm_gather_merge2(oi_index *h, const u64 *d_local2, uint32_t nq, uint32_t k, u64 *d_out_cos, u64 *d_out_bm25, cudaStream_t st) {
  if (nq == 0) return OI_OK;
  OiNcclApi *api = nccl_api();
  if (!h->comm || !api->lib) return h->fail(OI_ERR_STATE, "oi_index_comm_init was not called");
  const size_t half = (size_t)nq * k;
  ncclResult_t r = h->comm_skip ? ncclSuccess : api->AllGather(d_local2, h->d_gather, 2 * half, ncclUint64, h->comm->comm, st);
  if (r != ncclSuccess) return h->fail(OI_ERR_COMM, "ncclAllGather: %s", api->GetErrorString(r));
  cudaError_t e = oi_launch_merge_shards(h->d_gather, (uint32_t)h->world, nq, k, d_out_cos, st, &h->launches, 2 * half);
  if (e == cudaSuccess) e = oi_launch_merge_shards(h->d_gather + half, (uint32_t)h->world, nq, k, d_out_bm25, st, &h->launches, 2 * half);
  if (e != cudaSuccess) return h->fail(OI_ERR_CUDA, "merge_shards: %s", cudaGetErrorString(e));
  return OI_OK;
}
