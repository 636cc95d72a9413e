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

// Peer-to-peer exchange (SURVEY §5's alternative to the collective): every rank owns one exchange buffer, maps the
// buffers of all peers (CUDA IPC over NVLink) and, instead of calling ncclAllGather, PUSHES its local lists into slot
// `rank` of every peer's buffer with plain stores from one kernel, then raises a per-(peer, rank) flag with the
// batch's epoch; a rank merges as soon as the flags of all ranks carry the epoch.  Two slots (epoch parity): a peer can
// only be one batch ahead, because it needs this rank's flag of the batch in between.
struct OiP2p {
  u64 *sym = nullptr;              // this rank's buffer: [2 slots][world][slot_keys] keys, then [2][world] epoch flags, then 1 error word
  u64 *peer_base[64] = {};         // mapped base of every rank's buffer (peer_base[rank] == sym)
  u64 **d_peer_tbl = nullptr;      // the same table in device memory
  size_t slot_keys = 0;            // keys one rank may write per batch (2 * max_batch * max_k)
  uint64_t epoch = 0;
  bool attached = false;
};

struct OiComm {
  ncclComm_t comm = nullptr;
  OiP2p p2p;
};

// pushes n keys of d_local to every rank and waits for everybody's; *gathered = [world][slot_keys] keys (device)
static oi_status p2p_exchange(oi_index *h, const u64 *d_local, uint32_t n, const u64 **gathered, cudaStream_t st);

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

static void p2p_release(oi_index *h) {
  OiP2p &p = h->comm->p2p;
  if (p.attached)
    for (int r = 0; r < h->world; ++r)
      if (r != h->rank && p.peer_base[r]) cudaIpcCloseMemHandle(p.peer_base[r]);
  cudaFree(p.d_peer_tbl);
  cudaFree(p.sym);
  p = OiP2p();
  h->comm_exchange = 0;
}

void oi_comm_destroy(oi_index *h) {
  if (!h->comm) return;
  p2p_release(h);
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
  if (h->comm_exchange == 1 && h->comm->p2p.attached && !h->comm_skip) {
    const u64 *g = nullptr;
    oi_status s = p2p_exchange(h, d_local, nq * k, &g, st);
    if (s) return s;
    cudaError_t e2 = oi_launch_merge_shards(g, (uint32_t)h->world, nq, k, d_out, st, &h->launches, h->comm->p2p.slot_keys);
    if (e2 != cudaSuccess) return h->fail(OI_ERR_CUDA, "merge_shards: %s", cudaGetErrorString(e2));
    return OI_OK;
  }
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
  if (h->comm_exchange == 1 && h->comm->p2p.attached && !h->comm_skip) {
    const u64 *g = nullptr;
    oi_status s = p2p_exchange(h, d_local2, (uint32_t)(2 * half), &g, st);
    if (s) return s;
    const size_t stride = h->comm->p2p.slot_keys;
    cudaError_t e2 = oi_launch_merge_shards(g, (uint32_t)h->world, nq, k, d_out_cos, st, &h->launches, stride);
    if (e2 == cudaSuccess) e2 = oi_launch_merge_shards(g + half, (uint32_t)h->world, nq, k, d_out_bm25, st, &h->launches, stride);
    if (e2 != cudaSuccess) return h->fail(OI_ERR_CUDA, "merge_shards: %s", cudaGetErrorString(e2));
    return OI_OK;
  }
  ncclResult_t r = h->comm_skip ? ncclSuccess : api->AllGather(d_local2, h->d_gather, 2 * half, ncclUint64, h->comm->comm, st);
  if (r != ncclSuccess) return h->fail(OI_ERR_COMM, "ncclAllGather: %s", api->GetErrorString(r));
  cudaError_t e = oi_launch_merge_shards(h->d_gather, (uint32_t)h->world, nq, k, d_out_cos, st, &h->launches, 2 * half);
  if (e == cudaSuccess) e = oi_launch_merge_shards(h->d_gather + half, (uint32_t)h->world, nq, k, d_out_bm25, st, &h->launches, 2 * half);
  if (e != cudaSuccess) return h->fail(OI_ERR_CUDA, "merge_shards: %s", cudaGetErrorString(e));
  return OI_OK;
}

// ---- peer-to-peer exchange ------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ size_t p2p_flag_off(size_t slot_keys, int world, int slot, int r) {
  return 2 * (size_t)world * slot_keys + (size_t)slot * world + r;
}

// block b pushes this rank's n keys into slot `rank` of rank b's buffer and raises rank b's flag for this rank
__global__ void __launch_bounds__(256) p2p_push_kernel(const u64 *__restrict__ local, u64 *const *__restrict__ peers, int rank, int world,
                                                       uint32_t n, size_t slot_keys, int slot, u64 epoch) {
  u64 *base = peers[blockIdx.x];
  u64 *dst = base + ((size_t)slot * world + rank) * slot_keys;
  const uint4 *src4 = reinterpret_cast<const uint4 *>(local);
  uint4 *dst4 = reinterpret_cast<uint4 *>(dst);
  for (uint32_t i = threadIdx.x; i < n / 2; i += blockDim.x) dst4[i] = src4[i];
  if ((n & 1u) && threadIdx.x == 0) dst[n - 1] = local[n - 1];
  __threadfence_system();  // the keys are visible to the peer before its flag is
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 *flag = base + p2p_flag_off(slot_keys, world, slot, rank);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
  }
}

// one warp: lane r waits until rank r's keys of this batch have arrived (bounded: a lost peer must not hang the GPU)
__global__ void p2p_wait_kernel(u64 *sym, int world, size_t slot_keys, int slot, u64 epoch, long long max_cycles) {
  const int r = threadIdx.x;
  if (r < world) {
    const u64 *flag = sym + p2p_flag_off(slot_keys, world, slot, r);
    const long long t0 = clock64();
    for (;;) {
      u64 v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
      if (v == epoch) break;
      if (clock64() - t0 > max_cycles) {
        sym[2 * (size_t)world * slot_keys + 2 * (size_t)world] = 1ull;  // error word: the batch's lists are incomplete
        break;
      }
      __nanosleep(200);
    }
  }
}

}  // namespace

extern "C" oi_status oi_index_p2p_export(oi_index *h, uint8_t out[OI_P2P_HANDLE_BYTES]) {
  if (!h) return OI_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  static_assert(sizeof(cudaIpcMemHandle_t) == OI_P2P_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  if (!out) return h->fail(OI_ERR_INVALID_ARG, "out is NULL");
  if (!h->comm || h->world < 2) return h->fail(OI_ERR_STATE, "oi_index_comm_init (world >= 2) must come first");
  if (h->world > 64) return h->fail(OI_ERR_UNSUPPORTED, "the peer-to-peer exchange supports at most 64 ranks");
  cudaError_t e = cudaSetDevice(h->desc.device);
  if (e != cudaSuccess) return h->fail(OI_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  OiP2p &p = h->comm->p2p;
  if (!p.sym) {
    p.slot_keys = 2 * (size_t)h->desc.max_batch * h->desc.max_k;
    const size_t words = 2 * (size_t)h->world * p.slot_keys + 2 * (size_t)h->world + 1;
    if ((e = cudaMalloc(&p.sym, words * sizeof(u64))) != cudaSuccess) return h->fail(OI_ERR_OUT_OF_MEMORY, "cudaMalloc(exchange buffer): %s", cudaGetErrorString(e));
    if ((e = cudaMemset(p.sym, 0, words * sizeof(u64))) != cudaSuccess) return h->fail(OI_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e));
  }
  cudaIpcMemHandle_t hd;
  if ((e = cudaIpcGetMemHandle(&hd, p.sym)) != cudaSuccess) return h->fail(OI_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  memcpy(out, &hd, OI_P2P_HANDLE_BYTES);
  return OI_OK;
}

extern "C" oi_status oi_index_p2p_attach(oi_index *h, const uint8_t *handles) {
  if (!h) return OI_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!handles) return h->fail(OI_ERR_INVALID_ARG, "handles is NULL");
  if (!h->comm || !h->comm->p2p.sym) return h->fail(OI_ERR_STATE, "oi_index_p2p_export must come first");
  OiP2p &p = h->comm->p2p;
  if (p.attached) return h->fail(OI_ERR_STATE, "already attached");
  cudaError_t e = cudaSetDevice(h->desc.device);
  if (e != cudaSuccess) return h->fail(OI_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank) { p.peer_base[r] = p.sym; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handles + (size_t)r * OI_P2P_HANDLE_BYTES, OI_P2P_HANDLE_BYTES);
    void *ptr = nullptr;
    if ((e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) {
      for (int q = 0; q < r; ++q)
        if (q != h->rank && p.peer_base[q]) { cudaIpcCloseMemHandle(p.peer_base[q]); p.peer_base[q] = nullptr; }
      return h->fail(OI_ERR_COMM, "cudaIpcOpenMemHandle(rank %d): %s (peer access over NVLink is required)", r, cudaGetErrorString(e));
    }
    p.peer_base[r] = static_cast<u64 *>(ptr);
  }
  if ((e = cudaMalloc(&p.d_peer_tbl, (size_t)h->world * sizeof(u64 *))) != cudaSuccess ||
      (e = cudaMemcpy(p.d_peer_tbl, p.peer_base, (size_t)h->world * sizeof(u64 *), cudaMemcpyHostToDevice)) != cudaSuccess)
    return h->fail(OI_ERR_CUDA, "peer table: %s", cudaGetErrorString(e));
  p.attached = true;
  h->comm_exchange = 1;  // from now on the sharded calls push their lists instead of calling the collective
  return OI_OK;
}

// pushes n keys of d_local to every rank and waits for everybody's; returns the gathered array [world][n] (device)
static oi_status p2p_exchange(oi_index *h, const u64 *d_local, uint32_t n, const u64 **gathered, cudaStream_t st) {
  OiP2p &p = h->comm->p2p;
  if (n > p.slot_keys) return h->fail(OI_ERR_INVALID_ARG, "internal: exchange of %u keys exceeds the slot (%zu)", n, p.slot_keys);
  const u64 epoch = ++p.epoch;
  const int slot = (int)(epoch & 1ull);
  p2p_push_kernel<<<h->world, 256, 0, st>>>(d_local, p.d_peer_tbl, h->rank, h->world, n, p.slot_keys, slot, epoch);
  // ~2 s at 2 GHz before a missing peer is reported instead of waited for
  p2p_wait_kernel<<<1, 64, 0, st>>>(p.sym, h->world, p.slot_keys, slot, epoch, 4000000000ll);
  h->launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return h->fail(OI_ERR_CUDA, "p2p exchange: %s", cudaGetErrorString(e));
  *gathered = p.sym + (size_t)slot * h->world * p.slot_keys;
  return OI_OK;
}

extern "C" oi_status oi_index_p2p_status(oi_index *h, uint64_t *batches, uint32_t *timed_out) {
  if (!h) return OI_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->comm || !h->comm->p2p.attached) return h->fail(OI_ERR_STATE, "the peer-to-peer exchange is not attached");
  OiP2p &p = h->comm->p2p;
  u64 err = 0;
  cudaError_t e = cudaSetDevice(h->desc.device);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaMemcpy(&err, p.sym + 2 * (size_t)h->world * p.slot_keys + 2 * (size_t)h->world, sizeof(u64), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return h->fail(OI_ERR_CUDA, "p2p status: %s", cudaGetErrorString(e));
  if (batches) *batches = p.epoch;
  if (timed_out) *timed_out = err ? 1u : 0u;
  return OI_OK;
}
