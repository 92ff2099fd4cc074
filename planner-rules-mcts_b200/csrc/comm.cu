// Root-parallel merge behind the C ABI: one NCCL communicator per engine, one small in-place
// sum-allreduce of the packed root statistics on the engine's stream.  The reference has a
// single tree; what is summed is what ReturnBestAction / GetExpectedRewards read at the root
// (gridworld/grid_world_env.h:33,38).  NCCL is bound at run time (dlopen): a single-GPU host
// needs no NCCL at all, and a process that already carries one (PyTorch's) shares it.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"
#include "staging.cuh"

namespace brm {

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, n_ranks = 1;
};

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return BRM_OK;
  const char* names[] = {getenv("BRM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names)
    if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  BRM_REQUIRE(h != nullptr, "brm_comm: cannot load NCCL (libnccl.so.2; set BRM_NCCL_LIB): %s", dlerror());
  NcclApi api;
  api.handle = h;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  BRM_REQUIRE(api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString,
              "brm_comm: the NCCL library lacks an expected symbol");
  g_nccl = api;
  return BRM_OK;
}

#define BRM_NCCL_CHECK(expr)                                                             \
  do {                                                                                   \
    ncclResult_t _r = (expr);                                                            \
    if (_r != ncclSuccess) {                                                             \
      ::brm::set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r));               \
      return BRM_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

}  // namespace

// sum over the ranks of the engine's communicator, in place, on the engine's stream (a no-op
// without a communicator or with one rank)
int comm_allreduce_device(brm_engine* e, double* d_buf, size_t count) {
  Comm* c = static_cast<Comm*>(e->comm);
  if (!c || c->n_ranks <= 1) return BRM_OK;
  BRM_NCCL_CHECK(g_nccl.AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, c->comm, e->stream));
  e->collectives++;
  return BRM_OK;
}

void comm_release(brm_engine* e) {
  Comm* c = static_cast<Comm*>(e->comm);
  if (!c) return;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
  e->comm = nullptr;
}

}  // namespace brm

using namespace brm;

extern "C" {

int brm_comm_unique_id(uint8_t id[BRM_COMM_ID_BYTES]) {
  BRM_REQUIRE(id != nullptr, "brm_comm_unique_id: id is NULL");
  static_assert(BRM_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
  BRM_TRY(load_nccl());
  ncclUniqueId u;
  BRM_NCCL_CHECK(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, BRM_COMM_ID_BYTES);
  return BRM_OK;
}

int brm_comm_init(brm_engine* e, const uint8_t id[BRM_COMM_ID_BYTES], int32_t rank, int32_t n_ranks) {
  BRM_REQUIRE(e && id, "brm_comm_init: NULL argument");
  BRM_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "brm_comm_init: rank %d of %d", rank, n_ranks);
  BRM_TRY(load_nccl());
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  comm_release(e);
  ncclUniqueId u;
  memcpy(u.internal, id, BRM_COMM_ID_BYTES);
  Comm* c = new Comm();
  c->rank = rank;
  c->n_ranks = n_ranks;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, n_ranks, u, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    delete c;
    return BRM_ERR_CUDA;
  }
  e->comm = c;
  return BRM_OK;
}

int brm_comm_destroy(brm_engine* e) {
  BRM_REQUIRE(e != nullptr, "brm_comm_destroy: engine is NULL");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  comm_release(e);
  return BRM_OK;
}

int brm_comm_size(const brm_engine* e) {
  const Comm* c = e ? static_cast<const Comm*>(e->comm) : nullptr;
  return c ? c->n_ranks : 1;
}

int brm_allreduce_root_stats(brm_engine* e, int32_t mem, double* stats, int64_t count) {
  BRM_REQUIRE(e && stats && count >= 1, "brm_allreduce_root_stats: bad arguments");
  BRM_REQUIRE(mem == BRM_MEM_HOST || mem == BRM_MEM_DEVICE, "brm_allreduce_root_stats: bad mem %d", mem);
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  if (mem == BRM_MEM_DEVICE) return comm_allreduce_device(e, stats, static_cast<size_t>(count));
  if (brm_comm_size(e) <= 1) return BRM_OK;
  Stager st(e);
  double* d = nullptr;
  BRM_TRY(st.in(stats, static_cast<size_t>(count), &d));
  BRM_TRY(comm_allreduce_device(e, d, static_cast<size_t>(count)));
  BRM_TRY(st.out(stats, d, static_cast<size_t>(count)));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

}  // extern "C"
