// Host <-> device staging for BRM_MEM_HOST calls: stream-ordered allocations that are
// released when the call returns.
#pragma once
#include "common.cuh"

namespace brm {

struct Stager {
  brm_engine* e;
  std::vector<void*> ptrs;
  uint64_t h2d_bytes = 0, d2h_bytes = 0;
  explicit Stager(brm_engine* eng) : e(eng) {}
  ~Stager() {
    for (void* p : ptrs) cudaFreeAsync(p, e->stream);
  }
  template <class T>
  int alloc(size_t count, T** dev, bool zero = false) {
    void* p = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    BRM_CUDA_CHECK(cudaMallocAsync(&p, bytes, e->stream));
    ptrs.push_back(p);
    if (zero) BRM_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, e->stream));
    *dev = static_cast<T*>(p);
    return BRM_OK;
  }
  template <class T>
  int in(const T* host, size_t count, T** dev) {
    int rc = alloc(count, dev);
    if (rc != BRM_OK) return rc;
    BRM_CUDA_CHECK(cudaMemcpyAsync(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice, e->stream));
    h2d_bytes += count * sizeof(T);
    return BRM_OK;
  }
  template <class T>
  int out(T* host, const T* dev, size_t count) {
    BRM_CUDA_CHECK(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, e->stream));
    d2h_bytes += count * sizeof(T);
    return BRM_OK;
  }
};

#define BRM_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != BRM_OK) return _rc; \
  } while (0)

}  // namespace brm
