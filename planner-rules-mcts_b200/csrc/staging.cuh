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

// One decision on host buffers with ONE copy each way: the engine keeps a pinned host arena and
// a device arena; a call lays its inputs, zero-initialised outputs and plain outputs out in
// them (same offsets on both sides), gathers the inputs into the pinned arena, uploads them
// with one cudaMemcpyAsync, clears the zeroed region with one memset, and after the kernels
// downloads all outputs with one copy and scatters them to the caller's arrays.
class Arena {
 public:
  explicit Arena(brm_engine* e) : e_(e) {}
  template <class T>
  int in(const T* host, size_t count) { return add(kIn, host, nullptr, count * sizeof(T)); }
  template <class T>
  int out(T* host, size_t count, bool zero) { return add(zero ? kOutZero : kOut, nullptr, host, count * sizeof(T)); }
  int tmp_zero(size_t bytes) { return add(kTmpZero, nullptr, nullptr, bytes); }
  template <class T>
  T* dev(int slot) const { return reinterpret_cast<T*>(e_->d_arena + slots_[slot].off); }

  // lay out, gather, upload, clear
  int commit() {
    size_t off = 0;
    const int order[4] = {kIn, kTmpZero, kOutZero, kOut};
    size_t begin[5];
    for (int k = 0; k < 4; ++k) {
      begin[k] = off;
      for (Slot& s : slots_)
        if (s.kind == order[k]) {
          s.off = off;
          off += (s.bytes + 255) & ~static_cast<size_t>(255);
        }
    }
    begin[4] = off;
    in_bytes_ = begin[1];
    zero_off_ = begin[1];
    zero_bytes_ = begin[3] - begin[1];
    out_off_ = begin[2];
    out_bytes_ = begin[4] - begin[2];
    if (off > e_->arena_bytes) {
      BRM_CUDA_CHECK(cudaStreamSynchronize(e_->stream));
      if (e_->d_arena) BRM_CUDA_CHECK(cudaFree(e_->d_arena));
      if (e_->h_arena) BRM_CUDA_CHECK(cudaFreeHost(e_->h_arena));
      e_->d_arena = e_->h_arena = nullptr;
      e_->arena_bytes = 0;
      const size_t want = off + off / 2 + 4096;
      BRM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&e_->d_arena), want));
      BRM_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&e_->h_arena), want, cudaHostAllocDefault));
      e_->arena_bytes = want;
    }
    for (const Slot& s : slots_)
      if (s.kind == kIn) {
        memcpy(e_->h_arena + s.off, s.src, s.bytes);
        h2d_bytes += s.bytes;
      }
    if (in_bytes_)
      BRM_CUDA_CHECK(cudaMemcpyAsync(e_->d_arena, e_->h_arena, in_bytes_, cudaMemcpyHostToDevice, e_->stream));
    if (zero_bytes_) BRM_CUDA_CHECK(cudaMemsetAsync(e_->d_arena + zero_off_, 0, zero_bytes_, e_->stream));
    return BRM_OK;
  }

  // download, wait, scatter
  int finish() {
    if (out_bytes_)
      BRM_CUDA_CHECK(cudaMemcpyAsync(e_->h_arena + out_off_, e_->d_arena + out_off_, out_bytes_,
                                     cudaMemcpyDeviceToHost, e_->stream));
    BRM_CUDA_CHECK(cudaStreamSynchronize(e_->stream));
    for (const Slot& s : slots_)
      if (s.dst) {
        memcpy(s.dst, e_->h_arena + s.off, s.bytes);
        d2h_bytes += s.bytes;
      }
    return BRM_OK;
  }

  uint64_t h2d_bytes = 0, d2h_bytes = 0;

 private:
  enum { kIn = 0, kTmpZero = 1, kOutZero = 2, kOut = 3 };
  struct Slot {
    int kind;
    const void* src;
    void* dst;
    size_t bytes, off;
  };
  int add(int kind, const void* src, void* dst, size_t bytes) {
    slots_.push_back(Slot{kind, src, dst, bytes ? bytes : 1, 0});
    return static_cast<int>(slots_.size()) - 1;
  }
  brm_engine* e_;
  std::vector<Slot> slots_;
  size_t in_bytes_ = 0, zero_off_ = 0, zero_bytes_ = 0, out_off_ = 0, out_bytes_ = 0;
};

#define BRM_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != BRM_OK) return _rc; \
  } while (0)

}  // namespace brm
