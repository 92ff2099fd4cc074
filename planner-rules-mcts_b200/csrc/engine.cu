// Engine life-cycle, error reporting, timing, scratch memory and the root-statistics
// reduction (K5).  Host side of the C ABI declared in include/brm.h.
#include <cstdarg>

#include "common.cuh"
#include "staging.cuh"

namespace brm {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// K5: fold per-leaf rollout sums into root statistics.  One thread per
// (agent, action, component); leaves are visited in index order, so the result is
// deterministic.  stats[a][act][0] += rollouts of leaves whose first action of agent a
// is act; stats[a][act][1+v] += their return sums.
__global__ void root_stats_kernel(int32_t n_leaves, int32_t A, int32_t n_actions, int32_t V,
                                  const uint8_t* __restrict__ first_action,
                                  const double* __restrict__ returns_sum, int32_t rollouts_per_leaf,
                                  const double* stats_in, double* stats) {
  // one warp per output element; lanes stride over the leaves and the partial sums are
  // combined in a fixed butterfly order, so the result is deterministic
  const int elem = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = A * n_actions * (1 + V);
  if (elem >= total) return;
  const int c = elem % (1 + V);
  const int act = (elem / (1 + V)) % n_actions;
  const int a = elem / ((1 + V) * n_actions);
  double x = 0.0;
  for (int l = lane; l < n_leaves; l += 32) {
    if (first_action[l * A + a] != act) continue;
    x += (c == 0) ? static_cast<double>(rollouts_per_leaf)
                  : returns_sum[(static_cast<int64_t>(l) * A + a) * V + (c - 1)];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if (lane == 0) stats[elem] = stats_in[elem] + x;  // accumulated into (stats_in may be stats itself)
}

}  // namespace brm

using brm::set_error;

int brm_engine::ensure_ws(size_t bytes) {
  if (bytes <= ws_bytes) return BRM_OK;
  if (d_ws) {
    BRM_CUDA_CHECK(cudaStreamSynchronize(stream));
    BRM_CUDA_CHECK(cudaFree(d_ws));
    d_ws = nullptr;
    ws_bytes = 0;
  }
  size_t want = bytes + bytes / 4 + 4096;
  BRM_CUDA_CHECK(cudaMalloc(&d_ws, want));
  ws_bytes = want;
  return BRM_OK;
}

void brm_engine::time_begin() {
  if (!timing) return;
  if (ev_used == ev_pool.size()) {
    brm::EventPair p;
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
    ev_pool.push_back(p);
  }
  cudaEventRecord(ev_pool[ev_used].a, stream);
}

void brm_engine::time_end() {
  if (!timing) return;
  cudaEventRecord(ev_pool[ev_used].b, stream);
  ++ev_used;
}

extern "C" {

const char* brm_last_error(void) { return brm::g_error; }
int brm_abi_version(void) { return BRM_ABI_VERSION; }

int brm_abi_struct_sizes(int32_t* out, int32_t n) {
  const int32_t sizes[] = {sizeof(brm_rule), sizeof(brm_rollout_params), sizeof(brm_gw_params), sizeof(brm_gw_states),
                           sizeof(brm_cw_params), sizeof(brm_cw_agent), sizeof(brm_cw_lane), sizeof(brm_cw_scenario),
                           sizeof(brm_cw_states), sizeof(brm_cw_rollout_out), sizeof(brm_cw_trace)};
  const int32_t count = static_cast<int32_t>(sizeof(sizes) / sizeof(sizes[0]));
  for (int32_t i = 0; out && i < n && i < count; ++i) out[i] = sizes[i];
  return count;
}

int brm_create(int device, void* cuda_stream, brm_engine** out) {
  BRM_REQUIRE(out != nullptr, "brm_create: out is NULL");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t err = cudaGetDeviceCount(&n_dev);
  if (err != cudaSuccess || n_dev == 0) {
    set_error("brm_create: no CUDA device available (%s); this engine has no CPU fallback",
              err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
    return BRM_ERR_NO_DEVICE;
  }
  BRM_REQUIRE(device >= 0 && device < n_dev, "brm_create: device %d out of range [0,%d)", device,
              n_dev);
  cudaDeviceProp prop;
  BRM_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("brm_create: device %d is sm_%d%d; kernels are built for sm_100a only", device,
              prop.major, prop.minor);
    return BRM_ERR_NO_DEVICE;
  }
  BRM_CUDA_CHECK(cudaSetDevice(device));
  {
    // BRM_MEM_HOST calls stage through stream-ordered allocations; keep freed blocks in the
    // pool across synchronisations (the default threshold 0 returns them to the OS at every
    // sync, which costs milliseconds per call)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  brm_engine* e = new brm_engine();
  e->device = device;
  e->sm_count = prop.multiProcessorCount;
  e->cc_major = prop.major;
  e->cc_minor = prop.minor;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  e->clock_khz = khz;
  if (cuda_stream) {
    e->stream = static_cast<cudaStream_t>(cuda_stream);
    e->own_stream = false;
  } else {
    cudaError_t se = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
      set_error("cudaStreamCreate failed: %s", cudaGetErrorString(se));
      delete e;
      return BRM_ERR_CUDA;
    }
    e->own_stream = true;
  }
  *out = e;
  return BRM_OK;
}

void brm_cw_release(brm_engine* e);  // cw_engine.cu

void brm_destroy(brm_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  brm_cw_release(e);
  brm::comm_release(e);
  if (e->d_gw_blob) cudaFree(e->d_gw_blob);
  if (e->d_ws) cudaFree(e->d_ws);
  if (e->d_arena) cudaFree(e->d_arena);
  if (e->h_arena) cudaFreeHost(e->h_arena);
  for (auto& p : e->ev_pool) {
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  if (e->own_stream) cudaStreamDestroy(e->stream);
  delete e;
}

int brm_synchronize(brm_engine* e) {
  BRM_REQUIRE(e != nullptr, "brm_synchronize: engine is NULL");
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

uint64_t brm_launch_count(const brm_engine* e) { return e ? e->launches : 0; }

int brm_set_kernel_timing(brm_engine* e, int enabled) {
  BRM_REQUIRE(e != nullptr, "brm_set_kernel_timing: engine is NULL");
  e->timing = enabled != 0;
  return BRM_OK;
}

int brm_kernel_time(brm_engine* e, double* ms_total, uint64_t* launches, int reset) {
  BRM_REQUIRE(e != nullptr, "brm_kernel_time: engine is NULL");
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  for (size_t i = 0; i < e->ev_used; ++i) {
    float ms = 0.f;
    BRM_CUDA_CHECK(cudaEventElapsedTime(&ms, e->ev_pool[i].a, e->ev_pool[i].b));
    e->kernel_ms += ms;
    e->timed_launches++;
  }
  e->ev_used = 0;
  if (ms_total) *ms_total = e->kernel_ms;
  if (launches) *launches = e->timed_launches;
  if (reset) {
    e->kernel_ms = 0.0;
    e->timed_launches = 0;
  }
  return BRM_OK;
}

int brm_device_info(brm_engine* e, int* sm_count, int* sm_clock_khz, int* cc_major, int* cc_minor) {
  BRM_REQUIRE(e != nullptr, "brm_device_info: engine is NULL");
  if (sm_count) *sm_count = e->sm_count;
  if (sm_clock_khz) *sm_clock_khz = e->clock_khz;
  if (cc_major) *cc_major = e->cc_major;
  if (cc_minor) *cc_minor = e->cc_minor;
  return BRM_OK;
}

}  // extern "C"

namespace brm {
// K5 on device pointers (also used by brm_*_rollout_root_stats)
int launch_root_stats(brm_engine* e, int32_t n_leaves, int32_t n_agents, int32_t n_actions, int32_t V,
                      const uint8_t* d_first_action, const double* d_returns_sum, int32_t rollouts_per_leaf,
                      double* d_stats, const double* d_stats_in) {
  const int total = n_agents * n_actions * (1 + V);
  root_stats_kernel<<<(total + 3) / 4, 128, 0, e->stream>>>(n_leaves, n_agents, n_actions, V, d_first_action,
                                                             d_returns_sum, rollouts_per_leaf,
                                                             d_stats_in ? d_stats_in : d_stats, d_stats);
  e->launches++;
  BRM_CUDA_CHECK(cudaGetLastError());
  return BRM_OK;
}
}  // namespace brm

extern "C" {

int brm_root_stats_reduce(brm_engine* e, int32_t mem, int32_t n_leaves, int32_t n_agents,
                          int32_t n_actions, int32_t reward_vec_size,
                          const uint8_t* leaf_first_action, const double* returns_sum,
                          int32_t rollouts_per_leaf, double* stats) {
  BRM_REQUIRE(e != nullptr, "brm_root_stats_reduce: engine is NULL");
  BRM_REQUIRE(n_leaves > 0 && n_agents > 0 && n_actions > 0 && reward_vec_size > 0,
              "brm_root_stats_reduce: bad sizes");
  BRM_CUDA_CHECK(cudaSetDevice(e->device));
  const int total = n_agents * n_actions * (1 + reward_vec_size);
  const size_t fa_bytes = static_cast<size_t>(n_leaves) * n_agents;
  const size_t rs_bytes = static_cast<size_t>(n_leaves) * n_agents * reward_vec_size * sizeof(double);
  const size_t st_bytes = static_cast<size_t>(total) * sizeof(double);
  if (mem == BRM_MEM_DEVICE) {
    brm::root_stats_kernel<<<(total + 3) / 4, 128, 0, e->stream>>>(
        n_leaves, n_agents, n_actions, reward_vec_size, leaf_first_action, returns_sum,
        rollouts_per_leaf, stats, stats);
    e->launches++;
    BRM_CUDA_CHECK(cudaGetLastError());
    return BRM_OK;
  }
  brm::Stager st(e);  // releases the staged buffers on every path out of here
  uint8_t* d_fa = nullptr;
  double *d_rs = nullptr, *d_st = nullptr;
  BRM_TRY(st.in(leaf_first_action, fa_bytes, &d_fa));
  BRM_TRY(st.in(returns_sum, rs_bytes / sizeof(double), &d_rs));
  BRM_TRY(st.in(stats, st_bytes / sizeof(double), &d_st));
  brm::root_stats_kernel<<<(total + 3) / 4, 128, 0, e->stream>>>(
      n_leaves, n_agents, n_actions, reward_vec_size, d_fa, d_rs, rollouts_per_leaf, d_st, d_st);
  e->launches++;
  BRM_CUDA_CHECK(cudaGetLastError());
  BRM_TRY(st.out(stats, d_st, st_bytes / sizeof(double)));
  BRM_CUDA_CHECK(cudaStreamSynchronize(e->stream));
  return BRM_OK;
}

}  // extern "C"
