// Shared host/device declarations of the rollout engine (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "brm.h"

namespace brm {

void set_error(const char* fmt, ...);

#define BRM_CUDA_CHECK(expr)                                                          \
  do {                                                                                \
    cudaError_t _err = (expr);                                                        \
    if (_err != cudaSuccess) {                                                        \
      ::brm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_err), __FILE__, \
                       __LINE__);                                                     \
      return BRM_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define BRM_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::brm::set_error(__VA_ARGS__);      \
      return BRM_ERR_INVALID;             \
    }                                     \
  } while (0)

// ---- GridWorld scenario blob: staged into shared memory by one TMA bulk copy ----
struct alignas(16) GwBlob {
  int32_t A, V, R, n_actions;
  int32_t merging_point, x_len, terminal_depth, n_rules;
  int32_t action_map[BRM_MAX_ACTIONS];
  double w_speed, w_acc;
  uint8_t slot_rule[BRM_GW_MAX_AGENTS][BRM_GW_MAX_SLOTS];
  int8_t slot_other[BRM_GW_MAX_AGENTS][BRM_GW_MAX_SLOTS];
  // per rule slot, for the automaton step: the bit of the agent's label word each of the 4 APs
  // reads (a byte each; 31 = always zero), and rule | n_aps << 8 | priority << 16 |
  // init_state << 24 | on_violation << 31.  Label word: bit 0 collision, 1 merged_e,
  // 2 + j merged_x#j, 10 + j in_direct_front_x#j.  `rules[k].trans` is stored with AP k of the
  // letter at bit k (the ABI has AP 0 at the top bit).
  uint32_t slot_bits[BRM_GW_MAX_AGENTS][BRM_GW_MAX_SLOTS];
  uint32_t slot_meta[BRM_GW_MAX_AGENTS][BRM_GW_MAX_SLOTS];
  brm_rule rules[BRM_MAX_RULES];
  uint8_t _tail[16 - (sizeof(brm_rule) * BRM_MAX_RULES) % 16];
};
static_assert(sizeof(GwBlob) % 16 == 0, "TMA bulk copies move multiples of 16 bytes");

struct EventPair {
  cudaEvent_t a, b;
};

}  // namespace brm

struct CwScenario;  // cw_kernels.cu
struct brm_engine;
namespace brm {
int launch_root_stats(brm_engine* e, int32_t n_leaves, int32_t n_agents, int32_t n_actions, int32_t V,
                      const uint8_t* d_first_action, const double* d_returns_sum, int32_t rollouts_per_leaf,
                      double* d_stats, const double* d_stats_in = nullptr);  // stats = stats_in (default: stats) + ...
// comm.cu: in-place sum over the ranks of the engine's communicator (no-op without one)
int comm_allreduce_device(brm_engine* e, double* d_buf, size_t count);
void comm_release(brm_engine* e);
}

struct brm_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 0, clock_khz = 0, cc_major = 0, cc_minor = 0;
  uint64_t launches = 0;
  // kernel timing
  bool timing = false;
  std::vector<brm::EventPair> ev_pool;
  size_t ev_used = 0;
  double kernel_ms = 0.0;
  uint64_t timed_launches = 0;
  // gridworld
  bool gw_ready = false;
  brm::GwBlob gw_blob;
  brm::GwBlob* d_gw_blob = nullptr;
  // continuous world (cw_engine.cu)
  void* cw = nullptr;
  // root-parallel communicator (comm.cu)
  void* comm = nullptr;
  uint64_t collectives = 0;
  // scratch
  void* d_ws = nullptr;
  size_t ws_bytes = 0;
  // one-copy staging of host-buffer calls (staging.cuh, Arena): pinned host + device, same layout
  uint8_t* h_arena = nullptr;
  uint8_t* d_arena = nullptr;
  size_t arena_bytes = 0;

  int ensure_ws(size_t bytes);
  // timing helpers around the dominant kernels
  void time_begin();
  void time_end();
};
