// Launch interface between the continuous-world host API (cw_engine.cu) and the kernels
// (cw_kernels.cu).
#pragma once
#include "common.cuh"
#include "cw_core.cuh"

namespace brm {

// The two launch shapes of the continuous-world kernels (threads per CTA x resident CTAs per
// SM).  Fourteen warps of 24 slots are what fits beside the blob for four agents at 616 bytes per
// rollout (the kernel needs 128 registers: not the limit); measured 288 ... 512 threads with
// the pool sizes they leave: throughput follows warps x slots
#ifndef BRM_CW_THREADS_LARGE
#define BRM_CW_THREADS_LARGE 448
#endif
constexpr int kCwThreadsLarge = BRM_CW_THREADS_LARGE;  // x 1
constexpr int kCwThreadsSmall = 128;  // x 3
constexpr int kCwCtasSmall = 3;

enum { MODE_ROLLOUT = 0, MODE_EXECUTE = 1, MODE_RULES = 2, MODE_QUERY = 3, MODE_REPLAY = 4 };

// what the engine knows about the uploaded scenarios
struct CwLaunchCtx {
  const uint8_t* blobs = nullptr;       // [n_scenarios] CwBlob, stride blob_stride
  const int32_t* blob_bytes = nullptr;  // used bytes of each blob (multiple of 16)
  size_t blob_stride = 0;
  int32_t blob_cap = 0;  // bytes reserved for the blob in shared memory
  int32_t A = 0, M = 0, V = 0, R = 0;
  int32_t threads = 0;       // CTA size: kCwThreadsLarge or kCwThreadsSmall
  int32_t P = 0;             // rollout slots of a warp's pool (<= 32)
  int32_t target_items = 0;  // phase-1 items a warp aims at per iteration (multiple of 32)
  size_t warp_bytes = 0;     // shared memory of one warp: pool + finalisation records + item lists
  size_t smem_bytes = 0;     // dynamic shared memory per CTA: blob + the warps' pieces
};

// bytes of the arrays that follow a warp's pool in shared memory (cw_pool_kernel): what a
// finished rollout leaves behind for its finalisation (8 + 6 * 4 bytes) and the three item lists
inline size_t cw_sched_bytes(int P, int A, int M, int R) {  // (does not depend on V)
  size_t b = static_cast<size_t>(P) * (8 + 6 * 4);
  b += static_cast<size_t>(A) * (kAgWords * 8 + 4) + static_cast<size_t>(M) * (1 + (R + 7) / 8) * 4;  // cached leaf
  b = (b + 7) & ~static_cast<size_t>(7);
  b += static_cast<size_t>(P) * (2 * A + M) * 2;
  return (b + 15) & ~static_cast<size_t>(15);
}

struct CwArgs {
  const uint8_t* blobs;
  const int32_t* blob_bytes;
  size_t blob_stride;
  int32_t blob_cap;
  int32_t A, M, V, R, P;
  int32_t warp_bytes, target_items, n_warps_total;
  brm_cw_states in;  // leaves (rollouts) / states (batch modes)
  unsigned long long* work_counter;  // next unclaimed rollout / state (one scenario) or leaf (several); zeroed
  int32_t* leaf_next;                // several scenarios: per leaf, its next unclaimed rollout; zeroed
  int64_t total;                     // work units behind the counter
  int32_t rollouts_per_leaf, max_steps, first_exp;
  int32_t single_scenario;  // 1: every unit lives in one scenario -> slots draw from the flat range
  double gamma;
  uint64_t seed, rollout_base;
  // rollouts
  brm_cw_rollout_out out;
  long long* returns_fx;  // [leaves][M*V] fixed-point sums, zeroed before the launch
  // batch modes
  brm_cw_states bout;
  const uint8_t* actions;  // execute: [M][n]; replay: [B][T][M]
  double* rewards;         // execute / rules: [M][V][n]
  uint8_t* num_actions;    // query: [M][n]
  double* term_reward;     // query: [M][V][n]
  int32_t do_check_terminal;
  int32_t T;
  brm_cw_trace tr;
};

// 64-bit words of workspace a rollout launch needs (fixed-point sums, the work counter, per-leaf counters)
inline size_t cw_rollout_ws_words(size_t leaves, int M, int V) { return leaves * M * V + 1 + (leaves + 1) / 2; }

// zeroed_ws: NULL, or cw_rollout_ws_words() already cleared 64-bit words (the caller then also
// cleared out.viol_sum / out.agent_steps)
int cw_launch_rollout(brm_engine* e, const CwLaunchCtx& ctx, int n_scenarios, const brm_cw_states& leaves,
                      const brm_rollout_params& rp, const brm_cw_rollout_out& out, long long* zeroed_ws = nullptr);
int cw_launch_batch(brm_engine* e, const CwLaunchCtx& ctx, int mode, const brm_cw_states& in,
                    const brm_cw_states* out, const uint8_t* actions, double* rewards, uint8_t* num_actions,
                    double* term_reward, int do_check_terminal, int32_t T, const brm_cw_trace* tr);

}  // namespace brm
